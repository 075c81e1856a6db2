"""Multi-GPU plumbing: one process per GPU, torch.distributed (NCCL over NVLink on the B200 box, gloo on CPU).

The path shards without any data-path collective: rays, images and lattice slabs are independent.
Two small exchanges exist around it and are the only collectives used:
  * render: all-gather of the finished pixels (rgb, disp, acc = 20 B / ray);
  * training: one all-reduce(sum) of the flat fp32 gradient bucket (2 455 876 floats), then local Adam.
The reference does this with nn.DataParallel inside one process (core/raycasters.py:116).
"""
import os

import torch
import torch.distributed as dist

REF_CHUNK = 4096        # the reference renders in 4096-ray chunks; shards stay aligned to them


def init_distributed(backend=None):
    """Initialise from the torchrun environment (RANK / WORLD_SIZE / LOCAL_RANK / MASTER_*).  No-op for 1 process."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, world, local


def world():
    return (dist.get_rank(), dist.get_world_size()) if dist.is_initialized() else (0, 1)


def shard_range(n, rank, world_size, align=REF_CHUNK):
    """Contiguous [lo, hi) of `n` units for `rank`, boundaries on multiples of `align` (last shard takes the tail)."""
    n_blocks = (n + align - 1) // align
    per, extra = divmod(n_blocks, world_size)
    b0 = rank * per + min(rank, extra)
    b1 = b0 + per + (1 if rank < extra else 0)
    return min(b0 * align, n), min(b1 * align, n)


def gather_sizes(n_local, device):
    """First-dimension sizes of every rank's shard (one tiny all-gather + host read): call once, outside hot loops."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return [int(n_local)]
    ws = dist.get_world_size()
    sz = torch.tensor([int(n_local)], device=device, dtype=torch.int64)
    all_sz = torch.zeros(ws, device=device, dtype=torch.int64)
    dist.all_gather_into_tensor(all_sz, sz)
    return [int(v) for v in all_sz.tolist()]


def allgather_rows(local, sizes=None):
    """All-gather tensors that differ in their first dimension -> the concatenation, on every rank.  With `sizes` given
    (see gather_sizes) the call is one collective and involves no host synchronisation."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return local
    ws = dist.get_world_size()
    if sizes is None:
        sizes = gather_sizes(local.shape[0], local.device)
    m = max(sizes)
    pad = local.contiguous()
    if local.shape[0] < m:
        pad = torch.cat([local, local.new_zeros((m - local.shape[0],) + tuple(local.shape[1:]))], 0)
    out = torch.empty((ws * m,) + tuple(local.shape[1:]), device=local.device, dtype=local.dtype)
    dist.all_gather_into_tensor(out, pad)
    if all(s == m for s in sizes):
        return out
    return torch.cat([out[r * m:r * m + s] for r, s in enumerate(sizes)], 0)


def pack_pixels(ret):
    """(rgb, disp, acc) -> (n,5) fp32, the 20 B / ray that cross NVLink."""
    return torch.cat([ret["rgb_map"], ret["disp_map"][:, None], ret["acc_map"][:, None]], -1).contiguous()


def render_sharded(caster, ray_batch, tensor_kwargs, other_kwargs, chunk=REF_CHUNK):
    """Render rays [lo,hi) of this rank (chunk-aligned) and all-gather the pixels: every rank returns (N,5)."""
    rank, ws = world()
    n = ray_batch.shape[0]
    lo, hi = shard_range(n, rank, ws, align=chunk)
    sizes = [shard_range(n, r, ws, align=chunk) for r in range(ws)]
    sizes = [b - a for a, b in sizes]
    if hi > lo:
        kw = {k: v[lo:hi] for k, v in tensor_kwargs.items()}
        ret = caster(ray_batch[lo:hi], nanmean_chunk=chunk, **kw, **other_kwargs)
        pix = pack_pixels(ret)
    else:
        dev = next(caster.parameters()).device
        pix = torch.zeros(0, 5, device=dev)
    return allgather_rows(pix, sizes)


class GradBucket:
    """Flat fp32 view of all trainable gradients -> a single all-reduce per iteration."""

    def __init__(self, params):
        self.params = [p for p in params if p.requires_grad]
        n = sum(p.numel() for p in self.params)
        dev = self.params[0].device
        self.flat = torch.zeros(n, device=dev, dtype=torch.float32)
        off = 0
        for p in self.params:
            p.grad = self.flat[off:off + p.numel()].view_as(p)
            off += p.numel()

    def zero(self):
        self.flat.zero_()

    def allreduce(self, average=True):
        if dist.is_initialized() and dist.get_world_size() > 1:
            if average and dist.get_backend() == "nccl":
                dist.all_reduce(self.flat, op=dist.ReduceOp.AVG)        # one collective, no separate scaling launch
                return
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM)
            if average:
                self.flat.div_(dist.get_world_size())
