"""world_size-2 gloo tests (CPU) of the multi-process plumbing: chunk-aligned ray shards, the pixel all-gather and the
flat gradient bucket all-reduce used for data-parallel training."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import danbo_b200
from danbo_b200 import parallel


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    r, w, _ = parallel.init_distributed(backend="gloo")
    assert (r, w) == (rank, world)
    n = 10000                                   # rays; shards must sit on 4096-ray chunk boundaries
    lo, hi = parallel.shard_range(n, rank, world)
    assert lo % parallel.REF_CHUNK == 0 and (hi % parallel.REF_CHUNK == 0 or hi == n)
    local = torch.arange(lo, hi, dtype=torch.float32)[:, None].repeat(1, 5)
    full = parallel.allgather_rows(local)
    assert full.shape == (n, 5) and torch.equal(full[:, 0], torch.arange(n, dtype=torch.float32))
    # flat gradient bucket: one all-reduce, averaged
    torch.manual_seed(0)
    lin = torch.nn.Linear(7, 3)
    bucket = parallel.GradBucket(list(lin.parameters()))
    bucket.zero()
    x = torch.full((4, 7), float(rank + 1))
    lin(x).sum().backward()
    mine = bucket.flat.clone()
    bucket.allreduce(average=True)
    gathered = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(gathered, mine)
    assert torch.allclose(bucket.flat, sum(gathered) / world)
    assert lin.weight.grad.data_ptr() == bucket.flat.data_ptr()       # grads are views of the bucket
    if rank == 0:
        ret["ok"] = True
    dist.barrier()
    dist.destroy_process_group()


def test_two_process_gloo():
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, _free_port(), ret), nprocs=2, join=True)
    assert ret.get("ok")


def test_shard_ranges_cover_everything():
    for n in (1, 4095, 4096, 4097, 261121):
        for w in (1, 2, 4, 8):
            spans = [parallel.shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))


def _pose_worker(rank, world, port, ret):
    """Data-parallel --opt_pose iteration on a stand-in caster: every rank sees other frames; after the step the network
    AND the pose layer must be identical on all ranks (flat-bucket all-reduce + averaged pose gradients)."""
    import types
    import numpy as np
    from danbo_b200 import pose_opt as po, training, synthetic as syn
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    parallel.init_distributed(backend="gloo")

    class Net(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.w = torch.nn.Parameter(torch.full((3,), 0.5))

    class Caster(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.network = Net()

        def forward(self, ray_batch, kp_batch=None, skts=None, bones=None, N_uniques=1, **kw):
            rgb = torch.sigmoid(skts[:, 3, :3, 3] * self.network.w + bones[:, 5, :3].sum(-1, keepdim=True))
            return {"rgb_map": rgb, "acc_map": torch.sigmoid(kp_batch[:, 7, 0])}

    n_frames, R = 6, 4
    rest = syn.rest_pose()
    poses = [syn.make_pose(60 + i, rest, render_cylinder=False) for i in range(n_frames)]
    attrs = {"rest_pose": rest[None], "betas": np.zeros((1, 10), np.float32),
             "kp3d": np.stack([p["kps"] for p in poses]), "bones": np.stack([p["bones"] for p in poses])}
    args = types.SimpleNamespace(opt_rot6d=False, opt_pose_lrate=1e-2, init_poseopt=None, no_poseopt_reload=False,
                                 use_ckpt_anchor=False, opt_pose_cache=False, opt_pose_tol=0.0, opt_pose_coef=1.0,
                                 use_temp_loss=False, ext_scale=0.001, lrate=1e-2, loss_fn="L1", agg_type="sigmoid",
                                 N_samples=8, N_importance=4, perturb=1.0, raw_noise_std=0., use_background=False,
                                 opt_vol_scale=False)
    pose_optimizer, kw = po.create_popt(args, attrs)
    caster = Caster()
    step = training.TrainStep(caster, args, world_size=world, popt_kwargs=kw, pose_optimizer=pose_optimizer)
    mine = torch.tensor([[0, 3], [2, 5]][rank]).repeat_interleave(R)          # two frames per rank, disjoint
    batch = {"ray_batch": torch.zeros(2 * R, 11), "kp_idx": mine, "N_uniques": 2, "cams": torch.zeros(2 * R, 1),
             "cyls": torch.zeros(2 * R, 5), "target_s": torch.full((2 * R, 3), 0.25 + 0.1 * rank)}
    for _ in range(3):
        step(batch)
    layer = kw["popt_layer"]
    flat = torch.cat([layer.pelvis.detach().reshape(-1), layer.bones.detach().reshape(-1), caster.network.w.detach()])
    gathered = [torch.zeros_like(flat) for _ in range(world)]
    dist.all_gather(gathered, flat)
    assert torch.equal(gathered[0], gathered[1]), "replicas diverged"
    moved = (layer.bones.detach() - torch.as_tensor(attrs["bones"])).abs().amax(dim=(1, 2))
    assert all(float(moved[i]) > 0 for i in (0, 2, 3, 5)) and all(float(moved[i]) == 0 for i in (1, 4))
    if rank == 0:
        ret["ok"] = True
    dist.barrier()
    dist.destroy_process_group()


def test_two_process_pose_layer_training_stays_in_sync():
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_pose_worker, args=(2, _free_port(), ret), nprocs=2, join=True)
    assert ret.get("ok")
