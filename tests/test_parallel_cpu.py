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
