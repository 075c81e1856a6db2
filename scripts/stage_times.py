"""Per-stage CUDA-event timing of one danbo_fast 512x512 render (profiling aid).  `--softmax`: agg_type=softmax."""
import sys, os, collections
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, ROOT + "/oracle", ROOT + "/tests"): sys.path.insert(0, p)
import torch
import bench
import danbo_b200
from danbo_b200 import kernels
dev = torch.device("cuda", 0)
caster, args, batch = bench.build_scene(0, dev)
if "--softmax" in sys.argv:
    caster.network.agg_type = "softmax"          # dense pair lists: every bone of an active row is evaluated
rays = batch["ray_batch"].to(dev)
kw = bench.caster_kwargs(args, batch, dev)
for _ in range(3): caster(rays, **kw)
torch.cuda.synchronize()
kernels.PROFILE = {"mlp": [], "launches": 0, "stages": []}
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); st = {}; out = caster(rays, _stages=st, **kw); e1.record(); torch.cuda.synchronize()
prof = kernels.PROFILE; kernels.PROFILE = None
agg = collections.OrderedDict()
for name, a, b in prof["stages"]: agg.setdefault(name, []).append(a.elapsed_time(b))
for a, b, c in prof["mlp"]: agg.setdefault("mlp_forward", []).append(a.elapsed_time(b))
tot = e0.elapsed_time(e1)
for k, v in agg.items(): print(f"{k:22s} n={len(v)} total {sum(v):7.3f} ms  ({100*sum(v)/tot:4.1f}%)  each {['%.3f' % x for x in v]}")
print(f"step total {tot:.3f} ms; sum of stages {sum(sum(v) for v in agg.values()):.3f} ms; rows coarse/fine (last block): {int(st['n_active0'])}, {int(st['n_active1'])}")
