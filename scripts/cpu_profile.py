import sys, os, time, cProfile, pstats
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, ROOT + "/oracle", ROOT + "/tests"): sys.path.insert(0, p)
import torch, bench
dev = torch.device("cuda", 0)
caster, args, batch = bench.build_scene(0, dev)
rays = batch["ray_batch"].to(dev); kw = bench.caster_kwargs(args, batch, dev)
for _ in range(3): caster(rays, **kw)
torch.cuda.synchronize()
pr = cProfile.Profile(); pr.enable()
for _ in range(5): caster(rays, **kw)
pr.disable(); torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
