import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, ROOT + "/oracle", ROOT + "/tests"): sys.path.insert(0, p)
import torch, bench
dev = torch.device("cuda", 0)
caster, args, batch = bench.build_scene(0, dev)
rays = batch["ray_batch"].to(dev); kw = bench.caster_kwargs(args, batch, dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for _ in range(3): caster(rays, **kw)
torch.cuda.synchronize()
for mode in ("noflush", "flush", "sync_each"):
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(12)]
    for i in range(12):
        ev[i][0].record(); caster(rays, **kw); ev[i][1].record()
        if mode == "flush": flush.fill_(i)
        if mode == "sync_each": torch.cuda.synchronize()
    torch.cuda.synchronize()
    print(mode, ["%.2f" % a.elapsed_time(b) for a, b in ev], "mem GB", torch.cuda.memory_reserved() / 1e9)
