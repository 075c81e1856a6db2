"""Wall-clock and allocator diagnostics of a full 512x512 danbo_fast render with agg_type=softmax (profiling aid)."""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, ROOT + "/oracle", ROOT + "/tests"): sys.path.insert(0, p)
import torch
import bench
dev = torch.device("cuda", 0)
caster, args, batch = bench.build_scene(0, dev)
rays = batch["ray_batch"].to(dev)
kw = bench.caster_kwargs(args, batch, dev)
for agg in ("sigmoid", "softmax"):
    caster.network.agg_type = agg
    for _ in range(3): caster(rays, **kw)
    torch.cuda.synchronize()
    s0 = torch.cuda.memory_stats()
    t0 = time.perf_counter()
    for _ in range(5): out = caster(rays, **kw)
    t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    s1 = torch.cuda.memory_stats()
    print(f"{agg}: host issue {(t1 - t0) * 200:.2f} ms/call, total {(t2 - t0) * 200:.2f} ms/call, "
          f"cudaMalloc calls {s1['num_device_alloc'] - s0['num_device_alloc']}, cudaFree {s1['num_device_free'] - s0['num_device_free']}, "
          f"retries {s1['num_alloc_retries'] - s0['num_alloc_retries']}, reserved {s1['reserved_bytes.all.current'] / 2**30:.1f} GiB, "
          f"rgb mean {float(out['rgb_map'].mean()):.4f}")
