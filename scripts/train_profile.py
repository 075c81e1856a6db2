import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, ROOT + "/oracle", ROOT + "/tests"): sys.path.insert(0, p)
import torch
import danbo_b200 as db
from danbo_b200 import synthetic as syn, skeleton as sk, training
from torch.profiler import profile, ProfilerActivity
dev = torch.device("cuda", 0)
args = db.make_args("danbo_cfg3", no_reload=True)
data_attrs = {"skel_type": sk.SMPLSkeleton, "near": syn.NEAR, "far": syn.FAR, "n_views": 8, "rest_pose": syn.rest_pose()}
_, kw_test, *_ = db.create_raycaster(args, data_attrs, device=dev)
caster = kw_test["ray_caster"]; caster.network.load_state_dict(syn.synthetic_params(0))
full = syn.training_batch(16, 192, seed=0)
batch = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in full.items()}
step = training.TrainStep(caster, args)
for _ in range(3): step(batch)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(10): step(batch)
t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
print("cpu issue ms/iter", (t1 - t0) * 100, "total ms/iter", (t2 - t0) * 100)
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    step(batch); torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=22, max_name_column_width=60))
# kernels only, by launch count
rows = [(e.key, e.count, e.device_time_total) for e in prof.key_averages() if e.device_type == torch.autograd.DeviceType.CUDA]
rows.sort(key=lambda r: -r[1])
tot_n, tot_t = sum(r[1] for r in rows), sum(r[2] for r in rows)
print(f"--- {tot_n} kernel launches, {tot_t:.0f} us of kernel time; by count:")
for k, n, t in rows[:40]:
    print(f"{n:4d} x {t / max(n, 1):7.1f} us = {t:8.1f} us  {k[:100]}")
