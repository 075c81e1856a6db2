"""Runs the BASELINE.json configs this round covers on the B200 and writes profiles/r1_configs.json.
  #1 danbo_base 64x64 render (the reference's CPU-runnable case)      #2 danbo_fast 512x512 render (bench.py headline)
  #3 danbo_base 64+16 training step, 3072 rays                        #5 bullet-time 512x512 views + density lattice
(#4, A-NeRF: scripts/bench_anerf.py.)   Usage: python scripts/bench_configs.py [n_views] [grid_res]"""
import sys, os, json, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, ROOT + "/oracle", ROOT + "/tests"): sys.path.insert(0, p)
import torch
import danbo_b200 as db
from danbo_b200 import synthetic as syn, skeleton as sk, render, parallel
from util import make_caster

rank, world, local = parallel.init_distributed()
dev = torch.device("cuda", local); torch.cuda.set_device(dev)
n_views = int(sys.argv[1]) if len(sys.argv) > 1 else 16
grid_res = int(sys.argv[2]) if len(sys.argv) > 2 else 255
out = {"n_gpus": world}

def timed(fn, reps=3):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps): r = fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps, r

# config 1: danbo_base 64x64, one pose
caster, args, _ = make_caster("danbo_base", device=dev)
pose = syn.make_pose(3)
b = syn.render_batch(pose, 64, 64, full_image=True)
kw = dict(N_samples=args.N_samples, kp_batch=b["kp_batch"], skts=b["skts"], cyls=b["cyls"], bones=b["bones"], cams=b["cams"],
          N_uniques=1, perturb=False, N_importance=args.N_importance, raw_noise_std=0., nanmean_chunk=args.chunk)
dt, _ = timed(lambda: caster(b["ray_batch"], **kw), reps=10)
out["config1_danbo_base_64x64"] = {"rays": 4096, "samples_per_ray": 144, "ms": dt * 1e3, "rays_per_s": 4096 / dt}
# config 5: bullet time + density lattice
caster_f, args_f, _ = make_caster("danbo_fast", device=dev)
poses = [syn.make_pose(100 + k) for k in range(n_views)]
c2ws = syn.bullet_time_cameras(syn.camera(), n_views)
dt, imgs = timed(lambda: render.render_images(caster_f, args_f, list(c2ws), poses, 512, 512, distributed=world > 1), reps=2)
out["config5_bullet_time_512"] = {"views": n_views, "s_total": dt, "images_per_s": n_views / dt, "ms_per_image": dt * 1e3 / n_views,
                                  "mean_pixel": float(imgs.mean())}
t = lambda a: torch.as_tensor(a)[None].to(dev)
dt, grid = timed(lambda: render.density_grid(caster, t(pose["kps"]), t(pose["skts"]), t(pose["bones"]), radius=1.8, res=grid_res,
                                             distributed=world > 1), reps=2)
npts = (grid_res + 1) ** 3
out["config5_density_grid"] = {"points": npts, "s": dt, "points_per_s": npts / dt, "occupied_frac": float((grid > 10).float().mean())}
if rank == 0:
    print(json.dumps(out, indent=1))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", f"configs_{world}gpu.json"), "w"), indent=1)
if world > 1:
    torch.distributed.barrier(); torch.distributed.destroy_process_group()
