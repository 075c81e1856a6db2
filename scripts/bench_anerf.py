"""BASELINE config #4: anerf_base 1024x1024 render (96+48 samples), one synthetic pose, box-restricted rays.
Prints rays/s, per-kernel times and the tensor roofline fraction of the W=448 MLP kernel (4 536 000 FLOP/sample).
Usage: python scripts/bench_anerf.py [H] [steps]   (under torchrun: one image per rank, pixels all-gathered)"""
import sys, os, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, ROOT + "/oracle", ROOT + "/tests"): sys.path.insert(0, p)
import torch
import danbo_b200 as db
from danbo_b200 import synthetic as syn, skeleton as sk, params, parallel, kernels as K

rank, world, local = parallel.init_distributed()
dev = torch.device("cuda", local); torch.cuda.set_device(dev)
H = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
args = db.make_args("anerf_base", no_reload=True)
attrs = {"skel_type": sk.SMPLSkeleton, "near": syn.NEAR, "far": syn.FAR, "n_views": 8, "rest_pose": syn.rest_pose()}
_, kw_test, *_ = db.create_raycaster(args, attrs, device=dev)
caster = kw_test["ray_caster"]
caster.network.load_state_dict(syn.synth_state_dict(params.anerf_param_shapes(), 0), strict=False)
caster.eval()
pose = syn.make_pose(3 + rank)
b = syn.render_batch(pose, H, H)
b = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in b.items()}
n = b["ray_batch"].shape[0]
kw = dict(N_samples=args.N_samples, kp_batch=b["kp_batch"], skts=b["skts"], cyls=b["cyls"], bones=b["bones"], cams=b["cams"],
          N_uniques=1, perturb=False, N_importance=args.N_importance, raw_noise_std=0., nanmean_chunk=args.chunk, nerf_type="nerf")
for _ in range(2): ret = caster(b["ray_batch"], **kw)
torch.cuda.synchronize()
if world > 1: torch.distributed.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(steps):
    ret = caster(b["ray_batch"], **kw)
    if world > 1: pix = parallel.allgather_rows(parallel.pack_pixels(ret))
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / steps
if world > 1:
    t = torch.tensor([ms], device=dev); torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX); ms = float(t)
# per-kernel times (eager, events around every launch)
K.PROFILE = {"mlp": [], "launches": 0, "stages": []}
caster(b["ray_batch"], **kw)
torch.cuda.synchronize()
prof = {}
for name, a0, a1 in K.PROFILE["stages"]:
    prof.setdefault(name, []).append(a0.elapsed_time(a1))
launches = K.PROFILE["launches"]
K.PROFILE = None
S_t = args.N_samples + args.N_importance
out = {"config4_anerf_base": {"H": H, "n_gpus": world, "rays_per_image": n, "samples_per_ray": S_t, "ms_per_image": ms,
                              "rays_per_s": n * world / (ms / 1e3), "finite": bool(torch.isfinite(ret["rgb_map"]).all()),
                              "mean_acc": float(ret["acc_map"].mean()), "gpu_launches": launches}}
if prof:
    mlp_ms = sum(prof.get("anerf_mlp", []))
    out["config4_anerf_base"]["kernel_ms"] = {k: sum(v) for k, v in prof.items()}
    if mlp_ms > 0:
        out["config4_anerf_base"]["mlp_tflops"] = 4536000.0 * n * S_t / (mlp_ms / 1e3) / 1e12
if rank == 0:
    print(json.dumps(out, indent=1))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", f"anerf_{world}gpu.json"), "w"), indent=1)
if world > 1:
    torch.distributed.barrier(); torch.distributed.destroy_process_group()
