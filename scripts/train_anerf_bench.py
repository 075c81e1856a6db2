"""A-NeRF training iteration at the config's own size (anerf_base: 3 072 rays x (96 + 48) samples, 16 poses x 192 rays):
forward with saved activations (tcgen05), losses, compositing backward, generic-GEMM MLP backward (fp32 SIMT), Adam.
Prints ms / iteration and the loss curve (run on the GPU box)."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)
import danbo_b200 as db  # noqa: E402
from danbo_b200 import synthetic as syn, skeleton as sk, params, training  # noqa: E402

dev = torch.device("cuda", 0)
n_poses, rpp = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (16, 192)
args = db.make_args("anerf_base", no_reload=True)
attrs = {"skel_type": sk.SMPLSkeleton, "near": syn.NEAR, "far": syn.FAR, "n_views": 8, "rest_pose": syn.rest_pose()}
_, kw, *_ = db.create_raycaster(args, attrs, device=dev)
caster = kw["ray_caster"]
caster.network.load_state_dict(syn.synth_state_dict(params.anerf_param_shapes(), 0), strict=False)
step = training.TrainStep(caster, args)
b = syn.training_batch(n_poses, rpp, seed=0)
b = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in b.items()}
torch.manual_seed(0)
losses = [float(step(b)[0]) for _ in range(3)]
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
iters = 5
for _ in range(iters):
    losses.append(float(step(b)[0]))
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / iters
rows = n_poses * rpp * (args.N_samples + args.N_importance)
print(f"anerf_base training: {n_poses * rpp} rays x {args.N_samples + args.N_importance} samples = {rows} rows; "
      f"{ms:.1f} ms / iteration ({1e3 / ms:.2f} it/s); ~{3 * 4.536e6 * rows / (ms * 1e-3) / 1e12:.0f} TFLOP/s effective "
      f"(3 x forward FLOPs); peak memory {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB")
print("losses", ["%.4f" % l for l in losses])
