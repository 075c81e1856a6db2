"""Tensor-core MLP backward (mlp_bwd_tcgen05.cu) against the fp32 SIMT chain (backward_mlp.cu) on the same training
batch: per-parameter relative L2 difference of the gradients (run on the GPU box)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)
import danbo_b200 as db  # noqa: E402
from danbo_b200 import kernels as K, synthetic as syn, skeleton as sk, training  # noqa: E402

dev = torch.device("cuda", 0)
n_poses, rpp = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (4, 96)
args = db.make_args("danbo_cfg3", no_reload=True, perturb=0., raw_noise_std=0.)
attrs = {"skel_type": sk.SMPLSkeleton, "near": syn.NEAR, "far": syn.FAR, "n_views": 8, "rest_pose": syn.rest_pose()}
grads = {}
for impl in ("simt", "tc", "tc"):
    K.BACKWARD_IMPL = impl
    _, kw, *_ = db.create_raycaster(args, attrs, device=dev)
    caster = kw["ray_caster"]
    caster.network.load_state_dict(syn.synthetic_params(0))
    step = training.TrainStep(caster, args)
    b = syn.training_batch(n_poses, rpp, seed=5)
    b = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in b.items()}
    loss, _ = step._fwd_bwd(b)
    torch.cuda.synchronize()
    g = {n: p.grad.detach().clone() for n, p in caster.network.named_parameters() if p.grad is not None}
    tag = impl if impl not in grads else impl + "2"
    grads[tag] = g
    print(f"{tag}: loss {float(loss):.6f}")
for a, b_ in (("tc", "simt"), ("tc2", "tc")):
    print(f"--- {a} vs {b_}")
    for n in grads[a]:
        x, y = grads[a][n].reshape(-1).double(), grads[b_][n].reshape(-1).double()
        rel = float((x - y).norm() / y.norm().clamp_min(1e-30))
        if n.startswith(("pts_linears", "alpha", "feature", "views", "rgb", "framecodes")) or rel > 1e-3:
            print(f"{n:40s} |g| {float(y.norm()):.3e} rel {rel:.3e}")
