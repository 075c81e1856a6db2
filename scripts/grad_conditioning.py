"""How well conditioned are the end-to-end parameter gradients?  CPU only (the oracle): perturb the MLP input X, or the MLP
output raw, by a relative 1e-3 and report the relative change of each parameter gradient of one training step
(`python scripts/grad_conditioning.py [fixture]`).  Output under profiles/r2_grad_conditioning.txt: a 1e-3 perturbation of
X moves the head gradients by 0.5-0.7 % but the trunk's by 2 % (layer 7) ... 12 % (layer 0) and the graph / aggregation
nets' by 17-19 % - the random-init field's gradients amplify input differences 20-200 x, which is why bf16 operand
rounding (a 2e-3 relative change of X) shows up as 3-22 % in the end-to-end comparison while every kernel is exact to
rounding on its own inputs."""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle")]
import torch
import danbo_oracle as orc
from util import load_fixture, params_for, align_A, field_mlp_bf16_ste, agg_type_of
from danbo_b200 import synthetic as syn, skeleton as sk
name = sys.argv[1] if len(sys.argv) > 1 else "train_cfg3_nonoise"
fx = load_fixture(name)
b = syn.training_batch(int(fx["n_poses"]), int(fx["rays_per_pose"]), seed=int(fx["batch_seed"]))
rpp = int(fx["rays_per_pose"])
rand = {k: fx["rand." + k] for k in ("t_rand", "noise0", "u", "noise1") if ("rand." + k) in fx}
init_scale = sk.initial_axis_scale(sk.skeleton_profile(syn.rest_pose()), 0.4)
def grads(eps, mode):
    torch.manual_seed(0)
    P = {k: v.clone().requires_grad_(v.dtype.is_floating_point and not k.endswith(".adj")) for k, v in params_for(fx).items()}
    def mlp(x, v, PP):
        if mode == "x":
            x = x * (1 + eps * torch.randn(x.shape, generator=torch.Generator().manual_seed(1)))
        out = field_mlp_bf16_ste(x, v, PP)
        if mode == "raw":
            out = out * (1 + eps * torch.randn(out.shape, generator=torch.Generator().manual_seed(2))).detach()
        return out
    ref = orc.render_rays(b["ray_batch"], b["skts"][::rpp], b["bones"][::rpp], b["cyls"][::rpp], b["cams"], align_A(), P,
                          int(fx["N_samples"]), int(fx["N_importance"]), rays_per_pose=rpp,
                          use_volume_near_far=bool(fx["use_volume_near_far"]), training=True, rand=rand,
                          raw_noise_std=float(fx["raw_noise_std"]), z_samples=fx["st.z_samples.0"], mlp_fn=mlp)
    loss = orc.training_loss(ref, b["target_s"], b["bgs"], P, init_scale)
    loss.backward()
    return {k: v.grad.clone() for k, v in P.items() if v.grad is not None}
g0 = grads(0.0, "none")
for mode in ("x", "raw"):
    g1 = grads(1e-3, mode)
    print(f"--- relative change of the gradients for a 1e-3 relative perturbation of {mode}")
    for k in [f"pts_linears.{i}.weight" for i in (7, 6, 5, 4, 2, 0)] + ["alpha_linear.weight", "rgb_linear.weight", "feature_linear.weight", "graph_net.layers.0.lin.weight", "prob_linears.layers.1.weight"]:
        print(f"   {k:32s} {float((g1[k]-g0[k]).norm()/g0[k].norm()):.3e}")
