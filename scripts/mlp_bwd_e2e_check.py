"""Is the MLP's backward right under END-TO-END conditions?  Takes the d raw the compositing backward fed it and the X rows
the field wrote (both from the real training step of a fixture), replays the MLP of both passes as a bf16-emulating torch
graph on those very inputs, and compares its weight gradients with the ones the kernels produced.  Separates "MLP
backward" from "everything upstream" in the end-to-end gradient differences (run on the GPU box)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)
from util import load_fixture, make_caster, preset_of, agg_type_of  # noqa: E402
from danbo_b200 import kernels as K, synthetic as syn  # noqa: E402

torch.backends.cuda.matmul.allow_tf32 = False
DEV = "cuda"
ste = lambda t: t + (t.to(torch.bfloat16).float() - t).detach()
MLP = [f"pts_linears.{i}.{w}" for i in range(8) for w in ("weight", "bias")] + [
    "alpha_linear.weight", "alpha_linear.bias", "feature_linear.weight", "feature_linear.bias", "rgb_linear.weight", "rgb_linear.bias"]
for name in sys.argv[1:] or ["train_cfg3_nonoise", "train_fast_nonoise"]:
    fx = load_fixture(name)
    caster, args, P = make_caster(preset_of(fx), train=True, agg_type=agg_type_of(fx))
    caster._debug_bwd = {}
    b = syn.training_batch(int(fx["n_poses"]), int(fx["rays_per_pose"]), seed=int(fx["batch_seed"]))
    rand = {k: fx["rand." + k].to(DEV) for k in ("t_rand", "noise0", "u", "noise1") if ("rand." + k) in fx}
    out = caster.render_rays(b["ray_batch"], N_samples=args.N_samples, kp_batch=b["kp_batch"], skts=b["skts"], cyls=b["cyls"],
                             bones=b["bones"], cams=b["cams"], N_uniques=int(fx["n_poses"]), perturb=1.0,
                             N_importance=args.N_importance, raw_noise_std=float(fx["raw_noise_std"]), _rand=rand)
    tgt, bgs = b["target_s"].to(DEV), b["bgs"].to(DEV)
    l1 = lambda rgb, acc: torch.mean(torch.abs(rgb + (1. - acc)[..., None] * bgs - tgt))
    (l1(out["rgb_map"], out["acc_map"]) + l1(out["rgb0"], out["acc0"])).backward()
    torch.cuda.synchronize()
    got = {n: p.grad.detach().clone() for n, p in caster.network.named_parameters() if p.grad is not None}
    dbg = caster._debug_bwd
    keep = dbg["keep"]
    packed = caster._packed_mlp()
    rbias = K.ray_bias(keep["rays_v"], keep["cam_idx"], keep["codes"], packed)
    W = {k: P[k].float().clone().requires_grad_(True) for k in MLP}
    Wv = P["views_linears.0.weight"].float().clone().requires_grad_(True)
    total = 0.
    for act, fo, d_raw in ((keep["act0"], keep["f0"], dbg["d_raw0"]), (keep["act1"], keep["f1"], dbg["d_raw1"])):
        rows = int(act.count.item())
        x = fo.x_rows[:rows, :195].float()
        ray = fo.row_ray[:rows].long()
        h, a = x, None
        for L in range(8):
            a = torch.relu(h @ ste(W[f"pts_linears.{L}.weight"]).t() + W[f"pts_linears.{L}.bias"])
            h = ste(a)
            if L == 4:
                h = torch.cat([x, h], -1)
        sigma = a @ W["alpha_linear.weight"].t() + W["alpha_linear.bias"]
        feat = ste(h @ ste(W["feature_linear.weight"]).t() + W["feature_linear.bias"])
        g = torch.relu(feat @ ste(Wv[:, :256]).t() + rbias[ray])
        rgb = g @ W["rgb_linear.weight"].t() + W["rgb_linear.bias"]
        raw = torch.cat([rgb, sigma], -1)
        total = total + (raw * d_raw[act.ids[:rows].long()]).sum()
    total.backward()
    print(f"== {name}: MLP weight gradients, kernels vs a bf16-emulating torch graph on the kernels' own X rows and d raw")
    for n in MLP:
        a_, r_ = got[n].reshape(-1).double(), W[n].grad.reshape(-1).double()
        print(f"   {n:28s} |g| {float(r_.norm()):.3e} rel {float((a_ - r_).norm() / r_.norm().clamp_min(1e-30)):.3e}")
    a_, r_ = got["views_linears.0.weight"][:, :256].reshape(-1).double(), Wv.grad[:, :256].reshape(-1).double()
    print(f"   {'views_linears.0.weight[:, :256]':28s} |g| {float(r_.norm()):.3e} rel {float((a_ - r_).norm() / r_.norm().clamp_min(1e-30)):.3e}")
