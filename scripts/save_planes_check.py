"""Are the activations the train-mode forward SAVES (danbo_mlp_forward_save) the ones a bf16-emulating restatement
computes from the kernel's own X rows?  End-to-end conditions: the train_cfg3_nonoise / train_fast_nonoise batches, both
passes, CTA-pair kernel, ragged last tile (run on the GPU box)."""
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
bf = lambda t: t.to(torch.bfloat16).float()
for name in sys.argv[1:] or ["train_cfg3_nonoise", "train_fast_nonoise"]:
    fx = load_fixture(name)
    caster, args, P = make_caster(preset_of(fx), train=True, agg_type=agg_type_of(fx))
    b = syn.training_batch(int(fx["n_poses"]), int(fx["rays_per_pose"]), seed=int(fx["batch_seed"]))
    rand = {k: fx["rand." + k].to(DEV) for k in ("t_rand", "noise0", "u", "noise1") if ("rand." + k) in fx}
    out = caster.render_rays(b["ray_batch"], N_samples=args.N_samples, kp_batch=b["kp_batch"], skts=b["skts"], cyls=b["cyls"],
                             bones=b["bones"], cams=b["cams"], N_uniques=int(fx["n_poses"]), perturb=1.0,
                             N_importance=args.N_importance, raw_noise_std=float(fx["raw_noise_std"]), _rand=rand)
    torch.cuda.synchronize()
    keep = out["rgb_map"].grad_fn.keep
    packed = caster._packed_mlp()
    rbias = K.ray_bias(keep["rays_v"], keep["cam_idx"], keep["codes"], packed)
    W = {k: v.float() for k, v in P.items()}
    for tag, act, fo, sv, raw in (("coarse", keep["act0"], keep["f0"], keep["sv0"], keep["raw0"]),
                                  ("fine", keep["act1"], keep["f1"], keep["sv1"], keep["raw1"])):
        rows = int(act.count.item())
        X = fo.x_rows[:rows, :195].float()
        ray = fo.row_ray[:rows].long()
        h, x = X, X
        print(f"== {name} {tag}: {rows} rows ({(rows + 127) // 128} tiles)")
        for L in range(8):
            a = torch.relu(h @ bf(W[f"pts_linears.{L}.weight"]).t() + W[f"pts_linears.{L}.bias"])
            got = sv.act[L][:rows].float()
            want = bf(a)
            d = (got - want).abs()
            ulp = want.abs().clamp_min(1e-3) * 2 ** -7
            bad_rows = (d > 2 * ulp).any(-1)
            mask_diff = ((got > 0) != (want > 0)).float().mean()
            print(f"   a{L}: max |diff| {float(d.max()):.3e}  entries > 2 ulp {float((d > 2 * ulp).float().mean()):.2e}  rows touched "
                  f"{int(bad_rows.sum())}  relu-mask mismatch {float(mask_diff):.2e}")
            h = got                                   # continue from the kernel's own activation: isolates each layer
            if L == 4:
                h = torch.cat([x, h], -1)
        a7 = torch.relu(sv.act[6][:rows].float() @ bf(W["pts_linears.7.weight"]).t() + W["pts_linears.7.bias"])   # unrounded
        sigma = a7 @ W["alpha_linear.weight"].t() + W["alpha_linear.bias"]
        feat = bf(sv.act[7][:rows].float() @ bf(W["feature_linear.weight"]).t() + W["feature_linear.bias"])
        dfeat = (sv.act[8][:rows].float() - feat).abs()
        print(f"   feat: max |diff| {float(dfeat.max()):.3e}")
        g = torch.relu(feat @ bf(W["views_linears.0.weight"][:, :256]).t() + rbias[ray])
        dg = (sv.g[:rows].float() - bf(g)).abs()
        print(f"   g   : max |diff| {float(dg.max()):.3e}")
        rgb = g @ W["rgb_linear.weight"].t() + W["rgb_linear.bias"]
        ids = act.ids[:rows].long()
        got_raw = raw[ids]
        print(f"   raw : max |diff| rgb {float((got_raw[:, :3] - rgb).abs().max()):.3e} sigma {float((got_raw[:, 3:] - sigma).abs().max()):.3e}"
              f" (scale {float(got_raw.abs().max()):.2e})")
