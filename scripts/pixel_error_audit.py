"""Where do the rendered-pixel differences against the reference come from?  (run on the GPU box)

For every render fixture: (a) coarse pass (identical sample positions): pixel error = bf16 MLP error; (b) fine pass
evaluated at the REFERENCE's importance samples (test hook `_rand[z_fine]`): pixel error with the coarse -> fine coupling
cut; (c) the free-running render: rays above 5e-3 are listed with the shift of their importance samples and whether a
fine sample changed its bone-visibility mask."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)
from util import load_fixture, make_caster, preset_of, agg_type_of, pose_tensors, config_flags_of, lindisp_of  # noqa: E402

DEV = "cuda"
names = sys.argv[1:] or ["render_fast", "render_base", "render_fast_miss", "render_fast_softmax", "render_fast_lindisp",
                         "render_perfcap", "render_surreal"]
for name in names:
    fx = load_fixture(name)
    flags = config_flags_of(fx)
    caster, args, P = make_caster(preset_of(fx), agg_type=agg_type_of(fx), **flags)
    skts, bones, cyl = pose_tensors(fx)
    N = fx["ray_batch"].shape[0]
    ex = lambda t: t.expand(N, *t.shape[1:])
    cams = fx["cams"] if caster.network.opt_framecode else None
    kw = dict(N_samples=args.N_samples, kp_batch=ex(fx["pose_kps"][None]), skts=ex(skts), cyls=ex(cyl), bones=ex(bones),
              cams=cams, N_uniques=1, perturb=False, N_importance=args.N_importance, raw_noise_std=0.,
              lindisp=lindisp_of(fx))
    st_free, st_inj = {}, {}
    free = caster(fx["ray_batch"], _stages=st_free, **kw)
    inj = caster(fx["ray_batch"], _stages=st_inj, _rand={"z_fine": fx["st.z_samples.0"].to(DEV), "z_all": fx["st.z_all.0"].to(DEV),
                                                         "order": fx["st.sorted_idxs.0"].to(DEV)}, **kw)
    torch.cuda.synchronize()
    err = lambda o, k: (o[k].cpu() - fx["out." + k]).abs().reshape(N, -1).max(-1).values
    span = (fx["st.far.0"] - fx["st.near.0"]).reshape(N)
    print(f"== {name}: {N} rays")
    for k in ("rgb0", "acc0"):
        e = err(free, k)
        print(f"   coarse {k}: mean {float(e.mean()):.2e} max {float(e.max()):.2e}")
    raw_inj = st_inj["raw"].cpu()
    scale = float(fx["st.raw.1"].abs().max())
    print(f"   fine at reference samples: merged raw max err {float((raw_inj - fx['st.raw.1']).abs().max()):.3e} of scale {scale:.3e}")
    mask_inj = ((st_inj["mask1"].cpu().long()[..., None] >> torch.arange(24)) & 1) == 0          # invalid
    print(f"   fine at reference samples: visibility mask mismatches {int((mask_inj != (fx['st.invalid.1'] != 0)).sum())}")
    for k in ("rgb_map", "acc_map"):
        e = err(inj, k)
        print(f"   fine at reference samples {k}: mean {float(e.mean()):.2e} max {float(e.max()):.2e}  rays>5e-3: {int((e > 5e-3).sum())}")
    dz = (st_free["z_samples"].cpu() - fx["st.z_samples.0"]).abs().max(-1).values / span.clamp_min(1e-6)
    mask_free = ((st_free["mask1"].cpu().long()[..., None] >> torch.arange(24)) & 1) == 0
    flip = (mask_free != (fx["st.invalid.1"] != 0)).reshape(N, -1).any(-1)
    print(f"   free-running: importance-sample shift / (far-near): median {float(dz.median()):.2e} p99 {float(dz.quantile(0.99)):.2e} max {float(dz.max()):.2e}")
    for k in ("rgb_map", "acc_map"):
        e = err(free, k)
        big = e > 5e-3
        print(f"   free-running {k}: mean {float(e.mean()):.2e} p99 {float(e.quantile(0.99)):.2e} max {float(e.max()):.2e}  rays>5e-3: "
              f"{int(big.sum())}, of which mask flips {int((big & flip).sum())}, min shift among them "
              f"{float(dz[big].min()) if big.any() else 0:.2e}, injected-run err among them max "
              f"{float(err(inj, k)[big].max()) if big.any() else 0:.2e}")
    # how the error scales with the shift: correlation of log err and log dz over rays that hit something
    hit = fx["out.acc_map"] > 1e-3
    e = err(free, "acc_map")
    print(f"   rays with acc>1e-3: {int(hit.sum())}; acc err vs shift (those rays): "
          f"corr {float(torch.corrcoef(torch.stack([e[hit].clamp_min(1e-9).log(), dz[hit].clamp_min(1e-9).log()]))[0, 1]):.2f}")
