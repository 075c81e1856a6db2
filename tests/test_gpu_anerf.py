"""AN1 (BASELINE config #4, A-NeRF) on the GPU through the C ABI, against the oracle and the reference's golden."""
import numpy as np
import pytest
import torch

import danbo_oracle as orc
from util import load_fixture, align_A, pose_tensors, decode_tile_image, anerf_mlp_bf16_reference

pytestmark = pytest.mark.gpu
DEV = "cuda"


def K():
    import danbo_b200
    return danbo_b200.kernels


def anerf_params(fx):
    from danbo_b200 import params, synthetic as syn
    return syn.synth_state_dict(params.anerf_param_shapes(), int(fx["weight_seed"]))


def make_anerf_caster(fx):
    import danbo_b200 as db
    from danbo_b200 import synthetic as syn, skeleton as sk
    args = db.make_args("anerf_base", no_reload=True, N_samples=int(fx.get("N_samples", 96)),
                        N_importance=int(fx.get("N_importance", 48)))
    attrs = {"skel_type": sk.SMPLSkeleton, "near": syn.NEAR, "far": syn.FAR, "n_views": 8, "rest_pose": syn.rest_pose()}
    _, kw_test, *_ = db.create_raycaster(args, attrs, device=DEV)
    caster = kw_test["ray_caster"]
    P = anerf_params(fx)
    missing = caster.network.load_state_dict(P, strict=False)
    assert not missing.unexpected_keys and all("pe_fn" in k for k in missing.missing_keys), missing
    caster.eval()
    return caster, args, P


def close(a, b, tol, what):
    scale = max(float(b.abs().max()), 1e-6)
    err = float((a - b).abs().max())
    assert err <= tol * scale + 1e-7, f"{what}: err {err:.3e} scale {scale:.3e}"


def test_anerf_encodings():
    """ray_kernel + embed_kernel: the operand tile images hold the oracle's density / view inputs rounded to bf16."""
    fx = load_fixture("render_anerf")
    caster, args, P = make_anerf_caster(fx)
    skts, bones, cyl = pose_tensors(fx)
    rb = fx["ray_batch"].to(DEV)
    N, S = rb.shape[0], int(fx["N_samples"])
    packed = caster._packed_mlp()
    align = caster._align()
    cams = fx["cams"].reshape(-1).to(DEV).to(torch.int32)
    enc, cb = K().anerf_ray_encode(rb, skts.to(DEV).contiguous(), N, cams, caster._codes_with_mean(), packed)
    z = fx["st.z.0"].to(DEV).contiguous()
    xd, xv = K().anerf_embed(rb, S, z, skts.to(DEV).contiguous(), N, align, enc, float(fx["tau"]))
    got_d = decode_tile_image(xd, N * S, 7, 432).cpu()
    got_v = decode_tile_image(xv, N * S, 11, 648).cpu()
    pts = orc.ray_points(fx["ray_batch"][:, :3], fx["ray_batch"][:, 3:6], fx["st.z.0"])
    sk_all = skts.expand(N, -1, -1, -1)
    want_d, want_v, _ = orc.anerf_inputs(pts, fx["ray_batch"][:, 3:6], fx["cams"], sk_all, align_A(), P, False, float(fx["tau"]))
    # bf16 rounding (2^-9 relative) on top of the fp32 sensitivity of the top octave (see test_oracle_golden)
    assert float((got_d - want_d).abs().max()) <= 4.5e-3, float((got_d - want_d).abs().max())
    assert float((got_v - want_v[:, :648]).abs().max()) <= 4.5e-3
    assert float((got_d - want_d).abs().mean()) <= 6e-4
    # and against the reference's own tensors for the rays it kept
    n_st = fx["st.v.0"].shape[0]
    assert float((got_d[:n_st * S] - fx["st.density_inputs.0"]).abs().max()) <= 4.5e-3
    assert float((got_v[:n_st * S] - fx["st.view_inputs.0"][:, :648]).abs().max()) <= 4.5e-3
    # frame-code part of the view layer, fp32
    Wv = P["views_linears.0.weight"]
    want_cb = want_v[::S, 648:] @ Wv[:, 448 + 648:].t() + P["views_linears.0.bias"]
    close(cb.cpu(), want_cb, 2e-5, "code bias")


def test_anerf_mlp():
    """mlp_kernel (CTA pairs, W = 448) on the embed kernel's rows: against a bf16-emulating restatement (tight) and
    the fp32 oracle / reference raw (bf16 tolerance)."""
    fx = load_fixture("render_anerf")
    caster, args, P = make_anerf_caster(fx)
    skts, bones, cyl = pose_tensors(fx)
    rb = fx["ray_batch"].to(DEV)
    N, S = rb.shape[0], int(fx["N_samples"])
    packed = caster._packed_mlp()
    align = caster._align()
    cams = fx["cams"].reshape(-1).to(DEV).to(torch.int32)
    enc, cb = K().anerf_ray_encode(rb, skts.to(DEV).contiguous(), N, cams, caster._codes_with_mean(), packed)
    z = fx["st.z.0"].to(DEV).contiguous()
    xd, xv = K().anerf_embed(rb, S, z, skts.to(DEV).contiguous(), N, align, enc, float(fx["tau"]))
    raw = torch.full((N * S + N, 4), float("nan"), device=DEV)
    K().anerf_mlp(xd, xv, packed, cb, N * S, S, raw)
    torch.cuda.synchronize()
    got = raw[:N * S].cpu()
    assert torch.isfinite(got).all()
    xd_f = decode_tile_image(xd, N * S, 7, 432).cpu()
    xv_f = decode_tile_image(xv, N * S, 11, 648).cpu()
    emu = anerf_mlp_bf16_reference(xd_f, xv_f, cb.cpu().repeat_interleave(S, 0), P)
    scale = float(emu.abs().max())
    err = (got - emu).abs()
    print(f"[anerf mlp] vs bf16 emulation: max {float(err.max()):.3e} mean {float(err.mean()):.3e} scale {scale:.3e}")
    assert float(err.mean()) <= 2e-4 * scale and float(err.max()) <= 2e-2 * scale
    want = fx["st.raw.0"].reshape(N * S, 4)
    err32 = (got - want).abs()
    print(f"[anerf mlp] vs reference fp32: max {float(err32.max()):.3e} mean {float(err32.mean()):.3e}")
    assert float(err32.max()) <= 5e-2 * scale and float(err32.mean()) <= 6e-3 * scale


def test_anerf_render_end_to_end():
    """create_raycaster(nerf_type='nerf') -> ray_caster(...) against the reference's rendered pixels."""
    fx = load_fixture("render_anerf")
    caster, args, P = make_anerf_caster(fx)
    skts, bones, cyl = pose_tensors(fx)
    rb = fx["ray_batch"].to(DEV)
    N = rb.shape[0]
    e = lambda t: t.to(DEV).expand(N, *t.shape[1:])
    st = {}
    ret = caster(rb, N_samples=args.N_samples, kp_batch=e(fx["pose_kps"][None]), skts=e(skts), cyls=e(cyl), bones=e(bones),
                 cams=fx["cams"].to(DEV), N_uniques=1, perturb=False, N_importance=args.N_importance, raw_noise_std=0.,
                 nerf_type="nerf", _stages=st)
    close(st["near"].cpu(), fx["st.near.0"].reshape(-1), 1e-6, "near")
    assert torch.equal(st["z_coarse"].cpu(), fx["st.z.0"])
    for k, tol_mean, tol_max in (("rgb0", 3e-3, 3e-2), ("acc0", 3e-3, 3e-2), ("rgb_map", 4e-3, 5e-2), ("acc_map", 4e-3, 5e-2)):
        d = (ret[k].cpu() - fx["out." + k]).abs()
        print(f"[anerf e2e] {k}: mean {float(d.mean()):.3e} max {float(d.max()):.3e}")
        assert float(d.mean()) <= tol_mean and float(d.max()) <= tol_max, (k, float(d.mean()), float(d.max()))
    # train mode without random draws evaluates the same pixels (the autograd node's forward is the same launch sequence)
    caster.train()
    tr = caster(rb, N_samples=args.N_samples, kp_batch=e(fx["pose_kps"][None]), skts=e(skts), cyls=e(cyl), bones=e(bones),
                cams=fx["cams"].to(DEV), N_uniques=1, perturb=False, N_importance=args.N_importance, nerf_type="nerf")
    assert tr["rgb_map"].requires_grad
    for k in ("rgb_map", "acc_map", "rgb0"):
        assert torch.equal(tr[k].detach(), ret[k]), k


def test_anerf_full_size_properties():
    """Config #4 at a full-size image block (512x512 crop of the workload, 96+48 samples): finite outputs, ranges,
    and chunk invariance (a reference 4096-ray chunk rendered alone gives the same bits)."""
    import danbo_b200 as db
    from danbo_b200 import synthetic as syn, skeleton as sk, params
    args = db.make_args("anerf_base", no_reload=True)
    attrs = {"skel_type": sk.SMPLSkeleton, "near": syn.NEAR, "far": syn.FAR, "n_views": 8, "rest_pose": syn.rest_pose()}
    _, kw_test, *_ = db.create_raycaster(args, attrs, device=DEV)
    caster = kw_test["ray_caster"]
    caster.network.load_state_dict(syn.synth_state_dict(params.anerf_param_shapes(), 0), strict=False)
    caster.eval()
    b = syn.render_batch(syn.make_pose(3), 512, 512)
    b = {k: (v.to(DEV) if torch.is_tensor(v) else v) for k, v in b.items()}
    N = b["ray_batch"].shape[0]
    kw = dict(N_samples=args.N_samples, kp_batch=b["kp_batch"], skts=b["skts"], cyls=b["cyls"], bones=b["bones"],
              cams=b["cams"], N_uniques=1, perturb=False, N_importance=args.N_importance, raw_noise_std=0., nerf_type="nerf")
    out = caster(b["ray_batch"], nanmean_chunk=4096, **kw)
    torch.cuda.synchronize()
    assert out["rgb_map"].shape == (N, 3)
    for k, v in out.items():
        assert torch.isfinite(v).all(), k
    assert float(out["acc_map"].min()) >= 0 and float(out["acc_map"].max()) <= 1.0
    assert float(out["T_i"].sum(-1).max()) <= 1.0 + 1e-4
    sl = slice(8192, 8192 + 4096)
    sub = caster(b["ray_batch"][sl], **{k: (v[sl] if torch.is_tensor(v) and v.shape[0] == N else v) for k, v in kw.items()})
    assert torch.equal(sub["rgb_map"], out["rgb_map"][sl]) and torch.equal(sub["acc_map"], out["acc_map"][sl])


def test_anerf_density_grid_and_points():
    """fwd_type='mesh' / 'density' for the A-NeRF field (D1): the reference's render_mesh_density on a 10^3 lattice."""
    fx = load_fixture("grid_anerf")
    caster, args, P = make_anerf_caster(fx)
    t = lambda a: a[None].to(DEV)
    sig = caster(kps=t(fx["pose_kps"]), skts=t(fx["pose_skts"]), bones=t(fx["pose_bones"]), radius=float(fx["radius"]),
                 res=int(fx["res"]), fwd_type="mesh")
    want = fx["sigma"]
    assert sig.shape == want.shape
    err = (sig.cpu() - want).abs()
    print(f"[anerf grid] sigma: mean {float(err.mean()):.3e} max {float(err.max()):.3e} scale {float(want.abs().max()):.3e}")
    assert float(err.max()) <= 3e-2 * float(want.abs().max()) and float(err.mean()) <= 4e-3 * float(want.abs().max())
    # point queries ('density') return the same numbers as the lattice they were taken from
    import numpy as np
    r, res = float(fx["radius"]), int(fx["res"])
    tt = np.linspace(-r, r, res + 1)
    grid = np.stack(np.meshgrid(tt, tt, tt), axis=-1).astype(np.float32).reshape(-1, 3)
    pts = torch.tensor(grid).to(DEV) + fx["pose_kps"][0].to(DEV)
    pd = caster(pts.reshape(-1, 1, 3), t(fx["pose_kps"]), t(fx["pose_skts"]), t(fx["pose_bones"]), fwd_type="density")
    assert pd.shape == (pts.shape[0], 1, 1)
    assert torch.equal(pd.reshape(res + 1, res + 1, res + 1).transpose(1, 0), sig)


def test_anerf_lindisp():
    """--lindisp on the A-NeRF path: coarse depths bit-identical to the oracle's (whose lindisp z is pinned by the
    reference-generated render_fast_lindisp fixture), pixels within the bf16 tolerance of the oracle's."""
    import danbo_oracle as orc
    from util import align_A
    fx = load_fixture("render_anerf")
    caster, args, P = make_anerf_caster(fx)
    skts, bones, cyl = pose_tensors(fx)
    rb = fx["ray_batch"]
    N = rb.shape[0]
    e = lambda t: t.to(DEV).expand(N, *t.shape[1:])
    st = {}
    ret = caster(rb.to(DEV), N_samples=args.N_samples, kp_batch=e(fx["pose_kps"][None]), skts=e(skts), cyls=e(cyl),
                 bones=e(bones), cams=fx["cams"].to(DEV), N_uniques=1, perturb=False, N_importance=args.N_importance,
                 raw_noise_std=0., nerf_type="nerf", lindisp=True, _stages=st)
    with torch.no_grad():
        want = orc.anerf_render_rays(rb, skts, cyl, fx["cams"], align_A(), {k: v.cpu() for k, v in P.items()},
                                     int(fx["N_samples"]), int(fx["N_importance"]), rays_per_pose=N, tau=float(fx["tau"]),
                                     lindisp=True, return_stages=True, z_samples=st["z_samples"].cpu())
    z_ref = want["_stages"]["z_coarse"]
    assert float((st["z_coarse"].cpu() - z_ref).abs().max()) <= 2e-6 * float(z_ref.abs().max())
    assert not torch.equal(z_ref, fx["st.z.0"])                      # really sampled in inverse depth
    for k, tol_mean, tol_max in (("rgb0", 3e-3, 3e-2), ("acc0", 3e-3, 3e-2), ("rgb_map", 4e-3, 5e-2), ("acc_map", 4e-3, 5e-2)):
        d = (ret[k].cpu() - want[k]).abs()
        print(f"[anerf lindisp] {k}: mean {float(d.mean()):.3e} max {float(d.max()):.3e}")
        assert float(d.mean()) <= tol_mean and float(d.max()) <= tol_max, (k, float(d.mean()), float(d.max()))


def test_anerf_separate_fine_network():
    """single_net = False (configs/h36m_zju/anerf_h.txt): unsmoothed importance weights, a second network evaluated on
    all S_c + S_f merged samples, against the reference's pixels."""
    import danbo_b200 as db
    from danbo_b200 import params, synthetic as syn, skeleton as sk
    fx = load_fixture("render_anerf_h")
    args = db.make_args("anerf_base", no_reload=True, single_net=False, N_samples=int(fx["N_samples"]),
                        N_importance=int(fx["N_importance"]))
    attrs = {"skel_type": sk.SMPLSkeleton, "near": syn.NEAR, "far": syn.FAR, "n_views": 8, "rest_pose": syn.rest_pose()}
    _, kw_test, _, grad_vars, *_ = db.create_raycaster(args, attrs, device=DEV)
    caster = kw_test["ray_caster"]
    assert caster.network_fine is not caster.network and not caster.single_net
    assert len(grad_vars) == 2 * len([p for p in caster.network.parameters() if p.requires_grad])
    seed = int(fx["weight_seed"])
    caster.network.load_state_dict(syn.synth_state_dict(params.anerf_param_shapes(), seed), strict=False)
    caster.network_fine.load_state_dict(syn.synth_state_dict(params.anerf_param_shapes(), seed + 1), strict=False)
    caster.eval()
    sd = caster.state_dict()
    assert not torch.equal(sd["network_fn_state_dict"]["alpha_linear.weight"], sd["network_fine_state_dict"]["alpha_linear.weight"])
    skts, bones, cyl = pose_tensors(fx)
    rb = fx["ray_batch"].to(DEV)
    N = rb.shape[0]
    e = lambda t: t.to(DEV).expand(N, *t.shape[1:])
    st = {}
    ret = caster(rb, N_samples=args.N_samples, kp_batch=e(fx["pose_kps"][None]), skts=e(skts), cyls=e(cyl), bones=e(bones),
                 cams=fx["cams"].to(DEV), N_uniques=1, perturb=False, N_importance=args.N_importance, raw_noise_std=0.,
                 nerf_type="nerf", _stages=st)
    assert torch.equal(st["z_coarse"].cpu(), fx["st.z.0"])
    for k, tol_mean, tol_max in (("rgb0", 3e-3, 3e-2), ("acc0", 3e-3, 3e-2), ("rgb_map", 4e-3, 5e-2), ("acc_map", 4e-3, 5e-2)):
        d = (ret[k].cpu() - fx["out." + k]).abs()
        print(f"[anerf_h e2e] {k}: mean {float(d.mean()):.3e} max {float(d.max()):.3e}")
        assert float(d.mean()) <= tol_mean and float(d.max()) <= tol_max, (k, float(d.mean()), float(d.max()))
    assert ret["alpha"].shape == fx["out.alpha"].shape and ret["T_i"].shape == fx["out.T_i"].shape
    # R1 with the unsmoothed weights (smooth=False), given identical weights: the oracle on the weights the kernel itself
    # composited from the reference's raw.  Without the +0.01 floor the pdf is ~1e-5 / sum in empty space, where the
    # `denom < 1e-5 -> 1` branch of sample_pdf (ray_utils.py:196-197) sits within an ulp of flipping, so a few samples may
    # land elsewhere in their bin: at least 90 % within 2e-6 of scale (an indexing error would leave ~0 %), all inside the
    # ray's depth range.
    S, S_f = args.N_samples, args.N_importance
    raw_g = torch.cat([fx["st.raw.0"].reshape(N * S, 4), torch.zeros(N, 4)], 0).to(DEV).contiguous()
    ones = torch.ones(N, S, dtype=torch.int32, device=DEV)
    c = K().composite_resample(rb, S, S_f, raw_g, ones, fx["st.z.0"].to(DEV), smooth=False)
    close(c["weights"].cpu(), fx["st.weights.0"], 2e-6, "coarse weights")
    _, zs, _, _ = orc.importance_sample(fx["st.z.0"], c["weights"].cpu(), S_f, is_only=False)
    dz = (c["z_samples"].cpu() - zs).abs() / float(zs.abs().max())
    print(f"[anerf_h] z_samples vs oracle on the same weights: within 2e-6: {float((dz <= 2e-6).float().mean()):.4f} max {float(dz.max()):.3e}")
    assert float((dz <= 2e-6).float().mean()) >= 0.9
    za = c["z_all"].cpu()
    assert bool((za[:, 1:] >= za[:, :-1]).all()) and float(za.min()) >= float(fx["st.z.0"].min()) - 1e-6 \
        and float(za.max()) <= float(fx["st.z.0"].max()) + 1e-6


def test_anerf_training_step_gradients():
    """Train-mode A-NeRF step (reference draws of tests/golden/train_anerf.npz): outputs, loss and the gradients of all
    25 parameter tensors against the oracle's autograd on the same samples (the oracle is pinned to the reference's own
    gradients by test_oracle_golden.py::test_anerf_training_step)."""
    from danbo_b200 import synthetic as syn
    fx = load_fixture("train_anerf")
    caster, args, _ = make_anerf_caster(fx)
    caster.train()
    n_poses, rpp = int(fx["n_poses"]), int(fx["rays_per_pose"])
    b = syn.training_batch(n_poses, rpp, seed=int(fx["batch_seed"]))
    rand = {k: fx["rand." + k] for k in ("t_rand", "noise0", "u", "noise1")}
    stages = {}
    out = caster.render_rays(b["ray_batch"], N_samples=args.N_samples, kp_batch=b["kp_batch"], skts=b["skts"], cyls=b["cyls"],
                             bones=b["bones"], cams=b["cams"], N_uniques=n_poses, perturb=1.0, N_importance=args.N_importance,
                             raw_noise_std=float(fx["raw_noise_std"]), nerf_type="nerf",
                             _rand={k: v.to(DEV) for k, v in rand.items()}, _stages=stages)
    assert set(out) == {"rgb_map", "disp_map", "acc_map", "alpha", "T_i", "rgb0", "disp0", "acc0", "alpha0"}
    tgt, bgs = b["target_s"].to(DEV), b["bgs"].to(DEV)
    l1 = lambda rgb, acc, t, g: torch.mean(torch.abs(rgb + (1. - acc)[..., None] * g - t))
    loss = l1(out["rgb_map"], out["acc_map"], tgt, bgs) + l1(out["rgb0"], out["acc0"], tgt, bgs)
    loss.backward()
    torch.cuda.synchronize()
    got = {n: p.grad.detach().cpu() for n, p in caster.network.named_parameters() if p.grad is not None}
    P = {k: v.clone().requires_grad_(v.dtype.is_floating_point) for k, v in anerf_params(fx).items()}
    ref = orc.anerf_render_rays(b["ray_batch"], b["skts"][::rpp], b["cyls"][::rpp], b["cams"], align_A(), P, int(fx["N_samples"]),
                                int(fx["N_importance"]), rays_per_pose=rpp, training=True, rand=rand,
                                raw_noise_std=float(fx["raw_noise_std"]), z_samples=stages["z_samples"].cpu())
    ref_loss = l1(ref["rgb_map"], ref["acc_map"], b["target_s"], b["bgs"]) + l1(ref["rgb0"], ref["acc0"], b["target_s"], b["bgs"])
    ref_loss.backward()
    print(f"[anerf train] loss cuda {float(loss):.6f} oracle {float(ref_loss):.6f} reference {float(fx['loss.total']):.6f}")
    assert abs(float(loss) - float(ref_loss)) <= 5e-3
    big = max(float(v.grad.norm()) for v in P.values() if v.grad is not None)
    checked, bad = 0, []
    for k, v in P.items():
        if v.grad is None:
            continue
        assert k in got, (k, "no gradient")
        a, r = got[k].reshape(-1).double(), v.grad.reshape(-1).double()
        cos = float(torch.dot(a, r) / (a.norm() * r.norm() + 1e-30))
        rel = float((a - r).norm() / max(float(r.norm()), 1e-3 * big))
        print(f"[anerf train] {k:32s} |g| {float(r.norm()):.3e} cos {cos:.5f} rel {rel:.3e}")
        checked += 1
        # bf16 forward + the reference's raw-noise gate (raw_noise_std = 1), as for the DANBO field's noisy fixtures
        if (float(r.norm()) > 1e-3 * big and cos < 0.97) or rel > 0.3:
            bad.append((k, cos, rel))
    assert checked == 25 and not bad, bad


def test_anerf_train_step_trains():
    """training.TrainStep on the A-NeRF caster: loss kernel, backward, single-launch Adam; the loss must go down."""
    import danbo_b200 as db
    from danbo_b200 import synthetic as syn, training
    fx = load_fixture("train_anerf")
    caster, args, _ = make_anerf_caster(fx)
    targs = db.make_args("anerf_base", no_reload=True, N_samples=int(fx["N_samples"]), N_importance=int(fx["N_importance"]))
    b = syn.training_batch(2, 48, seed=4)
    b = {k: (v.to(DEV) if torch.is_tensor(v) else v) for k, v in b.items()}
    step = training.TrainStep(caster, targs)
    torch.manual_seed(3)
    losses = [float(step(b)[0]) for _ in range(15)]
    print("[anerf train] losses", ["%.4f" % l for l in losses])
    assert all(l == l for l in losses) and losses[-1] < losses[0] - 0.01
