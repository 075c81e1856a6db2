"""GPU checks of everything else that was written without GPU access at the end of round 1 (run with -m gpu):
the tensor-core aggregation net (csrc/field_mma.cu, `FieldConsts(pair_logits_impl="mma")`) against the fp32 FFMA kernel
and the oracle, the stream-pipelined ray blocks (`raycaster.BLOCK_STREAMS`), and mesh extraction on the device.

STATUS: written at the end of round 1 after the round's GPU minutes were spent; the kernel compiles for sm_100a (137
registers, no spills, 72 HMMA.16816 per instantiation) and its index arithmetic is checked on the CPU by
tests/test_pair_logits_mma_layout.py, but it has NOT run on hardware.  `xfail(strict=False)` and last in the `-m gpu`
order, so a failure cannot stop the `-x` run of the verified tests; remove the marker once green.  The FFMA kernel stays
the default until this one is both green and measured faster."""
import pytest
import torch

import danbo_oracle as orc
from util import load_fixture, params_for, align_A, make_caster, preset_of, agg_type_of, pose_tensors

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(300, method="thread"),   # a hung kernel must not hang the box
              pytest.mark.xfail(strict=False, reason="mma pair-logits kernel not yet run on hardware (written without GPU access)")]
DEV = "cuda"


def _consts(caster, impl):
    from danbo_b200 import kernels as K
    net = caster.network
    return K.FieldConsts(caster.transforms[0].to(DEV).float().contiguous(), net.graph_net.axis_scale, net.agg_tensors(),
                         pair_logits_impl=impl)


@pytest.mark.parametrize("name", ["render_fast", "render_base", "render_fast_softmax"])
def test_pair_logits_mma_matches_ffma(name):
    from danbo_b200 import kernels as K
    fx = load_fixture(name)
    agg = agg_type_of(fx)
    caster, args, _ = make_caster(preset_of(fx), agg_type=agg)
    Pc = params_for(fx)
    skts, bones, _ = pose_tensors(fx)
    rb, z = fx["ray_batch"], fx["st.z.0"]
    N, S = z.shape
    d = lambda t: t.to(DEV).contiguous()
    outs = {}
    for impl in ("ffma", "mma"):
        consts = _consts(caster, impl)
        assert (consts.frags is not None) == (impl == "mma")
        zg, mask, act = K.sample_mask(d(rb), S, d(skts), N, consts, z_in=d(z), append_empty=False)
        fo = K.field_agg(d(rb), S, zg, mask, act, d(skts), d(fx["st.vol.0"]), N, consts, want_hbar=True,
                         agg_mode=K.AGG_MODES[agg])
        torch.cuda.synchronize()
        n_act = int(act.count.item())
        outs[impl] = (fo.logits.cpu(), fo.hbar[:n_act].cpu(), act.ids[:n_act].cpu().long(), mask.cpu())
    lf, hf_, ids, mask = outs["ffma"]
    lm, hm, ids_m, _ = outs["mma"]
    assert torch.equal(ids, ids_m)
    vis = ((mask.long().reshape(-1, 1) >> torch.arange(24)) & 1).bool()                # (N*S, 24)
    sel = vis if agg == "sigmoid" else vis.any(-1, keepdim=True).expand(-1, 24)       # softmax: every bone of an active row
    scale = float(lf[sel].abs().max())
    err = float((lm[sel] - lf[sel]).abs().max())
    print(f"[mma] {name}: logits max |mma - ffma| {err:.3e} of scale {scale:.3e}; hbar {float((hm - hf_).abs().max()):.3e}")
    assert err <= 2e-5 * max(scale, 1.0)
    assert float((hm - hf_).abs().max()) <= 2e-5 * max(float(hf_.abs().max()), 1.0)
    # and against the oracle, at the tolerance the FFMA kernel is held to (tests/test_gpu_parity.py::test_field_agg)
    pts = orc.ray_points(rb[:, 0:3], rb[:, 3:6], z)
    pts_t = orc.world_to_bone(pts, skts.expand(N, -1, -1, -1), align_A())
    h, invalid, _ = orc.bone_features(pts_t, fx["st.vol.0"], Pc["graph_net.axis_scale"], rays_per_pose=N)
    a = orc.agg_net(h.reshape(N * S, 24, 15), Pc)
    same = (vis == (invalid.reshape(N * S, 24) == 0)).all(-1, keepdim=True).expand(-1, 24)
    ok = sel & same
    assert float((lm[ok] - a[ok]).abs().max()) <= 1e-4 * max(float(a[ok].abs().max()), 1.0)


def test_pair_logits_mma_multi_pose_and_speed():
    """Several poses in one call (per-lane pose tables, features from global memory) and a full-image timing of both
    kernels through the whole field_agg call (CUDA events; printed, not asserted)."""
    from danbo_b200 import kernels as K, synthetic as syn
    caster, args, _ = make_caster("danbo_fast")
    b = syn.training_batch(4, 96, seed=3)
    rpp = 96
    d = lambda t: t.to(DEV).contiguous().float()
    rays, skts = d(b["ray_batch"]), d(b["skts"][::rpp])
    vol = caster.network.bone_volumes(d(b["bones"][::rpp])).float().contiguous()
    res = {}
    for impl in ("ffma", "mma"):
        consts = _consts(caster, impl)
        near, far = K.nearfar(rays, d(b["cyls"][::rpp]), skts, rpp, consts.align, consts.axis_scale, use_box=True)
        zg, mask, act = K.sample_mask(rays, 32, skts, rpp, consts, near=near, far=far, append_empty=True)
        fo = K.field_agg(rays, 32, zg, mask, act, skts, vol, rpp, consts, want_hbar=True)
        torch.cuda.synchronize()
        res[impl] = (fo.logits.cpu(), mask.cpu())
    vis = ((res["ffma"][1].long().reshape(-1, 1) >> torch.arange(24)) & 1).bool()
    assert torch.equal(res["ffma"][1], res["mma"][1]) and int(vis.sum()) > 0
    err = float((res["mma"][0][vis] - res["ffma"][0][vis]).abs().max())
    assert err <= 2e-5 * max(float(res["ffma"][0][vis].abs().max()), 1.0), err
    # timing on a full image, one pose
    pose = syn.make_pose(3)
    rb = syn.render_batch(pose, 512, 512)
    rays = d(rb["ray_batch"])
    n = rays.shape[0]
    skts1 = d(torch.as_tensor(pose["skts"])[None])
    vol1 = caster.network.bone_volumes(d(torch.as_tensor(pose["bones"])[None])).float().contiguous()
    cyl1 = d(torch.as_tensor(pose["cyl"])[None])
    for impl in ("ffma", "mma"):
        consts = _consts(caster, impl)
        near, far = K.nearfar(rays, cyl1, skts1, n, consts.align, consts.axis_scale, use_box=True)
        zg, mask, act = K.sample_mask(rays, 32, skts1, n, consts, near=near, far=far, append_empty=True)
        for _ in range(3):
            K.field_agg(rays, 32, zg, mask, act, skts1, vol1, n, consts)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            K.field_agg(rays, 32, zg, mask, act, skts1, vol1, n, consts)
        e1.record()
        torch.cuda.synchronize()
        print(f"[mma] field_agg, 512x512 coarse pass, {impl}: {e0.elapsed_time(e1) / 10:.3f} ms")


def test_block_streams_same_pixels_and_speed():
    """Opt-in stream pipelining of ray blocks (raycaster.BLOCK_STREAMS): identical pixels, timing printed."""
    from danbo_b200 import synthetic as syn
    caster, args, _ = make_caster("danbo_fast")
    pose = syn.make_pose(3)
    rb = syn.render_batch(pose, 512, 512)
    kw = dict(N_samples=args.N_samples, kp_batch=rb["kp_batch"], skts=rb["skts"], cyls=rb["cyls"], bones=rb["bones"],
              cams=rb["cams"], N_uniques=1, perturb=False, N_importance=args.N_importance, raw_noise_std=0.,
              nanmean_chunk=4096)
    rays = rb["ray_batch"].to(DEV)
    res = {}
    for n in (1, 2, 3):
        caster.block_streams = n
        for _ in range(2):
            out = caster(rays, **kw)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            out = caster(rays, **kw)
        e1.record()
        torch.cuda.synchronize()
        res[n] = {k: v.clone() for k, v in out.items()}
        print(f"[streams] 512x512 danbo_fast eager, {n} stream(s): {e0.elapsed_time(e1) / 5:.3f} ms")
    for n in (2, 3):
        for k in ("rgb_map", "acc_map", "disp_map", "rgb0"):
            assert torch.equal(res[1][k], res[n][k]), (n, k)


def test_render_mesh_on_the_device():
    """`mesh.render_mesh` (run_render.py:1266-1281) on the real caster: the lattice of `fwd_type='mesh'` is extracted on
    the GPU; the surface must be closed and consistently oriented, and equal to the extraction of the same lattice on
    the CPU."""
    from danbo_b200 import mesh, synthetic as syn
    from test_mesh import check_closed_oriented
    caster, args, _ = make_caster("danbo_fast")
    pose = syn.make_pose(3)
    t = lambda a: torch.as_tensor(a)[None].to(DEV)
    res = 63
    raw = caster(kps=t(pose["kps"]), skts=t(pose["skts"]), bones=t(pose["bones"]), radius=1.0, res=res, fwd_type="mesh")
    sig = torch.relu(raw.reshape(res + 1, res + 1, res + 1))
    thr = float(sig.max()) * 0.25
    sig[0], sig[-1], sig[:, 0], sig[:, -1], sig[:, :, 0], sig[:, :, -1] = 0., 0., 0., 0., 0., 0.     # closed inside the lattice
    v, tri = mesh.marching_cubes(sig, thr)
    assert v.is_cuda and tri.shape[0] > 0
    check_closed_oriented(tri.cpu())
    v2, tri2 = mesh.marching_cubes(sig.cpu(), thr)
    assert torch.equal(tri.cpu(), tri2) and float((v.cpu() - v2).abs().max()) < 1e-5
    out = mesh.render_mesh(caster, t(pose["kps"]), t(pose["skts"]), t(pose["bones"]), radius=1.0, res=res, threshold=thr)
    assert len(out) == 1 and out[0][0].shape[1] == 3
    print(f"[mesh] {v.shape[0]} vertices, {tri.shape[0]} triangles at threshold {thr:.3f}")


@pytest.mark.parametrize("name", ["render_perfcap", "render_surreal"])
def test_render_other_shipped_configs(name):
    """configs/perfcap (view directions in the root joint's frame) and configs/surreal (no frame codes; `cams=None`) end
    to end against the reference's fixtures, at the bounds of test_gpu_parity.py::test_render_rays_end_to_end.  The view
    branch is the only difference from the verified path, and it enters through the per-ray bias of the view layer:
    that bias is checked directly first."""
    from danbo_b200 import kernels as K
    from util import config_flags_of
    fx = load_fixture(name)
    flags = config_flags_of(fx)
    caster, args, P = make_caster(preset_of(fx), **flags)
    assert caster.view_mode == ("root_local" if "perfcap" in name else "world")
    assert caster.network.opt_framecode == ("surreal" not in name)
    skts, bones, cyl = pose_tensors(fx)
    N = fx["ray_batch"].shape[0]
    ex = lambda t: t.expand(N, *t.shape[1:])
    stages = {}
    cams = fx["cams"] if caster.network.opt_framecode else None          # trainer.py:310
    out = caster(fx["ray_batch"], N_samples=args.N_samples, kp_batch=ex(fx["pose_kps"][None]), skts=ex(skts),
                 cyls=ex(cyl), bones=ex(bones), cams=cams, N_uniques=1, perturb=False,
                 N_importance=args.N_importance, raw_noise_std=0., _stages=stages)
    torch.cuda.synchronize()
    # V1: ray bias = W_v[:, 256:] . view_inputs + b_v, with view_inputs the reference's own per-ray tensor
    Pc = params_for(fx)
    Wv, bv = Pc["views_linears.0.weight"], Pc["views_linears.0.bias"]
    want_bias = fx["st.view_inputs.0"][:N] @ Wv[:, 256:].t() + bv
    got_bias = stages["ray_bias"].cpu()[: want_bias.shape[0]]
    err = float((got_bias - want_bias).abs().max())
    print(f"[configs] {name}: ray bias max err {err:.3e} of scale {float(want_bias.abs().max()):.3e}")
    assert err <= 2e-5 * max(float(want_bias.abs().max()), 1.0)
    assert float((stages["z_coarse"].cpu() - fx["st.z.0"]).abs().max()) <= 2e-6 * float(fx["st.z.0"].abs().max())
    for k in ("rgb0", "acc0", "rgb_map", "acc_map"):
        e = (out[k].cpu() - fx["out." + k]).abs()
        print(f"[configs] {name} {k}: mean {float(e.mean()):.3e} max {float(e.max()):.3e}")
        assert float(e.mean()) <= 4e-3 and float(e.flatten().quantile(0.99)) <= 5e-2 and float(e.max()) <= 0.2, k


@pytest.mark.parametrize("name", ["train_surreal", "train_perfcap"])
def test_training_step_other_shipped_configs(name):
    """configs/surreal training (MSE loss, no frame codes: the autograd node runs on a zero-padded view weight and the
    283-column gradient must come back) and configs/perfcap training (root-local view directions of four poses): loss and
    parameter gradients against the oracle's (bounds of test_gpu_training.py::test_training_step_gradients[train_fast])."""
    from danbo_b200 import synthetic as syn, skeleton as sk
    from util import config_flags_of, view_mode_of
    fx = load_fixture(name)
    loss_fn = str(fx.get("loss_fn", "L1"))
    has_codes = bool(int(fx.get("opt_framecode", 1)))
    caster, args, _ = make_caster(preset_of(fx), train=True, **config_flags_of(fx))
    n_poses, rpp = int(fx["n_poses"]), int(fx["rays_per_pose"])
    b = syn.training_batch(n_poses, rpp, seed=int(fx["batch_seed"]))
    rand = {k: fx["rand." + k] for k in ("t_rand", "noise0", "u", "noise1")}
    init_scale = sk.initial_axis_scale(sk.skeleton_profile(syn.rest_pose()), 0.4)
    stages = {}
    out = caster.render_rays(b["ray_batch"], N_samples=args.N_samples, kp_batch=b["kp_batch"], skts=b["skts"],
                             cyls=b["cyls"], bones=b["bones"], cams=b["cams"] if has_codes else None, N_uniques=n_poses,
                             perturb=1.0,
                             N_importance=args.N_importance, raw_noise_std=float(fx["raw_noise_std"]),
                             _rand={k: v.to(DEV) for k, v in rand.items()}, _stages=stages)
    P = {k: v.clone().requires_grad_(v.dtype.is_floating_point and not k.endswith(".adj")) for k, v in params_for(fx).items()}
    dev_out = {k: v for k, v in out.items()}
    loss = orc.training_loss(dev_out, b["target_s"].to(DEV), b["bgs"].to(DEV),
                             {"graph_net.axis_scale": caster.network.graph_net.axis_scale}, init_scale.to(DEV), loss_fn=loss_fn)
    loss.backward()
    torch.cuda.synchronize()
    ref = orc.render_rays(b["ray_batch"], b["skts"][::rpp], b["bones"][::rpp], b["cyls"][::rpp], b["cams"], align_A(), P,
                          int(fx["N_samples"]), int(fx["N_importance"]), rays_per_pose=rpp,
                          use_volume_near_far=bool(fx["use_volume_near_far"]), training=True, rand=rand,
                          raw_noise_std=float(fx["raw_noise_std"]), z_samples=stages["z_samples"].cpu(),
                          view_mode=view_mode_of(fx))
    ref_loss = orc.training_loss(ref, b["target_s"], b["bgs"], P, init_scale, loss_fn=loss_fn)
    ref_loss.backward()
    assert abs(float(loss) - float(ref_loss)) <= 5e-3
    g = caster.network.views_linears[0].weight.grad
    assert g is not None and g.shape == (128, 411 if has_codes else 283)
    big = max(float(v.grad.norm()) for v in P.values() if v.grad is not None)
    for k, v in P.items():
        if v.grad is None:
            continue
        a = dict(caster.network.named_parameters())[k].grad.detach().cpu().reshape(-1).double()
        r = v.grad.reshape(-1).double()
        cos = float(torch.dot(a, r) / (a.norm() * r.norm() + 1e-30))
        rel = float((a - r).norm() / max(float(r.norm()), 1e-3 * big))
        print(f"[{name}] {k:40s} cos {cos:.5f} rel {rel:.3e}")
        assert (float(r.norm()) <= 1e-3 * big or cos >= 0.985) and rel <= 0.2, (k, cos, rel)
