"""The tensor-core aggregation net (csrc/field_mma.cu, `FieldConsts(pair_logits_impl="mma")`) against the fp32 FFMA
kernel and the oracle (run with -m gpu).  First run on hardware in round 2 (gpurun_out/r2a_*): 0.85 -> 0.57 ms for the
coarse field_agg call of a 512x512 image."""
import pytest
import torch

import danbo_oracle as orc
from util import load_fixture, params_for, align_A, make_caster, preset_of, agg_type_of, pose_tensors

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(300, method="thread")]   # a hung kernel must not hang the box
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
    # the active list is compacted with one atomic per block: its ORDER differs from run to run, its content does not
    pf, pm = torch.argsort(ids), torch.argsort(ids_m)
    assert torch.equal(ids[pf], ids_m[pm])
    hf_, hm = hf_[pf], hm[pm]
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
