"""Gradients with respect to the poses (run with -m gpu): `danbo_field_agg_bwd`'s d skts output and
the whole training step's d loss / d (skts, bones), against torch autograd of the CPU oracle and against the gradients
the reference itself produced (tests/golden/train_fast_popt.npz; the oracle is pinned to them on the CPU by
test_oracle_golden.py::test_training_step_pose_gradients).

First run on hardware in round 2 (gpurun_out/r2a_gpu_unverified.log): all green."""
import pytest
import torch

import danbo_oracle as orc
from util import load_fixture, params_for, align_A, make_caster, preset_of, agg_type_of

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(300, method="thread")]   # a hung kernel must not hang the box
DEV = "cuda"


def _K():
    from danbo_b200 import kernels
    return kernels


@pytest.mark.parametrize("name,agg", [("render_fast", "sigmoid"), ("render_base", "sigmoid"), ("render_fast", "softmax")])
def test_field_backward_pose_gradient(name, agg):
    """Stage level, fp32 both sides: d skts of the feature gather + aggregation net for a random d X / d logits."""
    from util import pose_tensors
    fx = load_fixture(name)
    caster, args, Pdev = make_caster(preset_of(fx), train=True, agg_type=agg)
    K = _K()
    Pc = params_for(fx)
    skts, bones, _ = pose_tensors(fx)
    rb = fx["ray_batch"]
    N, z = rb.shape[0], fx["st.z.0"]
    S = z.shape[1]
    consts = caster._consts()
    vol = fx["st.vol.0"]
    d = lambda t: t.to(DEV).contiguous()
    zg, mask, act = K.sample_mask(d(rb), S, d(skts), N, consts, z_in=d(z), append_empty=1)
    fo = K.field_agg(d(rb), S, zg, mask, act, d(skts), d(vol), N, consts, want_hbar=True, want_xrows=True,
                     agg_mode=K.AGG_MODES[agg])
    n_act = int(act.count.item())
    ids = act.ids[:n_act].cpu().long()
    torch.manual_seed(5)
    dX = torch.zeros(act.capacity, 208); dX[:n_act, :195] = torch.randn(n_act, 195)
    g_ext = torch.randn(N * S, 24)
    keys = ["prob_linears.layers.0.lin.weight", "prob_linears.layers.0.adj_w", "prob_linears.layers.0.bias",
            "prob_linears.layers.1.weight", "prob_linears.layers.1.bias", "prob_linears.layers.2.weight",
            "prob_linears.layers.2.bias"]
    mk = lambda: [torch.zeros_like(Pdev[k]) for k in keys] + [torch.zeros(1, 24, 240, device=DEV), torch.zeros(24, 3, device=DEV)]
    grads, grads_plain = mk(), mk()
    d_skts = torch.zeros(1, 24, 4, 4, device=DEV)
    K.field_agg_bwd(d(rb), S, zg, mask, act, d(skts), d(vol), N, consts, fo, d(dX), d(g_ext), grads, d_skts=d_skts)
    K.field_agg_bwd(d(rb), S, zg, mask, act, d(skts), d(vol), N, consts, fo, d(dX), d(g_ext), grads_plain)
    torch.cuda.synchronize()
    for a, b in zip(grads, grads_plain):                              # the other gradients do not depend on the new output
        assert float((a - b).abs().max()) <= 1e-4 * max(float(b.abs().max()), 1e-6)
    # ---- oracle autograd with the world-to-bone matrices as a leaf
    P = {k: v.clone() for k, v in Pc.items()}
    skts_r = skts.clone().requires_grad_(True)
    pts = orc.ray_points(rb[:, 0:3], rb[:, 3:6], z)
    pts_t = orc.world_to_bone(pts, skts_r.expand(N, -1, -1, -1), align_A())
    h, invalid, _ = orc.bone_features(pts_t, vol, P["graph_net.axis_scale"], rays_per_pose=N)
    hf = h.reshape(N * S, 24, 15)
    a = orc.agg_net(hf, P)
    valid = 1 - invalid.reshape(N * S, 24)
    p = orc.agg_prob(a, invalid.reshape(N * S, 24), agg)
    X = orc.pe_embed((hf * p[..., None]).sum(-2), 6)
    real = ids < N * S
    rows_real = torch.nonzero(real).reshape(-1)
    loss = (X[ids[real]] * dX[rows_real, :195]).sum() + (a * valid * g_ext)[ids[real]].sum()
    loss.backward()
    want, got = skts_r.grad, d_skts.cpu()
    rel = float((got - want).norm() / want.norm())
    print(f"[popt] field bwd {name}/{agg} d skts |g| {float(want.norm()):.3e} rel {rel:.3e}")
    assert float(got[:, :, 3].abs().max()) == 0.0
    assert rel <= 2e-4, rel


def test_training_step_pose_gradients():
    """Whole step: the pose tensors enter `RayCaster` as leaves (what the pose layer's outputs are under --opt_pose);
    d loss / d skts comes from the kernels, d loss / d bones through the graph net's PyTorch ops."""
    from danbo_b200 import synthetic as syn, skeleton as sk
    from test_gpu_training import torch_loss
    fx = load_fixture("train_fast_popt")
    agg = agg_type_of(fx)
    caster, args, _ = make_caster(preset_of(fx), train=True, agg_type=agg)
    n_poses, rpp = int(fx["n_poses"]), int(fx["rays_per_pose"])
    b = syn.training_batch(n_poses, rpp, seed=int(fx["batch_seed"]))
    rand = {k: fx["rand." + k] for k in ("t_rand", "noise0", "u", "noise1")}
    init_scale = sk.initial_axis_scale(sk.skeleton_profile(syn.rest_pose()), 0.4)
    skts_leaf = b["skts"].clone().to(DEV).requires_grad_(True)           # (n,24,4,4) per ray, like the reference's batch
    bones_leaf = b["bones"].clone().to(DEV).requires_grad_(True)
    stages = {}
    out = caster.render_rays(b["ray_batch"], N_samples=args.N_samples, kp_batch=b["kp_batch"], skts=skts_leaf,
                             cyls=b["cyls"], bones=bones_leaf, cams=b["cams"], N_uniques=n_poses, perturb=1.0,
                             N_importance=args.N_importance, raw_noise_std=float(fx["raw_noise_std"]),
                             _rand={k: v.to(DEV) for k, v in rand.items()}, _stages=stages)
    loss = torch_loss(out, b["target_s"].to(DEV), b["bgs"].to(DEV), caster.network.graph_net.axis_scale,
                      init_scale.to(DEV), agg)
    loss.backward()
    torch.cuda.synchronize()
    per_pose = lambda g: g.reshape(n_poses, rpp, *g.shape[1:]).sum(1).cpu()
    got_skts, got_bones = per_pose(skts_leaf.grad), per_pose(bones_leaf.grad)
    # parameters still receive their gradients on this route
    assert all(p.grad is not None for n, p in caster.network.named_parameters() if n.startswith("graph_net.layers"))
    # ---- oracle autograd at the kernel path's own importance samples
    P = params_for(fx)
    pose_skts = b["skts"][::rpp].clone().requires_grad_(True)
    pose_bones = b["bones"][::rpp].clone().requires_grad_(True)
    ref = orc.render_rays(b["ray_batch"], pose_skts, pose_bones, b["cyls"][::rpp], b["cams"], align_A(), P,
                          int(fx["N_samples"]), int(fx["N_importance"]), rays_per_pose=rpp,
                          use_volume_near_far=bool(fx["use_volume_near_far"]), training=True, rand=rand,
                          raw_noise_std=float(fx["raw_noise_std"]), z_samples=stages["z_samples"].cpu(), agg_type=agg)
    ref_loss = orc.training_loss(ref, b["target_s"], b["bgs"], P, init_scale, agg_type=agg)
    ref_loss.backward()
    assert abs(float(loss) - float(ref_loss)) <= 5e-3
    for nm, got, want, fxk in (("skts", got_skts, pose_skts.grad, "pose_grad.skts"),
                               ("bones", got_bones, pose_bones.grad, "pose_grad.bones")):
        a, r, f = got.reshape(-1).double(), want.reshape(-1).double(), fx[fxk].reshape(-1).double()
        cos = float(torch.dot(a, r) / (a.norm() * r.norm() + 1e-30))
        rel = float((a - r).norm() / r.norm())
        cos_ref = float(torch.dot(a, f) / (a.norm() * f.norm() + 1e-30))
        print(f"[popt] step d {nm}: |g| {float(r.norm()):.3e} cos {cos:.5f} rel {rel:.3e} | cos vs reference {cos_ref:.5f}")
        # same bounds as the parameter gradients of train_fast (bf16 forward MLP, noise-gate flips)
        assert cos >= 0.985 and rel <= 0.2, (nm, cos, rel)
        assert cos_ref >= 0.95, (nm, cos_ref)


def test_train_step_with_pose_layer_and_feed():
    """--opt_pose iteration end to end: device-resident feed -> pose layer -> render block -> losses (+ pose regulariser)
    -> both optimisers.  The poses of the frames in the batch must move, stay finite, and the loss must not blow up."""
    from danbo_b200 import feed as fd, pose_opt as po, training, synthetic as syn
    caster, args, _ = make_caster("danbo_fast", train=True)
    arrays = fd.synthetic_arrays(n_images=6, H=64, W=64, seed=2)
    feed = fd.RayFeed.from_arrays(arrays, syn.NEAR, syn.FAR, N_rand=4 * 48, N_sample_images=4, device=DEV, seed=1)
    args.opt_pose_coef, args.opt_pose_tol, args.opt_pose_lrate = 1.0, 0.0, 1e-3
    attrs = {"rest_pose": syn.rest_pose()[None], "betas": torch.zeros(1, 10).numpy(), "kp3d": arrays["kp3d"],
             "bones": arrays["bones"]}
    pose_optimizer, kw = po.create_popt(args, attrs, device=DEV)
    layer = kw["popt_layer"]
    before = layer.bones.detach().clone()
    step = training.TrainStep(caster, args, popt_kwargs=kw, pose_optimizer=pose_optimizer)
    g = torch.Generator(device=DEV).manual_seed(0)
    losses, touched = [], set()
    for _ in range(6):
        batch = feed.next_batch(g)
        touched.update(feed.last_idxs[0].tolist())
        loss, preds = step(batch)
        losses.append(float(loss))
        assert preds["rgb_map"].shape == (4 * 48, 3)
    torch.cuda.synchronize()
    print("[popt] losses", [f"{v:.4f}" for v in losses], "MPJPC", float(step.last_stats["MPJPC"]))
    assert all(l == l and l < 10 for l in losses)
    moved = (layer.bones.detach() - before).abs().amax(dim=(1, 2)).cpu()
    assert bool(torch.isfinite(layer.bones).all()) and bool(torch.isfinite(layer.pelvis).all())
    for i in range(6):
        assert (float(moved[i]) > 0) == (i in touched), (i, float(moved[i]), sorted(touched))
