"""Pose layer (`pose_opt.PoseOptLayer`) against the reference's core/pose_opt.py.

CPU only (the layer is torch ops).  `tests/golden/popt.npz` holds state dicts, outputs, parameter gradients and the
trainer's pose regulariser from the UNMODIFIED reference (oracle/gen_golden_popt.py)."""
import os
import sys
import types

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import danbo_b200                                    # noqa: E402,F401
from danbo_b200 import pose_opt as po                # noqa: E402
from danbo_b200 import skeleton as sk                # noqa: E402

FX = np.load(os.path.join(ROOT, "tests", "golden", "popt.npz"))
T = lambda k: torch.tensor(FX[k])


def _layer(tag):
    sd = {k[len(tag) + 4:]: T(k) for k in FX.files if k.startswith(tag + ".sd.")}
    layer = po.load_poseopt_from_state_dict({"poseopt_layer_state_dict": sd})
    assert set(layer.state_dict().keys()) == set(sd.keys())                       # the reference's key scheme
    for k, v in layer.state_dict().items():
        assert torch.equal(v, sd[k])
    return layer


def _probe_weights(n, seed=0):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(n, 24, 3, generator=g), torch.randn(n, 24, 4, 4, generator=g)


@pytest.mark.parametrize("tag", ["axisang", "rot6d", "multiview"])
@pytest.mark.parametrize("uniques", [None, 4])
def test_forward_and_gradients_match_reference(tag, uniques):
    layer = _layer(tag)
    assert layer.use_rot6d == (tag == "rot6d")
    idxs = FX["idxs"]
    outs = layer(idxs, N_uniques=uniques)
    for nm, t in zip(("kps", "bones", "skts", "l2ws", "rots"), outs):
        ref = T(f"{tag}.out.{nm}")
        assert t.shape == ref.shape, nm
        err = float((t.detach() - ref).abs().max())
        assert err <= 2e-6 * max(1.0, float(ref.abs().max())), (nm, err)          # skts: rigid inverse vs torch.inverse
    wk, ws = _probe_weights(len(idxs))
    ((outs[0] * wk).sum() + (outs[2] * ws).sum()).backward()
    for nm, p in layer.named_parameters():
        ref = T(f"{tag}.grad.{nm}")
        err = float((p.grad - ref).abs().max())
        assert err <= 2e-5 * max(1.0, float(ref.abs().max())), (nm, err)


def test_chain_equals_numpy_forward_kinematics_and_inverse():
    """The level-parallel chain against the per-joint float64 chain the render data path uses (skeleton.forward_kinematics)."""
    rng = np.random.RandomState(0)
    bones = (rng.randn(5, 24, 3) * 0.4).astype(np.float32)
    rest = (sk.SMPL_REST_POSE * 0.5).astype(np.float32)
    kps, _, skts, l2ws, rots = po.get_kinematic_chain_T(torch.tensor(rest), torch.tensor(bones))
    ref = np.stack([sk.forward_kinematics(b, rest) for b in bones])
    assert np.abs(l2ws.numpy() - ref).max() < 2e-6
    eye = (skts @ l2ws).numpy()
    assert np.abs(eye - np.eye(4)).max() < 2e-6
    assert np.abs(kps.numpy() - ref[..., :3, 3]).max() < 2e-6
    assert [len(ids) for ids, _ in po._levels(sk.JOINT_PARENTS)] == [3, 3, 3, 5, 3, 2, 2, 2]   # pose_opt.py:374-413


def test_rotation_conversions_round_trip():
    rng = np.random.RandomState(1)
    aa = torch.tensor(rng.randn(200, 3).astype(np.float32))
    aa[:5] *= 1e-8                                                                    # small-angle branch
    aa[5] = torch.tensor([3.1, 0.2, -0.1])                                            # near pi
    R = po.axisang_to_rot(aa)
    assert float((R @ R.transpose(-1, -2) - torch.eye(3)).abs().max()) < 1e-5
    assert float((torch.linalg.det(R) - 1).abs().max()) < 1e-5
    assert np.abs(R.numpy() - sk.rodrigues(aa.numpy())).max() < 2e-6                  # quaternion route == Rodrigues
    assert float((po.axisang_to_rot(po.rot_to_axisang(R)) - R).abs().max()) < 1e-5
    assert float((po.rot6d_to_rotmat(po.rot_to_rot6d(R)) - R).abs().max()) < 1e-5
    back = po.rot_to_axisang(R)
    keep = aa.norm(dim=-1) < 3.0                                                      # below pi the vector itself returns
    assert float((back[keep] - aa[keep]).abs().max()) < 1e-5


@pytest.mark.parametrize("tag", ["axisang", "rot6d"])
def test_pose_regulariser_matches_trainer(tag):
    layer = _layer(tag)
    idxs = FX["idxs"]
    bones0 = T("init_bones")
    anchors = {"kps": T("init_kps"), "bones": bones0, "rots": po.axisang_to_rot(bones0.reshape(-1, 3)).reshape(-1, 24, 3, 3)}
    args = types.SimpleNamespace(opt_rot6d=tag == "rot6d", opt_pose_tol=0.002, opt_pose_coef=2.0, use_temp_loss=True,
                                 temp_coef=0.05, ext_scale=0.001)
    k, b, _, _, r = layer(idxs)
    losses, stats = po.kp_loss(args, anchors, idxs, {"kp_batch": k, "bones": b, "rots": r}, popt_layer=layer,
                               temp_val=T(f"{tag}.loss.temp_val"))
    for nm, got in (("kp_loss", losses["kp_loss"]), ("temp_loss", losses["temp_loss"]), ("MPJPC", stats["MPJPC"])):
        ref = float(FX[f"{tag}.loss.{nm}"])
        assert abs(float(got) - ref) <= 2e-5 * max(1.0, abs(ref)), (nm, float(got), ref)


def test_create_popt_and_checkpoint_round_trip(tmp_path):
    args = types.SimpleNamespace(opt_rot6d=True, opt_pose_lrate=1e-3, init_poseopt=None, no_poseopt_reload=False,
                                 use_ckpt_anchor=True, opt_pose_cache=True)
    attrs = {"rest_pose": FX["rest_pose"], "betas": np.zeros((1, 10), np.float32), "kp3d": FX["init_kps"],
             "bones": FX["init_bones"]}
    opt, kw = po.create_popt(args, attrs)
    layer = kw["popt_layer"]
    assert set(kw) == {"popt_anchors", "popt_layer", "skel_type"} and set(kw["popt_anchors"]) == {"kps", "bones", "rots", "beta"}
    assert isinstance(opt, torch.optim.Adam) and opt.param_groups[0]["lr"] == 1e-3
    assert layer.bones.shape == (7, 24, 6) and layer.use_cache
    k, b, s, l, r = layer(FX["idxs"])                                                 # served from the cache
    assert not k.requires_grad and k.shape == (12, 24, 3)
    # initial kps: pelvis = root keypoint, chain from the bones -> keypoints that differ from kp3d only by the synthetic
    # per-pose shift being applied at the root (make_pose keypoints + shift)
    assert float((k[:, 0] - torch.tensor(FX["init_kps"])[FX["idxs"], 0] - torch.tensor(FX["rest_pose"])[0, 0]).abs().max()) < 1e-6
    # pose checkpoint -> render data (pose_opt.py:415-451)
    path = os.path.join(tmp_path, "pose.tar")
    torch.save({"poseopt_layer_state_dict": layer.state_dict()}, path)
    kp3d, bones, skts, cyls, rest, pelvis = po.pose_ckpt_to_pose_data(path)
    assert kp3d.dtype == np.float32 and skts.shape == (7, 24, 4, 4) and cyls.shape == (7, 5) and bones.shape == (7, 24, 3)
    full = layer.calculate_kinematic(np.arange(7))
    assert np.abs(kp3d - full[0].detach().numpy()).max() < 2e-6
    assert np.abs(skts - full[2].detach().numpy()).max() < 5e-6
    assert np.abs(bones - FX["init_bones"]).max() < 1e-5                              # rot6d -> axis-angle recovers the input
    assert np.abs(po.load_bones_from_state_dict({"poseopt_layer_state_dict": layer.state_dict()}).numpy() - FX["init_bones"]).max() < 1e-5
    # anchors reloaded from a checkpoint (use_ckpt_anchor)
    args.init_poseopt = path
    _, kw2 = po.create_popt(args, attrs)
    assert float((kw2["popt_anchors"]["kps"] - full[0].detach()).abs().max()) < 1e-6
    with pytest.raises(NotImplementedError):
        po.pose_ckpt_to_pose_data(path, legacy=True)


def test_train_step_with_pose_layer_on_a_stand_in_caster():
    """`training.TrainStep(popt_kwargs=..., pose_optimizer=...)` plumbing on the CPU with a stand-in ray caster (the real
    one needs a GPU): poses come from the layer by `kp_idx`, the regulariser joins the loss, both optimisers step."""
    from danbo_b200 import training

    class Net(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.w = torch.nn.Parameter(torch.full((3,), 0.5))

    class Caster(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.network = Net()
            self.seen = None

        def forward(self, ray_batch, kp_batch=None, skts=None, bones=None, N_uniques=1, **kw):
            self.seen = dict(kp_batch=kp_batch, skts=skts, bones=bones, N_uniques=N_uniques)
            rgb = torch.sigmoid(skts[:, 3, :3, 3] * self.network.w + bones[:, 5, :3].sum(-1, keepdim=True))
            acc = torch.sigmoid(kp_batch[:, 7, 0])
            return {"rgb_map": rgb, "acc_map": acc}

    n_frames, B, R = 7, 3, 4
    args = types.SimpleNamespace(opt_rot6d=False, opt_pose_lrate=1e-2, init_poseopt=None, no_poseopt_reload=False,
                                 use_ckpt_anchor=False, opt_pose_cache=False, opt_pose_tol=0.0, opt_pose_coef=2.0,
                                 use_temp_loss=False, ext_scale=0.001, lrate=1e-2, loss_fn="L1", agg_type="sigmoid",
                                 N_samples=8, N_importance=4, perturb=1.0, raw_noise_std=0., use_background=False,
                                 opt_vol_scale=False)
    attrs = {"rest_pose": FX["rest_pose"], "betas": np.zeros((1, 10), np.float32), "kp3d": FX["init_kps"],
             "bones": FX["init_bones"]}
    pose_opt_, kw = po.create_popt(args, attrs)
    layer = kw["popt_layer"]
    caster = Caster()
    step = training.TrainStep(caster, args, popt_kwargs=kw, pose_optimizer=pose_opt_)
    kp_idx = torch.tensor([1, 4, 6]).repeat_interleave(R)
    batch = {"ray_batch": torch.zeros(B * R, 11), "kp_idx": kp_idx, "N_uniques": B, "cams": torch.zeros(B * R, 1),
             "cyls": torch.zeros(B * R, 5), "target_s": torch.full((B * R, 3), 0.25),
             # what the data feed put there; the layer's outputs must replace them
             "kp_batch": torch.full((B * R, 24, 3), 9.), "skts": torch.full((B * R, 24, 4, 4), 9.),
             "bones": torch.full((B * R, 24, 3), 9.)}
    before = {k: v.detach().clone() for k, v in layer.state_dict().items()}
    w0 = caster.network.w.detach().clone()
    with torch.no_grad():
        layer.bones[4, 9] += 0.3                                     # off the anchor -> the regulariser is active
    loss1, _ = step(batch)
    want = layer.calculate_kinematic(kp_idx.numpy())
    assert caster.seen["N_uniques"] == B and caster.seen["skts"].shape == (B * R, 24, 4, 4)
    assert float(caster.seen["kp_batch"].abs().max()) < 9.           # the layer's poses, not the feed's
    assert "MPJPC" in step.last_stats and float(step.last_stats["MPJPC"]) > 0
    after = layer.state_dict()
    touched = [1, 4, 6]
    rest = [i for i in range(n_frames) if i not in touched]
    assert float((after["bones"][touched] - before["bones"][touched]).abs().max()) > 0          # pose optimiser stepped
    assert float((after["pelvis"][touched] - before["pelvis"][touched]).abs().max()) > 0
    assert torch.equal(after["bones"][rest], before["bones"][rest])                             # frames not in the batch untouched
    assert not torch.equal(caster.network.w.detach(), w0)                                       # network optimiser stepped
    assert float((after["bones"][4, 9] - FX["init_bones"][4, 9]).abs().max()) < 0.3 + 1e-6      # pulled back towards the anchor
    losses = [float(loss1)] + [float(step(batch)[0]) for _ in range(20)]
    assert losses[-1] < losses[0]
    with pytest.raises(NotImplementedError):
        training.TrainStep(caster, args, popt_kwargs=kw, pose_optimizer=pose_opt_, graph=True)


def test_cache_follows_the_parameters_and_rest_pose_tables():
    kps, bones = T("init_kps"), T("init_bones")
    layer = po.PoseOptLayer(kps, bones, T("rest_pose"), use_cache=True)
    a = layer(FX["idxs"])
    with torch.no_grad():
        layer.pelvis += 1.0
    assert torch.equal(layer(FX["idxs"])[0], a[0])                               # served from the (stale) cache
    layer.update_cache()
    assert float((layer(FX["idxs"])[0] - a[0] - 1.0).abs().max()) < 1e-6
    layer.double()                                                                # _apply refreshes the cache
    assert layer(FX["idxs"])[0].dtype == torch.float64
    # several rest poses, chosen per frame (pose_opt.py:256-262)
    rest2 = torch.cat([T("rest_pose"), T("rest_pose") * 1.1], 0)
    multi = po.PoseOptLayer(kps, bones, rest2, rest_pose_idxs=np.array([0, 1, 0, 1, 0, 1, 0]))
    k = multi(np.array([1, 2]))[0]
    one = po.PoseOptLayer(kps, bones, T("rest_pose") * 1.1)(np.array([1]))[0]
    assert float((k[0] - one[0]).abs().max()) < 1e-6
    assert float((k[1] - po.PoseOptLayer(kps, bones, T("rest_pose"))(np.array([2]))[0][0]).abs().max()) < 1e-6
