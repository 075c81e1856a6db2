"""Synthetic 24-joint poses, cameras and ray batches — no dataset exists offline.

pose k   : bones = RandomState(k).randn(24,3)*0.2 (axis-angle), FK on rest_pose = SMPL_REST_POSE*0.5
camera   : c2w = I with t_z = 3, focal = 1.2*H, pinhole rays as in core/utils/ray_utils.py:7-29
render   : rays restricted to the image-space box of the pose's bounding cylinder, the way
           render_path does through kp_to_valid_rays (core/utils/ray_utils.py:84-138,
           core/utils/skeleton_utils.py:633-720)
training : 16 poses x 192 rays, image-major (core/dataset.py:980-987), cylinders with the dataset's
           expand ratios (core/pose_opt.py:448-449)
"""
import numpy as np
import torch

from . import skeleton as sk

NEAR, FAR = 1.0, 5.0            # data_attrs near/far used by every synthetic scene


def rest_pose():
    return (sk.SMPL_REST_POSE * 0.5).astype(np.float32)


def make_pose(seed, rest=None, render_cylinder=True):
    """One pose -> dict(bones (24,3), kps (24,3), skts (24,4,4), cyl (5,)) float32."""
    rest = rest_pose() if rest is None else rest
    bones = (np.random.RandomState(seed).randn(sk.N_JOINTS, 3) * 0.2).astype(np.float32)
    l2ws = sk.forward_kinematics(bones, rest, 1.0)
    kps = l2ws[:, :3, 3].astype(np.float32)
    skts = np.linalg.inv(l2ws).astype(np.float32)
    if render_cylinder:
        cyl = sk.bounding_cylinder(kps, ext_scale=0.001, extend_mm=250, top_expand_ratio=1.6,
                                   bot_expand_ratio=1.1, head="-y")
    else:
        cyl = sk.bounding_cylinder(kps, ext_scale=0.001, extend_mm=250, head="-y")
    return {"bones": bones, "kps": kps, "skts": skts, "cyl": cyl.astype(np.float32)}


def camera(tz=3.0):
    c2w = np.eye(4, dtype=np.float32)
    c2w[2, 3] = tz
    return c2w


def bullet_time_cameras(c2w, n_views):
    """Rotate the camera about the world y axis (core/load_data.py:56-71)."""
    out = []
    for a in np.linspace(0, np.radians(360), n_views + 1)[:-1]:
        c, s = np.cos(a), np.sin(a)
        ry = np.array([[c, 0, -s, 0], [0, 1, 0, 0], [s, 0, c, 0], [0, 0, 0, 1]], dtype=np.float32)
        out.append(ry @ c2w)
    return np.array(out)


def pinhole_rays(H, W, focal, c2w):
    """rays_o, rays_d (H,W,3) float32 torch; d = (x/f, -y/f, -1) rotated by c2w (ray_utils.py:7-29)."""
    c2w = torch.as_tensor(c2w, dtype=torch.float32)
    xs = torch.linspace(0, W - 1, W)
    ys = torch.linspace(0, H - 1, H)
    j, i = torch.meshgrid(ys, xs, indexing="ij")
    dirs = torch.stack([(i - W * 0.5) / focal, -(j - H * 0.5) / focal, -torch.ones_like(i)], -1)
    rays_d = torch.sum(dirs[..., None, :] * c2w[:3, :3], -1)
    rays_o = c2w[:3, -1].expand(rays_d.shape)
    return rays_o, rays_d


def cylinder_image_box(cyl, H, W, focal, c2w):
    """Pixel box (tl, br) that covers the projected cylinder caps (skeleton_utils.py:633-720)."""
    rad = np.linspace(0.0, 2 * np.pi, 50)
    x = cyl[0] + np.cos(rad) * cyl[2]
    z = cyl[1] + np.sin(rad) * cyl[2]
    one = np.ones_like(x)
    caps = np.concatenate([np.stack([x, cyl[3] * one, z, one], -1), np.stack([x, cyl[4] * one, z, one], -1)], 0)
    flip = np.concatenate([c2w[..., 0:1], -c2w[..., 1:2], -c2w[..., 2:3], c2w[..., 3:]], axis=-1)
    w2c = np.linalg.inv(flip)
    K = np.array([[focal, 0, 0, 0], [0, focal, 0, 0], [0, 0, 1, 0]], dtype=np.float32)
    p = caps @ w2c.T @ K.T
    uv = p[:, :2] / p[:, 2:3]
    lo = np.floor(uv.min(0)).astype(np.int32) + np.array([int(W * .5), int(H * .5)], dtype=np.int32)
    hi = np.ceil(uv.max(0)).astype(np.int32) + np.array([int(W * .5), int(H * .5)], dtype=np.int32)
    lo = np.clip(lo, 0, [W - 1, H - 1])
    hi = np.clip(hi, 0, [W - 1, H - 1])
    return lo, hi


def render_rays_for_pose(pose, H, W, focal=None, c2w=None):
    """Box-restricted rays of one image: rays_o, rays_d (n,3) and the flat pixel index of each ray."""
    focal = 1.2 * H if focal is None else focal
    c2w = camera() if c2w is None else c2w
    ro, rd = pinhole_rays(H, W, float(focal), c2w)
    tl, br = cylinder_image_box(pose["cyl"], H, W, float(focal), np.asarray(c2w))
    hh, ww = torch.meshgrid(torch.arange(tl[1], br[1]), torch.arange(tl[0], br[0]), indexing="ij")
    idx = (hh * W + ww).reshape(-1)
    return ro.reshape(-1, 3)[idx].contiguous(), rd.reshape(-1, 3)[idx].contiguous(), idx


def render_batch(pose, H, W, focal=None, c2w=None, cam_idx=0, full_image=False):
    """Everything `render()` hands the ray caster for one image of one pose (run_nerf.py:64-91):
    ray_batch (n,11) = [o, d, near, far, viewdir], and stride-0 expands of the pose tensors."""
    focal = 1.2 * H if focal is None else focal
    if full_image:
        ro, rd = pinhole_rays(H, W, float(focal), camera() if c2w is None else c2w)
        ro, rd = ro.reshape(-1, 3).contiguous(), rd.reshape(-1, 3).contiguous()
        idx = torch.arange(H * W)
    else:
        ro, rd, idx = render_rays_for_pose(pose, H, W, focal, c2w)
    n = ro.shape[0]
    view = rd / torch.norm(rd, dim=-1, keepdim=True)
    ones = torch.ones(n, 1)
    rays = torch.cat([ro, rd, NEAR * ones, FAR * ones, view], -1)
    t = lambda a: torch.as_tensor(a)
    return {
        "ray_batch": rays,
        "kp_batch": t(pose["kps"])[None].expand(n, -1, -1),
        "skts": t(pose["skts"])[None].expand(n, -1, -1, -1),
        "bones": t(pose["bones"])[None].expand(n, -1, -1),
        "cyls": t(pose["cyl"])[None].expand(n, -1),
        "cams": torch.full((n, 1), cam_idx, dtype=torch.long),
        "pixel_idx": idx,
        "N_uniques": 1,
    }


def training_batch(n_poses=16, rays_per_pose=192, H=512, W=512, seed=0, n_views=8):
    """Image-major batch of n_poses*rays_per_pose rays with targets; rays are drawn inside each pose's box."""
    rng = np.random.RandomState(1000 + seed)
    rest = rest_pose()
    parts = {k: [] for k in ("ray_batch", "kp_batch", "skts", "bones", "cyls", "cams")}
    for p in range(n_poses):
        pose = make_pose(seed * n_poses + p, rest, render_cylinder=False)
        pose_r = dict(pose, cyl=make_pose(seed * n_poses + p, rest, True)["cyl"])
        ro, rd, _ = render_rays_for_pose(pose_r, H, W)
        pick = torch.as_tensor(rng.choice(ro.shape[0], rays_per_pose, replace=False))
        ro, rd = ro[pick], rd[pick]
        view = rd / torch.norm(rd, dim=-1, keepdim=True)
        ones = torch.ones(rays_per_pose, 1)
        parts["ray_batch"].append(torch.cat([ro, rd, NEAR * ones, FAR * ones, view], -1))
        t = lambda a: torch.as_tensor(a)
        parts["kp_batch"].append(t(pose["kps"])[None].expand(rays_per_pose, -1, -1))
        parts["skts"].append(t(pose["skts"])[None].expand(rays_per_pose, -1, -1, -1))
        parts["bones"].append(t(pose["bones"])[None].expand(rays_per_pose, -1, -1))
        parts["cyls"].append(t(pose["cyl"])[None].expand(rays_per_pose, -1))
        parts["cams"].append(torch.full((rays_per_pose, 1), p % n_views, dtype=torch.long))
    batch = {k: torch.cat(v, 0).contiguous() for k, v in parts.items()}
    n = n_poses * rays_per_pose
    batch["target_s"] = torch.as_tensor(rng.rand(n, 3).astype(np.float32))
    batch["bgs"] = torch.as_tensor(rng.rand(n, 3).astype(np.float32))
    batch["N_uniques"] = n_poses
    return batch


def synth_state_dict(shapes, seed=0):
    """Deterministic weights for parity tests: numpy RandomState is stable across versions and machines,
    so the same tensors can be rebuilt wherever the golden fixtures are checked.
    `shapes` = ordered {name: shape}; every tensor is N(0,1)*s with s chosen so activations stay O(1)."""
    rng = np.random.RandomState(seed)
    out = {}
    for name, shape in shapes.items():
        shape = tuple(shape)
        a = rng.standard_normal(shape).astype(np.float32)
        if name.endswith("axis_scale") or name.endswith("adj") or name.endswith("cutoff_dist") or name.endswith("tau"):
            continue                                        # geometry-defined / buffers: keep as constructed
        if name.endswith("adj_w"):
            a = np.abs(a) * 0.05 + 0.02
        elif name.endswith("bias") or len(shape) == 1:
            a = a * 0.1
        elif "framecodes" in name:
            a = a * 0.3
        elif name.startswith("alpha_linear"):               # make densities large enough to saturate some rays
            a = a * (40.0 / np.sqrt(shape[1]))
        elif len(shape) == 3:                               # ParallelLinear (joint, in, out)
            a = a * (1.6 / np.sqrt(shape[1]))
        else:                                               # nn.Linear (out, in)
            a = a * (1.4 / np.sqrt(shape[1]))
        out[name] = torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32))
    return out


def synthetic_params(seed=0, n_framecodes=8, rest=None, opt_framecode=True):
    """Full parameter dict (reference state_dict names) for the DANBO field: synthetic weights plus the
    geometry-defined entries (tree adjacency buffers, initial per-bone half extents)."""
    from . import params as _params
    rest = rest_pose() if rest is None else rest
    shapes = _params.danbo_param_shapes(n_framecodes=n_framecodes)
    P = synth_state_dict(shapes, seed)
    adj = torch.from_numpy(sk.skeleton_adjacency())[None]
    for name in _params.BUFFER_NAMES:
        P[name] = adj.clone()
    P["graph_net.axis_scale"] = sk.initial_axis_scale(sk.skeleton_profile(rest), base_scale=0.4)
    P = {k: P[k] for k in shapes}
    if not opt_framecode:
        # the same weights without the frame-code part: the code columns are the last 128 inputs of the view layer
        # ([feature ; view PE ; code], nerf.py:200-209), so existing fixtures' random streams are untouched
        P.pop("framecodes.codes.weight")
        P["views_linears.0.weight"] = P["views_linears.0.weight"][:, :-128].contiguous()
    return P
