"""SMPL skeleton constants and the setup-time geometry of the hot path (host side, numpy).

Everything here runs once per model or once per pose on the CPU; none of it is a kernel target.
It restates, in this repo's own words, what the reference computes in
  core/utils/skeleton_utils.py:83-110   (SMPL joint tree)
  core/utils/skeleton_utils.py:259-282  (rest pose table)
  core/utils/skeleton_utils.py:334-376  (forward kinematics, get_smpl_l2ws)
  core/utils/skeleton_utils.py:544-566  (rotation that aligns a bone with +z)
  core/utils/skeleton_utils.py:568-618  (bounding cylinder of a pose)
  core/utils/skeleton_utils.py:1515-1568 (skeleton profile: widths and bone lengths)
  core/raycasters.py:548-591            (per-joint bone-align transforms, row S0 of SURVEY §8a)
  core/networks/misc.py:675-724         (initial per-bone volume half extents, row S1)
"""
from collections import namedtuple
import numpy as np

N_JOINTS = 24

JOINT_NAMES = (
    "pelvis", "left_hip", "right_hip", "spine1", "left_knee", "right_knee", "spine2", "left_ankle",
    "right_ankle", "spine3", "left_foot", "right_foot", "neck", "left_collar", "right_collar", "head",
    "left_shoulder", "right_shoulder", "left_elbow", "right_elbow", "left_wrist", "right_wrist",
    "left_hand", "right_hand",
)
# parent of every joint (root is its own parent)
JOINT_PARENTS = np.array([0, 0, 0, 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 9, 9, 12, 13, 14, 16, 17, 18, 19, 20, 21])

SkeletonType = namedtuple("SkeletonType", ["joint_names", "joint_trees", "root_id", "nonroot_id", "end_effectors"])
SMPLSkeleton = SkeletonType(joint_names=list(JOINT_NAMES), joint_trees=JOINT_PARENTS, root_id=0,
                            nonroot_id=list(range(1, N_JOINTS)), end_effectors=[10, 11, 15, 22, 23])

# SMPL rest pose (x, y, z) in the reference's unit (data table, skeleton_utils.py:259-282)
SMPL_REST_POSE = np.array([
    [0.00000000e+00, 2.30003661e-09, -9.86228770e-08], [1.63832515e-01, -2.17391014e-01, -2.89178602e-02],
    [-1.57855421e-01, -2.14761734e-01, -2.09642015e-02], [-7.04505108e-03, 2.50450850e-01, -4.11837511e-02],
    [2.42021069e-01, -1.08830070e+00, -3.14962119e-02], [-2.47206554e-01, -1.10715497e+00, -3.06970738e-02],
    [3.95125849e-03, 5.94849110e-01, -4.03754264e-02], [2.12680623e-01, -1.99382353e+00, -1.29327580e-01],
    [-2.10857525e-01, -2.01218796e+00, -1.23002514e-01], [9.39484313e-03, 7.19204426e-01, 2.06931755e-02],
    [2.63385147e-01, -2.12222481e+00, 1.46775618e-01], [-2.51970559e-01, -2.12153077e+00, 1.60450473e-01],
    [3.83779174e-03, 1.22592449e+00, -9.78838727e-02], [1.91201791e-01, 1.00385976e+00, -6.21964522e-02],
    [-1.77145526e-01, 9.96228695e-01, -7.55542740e-02], [1.68482102e-02, 1.38698268e+00, 2.44048554e-02],
    [4.01985168e-01, 1.07928419e+00, -7.47655183e-02], [-3.98825467e-01, 1.07523870e+00, -9.96334553e-02],
    [1.00236952e+00, 1.05217218e+00, -1.35129794e-01], [-9.86728609e-01, 1.04515052e+00, -1.40235111e-01],
    [1.56646240e+00, 1.06961894e+00, -1.37338534e-01], [-1.56946480e+00, 1.05935931e+00, -1.53905824e-01],
    [1.75282109e+00, 1.04682994e+00, -1.68231070e-01], [-1.75758195e+00, 1.04255080e+00, -1.77773550e-01],
], dtype=np.float32)


def joint_children(parents=JOINT_PARENTS):
    kids = [[] for _ in parents]
    for j, p in enumerate(parents):
        kids[p].append(j)          # note: the root lists itself as a child, as the reference does
    return kids


def skeleton_adjacency(parents=JOINT_PARENTS):
    """Identity plus parent/child edges, (24,24) float32 (gnn_backbone.py:18-34)."""
    n = len(parents)
    adj = np.eye(n, dtype=np.float32)
    for j, p in enumerate(parents):
        if j != p:
            adj[j, p] = adj[p, j] = 1.0
    return adj


def rodrigues(rotvec):
    """Axis-angle (…,3) -> rotation matrices (…,3,3), float64 (what scipy's from_rotvec returns)."""
    rv = np.asarray(rotvec, dtype=np.float64)
    theta = np.linalg.norm(rv, axis=-1, keepdims=True)
    small = theta < 1e-12
    axis = rv / np.where(small, 1.0, theta)
    x, y, z = axis[..., 0], axis[..., 1], axis[..., 2]
    zero = np.zeros_like(x)
    K = np.stack([zero, -z, y, z, zero, -x, -y, x, zero], -1).reshape(rv.shape[:-1] + (3, 3))
    s, c = np.sin(theta)[..., None], np.cos(theta)[..., None]
    return np.eye(3) + s * K + (1.0 - c) * (K @ K)


def forward_kinematics(bones, rest_pose, scale=1.0, parents=JOINT_PARENTS):
    """Local-to-world 4x4 per joint for one pose: chain of [R_j | rest_j - rest_parent] (skeleton_utils.py:334-376).
    Returns float64 (24,4,4); the reference gets float64 too because scipy rotations are float64."""
    rest = np.asarray(rest_pose) * scale
    R = rodrigues(bones)
    out = []
    for j in range(len(parents)):
        rel = np.eye(4)
        rel[:3, :3] = R[j]
        if j == 0:
            rel[:3, 3] = rest[0]
            out.append(rel)
        else:
            p = parents[j]
            rel[:3, 3] = rest[j] - rest[p]
            out.append(out[p] @ rel)
    return np.stack(out)


def _rot_about_y(t):
    c, s = np.cos(t), np.sin(t)
    return np.array([[c, 0, -s, 0], [0, 1, 0, 0], [s, 0, c, 0], [0, 0, 0, 1]], dtype=np.float32)


def _rot_about_x(t):
    c, s = np.cos(t), np.sin(t)
    return np.array([[1, 0, 0, 0], [0, c, -s, 0], [0, s, c, 0], [0, 0, 0, 1]], dtype=np.float32)


def _acos_clipped(a):
    return np.arccos(np.clip(a, -1.0 + 1e-8, 1.0 - 1e-8))


def bone_to_z_rotation(vec):
    """4x4 rotation that maps `vec` onto the +z axis: first about y, then about x (skeleton_utils.py:544-566)."""
    xz = vec[[0, 2]] / np.linalg.norm(vec[[0, 2]])
    ry = _rot_about_y(_acos_clipped(xz[-1]) * np.sign(xz[0]))
    v1 = ry[:3, :3] @ vec
    yz = v1[1:3] / np.linalg.norm(v1[1:3])
    rx = _rot_about_x(_acos_clipped(yz[-1]) * np.sign(yz[0]))
    return np.linalg.inv(rx @ ry).T


def bone_align_transforms(rest_pose, parents=JOINT_PARENTS):
    """Row S0: per-joint A_j (24,4,4) float32 and the child each joint aligns to (raycasters.py:548-591).
    Joints with no child or several children keep the identity."""
    rest = np.asarray(rest_pose).reshape(len(parents), 3)
    A = np.tile(np.eye(4, dtype=np.float32), (len(parents), 1, 1))
    child_of = []
    for j, kids in enumerate(joint_children(parents)):
        if len(kids) != 1:
            child_of.append(j)
            continue
        c = kids[0]
        d = rest[c] - rest[j]
        rot = bone_to_z_rotation(d)
        shift = -0.5 * np.linalg.norm(d[None], axis=-1)[..., None] * np.array([[0., 0., 1.]], dtype=np.float32)
        rot[:3, -1] = shift[0]
        A[j] = rot.astype(np.float32)
        child_of.append(c)
    return A, np.array(child_of)


def bounding_cylinder(kps, ext_scale=0.001, extend_mm=250, top_expand_ratio=1.0, bot_expand_ratio=0.25, head="-y"):
    """(…,5) = [root_x, root_z, radius, top, bottom] around the keypoints (skeleton_utils.py:568-618)."""
    if head.endswith("z"):
        g, h = [0, 1], 2
    elif head.endswith("y"):
        g, h = [0, 2], 1
    else:
        raise NotImplementedError(f"head orientation {head}")
    flip = -1 if head.startswith("-") else 1
    kps = np.asarray(kps)
    root = kps[..., 0, :]
    if kps.ndim == 2:
        dist = np.linalg.norm(kps[:, g] - root[g], axis=-1)
    else:
        dist = np.linalg.norm(kps[..., g] - root[:, None, g], axis=-1)
    ext = extend_mm * ext_scale
    radius = dist.max(-1) + ext
    hi = (flip * kps[..., h]).max(-1)
    lo = (flip * kps[..., h]).min(-1)
    top = flip * (hi + ext * top_expand_ratio)
    bot = flip * (lo - ext * bot_expand_ratio)
    return np.stack([root[..., g[0]], root[..., g[1]], radius, top, bot], axis=-1)


def skeleton_profile(rest_pose, parents=JOINT_PARENTS, names=JOINT_NAMES):
    """Widths, bone lengths and body-part index sets of a rest pose (skeleton_utils.py:1474-1568)."""
    rest = np.asarray(rest_pose)
    if rest.ndim == 2:
        rest = rest[None]
    prof = {}
    for w in ("shoulder", "hip", "collar", "knee"):
        idx = [i for i, n in enumerate(names) if w in n]
        prof[f"{w}_width"] = np.linalg.norm(rest[:, idx[0]] - rest[:, idx[1]], axis=-1)
    kids = joint_children(parents)
    lens, lens_child = [], []
    for r in rest:
        lens.append([float(((r[j] - r[parents[j]]) ** 2).sum() ** 0.5) for j in range(1, len(parents))])
        lc = []
        for j, c in enumerate(kids):
            if len(c) < 1:
                lc.append(-1.0)
                continue
            cs = c[1:] if j == 0 else c
            lc.append((((r[j:j + 1] - r[cs]) ** 2).sum(-1) ** 0.5).mean())
        lens_child.append(lc)
    prof["bone_lens"] = np.concatenate([np.zeros((len(rest), 1)), np.array(lens)], axis=-1)
    prof["bone_lens_to_child"] = np.array(lens_child)
    groups = {"head": ["head"], "torso": ["shoulder", "spine", "collar", "neck", "pelvis"],
              "arm": ["elbow", "wrist", "hand"], "leg": ["hip", "knee", "ankle", "foot"]}
    for g, keys in groups.items():
        prof[f"{g}_idxs"] = np.array([i for i, n in enumerate(names) if any(k in n for k in keys)])
    return prof


def initial_axis_scale(profile, base_scale=0.4):
    """Row S1: initial per-bone half extents (24,3) float32 (misc.py:675-724).
    x/y come from rest-pose widths (the reference reads knee_width for the arms too), z = 0.8 * bone-to-child
    length; end effectors take the longest bone; the head gets 1.1x of it."""
    import torch
    lens_child = profile["bone_lens_to_child"][0]
    shoulder = float(profile["shoulder_width"][0])
    knee = float(profile["knee_width"][0])
    collar = knee                                   # misc.py:692 reads 'knee_width' for collar_width
    x = torch.ones(N_JOINTS) * base_scale
    y = torch.ones(N_JOINTS) * base_scale
    x[profile["leg_idxs"]] = knee * 0.5
    y[profile["leg_idxs"]] = knee * 0.5
    x[profile["torso_idxs"]] = shoulder * 0.70
    y[profile["torso_idxs"]] = shoulder * 0.70
    x[profile["head_idxs"]] = shoulder * 0.60
    y[profile["head_idxs"]] = shoulder * 0.60
    x[profile["arm_idxs"]] = collar * 0.60
    y[profile["arm_idxs"]] = collar * 0.60
    z = torch.tensor(lens_child.copy().astype(np.float32)) * 0.8
    z[z < 0] = z.max()
    z[profile["head_idxs"]] = z.max() * 1.1
    return torch.stack([x, y, z], dim=-1)
