"""Shared helpers for the parity tests: fixture loading and parameter rebuilding."""
import os

import numpy as np
import torch

import danbo_b200  # noqa: F401
from danbo_b200 import skeleton as sk, synthetic as syn

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_fixture(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)
    out = {}
    for k in z.files:
        v = z[k]
        out[k] = torch.from_numpy(v) if v.dtype.kind in "fiub" and v.ndim > 0 else v.item() if v.ndim == 0 else v
    return out


def params_for(fx):
    return syn.synthetic_params(int(fx["weight_seed"]))


def align_A():
    A, _ = sk.bone_align_transforms(syn.rest_pose())
    return torch.from_numpy(A)


def pose_tensors(fx):
    return (fx["pose_skts"][None] if fx["pose_skts"].dim() == 3 else fx["pose_skts"],
            fx["pose_bones"][None] if fx["pose_bones"].dim() == 2 else fx["pose_bones"],
            fx["pose_cyl"][None] if fx["pose_cyl"].dim() == 1 else fx["pose_cyl"])


def mask_mismatch_report(x, got_invalid, want_invalid, ulps=16):
    """Bone-visibility masks are a bit-exact target except where |x| sits within a few ulp of the box face.
    Returns (#mismatches, #mismatches explained by the margin)."""
    bad = got_invalid != want_invalid
    n_bad = int(bad.sum())
    if n_bad == 0:
        return 0, 0
    margin = (x.abs() - 1.0).abs().min(-1).values            # distance of the closest coordinate to a face
    near = margin <= ulps * 1.1920929e-07
    return n_bad, int((bad & near).sum())
