"""danbo-pytorch_b200 — B200-native (sm_100a) implementation of DANBO's per-sample body-field hot path.

The directory name carries a hyphen, so import it through the alias module at the repo root:

    import danbo_b200                       # == importlib.import_module("danbo-pytorch_b200")

Host side: PyTorch for device memory, streams and torch.distributed.  Compute: hand-written CUDA kernels in
csrc/, reached only through the C ABI declared in include/danbo_b200.h (ctypes, raw pointers + sizes).
There is no CPU fallback: calling any kernel entry without the built library raises.
"""
from . import skeleton, synthetic, params, config  # noqa: F401
from . import _lib, build, kernels, networks, raycaster, anerf, parallel, training, render, feed, pose_opt, mesh, dropin  # noqa: F401
from .raycaster import RayCaster, GraphCaster, create_raycaster  # noqa: F401
from .config import make_args  # noqa: F401
from .dropin import install, uninstall  # noqa: F401

__all__ = ["skeleton", "synthetic", "params", "config", "kernels", "networks", "raycaster", "RayCaster",
           "GraphCaster", "create_raycaster", "make_args", "build"]
