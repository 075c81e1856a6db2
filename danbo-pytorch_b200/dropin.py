"""Drop-in hook: the reference's own scripts (run_nerf.py, run_render.py) build this repo's ray caster, unchanged.

The reference has one seam for the hot path: `create_raycaster` (core/raycasters.py:17-143), imported by name at
run_nerf.py:20 / run_render.py and called at run_nerf.py:606.  `install()` rebinds that name - in `core.raycasters` and in
every already-imported module that copied it - to `danbo_b200.create_raycaster`; `uninstall()` restores it.  Launcher:

    python -m danbo_b200.run  /path/to/DANBO-pytorch/run_nerf.py  --config configs/h36m_zju/danbo_fast.txt ...

runs the script as `__main__` with the hook installed (run.py).  Nothing in the reference tree is edited.
"""
import importlib
import sys

_STATE = {"orig": None}


def install(reference_root=None, anerf=True):
    """-> the reference's original create_raycaster.  `reference_root`: the DANBO-pytorch checkout (added to sys.path
    when `core` is not importable yet).  anerf=False keeps the reference's own path for nerf_type='nerf'."""
    from .raycaster import create_raycaster as ours
    if reference_root is not None and reference_root not in sys.path:
        sys.path.insert(0, reference_root)
    rc = importlib.import_module("core.raycasters")
    if _STATE["orig"] is None:
        _STATE["orig"] = rc.create_raycaster
    orig = _STATE["orig"]

    def create_raycaster(args, data_attrs, device=None):
        if not anerf and getattr(args, "nerf_type", None) == "nerf":
            return orig(args, data_attrs, device=device)
        return ours(args, data_attrs, device=device)        # unsupported flags raise NotImplementedError: no fallback

    create_raycaster.__danbo_b200__ = True
    rc.create_raycaster = create_raycaster
    for mod in list(sys.modules.values()):                  # `from core.raycasters import create_raycaster` copies
        # only real module globals: a getattr would wake lazy namespaces such as torch.classes
        if mod is not None and mod is not rc and getattr(mod, "__dict__", {}).get("create_raycaster") is orig:
            setattr(mod, "create_raycaster", create_raycaster)
    return orig


def uninstall():
    orig = _STATE["orig"]
    if orig is None:
        return
    rc = importlib.import_module("core.raycasters")
    rc.create_raycaster = orig
    for mod in list(sys.modules.values()):
        f = getattr(mod, "__dict__", {}).get("create_raycaster") if mod is not None else None
        if f is not None and getattr(f, "__dict__", {}).get("__danbo_b200__", False):
            setattr(mod, "create_raycaster", orig)
    _STATE["orig"] = None
