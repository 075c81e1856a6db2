"""python -m danbo_b200.run <reference script> [its arguments...]: runs run_nerf.py / run_render.py of an unmodified
DANBO-pytorch checkout with this repo's ray caster behind `create_raycaster` (see dropin.py)."""
import os
import runpy
import sys


def main(argv=None):
    argv = sys.argv[1:] if argv is None else argv
    if not argv:
        raise SystemExit("usage: python -m danbo_b200.run /path/to/DANBO-pytorch/run_nerf.py [args...]")
    script = os.path.abspath(argv[0])
    root = os.path.dirname(script)
    from . import dropin
    dropin.install(reference_root=root)
    sys.argv = [script] + list(argv[1:])
    os.chdir(root)                                          # the reference's configs use paths relative to its root
    runpy.run_path(script, run_name="__main__")


if __name__ == "__main__":
    main()
