"""CPU checks of the drop-in boundary: the C-ABI library builds, loads, and exports every symbol the header declares."""
import ctypes
import os
import re

import danbo_b200
from danbo_b200 import _lib, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "danbo_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\bint\s+(danbo_\w+)\s*\(", text)))


def test_header_and_library_agree():
    build.build()
    lib = ctypes.CDLL(_lib.path())
    syms = header_symbols()
    assert len(syms) >= 10
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/danbo_b200.h but not exported"
    assert sorted(_lib.EXPORTS) == syms, "ctypes signatures must cover exactly the header's entry points"
    assert lib.danbo_version() == 1


def test_kernels_refuse_cpu_tensors():
    import pytest
    import torch
    with pytest.raises(RuntimeError):
        danbo_b200.kernels.nearfar(torch.zeros(4, 8), torch.zeros(1, 5), torch.zeros(1, 24, 4, 4), 4,
                                   torch.zeros(24, 4, 4), torch.ones(24, 3))


def test_unsupported_flags_raise():
    import pytest
    for kw in ({"agg_type": "relu"}, {"agg_type": "sum"}, {"gnn_backbone": "PNBGNN"}, {"netwidth": 128},
               {"nerf_type": "nerf"}, {"single_net": False}):
        with pytest.raises(NotImplementedError):
            danbo_b200.raycaster.check_args(danbo_b200.make_args("danbo_base", **kw))
    # both aggregation types of danbo.py:431-435 that a config can reach pass the check; softmax needs mask_vol_prob
    for agg in ("sigmoid", "softmax"):
        danbo_b200.raycaster.check_args(danbo_b200.make_args("danbo_base", agg_type=agg, lindisp=True))
    from danbo_b200.networks import DanboField
    with pytest.raises(NotImplementedError):
        DanboField(agg_type="softmax", mask_vol_prob=False)


def test_anerf_flag_subset():
    """nerf_type='nerf' (A-NeRF, BASELINE config #4): the preset passes the flag check, anything outside the shipped
    anerf_base.txt subset raises (no silent fallback)."""
    import pytest
    import danbo_b200 as db
    from danbo_b200 import anerf
    args = db.make_args("anerf_base", no_reload=True)
    anerf.check_anerf_args(args)
    for k, v in (("netwidth", 256), ("multires", 10), ("cutoff_viewdir", False), ("view_type", "world"),
                 ("bone_type", "axisang"), ("cutoff_mm", 400.0)):
        bad = db.make_args("anerf_base", no_reload=True, **{k: v})
        with pytest.raises(NotImplementedError):
            anerf.check_anerf_args(bad)
    # parameter inventory = the reference's state_dict (probe in oracle/gen_golden.py), minus the embedder constants
    from danbo_b200 import params
    shapes = params.anerf_param_shapes()
    assert shapes["pts_linears.0.weight"] == (448, 432) and shapes["pts_linears.5.weight"] == (448, 880)
    assert shapes["views_linears.0.weight"] == (224, 1224) and shapes["rgb_linear.weight"] == (3, 224)
    net = anerf.AnerfField()
    sd = net.state_dict()
    assert set(shapes) <= set(sd) and {"pe_fn.cutoff_dist", "pe_fn.tau", "dirs_pe_fn.cutoff_dist", "dirs_pe_fn.tau"} <= set(sd)
    for k, shp in shapes.items():
        assert tuple(sd[k].shape) == tuple(shp), k
