"""Profiling aid: clock64 timeline of one CTA pair of the A-NeRF MLP kernel (issuer vs epilogue, per layer / pass)."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, ROOT + "/oracle", ROOT + "/tests"): sys.path.insert(0, p)
import torch
import danbo_b200 as db
from danbo_b200 import synthetic as syn, skeleton as sk, params, kernels as K
DEV = torch.device("cuda", 0)
args = db.make_args("anerf_base", no_reload=True)
attrs = {"skel_type": sk.SMPLSkeleton, "near": syn.NEAR, "far": syn.FAR, "n_views": 8, "rest_pose": syn.rest_pose()}
_, kw_test, *_ = db.create_raycaster(args, attrs, device=DEV)
caster = kw_test["ray_caster"]
caster.network.load_state_dict(syn.synth_state_dict(params.anerf_param_shapes(), 0), strict=False)
packed = caster._packed_mlp()
rows = 74 * 256 * 8
tiles = rows // 128
xd = torch.zeros(tiles * K.ANERF_XD_TILE_BYTES, dtype=torch.uint8, device=DEV)
xv = torch.zeros(tiles * K.ANERF_XV_TILE_BYTES, dtype=torch.uint8, device=DEV)
cb = torch.zeros(rows // 96 + 1, 224, device=DEV)
out = torch.empty(rows, 4, device=DEV)
for _ in range(2): K.anerf_mlp(xd, xv, packed, cb, rows, 96, out)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); K.anerf_mlp(xd, xv, packed, cb, rows, 96, out); e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
print(f"{rows} rows in {ms:.3f} ms -> {rows*4536000/ms/1e9:.1f} TFLOP/s; per tile pair {ms*1e3/8:.1f} us")
trace = torch.zeros(320, dtype=torch.int64, device=DEV)
K.anerf_mlp(xd, xv, packed, cb, rows, 96, out, trace=trace)
torch.cuda.synchronize()
t = trace.cpu()[:160].reshape(2, 2, 20, 2)
x = trace.cpu()[160:].reshape(2, 20, 4)
t0 = int(t[0, 0, 0, 0])
for it in range(2):
    print(f"--- tile iter {it}")
    for L in range(10):
        for h in range(2):
            if L == 9 and h == 1: continue
            m0, m1 = int(t[it, 0, L * 2 + h, 0]) - t0, int(t[it, 0, L * 2 + h, 1]) - t0
            p0, p1 = int(t[it, 1, L * 2 + h, 0]) - t0, int(t[it, 1, L * 2 + h, 1]) - t0
            d = [int(v) - t0 for v in x[it, L * 2 + h]]
            print(f"L{L} pass{h}: mma issue {m0:7d}..{m1:7d} ({m1-m0:5d}) | epilogue {p0:7d}..{p1:7d} ({p1-p0:5d}) acc@{d[0]:7d} ld0+{d[1]-d[0]:5d} round0+{d[2]-d[1]:5d} round1+{d[3]-d[2]:5d} tail+{p1-d[3]:5d}")
