"""Profiling aid: clock64 timeline of one CTA of the fused MLP kernel (MMA issuer vs epilogue, per layer/half)."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, ROOT + "/oracle", ROOT + "/tests"): sys.path.insert(0, p)
import torch
from util import *
import danbo_b200
K = danbo_b200.kernels
DEV = "cuda"
caster, args, P = make_caster("danbo_fast")
packed = caster._packed_mlp()
rows = 148 * 128 * 8
n_tiles = rows // 128
xt = torch.zeros(n_tiles * K.X_TILE_BYTES, dtype=torch.uint8, device=DEV)
act = K.ActiveList(rows, DEV); act.ids.copy_(torch.arange(rows, device=DEV, dtype=torch.int32)); act.count.fill_(rows)
row_ray = torch.zeros(rows, dtype=torch.int32, device=DEV)
rbias = torch.zeros(1, 128, device=DEV)
out = torch.empty(rows, 4, device=DEV)
for _ in range(3): K.mlp_forward(xt, packed, rbias, act, row_ray, out)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); K.mlp_forward(xt, packed, rbias, act, row_ray, out); e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
print(f"{rows} rows in {ms:.3f} ms -> {rows*1354752/ms/1e9:.1f} TFLOP/s; per tile per SM {ms*1e3/8:.1f} us")
trace = torch.zeros(480, dtype=torch.int64, device=DEV)
e0.record(); K.mlp_forward(xt, packed, rbias, act, row_ray, out, trace=trace); e1.record()
torch.cuda.synchronize()
ms_tr = e0.elapsed_time(e1)
t = trace.cpu()[:320].reshape(4, 2, 20, 2)
x = trace.cpu()[320:480].reshape(4, 20, 2)
t0 = int(t[0, 0, 0, 0])
clk_per_tile = (int(t[3, 0, 0, 0]) - t0) / 3
print(f"traced launch {ms_tr:.3f} ms; {clk_per_tile:.0f} clk per tile -> SM clock ~ {clk_per_tile * 8 / (ms_tr * 1e3):.0f} MHz (if the launch is 8 equal tiles)")
for it in range(2):
    print(f"--- tile iter {it} (clocks relative to first MMA)")
    for L in range(10):
        for h in range(2):
            if L == 9 and h == 1: continue
            m0, m1 = int(t[it, 0, L * 2 + h, 0]) - t0, int(t[it, 0, L * 2 + h, 1]) - t0
            p0, p1 = int(t[it, 1, L * 2 + h, 0]) - t0, int(t[it, 1, L * 2 + h, 1]) - t0
            l1, s1 = int(x[it, L * 2 + h, 0]) - t0, int(x[it, L * 2 + h, 1]) - t0
            print(f"L{L} h{h}: mma issue {m0:7d}..{m1:7d} ({m1-m0:5d}) | epi acc_full@{p0:7d} done@{p1:7d} ({p1-p0:5d}) ld+{l1-p0:5d} math+{s1-l1:5d} st+{p1-s1:5d}")

