import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, ROOT + "/oracle", ROOT + "/tests"): sys.path.insert(0, p)
import torch, numpy as np
import danbo_oracle as orc
from util import *
import danbo_b200
K = danbo_b200.kernels
DEV = "cuda"
fx = load_fixture("render_fast")
caster, args, P = make_caster("danbo_fast")
Pc = params_for(fx)
skts, bones, _ = pose_tensors(fx)
rb = fx["ray_batch"].to(DEV)
N, S, S_f = rb.shape[0], 32, 16
consts, packed = caster._consts(), caster._packed_mlp()
vol = fx["st.vol.0"].to(DEV).contiguous()
z, mask, act = K.sample_mask(rb, S, skts.to(DEV).contiguous(), N, consts, z_in=fx["st.z.0"].to(DEV), append_empty=1)
xt, row_ray, _, hbar = K.field_agg(rb, S, z, mask, act, skts.to(DEV).contiguous(), vol, N, consts, want_hbar=True)
cams = fx["cams"].reshape(-1).to(DEV).to(torch.int32)
rbias = K.ray_bias(rb, cams, caster._codes_with_mean(), packed)
raw = torch.full((N * S + N, 4), float("nan"), device=DEV)
K.mlp_forward(xt, packed, rbias, act, row_ray, raw)
torch.cuda.synchronize()
n_act = int(act.count.item()); ids = act.ids[:n_act].long()
got = raw[ids].cpu()
X = decode_xtiles(xt, n_act).cpu()
view = orc.view_inputs(fx["ray_batch"][:, 3:6], fx["cams"], Pc, training=False)
want_bias = view @ Pc["views_linears.0.weight"][:, 256:].t() + Pc["views_linears.0.bias"]
vb = want_bias[row_ray[:n_act].cpu().long()]
emu = mlp_bf16_reference(X, vb, Pc)
x32 = orc.pe_embed(hbar[:n_act, :15].cpu(), 6)
ref = orc.field_mlp(x32, view[row_ray[:n_act].cpu().long()], Pc)
for c in range(4):
    e1 = (got[:, c] - emu[:, c]).abs(); e2 = (got[:, c] - ref[:, c]).abs(); e3 = (emu[:, c] - ref[:, c]).abs()
    print(f"ch{c}: scale {emu[:,c].abs().max():.3f} | got-emu mean {e1.mean():.2e} p99 {e1.quantile(.99):.2e} max {e1.max():.2e} | got-ref mean {e2.mean():.2e} max {e2.max():.2e} | emu-ref mean {e3.mean():.2e} max {e3.max():.2e}")
print("n_act", n_act, "rows")
# ---- resample debug
raw0 = torch.cat([fx["st.raw.0"].reshape(N * S, 4), torch.zeros(N, 4)], 0).to(DEV).contiguous()
ones = torch.ones(N, S, dtype=torch.int32, device=DEV)
out = K.composite_resample(rb, S, S_f, raw0, ones, fx["st.z.0"].to(DEV), want_inds=True)
z_all, zs, order, inds = orc.importance_sample(fx["st.z.0"], out["weights"].cpu(), S_f)
gi = out["inds"].cpu().long()
print("inds mismatch frac", (gi != inds).float().mean().item())
bad = (gi != inds).nonzero()[:10]
print("first mismatches (ray,j,got,want):", [(int(r), int(j), int(gi[r, j]), int(inds[r, j])) for r, j in bad])
r = int(bad[0, 0])
w = out["weights"].cpu()[r]
dw = 0.5 * (torch.maximum(w[:-2], w[1:-1]) + torch.maximum(w[1:-1], w[2:])) + 0.01 + 1e-5
cdf = torch.cat([torch.zeros(1), torch.cumsum(dw / dw.sum(), -1)])
print("cdf", cdf)
print("got inds", gi[r], "want", inds[r])
print("zs got", out["z_samples"].cpu()[r], "want", zs[r])
print("zs max err", (out["z_samples"].cpu() - zs).abs().max().item())
