import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, ROOT + "/oracle", ROOT + "/tests"): sys.path.insert(0, p)
import torch
from util import *
import danbo_b200
K = danbo_b200.kernels
DEV = "cuda"
fx = load_fixture("render_fast")
caster, args, P = make_caster(preset_of(fx))
skts, bones, _ = pose_tensors(fx)
rb = fx["ray_batch"].to(DEV)
N, S = rb.shape[0], int(fx["N_samples"])
consts, packed = caster._consts(), caster._packed_mlp()
vol = fx["st.vol.0"].to(DEV).contiguous()
z, mask, act = K.sample_mask(rb, S, skts.to(DEV).contiguous(), N, consts, z_in=fx["st.z.0"].to(DEV), append_empty=1)
fo = K.field_agg(rb, S, z, mask, act, skts.to(DEV).contiguous(), vol, N, consts, want_hbar=True)
cams = fx["cams"].reshape(-1).to(DEV).to(torch.int32)
rbias = K.ray_bias(rb, cams, caster._codes_with_mean(), packed)
cnt = int(act.count.item())
print("N", N, "S", S, "count", cnt, "row_ray numel", fo.row_ray.numel(), "rbias", tuple(rbias.shape), rbias.data_ptr() % 256, rbias.is_contiguous())
rr = fo.row_ray[:cnt]
print("row_ray min/max", int(rr.min()), int(rr.max()), "xt bytes", fo.xtiles.numel(), "need", (cnt + 127) // 128 * K.X_TILE_BYTES)
print("ids max", int(act.ids[:cnt].max()), "cap", N * S + N)
