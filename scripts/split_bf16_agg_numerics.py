"""CPU experiment for DESIGN §8.3: accuracy of the aggregation net's blend logits when its two GEMM-shaped layers run as
split-bf16 tensor-core products (x = hi + lo with hi = bf16(x), lo = bf16(x - hi); products hi*hi + hi*lo + lo*hi
accumulated in fp32), on the reference-generated fixture samples.  No GPU needed.
    python scripts/split_bf16_agg_numerics.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, ROOT + "/oracle", ROOT + "/tests"): sys.path.insert(0, p)
import torch
import torch.nn.functional as F
import danbo_oracle as orc
from util import load_fixture, params_for


def split(x):
    hi = x.to(torch.bfloat16).float()
    lo = (x - hi).to(torch.bfloat16).float()
    return hi, lo


def mm_split(a, w, terms):
    """a (...,k) x w (k,n) per joint via einsum, with `terms` of the split product."""
    ah, al = split(a)
    wh, wl = split(w)
    out = torch.einsum("bkl,klj->bkj", ah, wh)
    if terms >= 2:
        out = out + torch.einsum("bkl,klj->bkj", ah, wl)
    if terms >= 3:
        out = out + torch.einsum("bkl,klj->bkj", al, wh)
    return out


def agg_net_split(h, P, terms, prefix="prob_linears"):
    o = mm_split(h, P[f"{prefix}.layers.0.lin.weight"], terms)
    adj = P[f"{prefix}.layers.0.adj_w"] * P[f"{prefix}.layers.0.adj"]
    o = F.relu(torch.matmul(adj, o) + P[f"{prefix}.layers.0.bias"])          # the 24x24 mix stays fp32 (70 FMAs per output)
    o = F.relu(mm_split(o, P[f"{prefix}.layers.1.weight"], terms) + P[f"{prefix}.layers.1.bias"])
    a = torch.einsum("bkl,klj->bkj", o, P[f"{prefix}.layers.2.weight"]) + P[f"{prefix}.layers.2.bias"]   # 32 -> 1: fp32
    return a[..., 0]


for name in ("render_fast", "render_base"):
    fx = load_fixture(name)
    P = params_for(fx)
    h = fx["st.h.0"].reshape(-1, 24, 15)
    # every (sample, bone) logit of the kept rays, visible or not (the window makes far bones' features tiny, not zero)
    valid = torch.ones(h.shape[0], 24)
    ref = orc.agg_net(h, P)
    scale = float(ref.abs().max())
    for terms in (1, 2, 3):
        got = agg_net_split(h, P, terms)
        err = ((got - ref) * valid).abs()
        p_err = ((torch.sigmoid(got) - torch.sigmoid(ref)) * valid).abs()
        print(f"{name}: {terms}-term split bf16: logit err max {float(err.max()):.2e} mean {float(err.mean()):.2e} "
              f"(scale {scale:.2e}, rel {float(err.max()) / scale:.1e}); blend weight err max {float(p_err.max()):.2e}")
