"""Index arithmetic of the tensor-core aggregation net (csrc/field_mma.cu) checked on the CPU.

The kernel cannot run here, so this test mirrors its data movement lane by lane - the shared-memory feature tiles and
their 32-bit fragment reads, the packed B-fragment table (`pack_frags_kernel`'s index decode), the re-use of layer-0
accumulator fragments as layer-1 A fragments, the quad reduction of layer 2 - around an emulation of
`mma.sync.m16n8k16.row.col.f32.bf16.bf16.f32` built from the PTX ISA's fragment layouts:

    A (16x16): a0,a1 (row g, cols 2t,2t+1)  a2,a3 (row g+8, same cols)  a4,a5 (row g, cols 2t+8,2t+9)  a6,a7 (row g+8, ...)
    B (16x8):  b0,b1 (k 2t,2t+1, n g)       b2,b3 (k 2t+8,2t+9, n g)
    C (16x8):  c0,c1 (row g, cols 2t,2t+1)  c2,c3 (row g+8, same cols)            g = lane / 4, t = lane % 4

and compares the logits with the fp32 oracle (`danbo_oracle.agg_net`) on random features: the 3-term split-bf16
products must stay within 1e-5 of the logits' scale.  It also checks that the kernel's tree-neighbour masks are the
skeleton's.  What this cannot catch is a wrong memory of the PTX layouts themselves; the GPU test
(tests/test_gpu_zzz_pair_logits_mma.py) is the final word."""
import os
import re
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import danbo_b200                                    # noqa: E402,F401
from danbo_b200 import skeleton as sk, synthetic as syn, params  # noqa: E402
import danbo_oracle as orc                           # noqa: E402

KHP = 24                                             # kHP in field_mma.cu
bf = lambda v: v.to(torch.bfloat16).to(torch.float32)


def split(v):
    hi = bf(v)
    return hi, bf(v - hi)


def mma(acc, a, b):
    """acc [32 lanes][4], a [32][4 regs][2 halves], b [32][2][2] -> acc + A @ B in the PTX fragment layout."""
    A = torch.zeros(16, 16)
    B = torch.zeros(16, 8)
    for lane in range(32):
        g, t = lane >> 2, lane & 3
        for reg in range(4):
            row = g + 8 * (reg & 1)
            col = 2 * t + 8 * (reg >> 1)
            A[row, col], A[row, col + 1] = a[lane, reg, 0], a[lane, reg, 1]
        for reg in range(2):
            k = 2 * t + 8 * reg
            B[k, g], B[k + 1, g] = b[lane, reg, 0], b[lane, reg, 1]
    D = A.double() @ B.double()
    out = acc.clone()
    for lane in range(32):
        g, t = lane >> 2, lane & 3
        for e in range(4):
            out[lane, e] += float(D[g + 8 * (e >> 1), 2 * t + (e & 1)])
    return out


def pack_frags(w0, w1):
    """Mirror of pack_frags_kernel: tables [bone][hl][nt][lane][2 regs][2 halves] and [bone][hl][s][nt][lane][2][2]."""
    f0 = torch.zeros(24, 2, 4, 32, 2, 2)
    f1 = torch.zeros(24, 2, 2, 4, 32, 2, 2)
    n0, n1 = 24 * 2 * 4 * 32, 24 * 2 * 2 * 4 * 32
    for idx in range(n0 + n1):
        first = idx < n0
        rest = idx if first else idx - n0
        lane = rest & 31; rest >>= 5
        nt = rest & 3; rest >>= 2
        s = 0
        if not first:
            s = rest & 1; rest >>= 1
        hl = rest & 1; rest >>= 1
        bone = rest
        g, t = lane >> 2, lane & 3
        o = nt * 8 + g
        v = []
        for e in range(4):
            i = s * 16 + 2 * t + (e & 1) + (e >> 1) * 8
            if first:
                v.append(float(w0[bone, i, o]) if i < 15 else 0.0)
            else:
                v.append(float(w1[bone, i, o]))
        parts = split(torch.tensor(v))[hl]
        dst = f0[bone, hl, nt, lane] if first else f1[bone, hl, s, nt, lane]
        dst[0, 0], dst[0, 1], dst[1, 0], dst[1, 1] = parts[0], parts[1], parts[2], parts[3]
    return f0, f1


def kernel_chunk(j, h_all, P, f0, f1, nbr_mask):
    """One warp's 32-pair chunk of bone j, as pair_logits_mma_kernel computes it.  h_all (32 pairs, 24, 15)."""
    pre = "prob_linears.layers."
    adjw, adjm = P[pre + "0.adj_w"].reshape(24, 24), P[pre + "0.adj"].reshape(24, 24)
    b0, b1 = P[pre + "0.bias"], P[pre + "1.bias"].reshape(24, 32)
    w2, b2 = P[pre + "2.weight"].reshape(24, 32, 1), P[pre + "2.bias"].reshape(24, 1)
    acc = torch.zeros(2, 4, 32, 4)                    # [mt][nt][lane][4]
    for k in range(24):
        if not (nbr_mask[j] >> k) & 1:
            continue
        adj = adjw[j, k] * adjm[j, k]
        tiles = torch.zeros(2, 32 * KHP)              # hi / lo, flat bf16 elements, row pitch kHP
        for lane in range(32):
            row = torch.cat([h_all[lane, k] * adj, torch.zeros(1)])
            hi, lo = split(row)
            tiles[0, lane * KHP:lane * KHP + 16] = hi
            tiles[1, lane * KHP:lane * KHP + 16] = lo
        a = torch.zeros(2, 2, 32, 4, 2)               # [hl][mt][lane][reg][half]
        for hl in range(2):
            words = tiles[hl].reshape(-1, 2)          # 32-bit words: two consecutive bf16
            for mt in range(2):
                for lane in range(32):
                    g, t = lane >> 2, lane & 3
                    r0 = mt * 16 + g
                    w00, w10 = (r0 * KHP) // 2 + t, ((r0 + 8) * KHP) // 2 + t
                    a[hl, mt, lane, 0], a[hl, mt, lane, 1] = words[w00], words[w10]
                    a[hl, mt, lane, 2], a[hl, mt, lane, 3] = words[w00 + 4], words[w10 + 4]
        for nt in range(4):
            b_hi, b_lo = f0[k, 0, nt], f0[k, 1, nt]
            for mt in range(2):
                acc[mt, nt] = mma(acc[mt, nt], a[0, mt], b_hi)
                acc[mt, nt] = mma(acc[mt, nt], a[1, mt], b_hi)
                acc[mt, nt] = mma(acc[mt, nt], a[0, mt], b_lo)
    # bias + relu -> layer-1 A fragments
    m = torch.zeros(2, 2, 2, 32, 4, 2)                # [hl][mt][s][lane][reg][half]
    for mt in range(2):
        for nt in range(4):
            s, half = nt >> 1, nt & 1
            for lane in range(32):
                t = lane & 3
                c = [nt * 8 + 2 * t, nt * 8 + 2 * t + 1]
                top = torch.relu(torch.stack([acc[mt, nt, lane, 0] + b0[c[0]], acc[mt, nt, lane, 1] + b0[c[1]]]))
                bot = torch.relu(torch.stack([acc[mt, nt, lane, 2] + b0[c[0]], acc[mt, nt, lane, 3] + b0[c[1]]]))
                for hl, parts in enumerate(zip(split(top), split(bot))):
                    m[hl, mt, s, lane, half * 2 + 0] = parts[0]
                    m[hl, mt, s, lane, half * 2 + 1] = parts[1]
    acc1 = torch.zeros(2, 4, 32, 4)
    for mt in range(2):
        for nt in range(4):
            for lane in range(32):
                t = lane & 3
                c0, c1 = b1[j, nt * 8 + 2 * t], b1[j, nt * 8 + 2 * t + 1]
                acc1[mt, nt, lane] = torch.stack([c0, c1, c0, c1])
    for s in range(2):
        for nt in range(4):
            b_hi, b_lo = f1[j, 0, s, nt], f1[j, 1, s, nt]
            for mt in range(2):
                acc1[mt, nt] = mma(acc1[mt, nt], m[0, mt, s], b_hi)
                acc1[mt, nt] = mma(acc1[mt, nt], m[1, mt, s], b_hi)
                acc1[mt, nt] = mma(acc1[mt, nt], m[0, mt, s], b_lo)
    # layer 2 + quad reduction
    part = torch.zeros(2, 2, 32)
    for nt in range(4):
        for lane in range(32):
            t = lane & 3
            w20, w21 = w2[j, nt * 8 + 2 * t, 0], w2[j, nt * 8 + 2 * t + 1, 0]
            for mt in range(2):
                r = torch.relu(acc1[mt, nt, lane])
                part[mt, 0, lane] += r[0] * w20 + r[1] * w21
                part[mt, 1, lane] += r[2] * w20 + r[3] * w21
    outs = torch.zeros(32)
    for mt in range(2):
        for hf in range(2):
            v = part[mt, hf].clone()
            v = v + v[torch.arange(32) ^ 1]
            v = v + v[torch.arange(32) ^ 2]
            for lane in range(32):
                if lane & 3 == 0:
                    outs[mt * 16 + hf * 8 + (lane >> 2)] = v[lane] + b2[j, 0]
    return outs


def header_neighbour_masks():
    src = open(os.path.join(ROOT, "danbo-pytorch_b200", "csrc", "field_common.cuh")).read()
    body = src[src.index("kNbrMask[DANBO_J] = {"):]
    body = body[:body.index("};")]
    return [int(v, 16) for v in re.findall(r"0x([0-9A-Fa-f]+)u", body)]


def test_neighbour_masks_are_the_skeleton_tree():
    masks = header_neighbour_masks()
    assert len(masks) == 24
    par = sk.JOINT_PARENTS
    for j in range(24):
        want = {j, int(par[j])} | {c for c in range(24) if par[c] == j and c != j}
        assert {k for k in range(24) if (masks[j] >> k) & 1} == want, j
    adj = torch.tensor(sk.skeleton_adjacency()) + torch.eye(24)
    assert all(((masks[j] >> k) & 1) == int(adj[j, k] > 0) for j in range(24) for k in range(24))


def test_split_bf16_mma_chunk_matches_fp32_aggregation_net():
    P = {k: torch.as_tensor(v) for k, v in syn.synthetic_params(0).items()}          # includes the adjacency buffers
    pre = "prob_linears.layers."
    f0, f1 = pack_frags(P[pre + "0.lin.weight"], P[pre + "1.weight"])
    masks = header_neighbour_masks()
    torch.manual_seed(3)
    h_all = torch.randn(32, 24, 15) * 0.7
    want = orc.agg_net(h_all, P)                       # (32, 24), fp32
    scale = float(want.abs().max())
    for j in (0, 9, 17, 23):                           # root (4 neighbours), spine3 (5), an arm joint (3), a leaf (2)
        got = kernel_chunk(j, h_all, P, f0, f1, masks)
        err = float((got - want[:, j]).abs().max())
        print(f"[mma layout] bone {j}: max err {err:.3e} of scale {scale:.3e}")
        assert err <= 1e-5 * scale, (j, err, scale)
