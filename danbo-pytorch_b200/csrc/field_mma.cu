// A1 on tensor cores (opt-in): the aggregation net of `pair_logits` as split-bf16 `mma.sync` products.
//
// Reference: DANBO.get_agg_logits core/networks/danbo.py:304-330 on MixGNN (gnn_backbone.py:225-274): per (sample, bone
// j) pair   mix = relu(sum_{k in tree(j)} adj[j][k] h_k W0[k] + b0),  l1 = relu(mix W1[j] + b1[j]),  a = l1 . w2[j] + b2[j].
//
// Why: the FFMA kernel (field.cu) is bound by the load/store unit, not by arithmetic - every weight is a warp-uniform
// load feeding 64 MACs, ~780 load wavefronts per 64-pair chunk (profiles/r1d_ncu_simt_summary.txt: LSU 74 % busy,
// FMA pipe 39 %).  Here a warp's 32 pairs are the M dimension of m16n8k16 MMAs (two M tiles), the weights are B
// fragments read as coalesced 8-byte loads from a table packed once per weight update (16 wavefronts per neighbour
// instead of 120), and the layer-0 accumulator fragments are re-used in place as the A fragments of layer 1 (the C
// layout of two adjacent n8 tiles is the A layout of one k16 step).
//
// Precision: every fp32 operand v is split into bf16 hi = rn(v), lo = rn(v - hi) and a product is accumulated in fp32 as
// a_hi b_hi + a_lo b_hi + a_hi b_lo (the lo.lo term is below 2^-16 of the product).  On the reference fixtures this
// keeps the logits within 3e-6 of their scale (scripts/split_bf16_agg_numerics.py), inside the 1e-4 `confd` tolerance;
// plain bf16 would be off by 1.7e-3.
//
// STATUS: written at the end of round 1 without GPU access; selected only when the caller passes a fragment table
// (consts[10] != NULL, `DANBO_PAIR_LOGITS=mma` in the Python layer).  The index arithmetic is checked on the CPU by a
// lane-level emulation of the MMA fragment layouts (tests/test_pair_logits_mma_layout.py).
#include "field_common.cuh"

namespace danbo {
namespace aggmma {

constexpr int kHP = 24;                              // bf16 pitch of a feature row in shared memory (48 B: conflict-free)
constexpr int kW0Words = DANBO_J * 2 * 4 * 32 * 2;   // [k][hi|lo][n tile][lane][2]
constexpr int kW1Words = DANBO_J * 2 * 2 * 4 * 32 * 2;   // [j][hi|lo][k step][n tile][lane][2]

__device__ __forceinline__ uint32_t pack_bf16(__nv_bfloat16 lo_half, __nv_bfloat16 hi_half) {
    __nv_bfloat162 v;
    v.x = lo_half; v.y = hi_half;                    // .x = low 16 bits = the element with the lower k / column index
    return *reinterpret_cast<uint32_t*>(&v);
}

__device__ __forceinline__ void split(float v, __nv_bfloat16& hi, __nv_bfloat16& lo) {
    hi = __float2bfloat16_rn(v);
    lo = __float2bfloat16_rn(v - __bfloat162float(hi));
}

__device__ __forceinline__ void mma16816(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

// B fragment of an m16n8k16 MMA for lane (g = lane / 4, t = lane % 4): reg 0 = {B[2t][g], B[2t+1][g]},
// reg 1 = {B[2t+8][g], B[2t+9][g]} with B[k][n] = W[i = 16 s + k][o = 8 nt + n].
__global__ void pack_frags_kernel(const float* __restrict__ w0 /* (24,15,32) */, const float* __restrict__ w1 /* (24,32,32) */,
                                  uint32_t* __restrict__ frags) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    const int n0 = kW0Words / 2, n1 = kW1Words / 2;
    if (idx >= n0 + n1) return;
    int rest = idx < n0 ? idx : idx - n0;
    const int lane = rest & 31; rest >>= 5;
    const int nt = rest & 3; rest >>= 2;
    int s = 0;
    if (idx >= n0) { s = rest & 1; rest >>= 1; }
    const int hl = rest & 1; rest >>= 1;
    const int bone = rest;
    const int g = lane >> 2, t = lane & 3, o = nt * 8 + g;
    float v[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const int i = s * 16 + 2 * t + (e & 1) + (e >> 1) * 8;
        if (idx < n0) v[e] = i < DANBO_FEAT ? w0[((size_t)bone * DANBO_FEAT + i) * DANBO_AGG_W + o] : 0.f;
        else v[e] = w1[((size_t)bone * DANBO_AGG_W + i) * DANBO_AGG_W + o];
    }
    __nv_bfloat16 h[4], l[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) split(v[e], h[e], l[e]);
    // the decode above is the inverse of the table index: W0 word pair ((k*2 + hl)*4 + nt)*32 + lane, W1 word pair
    // (((j*2 + hl)*2 + s)*4 + nt)*32 + lane after the W0 part
    uint32_t* dst = frags + (idx < n0 ? 2 * (size_t)idx : kW0Words + 2 * (size_t)(idx - n0));
    dst[0] = hl ? pack_bf16(l[0], l[1]) : pack_bf16(h[0], h[1]);
    dst[1] = hl ? pack_bf16(l[2], l[3]) : pack_bf16(h[2], h[3]);
}

// kBlocks = resident blocks per SM the register allocation is bounded for: the kernel is latency bound (dependent chain
// feature taps -> MMA -> mix -> MMA per 32-pair chunk; ncu at 12 warps/SM: tensor pipe 24 %, long-scoreboard stalls), so
// occupancy is what hides it.  3 -> 137 registers, 4 -> 128 (no spill), 5 -> 96 (20 bytes of spill).
template <bool kOnePose, int kBlocks>
__global__ void __launch_bounds__(128, kBlocks)
pair_logits_mma_kernel(const float* __restrict__ rays, int ray_stride, int S, const float* __restrict__ z,
                       const int* __restrict__ active_ids, const float* __restrict__ pose_skts,
                       const float* __restrict__ pose_vol, int rays_per_pose, int n_poses, FieldConsts fc, PairWork pw,
                       int pair_capacity, const uint32_t* __restrict__ frags, float* __restrict__ logits) {
    __shared__ __align__(16) float vol_s[kOnePose ? DANBO_J * DANBO_VOL : 4];
    __shared__ __align__(16) __nv_bfloat16 Hs[4][2][32 * kHP];          // per warp: hi / lo feature tiles [32 pairs][16]
    __shared__ float outs[4][32];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int g = lane >> 2, t = lane & 3;
    if (kOnePose) {
        const float4* src = reinterpret_cast<const float4*>(pose_vol);
        for (int i = threadIdx.x; i < DANBO_J * DANBO_VOL / 4; i += blockDim.x) reinterpret_cast<float4*>(vol_s)[i] = __ldg(src + i);
        __syncthreads();
    }
    const int* cnt = pw.count();
    int n_chunks = 0;
    for (int j = 0; j < DANBO_J; ++j) n_chunks += (cnt[j] + 31) >> 5;
    const int warps = (gridDim.x * blockDim.x) >> 5;
    const uint2* f0 = reinterpret_cast<const uint2*>(frags);
    const uint2* f1 = reinterpret_cast<const uint2*>(frags + kW0Words);
    float b0v[4][2];                                                     // layer-0 bias of this lane's columns (shared by all bones)
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) { b0v[nt][0] = __ldg(fc.agg_b0 + nt * 8 + 2 * t); b0v[nt][1] = __ldg(fc.agg_b0 + nt * 8 + 2 * t + 1); }
    for (int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; c < n_chunks; c += warps) {
        int j = 0, first = 0;
        for (;; ++j) { const int nc = (cnt[j] + 31) >> 5; if (c < first + nc) break; first += nc; }
        const int in_seg = (c - first) * 32 + lane;
        const int at = c * 32 + lane;                                    // segments are padded to 32 pairs
        const bool live = in_seg < cnt[j] && at < pair_capacity;
        float px = 0.f, py = 0.f, pz = 0.f;
        int id = 0, pose = 0;
        if (live) {
            id = active_ids[pw.pairs()[at]];
            const int n = id / S;
            const float* r = rays + (size_t)n * ray_stride;
            const float zz = z[id];
            px = __fadd_rn(r[0], __fmul_rn(r[3], zz));
            py = __fadd_rn(r[1], __fmul_rn(r[4], zz));
            pz = __fadd_rn(r[2], __fmul_rn(r[5], zz));
            pose = n / rays_per_pose; if (pose >= n_poses) pose = n_poses - 1;
        }
        float acc[2][4][4];
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
            for (int nt = 0; nt < 4; ++nt)
#pragma unroll
                for (int e = 0; e < 4; ++e) acc[mt][nt][e] = 0.f;
        // ---- layer 0: one k16 step per tree neighbour (15 features + one zero column)
        for (uint32_t nb = kNbrMask[j]; nb;) {
            const int k = __ffs(nb) - 1; nb &= nb - 1;
            const float adj = __ldg(fc.agg_adjw + j * DANBO_J + k) * __ldg(fc.agg_adj + j * DANBO_J + k);
            float h[DANBO_FEAT];
            {
                float x0, x1, x2;
                bone_coords(pose_skts + ((size_t)pose * DANBO_J + k) * 16, fc.align + k * 16, fc.axis_scale + k * 3,
                            px, py, pz, x0, x1, x2);
                if (kOnePose) bone_features<true>(vol_s + k * DANBO_VOL, x0, x1, x2, h);
                else bone_features<false>(pose_vol + ((size_t)pose * DANBO_J + k) * DANBO_VOL, x0, x1, x2, h);
            }
            uint32_t wh[8], wl[8];                                       // this pair's row, 16 bf16 each
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                __nv_bfloat16 h0, l0, h1, l1;
                split(live ? h[2 * q] * adj : 0.f, h0, l0);
                if (2 * q + 1 < DANBO_FEAT) split(live ? h[2 * q + 1] * adj : 0.f, h1, l1);
                else { h1 = __float2bfloat16_rn(0.f); l1 = h1; }
                wh[q] = pack_bf16(h0, h1); wl[q] = pack_bf16(l0, l1);
            }
            uint4* rh = reinterpret_cast<uint4*>(Hs[wib][0] + lane * kHP);
            uint4* rl = reinterpret_cast<uint4*>(Hs[wib][1] + lane * kHP);
            rh[0] = make_uint4(wh[0], wh[1], wh[2], wh[3]); rh[1] = make_uint4(wh[4], wh[5], wh[6], wh[7]);
            rl[0] = make_uint4(wl[0], wl[1], wl[2], wl[3]); rl[1] = make_uint4(wl[4], wl[5], wl[6], wl[7]);
            __syncwarp();
            uint32_t a_hi[2][4], a_lo[2][4];
#pragma unroll
            for (int mt = 0; mt < 2; ++mt) {
                const int r0 = mt * 16 + g;
                const uint32_t* Th = reinterpret_cast<const uint32_t*>(Hs[wib][0]);
                const uint32_t* Tl = reinterpret_cast<const uint32_t*>(Hs[wib][1]);
                const int w00 = (r0 * kHP) / 2 + t, w10 = ((r0 + 8) * kHP) / 2 + t;       // 32-bit word = columns 2t, 2t+1
                a_hi[mt][0] = Th[w00]; a_hi[mt][1] = Th[w10]; a_hi[mt][2] = Th[w00 + 4]; a_hi[mt][3] = Th[w10 + 4];
                a_lo[mt][0] = Tl[w00]; a_lo[mt][1] = Tl[w10]; a_lo[mt][2] = Tl[w00 + 4]; a_lo[mt][3] = Tl[w10 + 4];
            }
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) {
                const uint2 bh = __ldg(f0 + ((k * 2 + 0) * 4 + nt) * 32 + lane);
                const uint2 bl = __ldg(f0 + ((k * 2 + 1) * 4 + nt) * 32 + lane);
                const uint32_t b_hi[2] = {bh.x, bh.y}, b_lo[2] = {bl.x, bl.y};
#pragma unroll
                for (int mt = 0; mt < 2; ++mt) {
                    mma16816(acc[mt][nt], a_hi[mt], b_hi);
                    mma16816(acc[mt][nt], a_lo[mt], b_hi);
                    mma16816(acc[mt][nt], a_hi[mt], b_lo);
                }
            }
            __syncwarp();                                                // the tiles are rewritten for the next neighbour
        }
        // ---- bias + relu; the accumulator fragments of n tiles (2s, 2s+1) are the A fragment of k step s of layer 1
        uint32_t m_hi[2][2][4], m_lo[2][2][4];
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) {
                const int s = nt >> 1, half = nt & 1;
                __nv_bfloat16 h0, l0, h1, l1;
                split(fmaxf(acc[mt][nt][0] + b0v[nt][0], 0.f), h0, l0); split(fmaxf(acc[mt][nt][1] + b0v[nt][1], 0.f), h1, l1);
                m_hi[mt][s][half * 2 + 0] = pack_bf16(h0, h1); m_lo[mt][s][half * 2 + 0] = pack_bf16(l0, l1);   // row g
                split(fmaxf(acc[mt][nt][2] + b0v[nt][0], 0.f), h0, l0); split(fmaxf(acc[mt][nt][3] + b0v[nt][1], 0.f), h1, l1);
                m_hi[mt][s][half * 2 + 1] = pack_bf16(h0, h1); m_lo[mt][s][half * 2 + 1] = pack_bf16(l0, l1);   // row g + 8
            }
        // ---- layer 1 (32 -> 32)
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) {
                const float c0 = __ldg(fc.agg_b1 + j * DANBO_AGG_W + nt * 8 + 2 * t), c1 = __ldg(fc.agg_b1 + j * DANBO_AGG_W + nt * 8 + 2 * t + 1);
                acc[mt][nt][0] = c0; acc[mt][nt][1] = c1; acc[mt][nt][2] = c0; acc[mt][nt][3] = c1;
            }
#pragma unroll
        for (int s = 0; s < 2; ++s)
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) {
                const uint2 bh = __ldg(f1 + (((j * 2 + 0) * 2 + s) * 4 + nt) * 32 + lane);
                const uint2 bl = __ldg(f1 + (((j * 2 + 1) * 2 + s) * 4 + nt) * 32 + lane);
                const uint32_t b_hi[2] = {bh.x, bh.y}, b_lo[2] = {bl.x, bl.y};
#pragma unroll
                for (int mt = 0; mt < 2; ++mt) {
                    mma16816(acc[mt][nt], m_hi[mt][s], b_hi);
                    mma16816(acc[mt][nt], m_lo[mt][s], b_hi);
                    mma16816(acc[mt][nt], m_hi[mt][s], b_lo);
                }
            }
        // ---- layer 2 (32 -> 1) in fp32: this lane holds 8 of the 32 columns of rows g and g + 8 of both M tiles
        float part[2][2] = {{0.f, 0.f}, {0.f, 0.f}};
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
            const float w20 = __ldg(fc.agg_w2 + j * DANBO_AGG_W + nt * 8 + 2 * t), w21 = __ldg(fc.agg_w2 + j * DANBO_AGG_W + nt * 8 + 2 * t + 1);
#pragma unroll
            for (int mt = 0; mt < 2; ++mt) {
                part[mt][0] = fmaf(fmaxf(acc[mt][nt][0], 0.f), w20, fmaf(fmaxf(acc[mt][nt][1], 0.f), w21, part[mt][0]));
                part[mt][1] = fmaf(fmaxf(acc[mt][nt][2], 0.f), w20, fmaf(fmaxf(acc[mt][nt][3], 0.f), w21, part[mt][1]));
            }
        }
        const float b2 = __ldg(fc.agg_b2 + j);
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
            for (int hf = 0; hf < 2; ++hf) {
                float v = part[mt][hf];
                v += __shfl_xor_sync(0xffffffffu, v, 1);
                v += __shfl_xor_sync(0xffffffffu, v, 2);
                if (t == 0) outs[wib][mt * 16 + hf * 8 + g] = v + b2;
            }
        __syncwarp();
        if (live) logits[(size_t)id * DANBO_J + j] = outs[wib][lane];
        __syncwarp();
    }
}

}  // namespace aggmma

static int g_pair_logits_blocks = 4;     // resident blocks per SM of pair_logits_mma_kernel (3, 4 or 5)

int set_pair_logits_blocks(int b) {
    const int old = g_pair_logits_blocks;
    if (b >= 3 && b <= 5) g_pair_logits_blocks = b;
    return old;
}

int launch_pair_logits_mma(const float* rays, int ray_stride, int S, const float* z, const int* active_ids,
                           const float* pose_skts, const float* pose_vol, int rays_per_pose, int n_poses,
                           const FieldConsts& fc, PairWork pw, int pair_capacity, const void* frags, float* logits,
                           int num_sms, cudaStream_t st) {
    const int occ = g_pair_logits_blocks;
    int pblocks = (pair_capacity / 32 + 3) / 4;
    if (pblocks > num_sms * occ) pblocks = num_sms * occ;             // one resident wave, grid-stride over the chunks
    if (pblocks < 1) pblocks = 1;
#define DANBO_PLM_LAUNCH(ONE, OCC) aggmma::pair_logits_mma_kernel<ONE, OCC><<<pblocks, 128, 0, st>>>( \
        rays, ray_stride, S, z, active_ids, pose_skts, pose_vol, rays_per_pose, n_poses, fc, pw, pair_capacity, \
        (const uint32_t*)frags, logits)
    if (n_poses == 1) { if (occ == 3) DANBO_PLM_LAUNCH(true, 3); else if (occ == 4) DANBO_PLM_LAUNCH(true, 4); else DANBO_PLM_LAUNCH(true, 5); }
    else { if (occ == 3) DANBO_PLM_LAUNCH(false, 3); else if (occ == 4) DANBO_PLM_LAUNCH(false, 4); else DANBO_PLM_LAUNCH(false, 5); }
#undef DANBO_PLM_LAUNCH
    DANBO_CHECK_LAUNCH();
    return 0;
}

}  // namespace danbo

// Packs prob_linears' layer-0 / layer-1 weights (consts[2] = (24,15,32), consts[6] = (24,32,32), fp32) into split-bf16
// MMA B fragments: `frags` = danbo_agg_frag_bytes() bytes of device memory.  Re-run after every weight update.
// Resident blocks per SM (3, 4 or 5) the tensor-core aggregation-net kernel is compiled / launched for -> previous value.
extern "C" int danbo_pair_logits_set_blocks(int blocks) { return danbo::set_pair_logits_blocks(blocks); }

extern "C" int danbo_agg_frag_bytes(void) {
    return (danbo::aggmma::kW0Words + danbo::aggmma::kW1Words) * 4;
}

extern "C" int danbo_pack_agg_frags(const float* const* consts, void* frags, void* stream) {
    if (!consts || !frags) return -1;
    const int n = (danbo::aggmma::kW0Words + danbo::aggmma::kW1Words) / 2;
    danbo::aggmma::pack_frags_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(consts[2], consts[6], (uint32_t*)frags);
    DANBO_CHECK_LAUNCH();
    return 0;
}
