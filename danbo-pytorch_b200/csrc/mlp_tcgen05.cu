// M1: density/colour MLP of the DANBO field as ONE persistent tcgen05 kernel (sm_100a).
//
// Reference being replaced: core/networks/nerf.py:164-209 (inference_batchify / forward_density /
// forward_view, ten addmm launches + relu/cat per 65 536-point chunk, every activation round-tripping HBM).
//
// Design:
//   * one CTA per SM, 10 warps: warp 0 = bulk-copy (TMA) producer, warp 1 = MMA issuer (one elected lane),
//     warps 2-9 = epilogue (two warps per TMEM lane quarter, thread <-> sample row x 64-column slice).
//   * tile = 128 samples.  Accumulators live in TMEM as two N=128 halves (columns 0-255, fp32).  The activation
//     of every layer is written back to TMEM as packed bf16 (two ping-pong buffers, columns 256-511) and is the
//     A operand of the next layer's tcgen05.mma straight from TMEM -- activations never touch shared or global
//     memory.  Only the encoded input X (195 -> K 208, needed by layers 0 and 5) sits in shared memory.
//   * weights (bf16, 1.3 MB, L2 resident) stream through a 9-deep ring of 16 KB stages
//     ([128 out-rows x 64 k], canonical 128B-swizzled K-major) with cp.async.bulk + mbarrier complete_tx; the
//     pack kernel below writes them to global memory already in that shared-memory image, so no tensor map is
//     needed.
//   * the epilogue of output half 0 overlaps the MMAs of half 1; the next layer starts on the K range that half 0
//     produced while half 1 is still in its epilogue.
//   * heads: sigma = w_alpha . relu(h7) in the layer-7 epilogue (fp32 FFMA), rgb = W_rgb . relu(view layer) in the
//     last epilogue; the per-ray part of the view layer (view PE + frame code, 155 inputs) arrives as a per-ray
//     128-vector computed once per ray (ray_bias kernel).
#include "tc_common.cuh"

namespace danbo {
namespace mlp {

constexpr int kStages = 9;
constexpr int kStageBytes = 128 * 64 * 2;          // 16 KB
constexpr int kXBytes = DANBO_X_TILE_BYTES;        // 64 KB
constexpr int kNumHeadFloats = 9 * 256 + 256 + 3 * 128 + 4;   // biases L0..L8, w_alpha, W_rgb, b_alpha, b_rgb[3]
constexpr int kThreads = 320;            // producer warp, MMA warp, 8 epilogue warps

// TMEM column map
constexpr uint32_t kAccCol = 0;        // + 128*h
constexpr uint32_t kActCol = 256;      // + 128*buf

// stages consumed per tile (full / density-only)
__host__ __device__ constexpr int stages_per_tile(bool full) { return full ? 84 : 72; }

struct __align__(1024) Smem {
    uint8_t x[kXBytes];
    uint8_t w[kStages][kStageBytes];
    float heads[kNumHeadFloats + 4];
    float4 part[DANBO_TILE_M];          // second column slice's partial (rgb, sigma) of every row
    uint64_t w_full[kStages];
    uint64_t w_empty[kStages];
    uint64_t x_full, x_empty;
    uint64_t acc_full[2];
    uint64_t act_ready[2];
    uint32_t tmem_base;
};

using namespace danbo::tc;

// Layer plan (L = 0..9): 0 = pts_linears.0 (A = X), 1-4, 5 = skip layer (A = [X ; act]), 6, 7, 8 = feature_linear,
// 9 = views_linears.0 (feature part; N = 128).
__device__ __forceinline__ int n_halves(int L) { return L == 9 ? 1 : 2; }
__device__ __forceinline__ bool uses_x(int L) { return L == 0 || L == 5; }
__device__ __forceinline__ bool uses_act(int L) { return L != 0; }

template <bool kFull>
__global__ void __launch_bounds__(kThreads, 1)
mlp_kernel(const uint8_t* __restrict__ xtiles,       // [tiles][64 KB] swizzled bf16 X tiles
           const uint8_t* __restrict__ wstream,      // [84|72][16 KB] packed weight stages
           const float* __restrict__ heads,          // kNumHeadFloats
           const float* __restrict__ ray_bias,       // [n_rays][128] (full mode)
           const int* __restrict__ row_sample,       // [rows] destination sample id of every row
           const int* __restrict__ row_ray,          // [rows] ray of every row (full mode)
           const int* __restrict__ n_rows_ptr,       // device scalar: number of valid rows
           float* __restrict__ out,                  // full: raw [*,4]; density: sigma [*]
           int out_capacity,
           __nv_bfloat16* __restrict__ act_save,     // train mode: [9][save_cap][256] activations of L0..L8, or null
           __nv_bfloat16* __restrict__ g_save,       // train mode: [save_cap][128] relu(view layer), or null
           int save_cap,
           long long* __restrict__ trace) {          // optional clock64 timeline of CTA 0 (profiling aid) or null
    extern __shared__ uint8_t smem_raw[];
    Smem& S = *reinterpret_cast<Smem*>(smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_rows = *n_rows_ptr;
    const int n_tiles = (n_rows + DANBO_TILE_M - 1) / DANBO_TILE_M;
    constexpr int kLayers = kFull ? 10 : 8;

    for (int i = threadIdx.x; i < kNumHeadFloats; i += kThreads) S.heads[i] = heads[i];
    if (threadIdx.x == 0) {
        for (int s = 0; s < kStages; ++s) { mbar_init(&S.w_full[s], 1); mbar_init(&S.w_empty[s], 1); }
        mbar_init(&S.x_full, 1); mbar_init(&S.x_empty, 1);
        mbar_init(&S.acc_full[0], 1); mbar_init(&S.acc_full[1], 1);
        mbar_init(&S.act_ready[0], 8); mbar_init(&S.act_ready[1], 8);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&S.tmem_base)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = S.tmem_base;
    const bool tr = (trace != nullptr) && blockIdx.x == 0;
    // trace layout: [tile_iter < 4][role: 0 = mma, 1 = epilogue][(L*2+h)][begin, end]
#define DANBO_TRACE(it, role, L, h, which) do { if (tr && (it) < 4) trace[(((it) * 2 + (role)) * 20 + (L) * 2 + (h)) * 2 + (which)] = clock64(); } while (0)

    if (warp == 0) {
        // ===== producer: X tile + weight stages (one elected lane issues, the warp stays converged) =====
        const bool leader = elect_one();
        uint32_t ws = 0, wphase = 0, xphase = 0;
        for (int t = blockIdx.x, it = 0; t < n_tiles; t += gridDim.x, ++it) {
            if (it > 0) { mbar_wait(&S.x_empty, xphase); xphase ^= 1; }
            const uint8_t* xs = xtiles + (size_t)t * kXBytes;
            if (leader) {
                mbar_expect_tx(&S.x_full, kXBytes);
#pragma unroll
                for (int c = 0; c < 4; ++c) bulk_g2s(S.x + c * 16384, xs + c * 16384, 16384, &S.x_full);
            }
            for (int s = 0; s < stages_per_tile(kFull); ++s) {
                mbar_wait(&S.w_empty[ws], wphase ^ 1);
                if (leader) {
                    mbar_expect_tx(&S.w_full[ws], kStageBytes);
                    bulk_g2s(S.w[ws], wstream + (size_t)s * kStageBytes, kStageBytes, &S.w_full[ws]);
                }
                if (++ws == kStages) { ws = 0; wphase ^= 1; }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer: the whole warp runs the (uniform) schedule, one elected lane issues tcgen05 =====
        // Keeping the control flow warp-uniform lets descriptors and TMEM addresses live in uniform registers; a
        // lane-0-only loop costs ~230 clk per MMA in R2UR waterfalls (measured), 3.6x the MMA itself.
        const bool leader = elect_one();
        uint32_t ws = 0, wphase = 0, xphase = 0, r0 = 0, r1 = 0;
        const uint32_t x_base = smem_u32(S.x);
        const uint64_t desc_hi = make_desc(0);
        for (int t = blockIdx.x, it = 0; t < n_tiles; t += gridDim.x, ++it) {
            mbar_wait(&S.x_full, xphase); xphase ^= 1;
            if (it > 0) {                                                   // accumulators drained by the last epilogues
                mbar_wait(&S.act_ready[0], r0 & 1); ++r0;
                if (!kFull) { mbar_wait(&S.act_ready[1], r1 & 1); ++r1; }    // density-only ends on a two-half layer
            }
            tc_fence_after();
            for (int L = 0; L < kLayers; ++L) {
                if (L > 0) { mbar_wait(&S.act_ready[0], r0 & 1); ++r0; tc_fence_after(); }
                bool got_r1 = (L == 0);
                const uint32_t act_in = tmem + kActCol + 128u * ((L - 1) & 1);
                for (int h = 0; h < n_halves(L); ++h) {
                    const uint32_t d = tmem + kAccCol + 128u * h;
                    const int n_xc = uses_x(L) ? 4 : 0;
                    const int n_chunks = n_xc + (uses_act(L) ? 4 : 0);
                    for (int c = 0; c < n_chunks; ++c) {
                        const bool is_x = c < n_xc;
                        const int kc = is_x ? c : c - n_xc;
                        if (!is_x && kc == 2 && !got_r1) {
                            mbar_wait(&S.act_ready[1], r1 & 1); ++r1; got_r1 = true;
                        }
                        mbar_wait(&S.w_full[ws], wphase);
                        tc_fence_after();
                        if (c == 0) { DANBO_TRACE(it, 0, L, h, 0); }
                        const uint64_t bdesc = desc_hi | (uint64_t)((smem_u32(S.w[ws]) >> 4) & 0x3FFF);
                        const uint32_t acc0 = c > 0 ? 1u : 0u;
                        if (is_x) {
                            const uint64_t adesc = desc_hi | (uint64_t)(((x_base + kc * 16384) >> 4) & 0x3FFF);
                            if (leader) {
                                mma_ss(d, adesc, bdesc, kIdesc, acc0);
                                if (kc < 3) {                                   // K = 208: the last X chunk holds one k-step
                                    mma_ss(d, adesc + 2, bdesc + 2, kIdesc, 1u);
                                    mma_ss(d, adesc + 4, bdesc + 4, kIdesc, 1u);
                                    mma_ss(d, adesc + 6, bdesc + 6, kIdesc, 1u);
                                }
                            }
                        } else {
                            const uint32_t a_t = act_in + kc * 32;
                            if (leader) {
                                mma_ts(d, a_t, bdesc, kIdesc, acc0);
                                mma_ts(d, a_t + 8, bdesc + 2, kIdesc, 1u);
                                mma_ts(d, a_t + 16, bdesc + 4, kIdesc, 1u);
                                mma_ts(d, a_t + 24, bdesc + 6, kIdesc, 1u);
                            }
                        }
                        if (leader) {
                            tc_commit(&S.w_empty[ws]);
                            if (L == 5 && h == 1 && is_x && kc == 3) tc_commit(&S.x_empty);   // X no longer needed
                        }
                        __syncwarp();
                        if (++ws == kStages) { ws = 0; wphase ^= 1; }
                    }
                    if (leader) tc_commit(&S.acc_full[h]);
                    __syncwarp();
                    DANBO_TRACE(it, 0, L, h, 1);
                }
            }
        }
    } else {
        // ===== epilogue warps 2..9: two warps per TMEM lane quarter, each owns 64 of the 128 columns of a half =====
        const int q = warp & 3;                       // TMEM lane quarter this warp may access
        const int ch = (warp - 2) >> 2;               // which 64-column slice of every half
        const int row = q * 32 + lane;
        const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
        const float* bias = S.heads;                  // [9][256]
        const float* w_alpha = S.heads + 9 * 256;     // [256]
        const float* w_rgb = w_alpha + 256;           // [3][128]
        const float* tail = w_rgb + 3 * 128;          // b_alpha, b_rgb[3]
        uint32_t f0 = 0, f1 = 0;
        for (int t = blockIdx.x, it = 0; t < n_tiles; t += gridDim.x, ++it) {
            const int grow = t * DANBO_TILE_M + row;
            const bool valid = grow < n_rows;
            const int sample = valid ? row_sample[grow] : -1;
            const int ray = (kFull && valid) ? row_ray[grow] : 0;
            float alpha = ch == 0 ? tail[0] : 0.f;
            float rgb[3] = {ch == 0 ? tail[1] : 0.f, ch == 0 ? tail[2] : 0.f, ch == 0 ? tail[3] : 0.f};
            for (int L = 0; L < kLayers; ++L) {
                for (int h = 0; h < n_halves(L); ++h) {
                    if (h == 0) { mbar_wait(&S.acc_full[0], f0 & 1); ++f0; }
                    else        { mbar_wait(&S.acc_full[1], f1 & 1); ++f1; }
                    tc_fence_after();
                    if (warp == 2 && lane == 0) DANBO_TRACE(it, 1, L, h, 0);
                    const uint32_t acc = tmem + lane_addr + kAccCol + 128u * h + 64u * ch;
                    const uint32_t act_out = tmem + lane_addr + kActCol + 128u * (L & 1) + 64u * h + 32u * ch;
                    uint32_t v[2][32];
                    tmem_ld32(acc, v[0]);
                    tmem_ld32(acc + 32, v[1]);
                    tmem_wait_ld();
                    const int col0 = h * 128 + ch * 64;            // first output column of this thread's slice
                    if (L < 9) {
#pragma unroll
                        for (int g = 0; g < 2; ++g) {
                            const float4* b4 = reinterpret_cast<const float4*>(bias + L * 256 + col0 + 32 * g);
                            uint32_t pk[16];
#pragma unroll
                            for (int i = 0; i < 8; ++i) {
                                const float4 b = b4[i];
                                float a0 = __uint_as_float(v[g][4 * i + 0]) + b.x;
                                float a1 = __uint_as_float(v[g][4 * i + 1]) + b.y;
                                float a2 = __uint_as_float(v[g][4 * i + 2]) + b.z;
                                float a3 = __uint_as_float(v[g][4 * i + 3]) + b.w;
                                if (L == 7) {
                                    a0 = fmaxf(a0, 0.f); a1 = fmaxf(a1, 0.f); a2 = fmaxf(a2, 0.f); a3 = fmaxf(a3, 0.f);
                                    const float4 wa = *reinterpret_cast<const float4*>(w_alpha + col0 + 32 * g + 4 * i);
                                    alpha = fmaf(a0, wa.x, alpha); alpha = fmaf(a1, wa.y, alpha);
                                    alpha = fmaf(a2, wa.z, alpha); alpha = fmaf(a3, wa.w, alpha);
                                }
                                if (L < 8) { pk[2 * i] = pack_bf16_relu(a0, a1); pk[2 * i + 1] = pack_bf16_relu(a2, a3); }
                                else       { pk[2 * i] = pack_bf16(a0, a1);      pk[2 * i + 1] = pack_bf16(a2, a3); }
                            }
                            if (kFull || L < 7) tmem_st16(act_out + 16 * g, pk);
                            if (act_save != nullptr && valid) {          // saved for the backward pass (bf16, row-major)
                                uint4* dst = reinterpret_cast<uint4*>(act_save + ((size_t)L * save_cap + grow) * 256 + col0 + 32 * g);
#pragma unroll
                                for (int i = 0; i < 4; ++i) dst[i] = make_uint4(pk[4 * i], pk[4 * i + 1], pk[4 * i + 2], pk[4 * i + 3]);
                            }
                        }
                    } else {
#pragma unroll
                        for (int g = 0; g < 2; ++g) {
                            const float4* b4 = reinterpret_cast<const float4*>(ray_bias + (size_t)ray * 128 + col0 + 32 * g);
#pragma unroll
                            for (int i = 0; i < 8; ++i) {
                                const float4 b = __ldg(b4 + i);
                                const float a[4] = {fmaxf(__uint_as_float(v[g][4 * i + 0]) + b.x, 0.f),
                                                    fmaxf(__uint_as_float(v[g][4 * i + 1]) + b.y, 0.f),
                                                    fmaxf(__uint_as_float(v[g][4 * i + 2]) + b.z, 0.f),
                                                    fmaxf(__uint_as_float(v[g][4 * i + 3]) + b.w, 0.f)};
#pragma unroll
                                for (int k = 0; k < 3; ++k) {
                                    const float4 w = *reinterpret_cast<const float4*>(w_rgb + k * 128 + col0 + 32 * g + 4 * i);
                                    rgb[k] = fmaf(a[0], w.x, rgb[k]); rgb[k] = fmaf(a[1], w.y, rgb[k]);
                                    rgb[k] = fmaf(a[2], w.z, rgb[k]); rgb[k] = fmaf(a[3], w.w, rgb[k]);
                                }
                                if (g_save != nullptr && valid) {
                                    uint2* dst = reinterpret_cast<uint2*>(g_save + (size_t)grow * 128 + col0 + 32 * g + 4 * i);
                                    *dst = make_uint2(pack_bf16(a[0], a[1]), pack_bf16(a[2], a[3]));
                                }
                            }
                        }
                    }
                    tmem_wait_st();
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&S.act_ready[h]);
                    if (warp == 2 && lane == 0) DANBO_TRACE(it, 1, L, h, 1);
                }
            }
            // combine the two column slices of every row and write the row's output
            if (ch == 1) S.part[row] = make_float4(rgb[0], rgb[1], rgb[2], alpha);
            asm volatile("bar.sync 1, 256;" ::: "memory");
            if (ch == 0 && sample >= 0 && sample < out_capacity) {
                const float4 o = S.part[row];
                if (kFull) *reinterpret_cast<float4*>(out + (size_t)sample * 4) = make_float4(rgb[0] + o.x, rgb[1] + o.y, rgb[2] + o.z, alpha + o.w);
                else out[sample] = alpha + o.w;
            }
            asm volatile("bar.sync 1, 256;" ::: "memory");
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
    }
}

// -----------------------------------------------------------------------------------------------------------
// Weight packing: fp32 nn.Linear weights (out,in) -> bf16 stage stream in consumption order + fp32 head vector.
// One thread per bf16 element of the stream.
struct PackArgs {
    const float* w[8];        // pts_linears.0..7 weight
    const float* b[8];        // pts_linears.0..7 bias
    const float* w_alpha; const float* b_alpha;
    const float* w_feat;  const float* b_feat;
    const float* w_view;      // (128, 411): [feature 256 | view PE 27 | code 128]
    const float* b_view;
    const float* w_rgb;   const float* b_rgb;
};
constexpr int kRayBiasFloats = 155 * 128 + 128;      // W_v[:, 256:411]^T (155,128) then b_v (128)

// stage s of the per-tile stream -> (layer, half, is_x, kc)
__host__ __device__ inline void stage_to_layer(int s, int& L, int& h, int& is_x, int& kc) {
    const int per_layer[10] = {8, 8, 8, 8, 8, 16, 8, 8, 8, 4};
    L = 0;
    while (s >= per_layer[L]) { s -= per_layer[L]; ++L; }
    const int per_half = (L == 5) ? 8 : 4;
    h = s / per_half;
    int c = s % per_half;
    is_x = (L == 0) || (L == 5 && c < 4);
    kc = (L == 5 && c >= 4) ? c - 4 : c;
}

__global__ void pack_weights_kernel(PackArgs a, __nv_bfloat16* __restrict__ wstream, float* __restrict__ heads,
                                    float* __restrict__ wv_ray) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < kRayBiasFloats; i += gridDim.x * blockDim.x) {
        if (i < 155 * 128) { const int c = i / 128, o = i % 128; wv_ray[i] = a.w_view[(size_t)o * 411 + 256 + c]; }
        else wv_ray[i] = a.b_view[i - 155 * 128];
    }
    const int total = 84 * 128 * 64;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int s = i / (128 * 64), e = i % (128 * 64);
        const int n = e / 64, k = e % 64;
        int L, h, is_x, kc;
        stage_to_layer(s, L, h, is_x, kc);
        const int out_row = h * 128 + n;
        const int kin = kc * 64 + k;
        float val = 0.f;
        if (L <= 7) {
            const int in_dim = (L == 0) ? DANBO_X_COLS : (L == 5 ? DANBO_X_COLS + 256 : 256);
            if (L == 0) { if (kin < DANBO_X_COLS) val = a.w[0][(size_t)out_row * in_dim + kin]; }
            else if (L == 5) {
                if (is_x) { if (kin < DANBO_X_COLS) val = a.w[5][(size_t)out_row * in_dim + kin]; }
                else val = a.w[5][(size_t)out_row * in_dim + DANBO_X_COLS + kin];
            } else val = a.w[L][(size_t)out_row * in_dim + kin];
        } else if (L == 8) {
            val = a.w_feat[(size_t)out_row * 256 + kin];
        } else {
            val = a.w_view[(size_t)out_row * 411 + kin];          // out_row < 128 because h == 0
        }
        const uint32_t off = sw128_offset((uint32_t)n, (uint32_t)k);     // chunk index is 0 for k < 64
        wstream[(size_t)s * (128 * 64) + off / 2] = __float2bfloat16_rn(val);
    }
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < kNumHeadFloats; i += gridDim.x * blockDim.x) {
        float v;
        if (i < 8 * 256) v = a.b[i / 256][i % 256];
        else if (i < 9 * 256) v = a.b_feat[i - 8 * 256];
        else if (i < 10 * 256) v = a.w_alpha[i - 9 * 256];
        else if (i < 10 * 256 + 384) v = a.w_rgb[i - 10 * 256];
        else if (i == 10 * 256 + 384) v = a.b_alpha[0];
        else v = a.b_rgb[i - (10 * 256 + 385)];
        heads[i] = v;
    }
}

}  // namespace mlp
}  // namespace danbo

using namespace danbo;

extern "C" int danbo_mlp_workspace_bytes(long long* wstream_bytes, long long* heads_bytes, long long* raybias_bytes) {
    *wstream_bytes = 84LL * mlp::kStageBytes;
    *heads_bytes = (long long)mlp::kNumHeadFloats * 4;
    *raybias_bytes = (long long)mlp::kRayBiasFloats * 4;
    return 0;
}

extern "C" int danbo_pack_mlp_weights(const float* const* w_pts, const float* const* b_pts, const float* w_alpha,
                                      const float* b_alpha, const float* w_feat, const float* b_feat,
                                      const float* w_view, const float* b_view, const float* w_rgb,
                                      const float* b_rgb, void* wstream, float* heads, float* wv_ray, void* stream) {
    mlp::PackArgs a;
    for (int i = 0; i < 8; ++i) { a.w[i] = w_pts[i]; a.b[i] = b_pts[i]; }
    a.w_alpha = w_alpha; a.b_alpha = b_alpha; a.w_feat = w_feat; a.b_feat = b_feat;
    a.w_view = w_view; a.b_view = b_view; a.w_rgb = w_rgb; a.b_rgb = b_rgb;
    mlp::pack_weights_kernel<<<296, 256, 0, (cudaStream_t)stream>>>(a, (__nv_bfloat16*)wstream, heads, wv_ray);
    DANBO_CHECK_LAUNCH();
    return 0;
}

static int mlp_launch(const void* xtiles, const void* wstream, const float* heads, const float* ray_bias,
                      const int* row_sample, const int* row_ray, const int* n_rows_dev, int max_rows,
                      float* out, int out_capacity, int density_only, int num_sms, void* act_save, void* g_save,
                      int save_cap, long long* trace, void* stream) {
    if (max_rows <= 0) return 0;
    const int smem = (int)sizeof(mlp::Smem) + 1024;
    int max_tiles = (max_rows + DANBO_TILE_M - 1) / DANBO_TILE_M;
    int grid = num_sms < max_tiles ? num_sms : max_tiles;
    if (grid < 1) grid = 1;
    cudaError_t e;
    static bool attr_set[2] = {false, false};           // idempotent launch attribute, set on first use per variant
    if (density_only) {
        if (!attr_set[0]) {
            e = cudaFuncSetAttribute(mlp::mlp_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
            if (e != cudaSuccess) return (int)e;
            attr_set[0] = true;
        }
        mlp::mlp_kernel<false><<<grid, mlp::kThreads, smem, (cudaStream_t)stream>>>(
            (const uint8_t*)xtiles, (const uint8_t*)wstream, heads, ray_bias, row_sample, row_ray, n_rows_dev, out, out_capacity,
            (__nv_bfloat16*)act_save, (__nv_bfloat16*)g_save, save_cap, trace);
    } else {
        if (!attr_set[1]) {
            e = cudaFuncSetAttribute(mlp::mlp_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
            if (e != cudaSuccess) return (int)e;
            attr_set[1] = true;
        }
        mlp::mlp_kernel<true><<<grid, mlp::kThreads, smem, (cudaStream_t)stream>>>(
            (const uint8_t*)xtiles, (const uint8_t*)wstream, heads, ray_bias, row_sample, row_ray, n_rows_dev, out, out_capacity,
            (__nv_bfloat16*)act_save, (__nv_bfloat16*)g_save, save_cap, trace);
    }
    DANBO_CHECK_LAUNCH();
    return 0;
}

extern "C" int danbo_mlp_forward(const void* xtiles, const void* wstream, const float* heads, const float* ray_bias,
                                 const int* row_sample, const int* row_ray, const int* n_rows_dev, int max_rows,
                                 float* out, int out_capacity, int density_only, int num_sms, void* stream) {
    return mlp_launch(xtiles, wstream, heads, ray_bias, row_sample, row_ray, n_rows_dev, max_rows, out, out_capacity,
                      density_only, num_sms, nullptr, nullptr, 0, nullptr, stream);
}

// Train-mode forward: the same launch, and every layer's bf16 activation is kept for the backward pass:
// act_save [9][save_cap][256] = outputs of pts_linears.0..7 (after relu) and feature_linear; g_save [save_cap][128] =
// relu(views_linears.0).  save_cap >= max_rows.
extern "C" int danbo_mlp_forward_save(const void* xtiles, const void* wstream, const float* heads, const float* ray_bias,
                                      const int* row_sample, const int* row_ray, const int* n_rows_dev, int max_rows,
                                      float* out, int out_capacity, int num_sms, void* act_save, void* g_save,
                                      int save_cap, void* stream) {
    if (!act_save || !g_save || save_cap < max_rows) return -1;
    return mlp_launch(xtiles, wstream, heads, ray_bias, row_sample, row_ray, n_rows_dev, max_rows, out, out_capacity, 0,
                      num_sms, act_save, g_save, save_cap, nullptr, stream);
}

// Same launch, and CTA 0 writes a clock64 timeline of its first 4 tiles into trace[4*2*20*2] (profiling aid).
extern "C" int danbo_mlp_forward_trace(const void* xtiles, const void* wstream, const float* heads, const float* ray_bias,
                                       const int* row_sample, const int* row_ray, const int* n_rows_dev, int max_rows,
                                       float* out, int out_capacity, int density_only, int num_sms, long long* trace,
                                       void* stream) {
    return mlp_launch(xtiles, wstream, heads, ray_bias, row_sample, row_ray, n_rows_dev, max_rows, out, out_capacity,
                      density_only, num_sms, nullptr, nullptr, 0, trace, stream);
}
