// AN1: the A-NeRF field (nerf_type = nerf, BASELINE config #4) on sm_100a.
//
// Reference being replaced (per sample, fp32, every intermediate through HBM):
//   core/networks/nerf.py:222-279   encode_pts / encode_views
//   core/cutoff_embedder.py:151-214 CutoffEmbedder._embed   (cutoff positional encodings)
//   core/encoders.py:639-651 RelDistEncoder, :774-795 VecNormEncoder, :305-317 transform_batch_rays
//   core/networks/nerf.py:164-209   8 x 448 MLP with skip, alpha / feature / view (224) / rgb heads
//
// Kernels:
//   ray_kernel    per ray: direction encoding of the 24 bone frames (648 floats) and the frame-code part of the view
//                 layer (224 floats, W_v[:, 1096:1224] . code + b_v).
//   embed_kernel  per sample: 24 bone-frame distances -> cutoff PE (360) + unit vectors (72) and the per-sample cutoff
//                 weights applied to the ray's direction encoding (648); both written as bf16 K-major 128B-swizzled
//                 tile images (the operand layout tcgen05.mma reads from shared memory).
//   mlp_kernel    one persistent launch over CTA PAIRS (cta_group::2, M = 256): all ten layers of a 2 x 128-row tile
//                 back to back.  W = 448 does not leave room in TMEM for the activations (448 fp32 accumulator columns
//                 of 512), so every layer's bf16 activation tile goes to a per-CTA scratch in global memory (224 KB
//                 per CTA, 33 MB in all: L2 resident) in operand layout and is streamed back through the shared
//                 memory ring together with the weights.  A layer runs as two passes over N (224 columns each): the
//                 epilogue of pass 0 overlaps the MMAs of pass 1, and the next layer starts on the K range pass 0
//                 produced.
#include "tc_common.cuh"
#include "field_common.cuh"

namespace danbo {
namespace anerf {

using namespace danbo::tc;

constexpr int kW = 448;                  // netwidth
constexpr int kViewW = 224;
constexpr int kHalfN = 224;              // accumulator half (one pass over N)
constexpr int kQuartN = 112;             // B rows one CTA of the pair stages per pass
constexpr int kXD = 432, kChunksD = 7;   // density input, K padded to 448
constexpr int kXV = 648, kChunksV = 11;  // view input (without the frame code), K padded to 704
constexpr int kChunkBytes = 16384;       // [128 rows x 64 k] bf16
constexpr int kBBytes = kQuartN * 128;   // [112 rows x 64 k] bf16 = 14 336
constexpr int kSlotBytes = 2 * (kChunkBytes + kBBytes);   // a ring slot holds up to TWO k-chunks: A (32 KB) then B (28 KB)
constexpr int kSlots = 3;
constexpr int kStages = 158;             // k-chunks per tile: 14 + 4*14 + 28 + 14 + 14 + 14 + 18
constexpr int kGroups = 90;              // ring stages per tile (groups of one or two k-chunks), 30 per slot
constexpr int kThreads = 352;            // warp 0 A producer, 1 MMA issuer / relay, 2-9 epilogue, 10 B producer
constexpr int kActBytes = kChunksD * kChunkBytes;     // one activation tile image (114 688)
constexpr float kCutoff = 0.5f;          // cutoff_mm * ext_scale (run_nerf.py:498, encoders.py:62)

// heads buffer (fp32): biases of pts_linears.0-7 and feature_linear (9 x 448), w_alpha (448), W_rgb (3 x 224),
// b_alpha, b_rgb[3]
constexpr int kHeadBias = 0, kHeadWAlpha = 9 * kW, kHeadWRgb = kHeadWAlpha + kW, kHeadTail = kHeadWRgb + 3 * kViewW;
constexpr int kNumHeadFloats = kHeadTail + 4;
constexpr int kCodeFloats = 128 * kViewW + kViewW;    // W_v[:, 1096:1224]^T (128,224) then b_v (224)

__host__ __device__ constexpr int n_k(int L) { return L == 5 ? 14 : (L == 9 ? 18 : 7); }
__host__ __device__ constexpr int n_pass(int L) { return L == 9 ? 1 : 2; }
__host__ __device__ constexpr int stage_base(int L) {
    return L <= 5 ? 14 * L : (L == 6 ? 98 : (L == 7 ? 112 : (L == 8 ? 126 : 140)));
}
// A bulk copy costs its issuing warp ~330 clk whatever its size (scripts/micro/bulk_rate.cu), and one k-chunk is only
// 448 clk of MMAs: ring stages are therefore GROUPS of up to two k-chunks (one copy of A and one of B per group).
// The 7 chunks of an input run split (0,1)(2,3)(4,5)(6); the 7 chunks of an activation run split (0,1)(2)(3,4)(5,6) so
// that chunk 3 -- the first one holding columns of the previous layer's second pass -- starts a group; the 11 view
// chunks split (0,1)...(8,9)(10).
__host__ __device__ constexpr int groups_per_pass(int L) { return L == 5 ? 8 : (L == 9 ? 10 : 4); }
__host__ __device__ constexpr int group_base(int L) {
    return L <= 5 ? 8 * L : (L == 6 ? 56 : (L == 7 ? 64 : (L == 8 ? 72 : 80)));
}
__host__ __device__ constexpr bool first_run_is_input(int L) { return L == 0 || L == 5; }
__host__ __device__ constexpr int act_group_kc(int g) { return g == 0 ? 0 : (g == 1 ? 2 : (g == 2 ? 3 : 5)); }
// first k-chunk (layer-local) and number of chunks of group gi of a pass of layer L
__host__ __device__ constexpr int group_kc0(int L, int gi) {
    return gi < 4 ? (first_run_is_input(L) ? 2 * gi : act_group_kc(gi))
                  : 7 + (L == 5 ? act_group_kc(gi - 4) : 2 * (gi - 4));
}
__host__ __device__ constexpr int group_nch(int L, int gi) {
    return gi < 4 ? (first_run_is_input(L) ? (gi == 3 ? 1 : 2) : (gi == 1 ? 1 : 2))
                  : (L == 5 ? (gi - 4 == 1 ? 1 : 2) : (gi - 4 == 5 ? 1 : 2));
}

// ------------------------------------------------------------------------------------------------------------
// per-ray quantities
__global__ void __launch_bounds__(256)
ray_kernel(const float* __restrict__ rays, int ray_stride, int n_rays, const float* __restrict__ pose_skts,
           int rays_per_pose, int n_poses, const int* __restrict__ cam_idx, const float* __restrict__ codes, int n_codes,
           const float* __restrict__ w_code /* kCodeFloats */, float* __restrict__ ray_enc /* (n,648) */,
           float* __restrict__ code_bias /* (n,224) */) {
    const int n = blockIdx.x;
    if (n >= n_rays) return;
    const float* r = rays + (size_t)n * ray_stride;
    int pose = n / rays_per_pose; if (pose >= n_poses) pose = n_poses - 1;
    const int t = threadIdx.x;
    if (t < DANBO_J) {
        // transform_batch_rays: rotation part of the bone transform only, then VecNormEncoder (F.normalize, eps 1e-12)
        const float* skt = pose_skts + ((size_t)pose * DANBO_J + t) * 16;
        const float dx = r[3], dy = r[4], dz = r[5];
        const float l0 = fmaf(skt[2], dz, fmaf(skt[1], dy, skt[0] * dx));
        const float l1 = fmaf(skt[6], dz, fmaf(skt[5], dy, skt[4] * dx));
        const float l2 = fmaf(skt[10], dz, fmaf(skt[9], dy, skt[8] * dx));
        const float den = fmaxf(sqrtf(l0 * l0 + l1 * l1 + l2 * l2), 1e-12f);
        const float d[3] = {__fdiv_rn(l0, den), __fdiv_rn(l1, den), __fdiv_rn(l2, den)};
        float* e = ray_enc + (size_t)n * kXV;
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            e[t * 3 + a] = d[a];
#pragma unroll
            for (int f = 0; f < 4; ++f) {
                const float x = d[a] * (float)(1 << f);
                e[(1 + 2 * f) * 72 + t * 3 + a] = sinf(x);
                e[(2 + 2 * f) * 72 + t * 3 + a] = cosf(x);
            }
        }
    }
    if (t < kViewW) {
        // Optcodes (embedding.py:86-108): codes[cam] or the mean code (last row of `codes`) when cam < 0
        const int cam = cam_idx ? cam_idx[n] : -1;
        const float* code = codes + (size_t)((cam < 0 || cam >= n_codes) ? n_codes : cam) * 128;
        float acc = w_code[128 * kViewW + t];
        for (int c = 0; c < 128; ++c) acc = fmaf(code[c], w_code[c * kViewW + t], acc);
        code_bias[(size_t)n * kViewW + t] = acc;
    }
}

// ------------------------------------------------------------------------------------------------------------
// per-sample encodings -> operand tile images
// One block = one 128-row tile.  A lane owns a row, and a row's 16-byte units lie in different 128-byte lines of the
// image, so direct global stores cost one LSU pass per lane (the kernel was store-issue bound at 150 ms per 1024^2
// image).  The image is therefore assembled in shared memory (same swizzled layout: conflict-free 16-byte stores) and
// leaves as bulk copies: xd (112 KB), then the view image in two parts (112 KB + 64 KB).
constexpr int kEmbedStage = kChunksD * kChunkBytes;          // 112 KB of dynamic shared memory

__device__ __forceinline__ void emit_piece(uint8_t* __restrict__ tile, int rr, int col0, const float (&v)[8]) {
    const int chunk = col0 >> 6, k = col0 & 63;
    const uint32_t dst = smem_u32(tile) + chunk * kChunkBytes + (rr >> 3) * 1024 + (rr & 7) * 128 + (((k >> 3) ^ (rr & 7)) << 4);
    asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(dst), "r"(pack_bf16(v[0], v[1])), "r"(pack_bf16(v[2], v[3])),
                 "r"(pack_bf16(v[4], v[5])), "r"(pack_bf16(v[6], v[7])) : "memory");
}

// all threads: make the staged image visible to the async proxy, then one thread sends it to global memory
__device__ __forceinline__ void flush_stage(uint8_t* stage, uint8_t* gdst, uint32_t bytes) {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(stage)), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");     // the stage may be overwritten
    }
    __syncthreads();
}

__global__ void __launch_bounds__(128)
embed_kernel(const float* __restrict__ rays, int ray_stride, int S, const float* __restrict__ z, int n_rows,
             const float* __restrict__ pose_skts, int rays_per_pose, int n_poses, const float* __restrict__ align,
             const float* __restrict__ ray_enc, float tau, uint8_t* __restrict__ xd, uint8_t* __restrict__ xv) {
    extern __shared__ uint8_t embed_raw[];
    uint8_t* stage = embed_raw + ((1024u - (smem_u32(embed_raw) & 1023u)) & 1023u);
    const int tile = blockIdx.x, rr = threadIdx.x;
    const int e = tile * DANBO_TILE_M + rr;
    uint8_t* gd = xd + (size_t)tile * (kChunksD * kChunkBytes);
    uint8_t* gv = xv + (size_t)tile * (kChunksV * kChunkBytes);
    const float zero8[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    const bool live = e < n_rows;                                 // rows of the last tile beyond the data: zeros
    const int n = live ? e / S : 0;
    float w[DANBO_J], sn[DANBO_J], cs[DANBO_J];
    if (live) {
        const float* r = rays + (size_t)n * ray_stride;
        const float zz = z[e];
        const float px = __fadd_rn(r[0], __fmul_rn(r[3], zz));
        const float py = __fadd_rn(r[1], __fmul_rn(r[4], zz));
        const float pz = __fadd_rn(r[2], __fmul_rn(r[5], zz));
        int pose = n / rays_per_pose; if (pose >= n_poses) pose = n_poses - 1;
        const float* skt = pose_skts + (size_t)pose * DANBO_J * 16;
#pragma unroll
        for (int g = 0; g < 3; ++g) {
            float r8[3][8], iw[8];
#pragma unroll
            for (int jj = 0; jj < 8; ++jj) {
                const int j = 8 * g + jj;
                float q0, q1, q2;
                bone_aligned(skt + j * 16, align + j * 16, px, py, pz, q0, q1, q2);      // T1 + T2
                const float v = sqrtf(q0 * q0 + q1 * q1 + q2 * q2);                       // RelDistEncoder
                const float den = fmaxf(v, 1e-12f);
                r8[(3 * jj + 0) / 8][(3 * jj + 0) % 8] = __fdiv_rn(q0, den);            // VecNormEncoder (F.normalize)
                r8[(3 * jj + 1) / 8][(3 * jj + 1) % 8] = __fdiv_rn(q1, den);
                r8[(3 * jj + 2) / 8][(3 * jj + 2) % 8] = __fdiv_rn(q2, den);
                w[j] = 1.f - 1.f / (1.f + expf(-tau * (v - kCutoff)));                   // cutoff_embedder.py:177-184
                const float inp = kCutoff - v;                                            // cut_to_cutoff
                iw[jj] = inp * w[j];
                sincosf(inp * (2.f / kCutoff) - 1.f, &sn[j], &cs[j]);                     // shift_inputs, octave 0
            }
            emit_piece(stage, rr, 8 * g, iw);                                             // PE row 0: (c - v) w
            emit_piece(stage, rr, 360 + 24 * g, r8[0]);                                   // unit vectors (bone_type reldir)
            emit_piece(stage, rr, 360 + 24 * g + 8, r8[1]);
            emit_piece(stage, rr, 360 + 24 * g + 16, r8[2]);
        }
#pragma unroll
        for (int f = 0; f < 7; ++f) {               // rows 1+2f (sin) and 2+2f (cos); next octave by the double-angle step
#pragma unroll
            for (int g = 0; g < 3; ++g) {
                float a[8], b[8];
#pragma unroll
                for (int jj = 0; jj < 8; ++jj) { a[jj] = sn[8 * g + jj] * w[8 * g + jj]; b[jj] = cs[8 * g + jj] * w[8 * g + jj]; }
                emit_piece(stage, rr, (1 + 2 * f) * 24 + 8 * g, a);
                emit_piece(stage, rr, (2 + 2 * f) * 24 + 8 * g, b);
            }
            if (f < 6) {
#pragma unroll
                for (int j = 0; j < DANBO_J; ++j) {
                    const float s2 = 2.f * sn[j] * cs[j];
                    cs[j] = 1.f - 2.f * sn[j] * sn[j];
                    sn[j] = s2;
                }
            }
        }
        emit_piece(stage, rr, 432, zero8);
        emit_piece(stage, rr, 440, zero8);
    } else {
#pragma unroll
        for (int j = 0; j < DANBO_J; ++j) w[j] = 0.f;
        for (int c = 0; c < kChunksD * 64; c += 8) emit_piece(stage, rr, c, zero8);
    }
    flush_stage(stage, gd, kChunksD * kChunkBytes);
    // view input: the ray's direction encoding (9 rows x 72) times the sample's cutoff weight of each joint.
    // Columns 0..447 (7 chunks) first, then 448..703 (4 chunks, zero padded from 648).
    const float4* e4 = reinterpret_cast<const float4*>(ray_enc + (size_t)n * kXV);
#pragma unroll
    for (int part = 0; part < 2; ++part) {
#pragma unroll
        for (int p = 0; p < 9; ++p) {
#pragma unroll
            for (int i8 = 0; i8 < 9; ++i8) {
                const int col = p * 72 + 8 * i8;
                if ((col < 448) != (part == 0)) continue;
                float o[8];
                if (live) {
                    const float4 lo = __ldg(e4 + col / 4), hi = __ldg(e4 + col / 4 + 1);
                    const float ev[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
#pragma unroll
                    for (int t = 0; t < 8; ++t) o[t] = ev[t] * w[(8 * i8 + t) / 3];
                } else {
#pragma unroll
                    for (int t = 0; t < 8; ++t) o[t] = 0.f;
                }
                emit_piece(stage, rr, col - part * 448, o);
            }
        }
        if (part == 1) {
#pragma unroll
            for (int c = kXV; c < kChunksV * 64; c += 8) emit_piece(stage, rr, c - 448, zero8);
        }
        flush_stage(stage, gv + part * (7 * kChunkBytes), part == 0 ? 7 * kChunkBytes : 4 * kChunkBytes);
    }
    // bulk stores still in flight are completed before the grid ends; nothing reads them earlier (stream order)
    if (threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

// ------------------------------------------------------------------------------------------------------------
// weight packing
struct PackArgs {
    const float* w[8];        // pts_linears.0..7 weight
    const float* b[8];
    const float* w_alpha; const float* b_alpha;
    const float* w_feat;  const float* b_feat;
    const float* w_view;      // (224, 1224): [feature 448 | view PE 648 | code 128]
    const float* b_view;
    const float* w_rgb;   const float* b_rgb;
};

__host__ __device__ inline void stage_decode(int s, int& L, int& h, int& kc) {
    L = 9;
    for (int l = 0; l < 9; ++l) if (s < stage_base(l + 1)) { L = l; break; }
    const int r = s - stage_base(L);
    h = r / n_k(L);
    kc = r % n_k(L);
}

__global__ void pack_kernel(PackArgs a, __nv_bfloat16* __restrict__ wstream, float* __restrict__ heads,
                            float* __restrict__ w_code) {
    const int total = 2 * kStages * kQuartN * 64;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int rank = i / (kStages * kQuartN * 64);
        const int rem = i % (kStages * kQuartN * 64);
        const int s = rem / (kQuartN * 64), el = rem % (kQuartN * 64);
        const int n = el / 64, k = el % 64;
        int L, h, kc;
        stage_decode(s, L, h, kc);
        const int row = (L == 9 ? 0 : h * kHalfN) + rank * kQuartN + n;
        float val = 0.f;
        if (L == 0) {
            const int kin = kc * 64 + k;
            if (kin < kXD) val = a.w[0][(size_t)row * kXD + kin];
        } else if (L == 5) {
            if (kc < 7) { const int kin = kc * 64 + k; if (kin < kXD) val = a.w[5][(size_t)row * (kXD + kW) + kin]; }
            else val = a.w[5][(size_t)row * (kXD + kW) + kXD + (kc - 7) * 64 + k];
        } else if (L <= 7) {
            val = a.w[L][(size_t)row * kW + kc * 64 + k];
        } else if (L == 8) {
            val = a.w_feat[(size_t)row * kW + kc * 64 + k];
        } else {
            if (kc < 7) val = a.w_view[(size_t)row * 1224 + kc * 64 + k];
            else { const int kv = (kc - 7) * 64 + k; if (kv < kXV) val = a.w_view[(size_t)row * 1224 + kW + kv]; }
        }
        const uint32_t off = sw128_offset((uint32_t)n, (uint32_t)k);
        wstream[((size_t)rank * kStages + s) * (kQuartN * 64) + off / 2] = __float2bfloat16_rn(val);
    }
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < kNumHeadFloats; i += gridDim.x * blockDim.x) {
        float v;
        if (i < 8 * kW) v = a.b[i / kW][i % kW];
        else if (i < 9 * kW) v = a.b_feat[i - 8 * kW];
        else if (i < kHeadWRgb) v = a.w_alpha[i - kHeadWAlpha];
        else if (i < kHeadTail) v = a.w_rgb[i - kHeadWRgb];
        else if (i == kHeadTail) v = a.b_alpha[0];
        else v = a.b_rgb[i - kHeadTail - 1];
        heads[i] = v;
    }
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < kCodeFloats; i += gridDim.x * blockDim.x) {
        if (i < 128 * kViewW) { const int c = i / kViewW, o = i % kViewW; w_code[i] = a.w_view[(size_t)o * 1224 + kW + kXV + c]; }
        else w_code[i] = a.b_view[i - 128 * kViewW];
    }
}

// ------------------------------------------------------------------------------------------------------------
// the MLP
struct __align__(1024) Smem {
    uint8_t ring[kSlots][kSlotBytes];    // per slot: A chunk (16 KB) then this CTA's B rows (14 KB)
    uint8_t stage[8][4096];              // per epilogue warp: 32 rows x 8 units of 16 B, to turn row-per-lane stores into line-wide ones
    float head_w[kW + 3 * kViewW];       // w_alpha, W_rgb
    float4 part[DANBO_TILE_M];
    uint64_t w_full[kSlots];
    uint64_t w_empty[kSlots];
    uint64_t acc_full[2];
    uint64_t acc_free[2];                // leader CTA's copy is the one waited on (16 arrivals: both CTAs' epilogues)
    uint64_t act_written[2];             // local: this CTA's epilogue warps -> its A producer (8 arrivals)
    uint32_t tmem_base;
};

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* v) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t* v) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld32p(uint32_t taddr, uint32_t* v) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
          "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
          "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr) : "memory");
}

struct EpiCtx {
    int lane, row, ch;
    uint32_t tmem_lane;
    uint64_t* acc_bar; uint32_t acc_phase;
    uint32_t free_addr;        // leader's acc_free[h], shared::cluster address
    uint64_t* written_bar;     // local act_written[h] (null: this layer writes no activation)
    uint8_t* act_out;          // this CTA's activation tile image of the layer being produced
    uint32_t stage;            // shared::cta address of this warp's 4 KB staging block
    long long* dslot;          // profiling aid: 4 clock64 stamps of this call (null when not tracing)
};

// One thread's 112 columns (two rounds of 56) of one accumulator half.
// kKind 0: relu hidden; 1: relu + sigma head; 2: linear (feature_linear); 3: view layer (+ per-ray bias, relu) + rgb head
template <int kKind>
__device__ __forceinline__ void epi_half(const EpiCtx& E, int h, const float* __restrict__ bias /* at column h*224 + ch*112 */,
                                         const float* __restrict__ head_w /* w_alpha (same offset) or W_rgb + ch*112 */,
                                         float& alpha, float (&rgb)[3]) {
    const uint32_t acc = E.tmem_lane + 256u * h + (uint32_t)(E.ch * kQuartN);
    float part[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int rd = 0; rd < 2; ++rd) {
        float4 b[14];
#pragma unroll
        for (int i = 0; i < 14; ++i) b[i] = __ldg(reinterpret_cast<const float4*>(bias + rd * 56) + i);
        if (rd == 0) { mbar_wait(E.acc_bar, E.acc_phase); tc_fence_after(); if (E.dslot) E.dslot[0] = clock64(); }
        uint32_t v[56];
        tmem_ld32p(acc + rd * 56, v);
        tmem_ld16(acc + rd * 56 + 32, v + 32);
        tmem_ld8(acc + rd * 56 + 48, v + 48);
        tmem_wait_ld();
        if (E.dslot) E.dslot[1 + rd] = clock64();
        if (rd == 1) {                                  // accumulator half drained: the next pass may overwrite it
            tc_fence_before();
            __syncwarp();
            if (E.lane == 0) mbar_arrive_cluster(E.free_addr);
        }
        uint32_t pk[28];
#pragma unroll
        for (int i = 0; i < 14; ++i) {
            float a0 = __uint_as_float(v[4 * i + 0]) + b[i].x, a1 = __uint_as_float(v[4 * i + 1]) + b[i].y;
            float a2 = __uint_as_float(v[4 * i + 2]) + b[i].z, a3 = __uint_as_float(v[4 * i + 3]) + b[i].w;
            if (kKind == 1) {
                const float4 wa = reinterpret_cast<const float4*>(head_w + rd * 56)[i];
                part[0] = fmaf(fmaxf(a0, 0.f), wa.x, part[0]); part[1] = fmaf(fmaxf(a1, 0.f), wa.y, part[1]);
                part[2] = fmaf(fmaxf(a2, 0.f), wa.z, part[2]); part[3] = fmaf(fmaxf(a3, 0.f), wa.w, part[3]);
            }
            if (kKind == 3) {
                a0 = fmaxf(a0, 0.f); a1 = fmaxf(a1, 0.f); a2 = fmaxf(a2, 0.f); a3 = fmaxf(a3, 0.f);
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    const float4 wr = reinterpret_cast<const float4*>(head_w + k * kViewW + rd * 56)[i];
                    rgb[k] = fmaf(a0, wr.x, rgb[k]); rgb[k] = fmaf(a1, wr.y, rgb[k]);
                    rgb[k] = fmaf(a2, wr.z, rgb[k]); rgb[k] = fmaf(a3, wr.w, rgb[k]);
                }
            } else if (kKind == 2) {
                pk[2 * i] = pack_bf16(a0, a1); pk[2 * i + 1] = pack_bf16(a2, a3);
            } else {
                pk[2 * i] = pack_bf16_relu(a0, a1); pk[2 * i + 1] = pack_bf16_relu(a2, a3);
            }
        }
        if (kKind != 3) {
            // A lane owns a row, and a row's 16-byte units land in different 128-byte lines of the operand image: stored
            // directly, every STG touches 32 lines (the LSU needs a pass per line; 2 500 clk per round, measured).  Stage
            // the warp's 32 x 7 units in shared memory and copy them out 4 rows per instruction instead.
            const int col0 = h * kHalfN + E.ch * kQuartN + rd * 56;
#pragma unroll
            for (int pc = 0; pc < 7; ++pc) {
                const uint32_t a = E.stage + (uint32_t)(E.lane * 128 + ((pc ^ (E.lane & 7)) << 4));
                asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(a), "r"(pk[4 * pc]), "r"(pk[4 * pc + 1]), "r"(pk[4 * pc + 2]), "r"(pk[4 * pc + 3]) : "memory");
            }
            __syncwarp();
            const int u = E.lane & 7;
            const int c = col0 + 8 * u, chunk = c >> 6, k = c & 63;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int rr = 4 * i + (E.lane >> 3);
                if (u < 7) {
                    uint4 val;
                    const uint32_t a = E.stage + (uint32_t)(rr * 128 + ((u ^ (rr & 7)) << 4));
                    asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(val.x), "=r"(val.y), "=r"(val.z), "=r"(val.w) : "r"(a) : "memory");
                    const int rowg = (E.row & ~31) + rr;
                    uint8_t* dst = E.act_out + chunk * kChunkBytes + (rowg >> 3) * 1024 + (rowg & 7) * 128 + (((k >> 3) ^ (rowg & 7)) << 4);
                    *reinterpret_cast<uint4*>(dst) = val;
                }
            }
            __syncwarp();
        }
    }
    if (kKind == 1) alpha += (part[0] + part[1]) + (part[2] + part[3]);
    if (E.dslot) E.dslot[3] = clock64();
    if (kKind != 3) {
        // generic-proxy stores -> visible to the bulk (async proxy) loads of this CTA's A producer.  One lane fences for
        // the warp (bar.warp.sync orders the other lanes' stores before it); release only: an acq_rel fence also
        // invalidates the SM's L1 (CCTL.IVALL) and cost ~8 000 clk per pass when every thread issued it (measured).
        __syncwarp();
        if (E.lane == 0) {
            asm volatile("fence.release.gpu;" ::: "memory");
            asm volatile("fence.proxy.async.global;" ::: "memory");
            mbar_arrive(E.written_bar);
        }
    }
}

__global__ void __launch_bounds__(kThreads, 1)
mlp_kernel(const uint8_t* __restrict__ xd, const uint8_t* __restrict__ xv, const uint8_t* __restrict__ wstream,
           const float* __restrict__ heads, const float* __restrict__ code_bias, uint8_t* __restrict__ scratch,
           float* __restrict__ out /* raw (rows,4) */, int n_rows, int S, int out_capacity,
           long long* __restrict__ trace /* optional clock64 timeline of CTA 0: [2 tiles][2 roles][20][2], or null */,
           uint8_t* __restrict__ save /* train mode: [tile][9] activation tile images kept for the backward pass, or null */) {
    extern __shared__ uint8_t smem_raw[];
    Smem& Sm = *reinterpret_cast<Smem*>(smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_tiles = (n_rows + DANBO_TILE_M - 1) / DANBO_TILE_M;
    const uint32_t rank = cluster_ctarank();
    const bool lead_cta = rank == 0;
    const int unit = (int)(blockIdx.x >> 1), n_units = (int)(gridDim.x >> 1);
    const int n_work = (n_tiles + 1) / 2;
    uint8_t* act_buf = scratch + (size_t)blockIdx.x * (2 * kActBytes);

    for (int i = threadIdx.x; i < kW + 3 * kViewW; i += kThreads) Sm.head_w[i] = heads[kHeadWAlpha + i];
    if (threadIdx.x == 0) {
        for (int s = 0; s < kSlots; ++s) { mbar_init(&Sm.w_full[s], lead_cta ? 3 : 2); mbar_init(&Sm.w_empty[s], 1); }
        mbar_init(&Sm.acc_full[0], 1); mbar_init(&Sm.acc_full[1], 1);
        mbar_init(&Sm.acc_free[0], 16); mbar_init(&Sm.acc_free[1], 16);
        mbar_init(&Sm.act_written[0], 8); mbar_init(&Sm.act_written[1], 8);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&Sm.tmem_base)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem = Sm.tmem_base;
    const bool tr = trace != nullptr && blockIdx.x == 0;
#define ANERF_TRACE(it, role, L, h, which) do { if (tr && (it) < 2) trace[(((it) * 2 + (role)) * 20 + (L) * 2 + (h)) * 2 + (which)] = clock64(); } while (0)

    if (warp == 0 || warp == 10) {
        // ===== producers: warp 0 streams the A operand (input / activation chunks), warp 10 this CTA's B rows =====
        const bool is_a = warp == 0;
        const bool leader = elect_one();
        const uint8_t* wsrc = wstream + (size_t)rank * kStages * kBBytes;
        for (int w = unit, it = 0; w < n_work; w += n_units, ++it) {
            int t = 2 * w + (int)rank;
            if (t >= n_tiles) t = n_tiles - 1;                       // odd tail: a dummy tile (its rows are never written)
            const uint8_t* xd_t = xd + (size_t)t * (kChunksD * kChunkBytes);
            const uint8_t* xv_t = xv + (size_t)t * (kChunksV * kChunkBytes);
            int s = 0;
            const size_t t_act = (size_t)(2 * w + (int)rank);            // unclamped: where this CTA's epilogue writes
            for (int L = 0; L < 10; ++L) {
                // train mode keeps every layer's activation image ([tile][layer]) instead of the two ping-pong buffers
                const uint8_t* act_in = save ? save + (t_act * 9 + (size_t)(L > 0 ? L - 1 : 0)) * kActBytes
                                             : act_buf + ((L - 1) & 1) * kActBytes;
                const uint32_t wr_par = (uint32_t)(9 * it + L - 1) & 1u;       // act_written completions: 9 per tile
                for (int h = 0; h < n_pass(L); ++h) {
                    for (int gi = 0; gi < groups_per_pass(L); ++gi, ++s) {
                        const int slot = s % kSlots;
                        const uint32_t use = (uint32_t)(s / kSlots);            // 30 uses per slot per tile: parity repeats
                        mbar_wait(&Sm.w_empty[slot], (use & 1u) ^ 1u);
                        const int kc = group_kc0(L, gi), nch = group_nch(L, gi);
                        if (is_a) {
                            const uint8_t* src;
                            int ka = -1;                                        // activation chunk index, if any
                            if (L == 0) src = xd_t + kc * kChunkBytes;
                            else if (L == 5) { if (kc < 7) src = xd_t + kc * kChunkBytes; else { ka = kc - 7; src = act_in + ka * kChunkBytes; } }
                            else if (L == 9) { if (kc < 7) { ka = kc; src = act_in + ka * kChunkBytes; } else src = xv_t + (kc - 7) * kChunkBytes; }
                            else { ka = kc; src = act_in + ka * kChunkBytes; }
                            if (h == 0 && ka == 0) mbar_wait(&Sm.act_written[0], wr_par);   // columns 0..223 of the previous layer
                            if (h == 0 && ka == 3) mbar_wait(&Sm.act_written[1], wr_par);   // chunk 3 straddles the halves
                            if (ka >= 0) asm volatile("fence.proxy.async.global;" ::: "memory");
                            if (leader) {
                                mbar_expect_tx(&Sm.w_full[slot], nch * kChunkBytes);
                                bulk_g2s(Sm.ring[slot], src, nch * kChunkBytes, &Sm.w_full[slot]);
                            }
                        } else if (leader) {
                            const int chunk_idx = stage_base(L) + h * n_k(L) + kc;          // position in the B stream
                            mbar_expect_tx(&Sm.w_full[slot], nch * kBBytes);
                            bulk_g2s(Sm.ring[slot] + 2 * kChunkBytes, wsrc + (size_t)chunk_idx * kBBytes, nch * kBBytes, &Sm.w_full[slot]);
                        }
                        __syncwarp();
                    }
                }
            }
        }
    } else if (warp == 1 && !lead_cta) {
        // ===== peer CTA: forward "both of my copies of this stage have landed" to the leader =====
        const uint32_t w_full_lead = mapa_cluster(smem_u32(&Sm.w_full[0]), 0);
        for (int w = unit, it = 0; w < n_work; w += n_units, ++it) {
            for (int s = 0; s < kGroups; ++s) {
                const int slot = s % kSlots;
                mbar_wait(&Sm.w_full[slot], (uint32_t)(s / kSlots) & 1u);
                if (lane == 0) mbar_arrive_cluster(w_full_lead + 8u * slot);
                __syncwarp();
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer (leader CTA): fully unrolled per-tile schedule, compile-time slots and descriptor offsets =====
        const bool leader = elect_one();
        const uint64_t desc0 = make_desc(0) | (uint64_t)((smem_u32(&Sm.ring[0][0]) >> 4) & 0x3FFF);
        constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t(kHalfN) >> 3) << 17) | ((256u >> 4) << 24);
        for (int w = unit, it = 0; w < n_work; w += n_units, ++it) {
#pragma unroll
            for (int L = 0; L < 10; ++L) {
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    if (h >= n_pass(L)) continue;
                    // the epilogue of the previous user of this accumulator half has drained it
                    if (h == 0) {
                        if (L > 0) mbar_wait(&Sm.acc_free[0], (uint32_t)(L - 1) & 1u);
                        else if (it > 0) mbar_wait(&Sm.acc_free[0], 1u);             // (L9, 0) of the previous tile: 10 per tile
                    } else {
                        if (L > 0) mbar_wait(&Sm.acc_free[1], (uint32_t)(9 * it + L - 1) & 1u);
                        else if (it > 0) mbar_wait(&Sm.acc_free[1], (uint32_t)(9 * it - 1) & 1u);
                    }
                    tc_fence_after();
                    const uint32_t d = tmem + 256u * h;
#pragma unroll
                    for (int gi = 0; gi < 10; ++gi) {
                        if (gi >= groups_per_pass(L)) continue;
                        const int s = group_base(L) + h * groups_per_pass(L) + gi;
                        const int slot = s % kSlots;
                        mbar_wait(&Sm.w_full[slot], (uint32_t)(s / kSlots) & 1u);
                        tc_fence_after();
                        if (gi == 0) { ANERF_TRACE(it, 0, L, h, 0); }
                        const uint64_t adesc = desc0 + (uint64_t)((slot * kSlotBytes) >> 4);
                        const uint64_t bdesc = desc0 + (uint64_t)((slot * kSlotBytes + 2 * kChunkBytes) >> 4);
                        if (leader) {
#pragma unroll
                            for (int c = 0; c < 2; ++c) {
                                if (c >= group_nch(L, gi)) continue;
                                const uint64_t ad = adesc + (uint64_t)((c * kChunkBytes) >> 4), bd = bdesc + (uint64_t)((c * kBBytes) >> 4);
                                mma_ss_pair(d, ad, bd, idesc, (gi > 0 || c > 0) ? 1u : 0u);
                                mma_ss_pair(d, ad + 2, bd + 2, idesc, 1u);
                                mma_ss_pair(d, ad + 4, bd + 4, idesc, 1u);
                                mma_ss_pair(d, ad + 6, bd + 6, idesc, 1u);
                            }
                            tc_commit_pair(&Sm.w_empty[slot]);
                            if (gi == groups_per_pass(L) - 1) tc_commit_pair(&Sm.acc_full[h]);
                        }
                        __syncwarp();
                    }
                    ANERF_TRACE(it, 0, L, h, 1);
                }
            }
        }
    } else {
        // ===== epilogue warps 2..9: two warps per TMEM lane quarter, each owns 112 of the 224 columns of a half =====
        EpiCtx E;
        const int q = warp & 3;
        E.ch = (warp - 2) >> 2;
        E.lane = lane;
        E.row = q * 32 + lane;
        E.tmem_lane = tmem + ((uint32_t)(q * 32) << 16);
        E.stage = smem_u32(Sm.stage[warp - 2]);
        const uint32_t free_addr[2] = {mapa_cluster(smem_u32(&Sm.acc_free[0]), 0), mapa_cluster(smem_u32(&Sm.acc_free[1]), 0)};
        const float* tail = heads + kHeadTail;
        for (int w = unit, it = 0; w < n_work; w += n_units, ++it) {
            const int t = 2 * w + (int)rank;
            const int grow = t * DANBO_TILE_M + E.row;
            const bool valid = grow < n_rows;
            const int ray = valid ? grow / S : 0;
            float alpha = E.ch == 0 ? tail[0] : 0.f;
            float rgb[3] = {E.ch == 0 ? tail[1] : 0.f, E.ch == 0 ? tail[2] : 0.f, E.ch == 0 ? tail[3] : 0.f};
            for (int L = 0; L < 10; ++L) {
                for (int h = 0; h < n_pass(L); ++h) {
                    E.acc_bar = &Sm.acc_full[h];
                    E.acc_phase = h == 0 ? ((uint32_t)L & 1u) : ((uint32_t)(9 * it + L) & 1u);   // 10 / 9 completions per tile
                    E.free_addr = free_addr[h];
                    E.written_bar = &Sm.act_written[h];
                    E.act_out = save ? save + ((size_t)t * 9 + (size_t)(L < 9 ? L : 8)) * kActBytes : act_buf + (L & 1) * kActBytes;
                    const int coff = h * kHalfN + E.ch * kQuartN;
                    if (warp == 2 && lane == 0) { ANERF_TRACE(it, 1, L, h, 0); }
                    E.dslot = (tr && it < 2 && warp == 2 && lane == 0) ? trace + 160 + ((it * 20 + L * 2 + h) * 4) : nullptr;
                    if (L < 7) epi_half<0>(E, h, heads + kHeadBias + L * kW + coff, nullptr, alpha, rgb);
                    else if (L == 7) epi_half<1>(E, h, heads + kHeadBias + L * kW + coff, Sm.head_w + coff, alpha, rgb);
                    else if (L == 8) epi_half<2>(E, h, heads + kHeadBias + L * kW + coff, nullptr, alpha, rgb);
                    else epi_half<3>(E, 0, code_bias + (size_t)ray * kViewW + E.ch * kQuartN, Sm.head_w + kW + E.ch * kQuartN, alpha, rgb);
                    if (warp == 2 && lane == 0) { ANERF_TRACE(it, 1, L, h, 1); }
                }
            }
            if (E.ch == 1) Sm.part[E.row] = make_float4(rgb[0], rgb[1], rgb[2], alpha);
            asm volatile("bar.sync 1, 256;" ::: "memory");
            if (E.ch == 0 && valid && grow < out_capacity) {
                const float4 o = Sm.part[E.row];
                *reinterpret_cast<float4*>(out + (size_t)grow * 4) = make_float4(rgb[0] + o.x, rgb[1] + o.y, rgb[2] + o.z, alpha + o.w);
            }
            asm volatile("bar.sync 1, 256;" ::: "memory");
        }
    }

    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

}  // namespace anerf
}  // namespace danbo

using namespace danbo;

extern "C" int danbo_anerf_workspace_bytes(int num_sms, long long* wstream_bytes, long long* heads_bytes,
                                           long long* code_bytes, long long* scratch_bytes) {
    *wstream_bytes = 2LL * anerf::kStages * anerf::kBBytes;
    *heads_bytes = (long long)anerf::kNumHeadFloats * 4;
    *code_bytes = (long long)anerf::kCodeFloats * 4;
    *scratch_bytes = (long long)num_sms * 2 * anerf::kActBytes;
    return 0;
}

extern "C" int danbo_anerf_pack_weights(const float* const* w_pts, const float* const* b_pts, const float* w_alpha,
                                        const float* b_alpha, const float* w_feat, const float* b_feat,
                                        const float* w_view, const float* b_view, const float* w_rgb, const float* b_rgb,
                                        void* wstream, float* heads, float* w_code, void* stream) {
    anerf::PackArgs a;
    for (int i = 0; i < 8; ++i) { a.w[i] = w_pts[i]; a.b[i] = b_pts[i]; }
    a.w_alpha = w_alpha; a.b_alpha = b_alpha; a.w_feat = w_feat; a.b_feat = b_feat;
    a.w_view = w_view; a.b_view = b_view; a.w_rgb = w_rgb; a.b_rgb = b_rgb;
    anerf::pack_kernel<<<592, 256, 0, (cudaStream_t)stream>>>(a, (__nv_bfloat16*)wstream, heads, w_code);
    DANBO_CHECK_LAUNCH();
    return 0;
}

extern "C" int danbo_anerf_ray_encode(const float* rays, int ray_stride, int n_rays, const float* pose_skts,
                                      int rays_per_pose, int n_poses, const int* cam_idx, const float* codes,
                                      int n_codes, const float* w_code, float* ray_enc, float* code_bias, void* stream) {
    if (n_rays <= 0) return 0;
    anerf::ray_kernel<<<n_rays, 256, 0, (cudaStream_t)stream>>>(rays, ray_stride, n_rays, pose_skts, rays_per_pose, n_poses,
                                                                 cam_idx, codes, n_codes, w_code, ray_enc, code_bias);
    DANBO_CHECK_LAUNCH();
    return 0;
}

extern "C" int danbo_anerf_embed(const float* rays, int ray_stride, int S, const float* z, int n_rows,
                                 const float* pose_skts, int rays_per_pose, int n_poses, const float* align,
                                 const float* ray_enc, float tau, void* xd, void* xv, void* stream) {
    if (n_rows <= 0) return 0;
    const int tiles = (n_rows + DANBO_TILE_M - 1) / DANBO_TILE_M;
    const int esmem = anerf::kEmbedStage + 1024;
    static bool embed_attr = false;
    if (!embed_attr) {
        cudaError_t e = cudaFuncSetAttribute(anerf::embed_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, esmem);
        if (e != cudaSuccess) return (int)e;
        embed_attr = true;
    }
    anerf::embed_kernel<<<tiles, 128, esmem, (cudaStream_t)stream>>>(rays, ray_stride, S, z, n_rows, pose_skts, rays_per_pose,
                                                                  n_poses, align, ray_enc, tau, (uint8_t*)xd, (uint8_t*)xv);
    DANBO_CHECK_LAUNCH();
    return 0;
}

static int anerf_mlp_launch(const void* xd, const void* xv, const void* wstream, const float* heads,
                            const float* code_bias, void* scratch, int n_rows, int S, float* out, int out_capacity,
                            int num_sms, long long* trace, void* save, void* stream) {
    if (n_rows <= 0) return 0;
    if (num_sms < 2) return -1;
    const int smem = (int)sizeof(anerf::Smem) + 1024;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(anerf::mlp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return (int)e;
        attr_set = true;
    }
    const int tiles = (n_rows + DANBO_TILE_M - 1) / DANBO_TILE_M;
    const int pairs = (tiles + 1) / 2, sm_pairs = num_sms / 2;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2 * (sm_pairs < pairs ? sm_pairs : pairs));
    cfg.blockDim = dim3(anerf::kThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = (cudaStream_t)stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, anerf::mlp_kernel, (const uint8_t*)xd, (const uint8_t*)xv, (const uint8_t*)wstream,
                                       heads, code_bias, (uint8_t*)scratch, out, n_rows, S, out_capacity, trace, (uint8_t*)save);
    if (e != cudaSuccess) return (int)e;
    DANBO_CHECK_LAUNCH();
    return 0;
}

extern "C" int danbo_anerf_mlp(const void* xd, const void* xv, const void* wstream, const float* heads,
                               const float* code_bias, void* scratch, int n_rows, int S, float* out, int out_capacity,
                               int num_sms, long long* trace, void* stream) {
    return anerf_mlp_launch(xd, xv, wstream, heads, code_bias, scratch, n_rows, S, out, out_capacity, num_sms, trace, nullptr, stream);
}

extern "C" int danbo_anerf_save_bytes(int n_rows, long long* bytes) {
    if (!bytes || n_rows < 0) return -1;
    const long long tiles = (n_rows + DANBO_TILE_M - 1) / DANBO_TILE_M;
    *bytes = ((tiles + 1) / 2 * 2) * 9 * (long long)anerf::kActBytes;    // tiles rounded up to a CTA pair
    return 0;
}

extern "C" int danbo_anerf_mlp_save(const void* xd, const void* xv, const void* wstream, const float* heads,
                                    const float* code_bias, void* scratch, int n_rows, int S, float* out, int out_capacity,
                                    int num_sms, void* save, void* stream) {
    if (!save) return -1;
    return anerf_mlp_launch(xd, xv, wstream, heads, code_bias, scratch, n_rows, S, out, out_capacity, num_sms, nullptr, save, stream);
}

namespace danbo { namespace anerf {
// Operand tile images ([tile][chunk][128 rows x 64 k] bf16, 128-byte swizzle) -> row-major bf16 (rows, n_chunks * 64).
// One thread per 16-byte piece; `tile_stride` bytes between consecutive tiles of the image (so one layer's plane can be
// read out of the [tile][9] activation save).
__global__ void __launch_bounds__(256)
untile_kernel(const uint8_t* __restrict__ img, long long tile_stride, int n_rows, int n_chunks, uint4* __restrict__ out) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int per_row = n_chunks * 8;
    const long long total = (long long)n_rows * per_row;
    if (idx >= total) return;
    const int row = (int)(idx / per_row), piece = (int)(idx - (long long)row * per_row);
    const int c = piece >> 3, p = piece & 7, t = row >> 7, r = row & 127;
    const uint8_t* src = img + (size_t)t * tile_stride + (size_t)c * kChunkBytes + (r >> 3) * 1024 + (r & 7) * 128 + ((p ^ (r & 7)) << 4);
    out[idx] = *reinterpret_cast<const uint4*>(src);
}
}}

extern "C" int danbo_anerf_untile(const void* img, long long tile_stride, int n_rows, int n_chunks, void* out, void* stream) {
    if (n_rows <= 0) return 0;
    if (n_chunks <= 0 || (tile_stride & 15)) return -1;
    const long long total = (long long)n_rows * n_chunks * 8;
    anerf::untile_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>((const uint8_t*)img, tile_stride, n_rows,
                                                                                      n_chunks, (uint4*)out);
    DANBO_CHECK_LAUNCH();
    return 0;
}
