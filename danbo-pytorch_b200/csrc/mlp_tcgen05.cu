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

constexpr int kStages = 9;                         // 16 KB units of shared memory reserved for the weight ring
// Ring slots actually used.  Chosen so that the stages of one tile (84 full / 72 density-only) are an EVEN multiple of
// the slot count: the slot and the mbarrier parity of every stage are then compile-time constants of the fully
// unrolled MMA schedule (runtime slot arithmetic kept descriptors in vector registers: R2UR chains and a waterfall
// loop around every commit, ~430 clk per 4-MMA stage instead of 256; measured).
__host__ __device__ constexpr int ring_slots(bool full, bool pair) { return pair ? (full ? 7 : 9) : (full ? 7 : 9); }
// Every ring slot is 16 KB and is filled by ONE bulk copy: issuing a cp.async.bulk costs the producer warp ~330 clk
// whatever its size (measured, scripts/micro/bulk_rate.cu), so copies must be few and large.  Single CTA: a slot is
// one stage ([128 x 64] of B).  CTA pair: a slot is this CTA's 64-row half of TWO consecutive stages (a "group"); the
// pack kernel writes a second copy of the stream in that order behind the first.
__host__ __device__ constexpr int stages_per_slot(bool pair) { return pair ? 2 : 1; }
constexpr int kStageBytes = 128 * 64 * 2;          // 16 KB
constexpr int kXBytes = DANBO_X_TILE_BYTES;        // 64 KB
constexpr int kNumHeadFloats = 9 * 256 + 256 + 3 * 128 + 4;   // biases L0..L8, w_alpha, W_rgb, b_alpha, b_rgb[3]
constexpr int kEmptyOff = kNumHeadFloats + 4;                 // heads[kEmptyOff ..): c[128], sigma0 of a sample no bone sees
constexpr int kEmptyFloats = 132;
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
    uint64_t w_full[2 * kStages];       // kStages slots of 16 KB, or (CTA pair) 2 * kStages slots of 8 KB
    uint64_t w_empty[2 * kStages];
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
// position of chunk c of (layer L, half h) in the per-tile stage stream (see stage_to_layer below)
__host__ __device__ constexpr int stage_index(int L, int h, int c) {
    return (L <= 5 ? 8 * L : 8 * L + 8) + h * (L == 5 ? 8 : 4) + c;
}

// ---- epilogue bodies ---------------------------------------------------------------------------------------
// One call = one thread's 64-column slice of one accumulator half.  The layer kind is a template parameter and all
// operands that do not depend on the accumulator (bias, per-ray bias) are fetched BEFORE the wait on the MMA, so the
// code between the TMEM load and the TMEM store is branch-free register arithmetic (a runtime `L ==` test inside
// the unrolled loops cost 1 500 clk per half in exposed LDS latency and branches; measured).
struct EpiCtx {
    int q, ch, lane;
    uint32_t tmem_lane, acc, act_out, phase;
    uint64_t* acc_bar;
    uint32_t done_addr;    // act_ready barrier: shared::cta address, or the leader CTA's shared::cluster address
    bool remote;
    long long* tslot;      // profiling aid (null when not tracing)
    long long* dslot;
};

__device__ __forceinline__ void epi_finish(const EpiCtx& E) {
    if (E.dslot) E.dslot[1] = clock64();
    tmem_wait_st();
    tc_fence_before();
    __syncwarp();
    if (E.lane == 0) {
        if (E.remote) mbar_arrive_cluster(E.done_addr);
        else asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(E.done_addr) : "memory");
    }
    if (E.tslot) E.tslot[1] = clock64();
}

// kKind 0: relu hidden layer; 1: relu + sigma head (pts_linears.7); 2: linear (feature_linear)
template <int kKind>
__device__ __forceinline__ void epi_hidden(const EpiCtx& E, const float* __restrict__ bias, const float* __restrict__ w_alpha,
                                           float& alpha, bool write_act, __nv_bfloat16* __restrict__ save_row) {
    float4 b[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) b[i] = reinterpret_cast<const float4*>(bias)[i];
    mbar_wait(E.acc_bar, E.phase);
    tc_fence_after();
    if (E.tslot) E.tslot[0] = clock64();
    uint32_t v[2][32];
    tmem_ld32(E.acc, v[0]);
    tmem_ld32(E.acc + 32, v[1]);
    tmem_wait_ld();
    if (E.dslot) E.dslot[0] = clock64();
    uint32_t pk[2][16];
    float part[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int g = 0; g < 2; ++g) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float4 bb = b[g * 8 + i];
            float a0 = __uint_as_float(v[g][4 * i + 0]) + bb.x;
            float a1 = __uint_as_float(v[g][4 * i + 1]) + bb.y;
            float a2 = __uint_as_float(v[g][4 * i + 2]) + bb.z;
            float a3 = __uint_as_float(v[g][4 * i + 3]) + bb.w;
            if (kKind == 1) {
                const float4 wa = reinterpret_cast<const float4*>(w_alpha)[g * 8 + i];
                part[0] = fmaf(fmaxf(a0, 0.f), wa.x, part[0]); part[1] = fmaf(fmaxf(a1, 0.f), wa.y, part[1]);
                part[2] = fmaf(fmaxf(a2, 0.f), wa.z, part[2]); part[3] = fmaf(fmaxf(a3, 0.f), wa.w, part[3]);
            }
            if (kKind != 2) { pk[g][2 * i] = pack_bf16_relu(a0, a1); pk[g][2 * i + 1] = pack_bf16_relu(a2, a3); }
            else            { pk[g][2 * i] = pack_bf16(a0, a1);      pk[g][2 * i + 1] = pack_bf16(a2, a3); }
        }
    }
    if (kKind == 1) alpha += (part[0] + part[1]) + (part[2] + part[3]);
    if (write_act) { tmem_st16(E.act_out, pk[0]); tmem_st16(E.act_out + 16, pk[1]); }
    if (save_row != nullptr) {                           // saved for the backward pass (bf16, row-major)
        uint4* dst = reinterpret_cast<uint4*>(save_row);
#pragma unroll
        for (int g = 0; g < 2; ++g)
#pragma unroll
            for (int i = 0; i < 4; ++i) dst[g * 4 + i] = make_uint4(pk[g][4 * i], pk[g][4 * i + 1], pk[g][4 * i + 2], pk[g][4 * i + 3]);
    }
    epi_finish(E);
}

// views_linears.0 (feature part from the MMA + per-ray part) -> relu -> rgb head
__device__ __forceinline__ void epi_view(const EpiCtx& E, const float* __restrict__ ray_bias, int bias_off,
                                         const float* __restrict__ w_rgb, float (&rgb)[3], __nv_bfloat16* __restrict__ gsave_row) {
    float4 b[16];
    const float4* rb4 = reinterpret_cast<const float4*>(ray_bias + bias_off);
#pragma unroll
    for (int i = 0; i < 16; ++i) b[i] = __ldg(rb4 + i);
    mbar_wait(E.acc_bar, E.phase);
    tc_fence_after();
    if (E.tslot) E.tslot[0] = clock64();
    uint32_t v[2][32];
    tmem_ld32(E.acc, v[0]);
    tmem_ld32(E.acc + 32, v[1]);
    tmem_wait_ld();
    if (E.dslot) E.dslot[0] = clock64();
    float part[3][2] = {{0.f, 0.f}, {0.f, 0.f}, {0.f, 0.f}};
#pragma unroll
    for (int g = 0; g < 2; ++g) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float4 bb = b[g * 8 + i];
            const float a0 = fmaxf(__uint_as_float(v[g][4 * i + 0]) + bb.x, 0.f);
            const float a1 = fmaxf(__uint_as_float(v[g][4 * i + 1]) + bb.y, 0.f);
            const float a2 = fmaxf(__uint_as_float(v[g][4 * i + 2]) + bb.z, 0.f);
            const float a3 = fmaxf(__uint_as_float(v[g][4 * i + 3]) + bb.w, 0.f);
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const float4 w = reinterpret_cast<const float4*>(w_rgb + k * 128)[g * 8 + i];
                part[k][0] = fmaf(a0, w.x, part[k][0]); part[k][1] = fmaf(a1, w.y, part[k][1]);
                part[k][0] = fmaf(a2, w.z, part[k][0]); part[k][1] = fmaf(a3, w.w, part[k][1]);
            }
            if (gsave_row != nullptr)
                reinterpret_cast<uint2*>(gsave_row)[g * 8 + i] = make_uint2(pack_bf16(a0, a1), pack_bf16(a2, a3));
        }
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) rgb[k] += part[k][0] + part[k][1];
    epi_finish(E);
}

// kPair: the two CTAs of a cluster (one TPC) run every MMA as cta_group::2 over a PAIR of row tiles: M = 256 (128 rows
// from each CTA's X / TMEM) and each CTA stages only HALF of every weight stage (64 of the 128 B rows, 8 KB), which
// halves the L2 -> SM weight traffic that bounds the single-CTA kernel.  CTA rank 0 issues all MMAs; its commits
// arrive on the barriers of both CTAs (multicast); the peer CTA forwards "my half landed" / "my activations are in
// TMEM" to the leader's barriers with cluster-scope arrives.
template <bool kFull, bool kPair>
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
    constexpr int kSlots = ring_slots(kFull, kPair);                    // ring slots (16 KB each)
    constexpr int kGroup = stages_per_slot(kPair);                      // stages per slot
    constexpr int kGroupsPerTile = stages_per_tile(kFull) / kGroup;
    static_assert(kGroupsPerTile % (2 * kSlots) == 0, "slot parity must repeat every tile");
    static_assert(kSlots <= kStages, "ring exceeds its shared memory");
    constexpr uint32_t kSlotBytes = kStageBytes;
    constexpr uint32_t kHalfStage = kStageBytes / 2;                     // one CTA's 64 B-rows of a stage (pair mode)
    const uint32_t rank = kPair ? cluster_ctarank() : 0u;                // CTA rank inside the pair
    const bool lead_cta = rank == 0;
    const int unit = kPair ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;   // persistent worker (CTA or CTA pair)
    const int n_units = kPair ? (int)(gridDim.x >> 1) : (int)gridDim.x;
    const int n_work = kPair ? (n_tiles + 1) / 2 : n_tiles;              // row tiles or pairs of row tiles

    for (int i = threadIdx.x; i < kNumHeadFloats; i += kThreads) S.heads[i] = heads[i];
    if (threadIdx.x == 0) {
        const uint32_t n_prod = (kPair && lead_cta) ? 2u : 1u;            // own expect_tx + the peer's forwarded arrive
        for (int s = 0; s < kSlots; ++s) { mbar_init(&S.w_full[s], n_prod); mbar_init(&S.w_empty[s], 1); }
        mbar_init(&S.x_full, n_prod); mbar_init(&S.x_empty, 1);
        mbar_init(&S.acc_full[0], 1); mbar_init(&S.acc_full[1], 1);
        mbar_init(&S.act_ready[0], kPair ? 16 : 8); mbar_init(&S.act_ready[1], kPair ? 16 : 8);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        if (kPair) {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&S.tmem_base)), "r"(512u) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        } else {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&S.tmem_base)), "r"(512u) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
    }
    tc_fence_before();
    __syncthreads();
    if (kPair) cluster_sync_all();                   // both CTAs' barriers initialised before any remote arrive / multicast commit
    tc_fence_after();
    const uint32_t tmem = S.tmem_base;
    const bool tr = (trace != nullptr) && blockIdx.x == 0;
    // trace layout: [tile_iter < 4][role: 0 = mma, 1 = epilogue][(L*2+h)][begin, end]
#define DANBO_TRACE(it, role, L, h, which) do { if (tr && (it) < 4) trace[(((it) * 2 + (role)) * 20 + (L) * 2 + (h)) * 2 + (which)] = clock64(); } while (0)
    // extra epilogue detail after the first 320 entries: [tile_iter < 4][(L*2+h)][tmem loads landed, stores issued]

    if (warp == 0) {
        // ===== producer: X tile + (this CTA's share of the) weight stages; one elected lane issues =====
        const bool leader = elect_one();
        uint32_t ws = 0, wphase = 0, xphase = 0;
        for (int w = unit, it = 0; w < n_work; w += n_units, ++it) {
            int t = kPair ? 2 * w + (int)rank : w;
            if (t >= n_tiles) t = n_tiles - 1;                           // odd tail of a pair: a dummy tile (rows all invalid)
            if (it > 0) { mbar_wait(&S.x_empty, xphase); xphase ^= 1; }
            const uint8_t* xs = xtiles + (size_t)t * kXBytes;
            if (leader) {
                mbar_expect_tx(&S.x_full, kXBytes);
                bulk_g2s(S.x, xs, kXBytes, &S.x_full);
            }
            // pair mode reads the second copy of the stream: [group][rank][2 stages x 8 KB]
            const uint8_t* wsrc = kPair ? wstream + (size_t)84 * kStageBytes + rank * kSlotBytes : wstream;
            for (int g = 0; g < kGroupsPerTile; ++g) {
                mbar_wait(&S.w_empty[ws], wphase ^ 1);
                if (leader) {
                    mbar_expect_tx(&S.w_full[ws], kSlotBytes);
                    bulk_g2s(&S.w[0][0] + ws * kSlotBytes, wsrc + (size_t)g * (kPair ? 2 * kSlotBytes : kSlotBytes), kSlotBytes, &S.w_full[ws]);
                }
                if (++ws == kSlots) { ws = 0; wphase ^= 1; }
            }
        }
    } else if (warp == 1 && kPair && !lead_cta) {
        // ===== peer CTA: forward "my X tile / my half of the stage has landed" to the leader's barriers =====
        uint32_t ws = 0, wphase = 0, xphase = 0;
        const uint32_t x_full_lead = mapa_cluster(smem_u32(&S.x_full), 0);
        const uint32_t w_full_lead = mapa_cluster(smem_u32(&S.w_full[0]), 0);
        for (int w = unit; w < n_work; w += n_units) {
            mbar_wait(&S.x_full, xphase); xphase ^= 1;
            if (lane == 0) mbar_arrive_cluster(x_full_lead);
            for (int g = 0; g < kGroupsPerTile; ++g) {
                mbar_wait(&S.w_full[ws], wphase);
                if (lane == 0) mbar_arrive_cluster(w_full_lead + 8u * ws);
                __syncwarp();
                if (++ws == kSlots) { ws = 0; wphase ^= 1; }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer: the whole warp runs the (uniform) schedule, one elected lane issues tcgen05 =====
        // The schedule of one tile is fully unrolled: ring slot, barrier parity, descriptor offsets and TMEM columns
        // of every stage are compile-time constants added to uniform bases.
        const bool leader = elect_one();
        const uint32_t x_base = smem_u32(S.x);
        const uint32_t w_base = smem_u32(&S.w[0][0]);
        const uint64_t desc_hi = make_desc(0);
        const uint64_t x_desc0 = desc_hi | (uint64_t)((x_base >> 4) & 0x3FFF);
        const uint64_t w_desc0 = desc_hi | (uint64_t)((w_base >> 4) & 0x3FFF);
        constexpr uint32_t idesc = kPair ? kIdescPair : kIdesc;
        // plain (acquire.cta) waits also on the barriers the peer CTA arrives on, as CUTLASS's ClusterBarrier does: what
        // they order is TMEM / async-proxy state, and an acquire.cluster wait adds an L1 invalidate (CCTL.IVALL) per wait
        auto wait_in = [](uint64_t* b, uint32_t ph) { mbar_wait(b, ph); };
        auto commit = [](uint64_t* b) { if (kPair) tc_commit_pair(b); else tc_commit(b); };
        // act_ready[1] completes 9 times per full tile (odd): its parity also depends on the tile iteration
        constexpr int kReady1PerTile = kFull ? 9 : 8;
        for (int w = unit, it = 0; w < n_work; w += n_units, ++it) {
            wait_in(&S.x_full, it & 1);
            const uint32_t r1_base = (uint32_t)(kReady1PerTile * it);
            if (it > 0) {                                                   // accumulators drained by the last epilogues
                wait_in(&S.act_ready[0], 1);                                //   completion #(kLayers-1) of the previous tile
                if (!kFull) wait_in(&S.act_ready[1], (r1_base - 1) & 1);    // density-only ends on a two-half layer
            }
            tc_fence_after();
#pragma unroll
            for (int L = 0; L < kLayers; ++L) {
                if (L > 0) { wait_in(&S.act_ready[0], (L - 1) & 1); tc_fence_after(); }
                const uint32_t act_in = tmem + kActCol + 128u * ((L - 1) & 1);
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    if (h >= n_halves(L)) continue;
                    const uint32_t d = tmem + kAccCol + 128u * h;
                    const int n_xc = uses_x(L) ? 4 : 0;
                    const int n_chunks = n_xc + (uses_act(L) ? 4 : 0);
#pragma unroll
                    for (int c = 0; c < 8; ++c) {
                        if (c >= n_chunks) continue;
                        const bool is_x = c < n_xc;
                        const int kc = is_x ? c : c - n_xc;
                        const int s = stage_index(L, h, c);                 // position in the tile's stage stream
                        const int g = s / kGroup;                           // ring slot use this stage belongs to
                        const int slot = g % kSlots;
                        const uint32_t par = (uint32_t)((g / kSlots) & 1);
                        if (!is_x && kc == 2 && h == 0 && L > 0)            // K columns 128.. come from the other half's epilogue
                            wait_in(&S.act_ready[1], (r1_base + (uint32_t)(L - 1)) & 1);
                        if (s % kGroup == 0) wait_in(&S.w_full[slot], par);
                        tc_fence_after();
                        if (c == 0) { DANBO_TRACE(it, 0, L, h, 0); }
                        const uint64_t bdesc = w_desc0 + (uint64_t)((slot * kSlotBytes + (s % kGroup) * kHalfStage) >> 4);
                        const uint32_t acc0 = c > 0 ? 1u : 0u;
                        if (leader) {
                            if (is_x) {
                                const uint64_t adesc = x_desc0 + (uint64_t)((kc * 16384) >> 4);
                                if (kPair) {
                                    mma_ss_pair(d, adesc, bdesc, idesc, acc0);
                                    if (kc < 3) {                               // K = 208: the last X chunk holds one k-step
                                        mma_ss_pair(d, adesc + 2, bdesc + 2, idesc, 1u);
                                        mma_ss_pair(d, adesc + 4, bdesc + 4, idesc, 1u);
                                        mma_ss_pair(d, adesc + 6, bdesc + 6, idesc, 1u);
                                    }
                                } else {
                                    mma_ss(d, adesc, bdesc, idesc, acc0);
                                    if (kc < 3) {
                                        mma_ss(d, adesc + 2, bdesc + 2, idesc, 1u);
                                        mma_ss(d, adesc + 4, bdesc + 4, idesc, 1u);
                                        mma_ss(d, adesc + 6, bdesc + 6, idesc, 1u);
                                    }
                                }
                            } else {
                                const uint32_t a_t = act_in + kc * 32;
                                if (kPair) {
                                    mma_ts_pair(d, a_t, bdesc, idesc, acc0);
                                    mma_ts_pair(d, a_t + 8, bdesc + 2, idesc, 1u);
                                    mma_ts_pair(d, a_t + 16, bdesc + 4, idesc, 1u);
                                    mma_ts_pair(d, a_t + 24, bdesc + 6, idesc, 1u);
                                } else {
                                    mma_ts(d, a_t, bdesc, idesc, acc0);
                                    mma_ts(d, a_t + 8, bdesc + 2, idesc, 1u);
                                    mma_ts(d, a_t + 16, bdesc + 4, idesc, 1u);
                                    mma_ts(d, a_t + 24, bdesc + 6, idesc, 1u);
                                }
                            }
                            if (s % kGroup == kGroup - 1) commit(&S.w_empty[slot]);
                            if (L == 5 && h == 1 && is_x && kc == 3) commit(&S.x_empty);   // X no longer needed
                            if (c == n_chunks - 1) commit(&S.acc_full[h]);
                        }
                        __syncwarp();
                    }
                    DANBO_TRACE(it, 0, L, h, 1);
                }
            }
        }
    } else {
        // ===== epilogue warps 2..9: two warps per TMEM lane quarter, each owns 64 of the 128 columns of a half =====
        EpiCtx E;
        E.q = warp & 3;                               // TMEM lane quarter this warp may access
        E.ch = (warp - 2) >> 2;                       // which 64-column slice of every half
        const int q = E.q, ch = E.ch;
        const int row = q * 32 + lane;
        E.tmem_lane = tmem + ((uint32_t)(q * 32) << 16);
        E.lane = lane;
        E.remote = kPair;
        const uint32_t ready_addr[2] = {kPair ? mapa_cluster(smem_u32(&S.act_ready[0]), 0) : smem_u32(&S.act_ready[0]),
                                        kPair ? mapa_cluster(smem_u32(&S.act_ready[1]), 0) : smem_u32(&S.act_ready[1])};
        const float* bias = S.heads;                  // [9][256]
        const float* w_alpha = S.heads + 9 * 256;     // [256]
        const float* w_rgb = w_alpha + 256;           // [3][128]
        const float* tail = w_rgb + 3 * 128;          // b_alpha, b_rgb[3]
        uint32_t f0 = 0, f1 = 0;
        for (int w = unit, it = 0; w < n_work; w += n_units, ++it) {
            const int t = kPair ? 2 * w + (int)rank : w;
            const int grow = t * DANBO_TILE_M + row;
            const bool valid = grow < n_rows;                 // (a dummy tile of an odd tail has no valid row)
            const int sample = valid ? row_sample[grow] : -1;
            const int ray = (kFull && valid) ? row_ray[grow] : 0;
            float alpha = ch == 0 ? tail[0] : 0.f;
            float rgb[3] = {ch == 0 ? tail[1] : 0.f, ch == 0 ? tail[2] : 0.f, ch == 0 ? tail[3] : 0.f};
            const bool save = act_save != nullptr && valid;
            for (int L = 0; L < kLayers; ++L) {
                for (int h = 0; h < n_halves(L); ++h) {
                    const int col0 = h * 128 + ch * 64;            // first output column of this thread's slice
                    long long* tslot = (tr && it < 4 && warp == 2 && lane == 0) ? trace + (((it * 2 + 1) * 20 + L * 2 + h) * 2) : nullptr;
                    long long* dslot = tslot ? trace + 320 + ((it * 20 + L * 2 + h) * 2) : nullptr;
                    E.acc = E.tmem_lane + kAccCol + 128u * h + 64u * ch;
                    E.act_out = E.tmem_lane + kActCol + 128u * (L & 1) + 64u * h + 32u * ch;
                    E.acc_bar = &S.acc_full[h];
                    if (h == 0) { E.phase = f0 & 1; ++f0; } else { E.phase = f1 & 1; ++f1; }
                    E.done_addr = ready_addr[h];
                    E.tslot = tslot; E.dslot = dslot;
                    __nv_bfloat16* srow = save ? act_save + ((size_t)L * save_cap + grow) * 256 + col0 : nullptr;
                    if (L < 7) {
                        epi_hidden<0>(E, bias + L * 256 + col0, nullptr, alpha, true, srow);
                    } else if (L == 7) {
                        epi_hidden<1>(E, bias + L * 256 + col0, w_alpha + col0, alpha, kFull, srow);
                    } else if (L == 8) {
                        epi_hidden<2>(E, bias + L * 256 + col0, nullptr, alpha, true, srow);
                    } else {
                        epi_view(E, ray_bias, ray * 128 + col0, w_rgb + col0, rgb,
                                 (g_save != nullptr && valid) ? g_save + (size_t)grow * 128 + col0 : nullptr);
                    }
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
    if (kPair) cluster_sync_all();                   // the peer's TMEM / barriers stay alive until both CTAs are done
    if (warp == 1) {
        if (kPair) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
        else       asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
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
        const __nv_bfloat16 bv = __float2bfloat16_rn(val);
        wstream[(size_t)s * (128 * 64) + off / 2] = bv;
        // second copy for CTA pairs: [group of 2 stages][rank = n / 64][stage in group][64 rows x 64 k] (8 KB pieces)
        const uint32_t half_off = off & 8191u;                           // rows 64..127 start at byte 8192 of a stage
        wstream[(size_t)84 * (128 * 64) + (((size_t)(s >> 1) * 2 + (n >> 6)) * 2 + (s & 1)) * (64 * 64) + half_off / 2] = bv;
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

// The field's output for a sample NO bone sees: blended feature h = 0, so the MLP input is PE(0) for every such sample and
// only the per-ray view bias varies.  The density trunk, sigma and the feature part of the view layer are therefore
// constants of the weights: c[o] = (W_v[:, :256] . feat(PE(0)))[o] and sigma0, computed here in fp32 once per weight
// pack (one block; a warp per output row, coalesced row reads).  danbo_ray_bias then finishes each ray's empty-sample
// output as rgb = W_rgb . relu(c + ray_bias) + b_rgb - 261 121 of the 2.8 M MLP rows of a 512x512 image disappear.
// (nerf.py:176-209 on an all-zero blended feature; danbo.py:299-302.)
__global__ void __launch_bounds__(1024)
empty_trunk_kernel(PackArgs a, float* __restrict__ out /* c[128], sigma0, 3 unused */) {
    __shared__ float buf0[DANBO_X_COLS + 256], buf1[DANBO_X_COLS + 256];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // PE(0) = [0 x 15 | per octave: sin(0) x 15, cos(0) x 15]
    for (int i = threadIdx.x; i < DANBO_X_COLS; i += blockDim.x) {
        const float v = (i >= 15 && ((i - 15) % 30) >= 15) ? 1.f : 0.f;
        buf0[i] = v; buf1[i] = v;                            // both buffers start with x: layer 5 reads [x ; h4]
    }
    __syncthreads();
    // a warp owns up to 8 consecutive output rows and walks all of them together: 8 independent row loads per step keep
    // the single SM's memory pipeline full (one row at a time was a chain of exposed L2 latencies: 140 us for the trunk)
    auto layer = [&](const float* in, int n_in, const float* W, int ld, const float* bias, float* dst, int n_out, bool relu) {
        const int per = (n_out + 31) / 32;                  // rows per warp (<= 8)
        const int o0 = warp * per;
        float s[8];
#pragma unroll
        for (int r = 0; r < 8; ++r) s[r] = 0.f;
        for (int k = lane; k < n_in; k += 32) {
            const float x = in[k];
#pragma unroll
            for (int r = 0; r < 8; ++r)
                if (r < per && o0 + r < n_out) s[r] = fmaf(__ldg(W + (size_t)(o0 + r) * ld + k), x, s[r]);
        }
#pragma unroll
        for (int r = 0; r < 8; ++r) {
#pragma unroll
            for (int m = 16; m > 0; m >>= 1) s[r] += __shfl_xor_sync(0xffffffffu, s[r], m);
            if (lane == 0 && r < per && o0 + r < n_out) {
                const float v = s[r] + (bias ? __ldg(bias + o0 + r) : 0.f);
                dst[o0 + r] = relu ? fmaxf(v, 0.f) : v;
            }
        }
        __syncthreads();
    };
    // activations live behind the x prefix of the two buffers, ping-pong
    float* h[2] = {buf0 + DANBO_X_COLS, buf1 + DANBO_X_COLS};
    layer(buf0, DANBO_X_COLS, a.w[0], DANBO_X_COLS, a.b[0], h[0], 256, true);
    int cur = 0;
    for (int L = 1; L < 8; ++L) {
        if (L == 5) layer((cur ? buf1 : buf0), DANBO_X_COLS + 256, a.w[5], DANBO_X_COLS + 256, a.b[5], h[cur ^ 1], 256, true);
        else layer(h[cur], 256, a.w[L], 256, a.b[L], h[cur ^ 1], 256, true);
        cur ^= 1;
    }
    // sigma0 = w_alpha . h7 + b_alpha ; feat = W_f h7 + b_f ; c = W_v[:, :256] feat
    layer(h[cur], 256, a.w_alpha, 256, a.b_alpha, out + 128, 1, false);
    layer(h[cur], 256, a.w_feat, 256, a.b_feat, h[cur ^ 1], 256, false);
    layer(h[cur ^ 1], 256, a.w_view, 411, nullptr, out, 128, false);
}

// One warp per ray: raw_tail[ray] = [W_rgb . relu(c + ray_bias[ray]) + b_rgb, sigma0].
__global__ void __launch_bounds__(256)
empty_rows_kernel(const float* __restrict__ ray_bias, int n_rays, const float* __restrict__ heads, float4* __restrict__ raw_tail) {
    const int lane = threadIdx.x & 31;
    const float* ec = heads + kEmptyOff;
    const float4 c = *reinterpret_cast<const float4*>(ec + lane * 4);
    const float* wr = heads + 10 * 256;                     // W_rgb (3,128), then b_alpha, b_rgb[3]
    float4 w[3];
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) w[ch] = *reinterpret_cast<const float4*>(wr + ch * 128 + lane * 4);
    const float sigma0 = ec[128], b0 = wr[385], b1 = wr[386], b2 = wr[387];
    const int warps = (gridDim.x * blockDim.x) >> 5;
    for (int ray = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; ray < n_rays; ray += warps) {
        const float4 b = __ldg(reinterpret_cast<const float4*>(ray_bias + (size_t)ray * 128) + lane);
        const float g0 = fmaxf(c.x + b.x, 0.f), g1 = fmaxf(c.y + b.y, 0.f), g2 = fmaxf(c.z + b.z, 0.f), g3 = fmaxf(c.w + b.w, 0.f);
        float r[3];
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) r[ch] = w[ch].x * g0 + w[ch].y * g1 + w[ch].z * g2 + w[ch].w * g3;
#pragma unroll
        for (int m = 16; m > 0; m >>= 1) {
            r[0] += __shfl_xor_sync(0xffffffffu, r[0], m); r[1] += __shfl_xor_sync(0xffffffffu, r[1], m);
            r[2] += __shfl_xor_sync(0xffffffffu, r[2], m);
        }
        if (lane == 0) raw_tail[ray] = make_float4(r[0] + b0, r[1] + b1, r[2] + b2, sigma0);
    }
}

}  // namespace mlp
}  // namespace danbo

using namespace danbo;

extern "C" int danbo_mlp_empty_rows(const float* ray_bias, int n_rays, const float* heads, float* raw_tail, int num_sms,
                                    void* stream) {
    if (n_rays <= 0) return 0;
    if (!ray_bias || !heads || !raw_tail || (reinterpret_cast<uintptr_t>(raw_tail) & 15)) return -1;
    int blocks = (n_rays + 7) / 8;
    if (blocks > num_sms * 8) blocks = num_sms * 8;
    mlp::empty_rows_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(ray_bias, n_rays, heads, (float4*)raw_tail);
    DANBO_CHECK_LAUNCH();
    return 0;
}

extern "C" int danbo_mlp_workspace_bytes(long long* wstream_bytes, long long* heads_bytes, long long* raybias_bytes) {
    *wstream_bytes = 2 * 84LL * mlp::kStageBytes;        // single-CTA order, then the CTA-pair order
    *heads_bytes = (long long)(mlp::kEmptyOff + mlp::kEmptyFloats) * 4;     // + the empty-sample constants (empty_trunk_kernel)
    *raybias_bytes = (long long)mlp::kRayBiasFloats * 4;
    return 0;
}

static int g_pack_empty = 1;      // 1: danbo_pack_mlp_weights also computes the empty-sample constants (one-block trunk, ~80 us)

// Training packs the weights every iteration and never reads the empty-sample constants (its per-ray empty rows go
// through the MLP so that their gradient does): the caller switches the constants off for those packs.  -> previous.
extern "C" int danbo_mlp_set_pack_empty(int enable) {
    const int old = g_pack_empty;
    g_pack_empty = enable ? 1 : 0;
    return old;
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
    if (g_pack_empty) {
        mlp::empty_trunk_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(a, heads + mlp::kEmptyOff);
        DANBO_CHECK_LAUNCH();
    }
    return 0;
}

static int g_cta_pair = 1;        // 1: cta_group::2 kernel over CTA pairs (default), 0: one CTA per row tile

extern "C" int danbo_mlp_set_cta_pair(int enable) {
    const int old = g_cta_pair;
    g_cta_pair = enable ? 1 : 0;
    return old;
}

template <bool kFull, bool kPair>
static int mlp_launch_variant(int grid, int smem, cudaStream_t stream, const uint8_t* xtiles, const uint8_t* wstream,
                              const float* heads, const float* ray_bias, const int* row_sample, const int* row_ray,
                              const int* n_rows_dev, float* out, int out_capacity, __nv_bfloat16* act_save,
                              __nv_bfloat16* g_save, int save_cap, long long* trace) {
    static bool attr_set = false;             // idempotent launch attribute, set on first use per variant
    auto kern = mlp::mlp_kernel<kFull, kPair>;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return (int)e;
        attr_set = true;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(mlp::kThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = kPair ? 2 : 1; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, kern, xtiles, wstream, heads, ray_bias, row_sample, row_ray, n_rows_dev, out,
                                       out_capacity, act_save, g_save, save_cap, trace);
    if (e != cudaSuccess) return (int)e;
    DANBO_CHECK_LAUNCH();
    return 0;
}

static int mlp_launch(const void* xtiles, const void* wstream, const float* heads, const float* ray_bias,
                      const int* row_sample, const int* row_ray, const int* n_rows_dev, int max_rows,
                      float* out, int out_capacity, int density_only, int num_sms, void* act_save, void* g_save,
                      int save_cap, long long* trace, void* stream) {
    if (max_rows <= 0) return 0;
    const int smem = (int)sizeof(mlp::Smem) + 1024;
    const int max_tiles = (max_rows + DANBO_TILE_M - 1) / DANBO_TILE_M;
    const bool pair = g_cta_pair != 0 && num_sms >= 2;
    int grid;
    if (pair) {
        const int max_pairs = (max_tiles + 1) / 2, sm_pairs = num_sms / 2;
        grid = 2 * (sm_pairs < max_pairs ? sm_pairs : max_pairs);
    } else {
        grid = num_sms < max_tiles ? num_sms : max_tiles;
    }
    if (grid < 1) grid = 1;
#define DANBO_MLP_ARGS grid, smem, (cudaStream_t)stream, (const uint8_t*)xtiles, (const uint8_t*)wstream, heads, ray_bias, \
                       row_sample, row_ray, n_rows_dev, out, out_capacity, (__nv_bfloat16*)act_save, (__nv_bfloat16*)g_save, save_cap, trace
    if (density_only) return pair ? mlp_launch_variant<false, true>(DANBO_MLP_ARGS) : mlp_launch_variant<false, false>(DANBO_MLP_ARGS);
    return pair ? mlp_launch_variant<true, true>(DANBO_MLP_ARGS) : mlp_launch_variant<true, false>(DANBO_MLP_ARGS);
#undef DANBO_MLP_ARGS
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

// Same launch, and CTA 0 writes a clock64 timeline of its first 4 tiles into trace[480] (profiling aid).
extern "C" int danbo_mlp_forward_trace(const void* xtiles, const void* wstream, const float* heads, const float* ray_bias,
                                       const int* row_sample, const int* row_ray, const int* n_rows_dev, int max_rows,
                                       float* out, int out_capacity, int density_only, int num_sms, long long* trace,
                                       void* stream) {
    return mlp_launch(xtiles, wstream, heads, ray_bias, row_sample, row_ray, n_rows_dev, max_rows, out, out_capacity,
                      density_only, num_sms, nullptr, nullptr, 0, trace, stream);
}
