// C1 alpha compositing, R1 importance resampling + sorted merge order, R2 merge of coarse and fine samples.
// Reference: core/networks/nerf.py:281-347 (raw2outputs), core/utils/ray_utils.py:159-203,257-291
// (sample_pdf / isample_from_lineseg), core/raycasters.py:484-514,745-761 (_merge_encodings / merge_samples).
//
// One warp per ray; every lane owns a contiguous run of samples, products and prefix sums are done as
// lane-local sequential work plus one warp scan.  cumprod / cumsum accumulate in fp64 like ATen's CPU kernels
// (acc_type<float>) and round each output to fp32.  HBM bound: 16 B raw + 4 B z in, 8 B (weight, alpha) out
// per sample, 20 B per ray.
#include "common.cuh"
#include <math.h>

namespace danbo {

constexpr int kMaxS = 160;          // longest ray: 96 + 48 = 144 samples
constexpr int kEPL = 5;             // elements per lane at kMaxS
constexpr int kWarpsPerBlock = 4;

struct RaySmem {
    float z[kMaxS];
    float w[kMaxS];
    float cdf[kMaxS];
    float cat[kMaxS];
    float zs[kMaxS];
    int order[kMaxS];
};

__device__ __forceinline__ double warp_excl_prod(double v, int lane) {
    double inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const double u = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc *= u; }
    const double ex = __shfl_up_sync(0xffffffffu, inc, 1);
    return lane == 0 ? 1.0 : ex;
}
__device__ __forceinline__ double warp_excl_sum(double v, int lane) {
    double inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const double u = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += u; }
    const double ex = __shfl_up_sync(0xffffffffu, inc, 1);
    return lane == 0 ? 0.0 : ex;
}

// Composites one ray whose S samples are described by fetch(i) -> raw (rgb, sigma).  sm.z must hold z[0..S).
// Writes weights / alpha rows and the per-ray maps; leaves the weights in sm.w.
// kE = samples per lane = ceil(S / 32), a template parameter so that the per-lane loops have no dead iterations (the
// generic 5-element loops spent 4 of 5 iterations on predicated-off work at S = 32).
template <int kE, class Fetch>
__device__ __forceinline__ void composite_ray(RaySmem& sm, int S, int lane, float norm_d, float inv_B,
                                              const float* __restrict__ noise_row, Fetch fetch,
                                              float* __restrict__ w_row, float* __restrict__ alpha_row,
                                              float* __restrict__ rgb_out, float* __restrict__ disp_out,
                                              float* __restrict__ acc_out) {
    constexpr int epl = kE;
    float al[kE], cr[kE], cg[kE], cb[kE];
    double lp = 1.0;
#pragma unroll
    for (int e = 0; e < kE; ++e) {
        const int i = lane * epl + e;
        al[e] = 0.f; cr[e] = cg[e] = cb[e] = 0.f;
        if (i < S) {
            const float4 raw = fetch(i);
            float d = (i + 1 < S) ? __fsub_rn(sm.z[i + 1], sm.z[i]) : 1e10f;
            d = __fmul_rn(d, norm_d);
            cr[e] = (1.f / (1.f + expf(-raw.x))) * 1.002f - 0.001f;
            cg[e] = (1.f / (1.f + expf(-raw.y))) * 1.002f - 0.001f;
            cb[e] = (1.f / (1.f + expf(-raw.z))) * 1.002f - 0.001f;
            float sg = __fmul_rn(raw.w, inv_B);
            if (noise_row) sg = __fadd_rn(sg, noise_row[i]);
            sg = fmaxf(sg, 0.f);
            al[e] = __fsub_rn(1.f, expf(-__fmul_rn(sg, d)));
            lp *= (double)__fadd_rn(__fsub_rn(1.f, al[e]), 1e-10f);
        }
    }
    double T = warp_excl_prod(lp, lane);
    float sr = 0.f, sg_ = 0.f, sb = 0.f, sd = 0.f, sw = 0.f;
#pragma unroll
    for (int e = 0; e < kE; ++e) {
        const int i = lane * epl + e;
        if (i < S) {
            const float w = __fmul_rn(al[e], (float)T);
            T *= (double)__fadd_rn(__fsub_rn(1.f, al[e]), 1e-10f);
            sm.w[i] = w;
            if (w_row) w_row[i] = w;
            if (alpha_row) alpha_row[i] = al[e];
            sr = fmaf(w, cr[e], sr); sg_ = fmaf(w, cg[e], sg_); sb = fmaf(w, cb[e], sb);
            sd = fmaf(w, sm.z[i], sd); sw += w;
        }
    }
    sr = warp_sum(sr); sg_ = warp_sum(sg_); sb = warp_sum(sb); sd = warp_sum(sd); sw = warp_sum(sw);
    if (lane == 0) {
        rgb_out[0] = sr; rgb_out[1] = sg_; rgb_out[2] = sb;
        float disp = 1.f / fmaxf(1e-10f, sd / (sw + 1e-10f));
        if (fabsf(sw) <= 1e-8f) disp = 0.f;                         // torch.isclose(sum w, 0) -> disparity 0
        *disp_out = disp;
        *acc_out = fminf(sw, 1.f);
    }
    __syncwarp();
}

// Coarse pass: composite + inverse-CDF importance sampling + merge order.
template <int kE>
__global__ void __launch_bounds__(32 * kWarpsPerBlock)
composite_resample_kernel(const float* __restrict__ rays, int ray_stride, int n_rays, int S, int S_f,
                          const float* __restrict__ raw /* (n_rays*S + n_rays, 4); tail = per-ray empty sample */,
                          const uint32_t* __restrict__ mask, const float* __restrict__ z,
                          const float* __restrict__ noise, float inv_B, const float* __restrict__ u_vals /* (S_f) or null */,
                          const float* __restrict__ u_rand /* (n_rays,S_f) or null */,
                          float* __restrict__ weights, float* __restrict__ alpha, float* __restrict__ rgb0,
                          float* __restrict__ disp0, float* __restrict__ acc0, float* __restrict__ z_samples,
                          float* __restrict__ z_all, int* __restrict__ order_out, int* __restrict__ inds_out,
                          int smooth /* 1: single_net weights (is_only, ray_utils.py:272-279); 0: the interior weights */) {
    __shared__ RaySmem smem[kWarpsPerBlock];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int n = blockIdx.x * kWarpsPerBlock + wib;
    if (n >= n_rays) return;
    RaySmem& sm = smem[wib];
    const float* r = rays + (size_t)n * ray_stride;
    const float norm_d = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(r[3], r[3]), __fmul_rn(r[4], r[4])), __fmul_rn(r[5], r[5])));
    for (int i = lane; i < S; i += 32) sm.z[i] = z[(size_t)n * S + i];
    __syncwarp();
    const float4* raw4 = reinterpret_cast<const float4*>(raw);
    const float4 empty = raw4[(size_t)n_rays * S + n];
    auto fetch = [&](int i) { return mask[(size_t)n * S + i] ? raw4[(size_t)n * S + i] : empty; };
    composite_ray<kE>(sm, S, lane, norm_d, inv_B, noise ? noise + (size_t)n * S : nullptr, fetch,
                      weights + (size_t)n * S, alpha + (size_t)n * S, rgb0 + (size_t)n * 3, disp0 + n, acc0 + n);
    if (S_f <= 0) return;
    // ---- R1: pdf over the S-2 interior bins, cdf with a leading zero (ray_utils.py:161-164,272-279)
    const int nb = S - 2;
    const int epl = (nb + 31) / 32;                     // <= kE
    float dw[kE];
    double lsum = 0.0;
#pragma unroll
    for (int e = 0; e < kE; ++e) {
        const int k = lane * epl + e;
        dw[e] = 0.f;
        if (e < epl && k < nb) {
            if (smooth) {
                const float a = fmaxf(sm.w[k], sm.w[k + 1]), b = fmaxf(sm.w[k + 1], sm.w[k + 2]);
                dw[e] = __fadd_rn(__fadd_rn(__fmul_rn(0.5f, __fadd_rn(a, b)), 0.01f), 1e-5f);
            } else {
                dw[e] = __fadd_rn(sm.w[k + 1], 1e-5f);                     // weights[..., 1:-1] + 1e-5 (sample_pdf)
            }
            lsum += (double)dw[e];
        }
    }
    double tot = lsum;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, o);
    const float total = (float)tot;
    double lpdf = 0.0;
#pragma unroll
    for (int e = 0; e < kE; ++e) {
        const int k = lane * epl + e;
        if (e < epl && k < nb) { dw[e] = __fdiv_rn(dw[e], total); lpdf += (double)dw[e]; }
    }
    double run = warp_excl_sum(lpdf, lane);
    if (lane == 0) sm.cdf[0] = 0.f;
#pragma unroll
    for (int e = 0; e < kE; ++e) {
        const int k = lane * epl + e;
        if (e < epl && k < nb) { run += (double)dw[e]; sm.cdf[k + 1] = (float)run; }
    }
    __syncwarp();
    const int n_cdf = S - 1;
    for (int j = lane; j < S_f; j += 32) {
        const float u = u_rand ? u_rand[(size_t)n * S_f + j] : u_vals[j];
        // searchsorted(cdf, u, right=True) = #{k : cdf[k] <= u}; the cdf is non-decreasing (a running sum of
        // non-negative terms), so the count is the first index with cdf > u: binary search
        int lo = 0, hi = n_cdf;
        while (lo < hi) { const int mid = (lo + hi) >> 1; if (sm.cdf[mid] <= u) lo = mid + 1; else hi = mid; }
        const int ind = lo;
        const int below = ind - 1 > 0 ? ind - 1 : 0;
        const int above = ind < n_cdf - 1 ? ind : n_cdf - 1;
        const float cb = sm.cdf[below], ca = sm.cdf[above];
        const float bb = __fmul_rn(.5f, __fadd_rn(sm.z[below + 1], sm.z[below]));
        const float ba = __fmul_rn(.5f, __fadd_rn(sm.z[above + 1], sm.z[above]));
        float den = __fsub_rn(ca, cb);
        if (den < 1e-5f) den = 1.f;
        const float t = __fdiv_rn(__fsub_rn(u, cb), den);
        const float zsv = __fadd_rn(bb, __fmul_rn(t, __fsub_rn(ba, bb)));
        sm.zs[j] = zsv;
        z_samples[(size_t)n * S_f + j] = zsv;
        if (inds_out) inds_out[(size_t)n * S_f + j] = ind;
    }
    __syncwarp();
    // ---- sorted merge of [z ; z_samples] (stable: ties keep concatenation order)
    const int St = S + S_f;
    for (int i = lane; i < St; i += 32) sm.cat[i] = i < S ? sm.z[i] : sm.zs[i - S];
    __syncwarp();
    // rank of element i in the stable sort = #{k : cat[k] < cat[i] or (cat[k] == cat[i] and k < i)}.  When both halves are
    // already non-decreasing (always for z; for z_samples whenever u is sorted, i.e. eval) this is
    //   coarse i: i + #{z_samples < z_i}        fine j: S + j' ... = #{z <= zs_j} + j
    // by two binary searches; otherwise (train mode, random u) the general count.  Checked per ray, so exact either way.
    bool sorted = true;
    for (int i = lane; i < St; i += 32)
        if (i + 1 < St && i + 1 != S) sorted = sorted && !(sm.cat[i + 1] < sm.cat[i]);
    sorted = __all_sync(0xffffffffu, sorted);
    if (sorted) {
        for (int i = lane; i < St; i += 32) {
            const float v = sm.cat[i];
            int rank;
            if (i < S) {
                int lo = 0, hi = S_f;                               // first fine index with zs >= v
                while (lo < hi) { const int mid = (lo + hi) >> 1; if (sm.zs[mid] < v) lo = mid + 1; else hi = mid; }
                rank = i + lo;
            } else {
                int lo = 0, hi = S;                                 // first coarse index with z > v
                while (lo < hi) { const int mid = (lo + hi) >> 1; if (sm.z[mid] <= v) lo = mid + 1; else hi = mid; }
                rank = lo + (i - S);
            }
            sm.order[rank] = i;
        }
    } else {
        for (int i = lane; i < St; i += 32) {
            const float v = sm.cat[i];
            int rank = 0;
            for (int k = 0; k < St; ++k) { const float c = sm.cat[k]; rank += (c < v || (c == v && k < i)) ? 1 : 0; }
            sm.order[rank] = i;
        }
    }
    __syncwarp();
    for (int i = lane; i < St; i += 32) {
        const int src = sm.order[i];
        order_out[(size_t)n * St + i] = src;
        z_all[(size_t)n * St + i] = sm.cat[src];
    }
}

// Fine pass: gather coarse / fine raw by merge order, composite the S_t samples.
template <int kE>
__global__ void __launch_bounds__(32 * kWarpsPerBlock)
merge_composite_kernel(const float* __restrict__ rays, int ray_stride, int n_rays, int S_c, int S_f,
                       const float* __restrict__ raw0 /* (n_rays*S_c + n_rays, 4) */, const uint32_t* __restrict__ mask0,
                       const float* __restrict__ raw1 /* (n_rays*S_f, 4) */, const uint32_t* __restrict__ mask1,
                       const float* __restrict__ z_all, const int* __restrict__ order, const float* __restrict__ noise,
                       float inv_B, float* __restrict__ weights, float* __restrict__ alpha, float* __restrict__ rgb,
                       float* __restrict__ disp, float* __restrict__ acc, float* __restrict__ raw_merged /* (n,St,4) or null */,
                       const float* __restrict__ confd0, const float* __restrict__ confd1,
                       float* __restrict__ confd_merged /* (n,St,24) or null */, float* __restrict__ invalid_merged /* or null */) {
    __shared__ RaySmem smem[kWarpsPerBlock];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int n = blockIdx.x * kWarpsPerBlock + wib;
    if (n >= n_rays) return;
    RaySmem& sm = smem[wib];
    const int St = S_c + S_f;
    const float* r = rays + (size_t)n * ray_stride;
    const float norm_d = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(r[3], r[3]), __fmul_rn(r[4], r[4])), __fmul_rn(r[5], r[5])));
    for (int i = lane; i < St; i += 32) { sm.z[i] = z_all[(size_t)n * St + i]; sm.order[i] = order[(size_t)n * St + i]; }
    __syncwarp();
    const float4* r0 = reinterpret_cast<const float4*>(raw0);
    const float4* r1 = reinterpret_cast<const float4*>(raw1);
    const float4 empty = r0[(size_t)n_rays * S_c + n];
    auto fetch = [&](int i) {
        const int src = sm.order[i];
        float4 v;
        if (src < S_c) v = mask0[(size_t)n * S_c + src] ? r0[(size_t)n * S_c + src] : empty;
        else           v = mask1[(size_t)n * S_f + src - S_c] ? r1[(size_t)n * S_f + src - S_c] : empty;
        if (raw_merged) reinterpret_cast<float4*>(raw_merged)[(size_t)n * St + i] = v;
        return v;
    };
    composite_ray<kE>(sm, St, lane, norm_d, inv_B, noise ? noise + (size_t)n * St : nullptr, fetch,
                      weights + (size_t)n * St, alpha + (size_t)n * St, rgb + (size_t)n * 3, disp + n, acc + n);
    if (confd_merged || invalid_merged) {                          // training outputs (raycasters.py:710-716)
        for (int e = lane; e < St * DANBO_J; e += 32) {
            const int i = e / DANBO_J, j = e - i * DANBO_J;
            const int src = sm.order[i];
            const bool coarse = src < S_c;
            const size_t sidx = coarse ? (size_t)n * S_c + src : (size_t)n * S_f + src - S_c;
            const uint32_t m = coarse ? mask0[sidx] : mask1[sidx];
            const bool vis = (m >> j) & 1u;
            if (invalid_merged) invalid_merged[(size_t)n * St * DANBO_J + e] = vis ? 0.f : 1.f;
            if (confd_merged) confd_merged[(size_t)n * St * DANBO_J + e] = vis ? (coarse ? confd0 : confd1)[sidx * DANBO_J + j] : 0.f;
        }
    }
}

}  // namespace danbo

using namespace danbo;

extern "C" int danbo_composite_resample(const float* rays, int ray_stride, int n_rays, int S, int S_f, const float* raw,
                                        const unsigned int* mask, const float* z, const float* noise, float inv_B,
                                        const float* u_vals, const float* u_rand, float* weights, float* alpha,
                                        float* rgb0, float* disp0, float* acc0, float* z_samples, float* z_all,
                                        int* order, int* inds, int smooth_weights, void* stream) {
    if (n_rays <= 0) return 0;
    if (S < 3 || S > kMaxS || S + S_f > kMaxS) return -1;
    if (S_f > 0 && !u_vals && !u_rand) return -2;
    const int G = (n_rays + kWarpsPerBlock - 1) / kWarpsPerBlock;
#define DANBO_CR_LAUNCH(E) composite_resample_kernel<E><<<G, 32 * kWarpsPerBlock, 0, (cudaStream_t)stream>>>( \
        rays, ray_stride, n_rays, S, S_f, raw, mask, z, noise, inv_B, u_vals, u_rand, weights, alpha, rgb0, disp0, acc0, \
        z_samples, z_all, order, inds, smooth_weights)
    switch ((S + 31) / 32) {                                 // samples per lane: the kernel is specialised on it
        case 1: DANBO_CR_LAUNCH(1); break;
        case 2: DANBO_CR_LAUNCH(2); break;
        case 3: DANBO_CR_LAUNCH(3); break;
        case 4: DANBO_CR_LAUNCH(4); break;
        default: DANBO_CR_LAUNCH(5); break;
    }
#undef DANBO_CR_LAUNCH
    DANBO_CHECK_LAUNCH();
    return 0;
}

extern "C" int danbo_merge_composite(const float* rays, int ray_stride, int n_rays, int S_c, int S_f, const float* raw0,
                                     const unsigned int* mask0, const float* raw1, const unsigned int* mask1,
                                     const float* z_all, const int* order, const float* noise, float inv_B,
                                     float* weights, float* alpha, float* rgb, float* disp, float* acc,
                                     float* raw_merged, const float* confd0, const float* confd1, float* confd_merged,
                                     float* invalid_merged, void* stream) {
    if (n_rays <= 0) return 0;
    if (S_c + S_f > kMaxS) return -1;
    const int G = (n_rays + kWarpsPerBlock - 1) / kWarpsPerBlock;
#define DANBO_MC_LAUNCH(E) merge_composite_kernel<E><<<G, 32 * kWarpsPerBlock, 0, (cudaStream_t)stream>>>( \
        rays, ray_stride, n_rays, S_c, S_f, raw0, mask0, raw1, mask1, z_all, order, noise, inv_B, weights, alpha, rgb, \
        disp, acc, raw_merged, confd0, confd1, confd_merged, invalid_merged)
    switch ((S_c + S_f + 31) / 32) {
        case 1: DANBO_MC_LAUNCH(1); break;
        case 2: DANBO_MC_LAUNCH(2); break;
        case 3: DANBO_MC_LAUNCH(3); break;
        case 4: DANBO_MC_LAUNCH(4); break;
        default: DANBO_MC_LAUNCH(5); break;
    }
#undef DANBO_MC_LAUNCH
    DANBO_CHECK_LAUNCH();
    return 0;
}

// =========================================================================================================
// Backward of C1 (+ R2): gradients of the rendered maps w.r.t. raw.  Autograd of nerf.py:297-347 as it is reached
// from trainer.py:405-409: sources are rgb_map and acc_map (disp_map, alpha and weights carry no gradient), the
// importance samples are detached (ray_utils.py:287), so each ray is independent:
//   w_i = a_i T_i, T_i = prod_{k<i}(1 - a_k + 1e-10)
//   dL/dw_i   = g_rgb . c_i + g_acc [sum w < 1]
//   dL/da_i   = dw_i T_i - (sum_{k>i} dw_k w_k) / (1 - a_i + 1e-10)
//   dL/dsig_i = da_i * dist_i * (1 - a_i) for sigma_i > 0,   dL/dc_i = w_i g_rgb,  c = 1.002 sigmoid(raw) - 0.001
// One warp per ray, forward values recomputed from raw (nothing saved).  Gradients of samples no bone sees are
// summed into the ray's empty entry.
namespace danbo {

template <class Fetch, class Emit>
__device__ __forceinline__ void composite_ray_bwd(RaySmem& sm, int S, int lane, float norm_d, float inv_B,
                                                  const float* __restrict__ noise_row, Fetch fetch, float gr, float gg,
                                                  float gb, float gacc, Emit emit) {
    const int epl = (S + 31) / 32;
    float al[kEPL], sgm[kEPL], dist[kEPL], cr[kEPL], cg[kEPL], cb[kEPL];
    double lp = 1.0;
#pragma unroll
    for (int e = 0; e < kEPL; ++e) {
        const int i = lane * epl + e;
        al[e] = 0.f; sgm[e] = 0.f; dist[e] = 0.f; cr[e] = cg[e] = cb[e] = 0.f;
        if (e < epl && i < S) {
            const float4 raw = fetch(i);
            float d = (i + 1 < S) ? __fsub_rn(sm.z[i + 1], sm.z[i]) : 1e10f;
            d = __fmul_rn(d, norm_d);
            dist[e] = d;
            cr[e] = 1.f / (1.f + expf(-raw.x)); cg[e] = 1.f / (1.f + expf(-raw.y)); cb[e] = 1.f / (1.f + expf(-raw.z));
            float sg = __fmul_rn(raw.w, inv_B);
            if (noise_row) sg = __fadd_rn(sg, noise_row[i]);
            sgm[e] = sg;
            al[e] = __fsub_rn(1.f, expf(-__fmul_rn(fmaxf(sg, 0.f), d)));
            lp *= (double)__fadd_rn(__fsub_rn(1.f, al[e]), 1e-10f);
        }
    }
    double T = warp_excl_prod(lp, lane);
    float w[kEPL], Tf[kEPL], dw[kEPL];
    float wsum = 0.f;
#pragma unroll
    for (int e = 0; e < kEPL; ++e) {
        const int i = lane * epl + e;
        w[e] = 0.f; Tf[e] = 0.f;
        if (e < epl && i < S) {
            Tf[e] = (float)T;
            w[e] = al[e] * Tf[e];
            T *= (double)__fadd_rn(__fsub_rn(1.f, al[e]), 1e-10f);
            wsum += w[e];
        }
    }
    wsum = warp_sum(wsum);
    const float ga = (wsum < 1.f) ? gacc : 0.f;                 // acc = min(sum w, 1)
    // suffix sums of dw_k w_k: lane-local reverse pass + warp suffix scan
    float loc = 0.f;
#pragma unroll
    for (int e = 0; e < kEPL; ++e) {
        const float c0 = cr[e] * 1.002f - 0.001f, c1 = cg[e] * 1.002f - 0.001f, c2 = cb[e] * 1.002f - 0.001f;
        dw[e] = gr * c0 + gg * c1 + gb * c2 + ga;
        loc += dw[e] * w[e];
    }
    float inc = loc;                                            // inclusive suffix over lanes
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const float u = __shfl_down_sync(0xffffffffu, inc, o); if (lane + o < 32) inc += u; }
    float suffix = inc - loc;                                   // sum over lanes > this lane
#pragma unroll
    for (int e = kEPL - 1; e >= 0; --e) {
        const int i = lane * epl + e;
        if (e < epl && i < S) {
            const float om = __fadd_rn(__fsub_rn(1.f, al[e]), 1e-10f);
            const float da = dw[e] * Tf[e] - suffix / om;
            suffix += dw[e] * w[e];
            float4 g;
            g.x = w[e] * gr * 1.002f * cr[e] * (1.f - cr[e]);
            g.y = w[e] * gg * 1.002f * cg[e] * (1.f - cg[e]);
            g.z = w[e] * gb * 1.002f * cb[e] * (1.f - cb[e]);
            g.w = sgm[e] > 0.f ? da * dist[e] * (1.f - al[e]) * inv_B : 0.f;
            emit(i, g);
        }
    }
}

// d raw0 (+)= backward of the coarse composite; d raw0 is (n*S + n, 4), zero-initialised by the caller.
__global__ void __launch_bounds__(32 * kWarpsPerBlock)
composite_bwd_kernel(const float* __restrict__ rays, int ray_stride, int n_rays, int S, const float* __restrict__ raw,
                     const uint32_t* __restrict__ mask, const float* __restrict__ z, const float* __restrict__ noise,
                     float inv_B, const float* __restrict__ g_rgb, const float* __restrict__ g_acc,
                     float* __restrict__ d_raw) {
    __shared__ RaySmem smem[kWarpsPerBlock];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int n = blockIdx.x * kWarpsPerBlock + wib;
    if (n >= n_rays) return;
    RaySmem& sm = smem[wib];
    const float* r = rays + (size_t)n * ray_stride;
    const float norm_d = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(r[3], r[3]), __fmul_rn(r[4], r[4])), __fmul_rn(r[5], r[5])));
    for (int i = lane; i < S; i += 32) sm.z[i] = z[(size_t)n * S + i];
    __syncwarp();
    const float4* raw4 = reinterpret_cast<const float4*>(raw);
    const float4 empty = raw4[(size_t)n_rays * S + n];
    auto fetch = [&](int i) { return mask[(size_t)n * S + i] ? raw4[(size_t)n * S + i] : empty; };
    float4 ge = make_float4(0.f, 0.f, 0.f, 0.f);
    auto emit = [&](int i, float4 g) {
        if (mask[(size_t)n * S + i]) {
            float4* d = reinterpret_cast<float4*>(d_raw) + (size_t)n * S + i;
            float4 o = *d; o.x += g.x; o.y += g.y; o.z += g.z; o.w += g.w; *d = o;
        } else { ge.x += g.x; ge.y += g.y; ge.z += g.z; ge.w += g.w; }
    };
    composite_ray_bwd(sm, S, lane, norm_d, inv_B, noise ? noise + (size_t)n * S : nullptr, fetch,
                      g_rgb[n * 3], g_rgb[n * 3 + 1], g_rgb[n * 3 + 2], g_acc[n], emit);
    ge.x = warp_sum(ge.x); ge.y = warp_sum(ge.y); ge.z = warp_sum(ge.z); ge.w = warp_sum(ge.w);
    if (lane == 0) {
        float4* d = reinterpret_cast<float4*>(d_raw) + (size_t)n_rays * S + n;
        float4 o = *d; o.x += ge.x; o.y += ge.y; o.z += ge.z; o.w += ge.w; *d = o;
    }
}

// backward of the merged composite: scatters to d raw0 (coarse + empty entries) and d raw1 (fine); also routes the
// gradient of the merged blend logits (soft-softmax loss, trainer.py:507-536) back to the two sparse logit buffers.
__global__ void __launch_bounds__(32 * kWarpsPerBlock)
merge_composite_bwd_kernel(const float* __restrict__ rays, int ray_stride, int n_rays, int S_c, int S_f,
                           const float* __restrict__ raw0, const uint32_t* __restrict__ mask0,
                           const float* __restrict__ raw1, const uint32_t* __restrict__ mask1,
                           const float* __restrict__ z_all, const int* __restrict__ order,
                           const float* __restrict__ noise, float inv_B, const float* __restrict__ g_rgb,
                           const float* __restrict__ g_acc, const float* __restrict__ g_confd /* (n,St,24) or null */,
                           float* __restrict__ d_raw0, float* __restrict__ d_raw1,
                           float* __restrict__ d_logit0 /* (n*S_c,24) */, float* __restrict__ d_logit1 /* (n*S_f,24) */) {
    __shared__ RaySmem smem[kWarpsPerBlock];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int n = blockIdx.x * kWarpsPerBlock + wib;
    if (n >= n_rays) return;
    RaySmem& sm = smem[wib];
    const int St = S_c + S_f;
    const float* r = rays + (size_t)n * ray_stride;
    const float norm_d = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(r[3], r[3]), __fmul_rn(r[4], r[4])), __fmul_rn(r[5], r[5])));
    for (int i = lane; i < St; i += 32) { sm.z[i] = z_all[(size_t)n * St + i]; sm.order[i] = order[(size_t)n * St + i]; }
    __syncwarp();
    const float4* r0 = reinterpret_cast<const float4*>(raw0);
    const float4* r1 = reinterpret_cast<const float4*>(raw1);
    const float4 empty = r0[(size_t)n_rays * S_c + n];
    auto fetch = [&](int i) {
        const int src = sm.order[i];
        if (src < S_c) return mask0[(size_t)n * S_c + src] ? r0[(size_t)n * S_c + src] : empty;
        return mask1[(size_t)n * S_f + src - S_c] ? r1[(size_t)n * S_f + src - S_c] : empty;
    };
    float4 ge = make_float4(0.f, 0.f, 0.f, 0.f);
    auto emit = [&](int i, float4 g) {
        const int src = sm.order[i];
        float4* d = nullptr;
        if (src < S_c) { if (mask0[(size_t)n * S_c + src]) d = reinterpret_cast<float4*>(d_raw0) + (size_t)n * S_c + src; }
        else if (mask1[(size_t)n * S_f + src - S_c]) d = reinterpret_cast<float4*>(d_raw1) + (size_t)n * S_f + src - S_c;
        if (d) { float4 o = *d; o.x += g.x; o.y += g.y; o.z += g.z; o.w += g.w; *d = o; }
        else { ge.x += g.x; ge.y += g.y; ge.z += g.z; ge.w += g.w; }
    };
    composite_ray_bwd(sm, St, lane, norm_d, inv_B, noise ? noise + (size_t)n * St : nullptr, fetch,
                      g_rgb[n * 3], g_rgb[n * 3 + 1], g_rgb[n * 3 + 2], g_acc[n], emit);
    ge.x = warp_sum(ge.x); ge.y = warp_sum(ge.y); ge.z = warp_sum(ge.z); ge.w = warp_sum(ge.w);
    if (lane == 0) {
        float4* d = reinterpret_cast<float4*>(d_raw0) + (size_t)n_rays * S_c + n;
        float4 o = *d; o.x += ge.x; o.y += ge.y; o.z += ge.z; o.w += ge.w; *d = o;
    }
    if (g_confd) {
        for (int e = lane; e < St * DANBO_J; e += 32) {
            const int i = e / DANBO_J, j = e - i * DANBO_J;
            const int src = sm.order[i];
            const bool coarse = src < S_c;
            const size_t sidx = coarse ? (size_t)n * S_c + src : (size_t)n * S_f + src - S_c;
            const uint32_t m = coarse ? mask0[sidx] : mask1[sidx];
            if ((m >> j) & 1u) (coarse ? d_logit0 : d_logit1)[sidx * DANBO_J + j] = g_confd[(size_t)n * St * DANBO_J + e];
        }
    }
}

}  // namespace danbo

extern "C" int danbo_composite_bwd(const float* rays, int ray_stride, int n_rays, int S, const float* raw,
                                   const unsigned int* mask, const float* z, const float* noise, float inv_B,
                                   const float* g_rgb, const float* g_acc, float* d_raw, void* stream) {
    if (n_rays <= 0) return 0;
    if (S > danbo::kMaxS) return -1;
    const int G = (n_rays + danbo::kWarpsPerBlock - 1) / danbo::kWarpsPerBlock;
    danbo::composite_bwd_kernel<<<G, 32 * danbo::kWarpsPerBlock, 0, (cudaStream_t)stream>>>(
        rays, ray_stride, n_rays, S, raw, mask, z, noise, inv_B, g_rgb, g_acc, d_raw);
    DANBO_CHECK_LAUNCH();
    return 0;
}

extern "C" int danbo_merge_composite_bwd(const float* rays, int ray_stride, int n_rays, int S_c, int S_f,
                                         const float* raw0, const unsigned int* mask0, const float* raw1,
                                         const unsigned int* mask1, const float* z_all, const int* order,
                                         const float* noise, float inv_B, const float* g_rgb, const float* g_acc,
                                         const float* g_confd, float* d_raw0, float* d_raw1, float* d_logit0,
                                         float* d_logit1, void* stream) {
    if (n_rays <= 0) return 0;
    if (S_c + S_f > danbo::kMaxS) return -1;
    const int G = (n_rays + danbo::kWarpsPerBlock - 1) / danbo::kWarpsPerBlock;
    danbo::merge_composite_bwd_kernel<<<G, 32 * danbo::kWarpsPerBlock, 0, (cudaStream_t)stream>>>(
        rays, ray_stride, n_rays, S_c, S_f, raw0, mask0, raw1, mask1, z_all, order, noise, inv_B, g_rgb, g_acc, g_confd,
        d_raw0, d_raw1, d_logit0, d_logit1);
    DANBO_CHECK_LAUNCH();
    return 0;
}
