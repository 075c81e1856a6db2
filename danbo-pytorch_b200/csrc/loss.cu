// L*: the trainer's losses on the outputs of the path, value + gradients in ONE launch.
//
// core/trainer.py:396-422 (_compute_nerf_loss: L1 / MSE on rgb + (1 - acc) bg, fine and coarse), :507-536
// (_compute_soft_softmax_loss on confd / part_invalid / T_i / alpha) and :538-553 (_compute_volume_scale_loss).
// As PyTorch ops these are ~60 elementwise / reduction launches forward and as many backward over the (n,S_t,24)
// confd tensors; a training iteration at BASELINE config #3 is ~1.8 ms, so they were a tenth of it.  Every term is a
// mean (or a 72-element sum), so its gradient is known in closed form while the value is accumulated:
//   rgb:   pred = rgb + (1 - acc) bg, d = pred - target;  L1: mean|d|, g_rgb = c sign(d)/(3n);  MSE: mean d^2, g_rgb = 2 c d/(3n)
//          g_acc = -sum_c g_rgb_c bg_c
//   soft:  r = [T_i alpha > 0] - sum_j valid_j (1.002 s(a_j) - 0.001);  loss = c mean r^2
//          g_a_j = -2 c r/(n S_t) valid_j 1.002 s(a_j)(1 - s(a_j))
//   vol:   s = max(|axis_scale|, 0.05 init);  loss = c sum_j s_j0 s_j1 s_j2;  g = c prod_{b != a} s_b sign(axis_scale_a) [|.| >= min]
// Term sums are accumulated in fp64 (one atomic per block), so the value does not depend on the block order beyond
// fp64 rounding.
#include "common.cuh"
#include <math.h>

namespace danbo {

constexpr int kLossBlock = 256;

__device__ __forceinline__ double block_sum(double v, double* sh) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) sh[warp] = v;
    __syncthreads();
    double t = 0.0;
    if (warp == 0) {
        t = lane < (kLossBlock >> 5) ? sh[lane] : 0.0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    }
    __syncthreads();
    return t;                                            // valid in warp 0
}

struct RgbTerm { const float* rgb; const float* acc; float* g_rgb; float* g_acc; float coef; };

__global__ void __launch_bounds__(kLossBlock)
train_loss_kernel(RgbTerm fine, RgbTerm coarse, const float* __restrict__ target, const float* __restrict__ bgs,
                  float bg_scalar, int use_bg, int n_rays, int loss_kind,
                  const float* __restrict__ confd, const float* __restrict__ part_invalid, const float* __restrict__ T_i,
                  const float* __restrict__ alpha, int S_t, float soft_coef, float* __restrict__ g_confd,
                  const float* __restrict__ axis_scale, const float* __restrict__ init_scale, float vol_coef,
                  float* __restrict__ g_axis_scale, double* __restrict__ terms) {
    __shared__ double sh[kLossBlock >> 5];
    const long long idx = (long long)blockIdx.x * kLossBlock + threadIdx.x;
    // ---- rgb terms: one thread per ray
    double s_f = 0.0, s_c = 0.0;
    if (idx < n_rays) {
        float bg[3], tg[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            bg[c] = use_bg ? (bgs ? bgs[idx * 3 + c] : bg_scalar) : 0.f;
            tg[c] = target[idx * 3 + c];
        }
        const float inv = 1.f / (3.f * (float)n_rays);
#pragma unroll
        for (int which = 0; which < 2; ++which) {
            const RgbTerm& t = which ? coarse : fine;
            if (!t.rgb) continue;
            const float keep = use_bg ? 1.f - t.acc[idx] : 0.f;
            float ga = 0.f;
            double s = 0.0;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const float d = t.rgb[idx * 3 + c] + keep * bg[c] - tg[c];
                float g;
                if (loss_kind == 0) { s += (double)fabsf(d); g = d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f); }
                else { s += (double)d * d; g = 2.f * d; }
                g *= t.coef * inv;
                t.g_rgb[idx * 3 + c] = g;
                ga -= g * bg[c];
            }
            t.g_acc[idx] = ga;
            if (which) s_c = s * t.coef; else s_f = s * t.coef;
        }
    }
    // ---- soft-softmax term: one thread per (ray, sample) row of 24 logits
    double s_s = 0.0;
    const long long rows = confd ? (long long)n_rays * S_t : 0;
    if (idx < rows) {
        const float4* a4 = reinterpret_cast<const float4*>(confd + idx * DANBO_J);
        const float4* i4 = reinterpret_cast<const float4*>(part_invalid + idx * DANBO_J);
        float sg[DANBO_J], vl[DANBO_J];
        float sum_p = 0.f;
#pragma unroll
        for (int q = 0; q < DANBO_J / 4; ++q) {
            const float4 a = a4[q], iv = i4[q];
            const float av[4] = {a.x, a.y, a.z, a.w}, ivv[4] = {iv.x, iv.y, iv.z, iv.w};
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const float s = 1.f / (1.f + expf(-av[u]));
                const float v = 1.f - ivv[u];
                sg[4 * q + u] = s; vl[4 * q + u] = v;
                sum_p += (s * 1.002f - 0.001f) * v;
            }
        }
        const float label = (T_i[idx] * alpha[idx]) > 0.f ? 1.f : 0.f;
        const float r = label - sum_p;
        s_s = (double)r * r;
        const float k = -2.f * soft_coef * r / (float)rows * 1.002f;
        float4* g4 = reinterpret_cast<float4*>(g_confd + idx * DANBO_J);
#pragma unroll
        for (int q = 0; q < DANBO_J / 4; ++q) {
            float gv[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) { const float s = sg[4 * q + u]; gv[u] = k * vl[4 * q + u] * s * (1.f - s); }
            g4[q] = make_float4(gv[0], gv[1], gv[2], gv[3]);
        }
    }
    // ---- volume-scale penalty: 24 threads of block 0
    double s_v = 0.0;
    if (axis_scale && blockIdx.x == 0 && threadIdx.x < DANBO_J) {
        const int j = threadIdx.x;
        float s[3], raw[3];
        bool free_[3];
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            raw[a] = axis_scale[j * 3 + a];
            const float lo = init_scale[j * 3 + a] * 0.05f;
            free_[a] = fabsf(raw[a]) >= lo;                                  // clamp(min=lo) passes the gradient at and above lo
            s[a] = fmaxf(fabsf(raw[a]), lo);
        }
        s_v = (double)(s[0] * s[1] * s[2]) * vol_coef;
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            const float others = a == 0 ? s[1] * s[2] : (a == 1 ? s[0] * s[2] : s[0] * s[1]);
            const float sign = raw[a] > 0.f ? 1.f : (raw[a] < 0.f ? -1.f : 0.f);
            if (free_[a]) g_axis_scale[j * 3 + a] += vol_coef * others * sign;
        }
    }
    const double inv3n = 1.0 / (3.0 * (double)n_rays);
    double t;
    t = block_sum(s_f, sh); if (threadIdx.x == 0 && t != 0.0) atomicAdd(terms + 0, t * inv3n);
    t = block_sum(s_c, sh); if (threadIdx.x == 0 && t != 0.0) atomicAdd(terms + 1, t * inv3n);
    t = block_sum(s_s, sh); if (threadIdx.x == 0 && t != 0.0) atomicAdd(terms + 2, t * (double)soft_coef / (double)(rows > 0 ? rows : 1));
    t = block_sum(s_v, sh); if (threadIdx.x == 0 && t != 0.0) atomicAdd(terms + 3, t);
}

}  // namespace danbo

using namespace danbo;

extern "C" int danbo_train_loss(const float* rgb_map, const float* acc_map, const float* rgb0, const float* acc0,
                                const float* target, const float* bgs, float bg_scalar, int use_background, int n_rays,
                                int loss_kind, float rgb_coef, float coarse_weight, const float* confd,
                                const float* part_invalid, const float* T_i, const float* alpha, int S_t, float soft_coef,
                                const float* axis_scale, const float* init_scale, float vol_coef, double* terms,
                                float* g_rgb_map, float* g_acc_map, float* g_rgb0, float* g_acc0, float* g_confd,
                                float* g_axis_scale, void* stream) {
    if (n_rays <= 0) return 0;
    if (!rgb_map || !acc_map || !target || !terms || !g_rgb_map || !g_acc_map) return -1;
    if (rgb0 && (!acc0 || !g_rgb0 || !g_acc0)) return -1;
    if (confd && (!part_invalid || !T_i || !alpha || !g_confd || S_t <= 0)) return -1;
    if (axis_scale && (!init_scale || !g_axis_scale)) return -1;
    if (loss_kind < 0 || loss_kind > 1) return -2;
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e = cudaMemsetAsync(terms, 0, 4 * sizeof(double), st);
    if (e != cudaSuccess) return (int)e;
    const long long work = confd ? (long long)n_rays * S_t : (long long)n_rays;
    const int G = (int)((work + kLossBlock - 1) / kLossBlock);
    RgbTerm fine{rgb_map, acc_map, g_rgb_map, g_acc_map, rgb_coef};
    RgbTerm coarse{rgb0, acc0, g_rgb0, g_acc0, rgb_coef * coarse_weight};
    train_loss_kernel<<<G, kLossBlock, 0, st>>>(fine, coarse, target, bgs, bg_scalar, use_background, n_rays, loss_kind, confd,
                                                 part_invalid, T_i, alpha, S_t, soft_coef, g_confd, axis_scale, init_scale,
                                                 vol_coef, g_axis_scale, terms);
    DANBO_CHECK_LAUNCH();
    return 0;
}
