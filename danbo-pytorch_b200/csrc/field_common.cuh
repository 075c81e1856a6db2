// Device helpers shared by the forward (field.cu) and backward (backward_field.cu) field kernels.
#pragma once
#include "common.cuh"
#include <math.h>

namespace danbo {

struct FieldConsts {
    const float* align;       // (24,4,4) bone-align transforms A_j                     raycasters.py:548-591
    const float* axis_scale;  // (24,3)  per-bone half extents (|.| applied here)       gnn_backbone.py:802
    const float* agg_w0;      // (24,15,32) prob_linears.layers.0.lin.weight
    const float* agg_adjw;    // (24,24) prob_linears.layers.0.adj_w
    const float* agg_adj;     // (24,24) prob_linears.layers.0.adj (tree + self, 0/1)
    const float* agg_b0;      // (32)
    const float* agg_w1;      // (24,32,32)
    const float* agg_b1;      // (24,32)
    const float* agg_w2;      // (24,32)
    const float* agg_b2;      // (24)
};

// tree neighbours (self, parent, children) of every SMPL joint as bit masks (gnn_backbone.py:18-34)
static __constant__ uint32_t kNbrMask[DANBO_J] = {
    0x0000000Fu, 0x00000013u, 0x00000025u, 0x00000049u, 0x00000092u, 0x00000124u, 0x00000248u, 0x00000490u,
    0x00000920u, 0x00007240u, 0x00000480u, 0x00000900u, 0x00009200u, 0x00012200u, 0x00024200u, 0x00009000u,
    0x00052000u, 0x000A4000u, 0x00150000u, 0x002A0000u, 0x00540000u, 0x00A80000u, 0x00500000u, 0x00A00000u};

// ---------------------------------------------------------------------------------------------------------
// x_j = (A_j (R_j p + t_j) + a_j) / |s_j| for one joint, in the reference's two-step order with a true divide.
// encoders.py:288-303 (transform_batch_pts), :442-444 (bone align), gnn_backbone.py:802 (scale)
__device__ __forceinline__ void bone_aligned(const float* __restrict__ skt, const float* __restrict__ A,
                                             float px, float py, float pz, float& t0, float& t1, float& t2) {
    const float4 r0 = __ldg(reinterpret_cast<const float4*>(skt));
    const float4 r1 = __ldg(reinterpret_cast<const float4*>(skt) + 1);
    const float4 r2 = __ldg(reinterpret_cast<const float4*>(skt) + 2);
    const float l0 = fmaf(r0.z, pz, fmaf(r0.y, py, fmaf(r0.x, px, r0.w)));
    const float l1 = fmaf(r1.z, pz, fmaf(r1.y, py, fmaf(r1.x, px, r1.w)));
    const float l2 = fmaf(r2.z, pz, fmaf(r2.y, py, fmaf(r2.x, px, r2.w)));
    const float4 a0 = __ldg(reinterpret_cast<const float4*>(A));
    const float4 a1 = __ldg(reinterpret_cast<const float4*>(A) + 1);
    const float4 a2 = __ldg(reinterpret_cast<const float4*>(A) + 2);
    t0 = __fadd_rn(fmaf(a0.z, l2, fmaf(a0.y, l1, __fmul_rn(a0.x, l0))), a0.w);
    t1 = __fadd_rn(fmaf(a1.z, l2, fmaf(a1.y, l1, __fmul_rn(a1.x, l0))), a1.w);
    t2 = __fadd_rn(fmaf(a2.z, l2, fmaf(a2.y, l1, __fmul_rn(a2.x, l0))), a2.w);
}

__device__ __forceinline__ void bone_coords(const float* __restrict__ skt, const float* __restrict__ A,
                                            const float* __restrict__ scale, float px, float py, float pz,
                                            float& x0, float& x1, float& x2) {
    const float4 r0 = __ldg(reinterpret_cast<const float4*>(skt));
    const float4 r1 = __ldg(reinterpret_cast<const float4*>(skt) + 1);
    const float4 r2 = __ldg(reinterpret_cast<const float4*>(skt) + 2);
    const float l0 = fmaf(r0.z, pz, fmaf(r0.y, py, fmaf(r0.x, px, r0.w)));
    const float l1 = fmaf(r1.z, pz, fmaf(r1.y, py, fmaf(r1.x, px, r1.w)));
    const float l2 = fmaf(r2.z, pz, fmaf(r2.y, py, fmaf(r2.x, px, r2.w)));
    const float4 a0 = __ldg(reinterpret_cast<const float4*>(A));
    const float4 a1 = __ldg(reinterpret_cast<const float4*>(A) + 1);
    const float4 a2 = __ldg(reinterpret_cast<const float4*>(A) + 2);
    const float t0 = __fadd_rn(fmaf(a0.z, l2, fmaf(a0.y, l1, __fmul_rn(a0.x, l0))), a0.w);
    const float t1 = __fadd_rn(fmaf(a1.z, l2, fmaf(a1.y, l1, __fmul_rn(a1.x, l0))), a1.w);
    const float t2 = __fadd_rn(fmaf(a2.z, l2, fmaf(a2.y, l1, __fmul_rn(a2.x, l0))), a2.w);
    x0 = __fdiv_rn(t0, fabsf(__ldg(scale + 0)));
    x1 = __fdiv_rn(t1, fabsf(__ldg(scale + 1)));
    x2 = __fdiv_rn(t2, fabsf(__ldg(scale + 2)));
}

#define DANBO_PAIR_OVERFLOW_WORD 48   // work[48] != 0: pair_bucket saw more visible pairs than pair_capacity
struct PairWork {            // int workspace: [0,24) count per bone, [24,48) scatter cursor, [48] overflow flag, [64, 64+cap) pairs
    int* base;
    __device__ __forceinline__ int* count() const { return base; }
    __device__ __forceinline__ int* cursor() const { return base + 24; }
    __device__ __forceinline__ int* pairs() const { return base + 64; }
};

__device__ __forceinline__ int seg_start(const int* __restrict__ count, int j) {
    int off = 0;
    for (int i = 0; i < j; ++i) off += (count[i] + 31) & ~31;
    return off;
}

// features of bone k at local coordinates x (closed form of misc.py:331-351 + window, gnn_backbone.py:802-826).
// kShared: vol_k points into shared memory (plain loads) instead of global memory (read-only cache loads).
template <bool kShared = false>
__device__ __forceinline__ void bone_features(const float* __restrict__ vol_k, float x0, float x1, float x2, float (&h)[DANBO_FEAT]) {
    const float a2 = x0 * x0, b2 = x1 * x1, c2 = x2 * x2;
    const float win = expf(-2.f * (a2 * a2 * a2 + b2 * b2 * b2 + c2 * c2 * c2));
    const float xs[3] = {x0, x1, x2};
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        const float iy = ((xs[a] + 1.f) * (float)DANBO_RES - 1.f) * 0.5f;
        const float fl = floorf(iy);
        const float w1 = iy - fl, w0 = 1.f - w1;
        const int i0 = (int)fl, i1 = i0 + 1;
        const bool ok0 = i0 >= 0 && i0 < DANBO_RES, ok1 = i1 >= 0 && i1 < DANBO_RES;
#pragma unroll
        for (int f = 0; f < 5; ++f) {
            const float* line = vol_k + f * (DANBO_RES * 3) + a;
            const float v0 = ok0 ? (kShared ? line[i0 * 3] : __ldg(line + i0 * 3)) : 0.f;
            const float v1 = ok1 ? (kShared ? line[i1 * 3] : __ldg(line + i1 * 3)) : 0.f;
            h[f * 3 + a] = (v0 * w0 + v1 * w1) * win;
        }
    }
}

// A2, agg_type = softmax with mask_vol_prob (danbo.py:388-404): p_j = v_j exp(a_j - M) / max(sum_k (v_k exp(a_k - M) + eps), eps)
// with M the max over ALL 24 logits (invisible bones included) and eps = 1e-7.  Returns M and 1 / denominator.
constexpr float kSoftmaxEps = 1e-7f;
__device__ __forceinline__ void softmax_terms(const float* __restrict__ a, uint32_t visible, float& amax, float& inv_den,
                                              int* argmax = nullptr) {
    float m = a[0];
    int am = 0;
#pragma unroll
    for (int k = 1; k < DANBO_J; ++k) { const float v = a[k]; if (v > m) { m = v; am = k; } }     // first maximum, like torch.max
    float den = 0.f;
#pragma unroll
    for (int k = 0; k < DANBO_J; ++k) den += (((visible >> k) & 1u) ? expf(a[k] - m) : 0.f) + kSoftmaxEps;
    amax = m;
    inv_den = 1.f / fmaxf(den, kSoftmaxEps);
    if (argmax) *argmax = am;
}

}  // namespace danbo
