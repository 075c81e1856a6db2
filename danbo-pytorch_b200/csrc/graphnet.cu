// GN1 + GN2: the per-pose skeleton graph net (pose -> 24 x 240 bone feature lines), forward and backward.
//
// Reference: encode_graph_inputs core/encoders.py:460-473, AxisAngtoRot6DEncoder :859-877 (axis-angle -> quaternion ->
// matrix, first two columns), Embedder core/cutoff_embedder.py:62-73 (5 octaves), BodyGNN.forward
// core/networks/gnn_backbone.py:683-704 on FactorizeGNN: DensePNGCN :225-274 (per-joint linear, then the learned
// 24x24 mix adj_w * adj, shared bias), ParallelLinear core/networks/misc.py:174-183.
//     n0 = mask_root * w            lin0[j] = n0[j] W0[j]     pre0 = A0 lin0 + b0      n1 = relu(2 pre0)   (the x2: SURVEY F3)
//     lin1[j] = n1[j] W1[j]         pre1 = A1 lin1 + b1       n2 = relu(pre1)
//     n3 = relu(n2[j] W2[j] + b2[j])                          out[j] = n3[j] W3[j] + b3[j]
// The work is tiny (<= 16 poses x 24 nodes, 28 MFLOP) but as PyTorch ops it was ~50 launches forward and ~100
// backward of a 1.8 ms training iteration.  Here: one block per (node, 16-pose group), a thread per output column with
// one accumulator per pose, so every weight is read once per block with coalesced loads; a launch per mix (the mix reads
// every node's output of the previous launch): 3 launches forward, 3 backward.  fp32 throughout.
#include "common.cuh"
#include <math.h>

namespace danbo {
namespace gn {

constexpr int kG = 16;             // poses per block
constexpr int kW = 128;            // node width
constexpr int kIn = 66;            // 6 * (1 + 2 * 5)
constexpr int kOut = DANBO_VOL;    // 240
constexpr int kThreads = 256;

struct Params {
    const float* w0; const float* adjw0; const float* adj0; const float* b0;     // (24,66,128) (24,24) (24,24) (128)
    const float* w1; const float* adjw1; const float* adj1; const float* b1;     // (24,128,128) ... (128)
    const float* w2; const float* b2;                                            // (24,128,128) (24,128)
    const float* w3; const float* b3;                                            // (24,128,240) (24,240)
};
struct Saved {                     // activations kept for the backward pass, all (G,24,*)
    float* w_in; float* lin0; float* n1; float* lin1; float* n2; float* n3;
};
struct Grads {                     // accumulated into (+=)
    float* w0; float* adjw0; float* b0; float* w1; float* adjw1; float* b1; float* w2; float* b2; float* w3; float* b3;
};

// out[g][c] = sum_i in_s[g][i] W[i][c] for this thread's column c; W row-major (n_in, ld), in_s in shared memory
template <int kPitch>
__device__ __forceinline__ void colmm(const float* __restrict__ W, int n_in, int ld, int c, const float* in_s, float (&acc)[kG]) {
#pragma unroll 4
    for (int i = 0; i < n_in; ++i) {
        const float w = __ldg(W + (size_t)i * ld + c);
#pragma unroll
        for (int g = 0; g < kG; ++g) acc[g] = fmaf(in_s[g * kPitch + i], w, acc[g]);
    }
}

// mixed[g][c] = sum_k A[j][k] src[g][k][c]   (A = adj_w * adj; zero entries skipped)
__device__ __forceinline__ void mix_in(const float* __restrict__ adjw, const float* __restrict__ adj, int j,
                                       const float* __restrict__ src, int g0, int ng, int c, float (&acc)[kG]) {
    for (int k = 0; k < DANBO_J; ++k) {
        const float m = __ldg(adj + j * DANBO_J + k);
        if (m == 0.f) continue;
        const float a = __ldg(adjw + j * DANBO_J + k) * m;
#pragma unroll
        for (int g = 0; g < kG; ++g)
            if (g < ng) acc[g] = fmaf(a, src[((size_t)(g0 + g) * DANBO_J + k) * kW + c], acc[g]);
    }
}

// ---- forward ----------------------------------------------------------------------------------------------------
// F0: graph inputs (GN1) + layer-0 per-joint linear
__global__ void __launch_bounds__(kThreads)
fwd0_kernel(const float* __restrict__ bones /* (G,24,3) */, int G, Params P, Saved S) {
    __shared__ float in_s[kG * (kIn + 1)];
    const int j = blockIdx.x, g0 = blockIdx.y * kG, ng = min(kG, G - g0);
    for (int e = threadIdx.x; e < kG * (kIn + 1); e += kThreads) in_s[e] = 0.f;
    __syncthreads();
    if (threadIdx.x < ng) {
        const int g = g0 + threadIdx.x;
        const float* aa = bones + ((size_t)g * DANBO_J + j) * 3;
        // axis-angle -> quaternion -> matrix (pytorch3d's documented route, skeleton_utils.py:411-418)
        const float ax = aa[0], ay = aa[1], az = aa[2];
        const float ang = sqrtf(ax * ax + ay * ay + az * az), half = 0.5f * ang;
        const float k = fabsf(ang) < 1e-6f ? 0.5f - ang * ang / 48.f : sinf(half) / ang;
        const float qr = cosf(half), qi = ax * k, qj = ay * k, qk = az * k;
        const float two_s = 2.f / (qr * qr + qi * qi + qj * qj + qk * qk);
        const float r6[6] = {1.f - two_s * (qj * qj + qk * qk), two_s * (qi * qj - qk * qr),       // R00 R01
                             two_s * (qi * qj + qk * qr), 1.f - two_s * (qi * qi + qk * qk),       // R10 R11
                             two_s * (qi * qk - qj * qr), two_s * (qj * qk + qi * qr)};            // R20 R21
        const float root = j == 0 ? 0.f : 1.f;                                                     // mask_root
        float* row = in_s + threadIdx.x * (kIn + 1);
        float* keep = S.w_in + ((size_t)g * DANBO_J + j) * kIn;
#pragma unroll
        for (int i = 0; i < 6; ++i) {
            float v[11];
            v[0] = r6[i];
#pragma unroll
            for (int f = 0; f < 5; ++f) { const float x = r6[i] * (float)(1 << f); v[1 + 2 * f] = sinf(x); v[2 + 2 * f] = cosf(x); }
#pragma unroll
            for (int q = 0; q < 11; ++q) {
                const int col = q == 0 ? i : 6 + 6 * (q - 1) + i;             // [x, sin 2^0, cos 2^0, sin 2^1, ...] blocks of 6
                row[col] = v[q] * root;
                keep[col] = v[q] * root;
            }
        }
    }
    __syncthreads();
    const int c = threadIdx.x;
    if (c < kW) {
        float acc[kG];
#pragma unroll
        for (int g = 0; g < kG; ++g) acc[g] = 0.f;
        colmm<kIn + 1>(P.w0 + (size_t)j * kIn * kW, kIn, kW, c, in_s, acc);
#pragma unroll
        for (int g = 0; g < kG; ++g)
            if (g < ng) S.lin0[((size_t)(g0 + g) * DANBO_J + j) * kW + c] = acc[g];
    }
}

// F1: mix 0 + bias + relu(2 x) -> n1; layer-1 per-joint linear -> lin1
__global__ void __launch_bounds__(kThreads)
fwd1_kernel(int G, Params P, Saved S) {
    __shared__ float in_s[kG * (kW + 1)];
    const int j = blockIdx.x, g0 = blockIdx.y * kG, ng = min(kG, G - g0);
    const int c = threadIdx.x;
    if (c < kW) {
        float acc[kG];
        const float b = __ldg(P.b0 + c);
#pragma unroll
        for (int g = 0; g < kG; ++g) acc[g] = 0.f;
        mix_in(P.adjw0, P.adj0, j, S.lin0, g0, ng, c, acc);
#pragma unroll
        for (int g = 0; g < kG; ++g) {
            const float pre = acc[g] + b;
            const float v = g < ng ? fmaxf(pre + pre, 0.f) : 0.f;
            in_s[g * (kW + 1) + c] = v;
            if (g < ng) S.n1[((size_t)(g0 + g) * DANBO_J + j) * kW + c] = v;
        }
    }
    __syncthreads();
    if (c < kW) {
        float acc[kG];
#pragma unroll
        for (int g = 0; g < kG; ++g) acc[g] = 0.f;
        colmm<kW + 1>(P.w1 + (size_t)j * kW * kW, kW, kW, c, in_s, acc);
#pragma unroll
        for (int g = 0; g < kG; ++g)
            if (g < ng) S.lin1[((size_t)(g0 + g) * DANBO_J + j) * kW + c] = acc[g];
    }
}

// F2: mix 1 + bias + relu -> n2; layer 2 (+ per-joint bias, relu) -> n3; layer 3 -> out
__global__ void __launch_bounds__(kThreads)
fwd2_kernel(int G, Params P, Saved S, float* __restrict__ out /* (G,24,240) */) {
    __shared__ float a_s[kG * (kW + 1)];
    __shared__ float b_s[kG * (kW + 1)];
    const int j = blockIdx.x, g0 = blockIdx.y * kG, ng = min(kG, G - g0);
    const int c = threadIdx.x;
    if (c < kW) {
        float acc[kG];
        const float b = __ldg(P.b1 + c);
#pragma unroll
        for (int g = 0; g < kG; ++g) acc[g] = 0.f;
        mix_in(P.adjw1, P.adj1, j, S.lin1, g0, ng, c, acc);
#pragma unroll
        for (int g = 0; g < kG; ++g) {
            const float v = g < ng ? fmaxf(acc[g] + b, 0.f) : 0.f;
            a_s[g * (kW + 1) + c] = v;
            if (g < ng) S.n2[((size_t)(g0 + g) * DANBO_J + j) * kW + c] = v;
        }
    }
    __syncthreads();
    if (c < kW) {
        float acc[kG];
        const float b = __ldg(P.b2 + j * kW + c);
#pragma unroll
        for (int g = 0; g < kG; ++g) acc[g] = b;
        colmm<kW + 1>(P.w2 + (size_t)j * kW * kW, kW, kW, c, a_s, acc);
#pragma unroll
        for (int g = 0; g < kG; ++g) {
            const float v = g < ng ? fmaxf(acc[g], 0.f) : 0.f;
            b_s[g * (kW + 1) + c] = v;
            if (g < ng) S.n3[((size_t)(g0 + g) * DANBO_J + j) * kW + c] = v;
        }
    }
    __syncthreads();
    if (c < kOut) {
        float acc[kG];
        const float b = __ldg(P.b3 + j * kOut + c);
#pragma unroll
        for (int g = 0; g < kG; ++g) acc[g] = b;
        colmm<kW + 1>(P.w3 + (size_t)j * kW * kOut, kW, kOut, c, b_s, acc);
#pragma unroll
        for (int g = 0; g < kG; ++g)
            if (g < ng) out[((size_t)(g0 + g) * DANBO_J + j) * kOut + c] = acc[g];
    }
}

// ---- backward ---------------------------------------------------------------------------------------------------
// dW[i][c] += sum_g a_s[g][i] d_s[g][c]: one owner thread per entry when the grid has one pose group, atomics otherwise
template <int kPa, int kPd>
__device__ __forceinline__ void outer_acc(float* __restrict__ dW, int n_in, int n_out, const float* a_s, const float* d_s,
                                          bool exclusive) {
    for (int e = threadIdx.x; e < n_in * n_out; e += kThreads) {
        const int i = e / n_out, c = e - i * n_out;
        float s = 0.f;
#pragma unroll
        for (int g = 0; g < kG; ++g) s = fmaf(a_s[g * kPa + i], d_s[g * kPd + c], s);
        if (exclusive) dW[e] += s; else atomicAdd(dW + e, s);
    }
}

// d_in[g][i] = sum_c d_s[g][c] W[i][c] for i < n_in (n_in <= 128): W is staged through shared memory in [128 x 32]
// column tiles so that global reads are whole lines and each thread then walks its own row
template <int kPd>
__device__ __forceinline__ void rowmm(const float* __restrict__ W, int n_in, int n_out, const float* d_s, float* tile /* [128*33] */,
                                      float (&acc)[kG]) {
    const int i = threadIdx.x;
#pragma unroll
    for (int g = 0; g < kG; ++g) acc[g] = 0.f;
    for (int c0 = 0; c0 < n_out; c0 += 32) {
        const int nc = min(32, n_out - c0);
        __syncthreads();
        for (int e = threadIdx.x; e < n_in * 32; e += kThreads) {
            const int r = e >> 5, cc = e & 31;
            tile[r * 33 + cc] = cc < nc ? __ldg(W + (size_t)r * n_out + c0 + cc) : 0.f;
        }
        __syncthreads();
        if (i < n_in) {
#pragma unroll 8
            for (int cc = 0; cc < 32; ++cc) {
                const float w = tile[i * 33 + cc];
                if (c0 + cc < n_out) {
#pragma unroll
                    for (int g = 0; g < kG; ++g) acc[g] = fmaf(d_s[g * kPd + c0 + cc], w, acc[g]);
                }
            }
        }
    }
    __syncthreads();
}

__device__ __forceinline__ void bias_acc(float* __restrict__ db, int n, const float* d_s, int pitch, bool exclusive) {
    for (int c = threadIdx.x; c < n; c += kThreads) {
        float s = 0.f;
#pragma unroll
        for (int g = 0; g < kG; ++g) s += d_s[g * pitch + c];
        if (exclusive) db[c] += s; else atomicAdd(db + c, s);
    }
}

// B2: layers 3 and 2 of node j: d out -> dW3, db3, d n3 -> d pre2 -> dW2, db2, d n2 -> d pre1 (written), db1 (atomics)
__global__ void __launch_bounds__(kThreads)
bwd2_kernel(int G, Params P, Saved S, const float* __restrict__ d_out, Grads Gr, float* __restrict__ d_pre1) {
    __shared__ float d_s[kG * (kOut + 1)];          // d out, later d pre2 (pitch kW + 1)
    __shared__ float a_s[kG * (kW + 1)];            // n3, later n2
    __shared__ float tile[128 * 33];
    const int j = blockIdx.x, g0 = blockIdx.y * kG, ng = min(kG, G - g0);
    const bool excl = gridDim.y == 1;
    for (int e = threadIdx.x; e < kG * kOut; e += kThreads) {
        const int g = e / kOut, c = e - g * kOut;
        d_s[g * (kOut + 1) + c] = g < ng ? d_out[((size_t)(g0 + g) * DANBO_J + j) * kOut + c] : 0.f;
    }
    for (int e = threadIdx.x; e < kG * kW; e += kThreads) {
        const int g = e / kW, c = e - g * kW;
        a_s[g * (kW + 1) + c] = g < ng ? S.n3[((size_t)(g0 + g) * DANBO_J + j) * kW + c] : 0.f;
    }
    __syncthreads();
    outer_acc<kW + 1, kOut + 1>(Gr.w3 + (size_t)j * kW * kOut, kW, kOut, a_s, d_s, excl);
    bias_acc(Gr.b3 + j * kOut, kOut, d_s, kOut + 1, excl);
    float acc[kG];
    rowmm<kOut + 1>(P.w3 + (size_t)j * kW * kOut, kW, kOut, d_s, tile, acc);           // d n3
    // d pre2 = d n3 * [n3 > 0]; reuse d_s with pitch kW + 1 (all reads of d out are behind the barrier inside rowmm)
    if (threadIdx.x < kW) {
#pragma unroll
        for (int g = 0; g < kG; ++g) d_s[g * (kW + 1) + threadIdx.x] = a_s[g * (kW + 1) + threadIdx.x] > 0.f ? acc[g] : 0.f;
    }
    __syncthreads();
    for (int e = threadIdx.x; e < kG * kW; e += kThreads) {
        const int g = e / kW, c = e - g * kW;
        a_s[g * (kW + 1) + c] = g < ng ? S.n2[((size_t)(g0 + g) * DANBO_J + j) * kW + c] : 0.f;
    }
    __syncthreads();
    outer_acc<kW + 1, kW + 1>(Gr.w2 + (size_t)j * kW * kW, kW, kW, a_s, d_s, excl);
    bias_acc(Gr.b2 + j * kW, kW, d_s, kW + 1, excl);
    rowmm<kW + 1>(P.w2 + (size_t)j * kW * kW, kW, kW, d_s, tile, acc);                 // d n2
    if (threadIdx.x < kW) {
        const int c = threadIdx.x;
        float sb = 0.f;
#pragma unroll
        for (int g = 0; g < kG; ++g) {
            const float v = a_s[g * (kW + 1) + c] > 0.f ? acc[g] : 0.f;               // d pre1
            if (g < ng) d_pre1[((size_t)(g0 + g) * DANBO_J + j) * kW + c] = v;
            sb += v;
        }
        atomicAdd(Gr.b1 + c, sb);                                                      // shared bias: summed over nodes
    }
}

// B1 / B0: un-mix d pre (node k collects from the nodes j it feeds), adjacency-weight gradient, per-joint linear backward.
//   kLayer 1: d lin1 -> dW1, d n1 -> d pre0 = 2 d n1 [n1 > 0] (written), db0 (atomics)
//   kLayer 0: d lin0 -> dW0
template <int kLayer>
__global__ void __launch_bounds__(kThreads)
bwd_mix_kernel(int G, Params P, Saved S, const float* __restrict__ d_pre /* (G,24,128) of this layer */, Grads Gr,
               float* __restrict__ d_pre_prev /* (G,24,128), layer 1 only */) {
    constexpr int kNin = kLayer == 1 ? kW : kIn;
    __shared__ float d_s[kG * (kW + 1)];            // d lin[g][c] of node k
    __shared__ float a_s[kG * (kW + 1)];            // input activations of node k (n1 or w_in)
    __shared__ float tile[128 * 33];
    __shared__ float red[kThreads / 32];
    const int k = blockIdx.x, g0 = blockIdx.y * kG, ng = min(kG, G - g0);
    const bool excl = gridDim.y == 1;
    const float* adjw = kLayer == 1 ? P.adjw1 : P.adjw0;
    const float* adj = kLayer == 1 ? P.adj1 : P.adj0;
    const float* lin = kLayer == 1 ? S.lin1 : S.lin0;
    float* d_adjw = kLayer == 1 ? Gr.adjw1 : Gr.adjw0;
    const int c = threadIdx.x & (kW - 1), half = threadIdx.x >> 7;      // two threads per column: poses split in halves
    // lin[g][k][c] of this node, for the adjacency gradient
    float lk[kG / 2];
#pragma unroll
    for (int u = 0; u < kG / 2; ++u) {
        const int g = half * (kG / 2) + u;
        lk[u] = g < ng ? lin[((size_t)(g0 + g) * DANBO_J + k) * kW + c] : 0.f;
    }
    float dl[kG / 2];
#pragma unroll
    for (int u = 0; u < kG / 2; ++u) dl[u] = 0.f;
    for (int j = 0; j < DANBO_J; ++j) {
        const float m = __ldg(adj + j * DANBO_J + k);                    // block-uniform
        if (m == 0.f) continue;
        const float a = __ldg(adjw + j * DANBO_J + k) * m;
        float dot = 0.f;
#pragma unroll
        for (int u = 0; u < kG / 2; ++u) {
            const int g = half * (kG / 2) + u;
            const float d = g < ng ? d_pre[((size_t)(g0 + g) * DANBO_J + j) * kW + c] : 0.f;
            dl[u] = fmaf(a, d, dl[u]);
            dot = fmaf(d, lk[u], dot);
        }
        // d A[j][k] = sum_{g,c} d pre[g][j][c] lin[g][k][c];  d adj_w = d A * adj
        dot = warp_sum(dot);
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = dot;
        __syncthreads();
        if (threadIdx.x == 0) {
            float s = 0.f;
#pragma unroll
            for (int w = 0; w < kThreads / 32; ++w) s += red[w];
            if (excl) d_adjw[j * DANBO_J + k] += s * m; else atomicAdd(d_adjw + j * DANBO_J + k, s * m);
        }
        __syncthreads();
    }
#pragma unroll
    for (int u = 0; u < kG / 2; ++u) d_s[(half * (kG / 2) + u) * (kW + 1) + c] = dl[u];
    const float* act = kLayer == 1 ? S.n1 : S.w_in;
    for (int e = threadIdx.x; e < kG * kNin; e += kThreads) {
        const int g = e / kNin, i = e - g * kNin;
        a_s[g * (kW + 1) + i] = g < ng ? act[((size_t)(g0 + g) * DANBO_J + k) * kNin + i] : 0.f;
    }
    __syncthreads();
    float* dW = kLayer == 1 ? Gr.w1 + (size_t)k * kW * kW : Gr.w0 + (size_t)k * kIn * kW;
    outer_acc<kW + 1, kW + 1>(dW, kNin, kW, a_s, d_s, excl);
    if (kLayer == 1) {
        float acc[kG];
        rowmm<kW + 1>(P.w1 + (size_t)k * kW * kW, kW, kW, d_s, tile, acc);            // d n1
        if (threadIdx.x < kW) {
            const int cc = threadIdx.x;
            float sb = 0.f;
#pragma unroll
            for (int g = 0; g < kG; ++g) {
                const float v = a_s[g * (kW + 1) + cc] > 0.f ? 2.f * acc[g] : 0.f;    // d pre0 (n1 = relu(2 pre0))
                if (g < ng) d_pre_prev[((size_t)(g0 + g) * DANBO_J + k) * kW + cc] = v;
                sb += v;
            }
            atomicAdd(Gr.b0 + cc, sb);
        }
    }
}

}  // namespace gn
}  // namespace danbo

using namespace danbo;

static gn::Params gn_params(const float* const* p) {
    gn::Params P;
    P.w0 = p[0]; P.adjw0 = p[1]; P.adj0 = p[2]; P.b0 = p[3]; P.w1 = p[4]; P.adjw1 = p[5]; P.adj1 = p[6]; P.b1 = p[7];
    P.w2 = p[8]; P.b2 = p[9]; P.w3 = p[10]; P.b3 = p[11];
    return P;
}
static gn::Saved gn_saved(float* const* s) { return gn::Saved{s[0], s[1], s[2], s[3], s[4], s[5]}; }

extern "C" int danbo_graph_net_fwd(const float* pose_bones, int n_poses, const float* const* params, float* const* saved,
                                   float* vol_out, void* stream) {
    if (n_poses <= 0) return 0;
    if (!pose_bones || !params || !saved || !vol_out) return -1;
    cudaStream_t st = (cudaStream_t)stream;
    const gn::Params P = gn_params(params);
    const gn::Saved S = gn_saved(saved);
    const dim3 grid(DANBO_J, (n_poses + gn::kG - 1) / gn::kG);
    gn::fwd0_kernel<<<grid, gn::kThreads, 0, st>>>(pose_bones, n_poses, P, S);
    DANBO_CHECK_LAUNCH();
    gn::fwd1_kernel<<<grid, gn::kThreads, 0, st>>>(n_poses, P, S);
    DANBO_CHECK_LAUNCH();
    gn::fwd2_kernel<<<grid, gn::kThreads, 0, st>>>(n_poses, P, S, vol_out);
    DANBO_CHECK_LAUNCH();
    return 0;
}

extern "C" int danbo_graph_net_bwd(int n_poses, const float* const* params, float* const* saved, const float* d_vol,
                                   float* const* grads, float* work /* 2 * n_poses*24*128 floats */, void* stream) {
    if (n_poses <= 0) return 0;
    if (!params || !saved || !d_vol || !grads || !work) return -1;
    cudaStream_t st = (cudaStream_t)stream;
    const gn::Params P = gn_params(params);
    const gn::Saved S = gn_saved(saved);
    const gn::Grads Gr{grads[0], grads[1], grads[2], grads[3], grads[4], grads[5], grads[6], grads[7], grads[8], grads[9]};
    float* d_pre1 = work;
    float* d_pre0 = work + (size_t)n_poses * DANBO_J * gn::kW;
    const dim3 grid(DANBO_J, (n_poses + gn::kG - 1) / gn::kG);
    gn::bwd2_kernel<<<grid, gn::kThreads, 0, st>>>(n_poses, P, S, d_vol, Gr, d_pre1);
    DANBO_CHECK_LAUNCH();
    gn::bwd_mix_kernel<1><<<grid, gn::kThreads, 0, st>>>(n_poses, P, S, d_pre1, Gr, d_pre0);
    DANBO_CHECK_LAUNCH();
    gn::bwd_mix_kernel<0><<<grid, gn::kThreads, 0, st>>>(n_poses, P, S, d_pre0, Gr, nullptr);
    DANBO_CHECK_LAUNCH();
    return 0;
}
