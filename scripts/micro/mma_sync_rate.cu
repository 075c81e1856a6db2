// Microbenchmark (for DESIGN §8.3, `pair_logits` on tensor cores): issue rate of the legacy warp-level
// mma.sync.m16n8k16 (bf16 x bf16 -> f32) on one SM of a B200, as a function of resident warps and of the number of
// independent accumulator chains per warp.  The aggregation net is [pairs x 80] x [80 x 32] and [pairs x 32] x [32 x 32] per
// bone: far too small for a tcgen05 tile pipeline to pay off per bone, so the question is whether mma.sync (weights held
// as B fragments in registers for a whole segment of pairs) beats the FFMA version (~77 warp-FFMA per pair, LSU bound).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o micro/bin/mma_sync_rate micro/mma_sync_rate.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int CHAINS>
__global__ void k(int n_iter, long long* out, float* sink) {
    uint32_t a[4] = {0x3f803f80u + threadIdx.x, 0x3f803f80u, 0x3f803f80u, 0x3f803f80u};   // bf16 pairs (values do not matter)
    uint32_t b[2] = {0x3f803f80u, 0x3f803f80u ^ threadIdx.x};
    float acc[CHAINS][4];
#pragma unroll
    for (int c = 0; c < CHAINS; ++c)
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[c][i] = 0.f;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < n_iter; ++it) {
#pragma unroll
        for (int c = 0; c < CHAINS; ++c)
            asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                         : "+f"(acc[c][0]), "+f"(acc[c][1]), "+f"(acc[c][2]), "+f"(acc[c][3])
                         : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
    }
    __syncthreads();
    const long long t1 = clock64();
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) s += acc[c][0] + acc[c][1] + acc[c][2] + acc[c][3];
    if (s == 12345.f) sink[0] = s;
    if (threadIdx.x == 0) out[0] = t1 - t0;
}

template <int CHAINS>
void run(int warps, long long* out, float* sink) {
    const int n_iter = 4096;
    k<CHAINS><<<1, warps * 32>>>(n_iter, out, sink);
    cudaDeviceSynchronize();
    long long clk;
    cudaMemcpy(&clk, out, sizeof(clk), cudaMemcpyDeviceToHost);
    const double mmas = (double)n_iter * CHAINS * warps;
    // one m16n8k16 = 2048 MAC; a [64 pairs x 80] x [80 x 32] + [64 x 32] x [32 x 32] chunk in 3-term split bf16 = 336 of them
    printf("warps %2d chains %d: %7.2f clk per MMA per SM  -> %6.0f dense bf16 MAC/clk/SM, %5.1f clk per pair (split-bf16 agg net)\n",
           warps, CHAINS, clk / mmas, 2048.0 * mmas / clk, 5.25 * clk / mmas);
}

int main() {
    long long* out; float* sink;
    cudaMalloc(&out, 64); cudaMalloc(&sink, 64);
    for (int warps : {1, 4, 8, 16, 32}) { run<1>(warps, out, sink); run<4>(warps, out, sink); run<8>(warps, out, sink); }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
    return 0;
}
