// Microbenchmark: tcgen05.mma execute rate on one SM (bf16, M=128 per CTA), as a function of N, SS vs TS operands and of
// how often tcgen05.commit is issued.  Operands are whatever is in shared memory / TMEM (values do not matter).
// The issue loop is straight-line (16 MMAs unrolled, all descriptor offsets compile-time) so that it is not the limit.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../danbo-pytorch_b200/csrc/tc_common.cuh"
using namespace danbo::tc;

template <int TS, int N, int CE>
__global__ void k(int n_iter, long long* out) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* sm = smem_raw + ((1024u - (danbo::smem_u32(smem_raw) & 1023u)) & 1023u);
    __shared__ uint32_t tbase;
    __shared__ uint64_t bar[2];
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(danbo::smem_u32(&tbase)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tmem = tbase;
    if (warp == 0) {
        const bool leader = elect_one();
        constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t(N) >> 3) << 17) | ((128u >> 4) << 24);
        const uint64_t hi = make_desc(0);
        const uint32_t a_s = danbo::smem_u32(sm), b_s = danbo::smem_u32(sm + 65536);
        const uint64_t a0 = hi | (uint64_t)((a_s >> 4) & 0x3FFF), b0 = hi | (uint64_t)((b_s >> 4) & 0x3FFF);
        long long t0 = clock64();
        for (int it = 0; it < n_iter; ++it) {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                // chunk c = i / 4 (64 k each: +16 KB for A in smem, +32 TMEM columns; B rotates over 4 stage buffers of N*128 B)
                const uint32_t d = tmem + ((N <= 128 && (it & 1)) ? 128u : 0u);
                const uint64_t bdesc = b0 + (uint64_t)(((i / 4) * (N * 128)) >> 4) + (i % 4) * 2;
                if (leader) {
                    if (TS) mma_ts(d, tmem + 256 + (i / 4) * 32 + (i % 4) * 8, bdesc, idesc, i ? 1u : 0u);
                    else mma_ss(d, a0 + (uint64_t)(((i / 4) * 16384) >> 4) + (i % 4) * 2, bdesc, idesc, i ? 1u : 0u);
                    if (CE > 0 && (i % CE) == CE - 1) tc_commit(&bar[1]);
                }
            }
            __syncwarp();
        }
        if (leader) tc_commit(&bar[0]);
        __syncwarp();
        long long t1 = clock64();
        mbar_wait(&bar[0], 0);
        long long t2 = clock64();
        if (threadIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
    }
    tc_fence_before(); __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

template <int TS, int N, int CE>
void run(long long* out) {
    const int smem = 65536 + 4 * 32768 + 1024;
    cudaFuncSetAttribute(k<TS, N, CE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    const int n_iter = 64;
    k<TS, N, CE><<<1, 128, smem>>>(n_iter, out);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); exit(1); }
    long long h[2]; cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost);
    printf("%s N=%3d commit_every=%2d: issue %.1f clk/MMA, complete %.1f clk/MMA (ideal %d)\n", TS ? "TS" : "SS", N, CE,
           (double)h[0] / (16 * n_iter), (double)h[1] / (16 * n_iter), N / 2);
}

int main() {
    long long* out; cudaMalloc(&out, 64);
    run<0, 64, 0>(out); run<0, 128, 0>(out); run<0, 256, 0>(out);
    run<0, 128, 16>(out); run<0, 128, 4>(out); run<0, 128, 1>(out); run<0, 256, 4>(out);
    run<1, 64, 0>(out); run<1, 128, 0>(out); run<1, 256, 0>(out);
    run<1, 128, 16>(out); run<1, 128, 4>(out); run<1, 128, 1>(out); run<1, 256, 4>(out);
    return 0;
}
