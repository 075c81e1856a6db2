// Microbenchmark: tcgen05.ld / tcgen05.st throughput per SM as a function of the number of warps, and the cost of the
// bf16x2 conversion.  One CTA.  nvcc -gencode arch=compute_100a,code=sm_100a -o tmem_bw tmem_bw.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../danbo-pytorch_b200/csrc/tc_common.cuh"
using namespace danbo::tc;

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr) : "memory");
}

__global__ void k(int mode, int iters, long long* out, uint32_t* sink) {
    __shared__ uint32_t tbase;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(danbo::smem_u32(&tbase)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t base = tbase + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((warp >> 2) * 64) % 512;
    uint32_t acc = 0;
    uint32_t v[32]; uint32_t w[32];
    __syncthreads();
    long long t0 = clock64();
    if (mode == 0) {            // ld x32, wait every 2 loads
        for (int i = 0; i < iters; ++i) {
            tmem_ld32(base, v); tmem_ld32(base + 32, w); tmem_wait_ld();
#pragma unroll
            for (int j = 0; j < 32; ++j) acc ^= v[j] ^ w[j];
        }
    } else if (mode == 1) {     // st x16 x2, wait
        uint32_t p[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) p[j] = lane + j;
        for (int i = 0; i < iters; ++i) { tmem_st16(base, p); tmem_st16(base + 16, p); tmem_wait_st(); }
    } else if (mode == 2) {     // ld x32 pair + add + cvt relu pack + st (epilogue-like, no bias)
        for (int i = 0; i < iters; ++i) {
            tmem_ld32(base, v); tmem_ld32(base + 32, w); tmem_wait_ld();
            uint32_t p[16], q[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                p[j] = pack_bf16_relu(__uint_as_float(v[2 * j]) + 1.f, __uint_as_float(v[2 * j + 1]) + 1.f);
                q[j] = pack_bf16_relu(__uint_as_float(w[2 * j]) + 1.f, __uint_as_float(w[2 * j + 1]) + 1.f);
            }
            tmem_st16(base + 256 % 512, p); tmem_st16(base + 256 % 512 + 16, q); tmem_wait_st();
        }
    } else if (mode == 3) {     // pure cvt throughput
        float a = lane, b = warp;
        for (int i = 0; i < iters; ++i) {
#pragma unroll
            for (int j = 0; j < 32; ++j) { uint32_t r = pack_bf16_relu(a, b); a += __uint_as_float(r & 0x3f800000); b += 1.f; acc ^= r; }
        }
    } else if (mode == 4) {     // ld x16 pipelined: 4 loads in flight, wait
        uint32_t a0[16], a1[16], a2[16], a3[16];
        for (int i = 0; i < iters; ++i) {
            tmem_ld16(base, a0); tmem_ld16(base + 16, a1); tmem_ld16(base + 32, a2); tmem_ld16(base + 48, a3); tmem_wait_ld();
#pragma unroll
            for (int j = 0; j < 16; ++j) acc ^= a0[j] ^ a1[j] ^ a2[j] ^ a3[j];
        }
    }
    long long t1 = clock64();
    __syncthreads();
    if (lane == 0) out[warp] = t1 - t0;
    sink[threadIdx.x] = acc;
    tc_fence_before(); __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tbase), "r"(512u) : "memory");
}

int main() {
    long long* out; uint32_t* sink;
    cudaMalloc(&out, 64 * 8); cudaMalloc(&sink, 4096 * 4);
    const int iters = 2000;
    const char* names[] = {"ld 2x(x32)+wait (8 KB/warp/iter)", "st 2x(x16)+wait (4 KB/warp/iter)", "epilogue-like ld+add+cvt+st",
                           "cvt.relu.bf16x2 x32/iter", "ld 4x(x16)+wait (8 KB/warp/iter)"};
    for (int mode = 0; mode < 5; ++mode)
        for (int warps : {1, 4, 8, 16}) {
            k<<<1, warps * 32>>>(mode, iters, out, sink);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
            long long h[32]; cudaMemcpy(h, out, warps * 8, cudaMemcpyDeviceToHost);
            long long mx = 0; for (int i = 0; i < warps; ++i) mx = h[i] > mx ? h[i] : mx;
            printf("mode %d [%s] warps %2d: %.1f clk/iter/warp (max over warps)\n", mode, names[mode], warps, (double)mx / iters);
        }
    return 0;
}
