// Microbenchmark: cp.async.bulk (1-D global -> shared) issue interval and per-SM ingest rate, all SMs streaming the
// same 1.3 MB L2-resident region through a shared-memory ring, as a function of copy size and of the number of
// issuing warps.  No consumer: a slot is reused as soon as its copy has landed.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../danbo-pytorch_b200/csrc/tc_common.cuh"
using namespace danbo::tc;

constexpr int kRing = 128 * 1024;

__global__ void k(const uint8_t* src, int src_bytes, int copy_bytes, int n_copies, int n_warps, long long* out) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* sm = smem_raw + ((1024u - (danbo::smem_u32(smem_raw) & 1023u)) & 1023u);
    __shared__ uint64_t full[64];
    const int warp = threadIdx.x >> 5;
    const int slots = kRing / copy_bytes;           // <= 64
    if (threadIdx.x == 0) { for (int i = 0; i < 64; ++i) mbar_init(&full[i], 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    __syncthreads();
    long long t0 = clock64(), t_issue = 0;
    if (warp < n_warps) {
        const bool leader = elect_one();
        // warp w handles copies w, w + n_warps, ...
        for (int i = warp; i < n_copies; i += n_warps) {
            const int slot = i % slots;
            const int use = i / slots;
            if (use > 0) mbar_wait(&full[slot], (use - 1) & 1);        // previous copy into this slot has landed
            if (leader) {
                mbar_expect_tx(&full[slot], copy_bytes);
                bulk_g2s(sm + slot * copy_bytes, src + ((size_t)i * copy_bytes) % src_bytes, copy_bytes, &full[slot]);
            }
            __syncwarp();
            if (i < slots) t_issue = clock64();      // free-running issue phase (no waits yet)
        }
        // drain
        for (int s = 0; s < slots; ++s) {
            int last = -1;
            for (int i = s; i < n_copies; i += slots) if (i % n_warps == warp) last = i;
            if (last >= 0) mbar_wait(&full[s], (last / slots) & 1);
        }
    }
    long long t1 = clock64();
    __syncthreads();
    if (threadIdx.x == 0 && blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t_issue - t0; }
}

int main() {
    const int src_bytes = 84 * 16384;
    uint8_t* src; cudaMalloc(&src, src_bytes); cudaMemset(src, 1, src_bytes);
    long long* out; cudaMalloc(&out, 64);
    const int smem = kRing + 1024;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    for (int grid : {1, 148})
        for (int cb : {2048, 4096, 8192, 16384, 32768})
            for (int nw : {1, 2, 4}) {
                const int total = 8 << 20;          // 8 MB per SM
                const int n = total / cb;
                const int slots = kRing / cb;
                if (slots % nw) continue;
                k<<<grid, 160, smem>>>(src, src_bytes, cb, n, nw, out);     // warm-up (L2)
                k<<<grid, 160, smem>>>(src, src_bytes, cb, n, nw, out);
                cudaError_t e = cudaDeviceSynchronize();
                if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
                long long h[2]; cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost);
                const int free_issues = (slots + nw - 1) / nw;
                printf("grid %3d copy %5d B x %d warps: %.1f B/clk/SM, %.0f clk per copy (steady), free-running issue interval %.0f clk\n",
                       grid, cb, nw, (double)total / h[0], (double)h[0] / n, (double)h[1] / free_issues);
            }
    return 0;
}
