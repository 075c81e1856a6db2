// Shared device helpers for the danbo_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>

#define DANBO_J 24          // SMPL joints
#define DANBO_FEAT 15       // per-bone feature: 5 channels x 3 axis lines (FGNNcat)
#define DANBO_VOL 240       // floats per bone volume: feat(5) x bin(16) x axis(3)
#define DANBO_RES 16
#define DANBO_AGG_W 32
#define DANBO_X_COLS 195    // PE(15) with 6 frequencies
#define DANBO_X_KPAD 256    // padded K of one X row in its tile
#define DANBO_TILE_M 128    // samples per MLP tile
#define DANBO_X_TILE_BYTES (DANBO_TILE_M * DANBO_X_KPAD * 2)

#define DANBO_CHECK_LAUNCH() do { cudaError_t e_ = cudaGetLastError(); if (e_ != cudaSuccess) return (int)e_; } while (0)

namespace danbo {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// Byte offset of element (row r, column k) of a [128 x 256] bf16 K-major tile stored as 4 chunks of
// [128 rows x 64 cols] in the canonical 128-byte-swizzled UMMA layout (8-row x 128 B atoms, 16 B units XORed
// with the row index inside the atom).  The same formula packs weight stages ([128 n-rows x 64 k]).
__host__ __device__ __forceinline__ uint32_t sw128_offset(uint32_t r, uint32_t k) {
    uint32_t chunk = k >> 6, kk = k & 63;
    return chunk * (128u * 128u) + (r >> 3) * 1024u + (r & 7) * 128u + (((kk >> 3) ^ (r & 7)) << 4) + ((kk & 7) << 1);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

}  // namespace danbo
