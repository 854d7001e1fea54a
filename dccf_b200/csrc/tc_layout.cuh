// tc_layout.cuh — operand-tile geometry shared by the tensor-core scorers (tc_scores.cu, tc_train.cu):
// 128-row x 32-K stages of K-major, no-swizzle core matrices, hi/lo TF32 split.
#pragma once
#include "common.cuh"

namespace dccf {

constexpr int TC_BM = 128;                 // rows per tile
constexpr int TC_KC = 32;                  // K per stage (8 core-matrix columns of 4 tf32)
constexpr int TC_STAGES = 2;
constexpr int TC_PRODUCERS = 512;          // 16 warps: enough independent Philox chains per SM to fill the issue slots
constexpr int TC_NT = TC_PRODUCERS + 64;   // + MMA warp + TMA warp
constexpr uint32_t TC_A_BYTES = TC_BM * TC_KC * 4;   // 16 KB (one of hi / lo)
constexpr uint32_t TC_B_BYTES = D * TC_KC * 4;       //  8 KB (one of hi / lo)
constexpr uint32_t TC_STAGE_BYTES = 2 * TC_A_BYTES + 2 * TC_B_BYTES;   // 48 KB
constexpr uint32_t TC_SMEM_BYTES = TC_STAGES * TC_STAGE_BYTES + 256;
constexpr int TC_NACC = 4;                 // TMEM accumulators per tile (see the MMA issuer)
constexpr uint32_t TC_TMEM_COLS = TC_NACC * D;   // 256 of the 512 columns: two CTAs per SM
constexpr uint32_t TC_LBO = 128;           // K-adjacent core matrices are contiguous
constexpr uint32_t TC_SBO = 1024;          // 8 core matrices (32 K values) per 8-row group

// byte offset of element (row r, k) inside a [rows x 32] K-major no-swizzle operand tile
__host__ __device__ __forceinline__ uint32_t core_offset(int r, int k) {
    return (uint32_t)((r >> 3) * TC_SBO + (k >> 2) * TC_LBO + (r & 7) * 16 + (k & 3) * 4);
}

__device__ __forceinline__ float tf32_hi(float v) { return __uint_as_float(__float_as_uint(v) & 0xFFFFE000u); }

}  // namespace dccf
