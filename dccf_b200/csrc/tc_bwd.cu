// tc_bwd.cu — the feature part of d loss / d W on the tcgen05 tensor cores.
//
//   gW[:, 64 + f] = sum_r dpre[r, :] * x[r, f],   x[r, :] = Feat[i_r, :] + eps[r, :]     (768 of the 832 columns)
//
// is the one large contraction of the backward pass: [768 x N] x [N x 64] with the reduction over the N
// predictor rows of the batch (autograd computes it as addmm's weight gradient, src/runners/BaseRunner.py:183).
// Here the reduction dimension (rows) is the MMA's K, so both operands are "MN-major" in UMMA terms: the
// producers write x[row, 4 features] and dpre[row, 4 outputs] as 16-byte pieces exactly as they compute them,
// 8 rows x 16 bytes forming one 128-byte core matrix.  Same 3xTF32 error compensation and accumulator
// rotation as the forward scorer (tc_scores.cu); eps is regenerated from the library's Philox stream, so the
// [N,768] noise tensor never exists in memory.
//
// CTA = (block of 128 feature columns) x (split of the rows); 9 warps:
//   warps 0-7  producers: per 32-row chunk, x (Feat + eps) and dpre (= ds * E_user * gate), hi/lo split
//   warp  8    one thread issues tcgen05.mma.kind::tf32 (M = 128 features, N = 64 outputs, K = 8 rows)
//   warps 0-3  epilogue: TMEM lane = feature; writes gW_part[split][j][64 + f]
// The 64 item-embedding columns, db, the embedding-gradient records and the loss stay in k_bpr_bwd.
#include "bwd_common.cuh"
#include "tc_common.cuh"

namespace dccf {

constexpr int TB_FB = 128;                       // feature columns per CTA (MMA M)
constexpr int TB_RC = 32;                        // rows per stage (4 MMA k-steps)
constexpr int TB_STAGES = 2;
constexpr int TB_NT = 288;
constexpr uint32_t TB_A_BYTES = TB_RC * TB_FB * 4;   // 16 KB (one of hi / lo)
constexpr uint32_t TB_B_BYTES = TB_RC * D * 4;       //  8 KB
constexpr uint32_t TB_STAGE_BYTES = 2 * TB_A_BYTES + 2 * TB_B_BYTES;
constexpr uint32_t TB_SMEM = TB_STAGES * TB_STAGE_BYTES + 256;
constexpr uint32_t TB_TMEM_COLS = 256;
// MN-major, no swizzle: element (mn, k) at (k%8)*16 + (mn/4)*SBO + (k/8)*LBO + (mn%4)*4
constexpr uint32_t TB_SBO = 128;                 // consecutive 4-element MN blocks are contiguous core matrices
constexpr uint32_t TB_A_LBO = (TB_FB / 4) * 128; // next group of 8 rows: 4096 B
constexpr uint32_t TB_B_LBO = (D / 4) * 128;     // 2048 B

__host__ __device__ constexpr uint32_t make_idesc_tf32_mn(int M, int N) {
    return tc::make_idesc_tf32(M, N) | (1u << 15) | (1u << 16);   // a_major = b_major = MN
}

__device__ __forceinline__ void tb_store_split(uint8_t* hi_base, uint8_t* lo_base, uint32_t off, const float4& v) {
    float4 hi, lo;
    hi.x = __uint_as_float(__float_as_uint(v.x) & 0xFFFFE000u);
    hi.y = __uint_as_float(__float_as_uint(v.y) & 0xFFFFE000u);
    hi.z = __uint_as_float(__float_as_uint(v.z) & 0xFFFFE000u);
    hi.w = __uint_as_float(__float_as_uint(v.w) & 0xFFFFE000u);
    lo.x = __fsub_rn(v.x, hi.x); lo.y = __fsub_rn(v.y, hi.y); lo.z = __fsub_rn(v.z, hi.z); lo.w = __fsub_rn(v.w, hi.w);
    *reinterpret_cast<float4*>(hi_base + off) = hi;
    *reinterpret_cast<float4*>(lo_base + off) = lo;
}

__global__ void __launch_bounds__(TB_NT, 2) k_bwd_gw_tc(const BwdParams prm, int32_t rows_per_split) {
    extern __shared__ __align__(128) uint8_t smem[];
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + TB_STAGES * TB_STAGE_BYTES);
    uint64_t* empty_bar = full_bar + TB_STAGES;
    uint64_t* accum_bar = empty_bar + TB_STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum_bar + 1);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int fb = blockIdx.x;              // feature block
    const int split = blockIdx.y;
    const int64_t row_lo = (int64_t)split * rows_per_split;
    const int64_t row_hi = min(row_lo + rows_per_split, prm.n_rows);
    const int n_chunks = row_hi > row_lo ? (int)((row_hi - row_lo + TB_RC - 1) / TB_RC) : 0;
    const int K = D + prm.F;

    if (tid == 0) {
        for (int s = 0; s < TB_STAGES; ++s) {
            tc::mbar_init(&full_bar[s], 8);
            tc::mbar_init(&empty_bar[s], 1);
        }
        tc::mbar_init(accum_bar, 1);
        tc::fence_barrier_init();
    }
    if (warp == 8) tc::tmem_alloc(tmem_slot, TB_TMEM_COLS);
    tc::tc_fence_before_sync();
    __syncthreads();
    tc::tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;

    if (warp < 8) {
        // ===== producers: thread = (row k of the chunk, group of 4 feature quads / 2 output quads) =====
        const int k = tid & 31, grp = tid >> 5;
        const RngKey key_noise = resolve_rng_key(prm.rng, DOMAIN_NOISE);
        const uint32_t row_off = (uint32_t)((k & 7) * 16);
        for (int c = 0; c < n_chunks; ++c) {
            const int s = c % TB_STAGES;
            const uint32_t ph = (uint32_t)(c / TB_STAGES) & 1u;
            const int64_t r = row_lo + (int64_t)c * TB_RC + k;
            const bool live = r < row_hi;
            float4 xa[4], db[2];
#pragma unroll
            for (int i = 0; i < 4; ++i) xa[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            db[0] = db[1] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (live) {
                const int64_t p = r / prm.R;
                const int rem = (int)(r - p * prm.R);
                const int z = rem / prm.A;
                const int32_t u = checked_id(prm.X[2 * p], prm.n_users, nullptr);
                const int32_t fi = checked_id(prm.X[2 * p + 1], prm.n_items, nullptr);
                const float ds = dpred_of(prm, p) * __ldg(prm.save_w + p * prm.Z + z) * prm.inv_A;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int f = fb * TB_FB + (grp * 4 + i) * 4;
                    float4 v = ldg4(prm.Feat + (size_t)fi * prm.F + f);
                    if (prm.noise_mode == 1) {
                        const float4 e = ldg4(prm.noise + (size_t)r * prm.F + f);
                        v.x = __fadd_rn(v.x, e.x); v.y = __fadd_rn(v.y, e.y); v.z = __fadd_rn(v.z, e.z); v.w = __fadd_rn(v.w, e.w);
                    } else if (prm.noise_mode == 2) {
                        const float4 e = noise_quad(key_noise, (uint32_t)r, (uint32_t)(f >> 2), prm.noise_std);
                        v.x = __fadd_rn(v.x, e.x); v.y = __fadd_rn(v.y, e.y); v.z = __fadd_rn(v.z, e.z); v.w = __fadd_rn(v.w, e.w);
                    }
                    xa[i] = v;
                }
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    const int j0 = (grp * 2 + i) * 4;
                    const float4 h = ldg4(prm.save_h + (size_t)r * D + j0);
                    const float4 e = ldg4(prm.E_user + (size_t)u * D + j0);
                    float4 g;
                    if (prm.mask_mode == 1) {
                        const float4 m = ldg4(prm.mask + (size_t)r * D + j0);
                        g = make_float4(h.x > 0.f ? m.x : 0.f, h.y > 0.f ? m.y : 0.f, h.z > 0.f ? m.z : 0.f, h.w > 0.f ? m.w : 0.f);
                    } else {
                        const float sc = (prm.mask_mode == 2) ? prm.drop_scale : 1.f;
                        g = make_float4(h.x > 0.f ? sc : 0.f, h.y > 0.f ? sc : 0.f, h.z > 0.f ? sc : 0.f, h.w > 0.f ? sc : 0.f);
                    }
                    db[i] = make_float4(ds * e.x * g.x, ds * e.y * g.y, ds * e.z * g.z, ds * e.w * g.w);
                }
            }
            tc::mbar_wait(&empty_bar[s], ph ^ 1u);
            uint8_t* a_hi = smem + s * TB_STAGE_BYTES;
            uint8_t* a_lo = a_hi + TB_A_BYTES;
            uint8_t* b_hi = a_lo + TB_A_BYTES;
            uint8_t* b_lo = b_hi + TB_B_BYTES;
#pragma unroll
            for (int i = 0; i < 4; ++i)
                tb_store_split(a_hi, a_lo, row_off + (uint32_t)(grp * 4 + i) * TB_SBO + (uint32_t)(k >> 3) * TB_A_LBO, xa[i]);
#pragma unroll
            for (int i = 0; i < 2; ++i)
                tb_store_split(b_hi, b_lo, row_off + (uint32_t)(grp * 2 + i) * TB_SBO + (uint32_t)(k >> 3) * TB_B_LBO, db[i]);
            tc::fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(&full_bar[s]);
        }
    } else {
        // ===== MMA issuer =====
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc_tf32_mn(TB_FB, D);
            for (int c = 0; c < n_chunks; ++c) {
                const int s = c % TB_STAGES;
                const uint32_t ph = (uint32_t)(c / TB_STAGES) & 1u;
                tc::mbar_wait(&full_bar[s], ph);
                tc::tc_fence_after_sync();
                const uint32_t a_hi = tc::smem_u32(smem + s * TB_STAGE_BYTES);
                const uint32_t a_lo = a_hi + TB_A_BYTES;
                const uint32_t b_hi = a_lo + TB_A_BYTES;
                const uint32_t b_lo = b_hi + TB_B_BYTES;
#pragma unroll
                for (int j = 0; j < TB_RC / 8; ++j) {
                    const uint64_t da_hi = tc::make_smem_desc(a_hi + j * TB_A_LBO, TB_A_LBO, TB_SBO);
                    const uint64_t da_lo = tc::make_smem_desc(a_lo + j * TB_A_LBO, TB_A_LBO, TB_SBO);
                    const uint64_t db_hi = tc::make_smem_desc(b_hi + j * TB_B_LBO, TB_B_LBO, TB_SBO);
                    const uint64_t db_lo = tc::make_smem_desc(b_lo + j * TB_B_LBO, TB_B_LBO, TB_SBO);
                    const int ks = c * (TB_RC / 8) + j;
                    const uint32_t main_acc = tmem_base + (uint32_t)(1 + ks % 3) * D;   // see tc_scores.cu: truncating accumulate
                    tc::umma_tf32(tmem_base, da_lo, db_hi, idesc, ks != 0);
                    tc::umma_tf32(tmem_base, da_hi, db_lo, idesc, 1u);
                    tc::umma_tf32(main_acc, da_hi, db_hi, idesc, ks >= 3);
                }
                tc::umma_commit(&empty_bar[s]);
            }
            tc::umma_commit(accum_bar);
        }
        __syncwarp();
    }

    // ===== epilogue: warps 0-3, thread = feature column = TMEM lane =====
    if (warp < 4) {
        const int f = fb * TB_FB + tid;                       // feature index, always < F (F % 128 == 0 checked on host)
        float* out = prm.gW_part + (size_t)split * D * K + D + f;
        if (n_chunks > 0) {
            tc::mbar_wait(accum_bar, 0u);
            tc::tc_fence_after_sync();
        }
#pragma unroll 1
        for (int quarter = 0; quarter < 4; ++quarter) {
            float acc[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) acc[j] = 0.f;
            if (n_chunks > 0) {
                float part[16];
                const uint32_t lane_addr = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(quarter * 16);
                tc::tmem_ld_32x16(lane_addr + 1 * D, acc);
                tc::tmem_ld_32x16(lane_addr + 2 * D, part);
#pragma unroll
                for (int j = 0; j < 16; ++j) acc[j] += part[j];
                tc::tmem_ld_32x16(lane_addr + 3 * D, part);
#pragma unroll
                for (int j = 0; j < 16; ++j) acc[j] += part[j];
                tc::tmem_ld_32x16(lane_addr, part);
#pragma unroll
                for (int j = 0; j < 16; ++j) acc[j] += part[j];
            }
#pragma unroll
            for (int j = 0; j < 16; ++j) out[(size_t)(quarter * 16 + j) * K] = acc[j];
        }
    }

    tc::tc_fence_before_sync();
    __syncthreads();
    if (warp == 8) tc::tmem_dealloc(tmem_base, TB_TMEM_COLS);
}

// host side, called from dccf_bpr_bwd when the feature part goes to the tensor cores
int launch_bwd_gw_tc(const BwdParams& prm, cudaStream_t stream) {
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(k_bwd_gw_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TB_SMEM);
        if (e != cudaSuccess) {
            set_error("dccf_bpr_bwd: cannot opt in to %u bytes of shared memory: %s", TB_SMEM, cudaGetErrorString(e));
            return DCCF_ERR_CUDA;
        }
        attr_set = true;
    }
    const int64_t rps = (prm.n_rows + prm.n_splits - 1) / prm.n_splits;
    const int32_t rows_per_split = (int32_t)(((rps + TB_RC - 1) / TB_RC) * TB_RC);
    dim3 grid((unsigned)(prm.F / TB_FB), (unsigned)prm.n_splits);
    k_bwd_gw_tc<<<grid, TB_NT, TB_SMEM, stream>>>(prm, rows_per_split);
    return DCCF_OK;
}

}  // namespace dccf
