// full_scores.cu — full-catalogue scoring: the one place on this path where the work is a dense
// user x item GEMM, so it runs on the tcgen05 tensor cores.
//
//   score[u,i] = ( <A[u,:], B[i,:]> + row_bias[u] + col_bias[i] + g ) * col_scale[i]        A [U,64], B [I,64]
//
// Instances:
//   * the IPSBiasedMF exposure matrix (src/models/IPSBiasedMF.py:37-57; README.md:27-29: the full U x I
//     prediction matrix of that model IS <ds>.ips_expo_prob.npy): A = user factors, B = item factors,
//     row/col bias = user/item bias, g = global bias, col_scale = 1 / max(propensity, M);
//   * deterministic DCCF scoring of the whole catalogue (std 0, no confounders): A = E_user,
//     B = relu(PI + PF) (the per-item hidden vector), no biases.
// Two outputs: the materialised [U,I] matrix (HBM-write-bound: 4 B per score) and/or a fused per-user
// top-k (score desc, item id asc) that never writes the matrix — each epilogue thread owns one user row
// and keeps its running top-k in registers while the item tiles stream through TMEM.
//
// CTA = 128 users x a range of 64-item tiles, 9 warps:
//   warps 0-3  epilogue (thread = user row = TMEM lane), double-buffered accumulators
//   warps 4-7  stage the next item tile: load [64 x 64] factors, split hi/lo (3xTF32), core-matrix layout
//   warp  8    one thread issues tcgen05.mma.kind::tf32 (M=128, N=64, K=8) x 8 k-steps x 3 products
#include "common.cuh"
#include "tc_common.cuh"

namespace dccf {

constexpr int FS_BM = 128, FS_BN = 64;
constexpr int FS_NT = 288;
constexpr int FS_KMAX = 16;                              // largest fused top-k
constexpr uint32_t FS_A_IMG = FS_BM * 32 * 4;            // one [128 x 32] chunk image, 16 KB
constexpr uint32_t FS_B_IMG = FS_BN * 32 * 4;            // one [64 x 32] chunk image, 8 KB
constexpr uint32_t FS_A_BYTES = 4 * FS_A_IMG;            // hi k0-31, hi k32-63, lo k0-31, lo k32-63
constexpr uint32_t FS_B_BYTES = 4 * FS_B_IMG;            // same for the item tile, 32 KB per stage
constexpr uint32_t FS_COL_BYTES = 2 * FS_BN * 4;         // col_bias + col_scale of the tile
constexpr uint32_t FS_STAGE = FS_B_BYTES + FS_COL_BYTES;
constexpr uint32_t FS_SMEM = FS_A_BYTES + 2 * FS_STAGE + 256;
constexpr uint32_t FS_TMEM_COLS = 256;                   // 2 buffers x (main + correction) x 64 columns
constexpr uint32_t FS_LBO = 128, FS_SBO = 1024;

__device__ __forceinline__ uint32_t fs_core_offset(int r, int k) {   // inside a [rows x 32] image
    return (uint32_t)((r >> 3) * FS_SBO + (k >> 2) * FS_LBO + (r & 7) * 16 + (k & 3) * 4);
}
__device__ __forceinline__ float fs_hi(float v) { return __uint_as_float(__float_as_uint(v) & 0xFFFFE000u); }

struct FsParams {
    const float* A;          // [U,64]
    const float* B;          // [I,64]
    const float* row_bias;   // [U] or null
    const float* col_bias;   // [I] or null
    const float* col_scale;  // [I] or null
    float g;
    float* out;              // [U,I] or null
    float* topk_score;       // [splits,U,k] or null
    int64_t* topk_id;        // [splits,U,k]
    int32_t n_users, n_items, k;
    int32_t tiles_per_split, n_tiles;
};

// store a 4-float group of one row (split hi / lo into the chunk images at `base`)
__device__ __forceinline__ void fs_store_split(uint8_t* base, uint32_t img_bytes, int r, int k, const float4& v) {
    const int chunk = k >> 5, kk = k & 31;
    const uint32_t off = fs_core_offset(r, kk);
    float4 hi, lo;
    hi.x = fs_hi(v.x); hi.y = fs_hi(v.y); hi.z = fs_hi(v.z); hi.w = fs_hi(v.w);
    lo.x = __fsub_rn(v.x, hi.x); lo.y = __fsub_rn(v.y, hi.y); lo.z = __fsub_rn(v.z, hi.z); lo.w = __fsub_rn(v.w, hi.w);
    *reinterpret_cast<float4*>(base + chunk * img_bytes + off) = hi;
    *reinterpret_cast<float4*>(base + (2 + chunk) * img_bytes + off) = lo;
}

__global__ void __launch_bounds__(FS_NT, 1) k_full_scores(const FsParams prm) {
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t* a_s = smem;
    uint8_t* stage0 = smem + FS_A_BYTES;
    uint64_t* b_full = reinterpret_cast<uint64_t*>(smem + FS_A_BYTES + 2 * FS_STAGE);
    uint64_t* b_empty = b_full + 2;
    uint64_t* acc_full = b_empty + 2;
    uint64_t* acc_empty = acc_full + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int64_t u0 = (int64_t)blockIdx.x * FS_BM;
    const int split = blockIdx.y;
    const int t_lo = split * prm.tiles_per_split;
    const int t_hi = min(t_lo + prm.tiles_per_split, prm.n_tiles);

    if (tid == 0) {
        for (int s = 0; s < 2; ++s) {
            tc::mbar_init(&b_full[s], 4);      // four staging warps
            tc::mbar_init(&b_empty[s], 1);     // tcgen05.commit
            tc::mbar_init(&acc_full[s], 1);    // tcgen05.commit
            tc::mbar_init(&acc_empty[s], 4);   // four epilogue warps
        }
        tc::fence_barrier_init();
    }
    if (warp == 8) tc::tmem_alloc(tmem_slot, FS_TMEM_COLS);

    // user tile: 128 rows x 64 factors, split once (all threads)
    for (int i = tid; i < FS_BM * 16; i += FS_NT) {
        const int r = i >> 4, q = i & 15;
        const int64_t u = min(u0 + r, (int64_t)prm.n_users - 1);
        fs_store_split(a_s, FS_A_IMG, r, q * 4, ldg4(prm.A + (size_t)u * D + q * 4));
    }
    tc::fence_proxy_async_smem();
    tc::tc_fence_before_sync();
    __syncthreads();
    tc::tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;

    if (warp >= 4 && warp < 8) {
        // ===== item-tile staging =====
        const int t128 = tid - 128;
        const int item = t128 & 63, khalf = t128 >> 6;
        for (int t = t_lo; t < t_hi; ++t) {
            const int n = t - t_lo, s = n & 1;
            const uint32_t ph = (uint32_t)(n >> 1) & 1u;
            const int64_t i = min((int64_t)t * FS_BN + item, (int64_t)prm.n_items - 1);
            float4 v[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) v[q] = ldg4(prm.B + (size_t)i * D + khalf * 32 + q * 4);
            tc::mbar_wait(&b_empty[s], ph ^ 1u);
            uint8_t* st = stage0 + s * FS_STAGE;
#pragma unroll
            for (int q = 0; q < 8; ++q) fs_store_split(st, FS_B_IMG, item, khalf * 32 + q * 4, v[q]);
            if (khalf == 0) {
                // the epilogue of the tile that used this stage two tiles ago still reads its column vectors
                // until it hands the accumulator back
                tc::mbar_wait(&acc_empty[s], ph ^ 1u);
                float* col = reinterpret_cast<float*>(st + FS_B_BYTES);
                col[item] = prm.col_bias ? __ldg(prm.col_bias + i) : 0.f;
                col[FS_BN + item] = prm.col_scale ? __ldg(prm.col_scale + i) : 1.f;
            }
            tc::fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(&b_full[s]);
        }
    } else if (warp == 8) {
        // ===== MMA issuer =====
        if (lane == 0) {
            constexpr uint32_t idesc = tc::make_idesc_tf32(FS_BM, FS_BN);
            const uint32_t a_base = tc::smem_u32(a_s);
            for (int t = t_lo; t < t_hi; ++t) {
                const int n = t - t_lo, s = n & 1;
                const uint32_t ph = (uint32_t)(n >> 1) & 1u;
                tc::mbar_wait(&b_full[s], ph);
                tc::mbar_wait(&acc_empty[s], ph ^ 1u);     // the epilogue has drained this accumulator pair
                tc::tc_fence_after_sync();
                const uint32_t b_base = tc::smem_u32(stage0 + s * FS_STAGE);
                const uint32_t d_main = tmem_base + (uint32_t)(s * 2) * FS_BN;
                const uint32_t d_corr = d_main + FS_BN;
#pragma unroll
                for (int ks = 0; ks < 8; ++ks) {
                    const uint32_t chunk = ks >> 2, ko = (uint32_t)(ks & 3) * 2 * FS_LBO;
                    const uint64_t a_hi = tc::make_smem_desc(a_base + chunk * FS_A_IMG + ko, FS_LBO, FS_SBO);
                    const uint64_t a_lo = tc::make_smem_desc(a_base + (2 + chunk) * FS_A_IMG + ko, FS_LBO, FS_SBO);
                    const uint64_t b_hi = tc::make_smem_desc(b_base + chunk * FS_B_IMG + ko, FS_LBO, FS_SBO);
                    const uint64_t b_lo = tc::make_smem_desc(b_base + (2 + chunk) * FS_B_IMG + ko, FS_LBO, FS_SBO);
                    tc::umma_tf32(d_corr, a_lo, b_hi, idesc, ks != 0);
                    tc::umma_tf32(d_corr, a_hi, b_lo, idesc, 1u);
                    tc::umma_tf32(d_main, a_hi, b_hi, idesc, ks != 0);
                }
                tc::umma_commit(&b_empty[s]);
                tc::umma_commit(&acc_full[s]);
            }
        }
        __syncwarp();
    } else {
        // ===== epilogue: thread = user row =====
        const int row = tid;
        const int64_t u = u0 + row;
        const bool valid = u < prm.n_users;
        const float rb = (valid && prm.row_bias) ? __ldg(prm.row_bias + u) : 0.f;
        const float add = rb + prm.g;
        float best_s[FS_KMAX];
        int32_t best_i[FS_KMAX];
#pragma unroll
        for (int j = 0; j < FS_KMAX; ++j) { best_s[j] = -INFINITY; best_i[j] = -1; }
        const int k = prm.k;
        float thr_s = -INFINITY;      // score and id of the current k-th entry (kept in registers: the
        int32_t thr_i = -1;           // arrays are only touched through fully unrolled loops)
        for (int t = t_lo; t < t_hi; ++t) {
            const int n = t - t_lo, s = n & 1;
            const uint32_t ph = (uint32_t)(n >> 1) & 1u;
            tc::mbar_wait(&acc_full[s], ph);
            tc::tc_fence_after_sync();
            const float* col = reinterpret_cast<const float*>(stage0 + s * FS_STAGE + FS_B_BYTES);
            const uint32_t lane_addr = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(s * 2) * FS_BN;
            const int64_t i0 = (int64_t)t * FS_BN;
#pragma unroll 1
            for (int quarter = 0; quarter < 4; ++quarter) {
                float acc[16], corr[16];
                tc::tmem_ld_32x16(lane_addr + quarter * 16, acc);
                tc::tmem_ld_32x16(lane_addr + FS_BN + quarter * 16, corr);
                float v[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const int c = quarter * 16 + j;
                    v[j] = ((acc[j] + corr[j]) + (add + col[c])) * col[FS_BN + c];
                }
                if (prm.out != nullptr && valid) {
#pragma unroll
                    for (int j = 0; j < 16; j += 4) {
                        const int64_t i = i0 + quarter * 16 + j;
                        float* dst = prm.out + (size_t)u * prm.n_items + i;
                        if (i + 3 < prm.n_items && ((prm.n_items & 3) == 0)) {
                            st4(dst, make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]));
                        } else {
#pragma unroll
                            for (int e = 0; e < 4; ++e)
                                if (i + e < prm.n_items) dst[e] = v[j + e];
                        }
                    }
                }
                if (prm.topk_score != nullptr) {
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const int64_t i = i0 + quarter * 16 + j;
                        float sc = (v[j] != v[j]) ? -INFINITY : v[j];
                        if (i < prm.n_items && (sc > thr_s || thr_i < 0)) {
                            // insert, keeping (score desc, id asc): items arrive in ascending id, so an
                            // equal score never displaces an earlier one
                            int32_t ci = (int32_t)i;
#pragma unroll
                            for (int q = 0; q < FS_KMAX; ++q) {
                                if (q < k && (sc > best_s[q] || best_i[q] < 0)) {
                                    const float ts = best_s[q];
                                    const int32_t ti = best_i[q];
                                    best_s[q] = sc;
                                    best_i[q] = ci;
                                    sc = ts;
                                    ci = ti;
                                }
                            }
#pragma unroll
                            for (int q = 0; q < FS_KMAX; ++q)
                                if (q == k - 1) { thr_s = best_s[q]; thr_i = best_i[q]; }
                        }
                    }
                }
            }
            tc::tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(&acc_empty[s]);
        }
        if (prm.topk_score != nullptr && valid) {
            float* ds = prm.topk_score + ((size_t)split * prm.n_users + u) * k;
            int64_t* di = prm.topk_id + ((size_t)split * prm.n_users + u) * k;
#pragma unroll
            for (int q = 0; q < FS_KMAX; ++q)
                if (q < k) { ds[q] = best_s[q]; di[q] = best_i[q]; }
        }
    }

    tc::tc_fence_before_sync();
    __syncthreads();
    if (warp == 8) tc::tmem_dealloc(tmem_base, FS_TMEM_COLS);
}

// merge the per-split partial top-k lists: one thread per user
__global__ void k_topk_merge(const float* __restrict__ ps, const int64_t* __restrict__ pi, int32_t n_splits,
                             int32_t n_users, int32_t k, float* __restrict__ out_s, int64_t* __restrict__ out_i) {
    const int u = blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= n_users) return;
    float last_s = INFINITY;
    int64_t last_i = -1;
    for (int q = 0; q < k; ++q) {
        float bs = -INFINITY;
        int64_t bi = -1;
        for (int s = 0; s < n_splits; ++s)
            for (int j = 0; j < k; ++j) {
                const size_t o = ((size_t)s * n_users + u) * k + j;
                const float sc = ps[o];
                const int64_t id = pi[o];
                if (id < 0) continue;
                const bool below = (q == 0) || sc < last_s || (sc == last_s && id > last_i);
                if (below && (bi < 0 || sc > bs || (sc == bs && id < bi))) { bs = sc; bi = id; }
            }
        out_s[(size_t)u * k + q] = bs;
        out_i[(size_t)u * k + q] = bi;
        last_s = bs;
        last_i = bi;
    }
}

}  // namespace dccf

using namespace dccf;

extern "C" int32_t dccf_full_scores_splits(int32_t n_users, int32_t n_items) {
    const int ut = (n_users + FS_BM - 1) / FS_BM, it = (n_items + FS_BN - 1) / FS_BN;
    if (ut <= 0 || it <= 0) return 1;
    int splits = (2 * 148 + ut - 1) / ut;      // about two CTAs' worth of work per SM
    if (splits > it) splits = it;
    if (splits < 1) splits = 1;
    return splits;
}

extern "C" int dccf_full_scores(int32_t n_users, int32_t n_items, const float* A, const float* B,
                                const float* row_bias, const float* col_bias, const float* col_scale, float g,
                                float* out, int32_t k, float* topk_score, int64_t* topk_id, float* ws_score,
                                int64_t* ws_id, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    DCCF_CHECK_ARG(A && B, "dccf_full_scores: null factor matrix");
    DCCF_CHECK_ARG(out || (topk_score && topk_id), "dccf_full_scores: nothing to produce (out and top-k both null)");
    DCCF_CHECK_ARG((topk_score == nullptr) == (topk_id == nullptr), "dccf_full_scores: top-k needs both score and id outputs");
    DCCF_CHECK_ARG(topk_score == nullptr || (k >= 1 && k <= FS_KMAX), "dccf_full_scores: fused top-k supports 1 <= k <= %d, got %d", FS_KMAX, k);
    if (n_users <= 0 || n_items <= 0) return DCCF_OK;
    const int splits = dccf_full_scores_splits(n_users, n_items);
    DCCF_CHECK_ARG(topk_score == nullptr || splits == 1 || (ws_score && ws_id), "dccf_full_scores: top-k with %d item splits needs the [splits,U,k] workspaces", splits);
    FsParams prm;
    prm.A = A; prm.B = B; prm.row_bias = row_bias; prm.col_bias = col_bias; prm.col_scale = col_scale; prm.g = g;
    prm.out = out; prm.n_users = n_users; prm.n_items = n_items; prm.k = topk_score ? k : 1;
    prm.n_tiles = (n_items + FS_BN - 1) / FS_BN;
    prm.tiles_per_split = (prm.n_tiles + splits - 1) / splits;
    prm.topk_score = topk_score ? (splits == 1 ? topk_score : ws_score) : nullptr;
    prm.topk_id = topk_score ? (splits == 1 ? topk_id : ws_id) : nullptr;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(k_full_scores, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FS_SMEM);
        if (e != cudaSuccess) {
            set_error("dccf_full_scores: cannot opt in to %u bytes of shared memory: %s", FS_SMEM, cudaGetErrorString(e));
            return DCCF_ERR_CUDA;
        }
        attr_set = true;
    }
    dim3 grid((unsigned)((n_users + FS_BM - 1) / FS_BM), (unsigned)splits);
    k_full_scores<<<grid, FS_NT, FS_SMEM, stream>>>(prm);
    DCCF_CHECK_LAUNCH("k_full_scores");
    if (topk_score && splits > 1) {
        k_topk_merge<<<(unsigned)((n_users + 127) / 128), 128, 0, stream>>>(ws_score, ws_id, splits, n_users, k, topk_score, topk_id);
        DCCF_CHECK_LAUNCH("k_topk_merge");
    }
    return DCCF_OK;
}
