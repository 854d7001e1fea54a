// tc_scores.cu — tensor-core (tcgen05 / TMEM) variant of the evaluation scorer.
//
// Same math as k_row_scores (src/models/DCCF.py:84-96 of the reference), re-associated so that the only
// large contraction is the one that cannot be shared between rows:
//
//   pre[p,z,a,:] = W_i·E_item[item_z]  +  (W_f·Feat[i_p] + b)  +  W_f·eps[p,z,a,:]
//                  `---- PI[item_z] ---'   `------ PF[i_p] -----'   `- [128 x 768]x[768 x 64] per tile -'
//
// PI and PF are [I,64] tables projected once per parameter version (k_table_gemm, FP32 SIMT).  The noise
// product runs on the 5th-generation tensor cores as an error-compensated 3xTF32 product
// (hi·hi + hi·lo + lo·hi with hi = top 19 bits, lo = exact remainder; relative error ~2^-21 per product,
// FP32 accumulation in TMEM), which keeps the 1e-5 parity bound of the FP32 reference.
//
// CTA = one 128-row tile, 10 warps:
//   warps 0-15 producers: generate eps (Philox4x32-10 + Box-Muller, the library's stream — bit-identical to
//              the SIMT kernels) or read the explicit noise tensor, split hi/lo, store the K-major
//              core-matrix layout the UMMA descriptors describe, fence.proxy.async, arrive on `full`
//   warp  16   one thread issues tcgen05.mma.kind::tf32 (M=128, N=64, K=8), tcgen05.commit frees the stage
//   warp  17   one thread streams the pre-split W_f chunks with cp.async.bulk (TMA engine) into the stage
//   warps 0-3  epilogue: tcgen05.ld the accumulator (lane = row), + PI + PF, ReLU, dropout, dot with E_user
#include "common.cuh"
#include "tc_common.cuh"
#include "tc_layout.cuh"

namespace dccf {

// ---------------------------------------------------------------------------------------------
// W_f -> per-chunk operand images [chunk][ hi 8 KB | lo 8 KB ] in the core-matrix layout
// ---------------------------------------------------------------------------------------------
__global__ void k_prep_wf(const float* __restrict__ W, int F, float* __restrict__ gB) {
    const int K = D + F;
    const int total = D * F;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int n = i / F, f = i - n * F;
        const float w = __ldg(W + (size_t)n * K + D + f);
        const float hi = tf32_hi(w);
        const int c = f / TC_KC, kk = f - c * TC_KC;
        float* base = gB + (size_t)c * (2 * TC_B_BYTES / 4);
        const uint32_t off = core_offset(n, kk) / 4;
        base[off] = hi;
        base[TC_B_BYTES / 4 + off] = __fsub_rn(w, hi);
    }
}

// Which table row feeds output row i:  0 = row i itself (whole-table projection);
// 1 = the item in slot (p, z) of the batch, i = p*Z + z;  2 = the true item of pair i.
struct RowSource {
    int32_t mode;
    int32_t Z, S, n_items;
    const int64_t* X;
    const int64_t* sample_item;
};
__device__ __forceinline__ int64_t source_row(const RowSource& rs, int64_t i) {
    if (rs.mode == 0) return i;
    if (rs.mode == 2) return checked_id(rs.X[2 * i + 1], rs.n_items, nullptr);
    const int64_t p = i / rs.Z;
    const int z = (int)(i - p * rs.Z);
    return checked_id(z == 0 ? rs.X[2 * p + 1] : rs.sample_item[p * rs.S + (z - 1)], rs.n_items, nullptr);
}

// out[i,:] = X[src(i),:] · Wt_sub + bias      X [*,Kx] row-major, Wt_sub = rows of the k-major W copy ([Kx,64])
struct TableJob {
    const float* X;
    int64_t n;
    int32_t Kx;
    const float* Wt_sub;
    const float* bias;
    float* out;
    RowSource src;
    int32_t block_lo, block_n;
};
struct TableJobs {
    TableJob job[2];
    int32_t n_jobs;
};

__global__ void __launch_bounds__(256) k_table_gemm(const TableJobs jobs) {
    __shared__ __align__(16) float Xs[16][64 + 4];
    __shared__ __align__(16) float Ws[16][D];
    int ji = 0;
    if (jobs.n_jobs > 1 && (int)blockIdx.x >= jobs.job[1].block_lo) ji = 1;
    const TableJob& jb = jobs.job[ji];
    const float* __restrict__ X = jb.X;
    const float* __restrict__ Wt_sub = jb.Wt_sub;
    const float* __restrict__ bias = jb.bias;
    float* __restrict__ out = jb.out;
    const int64_t n = jb.n;
    const int Kx = jb.Kx;
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;          // 4 cols x 4 rows per thread
    const int64_t row0 = (int64_t)((int)blockIdx.x - jb.block_lo) * 64;
    const int64_t my_src = source_row(jb.src, min(row0 + (tid >> 2), n - 1));
    float acc[4][4] = {};
    for (int k0 = 0; k0 < Kx; k0 += 16) {
        {   // stage X: 64 rows x 16 k  (thread: row = tid>>2, 4 k)
            const int r = tid >> 2, q = tid & 3;
            const int64_t gr = my_src;
            const float4 v = ldg4(X + (size_t)gr * Kx + k0 + q * 4);
            Xs[q * 4 + 0][r] = v.x; Xs[q * 4 + 1][r] = v.y; Xs[q * 4 + 2][r] = v.z; Xs[q * 4 + 3][r] = v.w;
            st4(&Ws[tid >> 4][(tid & 15) * 4], ldg4(Wt_sub + (size_t)(k0 + (tid >> 4)) * D + (tid & 15) * 4));
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            const float4 a = ld4(&Xs[k][ty * 4]);
            const float4 b = ld4(&Ws[k][tx * 4]);
            const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        __syncthreads();
    }
    float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
    if (bias != nullptr) bv = ldg4(bias + tx * 4);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int64_t gr = row0 + ty * 4 + i;
        if (gr < n) st4(out + (size_t)gr * D + tx * 4, make_float4(acc[i][0] + bv.x, acc[i][1] + bv.y, acc[i][2] + bv.z, acc[i][3] + bv.w));
    }
}

struct TcParams {
    const float* E_user;
    const float* PI;     // [I,64]  W_i·E_item
    const float* PF;     // [I,64]  W_f·Feat + b
    const float* gB;     // pre-split W_f operand images
    const int64_t* X;
    const int64_t* sample_item;
    const float* noise;  // mode 1
    const float* mask;   // mode 1
    float* ws_rows;
    float* save_h;       // optional [N,64]: post-dropout activations (training)
    int32_t batch_tables;  // 0: PI/PF are [I,64] tables indexed by item id; 1: PI is [P*Z,64] (slot order), PF [P,64]
    float* dbg_pre;      // optional [N,64]: raw accumulator (W_f·eps), for tests
    int32_t* err_flag;
    int64_t n_rows;
    int32_t n_users, n_items, F, S, A, R;
    int32_t user_base;
    int32_t mask_mode;
    float noise_std, keep_prob, drop_scale;
    RngSpec rng;
};

template <int NOISE_MODE>
__global__ void __launch_bounds__(TC_NT, 2) k_row_scores_tc(const TcParams prm) {
    extern __shared__ __align__(128) uint8_t smem[];
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + TC_STAGES * TC_STAGE_BYTES);
    uint64_t* empty_bar = full_bar + TC_STAGES;
    uint64_t* accum_bar = empty_bar + TC_STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum_bar + 1);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int64_t row_base = (int64_t)blockIdx.x * TC_BM;
    const int n_chunks = prm.F / TC_KC;

    if (tid == 0) {
        for (int s = 0; s < TC_STAGES; ++s) {
            tc::mbar_init(&full_bar[s], TC_PRODUCERS / 32 + 1);   // 8 producer warps + the expect_tx arrival
            tc::mbar_init(&empty_bar[s], 1);                      // one tcgen05.commit
        }
        tc::mbar_init(accum_bar, 1);
        tc::fence_barrier_init();
    }
    if (warp == TC_PRODUCERS / 32) tc::tmem_alloc(tmem_slot, TC_TMEM_COLS);
    tc::tc_fence_before_sync();
    __syncthreads();
    tc::tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;

    if (warp < TC_PRODUCERS / 32) {
        // ===== producers: thread = (row, quarter of the 32-wide K chunk) =====
        const int row = tid & (TC_BM - 1), kq = tid >> 7;
        const int64_t grow = min(row_base + row, prm.n_rows - 1);
        const RngKey key_noise = resolve_rng_key(prm.rng, DOMAIN_NOISE);
        const float* nptr = (NOISE_MODE == 1) ? prm.noise + (size_t)grow * prm.F + kq * 8 : nullptr;
        for (int c = 0; c < n_chunks; ++c) {
            const int s = c % TC_STAGES;
            const uint32_t ph = (uint32_t)(c / TC_STAGES) & 1u;
            float4 e[2];
            if (NOISE_MODE == 1) {
#pragma unroll
                for (int q = 0; q < 2; ++q) e[q] = ldg4(nptr + c * TC_KC + 4 * q);
            } else {
#pragma unroll
                for (int q = 0; q < 2; ++q)
                    e[q] = noise_quad(key_noise, (uint32_t)grow, (uint32_t)(c * 8 + kq * 2 + q), prm.noise_std);
            }
            tc::mbar_wait(&empty_bar[s], ph ^ 1u);   // the MMAs that read this stage have completed
            uint8_t* a_hi = smem + s * TC_STAGE_BYTES;
            uint8_t* a_lo = a_hi + TC_A_BYTES;
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                const uint32_t off = core_offset(row, kq * 8 + 4 * q);
                float4 hi, lo;
                hi.x = tf32_hi(e[q].x); hi.y = tf32_hi(e[q].y); hi.z = tf32_hi(e[q].z); hi.w = tf32_hi(e[q].w);
                lo.x = __fsub_rn(e[q].x, hi.x); lo.y = __fsub_rn(e[q].y, hi.y);
                lo.z = __fsub_rn(e[q].z, hi.z); lo.w = __fsub_rn(e[q].w, hi.w);
                *reinterpret_cast<float4*>(a_hi + off) = hi;
                *reinterpret_cast<float4*>(a_lo + off) = lo;
            }
            tc::fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(&full_bar[s]);
        }
    } else if (warp == TC_PRODUCERS / 32) {
        // ===== MMA issuer =====
        if (lane == 0) {
            constexpr uint32_t idesc = tc::make_idesc_tf32(TC_BM, D);
            for (int c = 0; c < n_chunks; ++c) {
                const int s = c % TC_STAGES;
                const uint32_t ph = (uint32_t)(c / TC_STAGES) & 1u;
                tc::mbar_wait(&full_bar[s], ph);
                tc::tc_fence_after_sync();
                const uint32_t a_hi = tc::smem_u32(smem + s * TC_STAGE_BYTES);
                const uint32_t a_lo = a_hi + TC_A_BYTES;
                const uint32_t b_hi = a_hi + 2 * TC_A_BYTES;
                const uint32_t b_lo = b_hi + TC_B_BYTES;
#pragma unroll
                for (int j = 0; j < TC_KC / 8; ++j) {
                    const uint32_t ko = (uint32_t)j * 2 * TC_LBO;   // two core matrices per K=8 step
                    const uint64_t da_hi = tc::make_smem_desc(a_hi + ko, TC_LBO, TC_SBO);
                    const uint64_t da_lo = tc::make_smem_desc(a_lo + ko, TC_LBO, TC_SBO);
                    const uint64_t db_hi = tc::make_smem_desc(b_hi + ko, TC_LBO, TC_SBO);
                    const uint64_t db_lo = tc::make_smem_desc(b_lo + ko, TC_LBO, TC_SBO);
                    // The tensor core adds into the FP32 accumulator with truncation, a bias of ~2^-24 per
                    // accumulation that 288 chained MMAs would grow to ~8e-6.  The two correction products
                    // (2^-11 smaller) share accumulator 0; the main hi·hi product rotates over accumulators
                    // 1..3 (32 accumulations each); the epilogue adds the four in round-to-nearest FP32.
                    const int ks = c * (TC_KC / 8) + j;
                    const uint32_t main_acc = tmem_base + (uint32_t)(1 + ks % (TC_NACC - 1)) * D;
                    tc::umma_tf32(tmem_base, da_lo, db_hi, idesc, ks != 0);
                    tc::umma_tf32(tmem_base, da_hi, db_lo, idesc, 1u);
                    tc::umma_tf32(main_acc, da_hi, db_hi, idesc, ks >= TC_NACC - 1);
                }
                tc::umma_commit(&empty_bar[s]);   // frees the stage when these MMAs have read it
            }
            tc::umma_commit(accum_bar);           // accumulator complete
        }
        __syncwarp();
    } else {
        // ===== W_f streamer (TMA engine bulk copies) =====
        if (lane == 0) {
            for (int c = 0; c < n_chunks; ++c) {
                const int s = c % TC_STAGES;
                const uint32_t ph = (uint32_t)(c / TC_STAGES) & 1u;
                tc::mbar_wait(&empty_bar[s], ph ^ 1u);
                tc::mbar_arrive_expect_tx(&full_bar[s], 2 * TC_B_BYTES);
                tc::bulk_g2s(smem + s * TC_STAGE_BYTES + 2 * TC_A_BYTES, prm.gB + (size_t)c * (2 * TC_B_BYTES / 4),
                             2 * TC_B_BYTES, &full_bar[s]);
            }
        }
        __syncwarp();
    }

    // ===== epilogue: warps 0-3, thread = row = TMEM lane =====
    if (warp < 4) {
        tc::mbar_wait(accum_bar, 0u);
        tc::tc_fence_after_sync();
        const int row = tid;
        const int64_t grow_raw = row_base + row;
        const bool valid = grow_raw < prm.n_rows;
        const int64_t grow = valid ? grow_raw : prm.n_rows - 1;
        const int64_t p = grow / prm.R;
        const int rem = (int)(grow - p * prm.R);
        const int z = rem / prm.A;
        const int32_t u = checked_id(prm.X[2 * p] - prm.user_base, prm.n_users, prm.err_flag);
        const int32_t fi = checked_id(prm.X[2 * p + 1], prm.n_items, prm.err_flag);
        const int32_t it = (z == 0) ? fi : checked_id(prm.sample_item[p * prm.S + (z - 1)], prm.n_items, prm.err_flag);
        const float* pi = prm.PI + (size_t)(prm.batch_tables ? (p * (prm.S + 1) + z) : (int64_t)it) * D;
        const float* pf = prm.PF + (size_t)(prm.batch_tables ? p : (int64_t)fi) * D;
        const float* eu = prm.E_user + (size_t)u * D;
        const RngKey key_drop = resolve_rng_key(prm.rng, DOMAIN_DROPOUT);
        float dot = 0.f;
#pragma unroll 1
        for (int quarter = 0; quarter < 4; ++quarter) {
            float acc[16], part[16];
            const uint32_t lane_addr = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(quarter * 16);
            tc::tmem_ld_32x16(lane_addr + 1 * D, acc);
            tc::tmem_ld_32x16(lane_addr + 2 * D, part);
#pragma unroll
            for (int j = 0; j < 16; ++j) acc[j] += part[j];
            tc::tmem_ld_32x16(lane_addr + 3 * D, part);
#pragma unroll
            for (int j = 0; j < 16; ++j) acc[j] += part[j];
            tc::tmem_ld_32x16(lane_addr, part);          // the small correction terms last
#pragma unroll
            for (int j = 0; j < 16; ++j) acc[j] += part[j];
            if (prm.dbg_pre != nullptr && valid) {
#pragma unroll
                for (int j = 0; j < 16; j += 4)
                    st4(prm.dbg_pre + (size_t)grow * D + quarter * 16 + j, make_float4(acc[j], acc[j + 1], acc[j + 2], acc[j + 3]));
            }
#pragma unroll
            for (int j = 0; j < 16; j += 4) {
                const int col = quarter * 16 + j;
                const float4 a = ldg4(pi + col), f = ldg4(pf + col), e = ldg4(eu + col);
                float h0 = fmaxf(acc[j] + (a.x + f.x), 0.f), h1 = fmaxf(acc[j + 1] + (a.y + f.y), 0.f);
                float h2 = fmaxf(acc[j + 2] + (a.z + f.z), 0.f), h3 = fmaxf(acc[j + 3] + (a.w + f.w), 0.f);
                if (prm.mask_mode == 1) {
                    const float4 m = ldg4(prm.mask + (size_t)grow * D + col);
                    h0 *= m.x; h1 *= m.y; h2 *= m.z; h3 *= m.w;
                } else if (prm.mask_mode == 2) {
                    const float4 m = dropout_quad(key_drop, (uint32_t)grow, (uint32_t)(col >> 2), prm.keep_prob, prm.drop_scale);
                    h0 *= m.x; h1 *= m.y; h2 *= m.z; h3 *= m.w;
                }
                if (prm.save_h != nullptr && valid) st4(prm.save_h + (size_t)grow * D + col, make_float4(h0, h1, h2, h3));
                dot = fmaf(h0, e.x, dot);
                dot = fmaf(h1, e.y, dot);
                dot = fmaf(h2, e.z, dot);
                dot = fmaf(h3, e.w, dot);
            }
        }
        if (valid) prm.ws_rows[grow] = dot;
    }

    tc::tc_fence_before_sync();
    __syncthreads();
    if (warp == TC_PRODUCERS / 32) tc::tmem_dealloc(tmem_base, TC_TMEM_COLS);
}

// defined in score_fwd.cu
void launch_backdoor(const dccf_expo* expo, const int64_t* X, const int64_t* sample_item, int64_t n_pairs,
                     const dccf_dims* dims, const float* ws_rows, float* out_pred, float* save_w, int32_t* err_flag,
                     cudaStream_t stream);
void launch_transpose_w(const float* W, float* Wt, int K, cudaStream_t stream);

}  // namespace dccf

using namespace dccf;

extern "C" int64_t dccf_tc_operand_floats(int32_t feat_dim) {
    return (int64_t)(feat_dim / TC_KC) * (2 * TC_B_BYTES / 4);
}

static void fill_job(TableJob& j, const float* X, int64_t n, int Kx, const float* Wt_sub, const float* bias, float* out,
                     int mode, const dccf_dims* dims, const int64_t* Xids, const int64_t* sample_item, int32_t& blocks) {
    j.X = X; j.n = n; j.Kx = Kx; j.Wt_sub = Wt_sub; j.bias = bias; j.out = out;
    j.src.mode = mode; j.src.Z = dims->n_samples + 1; j.src.S = dims->n_samples; j.src.n_items = dims->n_items;
    j.src.X = Xids; j.src.sample_item = sample_item;
    j.block_lo = blocks; j.block_n = (int32_t)((n + 63) / 64); blocks += j.block_n;
}

extern "C" int dccf_tc_prepare(const dccf_dims* dims, const float* E_item, const float* Feat, const float* W,
                               const float* b, float* ws_wt, float* PI, float* PF, float* gB, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    DCCF_CHECK_ARG(dims && E_item && Feat && W && b && ws_wt && PI && PF && gB, "dccf_tc_prepare: null argument");
    DCCF_CHECK_ARG(dims->dim == D, "dccf_tc_prepare: dim=%d but this build has D=%d", dims->dim, D);
    DCCF_CHECK_ARG(dims->feat_dim > 0 && dims->feat_dim % 64 == 0, "dccf_tc_prepare: feat_dim=%d must be a positive multiple of 64", dims->feat_dim);
    const int F = dims->feat_dim, K = D + F;
    launch_transpose_w(W, ws_wt, K, stream);
    DCCF_CHECK_LAUNCH("k_transpose_w");
    TableJobs jobs;
    int32_t blocks = 0;
    fill_job(jobs.job[0], E_item, dims->n_items, D, ws_wt, nullptr, PI, 0, dims, nullptr, nullptr, blocks);
    fill_job(jobs.job[1], Feat, dims->n_items, F, ws_wt + (size_t)D * D, b, PF, 0, dims, nullptr, nullptr, blocks);
    jobs.n_jobs = 2;
    k_table_gemm<<<(unsigned)blocks, 256, 0, stream>>>(jobs);
    DCCF_CHECK_LAUNCH("k_table_gemm");
    k_prep_wf<<<96, 256, 0, stream>>>(W, F, gB);
    DCCF_CHECK_LAUNCH("k_prep_wf");
    return DCCF_OK;
}

static int tc_launch(const dccf_dims* dims, const float* E_user, const float* PI, const float* PF, const float* gB,
                     int batch_tables, const dccf_expo* expo, const int64_t* X, const int64_t* sample_item,
                     int64_t n_pairs, const dccf_rng* rng, float* out_pred, float* ws_rows, float* save_h, float* save_w,
                     float* dbg_pre, int32_t* err_flag, cudaStream_t stream);

extern "C" int dccf_score_fwd_tc(const dccf_dims* dims, const float* E_user, const float* PI, const float* PF,
                                 const float* gB, const dccf_expo* expo, const int64_t* X, const int64_t* sample_item,
                                 int64_t n_pairs, const dccf_rng* rng, float* out_pred, float* ws_rows, float* dbg_pre,
                                 int32_t* err_flag, void* stream_) {
    DCCF_CHECK_ARG(dims && expo && rng, "dccf_score_fwd_tc: null struct argument");
    DCCF_CHECK_ARG(E_user && PI && PF && gB && X && out_pred && ws_rows, "dccf_score_fwd_tc: null buffer");
    return tc_launch(dims, E_user, PI, PF, gB, 0, expo, X, sample_item, n_pairs, rng, out_pred, ws_rows, nullptr, nullptr,
                     dbg_pre, err_flag, (cudaStream_t)stream_);
}

static int tc_launch(const dccf_dims* dims, const float* E_user, const float* PI, const float* PF, const float* gB,
                     int batch_tables, const dccf_expo* expo, const int64_t* X, const int64_t* sample_item,
                     int64_t n_pairs, const dccf_rng* rng, float* out_pred, float* ws_rows, float* save_h, float* save_w,
                     float* dbg_pre, int32_t* err_flag, cudaStream_t stream) {
    DCCF_CHECK_ARG(dims->dim == D, "dccf_score_fwd_tc: dim=%d but this build has D=%d", dims->dim, D);
    DCCF_CHECK_ARG(dims->feat_dim > 0 && dims->feat_dim % 64 == 0, "dccf_score_fwd_tc: feat_dim=%d must be a positive multiple of 64", dims->feat_dim);
    DCCF_CHECK_ARG(dims->n_samples == 0 || sample_item, "dccf_score_fwd_tc: sample_item is null");
    DCCF_CHECK_ARG(rng->noise_mode == 1 || rng->noise_mode == 2, "dccf_score_fwd_tc: needs feature noise (mode 1 or 2); without noise use dccf_score_fwd");
    DCCF_CHECK_ARG(rng->noise_mode != 1 || rng->noise, "dccf_score_fwd_tc: noise_mode 1 needs a noise tensor");
    DCCF_CHECK_ARG(rng->mask_mode >= 0 && rng->mask_mode <= 2 && (rng->mask_mode != 1 || rng->mask), "dccf_score_fwd_tc: bad dropout mask mode");
    if (n_pairs <= 0) return DCCF_OK;
    const int Z = dims->n_samples + 1, R = Z * dims->n_attr;
    const int64_t n_rows = n_pairs * R;
    DCCF_CHECK_ARG(n_rows < (int64_t)1 << 31, "dccf_score_fwd_tc: %lld rows in one call (max 2^31-1)", (long long)n_rows);

    TcParams prm;
    prm.E_user = E_user; prm.PI = PI; prm.PF = PF; prm.gB = gB; prm.X = X; prm.sample_item = sample_item;
    prm.noise = rng->noise; prm.mask = rng->mask; prm.ws_rows = ws_rows; prm.save_h = save_h; prm.batch_tables = batch_tables;
    prm.dbg_pre = dbg_pre; prm.err_flag = err_flag;
    prm.n_rows = n_rows; prm.n_users = dims->n_users; prm.user_base = dims->user_base; prm.n_items = dims->n_items; prm.F = dims->feat_dim;
    prm.S = dims->n_samples; prm.A = dims->n_attr; prm.R = R; prm.mask_mode = rng->mask_mode;
    prm.noise_std = rng->noise_std; prm.keep_prob = 1.0f - rng->p_drop;
    prm.drop_scale = (rng->p_drop < 1.0f) ? 1.0f / (1.0f - rng->p_drop) : 0.0f;
    prm.rng.seed = rng->seed; prm.rng.offset = rng->offset; prm.rng.offset_dev = rng->offset_dev;

    static PerDeviceOnce attr_once;
    if (attr_once.need()) {
        cudaError_t e1 = cudaFuncSetAttribute(k_row_scores_tc<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TC_SMEM_BYTES);
        cudaError_t e2 = cudaFuncSetAttribute(k_row_scores_tc<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TC_SMEM_BYTES);
        if (e1 != cudaSuccess || e2 != cudaSuccess) {
            set_error("dccf_score_fwd_tc: cannot opt in to %u bytes of shared memory: %s", TC_SMEM_BYTES,
                      cudaGetErrorString(e1 != cudaSuccess ? e1 : e2));
            return DCCF_ERR_CUDA;
        }
        attr_once.mark();
    }
    const unsigned grid = (unsigned)((n_rows + TC_BM - 1) / TC_BM);
    if (rng->noise_mode == 1) k_row_scores_tc<1><<<grid, TC_NT, TC_SMEM_BYTES, stream>>>(prm);
    else k_row_scores_tc<2><<<grid, TC_NT, TC_SMEM_BYTES, stream>>>(prm);
    DCCF_CHECK_LAUNCH("k_row_scores_tc");
    launch_backdoor(expo, X, sample_item, n_pairs, dims, ws_rows, out_pred, save_w, err_flag, stream);
    DCCF_CHECK_LAUNCH("k_backdoor");
    return DCCF_OK;
}
