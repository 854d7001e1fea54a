// score_fwd.cu — kernels (a)+(b) of the DCCF hot path: fused gather of user / item / feature rows,
// the per-(u,i,z,a) predictor rows and the backdoor-adjusted score.
//
// Replaces src/models/DCCF.py:74-100 of the reference.  The reference materialises
// mlp_input[N, D+F] (N = P*R rows) and runs cuBLAS; here a CTA owns 128 consecutive rows, builds
// their [E_item[item_z] | Feat[i] + eps] tile chunk by chunk in shared memory (noise generated in
// registers when rng mode 2) and keeps the 128x64 pre-activations in registers.
#include "common.cuh"
#include "backdoor.cuh"

namespace dccf {

constexpr int FWD_BM = 128;  // rows per CTA
constexpr int FWD_KC = 16;   // K chunk
constexpr int FWD_NT = 128;  // threads per CTA

// W [D, D+F] row-major  ->  Wt [(D+F), D]  (k-major) so a K chunk of W is one contiguous 4 KB block.
__global__ void k_transpose_w(const float* __restrict__ W, float* __restrict__ Wt, int K) {
    __shared__ float tile[32][33];
    const int k0 = blockIdx.x * 32, j0 = blockIdx.y * 32;
    const int tx = threadIdx.x, ty = threadIdx.y;  // 32 x 8
    for (int i = ty; i < 32; i += 8) {
        const int j = j0 + i, k = k0 + tx;
        tile[i][tx] = (k < K) ? W[(size_t)j * K + k] : 0.f;
    }
    __syncthreads();
    for (int i = ty; i < 32; i += 8) {
        const int k = k0 + i, j = j0 + tx;
        if (k < K) Wt[(size_t)k * D + j] = tile[tx][i];
    }
}

struct FwdParams {
    const float* E_user;
    const float* E_item;
    const float* Feat;
    const float* Wt;    // [(D+F), D]
    const float* bias;  // [D]
    const int64_t* X;   // [P,2]
    const int64_t* sample_item;  // [P,S]
    const float* noise;  // mode 1
    const float* mask;   // mode 1
    float* ws_rows;      // [N]
    float* save_h;       // [N,D] or null
    int32_t* err_flag;
    int64_t n_rows;
    int32_t n_users, n_items, F, S, A, R;
    int32_t user_base;
    int32_t mask_mode;
    float noise_std, keep_prob, drop_scale;
    RngSpec rng;
};

// s[r] = < E_user[u_r], dropout(relu(W [E_item[item_r] | Feat[i_r] + eps_r] + b)) >
template <int NOISE_MODE>
__global__ void __launch_bounds__(FWD_NT) k_row_scores(const FwdParams prm) {
    __shared__ __align__(16) float As[2][FWD_KC][FWD_BM];
    __shared__ __align__(16) float Ws[2][FWD_KC][D];
    __shared__ int32_t u_row_s[FWD_BM];

    const int tid = threadIdx.x;
    const int tx = tid & 7, ty = tid >> 3;
    const int64_t row_base = (int64_t)blockIdx.x * FWD_BM;
    const RngKey key_noise = resolve_rng_key(prm.rng, DOMAIN_NOISE);
    const RngKey key_drop = resolve_rng_key(prm.rng, DOMAIN_DROPOUT);

    // ---- per-thread fill row: thread t stages row t of the tile -------------------------------
    const int64_t my_row = min(row_base + tid, prm.n_rows - 1);
    const float* item_ptr;
    const float* feat_ptr;
    {
        const int64_t p = my_row / prm.R;
        const int rem = (int)(my_row - p * prm.R);
        const int z = rem / prm.A;
        const int32_t u = checked_id(prm.X[2 * p] - prm.user_base, prm.n_users, prm.err_flag);
        const int32_t fi = checked_id(prm.X[2 * p + 1], prm.n_items, prm.err_flag);
        const int32_t it = (z == 0) ? fi : checked_id(prm.sample_item[p * prm.S + (z - 1)], prm.n_items, prm.err_flag);
        u_row_s[tid] = u;
        item_ptr = prm.E_item + (size_t)it * D;
        feat_ptr = prm.Feat + (size_t)fi * prm.F;
    }
    const float* noise_ptr = (NOISE_MODE == 1) ? prm.noise + (size_t)my_row * prm.F : nullptr;
    const uint32_t my_row32 = (uint32_t)my_row;

    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

    float4 a_reg[4], n_reg[4], w_reg[2];
    const int n_chunks = (D + prm.F) / FWD_KC;

    auto prefetch = [&](int c) {
        const int k0 = c * FWD_KC;
        const float* src = (k0 < D) ? item_ptr + k0 : feat_ptr + (k0 - D);
#pragma unroll
        for (int q = 0; q < 4; ++q) a_reg[q] = ldg4(src + 4 * q);
        if (NOISE_MODE == 1 && k0 >= D) {
#pragma unroll
            for (int q = 0; q < 4; ++q) n_reg[q] = ldg4(noise_ptr + (k0 - D) + 4 * q);
        }
        const float* wsrc = prm.Wt + (size_t)k0 * D;
        w_reg[0] = ldg4(wsrc + tid * 4);
        w_reg[1] = ldg4(wsrc + 512 + tid * 4);
    };
    auto stage = [&](int c, int buf) {
        const int k0 = c * FWD_KC;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            float4 v = a_reg[q];
            if (NOISE_MODE != 0 && k0 >= D) {
                float4 e;
                if (NOISE_MODE == 1) e = n_reg[q];
                else e = noise_quad(key_noise, my_row32, (uint32_t)((k0 - D) / 4 + q), prm.noise_std);
                v.x = __fadd_rn(v.x, e.x);
                v.y = __fadd_rn(v.y, e.y);
                v.z = __fadd_rn(v.z, e.z);
                v.w = __fadd_rn(v.w, e.w);
            }
            As[buf][4 * q + 0][tid] = v.x;
            As[buf][4 * q + 1][tid] = v.y;
            As[buf][4 * q + 2][tid] = v.z;
            As[buf][4 * q + 3][tid] = v.w;
        }
        float* wdst = &Ws[buf][0][0];
        st4(wdst + tid * 4, w_reg[0]);
        st4(wdst + 512 + tid * 4, w_reg[1]);
    };

    prefetch(0);
    stage(0, 0);
    __syncthreads();

    for (int c = 0; c < n_chunks; ++c) {
        const int buf = c & 1;
        if (c + 1 < n_chunks) prefetch(c + 1);
#pragma unroll
        for (int k = 0; k < FWD_KC; ++k) {
            const float4 a0 = ld4(&As[buf][k][ty * 4]);
            const float4 a1 = ld4(&As[buf][k][64 + ty * 4]);
            const float4 b0 = ld4(&Ws[buf][k][tx * 4]);
            const float4 b1 = ld4(&Ws[buf][k][32 + tx * 4]);
            const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        if (c + 1 < n_chunks) stage(c + 1, buf ^ 1);
        __syncthreads();
    }

    // ---- epilogue: bias, relu, dropout, dot with the user row, reduce over the 8 column lanes ----
    const float4 bias0 = ldg4(prm.bias + tx * 4);
    const float4 bias1 = ldg4(prm.bias + 32 + tx * 4);
    const float bia[8] = {bias0.x, bias0.y, bias0.z, bias0.w, bias1.x, bias1.y, bias1.z, bias1.w};
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int m = (i < 4) ? (ty * 4 + i) : (64 + ty * 4 + (i - 4));
        const int64_t grow_raw = row_base + m;
        const bool valid = grow_raw < prm.n_rows;
        const int64_t grow = valid ? grow_raw : prm.n_rows - 1;
        float h[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) h[j] = fmaxf(acc[i][j] + bia[j], 0.f);
        if (prm.mask_mode == 1) {
            const float4 m0 = ldg4(prm.mask + (size_t)grow * D + tx * 4);
            const float4 m1 = ldg4(prm.mask + (size_t)grow * D + 32 + tx * 4);
            h[0] *= m0.x; h[1] *= m0.y; h[2] *= m0.z; h[3] *= m0.w;
            h[4] *= m1.x; h[5] *= m1.y; h[6] *= m1.z; h[7] *= m1.w;
        } else if (prm.mask_mode == 2) {
            const float4 m0 = dropout_quad(key_drop, (uint32_t)grow, (uint32_t)tx, prm.keep_prob, prm.drop_scale);
            const float4 m1 = dropout_quad(key_drop, (uint32_t)grow, (uint32_t)(8 + tx), prm.keep_prob, prm.drop_scale);
            h[0] *= m0.x; h[1] *= m0.y; h[2] *= m0.z; h[3] *= m0.w;
            h[4] *= m1.x; h[5] *= m1.y; h[6] *= m1.z; h[7] *= m1.w;
        }
        if (prm.save_h != nullptr && valid) {
            st4(prm.save_h + (size_t)grow * D + tx * 4, make_float4(h[0], h[1], h[2], h[3]));
            st4(prm.save_h + (size_t)grow * D + 32 + tx * 4, make_float4(h[4], h[5], h[6], h[7]));
        }
        const float* eu = prm.E_user + (size_t)u_row_s[m] * D;
        const float4 e0 = ldg4(eu + tx * 4);
        const float4 e1 = ldg4(eu + 32 + tx * 4);
        float part = h[0] * e0.x;
        part = fmaf(h[1], e0.y, part);
        part = fmaf(h[2], e0.z, part);
        part = fmaf(h[3], e0.w, part);
        part = fmaf(h[4], e1.x, part);
        part = fmaf(h[5], e1.y, part);
        part = fmaf(h[6], e1.z, part);
        part = fmaf(h[7], e1.w, part);
        part += __shfl_xor_sync(0xffffffffu, part, 1);
        part += __shfl_xor_sync(0xffffffffu, part, 2);
        part += __shfl_xor_sync(0xffffffffu, part, 4);
        if (tx == 0 && valid) prm.ws_rows[grow] = part;
    }
}

// ----------------------------------------------------------------------------------------------
// Small-batch variant (training step: 256 pairs = 5 632 rows would give the 128-row kernel only 44 CTAs
// for 148 SMs).  A CTA owns 32 rows; its 4 warps split K four ways and each runs a private
// register-tiled pipeline (8x8 outputs per lane, warp-private staging buffers, __syncwarp only);
// the four partial 32x64 tiles are summed through shared memory before the fused epilogue.
// ----------------------------------------------------------------------------------------------
constexpr int SK_BM = 32;     // rows per CTA
constexpr int SK_WARPS = 4;   // K splits
constexpr int SK_NT = SK_WARPS * 32;

template <int NOISE_MODE>
__global__ void __launch_bounds__(SK_NT) k_row_scores_splitk(const FwdParams prm) {
    // staging: per warp As[16][32] + Ws[16][64] = 1536 floats; reduction: 4 x 32 x 64 floats (aliased)
    __shared__ __align__(16) float smem[SK_WARPS * SK_BM * D];
    __shared__ int32_t u_row_s[SK_BM];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tx = lane & 7, ty = lane >> 3;
    const int64_t row_base = (int64_t)blockIdx.x * SK_BM;
    const RngKey key_noise = resolve_rng_key(prm.rng, DOMAIN_NOISE);
    const RngKey key_drop = resolve_rng_key(prm.rng, DOMAIN_DROPOUT);
    float* As = smem + warp * (FWD_KC * SK_BM + FWD_KC * D);   // [16][32]
    float* Ws = As + FWD_KC * SK_BM;                           // [16][64]

    // lane l stages row l of the tile
    const int64_t my_row = min(row_base + lane, prm.n_rows - 1);
    const float* item_ptr;
    const float* feat_ptr;
    {
        const int64_t p = my_row / prm.R;
        const int rem = (int)(my_row - p * prm.R);
        const int z = rem / prm.A;
        const int32_t u = checked_id(prm.X[2 * p] - prm.user_base, prm.n_users, prm.err_flag);
        const int32_t fi = checked_id(prm.X[2 * p + 1], prm.n_items, prm.err_flag);
        const int32_t it = (z == 0) ? fi : checked_id(prm.sample_item[p * prm.S + (z - 1)], prm.n_items, prm.err_flag);
        if (warp == 0) u_row_s[lane] = u;
        item_ptr = prm.E_item + (size_t)it * D;
        feat_ptr = prm.Feat + (size_t)fi * prm.F;
    }
    const float* noise_ptr = (NOISE_MODE == 1) ? prm.noise + (size_t)my_row * prm.F : nullptr;
    const uint32_t my_row32 = (uint32_t)my_row;

    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

    const int n_chunks = (D + prm.F) / FWD_KC;
    const int per_warp = n_chunks / SK_WARPS;          // (D+F)/16 is a multiple of 4 because F % 64 == 0
    const int c_lo = warp * per_warp, c_hi = c_lo + per_warp;

    float4 a_reg[4], n_reg[4], w_reg[8];
    auto prefetch = [&](int c) {
        const int k0 = c * FWD_KC;
        const float* src = (k0 < D) ? item_ptr + k0 : feat_ptr + (k0 - D);
#pragma unroll
        for (int q = 0; q < 4; ++q) a_reg[q] = ldg4(src + 4 * q);
        if (NOISE_MODE == 1 && k0 >= D) {
#pragma unroll
            for (int q = 0; q < 4; ++q) n_reg[q] = ldg4(noise_ptr + (k0 - D) + 4 * q);
        }
        const float* wsrc = prm.Wt + (size_t)k0 * D;
#pragma unroll
        for (int q = 0; q < 8; ++q) w_reg[q] = ldg4(wsrc + (q * 32 + lane) * 4);
    };
    auto stage = [&](int c) {
        const int k0 = c * FWD_KC;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            float4 v = a_reg[q];
            if (NOISE_MODE != 0 && k0 >= D) {
                float4 e;
                if (NOISE_MODE == 1) e = n_reg[q];
                else e = noise_quad(key_noise, my_row32, (uint32_t)((k0 - D) / 4 + q), prm.noise_std);
                v.x = __fadd_rn(v.x, e.x);
                v.y = __fadd_rn(v.y, e.y);
                v.z = __fadd_rn(v.z, e.z);
                v.w = __fadd_rn(v.w, e.w);
            }
            As[(4 * q + 0) * SK_BM + lane] = v.x;
            As[(4 * q + 1) * SK_BM + lane] = v.y;
            As[(4 * q + 2) * SK_BM + lane] = v.z;
            As[(4 * q + 3) * SK_BM + lane] = v.w;
        }
#pragma unroll
        for (int q = 0; q < 8; ++q) st4(Ws + (q * 32 + lane) * 4, w_reg[q]);
    };

    prefetch(c_lo);
    stage(c_lo);
    __syncwarp();
    for (int c = c_lo; c < c_hi; ++c) {
        if (c + 1 < c_hi) prefetch(c + 1);
#pragma unroll
        for (int k = 0; k < FWD_KC; ++k) {
            const float4 a0 = ld4(As + k * SK_BM + ty * 4);
            const float4 a1 = ld4(As + k * SK_BM + 16 + ty * 4);
            const float4 b0 = ld4(Ws + k * D + tx * 4);
            const float4 b1 = ld4(Ws + k * D + 32 + tx * 4);
            const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncwarp();
        if (c + 1 < c_hi) stage(c + 1);
        __syncwarp();
    }

    // ---- cross-warp reduction of the 4 partial tiles (staging buffers are dead now) ---------------
    __syncthreads();
    float* part = smem + warp * (SK_BM * D);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int m = (i < 4) ? (ty * 4 + i) : (16 + ty * 4 + (i - 4));
        st4(part + m * D + tx * 4, make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]));
        st4(part + m * D + 32 + tx * 4, make_float4(acc[i][4], acc[i][5], acc[i][6], acc[i][7]));
    }
    __syncthreads();

    // ---- epilogue: warp w finishes rows w*8 .. w*8+7; lane owns columns 2*lane, 2*lane+1 -------------
    const int c2 = lane * 2;
    const float2 bia = __ldg(reinterpret_cast<const float2*>(prm.bias + c2));
#pragma unroll 2
    for (int rr = 0; rr < SK_BM / SK_WARPS; ++rr) {
        const int m = warp * (SK_BM / SK_WARPS) + rr;
        const int64_t grow_raw = row_base + m;
        const bool valid = grow_raw < prm.n_rows;
        const int64_t grow = valid ? grow_raw : prm.n_rows - 1;
        float2 pre = make_float2(0.f, 0.f);
#pragma unroll
        for (int w = 0; w < SK_WARPS; ++w) {
            const float2 v = *reinterpret_cast<const float2*>(smem + w * (SK_BM * D) + m * D + c2);
            pre.x += v.x;
            pre.y += v.y;
        }
        float hx = fmaxf(pre.x + bia.x, 0.f), hy = fmaxf(pre.y + bia.y, 0.f);
        if (prm.mask_mode == 1) {
            const float2 mk = __ldg(reinterpret_cast<const float2*>(prm.mask + (size_t)grow * D + c2));
            hx *= mk.x;
            hy *= mk.y;
        } else if (prm.mask_mode == 2) {
            const float4 mk = dropout_quad(key_drop, (uint32_t)grow, (uint32_t)(lane >> 1), prm.keep_prob, prm.drop_scale);
            hx *= (lane & 1) ? mk.z : mk.x;
            hy *= (lane & 1) ? mk.w : mk.y;
        }
        if (prm.save_h != nullptr && valid)
            *reinterpret_cast<float2*>(prm.save_h + (size_t)grow * D + c2) = make_float2(hx, hy);
        const float2 eu = __ldg(reinterpret_cast<const float2*>(prm.E_user + (size_t)u_row_s[m] * D + c2));
        float part_dot = fmaf(hy, eu.y, hx * eu.x);
        part_dot = warp_sum(part_dot);
        if (lane == 0 && valid) prm.ws_rows[grow] = part_dot;
    }
}

// pred[p] = (1/A) sum_z softmax_z(expo[u, item_z]) sum_a s[p,z,a]   — one warp per pair, shuffle reductions
// (body in backdoor.cuh, shared with the fused epilogue of the tensor-core training forward).
__global__ void __launch_bounds__(256) k_backdoor(const dccf_expo ex, const int64_t* __restrict__ X,
                                                  const int64_t* __restrict__ sample_item, int64_t n_pairs,
                                                  int32_t n_users, int32_t user_base, int32_t n_items, int32_t S, int32_t A,
                                                  const float* __restrict__ ws_rows, float* __restrict__ out_pred,
                                                  float* __restrict__ save_w, int32_t* err_flag) {
    const int lane = threadIdx.x & 31;
    const int64_t p = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (p >= n_pairs) return;  // warp-uniform
    backdoor_pair(ex, X, sample_item, p, lane, n_users, user_base, n_items, S, A, ws_rows + p * ((S + 1) * A), out_pred + p,
                  save_w != nullptr ? save_w + p * (S + 1) : nullptr, err_flag);
}

void launch_transpose_w(const float* W, float* Wt, int K, cudaStream_t stream) {
    k_transpose_w<<<dim3((K + 31) / 32, D / 32), dim3(32, 8), 0, stream>>>(W, Wt, K);
}

void launch_backdoor(const dccf_expo* expo, const int64_t* X, const int64_t* sample_item, int64_t n_pairs,
                     const dccf_dims* dims, const float* ws_rows, float* out_pred, float* save_w, int32_t* err_flag,
                     cudaStream_t stream) {
    const int warps_per_cta = 8;
    k_backdoor<<<(unsigned)((n_pairs + warps_per_cta - 1) / warps_per_cta), warps_per_cta * 32, 0, stream>>>(
        *expo, X, sample_item, n_pairs, dims->n_users, dims->user_base, dims->n_items, dims->n_samples, dims->n_attr, ws_rows, out_pred,
        save_w, err_flag);
}

}  // namespace dccf

using namespace dccf;

extern "C" int dccf_score_fwd(const dccf_dims* dims, const float* E_user, const float* E_item, const float* Feat,
                              const float* W, const float* b, const dccf_expo* expo, const int64_t* X,
                              const int64_t* sample_item, int64_t n_pairs, const dccf_rng* rng, float* out_pred,
                              float* ws_rows, float* ws_wt, float* save_h, float* save_w, int32_t* err_flag,
                              void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    DCCF_CHECK_ARG(dims && expo && rng, "dccf_score_fwd: null struct argument");
    DCCF_CHECK_ARG(dims->dim == D, "dccf_score_fwd: dim=%d but this build has D=%d", dims->dim, D);
    DCCF_CHECK_ARG(dims->feat_dim > 0 && dims->feat_dim % 64 == 0, "dccf_score_fwd: feat_dim=%d must be a positive multiple of 64", dims->feat_dim);
    DCCF_CHECK_ARG(dims->n_samples >= 0 && dims->n_attr >= 1, "dccf_score_fwd: bad n_samples/n_attr");
    DCCF_CHECK_ARG(E_user && E_item && Feat && W && b && X && out_pred && ws_rows && ws_wt, "dccf_score_fwd: null buffer");
    DCCF_CHECK_ARG(dims->n_samples == 0 || sample_item, "dccf_score_fwd: sample_item is null");
    DCCF_CHECK_ARG(rng->noise_mode >= 0 && rng->noise_mode <= 2 && rng->mask_mode >= 0 && rng->mask_mode <= 2, "dccf_score_fwd: bad rng mode");
    DCCF_CHECK_ARG(rng->noise_mode != 1 || rng->noise, "dccf_score_fwd: noise_mode 1 needs a noise tensor");
    DCCF_CHECK_ARG(rng->mask_mode != 1 || rng->mask, "dccf_score_fwd: mask_mode 1 needs a mask tensor");
    DCCF_CHECK_ARG(expo->mode == 0 ? expo->dense != nullptr
                                   : (expo->mode == 1 && expo->mf_user && expo->mf_item && expo->mf_user_bias && expo->mf_item_bias && expo->propensity),
                   "dccf_score_fwd: exposure source incomplete (mode %d)", expo->mode);
    if (n_pairs <= 0) return DCCF_OK;
    const int Z = dims->n_samples + 1, R = Z * dims->n_attr;
    const int64_t n_rows = n_pairs * R;
    DCCF_CHECK_ARG(n_rows < (int64_t)1 << 31, "dccf_score_fwd: %lld rows in one call (max 2^31-1); split the batch", (long long)n_rows);

    const int K = D + dims->feat_dim;
    k_transpose_w<<<dim3((K + 31) / 32, D / 32), dim3(32, 8), 0, stream>>>(W, ws_wt, K);
    DCCF_CHECK_LAUNCH("k_transpose_w");

    FwdParams prm;
    prm.E_user = E_user; prm.E_item = E_item; prm.Feat = Feat; prm.Wt = ws_wt; prm.bias = b;
    prm.X = X; prm.sample_item = sample_item; prm.noise = rng->noise; prm.mask = rng->mask;
    prm.ws_rows = ws_rows; prm.save_h = save_h; prm.err_flag = err_flag; prm.n_rows = n_rows;
    prm.n_users = dims->n_users; prm.user_base = dims->user_base; prm.n_items = dims->n_items; prm.F = dims->feat_dim;
    prm.S = dims->n_samples; prm.A = dims->n_attr; prm.R = R;
    prm.mask_mode = rng->mask_mode;
    prm.noise_std = rng->noise_std;
    prm.keep_prob = 1.0f - rng->p_drop;
    prm.drop_scale = (rng->p_drop < 1.0f) ? 1.0f / (1.0f - rng->p_drop) : 0.0f;
    prm.rng.seed = rng->seed; prm.rng.offset = rng->offset; prm.rng.offset_dev = rng->offset_dev;

    // few rows (a training step): 32-row CTAs with K split over the warps fill the 148 SMs;
    // many rows (an evaluation batch): 128-row CTAs reuse each staged W chunk four times more
    const unsigned grid_big = (unsigned)((n_rows + FWD_BM - 1) / FWD_BM);
    if (grid_big < 2 * 148) {
        const unsigned grid = (unsigned)((n_rows + SK_BM - 1) / SK_BM);
        switch (rng->noise_mode) {
            case 0: k_row_scores_splitk<0><<<grid, SK_NT, 0, stream>>>(prm); break;
            case 1: k_row_scores_splitk<1><<<grid, SK_NT, 0, stream>>>(prm); break;
            default: k_row_scores_splitk<2><<<grid, SK_NT, 0, stream>>>(prm); break;
        }
        DCCF_CHECK_LAUNCH("k_row_scores_splitk");
    } else {
        switch (rng->noise_mode) {
            case 0: k_row_scores<0><<<grid_big, FWD_NT, 0, stream>>>(prm); break;
            case 1: k_row_scores<1><<<grid_big, FWD_NT, 0, stream>>>(prm); break;
            default: k_row_scores<2><<<grid_big, FWD_NT, 0, stream>>>(prm); break;
        }
        DCCF_CHECK_LAUNCH("k_row_scores");
    }

    const int warps_per_cta = 8;
    k_backdoor<<<(unsigned)((n_pairs + warps_per_cta - 1) / warps_per_cta), warps_per_cta * 32, 0, stream>>>(
        *expo, X, sample_item, n_pairs, dims->n_users, dims->user_base, dims->n_items, dims->n_samples, dims->n_attr, ws_rows, out_pred,
        save_w, err_flag);
    DCCF_CHECK_LAUNCH("k_backdoor");
    return DCCF_OK;
}
