// bpr_bwd.cu — kernel (c) part 1 of the DCCF hot path: pairwise loss + the whole backward pass in ONE launch.
//
// Replaces DCCF.forward lines 116-125 (src/models/DCCF.py) and what autograd does for
// `loss.backward()` (src/runners/BaseRunner.py:183) on the DCCF.predict graph (DCCF.py:74-100):
//
//   dpred[j]      = -(1 - sigmoid(pred[j]-pred[j+b])),  dpred[j+b] = -dpred[j]        (BPR, sum)
//   ds[p,z,a]     = dpred[p] * w[p,z] / A
//   gE_user[u_p] += sum_{z,a} ds * h[p,z,a,:]
//   dpre[p,z,a,:] = ds * E_user[u_p,:] * gate[p,z,a,:]           gate = dropout mask where h > 0
//   gb = sum dpre ; gW = sum dpre (x) [E_item[item_z] | Feat[i_p] + eps[p,z,a,:]]
//   gE_item[item_z] += W_i^T sum_a dpre[p,z,a,:]
//
// The grid is heterogeneous: CTAs [0, n_gw) each own (one 64-column chunk of W) x (one split of
// the N rows) and write a partial 64x64 tile of gW (no atomics — the Adam kernel sums the splits
// in a fixed order); CTAs [n_gw, n_gw+n_rec) produce the per-pair embedding-gradient records
// with warp-shuffle reductions; the last CTA computes the scalar loss.
#include "common.cuh"

namespace dccf {

constexpr int BWD_NT = 128;     // threads per CTA
constexpr int BWD_RC = 16;      // rows per staged chunk
constexpr int BWD_CW = 64;      // W columns per CTA
constexpr int BWD_WARPS = BWD_NT / 32;

struct BwdParams {
    const float* E_user;
    const float* E_item;
    const float* Feat;
    const float* W;  // [D, D+F] row-major
    const int64_t* X;
    const int64_t* sample_item;
    const float* Y;
    const float* noise;  // mode 1
    const float* mask;   // mode 1
    const float* pred;
    const float* save_h;
    const float* save_w;
    float* out_loss;
    float* gW_part;
    float* gb_part;
    float* gu_rec;
    float* gi_rec;
    int32_t* rec_keys_u;
    int32_t* rec_keys_i;
    float* dpre_rows;        // optional [N, D]: d loss / d pre-activation of every row (operand of the tensor-core dW kernel)
    int64_t n_pairs, n_rows;
    int32_t n_users, n_items, F, S, A, Z, R;
    int32_t user_base;
    int32_t noise_mode, mask_mode, loss_mode;
    int32_t n_chunks;        // (D+F)/64
    int32_t n_splits;        // row splits
    int32_t rows_per_split;  // multiple of BWD_RC
    int32_t n_gw, n_rec_ctas;
    float noise_std, drop_scale, inv_A;
    RngSpec rng;
};

// d loss / d pred[p]   (DCCF.py:116-125)
__device__ __forceinline__ float dpred_of(const BwdParams& prm, int64_t p) {
    if (prm.loss_mode == 0) {
        const int64_t b = prm.n_pairs >> 1;
        if (p >= 2 * b) return 0.f;  // odd tail never enters the loss
        const bool is_pos = p < b;
        const float d = is_pos ? (__ldg(prm.pred + p) - __ldg(prm.pred + p + b))
                               : (__ldg(prm.pred + p - b) - __ldg(prm.pred + p));
        const float sg = 1.f / (1.f + expf(-d));
        const float g = -(1.f - sg);
        return is_pos ? g : -g;
    }
    if (prm.loss_mode == 2) return __ldg(prm.Y + p);  // upstream gradient supplied by the caller (autograd)
    return 2.f * (__ldg(prm.pred + p) - __ldg(prm.Y + p)) / (float)prm.n_pairs;
}

// ---------------------------------------------------------------------------------------------
// role 1: partial gW tile.  CTA (chunk c, split s): acc[j][col] = sum_{r in split} dpre[r][j] * x[r][c*64+col]
// ---------------------------------------------------------------------------------------------
__device__ void bwd_gw_role(const BwdParams& prm, int chunk, int split) {
    __shared__ __align__(16) float dp_s[BWD_RC][D];
    __shared__ __align__(16) float x_s[BWD_RC][BWD_CW];

    const int tid = threadIdx.x;
    const int fr = tid >> 3;         // fill: row within chunk (0..15)
    const int fc = (tid & 7) * 4;    // fill: columns fc..fc+3 and 32+fc..32+fc+3 (conflict-free float4 stores)
    const int jg = (tid >> 4) * 8;   // compute: first of 8 j
    const int cg = (tid & 15) * 4;   // compute: first of 4 columns
    const RngKey key_noise = resolve_rng_key(prm.rng, DOMAIN_NOISE);

    float acc[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    float gb_acc = 0.f;

    const int64_t row_lo = (int64_t)split * prm.rows_per_split;
    const int64_t row_hi = min(row_lo + prm.rows_per_split, prm.n_rows);
    const int col0 = chunk * BWD_CW;  // column of W; < D means item-embedding part
    const bool item_part = col0 < D;
    const int f0 = col0 - D + fc;     // feature column of x0 (x1 is 32 further)

    // raw operands of one staged row, loaded a chunk ahead of their use
    struct Raw {
        float4 h0, h1, e0, e1, m0, m1, x0, x1, n0, n1;
        float ds;
        bool live;
    };
    auto fetch = [&](int64_t rb, Raw& w) {
        const int64_t r = rb + fr;
        w.live = r < row_hi;
        if (!w.live) return;
        const int64_t p = r / prm.R;
        const int rem = (int)(r - p * prm.R);
        const int z = rem / prm.A;
        const int32_t u = checked_id(prm.X[2 * p] - prm.user_base, prm.n_users, nullptr);
        const int32_t fi = checked_id(prm.X[2 * p + 1], prm.n_items, nullptr);
        w.ds = dpred_of(prm, p) * __ldg(prm.save_w + p * prm.Z + z) * prm.inv_A;
        w.h0 = ldg4(prm.save_h + (size_t)r * D + fc);
        w.h1 = ldg4(prm.save_h + (size_t)r * D + 32 + fc);
        w.e0 = ldg4(prm.E_user + (size_t)u * D + fc);
        w.e1 = ldg4(prm.E_user + (size_t)u * D + 32 + fc);
        if (prm.mask_mode == 1) {
            w.m0 = ldg4(prm.mask + (size_t)r * D + fc);
            w.m1 = ldg4(prm.mask + (size_t)r * D + 32 + fc);
        }
        if (item_part) {
            const int32_t it = (z == 0) ? fi : checked_id(prm.sample_item[p * prm.S + (z - 1)], prm.n_items, nullptr);
            w.x0 = ldg4(prm.E_item + (size_t)it * D + fc);
            w.x1 = ldg4(prm.E_item + (size_t)it * D + 32 + fc);
        } else {
            w.x0 = ldg4(prm.Feat + (size_t)fi * prm.F + f0);
            w.x1 = ldg4(prm.Feat + (size_t)fi * prm.F + f0 + 32);
            if (prm.noise_mode == 1) {
                w.n0 = ldg4(prm.noise + (size_t)r * prm.F + f0);
                w.n1 = ldg4(prm.noise + (size_t)r * prm.F + f0 + 32);
            }
        }
    };
    auto gate = [&](const float4& h, const float4& m) {
        if (prm.mask_mode == 1)
            return make_float4(h.x > 0.f ? m.x : 0.f, h.y > 0.f ? m.y : 0.f, h.z > 0.f ? m.z : 0.f, h.w > 0.f ? m.w : 0.f);
        const float s = (prm.mask_mode == 2) ? prm.drop_scale : 1.f;
        return make_float4(h.x > 0.f ? s : 0.f, h.y > 0.f ? s : 0.f, h.z > 0.f ? s : 0.f, h.w > 0.f ? s : 0.f);
    };
    auto add4 = [](float4& a, const float4& b) {
        a.x = __fadd_rn(a.x, b.x); a.y = __fadd_rn(a.y, b.y); a.z = __fadd_rn(a.z, b.z); a.w = __fadd_rn(a.w, b.w);
    };
    auto store = [&](int64_t rb, const Raw& w) {
        float4 d0 = make_float4(0.f, 0.f, 0.f, 0.f), d1 = d0, x0 = d0, x1 = d0;
        if (w.live) {
            const float4 g0 = gate(w.h0, w.m0), g1 = gate(w.h1, w.m1);
            d0 = make_float4(w.ds * w.e0.x * g0.x, w.ds * w.e0.y * g0.y, w.ds * w.e0.z * g0.z, w.ds * w.e0.w * g0.w);
            d1 = make_float4(w.ds * w.e1.x * g1.x, w.ds * w.e1.y * g1.y, w.ds * w.e1.z * g1.z, w.ds * w.e1.w * g1.w);
            x0 = w.x0;
            x1 = w.x1;
            if (!item_part) {
                if (prm.noise_mode == 1) {
                    add4(x0, w.n0);
                    add4(x1, w.n1);
                } else if (prm.noise_mode == 2) {
                    const uint32_t r32 = (uint32_t)(rb + fr);
                    add4(x0, noise_quad(key_noise, r32, (uint32_t)(f0 / 4), prm.noise_std));
                    add4(x1, noise_quad(key_noise, r32, (uint32_t)(f0 / 4 + 8), prm.noise_std));
                }
            }
        }
        st4(&dp_s[fr][fc], d0);
        st4(&dp_s[fr][32 + fc], d1);
        st4(&x_s[fr][fc], x0);
        st4(&x_s[fr][32 + fc], x1);
    };

    Raw cur;
    cur.live = false;
    if (row_lo < row_hi) fetch(row_lo, cur);
    for (int64_t rb = row_lo; rb < row_hi; rb += BWD_RC) {
        __syncthreads();  // previous chunk fully consumed
        store(rb, cur);
        __syncthreads();
        if (rb + BWD_RC < row_hi) fetch(rb + BWD_RC, cur);   // in flight during the FMAs below

        // ---- rank-16 update of the 64x64 tile -----------------------------------------------
#pragma unroll
        for (int k = 0; k < BWD_RC; ++k) {
            const float4 a0 = ld4(&dp_s[k][jg]);
            const float4 a1 = ld4(&dp_s[k][jg + 4]);
            const float4 b0 = ld4(&x_s[k][cg]);
            const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const float b[4] = {b0.x, b0.y, b0.z, b0.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        if (chunk == 0 && tid < D) {
#pragma unroll
            for (int k = 0; k < BWD_RC; ++k) gb_acc += dp_s[k][tid];
        }
    }

    const int K = D + prm.F;
    float* out = prm.gW_part + (size_t)split * D * K;
#pragma unroll
    for (int i = 0; i < 8; ++i)
        st4(out + (size_t)(jg + i) * K + col0 + cg, make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]));
    if (chunk == 0 && tid < D) prm.gb_part[(size_t)split * D + tid] = gb_acc;
}

// ---------------------------------------------------------------------------------------------
// role 2: embedding-gradient records.  One warp per pair.
// ---------------------------------------------------------------------------------------------
__device__ void bwd_record_role(const BwdParams& prm, int cta) {
    __shared__ __align__(16) float Wi_s[D][D];  // Wi_s[j][k] = W[j][k], k < D

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int K = D + prm.F;
    for (int i = tid; i < D * D / 4; i += BWD_NT) {
        const int j = i / (D / 4), q = i % (D / 4);
        st4(&Wi_s[j][q * 4], ldg4(prm.W + (size_t)j * K + q * 4));
    }
    __syncthreads();

    const int64_t p = (int64_t)cta * BWD_WARPS + warp;
    if (p >= prm.n_pairs) return;
    const int32_t u = checked_id(prm.X[2 * p] - prm.user_base, prm.n_users, nullptr);
    const int32_t fi = checked_id(prm.X[2 * p + 1], prm.n_items, nullptr);
    const float dp = dpred_of(prm, p);
    const int c = lane * 2;  // this lane's two columns
    const float2 eu = __ldg(reinterpret_cast<const float2*>(prm.E_user + (size_t)u * D + c));

    float2 gu = make_float2(0.f, 0.f);
    for (int z = 0; z < prm.Z; ++z) {
        const float ds = dp * __ldg(prm.save_w + p * prm.Z + z) * prm.inv_A;
        float2 dsum = make_float2(0.f, 0.f);
        for (int a = 0; a < prm.A; ++a) {
            const int64_t r = (p * prm.Z + z) * prm.A + a;
            const float2 h = __ldg(reinterpret_cast<const float2*>(prm.save_h + (size_t)r * D + c));
            gu.x = fmaf(ds, h.x, gu.x);
            gu.y = fmaf(ds, h.y, gu.y);
            float gx, gy;
            if (prm.mask_mode == 1) {
                const float2 m = __ldg(reinterpret_cast<const float2*>(prm.mask + (size_t)r * D + c));
                gx = h.x > 0.f ? m.x : 0.f;
                gy = h.y > 0.f ? m.y : 0.f;
            } else {
                const float s = (prm.mask_mode == 2) ? prm.drop_scale : 1.f;
                gx = h.x > 0.f ? s : 0.f;
                gy = h.y > 0.f ? s : 0.f;
            }
            const float dx = ds * eu.x * gx, dy = ds * eu.y * gy;
            dsum.x += dx;
            dsum.y += dy;
            if (prm.dpre_rows != nullptr) *reinterpret_cast<float2*>(prm.dpre_rows + (size_t)r * D + c) = make_float2(dx, dy);
        }
        // gi[k] = sum_j W[j][k] * dsum[j]
        float2 gi = make_float2(0.f, 0.f);
#pragma unroll 8
        for (int j = 0; j < D; j += 2) {
            const float dj0 = __shfl_sync(0xffffffffu, dsum.x, j >> 1);
            const float dj1 = __shfl_sync(0xffffffffu, dsum.y, j >> 1);
            const float2 w0 = *reinterpret_cast<const float2*>(&Wi_s[j][c]);
            const float2 w1 = *reinterpret_cast<const float2*>(&Wi_s[j + 1][c]);
            gi.x = fmaf(w0.x, dj0, gi.x);
            gi.y = fmaf(w0.y, dj0, gi.y);
            gi.x = fmaf(w1.x, dj1, gi.x);
            gi.y = fmaf(w1.y, dj1, gi.y);
        }
        const int64_t rec = p * prm.Z + z;
        *reinterpret_cast<float2*>(prm.gi_rec + (size_t)rec * D + c) = gi;
        if (lane == 0) {
            const int32_t it = (z == 0) ? fi : checked_id(prm.sample_item[p * prm.S + (z - 1)], prm.n_items, nullptr);
            prm.rec_keys_i[rec] = it;
        }
    }
    *reinterpret_cast<float2*>(prm.gu_rec + (size_t)p * D + c) = gu;
    if (lane == 0) prm.rec_keys_u[p] = u;
}

// ---------------------------------------------------------------------------------------------
// role 3: scalar loss, fixed summation order (one warp, lane-strided then shuffle tree)
// ---------------------------------------------------------------------------------------------
__device__ void bwd_loss_role(const BwdParams& prm) {
    if (threadIdx.x >= 32 || prm.loss_mode == 2 || prm.out_loss == nullptr) return;
    const int lane = threadIdx.x;
    float s = 0.f;
    if (prm.loss_mode == 0) {
        const int64_t b = prm.n_pairs >> 1;
        for (int64_t j = lane; j < b; j += 32) {
            const float d = __ldg(prm.pred + j) - __ldg(prm.pred + j + b);
            // -log(sigmoid(d)) = softplus(-d), evaluated the stable way
            s += (d > 0.f) ? log1pf(expf(-d)) : (-d + log1pf(expf(d)));
        }
    } else {
        for (int64_t p = lane; p < prm.n_pairs; p += 32) {
            const float d = __ldg(prm.pred + p) - __ldg(prm.Y + p);
            s = fmaf(d, d, s);
        }
    }
    s = warp_sum(s);
    if (lane == 0) prm.out_loss[0] = (prm.loss_mode == 0) ? s : s / (float)prm.n_pairs;
}

__global__ void __launch_bounds__(BWD_NT) k_bpr_bwd(const BwdParams prm) {
    const int b = blockIdx.x;
    if (b < prm.n_gw) {
        bwd_gw_role(prm, b % prm.n_chunks, b / prm.n_chunks);
    } else if (b < prm.n_gw + prm.n_rec_ctas) {
        bwd_record_role(prm, b - prm.n_gw);
    } else {
        bwd_loss_role(prm);
    }
}

// number of row splits for N rows: enough CTAs to cover the 148 SMs about twice, at least 64 rows each
// (four per SM was measured: no faster, and the Adam kernel then reads twice the partial sums)
int bpr_bwd_launch(const dccf_dims* dims, const float* E_user, const float* E_item, const float* Feat, const float* W,
                   const int64_t* X, const int64_t* sample_item, const float* Y, int64_t n_pairs, const dccf_rng* rng,
                   int32_t loss_mode, const float* pred, const float* save_h, const float* save_w, float* out_loss,
                   float* gW_part, float* gb_part, float* gu_rec, float* gi_rec, int32_t* rec_keys_u, int32_t* rec_keys_i,
                   float* dpre_rows, bool with_gw, cudaStream_t stream);

static int32_t bwd_splits_for(int64_t n_rows, int n_chunks) {
    if (n_rows <= 0) return 1;
    const int64_t max_by_rows = (n_rows + 63) / 64;
    int64_t want = (2 * 148 + n_chunks - 1) / n_chunks;
    if (want > max_by_rows) want = max_by_rows;
    if (want < 1) want = 1;
    return (int32_t)want;
}

}  // namespace dccf

using namespace dccf;

extern "C" int32_t dccf_bwd_splits(int64_t n_rows) {
    // independent of F so that callers can size buffers before knowing the chunk count: use the
    // reference geometry (F = 768 -> 13 chunks); any F only changes the CTA count, not the result layout
    return bwd_splits_for(n_rows, 13);
}

// with_gw = false: only the record and loss roles run (the tensor-core path computes dW / db itself from
// dpre_rows, see tc_train.cu); gW_part / gb_part may then be null.
int dccf::bpr_bwd_launch(const dccf_dims* dims, const float* E_user, const float* E_item, const float* Feat,
                         const float* W, const int64_t* X, const int64_t* sample_item, const float* Y,
                         int64_t n_pairs, const dccf_rng* rng, int32_t loss_mode, const float* pred,
                         const float* save_h, const float* save_w, float* out_loss, float* gW_part, float* gb_part,
                         float* gu_rec, float* gi_rec, int32_t* rec_keys_u, int32_t* rec_keys_i, float* dpre_rows,
                         bool with_gw, cudaStream_t stream) {
    DCCF_CHECK_ARG(dims && rng, "dccf_bpr_bwd: null struct argument");
    DCCF_CHECK_ARG(dims->dim == D, "dccf_bpr_bwd: dim=%d but this build has D=%d", dims->dim, D);
    DCCF_CHECK_ARG(dims->feat_dim > 0 && dims->feat_dim % 64 == 0, "dccf_bpr_bwd: feat_dim=%d must be a positive multiple of 64", dims->feat_dim);
    DCCF_CHECK_ARG(dims->n_samples >= 0 && dims->n_attr >= 1, "dccf_bpr_bwd: bad n_samples/n_attr");
    DCCF_CHECK_ARG(loss_mode >= 0 && loss_mode <= 2, "dccf_bpr_bwd: loss_mode must be 0 (BPR), 1 (MSE) or 2 (external d loss/d pred in Y)");
    DCCF_CHECK_ARG(E_user && E_item && Feat && W && X && pred && save_h && save_w && gu_rec && gi_rec && rec_keys_u &&
                       rec_keys_i && (!with_gw || (gW_part && gb_part)),
                   "dccf_bpr_bwd: null buffer");
    DCCF_CHECK_ARG(loss_mode == 2 || out_loss, "dccf_bpr_bwd: out_loss is null");
    DCCF_CHECK_ARG(dims->n_samples == 0 || sample_item, "dccf_bpr_bwd: sample_item is null");
    DCCF_CHECK_ARG(loss_mode == 0 || Y, "dccf_bpr_bwd: loss_mode %d needs Y", loss_mode);
    DCCF_CHECK_ARG(rng->noise_mode >= 0 && rng->noise_mode <= 2 && rng->mask_mode >= 0 && rng->mask_mode <= 2, "dccf_bpr_bwd: bad rng mode");
    DCCF_CHECK_ARG(rng->noise_mode != 1 || rng->noise, "dccf_bpr_bwd: noise_mode 1 needs the noise tensor");
    DCCF_CHECK_ARG(rng->mask_mode != 1 || rng->mask, "dccf_bpr_bwd: mask_mode 1 needs the mask tensor");
    DCCF_CHECK_ARG(loss_mode != 0 || n_pairs % 2 == 0, "dccf_bpr_bwd: BPR needs an even number of pairs, got %lld", (long long)n_pairs);
    if (n_pairs <= 0) return DCCF_OK;

    BwdParams prm;
    prm.E_user = E_user; prm.E_item = E_item; prm.Feat = Feat; prm.W = W; prm.X = X; prm.sample_item = sample_item;
    prm.Y = Y; prm.noise = rng->noise; prm.mask = rng->mask; prm.pred = pred; prm.save_h = save_h; prm.save_w = save_w;
    prm.out_loss = out_loss; prm.gW_part = gW_part; prm.gb_part = gb_part; prm.gu_rec = gu_rec; prm.gi_rec = gi_rec;
    prm.rec_keys_u = rec_keys_u; prm.rec_keys_i = rec_keys_i; prm.dpre_rows = dpre_rows;
    prm.n_pairs = n_pairs;
    prm.n_users = dims->n_users; prm.user_base = dims->user_base; prm.n_items = dims->n_items; prm.F = dims->feat_dim; prm.S = dims->n_samples;
    prm.A = dims->n_attr; prm.Z = dims->n_samples + 1; prm.R = prm.Z * prm.A;
    prm.n_rows = n_pairs * prm.R;
    DCCF_CHECK_ARG(prm.n_rows < (int64_t)1 << 31, "dccf_bpr_bwd: %lld rows in one call (max 2^31-1)", (long long)prm.n_rows);
    prm.noise_mode = rng->noise_mode; prm.mask_mode = rng->mask_mode; prm.loss_mode = loss_mode;
    prm.n_chunks = (D + prm.F) / BWD_CW;
    prm.n_splits = dccf_bwd_splits(prm.n_rows);
    const int64_t rps = (prm.n_rows + prm.n_splits - 1) / prm.n_splits;
    prm.rows_per_split = (int32_t)(((rps + BWD_RC - 1) / BWD_RC) * BWD_RC);
    prm.n_gw = with_gw ? prm.n_chunks * prm.n_splits : 0;
    prm.n_rec_ctas = (int32_t)((n_pairs + BWD_WARPS - 1) / BWD_WARPS);
    prm.noise_std = rng->noise_std;
    prm.drop_scale = (rng->p_drop < 1.0f) ? 1.0f / (1.0f - rng->p_drop) : 0.0f;
    prm.inv_A = 1.0f / (float)prm.A;
    prm.rng.seed = rng->seed; prm.rng.offset = rng->offset; prm.rng.offset_dev = rng->offset_dev;

    const unsigned grid = (unsigned)(prm.n_gw + prm.n_rec_ctas + 1);
    k_bpr_bwd<<<grid, BWD_NT, 0, stream>>>(prm);
    DCCF_CHECK_LAUNCH("k_bpr_bwd");
    return DCCF_OK;
}

extern "C" int dccf_bpr_bwd(const dccf_dims* dims, const float* E_user, const float* E_item, const float* Feat,
                            const float* W, const int64_t* X, const int64_t* sample_item, const float* Y,
                            int64_t n_pairs, const dccf_rng* rng, int32_t loss_mode, const float* pred,
                            const float* save_h, const float* save_w, float* out_loss, float* gW_part, float* gb_part,
                            float* gu_rec, float* gi_rec, int32_t* rec_keys_u, int32_t* rec_keys_i, void* stream_) {
    return bpr_bwd_launch(dims, E_user, E_item, Feat, W, X, sample_item, Y, n_pairs, rng, loss_mode, pred, save_h, save_w,
                          out_loss, gW_part, gb_part, gu_rec, gi_rec, rec_keys_u, rec_keys_i, nullptr, true,
                          (cudaStream_t)stream_);
}
