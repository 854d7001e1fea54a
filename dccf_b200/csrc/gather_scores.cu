// gather_scores.cu — noise-free scoring as a pure gather: the HBM/L2-bound regime of the scorer (SURVEY.md §8d, row
// "eval scorer, std = 0").
//
// With --std 0 and dropout 0 (every evaluation pass of a model trained without feature noise, and the parity
// configuration of the test-suite) the predictor of src/models/DCCF.py:84-100 loses its only per-row random input:
//     pre[p,z,a,:] = W_i·E_item[item(p,z)] + W_f·Feat[i_p] + b         (the same for every attribute copy a)
// and both terms are rows of the two projected tables dccf_tc_prepare already maintains for the tensor-core scorer,
//     PI [I,D] = E_item·W_i^T          PF [I,D] = Feat·W_f^T + b .
// A pair then costs Z + 2 row gathers of 256 B (its user row, the PF row of its item, one PI row per slot), Z exposure
// values, 64·Z max/fma — no contraction over the 832 inputs at all (2.35 MFLOP per pair in the general kernel).  The
// tables are 2 x 4 MB at the electronics shape and stay L2-resident; DRAM sees the ids and the exposure sectors.
//
//     s[p,z]  = < E_user[u_p], relu(PI[item(p,z)] + PF[i_p]) >
//     pred[p] = sum_z softmax_z(expo[u_p, item(p,z)]) · s[p,z]          (= the mean over the A identical copies)
//
// Layout: one pair per half-warp, lane = 4 of the 64 columns (one 128-bit load per row), the pair's slots 16 at a time:
// lane j of the half owns slot z0 + j (its item id, its exposure value, its score), the row gathers of four slots are
// in flight together, the dot products are reduced with xor-shuffles inside the half.  The softmax over the slots is
// the usual max-shifted one; for more than 16 slots it is carried across chunks (running max, rescaled sums).
#include "backdoor.cuh"

namespace dccf {

struct GatherParams {
    dccf_expo ex;
    const float* E_user;
    const float* PI;
    const float* PF;
    const int64_t* X;
    const int64_t* sample_item;
    float* out_pred;
    int32_t* err_flag;
    int64_t n_pairs;
    int32_t n_users, user_base, n_items, S;
};

__device__ __forceinline__ float half_sum(float v) {        // over the 16 lanes of a half-warp (xor < 16 stays inside)
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float half_max(float v) {
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// partial dot of this lane's four columns: < eu, relu(a + pf) >
__device__ __forceinline__ float relu_dot4(const float4& a, const float4& pf, const float4& eu) {
    float d = fmaxf(a.x + pf.x, 0.f) * eu.x;
    d = fmaf(fmaxf(a.y + pf.y, 0.f), eu.y, d);
    d = fmaf(fmaxf(a.z + pf.z, 0.f), eu.z, d);
    d = fmaf(fmaxf(a.w + pf.w, 0.f), eu.w, d);
    return d;
}

__global__ void __launch_bounds__(256) k_gather_scores(const GatherParams prm) {
    const int lane = threadIdx.x & 31, sub = lane & 15, half_base = lane & 16;
    const int64_t p_raw = ((int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * 2 + (lane >> 4);
    const bool active = p_raw < prm.n_pairs;
    const int64_t p = active ? p_raw : prm.n_pairs - 1;       // idle halves shadow the last pair: full-warp shuffles
    const int S = prm.S, Z = S + 1;

    const int32_t u = checked_id(prm.X[2 * p] - prm.user_base, prm.n_users, prm.err_flag);
    const int32_t fi = checked_id(prm.X[2 * p + 1], prm.n_items, prm.err_flag);
    const float4 eu = ldg4(prm.E_user + (size_t)u * D + 4 * sub);
    const float4 pf = ldg4(prm.PF + (size_t)fi * D + 4 * sub);

    float run_max = -INFINITY, num = 0.f, den = 0.f;
    for (int z0 = 0; z0 < Z; z0 += 16) {
        // lane `sub` owns slot z0 + sub: item id and exposure value
        const int z_mine = z0 + sub;
        int32_t it_mine = fi;
        float x_mine = -INFINITY;
        if (z_mine < Z) {
            if (z_mine > 0) it_mine = checked_id(prm.sample_item[p * S + (z_mine - 1)], prm.n_items, prm.err_flag);
            x_mine = expo_value(prm.ex, u, it_mine, prm.n_items);
        }
        // scores of the chunk's slots: four row gathers in flight, each reduced over the half
        const int nz = min(16, Z - z0);
        float s_mine = 0.f;
        for (int j0 = 0; j0 < nz; j0 += 4) {
            float4 a[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                // slots past the end repeat the chunk's first one (a valid row; the result is dropped)
                const int src = (j0 + q < nz) ? (j0 + q) : 0;
                const int32_t it = __shfl_sync(0xffffffffu, it_mine, half_base | src);
                a[q] = ldg4(prm.PI + (size_t)it * D + 4 * sub);
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float d = half_sum(relu_dot4(a[q], pf, eu));
                if (sub == j0 + q) s_mine = d;
            }
        }
        // max-shifted softmax over the slots, carried across chunks
        const float new_max = fmaxf(run_max, half_max(x_mine));
        const float rescale = (run_max == -INFINITY) ? 0.f : expf(run_max - new_max);
        const float e = (z_mine < Z) ? expf(x_mine - new_max) : 0.f;
        num = fmaf(num, rescale, half_sum(e * s_mine));
        den = fmaf(den, rescale, half_sum(e));
        run_max = new_max;
    }
    if (active && sub == 0) prm.out_pred[p_raw] = num / den;
}

}  // namespace dccf

using namespace dccf;

extern "C" int dccf_score_gather(const dccf_dims* dims, const float* E_user, const float* PI, const float* PF,
                                 const dccf_expo* expo, const int64_t* X, const int64_t* sample_item, int64_t n_pairs,
                                 float* out_pred, int32_t* err_flag, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    DCCF_CHECK_ARG(dims && expo, "dccf_score_gather: null struct argument");
    DCCF_CHECK_ARG(dims->dim == D, "dccf_score_gather: dim=%d but this build has D=%d", dims->dim, D);
    DCCF_CHECK_ARG(dims->n_samples >= 0, "dccf_score_gather: bad n_samples");
    DCCF_CHECK_ARG(E_user && PI && PF && X && out_pred, "dccf_score_gather: null buffer");
    DCCF_CHECK_ARG(dims->n_samples == 0 || sample_item, "dccf_score_gather: sample_item is null");
    DCCF_CHECK_ARG(expo->mode == 0 ? expo->dense != nullptr
                                   : (expo->mode == 1 && expo->mf_user && expo->mf_item && expo->mf_user_bias &&
                                      expo->mf_item_bias && expo->propensity),
                   "dccf_score_gather: exposure source incomplete (mode %d)", expo->mode);
    if (n_pairs <= 0) return DCCF_OK;
    GatherParams prm;
    prm.ex = *expo;
    prm.E_user = E_user; prm.PI = PI; prm.PF = PF; prm.X = X; prm.sample_item = sample_item;
    prm.out_pred = out_pred; prm.err_flag = err_flag; prm.n_pairs = n_pairs;
    prm.n_users = dims->n_users; prm.user_base = dims->user_base; prm.n_items = dims->n_items; prm.S = dims->n_samples;
    const int pairs_per_cta = 2 * (256 / 32);
    k_gather_scores<<<(unsigned)((n_pairs + pairs_per_cta - 1) / pairs_per_cta), 256, 0, stream>>>(prm);
    DCCF_CHECK_LAUNCH("k_gather_scores");
    return DCCF_OK;
}
