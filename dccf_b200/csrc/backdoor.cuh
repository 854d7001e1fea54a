// backdoor.cuh — the backdoor-adjusted sum of one (user, item) pair, shared by k_backdoor and the fused
// epilogue of the tensor-core training forward.
#pragma once
#include "common.cuh"

namespace dccf {

// ----------------------------------------------------------------------------------------------
// exposure value of (user u, item it):  expo_prob[u, it]  or the IPSBiasedMF formula
// (src/models/DCCF.py:98 lookup; src/models/IPSBiasedMF.py:42-53 on the fly)
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ float expo_value(const dccf_expo& ex, int32_t u, int32_t it, int32_t n_items) {
    if (ex.mode == 0) return __ldg(ex.dense + (size_t)u * n_items + it);
    const float* pu = ex.mf_user + (size_t)u * D;
    const float* qi = ex.mf_item + (size_t)it * D;
    float dot = 0.f;
#pragma unroll 4
    for (int k = 0; k < D; k += 4) {
        const float4 a = ldg4(pu + k), b = ldg4(qi + k);
        dot = fmaf(a.x, b.x, dot);
        dot = fmaf(a.y, b.y, dot);
        dot = fmaf(a.z, b.z, dot);
        dot = fmaf(a.w, b.w, dot);
    }
    const float pred = dot + __ldg(ex.mf_user_bias + u) + __ldg(ex.mf_item_bias + it) + ex.mf_global_bias;
    return pred / fmaxf(__ldg(ex.propensity + it), ex.mf_min_propensity);
}

// The backdoor-adjusted sum of one pair (src/models/DCCF.py:98-100), executed by one full warp, in two halves:
//   backdoor_weights: e[z] = exp(expo[u, item_z] - max_z), den = A * sum_z e[z]  — depends on the ids only
//   backdoor_apply:   pred = sum_{z,a} e[z] * s[z,a] / den ;  save_w[z] = e[z] * A / den
// backdoor_pair runs both.  The arithmetic (lane-strided accumulation, shuffle trees) is the same whichever way the
// halves are scheduled, so precomputing the weights in another kernel changes no bit of the result.
__device__ __forceinline__ void backdoor_weights(const dccf_expo& ex, const int64_t* X, const int64_t* sample_item,
                                                 int64_t p, int lane, int32_t n_users, int32_t user_base, int32_t n_items,
                                                 int32_t S, int32_t A, float* e_out, float* den_out, int32_t* err_flag) {
    const int Z = S + 1, R = Z * A;
    const int32_t u = checked_id(X[2 * p] - user_base, n_users, err_flag);
    float mx = -INFINITY;
    for (int l = lane; l < R; l += 32) {
        const int z = l / A;
        const int32_t it = checked_id(slot_item(X, sample_item, p, z, S), n_items, err_flag);
        mx = fmaxf(mx, expo_value(ex, u, it, n_items));
    }
    mx = warp_max(mx);
    float den = 0.f;
    for (int l = lane; l < R; l += 32) {
        const int z = l / A;
        const int32_t it = checked_id(slot_item(X, sample_item, p, z, S), n_items, err_flag);
        const float e = expf(expo_value(ex, u, it, n_items) - mx);
        den += e;
        if (l % A == 0) e_out[z] = e;
    }
    den = warp_sum(den);  // = A * sum_z exp(.)
    if (lane == 0) den_out[0] = den;
}

// `rows` = the R row scores of the pair (global or shared memory, plain loads: the caller may have written them
// earlier in the same kernel); e / den from backdoor_weights (visible to the whole warp).
__device__ __forceinline__ void backdoor_apply(const float* e, float den, int lane, int32_t S, int32_t A, const float* rows,
                                               float* out_pred, float* save_w) {
    const int Z = S + 1, R = Z * A;
    float num = 0.f;
    for (int l = lane; l < R; l += 32) num = fmaf(e[l / A], rows[l], num);
    num = warp_sum(num);
    if (lane == 0) out_pred[0] = num / den;
    if (save_w != nullptr)
        for (int z = lane; z < Z; z += 32) save_w[z] = e[z] * (float)A / den;
}

// Both halves in one call, for kernels that have no place to keep the weights between them (k_backdoor, the
// unfused epilogue): the exposure values are simply evaluated again for the second pass.
__device__ __forceinline__ void backdoor_pair(const dccf_expo& ex, const int64_t* X, const int64_t* sample_item,
                                              int64_t p, int lane, int32_t n_users, int32_t user_base, int32_t n_items,
                                              int32_t S, int32_t A, const float* rows, float* out_pred,
                                              float* save_w, int32_t* err_flag) {
    const int Z = S + 1, R = Z * A;
    const int32_t u = checked_id(X[2 * p] - user_base, n_users, err_flag);

    float mx = -INFINITY;
    for (int l = lane; l < R; l += 32) {
        const int z = l / A;
        const int32_t it = checked_id(slot_item(X, sample_item, p, z, S), n_items, err_flag);
        mx = fmaxf(mx, expo_value(ex, u, it, n_items));
    }
    mx = warp_max(mx);
    float num = 0.f, den = 0.f;
    for (int l = lane; l < R; l += 32) {
        const int z = l / A;
        const int32_t it = checked_id(slot_item(X, sample_item, p, z, S), n_items, err_flag);
        const float e = expf(expo_value(ex, u, it, n_items) - mx);
        num = fmaf(e, rows[l], num);
        den += e;
    }
    num = warp_sum(num);
    den = warp_sum(den);  // = A * sum_z exp(.)
    if (lane == 0) out_pred[0] = num / den;
    if (save_w != nullptr) {
        for (int l = lane; l < R; l += 32) {
            if (l % A == 0) {
                const int z = l / A;
                const int32_t it = checked_id(slot_item(X, sample_item, p, z, S), n_items, err_flag);
                const float e = expf(expo_value(ex, u, it, n_items) - mx);
                save_w[z] = e * (float)A / den;
            }
        }
    }
}

}  // namespace dccf
