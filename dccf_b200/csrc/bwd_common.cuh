// bwd_common.cuh — parameters and the loss derivative shared by the SIMT backward (bpr_bwd.cu) and the
// tensor-core dW_f kernel (tc_bwd.cu).
#pragma once
#include "common.cuh"

namespace dccf {

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
    int64_t n_pairs, n_rows;
    int32_t n_users, n_items, F, S, A, Z, R;
    int32_t noise_mode, mask_mode, loss_mode;
    int32_t n_chunks;        // (D+F)/64
    int32_t n_simt_chunks;   // chunks handled by k_bpr_bwd (1 when the feature columns go to the tensor cores)
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

}  // namespace dccf
