// adam.cu — kernel (c) part 2 of the DCCF hot path: l2 + clip + Adam, dense over every parameter.
//
// Replaces, per training step of the reference:
//   loss += l2 * model.l2()              src/runners/BaseRunner.py:181, src/models/BaseModel.py:179-187
//   clip_grad_value_(params, 50)          src/runners/BaseRunner.py:185
//   Adam(lr, weight_decay=l2).step()      src/runners/BaseRunner.py:100,187  (torch 2.11 optim/adam.py,
//                                         _multi_tensor_adam, non-capturable branch)
// which together put a non-zero gradient on EVERY row of both embedding tables each step, so the
// update is a full streaming sweep (24 B/param: read p,m,v, write p,m,v).  The sparse part of the
// gradient arrives as records (row id + 64 floats) linked per row through head/next; records of one
// row are summed in ascending record order so the result does not depend on atomics timing.
#include "common.cuh"

namespace dccf {

struct AdamScalars {
    float w1;          // 1 - beta1 (lerp weight)
    float beta2;       // beta2
    float omb2;        // 1 - beta2
    float step_size;   // -(lr / (1 - beta1^t))
    float bc2_sqrt;    // sqrt(1 - beta2^t)
    float eps;
    float two_l2;      // 2 * l2
    float wd;          // weight decay
    float clip;        // <= 0: off
};

struct AdamHost {
    double lr, beta1, beta2;
    float eps, l2, wd, clip;
    int32_t step;
    const int32_t* step_dev;
};

__device__ __forceinline__ AdamScalars resolve_adam(const AdamHost& h) {
    const int32_t t = (h.step_dev != nullptr) ? __ldg(h.step_dev) : h.step;
    AdamScalars s;
    s.w1 = (float)(1.0 - h.beta1);
    s.beta2 = (float)h.beta2;
    s.omb2 = (float)(1.0 - h.beta2);
    const double bc1 = 1.0 - pow(h.beta1, (double)t);
    const double bc2 = 1.0 - pow(h.beta2, (double)t);
    s.step_size = (float)(-(h.lr / bc1));
    s.bc2_sqrt = (float)sqrt(bc2);
    s.eps = h.eps;
    s.two_l2 = 2.0f * h.l2;
    s.wd = h.wd;
    s.clip = h.clip;
    return s;
}

// one element of the update; g_sparse is the data-dependent part of the gradient
__device__ __forceinline__ void adam_elem(float& p, float& m, float& v, float g_sparse, const AdamScalars& s) {
    float g = __fadd_rn(g_sparse, __fmul_rn(s.two_l2, p));       // + d/dp (l2 * sum p^2)
    if (s.clip > 0.f) g = fminf(fmaxf(g, -s.clip), s.clip);      // clip_grad_value_
    g = fmaf(s.wd, p, g);                                        // Adam weight_decay: g + wd * p
    m = fmaf(s.w1, __fsub_rn(g, m), m);                          // exp_avg.lerp_(g, 1 - beta1)
    v = __fmul_rn(v, s.beta2);                                   // exp_avg_sq.mul_(beta2)
    v = fmaf(__fmul_rn(s.omb2, g), g, v);                        //   .addcmul_(g, g, value = 1 - beta2)
    const float denom = __fadd_rn(__fdiv_rn(sqrtf(v), s.bc2_sqrt), s.eps);
    p = fmaf(s.step_size, __fdiv_rn(m, denom), p);               // param.addcdiv_(exp_avg, denom, value = step_size)
}

// Records may arrive in n_seg segments of seg_len records each (one segment per data-parallel rank after the
// all-gather): record r lives in segment r / seg_len at local index r % seg_len.
struct RecLayout {
    int64_t seg_len;
    int64_t key_seg_stride;   // int32 elements between the key arrays of consecutive segments
    int64_t grad_seg_stride;  // floats between the gradient arrays of consecutive segments
};
__device__ __forceinline__ size_t rec_key_index(const RecLayout& L, int64_t r) {
    const int64_t seg = r / L.seg_len;
    return (size_t)(seg * L.key_seg_stride + (r - seg * L.seg_len));
}
__device__ __forceinline__ size_t rec_grad_index(const RecLayout& L, int64_t r) {
    const int64_t seg = r / L.seg_len;
    return (size_t)(seg * L.grad_seg_stride + (r - seg * L.seg_len) * D);
}

__global__ void k_link_records(const int32_t* __restrict__ keys, int64_t n_rec, int64_t n_rows, RecLayout L,
                               int32_t* __restrict__ head, int32_t* __restrict__ next) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_rec) return;
    const int32_t k = keys[rec_key_index(L, r)];
    if (k < 0 || k >= n_rows) {
        next[r] = -1;
        return;
    }
    next[r] = atomicExch(&head[k], (int32_t)r);
}

// 16 lanes per table row (one float4 each); a warp covers two consecutive rows = 512 contiguous bytes.
__global__ void __launch_bounds__(256) k_adam_sweep(float* __restrict__ table, float* __restrict__ m_,
                                                    float* __restrict__ v_, int64_t n_rows,
                                                    const float* __restrict__ rec_grads, RecLayout L,
                                                    int32_t* __restrict__ head, const int32_t* __restrict__ next,
                                                    const AdamHost hp) {
    const AdamScalars s = resolve_adam(hp);
    const int sub = threadIdx.x & 15;
    const int half = (threadIdx.x >> 4) & 1;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    // warp-uniform trip count: a warp owns rows 2*w and 2*w+1
    for (int64_t w = warp; 2 * w < n_rows; w += n_warps) {
        const int64_t row = 2 * w + half;
        const bool valid = row < n_rows;
        const size_t off = (size_t)(valid ? row : 2 * w) * D + sub * 4;
        float4 p = ld4(table + off);
        float4 m = ld4(m_ + off);
        float4 v = ld4(v_ + off);
        float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
        const int32_t h = valid ? head[row] : -1;
        __syncwarp();  // every lane holds its list head before any lane resets it
        if (h >= 0) {
            // ascending-index traversal of this row's (unordered) record list
            int32_t last = -1;
            while (true) {
                int32_t best = 0x7fffffff;
                for (int32_t r = h; r >= 0; r = __ldg(next + r))
                    if (r > last && r < best) best = r;
                if (best == 0x7fffffff) break;
                const float4 rg = ldg4(rec_grads + rec_grad_index(L, best) + sub * 4);
                g.x += rg.x; g.y += rg.y; g.z += rg.z; g.w += rg.w;
                last = best;
            }
            if (sub == 0) head[row] = -1;
        }
        if (valid) {
            adam_elem(p.x, m.x, v.x, g.x, s);
            adam_elem(p.y, m.y, v.y, g.y, s);
            adam_elem(p.z, m.z, v.z, g.z, s);
            adam_elem(p.w, m.w, v.w, g.w, s);
            st4(table + off, p);
            st4(m_ + off, m);
            st4(v_ + off, v);
        }
    }
}

__global__ void __launch_bounds__(256) k_adam_dense(float* __restrict__ p_, float* __restrict__ m_,
                                                    float* __restrict__ v_, int64_t n,
                                                    const float* __restrict__ g_parts, int32_t n_parts,
                                                    int64_t part_stride, const AdamHost hp) {
    const AdamScalars s = resolve_adam(hp);
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        float g = 0.f;
        for (int32_t k = 0; k < n_parts; ++k) g += __ldg(g_parts + (size_t)k * part_stride + i);
        float p = p_[i], m = m_[i], v = v_[i];
        adam_elem(p, m, v, g, s);
        p_[i] = p; m_[i] = m; v_[i] = v;
    }
}

// out[i] = sum_k parts[k * stride + i], ascending k (fixed order)
__global__ void __launch_bounds__(256) k_sum_parts(const float* __restrict__ parts, int32_t n_parts, int64_t stride,
                                                   int64_t n, float* __restrict__ out) {
    const int64_t step = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += step) {
        float g = 0.f;
        for (int32_t k = 0; k < n_parts; ++k) g += __ldg(parts + (size_t)k * stride + i);
        out[i] = g;
    }
}

__global__ void k_state_advance(int32_t* step_dev, uint64_t* offset_dev, uint64_t inc) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        if (step_dev) step_dev[0] += 1;
        if (offset_dev) offset_dev[0] += inc;
    }
}

static int check_hp(const dccf_adam* hp, AdamHost* out, const char* who) {
    DCCF_CHECK_ARG(hp != nullptr, "%s: null hyper-parameter struct", who);
    DCCF_CHECK_ARG(hp->step_dev != nullptr || hp->step >= 1, "%s: step must be >= 1 (got %d)", who, hp->step);
    DCCF_CHECK_ARG(hp->beta1 >= 0.0 && hp->beta1 < 1.0 && hp->beta2 >= 0.0 && hp->beta2 < 1.0, "%s: betas out of range", who);
    out->lr = hp->lr; out->beta1 = hp->beta1; out->beta2 = hp->beta2;
    out->eps = (float)hp->eps; out->l2 = (float)hp->l2; out->wd = (float)hp->weight_decay; out->clip = (float)hp->clip;
    out->step = hp->step; out->step_dev = hp->step_dev;
    return DCCF_OK;
}

}  // namespace dccf

using namespace dccf;

extern "C" int dccf_adam_sweep_seg(float* table, float* m, float* v, int64_t n_table_rows, const int32_t* rec_keys,
                                   const float* rec_grads, int32_t n_seg, int64_t seg_len, int64_t key_seg_stride,
                                   int64_t grad_seg_stride, int32_t* head, int32_t* next, const dccf_adam* hp,
                                   void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    AdamHost h;
    int rc = check_hp(hp, &h, "dccf_adam_sweep");
    if (rc != DCCF_OK) return rc;
    DCCF_CHECK_ARG(table && m && v && head, "dccf_adam_sweep: null buffer");
    DCCF_CHECK_ARG(n_seg >= 0 && seg_len >= 0, "dccf_adam_sweep: negative record layout");
    const int64_t n_rec = (int64_t)n_seg * seg_len;
    DCCF_CHECK_ARG(n_rec == 0 || (rec_keys && rec_grads && next), "dccf_adam_sweep: %lld records but a null record buffer", (long long)n_rec);
    DCCF_CHECK_ARG(n_rec < ((int64_t)1 << 31) && n_table_rows < ((int64_t)1 << 31), "dccf_adam_sweep: sizes exceed int32 indexing");
    DCCF_CHECK_ARG(n_seg <= 1 || (key_seg_stride >= seg_len && grad_seg_stride >= seg_len * D), "dccf_adam_sweep: segment strides overlap");
    if (n_table_rows <= 0) return DCCF_OK;
    RecLayout L;
    L.seg_len = seg_len > 0 ? seg_len : 1;
    L.key_seg_stride = key_seg_stride;
    L.grad_seg_stride = grad_seg_stride;
    if (n_rec > 0) {
        k_link_records<<<(unsigned)((n_rec + 255) / 256), 256, 0, stream>>>(rec_keys, n_rec, n_table_rows, L, head, next);
        DCCF_CHECK_LAUNCH("k_link_records");
    }
    // 16 rows per 256-thread CTA per pass; cap the grid at 8 CTAs per SM and grid-stride beyond
    int64_t ctas = (n_table_rows + 15) / 16;
    const int64_t cap = 148 * 8;
    if (ctas > cap) ctas = cap;
    k_adam_sweep<<<(unsigned)ctas, 256, 0, stream>>>(table, m, v, n_table_rows, rec_grads, L, head, next, h);
    DCCF_CHECK_LAUNCH("k_adam_sweep");
    return DCCF_OK;
}

extern "C" int dccf_adam_sweep(float* table, float* m, float* v, int64_t n_table_rows, const int32_t* rec_keys,
                               const float* rec_grads, int64_t n_rec, int32_t* head, int32_t* next,
                               const dccf_adam* hp, void* stream_) {
    DCCF_CHECK_ARG(n_rec >= 0, "dccf_adam_sweep: negative record count");
    return dccf_adam_sweep_seg(table, m, v, n_table_rows, rec_keys, rec_grads, 1, n_rec, n_rec, n_rec * D, head, next,
                               hp, stream_);
}

extern "C" int dccf_sum_parts(const float* parts, int32_t n_parts, int64_t part_stride, int64_t n, float* out,
                              void* stream_) {
    DCCF_CHECK_ARG(out && (n_parts == 0 || parts), "dccf_sum_parts: null buffer");
    DCCF_CHECK_ARG(n_parts >= 0 && part_stride >= 0, "dccf_sum_parts: bad layout");
    if (n <= 0) return DCCF_OK;
    int64_t ctas = (n + 255) / 256;
    if (ctas > 148 * 8) ctas = 148 * 8;
    k_sum_parts<<<(unsigned)ctas, 256, 0, (cudaStream_t)stream_>>>(parts, n_parts, part_stride, n, out);
    DCCF_CHECK_LAUNCH("k_sum_parts");
    return DCCF_OK;
}

extern "C" int dccf_adam_dense(float* p, float* m, float* v, int64_t n, const float* g_parts, int32_t n_parts,
                               int64_t part_stride, const dccf_adam* hp, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    AdamHost h;
    int rc = check_hp(hp, &h, "dccf_adam_dense");
    if (rc != DCCF_OK) return rc;
    DCCF_CHECK_ARG(p && m && v, "dccf_adam_dense: null buffer");
    DCCF_CHECK_ARG(n_parts == 0 || g_parts, "dccf_adam_dense: null gradient buffer");
    DCCF_CHECK_ARG(n_parts >= 0 && part_stride >= 0, "dccf_adam_dense: bad partial layout");
    if (n <= 0) return DCCF_OK;
    int64_t ctas = (n + 255) / 256;
    if (ctas > 148 * 8) ctas = 148 * 8;
    k_adam_dense<<<(unsigned)ctas, 256, 0, stream>>>(p, m, v, n, g_parts, n_parts, part_stride, h);
    DCCF_CHECK_LAUNCH("k_adam_dense");
    return DCCF_OK;
}

extern "C" int dccf_state_advance(int32_t* step_dev, uint64_t* offset_dev, uint64_t offset_inc, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (!step_dev && !offset_dev) return DCCF_OK;
    k_state_advance<<<1, 32, 0, stream>>>(step_dev, offset_dev, offset_inc);
    DCCF_CHECK_LAUNCH("k_state_advance");
    return DCCF_OK;
}
