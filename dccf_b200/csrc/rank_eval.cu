// rank_eval.cu — kernel (d) of the DCCF hot path: per-user top-k over the candidate list (positives +
// test_neg_n sampled negatives) and the ranking metrics at k.
//
// Replaces the ranking branch of BaseModel.evaluate_method (src/models/BaseModel.py:82-126:
// global pandas sort by score, groupby uid, Python loop per user) and the helpers it calls,
// src/utils/rank_metrics.py:61-87 (precision_at_k) and :130-201 (dcg_at_k / ndcg_at_k, method=1).
//
// One warp per user.  The k best candidates are extracted by k rounds of "best key strictly below
// the previous winner": each lane scans its strided share of the candidates (they stay in L1),
// then a shuffle tournament picks the warp-wide winner.  Order: score descending, then item id
// ascending, then row ascending — a total order, so the result is deterministic (the reference's
// quicksort leaves ties unordered).  NaN scores rank last, as pandas does.
#include "common.cuh"

namespace dccf {

struct RankKey {
    float s;
    int64_t iid;
    int32_t row;
    float label;
};

// true when a ranks strictly before b
__device__ __forceinline__ bool key_before(const RankKey& a, const RankKey& b) {
    if (a.s != b.s) return a.s > b.s;
    if (a.iid != b.iid) return a.iid < b.iid;
    return a.row < b.row;
}

__device__ __forceinline__ RankKey shfl_key(const RankKey& k, int src_xor) {
    RankKey o;
    o.s = __shfl_xor_sync(0xffffffffu, k.s, src_xor);
    o.iid = __shfl_xor_sync(0xffffffffu, k.iid, src_xor);
    o.row = __shfl_xor_sync(0xffffffffu, k.row, src_xor);
    o.label = __shfl_xor_sync(0xffffffffu, k.label, src_xor);
    return o;
}

__device__ __forceinline__ float order_score(float s) { return (s != s) ? -INFINITY : s; }

constexpr int RANK_WARPS = 4;

__global__ void __launch_bounds__(RANK_WARPS * 32) k_rank_eval(
    const float* __restrict__ scores, const float* __restrict__ labels, const int64_t* __restrict__ iids,
    const int32_t* __restrict__ cand_rows, const int64_t* __restrict__ user_off, int64_t n_users, int32_t k,
    int64_t* __restrict__ out_topk_iid, int32_t* __restrict__ out_topk_row, double* __restrict__ out_metrics) {
    const int lane = threadIdx.x & 31;
    const int64_t g = (int64_t)blockIdx.x * RANK_WARPS + (threadIdx.x >> 5);
    if (g >= n_users) return;  // warp-uniform
    const int64_t lo = user_off[g], hi = user_off[g + 1];
    const int64_t n = hi - lo;

    // ---- total relevance and (for the ideal DCG) the k largest labels --------------------------
    double label_sum = 0.0;
    for (int64_t c = lo + lane; c < hi; c += 32) label_sum += (double)__ldg(labels + cand_rows[c]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) label_sum += __shfl_xor_sync(0xffffffffu, label_sum, o);

    const int kk = (int)min((int64_t)k, n);
    double idcg = 0.0;
    {
        // k rounds over (label desc, position asc)
        float last_l = INFINITY;
        int64_t last_c = -1;
        for (int t = 0; t < kk; ++t) {
            float best_l = -INFINITY;
            int64_t best_c = -1;
            for (int64_t c = lo + lane; c < hi; c += 32) {
                const float l = __ldg(labels + cand_rows[c]);
                const bool below = (t == 0) || (l < last_l) || (l == last_l && c > last_c);
                if (below && (best_c < 0 || l > best_l || (l == best_l && c < best_c))) {
                    best_l = l;
                    best_c = c;
                }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const float ol = __shfl_xor_sync(0xffffffffu, best_l, o);
                const int64_t oc = __shfl_xor_sync(0xffffffffu, best_c, o);
                if (oc >= 0 && (best_c < 0 || ol > best_l || (ol == best_l && oc < best_c))) {
                    best_l = ol;
                    best_c = oc;
                }
            }
            idcg += (double)best_l / log2((double)(t + 2));
            last_l = best_l;
            last_c = best_c;
        }
    }

    // ---- k rounds of selection over (score desc, iid asc, row asc) -----------------------------
    double dcg = 0.0, hit_sum = 0.0;
    int nonzero = 0;
    RankKey last;
    last.s = INFINITY; last.iid = -1; last.row = -1; last.label = 0.f;
    for (int t = 0; t < k; ++t) {
        RankKey best;
        best.s = -INFINITY; best.iid = INT64_MAX; best.row = INT32_MAX; best.label = 0.f;
        bool have = false;
        if (t < kk) {
            for (int64_t c = lo + lane; c < hi; c += 32) {
                RankKey cur;
                cur.row = cand_rows[c];
                cur.s = order_score(__ldg(scores + cur.row));
                cur.iid = __ldg(iids + cur.row);
                const bool below = (t == 0) || key_before(last, cur);
                if (below && (!have || key_before(cur, best))) {
                    best = cur;
                    have = true;
                }
            }
            // tournament; lanes without a candidate carry the sentinel (row == INT32_MAX)
            if (!have) { best.s = -INFINITY; best.iid = INT64_MAX; best.row = INT32_MAX; }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const RankKey other = shfl_key(best, o);
                if (other.row != INT32_MAX && (best.row == INT32_MAX || key_before(other, best))) best = other;
            }
            best.label = __ldg(labels + best.row);
            dcg += (double)best.label / log2((double)(t + 2));
            hit_sum += (double)best.label;
            nonzero += (best.label != 0.f) ? 1 : 0;
            last = best;
        }
        if (lane == 0) {
            if (out_topk_iid) out_topk_iid[g * k + t] = (t < kk) ? best.iid : -1;
            if (out_topk_row) out_topk_row[g * k + t] = (t < kk) ? best.row : -1;
        }
    }

    if (lane == 0) {
        double* o = out_metrics + g * 5;
        o[0] = (idcg != 0.0) ? dcg / idcg : 0.0;                         // ndcg@k   (rank_metrics.py:198-201)
        o[1] = (hit_sum > 0.0) ? 1.0 : 0.0;                              // hit@k    (BaseModel.py:98-102)
        o[2] = (double)nonzero / (double)k;                              // precision@k (rank_metrics.py:83-87)
        o[3] = hit_sum / label_sum;                                      // recall@k (BaseModel.py:108-112)
        o[4] = 2.0 * hit_sum / ((double)k + label_sum);                  // f1@k     (BaseModel.py:121-126)
    }
}

}  // namespace dccf

using namespace dccf;

extern "C" int dccf_rank_eval(const float* scores, const float* labels, const int64_t* iids, const int32_t* cand_rows,
                              const int64_t* user_off, int64_t n_users, int32_t k, int64_t* out_topk_iid,
                              int32_t* out_topk_row, double* out_metrics, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    DCCF_CHECK_ARG(scores && labels && iids && cand_rows && user_off && out_metrics, "dccf_rank_eval: null buffer");
    DCCF_CHECK_ARG(k >= 1 && k <= 1024, "dccf_rank_eval: k=%d outside [1,1024]", k);
    if (n_users <= 0) return DCCF_OK;
    k_rank_eval<<<(unsigned)((n_users + RANK_WARPS - 1) / RANK_WARPS), RANK_WARPS * 32, 0, stream>>>(
        scores, labels, iids, cand_rows, user_off, n_users, k, out_topk_iid, out_topk_row, out_metrics);
    DCCF_CHECK_LAUNCH("k_rank_eval");
    return DCCF_OK;
}
