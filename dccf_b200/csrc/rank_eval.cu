// rank_eval.cu — kernel (d) of the DCCF hot path: per-user top-k over the candidate list (positives +
// test_neg_n sampled negatives) and the ranking metrics at every requested k, in ONE launch.
//
// Replaces the ranking branch of BaseModel.evaluate_method (src/models/BaseModel.py:82-126:
// global pandas sort by score, groupby uid, one Python loop per user PER METRIC) and the helpers it
// calls, src/utils/rank_metrics.py:61-87 (precision_at_k) and :130-201 (dcg_at_k / ndcg_at_k, method=1).
//
// k_rank_stream (k <= 16): one warp per user, ONE streaming pass over the user's candidates.  Lane l reads
// candidates l, l+32, ... (coalesced; sixteen loads in flight) and keeps its own best KCAP entries (score, row)
// sorted in registers; a candidate is compared against the lane's current worst entry first.  Item ids are read
// ONLY to break exact score ties (a dependent load on the insertion path made the warp wait for the L2 once per
// candidate batch: 74 us instead of < 10 for 1 024 users).  The 32 sorted lists are merged by k shuffle
// tournaments over the list heads (the winning lane pops).  Labels are read once: total relevance, number of positives (ideal DCG
// of 0/1 labels is a count; other label values take a second pass that keeps the k largest labels), and the
// labels of the k winners.  Every metric (ndcg, hit, precision, recall, f1) at every requested k (up to 4
// values of k per launch) comes out of that one pass; per-user values are optional, the per-metric SUMS over
// users are accumulated in a fixed order (warp -> CTA -> last CTA), so a metric list such as
// "ndcg@5,recall@5,precision@5" costs one launch and one 120-byte read-back.
// Order: score descending, then item id ascending, then row ascending — a total order, so the result is
// deterministic (the reference's quicksort leaves ties unordered).  NaN scores rank last, as pandas does.
// k_rank_select (any k <= 1024): the k-rounds selection kernel, kept for k > 16.
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
constexpr int RANK_MAX_NK = 4;       // values of k per launch
constexpr int RANK_NCOL = 5;         // ndcg, hit, precision, recall, f1
constexpr int RANK_STREAM_MAX_K = 16;

struct RankKs {
    int32_t k[RANK_MAX_NK];
    int32_t n_k, kmax;
};

// entry of a lane's sorted list: score and row only.  The item id decides exact score ties and is fetched then
// (the rows are in L1 / L2: they were just streamed); the label is loaded for the k winners only.
struct Ent {
    float s;
    int32_t row;      // INT32_MAX = sentinel, ranks after every real candidate
};
__device__ __forceinline__ bool ent_before(const Ent& a, const Ent& b, const int64_t* __restrict__ iids) {
    if (b.row == INT32_MAX) return a.row != INT32_MAX;
    if (a.row == INT32_MAX) return false;
    if (a.s != b.s) return a.s > b.s;
    const int64_t ia = __ldg(iids + a.row), ib = __ldg(iids + b.row);
    if (ia != ib) return ia < ib;
    return a.row < b.row;
}
__device__ __forceinline__ Ent ent_sentinel() {
    Ent e;
    e.s = -INFINITY; e.row = INT32_MAX;
    return e;
}

template <int KCAP>
__device__ __forceinline__ void list_insert(Ent (&L)[KCAP], const Ent& cur, const int64_t* __restrict__ iids) {
    // precondition: cur ranks before L[KCAP-1]
    bool placed = false;
#pragma unroll
    for (int i = KCAP - 1; i >= 1; --i) {
        if (!placed) {
            if (ent_before(cur, L[i - 1], iids)) L[i] = L[i - 1];
            else { L[i] = cur; placed = true; }
        }
    }
    if (!placed) L[0] = cur;
}

template <int KCAP>
__device__ __forceinline__ void label_insert(float (&L)[KCAP], float v) {
    bool placed = false;
#pragma unroll
    for (int i = KCAP - 1; i >= 1; --i) {
        if (!placed) {
            if (v > L[i - 1]) L[i] = L[i - 1];
            else { L[i] = v; placed = true; }
        }
    }
    if (!placed) L[0] = v;
}

constexpr int RANK_CHUNK = 1024;     // candidates of a user staged in shared memory at a time (4 KB per warp)

template <int KCAP>
__global__ void __launch_bounds__(RANK_WARPS * 32) k_rank_stream(
    const float* __restrict__ scores, const float* __restrict__ labels, const int64_t* __restrict__ iids,
    const int32_t* __restrict__ cand_rows, const int64_t* __restrict__ user_off, int64_t n_users, const RankKs ks,
    int64_t* __restrict__ out_topk_iid, int32_t* __restrict__ out_topk_row, double* __restrict__ out_metrics,
    double* __restrict__ part_sums, int32_t* __restrict__ cta_counter, double* __restrict__ out_sums) {
    __shared__ double acc_s[RANK_WARPS][RANK_MAX_NK * RANK_NCOL];
    __shared__ float sc_s[RANK_WARPS][RANK_CHUNK];
    __shared__ int s_last;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int ncol = ks.n_k * RANK_NCOL;
    if (lane < RANK_MAX_NK * RANK_NCOL) acc_s[warp][lane] = 0.0;
    __syncwarp();
    const int kmax = ks.kmax;
    const int64_t n_warps = (int64_t)gridDim.x * RANK_WARPS;
    float* my_s = sc_s[warp];

    for (int64_t g = (int64_t)blockIdx.x * RANK_WARPS + warp; g < n_users; g += n_warps) {   // warp-uniform
        const int64_t lo = user_off[g], hi = user_off[g + 1];
        const int64_t n = hi - lo;
        Ent L[KCAP];
#pragma unroll
        for (int i = 0; i < KCAP; ++i) L[i] = ent_sentinel();
        double label_sum = 0.0;
        int n_pos = 0;
        bool nonbinary = false;

        for (int64_t base = lo; base < hi; base += RANK_CHUNK) {
            const int m = (int)min((int64_t)RANK_CHUNK, hi - base);
            // ---- phase 1, the streaming pass: scores -> shared memory, label statistics; eight candidates per lane in
            // flight.  (Deliberately SMALL code: a first version that kept 16 candidates in registers and inlined the
            // insertion 16 times was 13 k instructions long and spent 63 % of its stall samples waiting for the
            // instruction cache.)
            for (int c0 = lane; c0 < m; c0 += 32 * 8) {
                int32_t row[8];
                float s[8], l[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int c = c0 + 32 * j;
                    row[j] = (c < m) ? (cand_rows != nullptr ? __ldg(cand_rows + base + c) : (int32_t)(base + c)) : -1;
                }
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    if (row[j] >= 0) {
                        s[j] = __ldg(scores + row[j]);
                        l[j] = __ldg(labels + row[j]);
                    }
                }
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    if (row[j] >= 0) {
                        my_s[c0 + 32 * j] = order_score(s[j]);
                        label_sum += (double)l[j];
                        n_pos += (l[j] == 1.f) ? 1 : 0;
                        nonbinary |= (l[j] != 0.f && l[j] != 1.f);
                    }
                }
            }
            __syncwarp();
            // ---- phase 2, selection from shared memory: one copy of the insertion code ----
#pragma unroll 1
            for (int c = lane; c < m; c += 32) {
                Ent cur;
                cur.s = my_s[c];
                if (cur.s >= L[KCAP - 1].s) {
                    cur.row = cand_rows != nullptr ? __ldg(cand_rows + base + c) : (int32_t)(base + c);
                    if (ent_before(cur, L[KCAP - 1], iids)) list_insert<KCAP>(L, cur, iids);
                }
            }
            __syncwarp();
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            label_sum += __shfl_xor_sync(0xffffffffu, label_sum, o);
            n_pos += __shfl_xor_sync(0xffffffffu, n_pos, o);
        }
        nonbinary = __any_sync(0xffffffffu, nonbinary);

        // ---- merge: kmax tournaments over the heads of the 32 sorted lists; lane t keeps winner t ------
        const int kk = (int)min((int64_t)kmax, n);
        Ent mine = ent_sentinel();
#pragma unroll 1
        for (int t = 0; t < kk; ++t) {
            Ent best = L[0];
#pragma unroll 1
            for (int o = 16; o > 0; o >>= 1) {
                Ent other;
                other.s = __shfl_xor_sync(0xffffffffu, best.s, o);
                other.row = __shfl_xor_sync(0xffffffffu, best.row, o);
                if (ent_before(other, best, iids)) best = other;
            }
            if (L[0].row == best.row) {                      // rows are unique: exactly one lane pops
#pragma unroll
                for (int i = 0; i + 1 < KCAP; ++i) L[i] = L[i + 1];
                L[KCAP - 1] = ent_sentinel();
            }
            if (lane == t) mine = best;
        }
        const float my_label = (lane < kk) ? __ldg(labels + mine.row) : 0.f;
        if (out_topk_iid != nullptr && lane < kmax) out_topk_iid[g * kmax + lane] = (lane < kk) ? __ldg(iids + mine.row) : -1;
        if (out_topk_row != nullptr && lane < kmax) out_topk_row[g * kmax + lane] = (lane < kk) ? mine.row : -1;

        // ---- ideal ordering of the labels: 0/1 labels -> the first n_pos positions are 1; otherwise the kmax
        // largest labels by a second pass (values only) --------------------------------------------------------
        float ideal = (lane < min(kk, n_pos)) ? 1.f : 0.f;    // lane t: label at ideal position t
        if (nonbinary) {                                       // warp-uniform
            float T[KCAP];
#pragma unroll
            for (int i = 0; i < KCAP; ++i) T[i] = -INFINITY;
#pragma unroll 1
            for (int64_t c = lo + lane; c < hi; c += 32) {
                const float v = __ldg(labels + (cand_rows != nullptr ? __ldg(cand_rows + c) : (int32_t)c));
                if (v > T[KCAP - 1]) label_insert<KCAP>(T, v);
            }
            ideal = 0.f;
#pragma unroll 1
            for (int t = 0; t < kk; ++t) {
                float best = T[0];
                int who = lane;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    const float ob = __shfl_xor_sync(0xffffffffu, best, o);
                    const int ow = __shfl_xor_sync(0xffffffffu, who, o);
                    if (ob > best || (ob == best && ow < who)) { best = ob; who = ow; }
                }
                if (lane == who) {
#pragma unroll
                    for (int i = 0; i + 1 < KCAP; ++i) T[i] = T[i + 1];
                    T[KCAP - 1] = -INFINITY;
                }
                if (lane == t) ideal = best;
            }
        }

        // ---- metrics at every requested k: sequential float64 sums over positions 0 .. k-1 ---------------
        double dcg = 0.0, idcg = 0.0, hit_sum = 0.0;
        int nonzero = 0, next_k = 0;
#pragma unroll 1
        for (int t = 0; t < kmax; ++t) {
            const float lab = __shfl_sync(0xffffffffu, my_label, t);
            const float idl = __shfl_sync(0xffffffffu, ideal, t);
            if (t < kk) {
                const double disc = log2((double)(t + 2));
                dcg += (double)lab / disc;
                idcg += (double)idl / disc;
                hit_sum += (double)lab;
                nonzero += (lab != 0.f) ? 1 : 0;
            }
            while (next_k < ks.n_k && ks.k[next_k] == t + 1) {     // ks ascending
                if (lane == 0) {
                    const double k = (double)(t + 1);
                    double m[RANK_NCOL];
                    m[0] = (idcg != 0.0) ? dcg / idcg : 0.0;                    // ndcg@k   (rank_metrics.py:198-201)
                    m[1] = (hit_sum > 0.0) ? 1.0 : 0.0;                         // hit@k    (BaseModel.py:98-102)
                    m[2] = (double)nonzero / k;                                 // precision@k (rank_metrics.py:83-87)
                    m[3] = hit_sum / label_sum;                                 // recall@k (BaseModel.py:108-112)
                    m[4] = 2.0 * hit_sum / (k + label_sum);                     // f1@k     (BaseModel.py:121-126)
#pragma unroll
                    for (int q = 0; q < RANK_NCOL; ++q) {
                        if (out_metrics != nullptr) out_metrics[(g * ks.n_k + next_k) * RANK_NCOL + q] = m[q];
                        acc_s[warp][next_k * RANK_NCOL + q] += m[q];
                    }
                }
                ++next_k;
            }
        }
        __syncwarp();
    }

    // ---- sums over users in a fixed order: warps of the CTA, then (last CTA) the CTAs in index order ------
    if (out_sums == nullptr) return;
    __syncthreads();
    if ((int)threadIdx.x < ncol) {
        double v = 0.0;
        for (int w = 0; w < RANK_WARPS; ++w) v += acc_s[w][threadIdx.x];
        part_sums[(size_t)blockIdx.x * (RANK_MAX_NK * RANK_NCOL) + threadIdx.x] = v;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        const int32_t prev = atomicAdd(cta_counter, 1);
        s_last = (prev == (int32_t)gridDim.x - 1) ? 1 : 0;
        if (s_last) {
            __threadfence();
            *cta_counter = 0;
        }
    }
    __syncthreads();
    if (!s_last) return;
    // warp w sums columns w, w + RANK_WARPS, ...: lane-strided over the CTAs (ascending), then a shuffle tree
    for (int col = warp; col < ncol; col += RANK_WARPS) {
        double v = 0.0;
        for (int b = lane; b < (int)gridDim.x; b += 32) v += __ldcg(part_sums + (size_t)b * (RANK_MAX_NK * RANK_NCOL) + col);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) out_sums[col] = v;
    }
}

// The k-rounds selection kernel (k > 16): each round finds the best key strictly below the previous winner.
__global__ void __launch_bounds__(RANK_WARPS * 32) k_rank_select(
    const float* __restrict__ scores, const float* __restrict__ labels, const int64_t* __restrict__ iids,
    const int32_t* __restrict__ cand_rows, const int64_t* __restrict__ user_off, int64_t n_users, int32_t k,
    int64_t* __restrict__ out_topk_iid, int32_t* __restrict__ out_topk_row, double* __restrict__ out_metrics) {
    const int lane = threadIdx.x & 31;
    const int64_t g = (int64_t)blockIdx.x * RANK_WARPS + (threadIdx.x >> 5);
    if (g >= n_users) return;  // warp-uniform
    const int64_t lo = user_off[g], hi = user_off[g + 1];
    const int64_t n = hi - lo;
    auto row_of = [&](int64_t c) { return cand_rows != nullptr ? cand_rows[c] : (int32_t)c; };

    double label_sum = 0.0;
    for (int64_t c = lo + lane; c < hi; c += 32) label_sum += (double)__ldg(labels + row_of(c));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) label_sum += __shfl_xor_sync(0xffffffffu, label_sum, o);

    const int kk = (int)min((int64_t)k, n);
    double idcg = 0.0;
    {
        float last_l = INFINITY;
        int64_t last_c = -1;
        for (int t = 0; t < kk; ++t) {
            float best_l = -INFINITY;
            int64_t best_c = -1;
            for (int64_t c = lo + lane; c < hi; c += 32) {
                const float l = __ldg(labels + row_of(c));
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
                cur.row = row_of(c);
                cur.s = order_score(__ldg(scores + cur.row));
                cur.iid = __ldg(iids + cur.row);
                const bool below = (t == 0) || key_before(last, cur);
                if (below && (!have || key_before(cur, best))) {
                    best = cur;
                    have = true;
                }
            }
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
        double* o = out_metrics + g * RANK_NCOL;
        o[0] = (idcg != 0.0) ? dcg / idcg : 0.0;
        o[1] = (hit_sum > 0.0) ? 1.0 : 0.0;
        o[2] = (double)nonzero / (double)k;
        o[3] = hit_sum / label_sum;
        o[4] = 2.0 * hit_sum / ((double)k + label_sum);
    }
}

}  // namespace dccf

using namespace dccf;

extern "C" int64_t dccf_rank_eval_ws_bytes(int64_t n_users) {
    (void)n_users;
    // partial sums of at most 148 * 4 CTAs + the CTA counter
    return (int64_t)(148 * 4) * RANK_MAX_NK * RANK_NCOL * (int64_t)sizeof(double) + 16;
}

extern "C" int dccf_rank_eval_multi(const float* scores, const float* labels, const int64_t* iids, const int32_t* cand_rows,
                                    const int64_t* user_off, int64_t n_users, const int32_t* ks_host, int32_t n_k,
                                    int64_t* out_topk_iid, int32_t* out_topk_row, double* out_metrics, void* ws,
                                    double* out_sums, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    DCCF_CHECK_ARG(scores && labels && iids && user_off, "dccf_rank_eval_multi: null buffer");
    DCCF_CHECK_ARG(ks_host && n_k >= 1 && n_k <= RANK_MAX_NK, "dccf_rank_eval_multi: 1 to %d values of k per call", RANK_MAX_NK);
    DCCF_CHECK_ARG(out_metrics || out_sums, "dccf_rank_eval_multi: no output requested");
    DCCF_CHECK_ARG(out_sums == nullptr || ws != nullptr, "dccf_rank_eval_multi: the sums need the workspace (dccf_rank_eval_ws_bytes)");
    RankKs ks;
    ks.n_k = n_k;
    for (int i = 0; i < RANK_MAX_NK; ++i) ks.k[i] = 0;
    for (int i = 0; i < n_k; ++i) {
        DCCF_CHECK_ARG(ks_host[i] >= 1 && ks_host[i] <= RANK_STREAM_MAX_K, "dccf_rank_eval_multi: k=%d outside [1,%d]", ks_host[i], RANK_STREAM_MAX_K);
        DCCF_CHECK_ARG(i == 0 || ks_host[i] > ks_host[i - 1], "dccf_rank_eval_multi: the values of k must be strictly ascending");
        ks.k[i] = ks_host[i];
    }
    ks.kmax = ks.k[n_k - 1];
    if (n_users <= 0) {
        if (out_sums) cudaMemsetAsync(out_sums, 0, sizeof(double) * n_k * RANK_NCOL, stream);
        return DCCF_OK;
    }
    int64_t ctas = (n_users + RANK_WARPS - 1) / RANK_WARPS;
    if (ctas > 148 * 4) ctas = 148 * 4;
    double* part = reinterpret_cast<double*>(ws);
    int32_t* counter = ws ? reinterpret_cast<int32_t*>(reinterpret_cast<uint8_t*>(ws) + (size_t)(148 * 4) * RANK_MAX_NK * RANK_NCOL * sizeof(double)) : nullptr;
#define DCCF_RANK_LAUNCH(KC)                                                                                              \
    k_rank_stream<KC><<<(unsigned)ctas, RANK_WARPS * 32, 0, stream>>>(scores, labels, iids, cand_rows, user_off, n_users, \
                                                                      ks, out_topk_iid, out_topk_row, out_metrics, part,  \
                                                                      counter, out_sums)
    if (ks.kmax <= 4) DCCF_RANK_LAUNCH(4);
    else if (ks.kmax <= 8) DCCF_RANK_LAUNCH(8);
    else DCCF_RANK_LAUNCH(16);
#undef DCCF_RANK_LAUNCH
    DCCF_CHECK_LAUNCH("k_rank_stream");
    return DCCF_OK;
}

extern "C" int dccf_rank_eval(const float* scores, const float* labels, const int64_t* iids, const int32_t* cand_rows,
                              const int64_t* user_off, int64_t n_users, int32_t k, int64_t* out_topk_iid,
                              int32_t* out_topk_row, double* out_metrics, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    DCCF_CHECK_ARG(scores && labels && iids && user_off && out_metrics, "dccf_rank_eval: null buffer");
    DCCF_CHECK_ARG(k >= 1 && k <= 1024, "dccf_rank_eval: k=%d outside [1,1024]", k);
    if (n_users <= 0) return DCCF_OK;
    if (k <= RANK_STREAM_MAX_K)
        return dccf_rank_eval_multi(scores, labels, iids, cand_rows, user_off, n_users, &k, 1, out_topk_iid, out_topk_row,
                                    out_metrics, nullptr, nullptr, stream_);
    k_rank_select<<<(unsigned)((n_users + RANK_WARPS - 1) / RANK_WARPS), RANK_WARPS * 32, 0, stream>>>(
        scores, labels, iids, cand_rows, user_off, n_users, k, out_topk_iid, out_topk_row, out_metrics);
    DCCF_CHECK_LAUNCH("k_rank_select");
    return DCCF_OK;
}
