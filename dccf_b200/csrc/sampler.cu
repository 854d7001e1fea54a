// sampler.cu — host-side (CPU) exact replay of the reference's negative sampler.
//
// Replaces the per-interaction Python loop of src/data_processor/DataProcessor.py:446-524
// (_sample_neg_from_uid_list) draw for draw: the caller hands in numpy's legacy MT19937 state
// (np.random.get_state()) and gets it back advanced exactly as the reference would have advanced it,
// so negatives are bit-identical under the same seed.  Primitives:
//   np.random.randint(n)            : mask = 2^k-1 >= n-1 ; repeat v = next_u32 & mask until v <= n-1
//   np.random.choice(pool, k, False): legacy permutation(len(pool))[:k] = Fisher-Yates from the top,
//                                     j = masked-rejection draw in [0, i], swap(i, j)
// This is host code (no kernel): the sampler consumes a sequential generator.
#include <algorithm>
#include <vector>
#include "common.cuh"

namespace dccf {

struct MT {
    uint32_t* key;
    int pos;
    inline void twist() {
        const uint32_t N = 624, M = 397;
        uint32_t y;
        uint32_t i;
        for (i = 0; i < N - M; i++) {
            y = (key[i] & 0x80000000u) | (key[i + 1] & 0x7fffffffu);
            key[i] = key[i + M] ^ (y >> 1) ^ (-(int32_t)(y & 1) & 0x9908b0dfu);
        }
        for (; i < N - 1; i++) {
            y = (key[i] & 0x80000000u) | (key[i + 1] & 0x7fffffffu);
            key[i] = key[i + (M - N)] ^ (y >> 1) ^ (-(int32_t)(y & 1) & 0x9908b0dfu);
        }
        y = (key[N - 1] & 0x80000000u) | (key[0] & 0x7fffffffu);
        key[N - 1] = key[M - 1] ^ (y >> 1) ^ (-(int32_t)(y & 1) & 0x9908b0dfu);
        pos = 0;
    }
    inline uint32_t next() {
        if (pos >= 624) twist();
        uint32_t y = key[pos++];
        y ^= (y >> 11);
        y ^= (y << 7) & 0x9d2c5680u;
        y ^= (y << 15) & 0xefc60000u;
        y ^= (y >> 18);
        return y;
    }
    // uniform integer in [0, max] by masked rejection (numpy legacy rk_interval / _bounded_uint64 for max < 2^32)
    inline uint32_t interval(uint32_t max) {
        if (max == 0) return 0;
        uint32_t mask = max;
        mask |= mask >> 1; mask |= mask >> 2; mask |= mask >> 4; mask |= mask >> 8; mask |= mask >> 16;
        uint32_t v;
        while ((v = (next() & mask)) > max) {
        }
        return v;
    }
};

static inline bool in_sorted(const int64_t* a, int64_t n, int64_t x) { return std::binary_search(a, a + n, x); }

}  // namespace dccf

using namespace dccf;

// hist_*: CSR over user ids [0, n_users): sorted, de-duplicated item lists.  uids [n]: users in sampling order.
// out_iid [n * neg_n].  Returns 0, or -1 with dccf_last_error() set (e.g. the reference's
// `assert remain_iids_num >= neg_n` would have fired).
extern "C" int dccf_sample_negatives(uint32_t* mt_key, int32_t* mt_pos, const int64_t* uids, int64_t n, int32_t neg_n,
                                     int32_t train, int64_t item_num, int64_t n_users, const int64_t* train_off,
                                     const int64_t* train_items, const int64_t* vt_off, const int64_t* vt_items,
                                     int64_t* out_iid) {
    DCCF_CHECK_ARG(mt_key && mt_pos && out_iid && train_off && train_items, "dccf_sample_negatives: null argument");
    DCCF_CHECK_ARG(n == 0 || uids, "dccf_sample_negatives: null uid list");
    DCCF_CHECK_ARG(train || (vt_off && vt_items), "dccf_sample_negatives: evaluation sampling needs the validation/test history");
    DCCF_CHECK_ARG(item_num > 0 && item_num <= 0xffffffffLL && neg_n >= 0, "dccf_sample_negatives: bad item_num / neg_n");
    MT mt{mt_key, *mt_pos};
    // negatives already handed to a user during this call (kept across rows when train, DP:477-504,516-517)
    std::vector<std::vector<int64_t>> drawn;
    if (train) drawn.resize((size_t)n_users);
    std::vector<int64_t> local, taken_vt, pool;
    for (int64_t r = 0; r < n; ++r) {
        const int64_t u = uids[r];
        DCCF_CHECK_ARG(u >= 0 && u < n_users, "dccf_sample_negatives: uid %lld outside [0,%lld)", (long long)u, (long long)n_users);
        const int64_t* th = train_items + train_off[u];
        const int64_t tn = train_off[u + 1] - train_off[u];
        const int64_t* vh = nullptr;
        int64_t vn = 0;
        int64_t n_taken = tn;
        if (!train) {
            vh = vt_items + vt_off[u];
            vn = vt_off[u + 1] - vt_off[u];
            for (int64_t k = 0; k < vn; ++k) n_taken += in_sorted(th, tn, vh[k]) ? 0 : 1;   // |train U vt|
        }
        std::vector<int64_t>& mine = train ? drawn[(size_t)u] : local;
        if (!train) mine.clear();
        n_taken += (int64_t)mine.size();          // drawn items are never in the history: disjoint
        const int64_t remain = item_num - n_taken;
        auto is_taken = [&](int64_t i) {
            return in_sorted(th, tn, i) || (vn > 0 && in_sorted(vh, vn, i)) ||
                   std::find(mine.begin(), mine.end(), i) != mine.end();
        };
        const bool use_pool = (1.0 * (double)remain / (double)item_num) < 0.2;
        if (use_pool) {
            pool.clear();
            for (int64_t i = 1; i < item_num; ++i)          // item 0 is excluded on this branch only (DP:493)
                if (!is_taken(i)) pool.push_back(i);
        }
        if (remain < neg_n) {
            set_error("dccf_sample_negatives: user %lld has %lld items left but %d negatives were requested "
                      "(the reference asserts here, DataProcessor.py:495)", (long long)u, (long long)remain, neg_n);
            *mt_pos = mt.pos;
            return DCCF_ERR_ARG;
        }
        int64_t* out = out_iid + r * neg_n;
        if (!use_pool) {
            for (int k = 0; k < neg_n; ++k) {
                int64_t i = mt.interval((uint32_t)(item_num - 1));
                while (is_taken(i)) i = mt.interval((uint32_t)(item_num - 1));
                out[k] = i;
                mine.push_back(i);
            }
        } else {
            // np.random.choice(pool, neg_n, replace=False) == pool[permutation(len(pool))[:neg_n]]
            const int64_t m = (int64_t)pool.size();
            if (m < neg_n) {
                set_error("dccf_sample_negatives: user %lld: pool of %lld items < %d negatives (np.random.choice would raise)",
                          (long long)u, (long long)m, neg_n);
                *mt_pos = mt.pos;
                return DCCF_ERR_ARG;
            }
            std::vector<int64_t> perm((size_t)m);
            for (int64_t i = 0; i < m; ++i) perm[(size_t)i] = i;
            for (int64_t i = m - 1; i >= 1; --i) {
                const int64_t j = mt.interval((uint32_t)i);
                std::swap(perm[(size_t)i], perm[(size_t)j]);
            }
            for (int k = 0; k < neg_n; ++k) {
                out[k] = pool[(size_t)perm[(size_t)k]];
                mine.push_back(out[k]);
            }
        }
    }
    *mt_pos = mt.pos;
    return DCCF_OK;
}
