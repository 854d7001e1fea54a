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
    std::vector<int64_t> local, pool;
    // "is item i taken for the current row" in O(1): stamp[i] == the row's stamp.  Each row marks its user's
    // history and the negatives the user already holds (a few dozen stores into an L2-resident array) instead of
    // searching them on every draw — with 1000 negatives per evaluation user a search of `mine` per draw is quadratic.
    std::vector<uint32_t> stamp((size_t)item_num, 0u);
    uint32_t cur = 0;
    for (int64_t r = 0; r < n; ++r) {
        const int64_t u = uids[r];
        DCCF_CHECK_ARG(u >= 0 && u < n_users, "dccf_sample_negatives: uid %lld outside [0,%lld)", (long long)u, (long long)n_users);
        if (++cur == 0) {                         // stamp wrapped after 2^32 - 1 rows: start over
            std::fill(stamp.begin(), stamp.end(), 0u);
            cur = 1;
        }
        int64_t n_taken = 0;                      // |train U vt U drawn| (ids outside [0, item_num) can never be drawn
        auto mark = [&](const int64_t* a, int64_t len) {          //  but count, as in the reference's len(set))
            for (int64_t k = 0; k < len; ++k) {
                const int64_t i = a[k];
                if (i < 0 || i >= item_num) {
                    n_taken += 1;
                } else if (stamp[(size_t)i] != cur) {
                    stamp[(size_t)i] = cur;
                    n_taken += 1;
                }
            }
        };
        mark(train_items + train_off[u], train_off[u + 1] - train_off[u]);
        if (!train) mark(vt_items + vt_off[u], vt_off[u + 1] - vt_off[u]);
        std::vector<int64_t>& mine = train ? drawn[(size_t)u] : local;
        if (!train) mine.clear();
        mark(mine.data(), (int64_t)mine.size());
        const int64_t remain = item_num - n_taken;
        auto is_taken = [&](int64_t i) { return stamp[(size_t)i] == cur; };
        auto take = [&](int64_t i) {
            stamp[(size_t)i] = cur;
            mine.push_back(i);
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
                take(i);
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
                take(out[k]);
            }
        }
    }
    *mt_pos = mt.pos;
    return DCCF_OK;
}
