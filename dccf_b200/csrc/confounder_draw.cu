// confounder_draw.cu — host-side (CPU) exact replay of the confounder draw of src/models/DCCF.py:72,
//     sample_item = torch.randint(item_num, size=(P, S))            (torch CPU generator)
// i.e. element k (row-major) = next_u32 % item_num with next_u32 from at::mt19937, strictly sequential, the stream
// continuing across calls (verified against torch 2.11).  An evaluation pass needs 10 draws per
// scored row — 4.8e8 per 1000-negative test set at the electronics shape — and torch's element-at-a-time loop
// (≈3-4 ns per draw on one core) then out-lasts the GPU scorer (0.43 ms per 163 840-draw batch).  Here a whole
// generation of 624 words is twisted and tempered in vectorisable loops and the remainder is taken with a
// multiply-high instead of a division.  The caller hands in the generator's words and position
// (torch.get_rng_state()) and writes them back, so every other consumer of the torch generator continues from
// exactly where the reference would have left it.  This is host code (no kernel): MT19937 is a sequential generator.
#include "common.cuh"

namespace dccf {
namespace {

constexpr int MT_N = 624, MT_M = 397;

// one generation: the three ranges have no dependence shorter than MT_N - MT_M = 227 words, so each loop vectorises
inline void mt_twist(uint32_t* __restrict__ s) {
    auto mix = [](uint32_t a, uint32_t b, uint32_t far_word) {
        const uint32_t y = (a & 0x80000000u) | (b & 0x7fffffffu);
        return far_word ^ (y >> 1) ^ ((0u - (y & 1u)) & 0x9908b0dfu);
    };
    for (int i = 0; i < MT_N - MT_M; ++i) s[i] = mix(s[i], s[i + 1], s[i + MT_M]);
    for (int i = MT_N - MT_M; i < MT_N - 1; ++i) s[i] = mix(s[i], s[i + 1], s[i + MT_M - MT_N]);
    s[MT_N - 1] = mix(s[MT_N - 1], s[0], s[MT_M - 1]);
}

inline uint32_t mt_temper(uint32_t y) {
    y ^= (y >> 11);
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= (y >> 18);
    return y;
}

}  // namespace
}  // namespace dccf

using namespace dccf;

// mt_state [624]: at::mt19937's words; *mt_left: its `left_` counter (words still readable + 1; 1 = regenerate on the
// next draw, which is also the freshly seeded state); both advanced in place.  *mt_next receives the matching `next_`
// index.  out [n] = successive (next_u32 % high), 0 < high < 2^32 (torch switches to 64-bit draws above that).
extern "C" int dccf_confounder_draw(uint32_t* mt_state, int32_t* mt_left, int32_t* mt_next, int64_t high, int64_t n,
                                    int64_t* out) {
    DCCF_CHECK_ARG(mt_state && mt_left && mt_next, "dccf_confounder_draw: null generator state");
    DCCF_CHECK_ARG(n == 0 || out, "dccf_confounder_draw: null output");
    DCCF_CHECK_ARG(n >= 0 && high > 0 && high < (1LL << 32), "dccf_confounder_draw: need n >= 0 and 0 < high < 2^32");
    DCCF_CHECK_ARG(*mt_left >= 1 && *mt_left <= MT_N, "dccf_confounder_draw: corrupt generator state (left = %d)", *mt_left);
    int pos = MT_N + 1 - *mt_left;                      // index of the next unread word, MT_N = none left
    // v % high as a multiply-high (Lemire, Kaser, Kurz 2019): exact for every 32-bit v and high
    const uint64_t d = (uint64_t)high;
    const uint64_t magic = ~0ull / d + 1ull;
    uint32_t y[MT_N];
    while (n > 0) {
        if (pos == MT_N) {
            mt_twist(mt_state);
            pos = 0;
        }
        const int take = (int)((int64_t)(MT_N - pos) < n ? (MT_N - pos) : n);
        const uint32_t* w = mt_state + pos;
        for (int i = 0; i < take; ++i) y[i] = mt_temper(w[i]);
        for (int i = 0; i < take; ++i) out[i] = (int64_t)(uint64_t)(((unsigned __int128)(magic * y[i]) * d) >> 64);
        out += take;
        n -= take;
        pos += take;
    }
    *mt_left = MT_N + 1 - pos;
    *mt_next = pos;
    return DCCF_OK;
}
