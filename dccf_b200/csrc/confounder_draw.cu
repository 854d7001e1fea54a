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
#include "mt19937.cuh"

namespace dccf {

constexpr int MT_N = mt::N, MT_M = mt::M;

// one generation on the host: the three ranges have no dependence shorter than 227 words, so each loop vectorises
static inline void mt_twist(uint32_t* __restrict__ s) {
    for (int i = 0; i < MT_N - MT_M; ++i) s[i] = mt::mix(s[i], s[i + 1], s[i + MT_M]);
    for (int i = MT_N - MT_M; i < MT_N - 1; ++i) s[i] = mt::mix(s[i], s[i + 1], s[i + MT_M - MT_N]);
    s[MT_N - 1] = mt::mix(s[MT_N - 1], s[0], s[MT_M - 1]);
}

// ---------------------------------------------------------------------------------------------------------------
// The same stream continued ON THE DEVICE.  MT19937 is sequential from one generation of 624 words to the next, but
// inside a generation the recurrence only reaches back 227 words (mt19937.cuh), so one CTA regenerates with three
// barriers (old and new generation in two shared-memory buffers), tempers, takes the remainder and stores 624 ids per
// round, coalesced.  The words and the read position live in device memory between calls: an evaluation pass uploads
// torch's generator once, every batch's draw is one launch that lands directly in the scorer's `sample_item` buffer
// (no 1.3 MB host->device copy per 16 384-pair batch, no host generator in the loop), and the final state goes back
// into torch's generator at the end of the pass.  The thread schedule below is emulated on the host, phase by phase,
// by tests/mt_emulate.cpp against the sequential generator.
// state [625]: words[624] + index of the next unread word (624 = regenerate first).
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(mt::CTA) k_confounder_draw(uint32_t* state, uint64_t high, uint64_t magic, int64_t n,
                                                             int64_t* __restrict__ out) {
    __shared__ uint32_t buf[2][MT_N];
    const int t = threadIdx.x;
    for (int i = t; i < MT_N; i += mt::CTA) buf[0][i] = state[i];
    int pos = (int)state[MT_N];
    __syncthreads();
    int cur = 0;
    int64_t done = 0;
    while (true) {                             // every quantity that steers the loop is uniform over the CTA
        const int take = (int)min((int64_t)(MT_N - pos), n - done);
        mt::emit(buf[cur], pos, take, t, magic, high, out + done);
        done += take;
        pos += take;
        if (done >= n) break;
        for (int phase = 0; phase < 3; ++phase) {
            mt::regen_phase(phase, buf[cur], buf[cur ^ 1], t);
            __syncthreads();
        }
        cur ^= 1;
        pos = 0;
    }
    for (int i = t; i < MT_N; i += mt::CTA) state[i] = buf[cur][i];
    if (t == 0) state[MT_N] = (uint32_t)pos;
}

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
    const uint64_t magic = mt::fastmod_magic(d);
    uint32_t y[MT_N];
    while (n > 0) {
        if (pos == MT_N) {
            mt_twist(mt_state);
            pos = 0;
        }
        const int take = (int)((int64_t)(MT_N - pos) < n ? (MT_N - pos) : n);
        const uint32_t* w = mt_state + pos;
        for (int i = 0; i < take; ++i) y[i] = mt::temper(w[i]);
        for (int i = 0; i < take; ++i) out[i] = mt::fastmod(y[i], magic, d);
        out += take;
        n -= take;
        pos += take;
    }
    *mt_left = MT_N + 1 - pos;
    *mt_next = pos;
    return DCCF_OK;
}

// Device twin of dccf_confounder_draw.  state_dev: uint32[625] in device memory (624 words + next unread index,
// 624 = regenerate first), advanced in place; out_dev [n] int64 in device memory.  Calls on one stream continue the
// stream in order.  0 < high < 2^32.
extern "C" int dccf_confounder_draw_dev(uint32_t* state_dev, int64_t high, int64_t n, int64_t* out_dev, void* stream_) {
    DCCF_CHECK_ARG(state_dev, "dccf_confounder_draw_dev: null generator state");
    DCCF_CHECK_ARG(n == 0 || out_dev, "dccf_confounder_draw_dev: null output");
    DCCF_CHECK_ARG(n >= 0 && high > 0 && high < (1LL << 32), "dccf_confounder_draw_dev: need n >= 0 and 0 < high < 2^32");
    if (n == 0) return DCCF_OK;
    const uint64_t d = (uint64_t)high;
    k_confounder_draw<<<1, mt::CTA, 0, (cudaStream_t)stream_>>>(state_dev, d, mt::fastmod_magic(d), n, out_dev);
    DCCF_CHECK_LAUNCH("k_confounder_draw");
    return DCCF_OK;
}
