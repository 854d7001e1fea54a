// Host emulation of k_confounder_draw's thread schedule (dccf_b200/csrc/confounder_draw.cu): the same per-phase functions
// (dccf_b200/csrc/mt19937.cuh) executed for t = 0 .. 255 between the kernel's barriers, in an adversarial thread order,
// compared with the textbook sequential MT19937 (genrand_int32 of Matsumoto & Nishimura, written out independently
// below) for random generator states, read positions, draw counts and ranges.  Built and run by
// tests/test_host_parity.py::test_device_confounder_schedule_emulated (g++, no GPU).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../dccf_b200/csrc/mt19937.cuh"

using namespace dccf;

// ---- independent sequential reference ---------------------------------------------------------------------------
struct RefMT {
    uint32_t mtv[624];
    int mti;
    uint32_t next() {
        static const uint32_t mag01[2] = {0x0u, 0x9908b0dfu};
        if (mti >= 624) {
            int kk;
            uint32_t y;
            for (kk = 0; kk < 624 - 397; kk++) {
                y = (mtv[kk] & 0x80000000u) | (mtv[kk + 1] & 0x7fffffffu);
                mtv[kk] = mtv[kk + 397] ^ (y >> 1) ^ mag01[y & 1u];
            }
            for (; kk < 623; kk++) {
                y = (mtv[kk] & 0x80000000u) | (mtv[kk + 1] & 0x7fffffffu);
                mtv[kk] = mtv[kk + (397 - 624)] ^ (y >> 1) ^ mag01[y & 1u];
            }
            y = (mtv[623] & 0x80000000u) | (mtv[0] & 0x7fffffffu);
            mtv[623] = mtv[396] ^ (y >> 1) ^ mag01[y & 1u];
            mti = 0;
        }
        uint32_t y = mtv[mti++];
        y ^= (y >> 11);
        y ^= (y << 7) & 0x9d2c5680u;
        y ^= (y << 15) & 0xefc60000u;
        y ^= (y >> 18);
        return y;
    }
};

// ---- the kernel, with "for every thread" loops where the CTA runs in parallel; `order` permutes the threads -------
static void emulated_kernel(uint32_t* state, uint64_t high, uint64_t magic, int64_t n, int64_t* out, const int* order) {
    static uint32_t buf[2][mt::N];
    for (int k = 0; k < mt::CTA; ++k) {
        const int t = order[k];
        for (int i = t; i < mt::N; i += mt::CTA) buf[0][i] = state[i];
    }
    int pos = (int)state[mt::N];
    int cur = 0;
    int64_t done = 0;
    while (true) {
        const int64_t left = n - done;
        const int take = (int)((int64_t)(mt::N - pos) < left ? (mt::N - pos) : left);
        for (int k = 0; k < mt::CTA; ++k) mt::emit(buf[cur], pos, take, order[k], magic, high, out + done);
        done += take;
        pos += take;
        if (done >= n) break;
        for (int phase = 0; phase < 3; ++phase)            // __syncthreads() after each
            for (int k = 0; k < mt::CTA; ++k) mt::regen_phase(phase, buf[cur], buf[cur ^ 1], order[k]);
        cur ^= 1;
        pos = 0;
    }
    for (int k = 0; k < mt::CTA; ++k) {
        const int t = order[k];
        for (int i = t; i < mt::N; i += mt::CTA) state[i] = buf[cur][i];
    }
    state[mt::N] = (uint32_t)pos;
}

int main() {
    srand(12345);
    int order[mt::CTA];
    for (int trial = 0; trial < 400; ++trial) {
        // thread order: forward, backward, or a shuffle — the phases must not depend on it
        for (int k = 0; k < mt::CTA; ++k) order[k] = (trial % 3 == 1) ? mt::CTA - 1 - k : k;
        if (trial % 3 == 2)
            for (int k = mt::CTA - 1; k > 0; --k) { int j = rand() % (k + 1); int tmp = order[k]; order[k] = order[j]; order[j] = tmp; }
        uint32_t state[mt::N + 1];
        for (int i = 0; i < mt::N; ++i) state[i] = ((uint32_t)rand() << 17) ^ ((uint32_t)rand() << 3) ^ (uint32_t)rand();
        const int pos0 = (trial % 5 == 0) ? mt::N : (trial % 5 == 1) ? 0 : rand() % (mt::N + 1);
        state[mt::N] = (uint32_t)pos0;
        static const uint64_t highs[] = {1, 2, 3, 16000, 999983, (1ull << 28) - 1, (1ull << 32) - 1, 1ull << 31};
        const uint64_t high = highs[trial % 8];
        RefMT ref;
        memcpy(ref.mtv, state, sizeof(ref.mtv));
        ref.mti = pos0;
        int64_t total = 0;
        for (int call = 0; call < 3; ++call) {             // consecutive launches continue the stream
            const int64_t n = (call == 1) ? rand() % 5 : rand() % 4000;
            std::vector<int64_t> got((size_t)n + 1, -7), want((size_t)n + 1, -7);
            emulated_kernel(state, high, mt::fastmod_magic(high), n, got.data(), order);
            for (int64_t i = 0; i < n; ++i) want[(size_t)i] = (int64_t)(ref.next() % high);
            if (got != want) { printf("MISMATCH trial %d call %d n %lld high %llu\n", trial, call, (long long)n, (unsigned long long)high); return 1; }
            total += n;
        }
        // generator state afterwards: same words wherever both are defined, same position
        if (total > 0 || pos0 < mt::N) {
            if ((int)state[mt::N] != ref.mti && !(ref.mti == pos0 && total == 0)) { printf("POSITION trial %d: %u vs %d\n", trial, state[mt::N], ref.mti); return 1; }
            if (memcmp(state, ref.mtv, sizeof(ref.mtv)) != 0) { printf("STATE trial %d\n", trial); return 1; }
        }
    }
    printf("MT_EMULATION_OK\n");
    return 0;
}
