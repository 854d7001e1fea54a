// mt19937.cuh — the arithmetic of MT19937 as torch's CPU generator (at::mt19937) and numpy's legacy RandomState use
// it, written once for the host replay (dccf_confounder_draw), the device kernel (k_confounder_draw) and the host
// emulation of that kernel's thread schedule in tests/mt_emulate.cpp (plain C++: DCCF_HD expands to nothing there).
#pragma once
#include <stdint.h>

#ifndef DCCF_HD
#ifdef __CUDACC__
#define DCCF_HD __host__ __device__ __forceinline__
#else
#define DCCF_HD inline
#endif
#endif

namespace dccf {
namespace mt {

constexpr int N = 624, M = 397, BACK = N - M;      // BACK = 227: how far the recurrence reaches inside a generation
constexpr int CTA = 256;                           // threads of the device kernel (>= BACK)

DCCF_HD uint32_t mix(uint32_t a, uint32_t b, uint32_t far_word) {
    const uint32_t y = (a & 0x80000000u) | (b & 0x7fffffffu);
    return far_word ^ (y >> 1) ^ ((0u - (y & 1u)) & 0x9908b0dfu);
}

DCCF_HD uint32_t temper(uint32_t y) {
    y ^= (y >> 11);
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= (y >> 18);
    return y;
}

// v % d as a multiply-high (Lemire, Kaser, Kurz 2019), exact for every 32-bit v and 0 < d < 2^32:
// magic = 2^64 / d rounded up (0 for d = 1, which yields 0 as it must)
DCCF_HD uint64_t fastmod_magic(uint64_t d) { return ~0ull / d + 1ull; }
DCCF_HD uint64_t mulhi64(uint64_t a, uint64_t b) {
#ifdef __CUDA_ARCH__
    return __umul64hi(a, b);
#else
    return (uint64_t)(((unsigned __int128)a * b) >> 64);
#endif
}
DCCF_HD int64_t fastmod(uint32_t v, uint64_t magic, uint64_t d) { return (int64_t)mulhi64(magic * (uint64_t)v, d); }

// ---- the device kernel's schedule, one function per barrier-separated phase, executed by every thread t --------
// Inside a generation  new[i] = mix(old[i], old[i+1], far)  with
//     i in [  0, 227): far = old[i + 397]
//     i in [227, 454): far = new[i - 227]        (phase 0 wrote it)
//     i in [454, 623): far = new[i - 227]        (phase 1 wrote it)
//     i = 623        : mix(old[623], new[0], new[396])
DCCF_HD void regen_phase(int phase, const uint32_t* o, uint32_t* w, int t) {
    if (phase == 0) {
        if (t < BACK) w[t] = mix(o[t], o[t + 1], o[t + M]);
    } else if (phase == 1) {
        if (t < BACK) w[BACK + t] = mix(o[BACK + t], o[BACK + t + 1], w[t]);
    } else {
        if (t < N - 1 - 2 * BACK) w[2 * BACK + t] = mix(o[2 * BACK + t], o[2 * BACK + t + 1], w[BACK + t]);
        else if (t == N - 1 - 2 * BACK) w[N - 1] = mix(o[N - 1], w[0], w[M - 1]);
    }
}
// ids from words gen[pos .. pos + take) of the current generation
DCCF_HD void emit(const uint32_t* gen, int pos, int take, int t, uint64_t magic, uint64_t high, int64_t* out) {
    for (int i = t; i < take; i += CTA) out[i] = fastmod(temper(gen[pos + i]), magic, high);
}

}  // namespace mt
}  // namespace dccf
