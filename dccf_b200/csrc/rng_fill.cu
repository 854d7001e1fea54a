// rng_fill.cu — materialisers of the library's counter-based random streams.
//
// The fused kernels generate the feature noise (src/models/DCCF.py:87) and the dropout mask
// (DCCF.py:94) in registers (rng mode 2).  These two kernels write exactly the same values to
// memory, which (1) defines the streams for anyone who wants to reproduce a run with the reference
// code (feed the tensors in as mode 1 / monkey-patched draws) and (2) lets the tests check that the
// fused mode-2 kernels equal the mode-1 kernels fed with the materialised tensors.
#include "common.cuh"

namespace dccf {

__global__ void k_noise_fill(float* __restrict__ out, int64_t n_rows, int32_t F, float std, RngSpec spec, int64_t row0) {
    const RngKey key = resolve_rng_key(spec, DOMAIN_NOISE);
    const int quads = F / 4;
    const int64_t total = n_rows * quads;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const int64_t r = i / quads;
        const int q = (int)(i - r * quads);
        st4(out + (size_t)r * F + q * 4, noise_quad(key, (uint32_t)(row0 + r), (uint32_t)q, std));
    }
}

__global__ void k_mask_fill(float* __restrict__ out, int64_t n_rows, int32_t dim, float keep, float scale, RngSpec spec,
                            int64_t row0) {
    const RngKey key = resolve_rng_key(spec, DOMAIN_DROPOUT);
    const int quads = dim / 4;
    const int64_t total = n_rows * quads;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const int64_t r = i / quads;
        const int q = (int)(i - r * quads);
        st4(out + (size_t)r * dim + q * 4, dropout_quad(key, (uint32_t)(row0 + r), (uint32_t)q, keep, scale));
    }
}

static unsigned fill_grid(int64_t total) {
    int64_t g = (total + 255) / 256;
    if (g > 148 * 16) g = 148 * 16;
    if (g < 1) g = 1;
    return (unsigned)g;
}

}  // namespace dccf

using namespace dccf;

extern "C" int dccf_noise_fill(float* out, int64_t n_rows, int32_t feat_dim, float std, uint64_t seed, uint64_t offset,
                               int64_t row0, void* stream_) {
    DCCF_CHECK_ARG(out != nullptr, "dccf_noise_fill: null output");
    DCCF_CHECK_ARG(feat_dim > 0 && feat_dim % 4 == 0, "dccf_noise_fill: feat_dim=%d must be a positive multiple of 4", feat_dim);
    DCCF_CHECK_ARG(row0 >= 0 && row0 + n_rows < ((int64_t)1 << 32), "dccf_noise_fill: row range exceeds the 32-bit counter");
    if (n_rows <= 0) return DCCF_OK;
    RngSpec spec{seed, offset, nullptr};
    k_noise_fill<<<fill_grid(n_rows * (feat_dim / 4)), 256, 0, (cudaStream_t)stream_>>>(out, n_rows, feat_dim, std, spec, row0);
    DCCF_CHECK_LAUNCH("k_noise_fill");
    return DCCF_OK;
}

extern "C" int dccf_dropout_mask_fill(float* out, int64_t n_rows, int32_t dim, float p_drop, uint64_t seed,
                                      uint64_t offset, int64_t row0, void* stream_) {
    DCCF_CHECK_ARG(out != nullptr, "dccf_dropout_mask_fill: null output");
    DCCF_CHECK_ARG(dim > 0 && dim % 4 == 0, "dccf_dropout_mask_fill: dim=%d must be a positive multiple of 4", dim);
    DCCF_CHECK_ARG(p_drop >= 0.f && p_drop <= 1.f, "dccf_dropout_mask_fill: p_drop=%f outside [0,1]", (double)p_drop);
    DCCF_CHECK_ARG(row0 >= 0 && row0 + n_rows < ((int64_t)1 << 32), "dccf_dropout_mask_fill: row range exceeds the 32-bit counter");
    if (n_rows <= 0) return DCCF_OK;
    RngSpec spec{seed, offset, nullptr};
    const float keep = 1.0f - p_drop;
    const float scale = (p_drop < 1.0f) ? 1.0f / (1.0f - p_drop) : 0.0f;
    k_mask_fill<<<fill_grid(n_rows * (dim / 4)), 256, 0, (cudaStream_t)stream_>>>(out, n_rows, dim, keep, scale, spec, row0);
    DCCF_CHECK_LAUNCH("k_mask_fill");
    return DCCF_OK;
}
