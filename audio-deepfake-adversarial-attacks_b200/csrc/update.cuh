// Attack update rules and min-max scaling — see update.cu.
#pragma once
#include "common.cuh"

namespace advb {

constexpr int ROW_CHUNKS = 8;  // partial row reductions per clip (summed in fixed order by the consumer)

int pgd_start(const float* x, const float* noise, float* adv, int64_t n, cudaStream_t stream);
int fgsm_step(const float* x, const float* g, float* adv, float eps, int64_t n, cudaStream_t stream);
int pgd_step(const float* x, const float* g, float* adv, float eps, float alpha, int64_t n, cudaStream_t stream);
// PGDL2: partial (B, ROW_CHUNKS) scratch x2
int pgdl2_step(const float* x, const float* g, float* adv, float eps, float alpha, float eps_div, int B, int T,
               float* partial_g, float* partial_d, cudaStream_t stream);
int minmax_scale(const float* x, float* x01, float* mn, float* mx, int B, int T, cudaStream_t stream);
int minmax_revert(const float* x01, const float* mn, const float* mx, float* x, int B, int T, cudaStream_t stream);

}  // namespace advb
