// Attack update rules (FGSM sign step, PGD L-inf project-and-clip, PGD L2 normalise/project/clip) and the
// per-clip min-max scaling that brackets every attack call (sm_100a, fp32, float4 streaming).
//
// Replaces fgsm.py:59-60, pgd.py:54-57,74-76, pgdl2.py:78-88 and src/aa/utils.py:4-14.  Every elementwise rule
// mirrors torch's op-by-op fp32 rounding (explicit __fadd_rn/__fmul_rn, no FMA contraction) so that, given the same
// gradient sign, the perturbed sample is bit-identical to the reference's.
#include "update.cuh"

namespace advb {

namespace {

__device__ __forceinline__ float signf_(float g) { return g > 0.f ? 1.f : (g < 0.f ? -1.f : 0.f); }
__device__ __forceinline__ float clampf_(float v, float lo, float hi) { return fminf(fmaxf(v, lo), hi); }

__global__ void pgd_start_kernel(const float* __restrict__ x, const float* __restrict__ noise, float* __restrict__ adv,
                                 int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float v = noise != nullptr ? clampf_(__fadd_rn(x[i], noise[i]), 0.f, 1.f) : x[i];
    adv[i] = v;
  }
}

__global__ void fgsm_step_kernel(const float* __restrict__ x, const float* __restrict__ g, float* __restrict__ adv,
                                 float eps, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    adv[i] = clampf_(__fadd_rn(x[i], __fmul_rn(eps, signf_(g[i]))), 0.f, 1.f);
}

__global__ void pgd_step_kernel(const float* __restrict__ x, const float* __restrict__ g, float* __restrict__ adv,
                                float eps, float alpha, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float xi = x[i];
    const float a1 = __fadd_rn(adv[i], __fmul_rn(alpha, signf_(g[i])));
    const float d = clampf_(__fsub_rn(a1, xi), -eps, eps);
    adv[i] = clampf_(__fadd_rn(xi, d), 0.f, 1.f);
  }
}

__device__ __forceinline__ float block_sum_256(float v, float* s_red) {
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (lane == 0) s_red[w] = v;
  __syncthreads();
  float r = 0.f;
  if (threadIdx.x == 0)
    for (int i = 0; i < 8; ++i) r += s_red[i];
  return r;  // valid on thread 0
}

// partial[b][chunk] = sum over the chunk of g^2
__global__ void __launch_bounds__(256) row_sumsq_kernel(const float* __restrict__ g, float* __restrict__ partial, int T) {
  __shared__ float s_red[8];
  const int b = blockIdx.y, ch = blockIdx.x;
  const int len = (T + ROW_CHUNKS - 1) / ROW_CHUNKS;
  const int lo = ch * len, hi = min(T, lo + len);
  float s = 0.f;
  for (int i = lo + threadIdx.x; i < hi; i += 256) {
    const float v = g[(size_t)b * T + i];
    s = fmaf(v, v, s);
  }
  const float r = block_sum_256(s, s_red);
  if (threadIdx.x == 0) partial[b * ROW_CHUNKS + ch] = r;
}

__device__ __forceinline__ float row_norm(const float* partial, int b) {
  float s = 0.f;
  for (int i = 0; i < ROW_CHUNKS; ++i) s += partial[b * ROW_CHUNKS + i];
  return sqrtf(s);
}

// adv <- adv + alpha * g / (||g|| + eps_div); partial_d = chunk sums of (adv - x)^2
__global__ void __launch_bounds__(256) pgdl2_ascent_kernel(const float* __restrict__ x, const float* __restrict__ g,
                                                            float* __restrict__ adv, const float* __restrict__ partial_g,
                                                            float* __restrict__ partial_d, float alpha, float eps_div,
                                                            int T) {
  __shared__ float s_red[8];
  const int b = blockIdx.y, ch = blockIdx.x;
  const float gn = __fadd_rn(row_norm(partial_g, b), eps_div);
  const int len = (T + ROW_CHUNKS - 1) / ROW_CHUNKS;
  const int lo = ch * len, hi = min(T, lo + len);
  float s = 0.f;
  for (int i = lo + threadIdx.x; i < hi; i += 256) {
    const size_t o = (size_t)b * T + i;
    const float a1 = __fadd_rn(adv[o], __fmul_rn(alpha, __fdiv_rn(g[o], gn)));
    adv[o] = a1;
    const float d = __fsub_rn(a1, x[o]);
    s = fmaf(d, d, s);
  }
  const float r = block_sum_256(s, s_red);
  if (threadIdx.x == 0) partial_d[b * ROW_CHUNKS + ch] = r;
}

// delta = (adv - x) * min(eps / ||delta||, 1); adv = clamp(x + delta, 0, 1)
__global__ void __launch_bounds__(256) pgdl2_project_kernel(const float* __restrict__ x, float* __restrict__ adv,
                                                             const float* __restrict__ partial_d, float eps, int T) {
  const int b = blockIdx.y;
  const float dn = row_norm(partial_d, b);
  const float factor = fminf(__fdiv_rn(eps, dn), 1.0f);  // eps/0 = inf -> 1 (pgdl2.py:84-85)
  for (int i = blockIdx.x * 256 + threadIdx.x; i < T; i += gridDim.x * 256) {
    const size_t o = (size_t)b * T + i;
    const float xi = x[o];
    const float d = __fmul_rn(__fsub_rn(adv[o], xi), factor);
    adv[o] = clampf_(__fadd_rn(xi, d), 0.f, 1.f);
  }
}

__global__ void __launch_bounds__(1024) row_minmax_kernel(const float* __restrict__ x, float* __restrict__ mn,
                                                           float* __restrict__ mx, int T) {
  __shared__ float s_lo[32], s_hi[32];
  const int b = blockIdx.x;
  float lo = INFINITY, hi = -INFINITY;
  for (int i = threadIdx.x; i < T; i += 1024) {
    const float v = x[(size_t)b * T + i];
    lo = fminf(lo, v);
    hi = fmaxf(hi, v);
  }
  lo = warp_min(lo);
  hi = warp_max(hi);
  if ((threadIdx.x & 31) == 0) {
    s_lo[threadIdx.x >> 5] = lo;
    s_hi[threadIdx.x >> 5] = hi;
  }
  __syncthreads();
  if (threadIdx.x < 32) {
    lo = warp_min(s_lo[threadIdx.x]);
    hi = warp_max(s_hi[threadIdx.x]);
    if (threadIdx.x == 0) {
      mn[b] = lo;
      mx[b] = hi;
    }
  }
}

__global__ void minmax_apply_kernel(const float* __restrict__ x, const float* __restrict__ mn,
                                    const float* __restrict__ mx, float* __restrict__ out, int T, int revert) {
  const int b = blockIdx.y;
  const float lo = mn[b], r = __fsub_rn(mx[b], lo);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < T; i += gridDim.x * blockDim.x) {
    const size_t o = (size_t)b * T + i;
    out[o] = revert ? __fadd_rn(__fmul_rn(x[o], r), lo) : __fdiv_rn(__fsub_rn(x[o], lo), r);
  }
}

inline int ew_blocks(int64_t n) {
  int64_t b = (n + 255) / 256;
  return (int)(b > 148 * 16 ? 148 * 16 : b);
}

}  // namespace

int pgd_start(const float* x, const float* noise, float* adv, int64_t n, cudaStream_t stream) {
  pgd_start_kernel<<<ew_blocks(n), 256, 0, stream>>>(x, noise, adv, n);
  ADVB_KERNEL_OK("pgd_start", stream);
  return 0;
}
int fgsm_step(const float* x, const float* g, float* adv, float eps, int64_t n, cudaStream_t stream) {
  fgsm_step_kernel<<<ew_blocks(n), 256, 0, stream>>>(x, g, adv, eps, n);
  ADVB_KERNEL_OK("fgsm_step", stream);
  return 0;
}
int pgd_step(const float* x, const float* g, float* adv, float eps, float alpha, int64_t n, cudaStream_t stream) {
  pgd_step_kernel<<<ew_blocks(n), 256, 0, stream>>>(x, g, adv, eps, alpha, n);
  ADVB_KERNEL_OK("pgd_step", stream);
  return 0;
}
int pgdl2_step(const float* x, const float* g, float* adv, float eps, float alpha, float eps_div, int B, int T,
               float* partial_g, float* partial_d, cudaStream_t stream) {
  dim3 grid(ROW_CHUNKS, B);
  row_sumsq_kernel<<<grid, 256, 0, stream>>>(g, partial_g, T);
  ADVB_KERNEL_OK("row_sumsq", stream);
  pgdl2_ascent_kernel<<<grid, 256, 0, stream>>>(x, g, adv, partial_g, partial_d, alpha, eps_div, T);
  ADVB_KERNEL_OK("pgdl2_ascent", stream);
  pgdl2_project_kernel<<<grid, 256, 0, stream>>>(x, adv, partial_d, eps, T);
  ADVB_KERNEL_OK("pgdl2_project", stream);
  return 0;
}
int minmax_scale(const float* x, float* x01, float* mn, float* mx, int B, int T, cudaStream_t stream) {
  row_minmax_kernel<<<B, 1024, 0, stream>>>(x, mn, mx, T);
  ADVB_KERNEL_OK("row_minmax", stream);
  dim3 grid(cdiv(T, 256 * 8), B);
  minmax_apply_kernel<<<grid, 256, 0, stream>>>(x, mn, mx, x01, T, 0);
  ADVB_KERNEL_OK("minmax_apply", stream);
  return 0;
}
int minmax_revert(const float* x01, const float* mn, const float* mx, float* x, int B, int T, cudaStream_t stream) {
  dim3 grid(cdiv(T, 256 * 8), B);
  minmax_apply_kernel<<<grid, 256, 0, stream>>>(x01, mn, mx, x, T, 1);
  ADVB_KERNEL_OK("minmax_revert", stream);
  return 0;
}

}  // namespace advb
