// RawNet3 forward and exact input-gradient backward (contract: rawnet3.cuh).  Replaces src/models/rawnet3.py:73-137
// (forward), :151-158 (PreEmphasis), :176-182 (AFMS), :242-274 (Bottle2neck) and the autograd input gradient the attacks
// take through them (fgsm.py:56-57, pgd.py:71-72, pgdl2.py:76-77, fab.py:90-105); the sinc filters follow
// asteroid-filterbanks 0.4.0 ParamSincFB.filters (third party).
//
// All dense contractions (98 % of the model's 19.1 GMAC per 64 000-sample clip) go through gemm_run (tcgen05 3xTF32 by
// default); this file holds the memory-bound glue: pre-emphasis + InstanceNorm, log|.| + mean normalisation, max-pool +
// AFMS, attentive statistics pooling, and their backward counterparts.  Per-(clip, channel) reductions over time use one
// CTA of 32 channels x 32 row lanes (coalesced 128-byte rows, fixed summation order => deterministic).
#include "rawnet3.cuh"

#include <algorithm>

namespace advb {

namespace {

constexpr int C = 1024, W = 128, NS = 256, SK = 251, CA = 1536;
constexpr float PRE = 0.97f, IN_EPS = 1e-4f, BN_EPS = 1e-5f;

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }

// ---- reductions inside a (32, 32) block: sum / max over threadIdx.y for every threadIdx.x -----------------------
__device__ __forceinline__ float colsum32(float v, float (*red)[33]) {
  red[threadIdx.y][threadIdx.x] = v;
  __syncthreads();
  if (threadIdx.y == 0) {
    float s = 0.f;
    for (int i = 0; i < 32; ++i) s += red[i][threadIdx.x];
    red[0][threadIdx.x] = s;
  }
  __syncthreads();
  const float r = red[0][threadIdx.x];
  __syncthreads();
  return r;
}
__device__ __forceinline__ float colmax32(float v, float (*red)[33]) {
  red[threadIdx.y][threadIdx.x] = v;
  __syncthreads();
  if (threadIdx.y == 0) {
    float s = red[0][threadIdx.x];
    for (int i = 1; i < 32; ++i) s = fmaxf(s, red[i][threadIdx.x]);
    red[0][threadIdx.x] = s;
  }
  __syncthreads();
  const float r = red[0][threadIdx.x];
  __syncthreads();
  return r;
}
__device__ __forceinline__ double block_sum_d(double v, double* red) {  // blockDim.x threads (multiple of 32, <= 1024)
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) red[w] = v;
  __syncthreads();
  double s = 0.0;
  if (threadIdx.x == 0)
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) s += red[i];
  if (threadIdx.x == 0) red[0] = s;
  __syncthreads();
  s = red[0];
  __syncthreads();
  return s;
}

// ---- per-call parameter preparation -----------------------------------------------------------------------------
// BatchNorm(eval, affine): y = x * s + t, s = w / sqrt(rv + eps), t = b - rm * s.
__global__ void rn_fold_bn_kernel(const float* __restrict__ w, const float* __restrict__ b, const float* __restrict__ rm,
                                  const float* __restrict__ rv, float* __restrict__ s, float* __restrict__ t, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    const float sc = w[i] / sqrtf(rv[i] + BN_EPS);
    s[i] = sc;
    t[i] = b[i] - rm[i] * sc;
  }
}

// ParamSincFB.filters(): 128 cosine + 128 sine band-pass filters of 251 taps from the live low_hz_ / band_hz_.
__global__ void rn_sinc_filters_kernel(const float* __restrict__ low_hz, const float* __restrict__ band_hz,
                                       const float* __restrict__ window, const float* __restrict__ n_axis,
                                       float* __restrict__ filt) {
  const int c = blockIdx.x, k = threadIdx.x;
  if (k >= SK) return;
  const int i = c & 127;
  const bool is_sin = c >= 128;
  const float low = 50.f + fabsf(low_hz[i]);
  const float high = fminf(fmaxf(low + 50.f + fabsf(band_hz[i]), 50.f), 8000.f);
  const float band = high - low;
  float v;
  if (k == 125) {
    v = is_sin ? 0.f : 2.f * band;
  } else {
    const int kk = k < 125 ? k : 250 - k;
    const float n = n_axis[kk];
    const float fl = low * n, fh = high * n;
    const float num = is_sin ? (cosf(fl) - cosf(fh)) : (sinf(fh) - sinf(fl));
    v = (num / (n / 2.f)) * window[kk];
    if (is_sin && k > 125) v = -v;
  }
  filt[c * SK + k] = v / (2.f * band);
}

// ---- pre-emphasis + InstanceNorm1d(1, eps=1e-4, affine) ---------------------------------------------------------
__device__ __forceinline__ float pre_emph(const float* __restrict__ x, int t) {
  return x[t] - PRE * x[t == 0 ? 1 : t - 1];  // reflect pad: x[-1] := x[1]
}
__global__ void __launch_bounds__(1024) rn_pre_stats_kernel(const float* __restrict__ x, int T, double* __restrict__ stats) {
  __shared__ double red[32];
  const float* xb = x + (size_t)blockIdx.x * T;
  double s = 0.0, q = 0.0;
  for (int t = threadIdx.x; t < T; t += blockDim.x) {
    const double e = (double)pre_emph(xb, t);
    s += e;
    q += e * e;
  }
  s = block_sum_d(s, red);
  q = block_sum_d(q, red);
  if (threadIdx.x == 0) {
    const double mean = s / T, var = fmax(q / T - mean * mean, 0.0);
    stats[2 * blockIdx.x] = mean;
    stats[2 * blockIdx.x + 1] = 1.0 / sqrt(var + (double)IN_EPS);
  }
}
__global__ void rn_pre_apply_kernel(const float* __restrict__ x, const double* __restrict__ stats, const float* __restrict__ w,
                                    const float* __restrict__ b, float* __restrict__ out, int B, int T) {
  const size_t n = (size_t)B * T;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const int clip = (int)(i / T), t = (int)(i - (size_t)clip * T);
    const float mean = (float)stats[2 * clip], inv = (float)stats[2 * clip + 1];
    out[i] = (pre_emph(x + (size_t)clip * T, t) - mean) * inv * w[0] + b[0];
  }
}
// backward: reductions mean(g), mean(g * e_hat), then d/de and the transposed pre-emphasis
__global__ void __launch_bounds__(1024) rn_pre_bwd_stats_kernel(const float* __restrict__ x, const float* __restrict__ gn,
                                                                const double* __restrict__ stats, int T,
                                                                double* __restrict__ bst) {
  __shared__ double red[32];
  const float* xb = x + (size_t)blockIdx.x * T;
  const float* gb = gn + (size_t)blockIdx.x * T;
  const float mean = (float)stats[2 * blockIdx.x], inv = (float)stats[2 * blockIdx.x + 1];
  double s = 0.0, q = 0.0;
  for (int t = threadIdx.x; t < T; t += blockDim.x) {
    const float eh = (pre_emph(xb, t) - mean) * inv;
    s += (double)gb[t];
    q += (double)gb[t] * (double)eh;
  }
  s = block_sum_d(s, red);
  q = block_sum_d(q, red);
  if (threadIdx.x == 0) {
    bst[2 * blockIdx.x] = s / T;
    bst[2 * blockIdx.x + 1] = q / T;
  }
}
__global__ void rn_pre_bwd_apply_kernel(const float* __restrict__ x, const float* __restrict__ gn,
                                        const double* __restrict__ stats, const double* __restrict__ bst,
                                        const float* __restrict__ w, float* __restrict__ gx, int B, int T) {
  const size_t n = (size_t)B * T;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const int clip = (int)(i / T), t = (int)(i - (size_t)clip * T);
    const float* xb = x + (size_t)clip * T;
    const float* gb = gn + (size_t)clip * T;
    const float mean = (float)stats[2 * clip], inv = (float)stats[2 * clip + 1];
    const float mg = (float)bst[2 * clip], mge = (float)bst[2 * clip + 1];
    const float k = w[0] * inv;
    auto ge = [&](int u) { return k * (gb[u] - mg - (pre_emph(xb, u) - mean) * inv * mge); };
    float v = ge(t);
    if (t + 1 < T) v -= PRE * ge(t + 1);
    if (t == 1) v -= PRE * ge(0);
    gx[i] = v;
  }
}

// ---- sinc post-processing: u = log(|s| + 1e-6), x0 = u - mean_t(u) ----------------------------------------------
__global__ void __launch_bounds__(1024) rn_sinc_post_kernel(const float* __restrict__ S, float* __restrict__ x0, int L0, int Tp,
                                                            int pad) {
  __shared__ float red[32][33];
  const int c = blockIdx.x * 32 + threadIdx.x, b = blockIdx.y;
  const float* sb = S + (size_t)b * L0 * NS + c;
  float acc = 0.f;
  for (int l = threadIdx.y; l < L0; l += 32) acc += logf(fabsf(sb[(size_t)l * NS]) + 1e-6f);
  const float mean = colsum32(acc, red) / (float)L0;
  float* ob = x0 + ((size_t)b * Tp + pad) * NS + c;
  for (int l = threadIdx.y; l < L0; l += 32) ob[(size_t)l * NS] = logf(fabsf(sb[(size_t)l * NS]) + 1e-6f) - mean;
}
__global__ void __launch_bounds__(1024) rn_sinc_post_bwd_kernel(const float* __restrict__ S, const float* __restrict__ gx0,
                                                                float* __restrict__ GS, int L0, int Tp, int pad) {
  __shared__ float red[32][33];
  const int c = blockIdx.x * 32 + threadIdx.x, b = blockIdx.y;
  const float* gb = gx0 + ((size_t)b * Tp + pad) * NS + c;
  float acc = 0.f;
  for (int l = threadIdx.y; l < L0; l += 32) acc += gb[(size_t)l * NS];
  const float mean = colsum32(acc, red) / (float)L0;
  const float* sb = S + (size_t)b * L0 * NS + c;
  float* ob = GS + (size_t)b * L0 * NS + c;
  for (int l = threadIdx.y; l < L0; l += 32) {
    const float s = sb[(size_t)l * NS];
    const float sg = s > 0.f ? 1.f : (s < 0.f ? -1.f : 0.f);
    ob[(size_t)l * NS] = (gb[(size_t)l * NS] - mean) * sg / (fabsf(s) + 1e-6f);
  }
}
// transposed strided convolution: g_n[tau] = sum_l Z[l][tau - 10 l]
__global__ void rn_col2im_kernel(const float* __restrict__ Z, float* __restrict__ gn, int B, int T, int L0) {
  const size_t n = (size_t)B * T;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const int clip = (int)(i / T), tau = (int)(i - (size_t)clip * T);
    int lo = tau - (SK - 1);
    lo = lo <= 0 ? 0 : (lo + 9) / 10;
    const int hi = min(L0 - 1, tau / 10);
    const float* zb = Z + (size_t)clip * L0 * NS;
    float acc = 0.f;
    for (int l = lo; l <= hi; ++l) acc += zb[(size_t)l * NS + (tau - 10 * l)];
    gn[i] = acc;
  }
}

// ---- generic helpers ---------------------------------------------------------------------------------------------
// dst[r][0..ncols) = src[r][0..ncols) for every row (float4)
__global__ void rn_copy_cols_kernel(const float* __restrict__ src, int lds, float* __restrict__ dst, int ldd, size_t rows,
                                    int ncols) {
  const int c4n = ncols / 4;
  const size_t n = rows * c4n;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const size_t r = i / c4n;
    const int c = (int)(i - r * c4n) * 4;
    st4(dst + r * ldd + c, ld4(src + r * lds + c));
  }
}
// valid rows only: dst[row(b,t)][c] = src[row(b,t)][c] * scale[c] * (mask[row][c] != 0), same padded row space
__global__ void rn_gate_kernel(const float* __restrict__ src, int lds, const float* __restrict__ scale,
                               const unsigned char* __restrict__ mask, int ldm, float* __restrict__ dst, int ldd, int B, int T,
                               int Tp, int pad, int ncols) {
  const int c4n = ncols / 4;
  const size_t n = (size_t)B * T * c4n;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const size_t bt = i / c4n;
    const int c = (int)(i - bt * c4n) * 4;
    const int b = (int)(bt / T), t = (int)(bt - (size_t)b * T);
    const size_t r = (size_t)b * Tp + pad + t;
    const float4 v = ld4(src + r * lds + c), s = ld4(scale + c);
    const uchar4 m = *reinterpret_cast<const uchar4*>(mask + r * ldm + c);
    st4(dst + r * ldd + c, make_float4(m.x ? v.x * s.x : 0.f, m.y ? v.y * s.y : 0.f, m.z ? v.z * s.z : 0.f, m.w ? v.w * s.w : 0.f));
  }
}
// out[b][n] = act(bias[n] + sum_k W[n * ldw + k] * in[b * K + k]); one warp per n, all clips.  act: 0 none, 1 sigmoid
__global__ void __launch_bounds__(256) rn_linear_kernel(const float* __restrict__ Wm, int ldw, const float* __restrict__ bias,
                                                        const float* __restrict__ in, float* __restrict__ out, int B, int N,
                                                        int K, int act) {
  const int n = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (n >= N) return;
  const float* wr = Wm + (size_t)n * ldw;
  for (int b = 0; b < B; ++b) {
    const float* ib = in + (size_t)b * K;
    float acc = 0.f;
    for (int k = lane; k < K; k += 32) acc = fmaf(wr[k], ib[k], acc);
    acc = warp_sum(acc);
    if (lane == 0) {
      acc += bias != nullptr ? bias[n] : 0.f;
      out[(size_t)b * N + n] = act == 1 ? 1.f / (1.f + expf(-acc)) : acc;
    }
  }
}
// out[b][k] = scale * sum_n W[n * ldw + k] * in[b * N + n];  block = 64 k x 16 partitions of n, reduced in shared memory
__global__ void __launch_bounds__(1024) rn_linear_t_kernel(const float* __restrict__ Wm, int ldw, const float* __restrict__ in,
                                                           float* __restrict__ out, int N, int K, float scale) {
  __shared__ float red[16][65];
  const int kx = threadIdx.x, ny = threadIdx.y;
  const int k = blockIdx.x * 64 + kx, b = blockIdx.y;
  const float* ib = in + (size_t)b * N;
  float acc = 0.f;
  if (k < K)
    for (int n = ny; n < N; n += 16) acc = fmaf(Wm[(size_t)n * ldw + k], ib[n], acc);
  red[ny][kx] = acc;
  __syncthreads();
  if (ny == 0 && k < K) {
    float s = 0.f;
    for (int i = 0; i < 16; ++i) s += red[i][kx];
    out[(size_t)b * K + k] = s * scale;
  }
}

// ---- max-pool over time + AFMS ----------------------------------------------------------------------------------
// P[b][t'][c] = max_j Y[b][pad + p t' + j][c] (first maximum wins, as ATen), arg-max j, pm[b][c] = mean_t' P
__global__ void __launch_bounds__(1024) rn_pool_mean_kernel(const float* __restrict__ Y, int Tp, int pad, int p, int To,
                                                            float* __restrict__ P, unsigned char* __restrict__ arg,
                                                            float* __restrict__ pm) {
  __shared__ float red[32][33];
  const int c = blockIdx.x * 32 + threadIdx.x, b = blockIdx.y;
  const float* yb = Y + ((size_t)b * Tp + pad) * C + c;
  float acc = 0.f;
  for (int t = threadIdx.y; t < To; t += 32) {
    float best = yb[(size_t)(p * t) * C];
    int bj = 0;
    for (int j = 1; j < p; ++j) {
      const float v = yb[(size_t)(p * t + j) * C];
      if (v > best) best = v, bj = j;
    }
    P[((size_t)b * To + t) * C + c] = best;
    arg[((size_t)b * To + t) * C + c] = (unsigned char)bj;
    acc += best;
  }
  const float s = colsum32(acc, red);
  if (threadIdx.y == 0) pm[(size_t)b * C + c] = s / (float)To;
}
// out[row(b, t')][c] = (P + alpha[c]) * yv[b][c]
__global__ void rn_afms_scale_kernel(const float* __restrict__ P, const float* __restrict__ alpha, const float* __restrict__ yv,
                                     float* __restrict__ out, int ldo, int Tpo, int pado, int B, int To) {
  const size_t n = (size_t)B * To * (C / 4);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const size_t bt = i / (C / 4);
    const int c = (int)(i - bt * (C / 4)) * 4;
    const int b = (int)(bt / To), t = (int)(bt - (size_t)b * To);
    const float4 v = ld4(P + bt * C + c), a = ld4(alpha + c), g = ld4(yv + (size_t)b * C + c);
    st4(out + ((size_t)b * Tpo + pado + t) * ldo + c,
        make_float4((v.x + a.x) * g.x, (v.y + a.y) * g.y, (v.z + a.z) * g.z, (v.w + a.w) * g.w));
  }
}
// gyv[b][c] = sum_t' gout[b To + t'][c] * (P + alpha[c]);  gz = gyv * yv * (1 - yv)
__global__ void __launch_bounds__(1024) rn_afms_bwd_sum_kernel(const float* __restrict__ gout, int ldg, const float* __restrict__ P,
                                                               const float* __restrict__ alpha, const float* __restrict__ yv,
                                                               int To, float* __restrict__ gz) {
  __shared__ float red[32][33];
  const int c = blockIdx.x * 32 + threadIdx.x, b = blockIdx.y;
  const float al = alpha[c];
  float acc = 0.f;
  for (int t = threadIdx.y; t < To; t += 32)
    acc = fmaf(gout[((size_t)b * To + t) * ldg + c], P[((size_t)b * To + t) * C + c] + al, acc);
  const float s = colsum32(acc, red);
  if (threadIdx.y == 0) {
    const float g = yv[(size_t)b * C + c];
    gz[(size_t)b * C + c] = s * g * (1.f - g);
  }
}
// un-AFMS + un-pool: GY[row(b,t)][c] = [arg == t % p] (gout[b][t / p][c] * yv + gm[b][c]);  GC3 = GY * bn3_s * m3
__global__ void rn_pool_bwd_kernel(const float* __restrict__ gout, int ldg, const float* __restrict__ yv,
                                   const float* __restrict__ gm, const unsigned char* __restrict__ arg,
                                   const float* __restrict__ bn3_s, const unsigned char* __restrict__ m3, float* __restrict__ GY,
                                   float* __restrict__ GC3, int B, int T, int Tp, int pad, int p, int To) {
  const size_t n = (size_t)B * T * (C / 4);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const size_t bt = i / (C / 4);
    const int c = (int)(i - bt * (C / 4)) * 4;
    const int b = (int)(bt / T), t = (int)(bt - (size_t)b * T);
    const int to = t / p, j = t - to * p;
    float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
    if (to < To) {
      const size_t o = ((size_t)b * To + to);
      const uchar4 a = *reinterpret_cast<const uchar4*>(arg + o * C + c);
      const float4 go = ld4(gout + o * ldg + c), y = ld4(yv + (size_t)b * C + c), m = ld4(gm + (size_t)b * C + c);
      g.x = a.x == j ? fmaf(go.x, y.x, m.x) : 0.f;
      g.y = a.y == j ? fmaf(go.y, y.y, m.y) : 0.f;
      g.z = a.z == j ? fmaf(go.z, y.z, m.z) : 0.f;
      g.w = a.w == j ? fmaf(go.w, y.w, m.w) : 0.f;
    }
    const size_t r = (size_t)b * Tp + pad + t;
    st4(GY + r * C + c, g);
    const float4 s = ld4(bn3_s + c);
    const uchar4 mk = *reinterpret_cast<const uchar4*>(m3 + r * C + c);
    st4(GC3 + r * C + c, make_float4(mk.x ? g.x * s.x : 0.f, mk.y ? g.y * s.y : 0.f, mk.z ? g.z * s.z : 0.f, mk.w ? g.w * s.w : 0.f));
  }
}
// M1 = MaxPool1d(3)(x1): cat4[b T3 + t][0..1024) and its arg-max
__global__ void rn_pool3_kernel(const float* __restrict__ x1, int Tp, int pad, float* __restrict__ cat4, unsigned char* __restrict__ arg1,
                                int B, int T3) {
  const size_t n = (size_t)B * T3 * (C / 4);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const size_t bt = i / (C / 4);
    const int c = (int)(i - bt * (C / 4)) * 4;
    const int b = (int)(bt / T3), t = (int)(bt - (size_t)b * T3);
    const float* src = x1 + ((size_t)b * Tp + pad + 3 * t) * C + c;
    float4 best = ld4(src);
    uchar4 bj = make_uchar4(0, 0, 0, 0);
    for (int j = 1; j < 3; ++j) {
      const float4 v = ld4(src + (size_t)j * C);
      if (v.x > best.x) best.x = v.x, bj.x = j;
      if (v.y > best.y) best.y = v.y, bj.y = j;
      if (v.z > best.z) best.z = v.z, bj.z = j;
      if (v.w > best.w) best.w = v.w, bj.w = j;
    }
    st4(cat4 + bt * (3 * C) + c, best);
    *reinterpret_cast<uchar4*>(arg1 + bt * C + c) = bj;
  }
}
// layer3 input = mp3(x1) + x2
__global__ void rn_add12_kernel(const float* __restrict__ cat4, float* __restrict__ xin, int Tp, int pad, int B, int T3) {
  const size_t n = (size_t)B * T3 * (C / 4);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const size_t bt = i / (C / 4);
    const int c = (int)(i - bt * (C / 4)) * 4;
    const int b = (int)(bt / T3), t = (int)(bt - (size_t)b * T3);
    const float4 u = ld4(cat4 + bt * (3 * C) + c), v = ld4(cat4 + bt * (3 * C) + C + c);
    st4(xin + ((size_t)b * Tp + pad + t) * C + c, make_float4(u.x + v.x, u.y + v.y, u.z + v.z, u.w + v.w));
  }
}
// backward of the sum: gcat4[:, 0:1024] += g, gcat4[:, 1024:2048] += g
__global__ void rn_add12_bwd_kernel(const float* __restrict__ gx3, int Tp, int pad, float* __restrict__ gcat4, int B, int T3) {
  const size_t n = (size_t)B * T3 * (C / 4);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const size_t bt = i / (C / 4);
    const int c = (int)(i - bt * (C / 4)) * 4;
    const int b = (int)(bt / T3), t = (int)(bt - (size_t)b * T3);
    const float4 g = ld4(gx3 + ((size_t)b * Tp + pad + t) * C + c);
    float* p0 = gcat4 + bt * (3 * C) + c;
    float4 u = ld4(p0), v = ld4(p0 + C);
    st4(p0, make_float4(u.x + g.x, u.y + g.y, u.z + g.z, u.w + g.w));
    st4(p0 + C, make_float4(v.x + g.x, v.y + g.y, v.z + g.z, v.w + g.w));
  }
}
// g_x1[b T2 + t][c] = gx2[row(b,t)][c] + [arg1 == t % 3] gcat4[b T3 + t / 3][c]
__global__ void rn_x1_grad_kernel(const float* __restrict__ gx2, int Tp, int pad, const float* __restrict__ gcat4,
                                  const unsigned char* __restrict__ arg1, float* __restrict__ gx1, int B, int T2, int T3) {
  const size_t n = (size_t)B * T2 * (C / 4);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const size_t bt = i / (C / 4);
    const int c = (int)(i - bt * (C / 4)) * 4;
    const int b = (int)(bt / T2), t = (int)(bt - (size_t)b * T2);
    float4 g = ld4(gx2 + ((size_t)b * Tp + pad + t) * C + c);
    const int to = t / 3, j = t - 3 * to;
    if (to < T3) {
      const size_t o = (size_t)b * T3 + to;
      const uchar4 a = *reinterpret_cast<const uchar4*>(arg1 + o * C + c);
      const float4 m = ld4(gcat4 + o * (3 * C) + c);
      if (a.x == j) g.x += m.x;
      if (a.y == j) g.y += m.y;
      if (a.z == j) g.z += m.z;
      if (a.w == j) g.w += m.w;
    }
    st4(gx1 + bt * C + c, g);
  }
}

// ---- attentive statistics pooling -------------------------------------------------------------------------------
// stats[b][0][c] = mean_t H, stats[b][1][c] = sqrt(clamp(var_unbiased, 1e-4, 1e4)), var[b][c]
__global__ void __launch_bounds__(1024) rn_asp_stats_kernel(const float* __restrict__ H, int T3, float* __restrict__ stats,
                                                            float* __restrict__ var) {
  __shared__ float red[32][33];
  const int c = blockIdx.x * 32 + threadIdx.x, b = blockIdx.y;
  const float* hb = H + (size_t)b * T3 * CA + c;
  float acc = 0.f;
  for (int t = threadIdx.y; t < T3; t += 32) acc += hb[(size_t)t * CA];
  const float mean = colsum32(acc, red) / (float)T3;
  acc = 0.f;
  for (int t = threadIdx.y; t < T3; t += 32) {
    const float d = hb[(size_t)t * CA] - mean;
    acc = fmaf(d, d, acc);
  }
  const float v = colsum32(acc, red) / (float)(T3 - 1);
  if (threadIdx.y == 0) {
    stats[(size_t)b * 2 * CA + c] = mean;
    stats[(size_t)b * 2 * CA + CA + c] = sqrtf(fminf(fmaxf(v, 1e-4f), 1e4f));
    var[(size_t)b * CA + c] = v;
  }
}
// softmax over time of E, mu = sum H w, sg = sqrt(clamp(sum H^2 w - mu^2, 1e-4, 1e4))
__global__ void __launch_bounds__(1024) rn_asp_pool_kernel(const float* __restrict__ H, const float* __restrict__ E, int T3,
                                                           float* __restrict__ emax, float* __restrict__ Zs, float* __restrict__ vq,
                                                           float* __restrict__ m2s, float* __restrict__ pooled) {
  __shared__ float red[32][33];
  const int c = blockIdx.x * 32 + threadIdx.x, b = blockIdx.y;
  const float* hb = H + (size_t)b * T3 * CA + c;
  const float* eb = E + (size_t)b * T3 * CA + c;
  float mx = -INFINITY;
  for (int t = threadIdx.y; t < T3; t += 32) mx = fmaxf(mx, eb[(size_t)t * CA]);
  mx = colmax32(mx, red);
  float z = 0.f, s1 = 0.f, s2 = 0.f;
  for (int t = threadIdx.y; t < T3; t += 32) {
    const float e = expf(eb[(size_t)t * CA] - mx), h = hb[(size_t)t * CA];
    z += e;
    s1 = fmaf(h, e, s1);
    s2 = fmaf(h * h, e, s2);
  }
  z = colsum32(z, red);
  s1 = colsum32(s1, red);
  s2 = colsum32(s2, red);
  if (threadIdx.y == 0) {
    const float mu = s1 / z, m2 = s2 / z, v = m2 - mu * mu;
    const size_t o = (size_t)b * CA + c;
    emax[o] = mx, Zs[o] = z, vq[o] = v, m2s[o] = m2;
    pooled[(size_t)b * 2 * CA + c] = mu;
    pooled[(size_t)b * 2 * CA + CA + c] = sqrtf(fminf(fmaxf(v, 1e-4f), 1e4f));
  }
}
// logit[b] = fc6(bn5(pooled))
__global__ void __launch_bounds__(256) rn_head_kernel(const float* __restrict__ pooled, const float* __restrict__ s5,
                                                      const float* __restrict__ t5, const float* __restrict__ w6,
                                                      const float* __restrict__ b6, float* __restrict__ logits) {
  __shared__ double red[8];
  const int b = blockIdx.x;
  double acc = 0.0;
  for (int k = threadIdx.x; k < 2 * CA; k += 256) acc += (double)(fmaf(pooled[(size_t)b * 2 * CA + k], s5[k], t5[k]) * w6[k]);
  acc = block_sum_d(acc, red);
  if (threadIdx.x == 0) logits[b] = (float)acc + b6[0];
}
// d loss / d logit -> d / d mu, d / d m2 and the softmax-backward inner product
__global__ void rn_head_bwd_kernel(const float* __restrict__ logits, const long long* __restrict__ y, const float* __restrict__ coef,
                                   int mode, float inv_n, const float* __restrict__ s5, const float* __restrict__ w6,
                                   const float* __restrict__ pooled, const float* __restrict__ vq, const float* __restrict__ m2s,
                                   float* __restrict__ gmu, float* __restrict__ gm2, float* __restrict__ dot, int B) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * CA) return;
  const int b = i / CA, c = i - b * CA;
  float go = 1.0f;
  if (mode == 2) go = coef[b];
  if (mode == 0) {
    const float o = logits[b];
    const float p1 = 1.0f / (1.0f + expf(-2.0f * o));
    go = 2.0f * (p1 - (float)y[b]) * inv_n;
  }
  const float mu = pooled[(size_t)b * 2 * CA + c], sg = pooled[(size_t)b * 2 * CA + CA + c];
  const float g_mu = go * w6[c] * s5[c], g_sg = go * w6[CA + c] * s5[CA + c];
  const float v = vq[i];
  const float g_v = (v >= 1e-4f && v <= 1e4f) ? g_sg / (2.f * sg) : 0.f;
  const float a = g_mu - 2.f * mu * g_v;
  gmu[i] = a;
  gm2[i] = g_v;
  dot[i] = mu * a + m2s[i] * g_v;
}
// GE = w (H gmu + H^2 gm2 - dot), w = softmax_t(E)
__global__ void rn_asp_ge_kernel(const float* __restrict__ H, const float* __restrict__ E, const float* __restrict__ emax,
                                 const float* __restrict__ Zs, const float* __restrict__ gmu, const float* __restrict__ gm2,
                                 const float* __restrict__ dot, float* __restrict__ GE, int B, int T3) {
  const size_t n = (size_t)B * T3 * CA;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const size_t bt = i / CA;
    const int c = (int)(i - bt * CA), b = (int)(bt / T3);
    const size_t o = (size_t)b * CA + c;
    const float w = expf(E[i] - emax[o]) / Zs[o], h = H[i];
    GE[i] = w * (h * gmu[o] + h * h * gm2[o] - dot[o]);
  }
}
// gasum[b][j] = sum_t GA[b T3 + t][j]   (128 columns)
__global__ void __launch_bounds__(1024) rn_colsum_kernel(const float* __restrict__ G, int ld, int T3, float* __restrict__ out, int N) {
  __shared__ float red[32][33];
  const int c = blockIdx.x * 32 + threadIdx.x, b = blockIdx.y;
  float acc = 0.f;
  for (int t = threadIdx.y; t < T3; t += 32) acc += G[((size_t)b * T3 + t) * ld + c];
  const float s = colsum32(acc, red);
  if (threadIdx.y == 0) out[(size_t)b * N + c] = s;
}
// GH = (G1 + w (gmu + 2 H gm2) + g_mean / T + [var in range] g_sd / sd * (H - mean) / (T - 1)) * [H > 0]   (in place in G1)
__global__ void rn_asp_combine_kernel(const float* __restrict__ H, const float* __restrict__ E, const float* __restrict__ emax,
                                      const float* __restrict__ Zs, const float* __restrict__ gmu, const float* __restrict__ gm2,
                                      const float* __restrict__ stats, const float* __restrict__ var, const float* __restrict__ gstat,
                                      float* __restrict__ G1, int B, int T3) {
  const size_t n = (size_t)B * T3 * CA;
  const float invT = 1.f / (float)T3, invT1 = 1.f / (float)(T3 - 1);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const size_t bt = i / CA;
    const int c = (int)(i - bt * CA), b = (int)(bt / T3);
    const size_t o = (size_t)b * CA + c;
    const float h = H[i];
    float g = 0.f;
    if (h > 0.f) {
      const float w = expf(E[i] - emax[o]) / Zs[o];
      const float mean = stats[(size_t)b * 2 * CA + c], sd = stats[(size_t)b * 2 * CA + CA + c];
      const float v = var[o];
      const float g_sd = gstat[(size_t)b * 2 * CA + CA + c];
      const float cvar = (v >= 1e-4f && v <= 1e4f) ? g_sd / sd * invT1 : 0.f;
      g = G1[i] + w * (gmu[o] + 2.f * h * gm2[o]) + gstat[(size_t)b * 2 * CA + c] * invT + cvar * (h - mean);
    }
    G1[i] = g;
  }
}

inline int ew_grid(size_t n, int block = 256) { return (int)std::min<size_t>((n + block - 1) / block, (size_t)148 * 16); }

// ---- GEMM descriptors ----------------------------------------------------------------------------------------------
GemmW wview(const float* w, long long s_n, long long s_k, long long s_tap, int n_valid, int k_valid) {
  GemmW v;
  v.w = w, v.s_n = s_n, v.s_k = s_k, v.s_tap = s_tap, v.n_valid = n_valid, v.k_valid = k_valid;
  return v;
}
struct WSpec {
  GemmW v;
  int N, K, ntap;
};
// forward / backward weight views of every contraction
WSpec spec_sinc_f(const RnModel& m) { return {wview(m.filt, SK, 1, 0, NS, SK), NS, 256, 1}; }
WSpec spec_sinc_b(const RnModel& m) { return {wview(m.filt, 1, SK, 0, SK, NS), 256, NS, 1}; }
WSpec spec_1x1_f(const float* w, int Cout, int Cin, int ldw) { return {wview(w, ldw, 1, 0, Cout, Cin), Cout, Cin, 1}; }
WSpec spec_1x1_b(const float* w, int Cout, int Cin, int ldw) { return {wview(w, 1, ldw, 0, Cin, Cout), Cin, Cout, 1}; }
WSpec spec_k3_f(const float* w) { return {wview(w, W * 3, 3, 1, W, W), W, W, 3}; }
WSpec spec_k3_b(const float* w) { return {wview(w, 3, W * 3, 1, W, W), W, W, 3}; }

int alloc_pack(const RnAllocFn& alloc, unsigned char** p, int N, int K, int ntap) {
  return alloc(reinterpret_cast<void**>(p), gemm_pack_bytes(N, K, ntap));
}
template <typename Tp>
int alloc_n(const RnAllocFn& alloc, Tp** p, size_t count) {
  return alloc(reinterpret_cast<void**>(p), count * sizeof(Tp));
}

GemmArgs base_args(const float* A, int lda, int M, const WSpec& s, const unsigned char* wpack, int Tp, int pad, int Tv,
                   const char* tag) {
  GemmArgs a;
  a.A = A, a.lda = lda, a.M = M, a.K = s.K, a.ntap = s.ntap, a.w = s.v, a.wpack = wpack, a.N = s.N;
  a.Tp = Tp, a.pad = pad, a.Tv = Tv, a.tag = tag;
  return a;
}

const char* kLayerNames[3] = {"layer1", "layer2", "layer3"};

}  // namespace

int rn_check_tensors(const std::function<int(const std::string&, long long)>& require) {
  ADVB_TRY(require("preprocess.1.weight", 1));
  ADVB_TRY(require("preprocess.1.bias", 1));
  ADVB_TRY(require("conv1.filterbank.low_hz_", 128));
  ADVB_TRY(require("conv1.filterbank.band_hz_", 128));
  ADVB_TRY(require("conv1.filterbank.window_", 125));
  ADVB_TRY(require("conv1.filterbank.n_", 125));
  auto bn = [&](const std::string& p, int n) {
    for (const char* q : {".weight", ".bias", ".running_mean", ".running_var"}) ADVB_TRY(require(p + q, n));
    return 0;
  };
  for (int l = 0; l < 3; ++l) {
    const std::string p = kLayerNames[l];
    const int cin = l == 0 ? NS : C;
    ADVB_TRY(require(p + ".conv1.weight", (long long)C * cin));
    ADVB_TRY(require(p + ".conv1.bias", C));
    ADVB_TRY(bn(p + ".bn1", C));
    for (int i = 0; i < 7; ++i) {
      ADVB_TRY(require(p + ".convs." + std::to_string(i) + ".weight", W * W * 3));
      ADVB_TRY(require(p + ".convs." + std::to_string(i) + ".bias", W));
      ADVB_TRY(bn(p + ".bns." + std::to_string(i), W));
    }
    ADVB_TRY(require(p + ".conv3.weight", (long long)C * C));
    ADVB_TRY(require(p + ".conv3.bias", C));
    ADVB_TRY(bn(p + ".bn3", C));
    ADVB_TRY(require(p + ".afms.alpha", C));
    ADVB_TRY(require(p + ".afms.fc.weight", (long long)C * C));
    ADVB_TRY(require(p + ".afms.fc.bias", C));
    if (l == 0) ADVB_TRY(require(p + ".residual.0.weight", (long long)C * NS));
  }
  ADVB_TRY(require("layer4.weight", (long long)CA * 3 * C));
  ADVB_TRY(require("layer4.bias", CA));
  ADVB_TRY(require("attention.0.weight", (long long)W * 3 * CA));
  ADVB_TRY(require("attention.0.bias", W));
  ADVB_TRY(bn("attention.2", W));
  ADVB_TRY(require("attention.3.weight", (long long)CA * W));
  ADVB_TRY(require("attention.3.bias", CA));
  ADVB_TRY(bn("bn5", 2 * CA));
  ADVB_TRY(require("fc6.weight", 2 * CA));
  ADVB_TRY(require("fc6.bias", 1));
  return 0;
}

void rn_bind(RnModel& m, const RnLookupFn& t) {
  auto bn = [&](const std::string& p, const float* (&dst)[4]) {
    dst[0] = t(p + ".weight"), dst[1] = t(p + ".bias"), dst[2] = t(p + ".running_mean"), dst[3] = t(p + ".running_var");
  };
  m.in_w = t("preprocess.1.weight"), m.in_b = t("preprocess.1.bias");
  m.low_hz = t("conv1.filterbank.low_hz_"), m.band_hz = t("conv1.filterbank.band_hz_");
  m.window = t("conv1.filterbank.window_"), m.n_axis = t("conv1.filterbank.n_");
  for (int l = 0; l < 3; ++l) {
    RnLayer& k = m.layer[l];
    const std::string p = kLayerNames[l];
    k.w1 = t(p + ".conv1.weight"), k.b1 = t(p + ".conv1.bias");
    k.w3 = t(p + ".conv3.weight"), k.b3 = t(p + ".conv3.bias");
    k.wres = l == 0 ? t(p + ".residual.0.weight") : nullptr;
    bn(p + ".bn1", k.bn1);
    bn(p + ".bn3", k.bn3);
    for (int i = 0; i < 7; ++i) {
      k.wc[i] = t(p + ".convs." + std::to_string(i) + ".weight"), k.bc[i] = t(p + ".convs." + std::to_string(i) + ".bias");
      bn(p + ".bns." + std::to_string(i), k.bns[i]);
    }
    k.alpha = t(p + ".afms.alpha"), k.afc_w = t(p + ".afms.fc.weight"), k.afc_b = t(p + ".afms.fc.bias");
  }
  m.w4 = t("layer4.weight"), m.b4 = t("layer4.bias");
  m.wa = t("attention.0.weight"), m.ba = t("attention.0.bias");
  m.wb = t("attention.3.weight"), m.bb = t("attention.3.bias");
  bn("attention.2", m.abn);
  bn("bn5", m.bn5);
  m.w6 = t("fc6.weight"), m.b6 = t("fc6.bias");
}

int rn_build(RnModel& m, int Bmax, int T, const RnAllocFn& alloc) {
  ADVB_CHECK(T >= 4000, "clip too short for RawNet3's pooling stack");
  m.Bmax = Bmax, m.T = T;
  m.L0 = (T - SK) / 10 + 1;
  const size_t B = (size_t)Bmax;
  const int dil[3] = {2, 3, 4}, pool[3] = {5, 3, 1};
  int Tin = m.L0;
  for (int l = 0; l < 3; ++l) {
    RnLayer& k = m.layer[l];
    k.name = kLayerNames[l];
    k.Cin = l == 0 ? NS : C, k.d = dil[l], k.pool = pool[l];
    k.T = Tin, k.Tp = Tin + 2 * k.d, k.To = Tin / k.pool;
    if (l == 0) Tin = k.To;                  // layer2 consumes x1
    if (l == 1) Tin = k.To;                  // layer3 consumes mp3(x1) + x2, same length as x2
    const size_t R = B * k.Tp;
    ADVB_TRY(alloc_n(alloc, &k.bn1_s, C));
    ADVB_TRY(alloc_n(alloc, &k.bn1_t, C));
    ADVB_TRY(alloc_n(alloc, &k.bn3_s, C));
    ADVB_TRY(alloc_n(alloc, &k.bn3_t, C));
    ADVB_TRY(alloc_n(alloc, &k.bns_s, 7 * W));
    ADVB_TRY(alloc_n(alloc, &k.bns_t, 7 * W));
    ADVB_TRY(alloc_pack(alloc, &k.p_w1, C, k.Cin, 1));
    ADVB_TRY(alloc_pack(alloc, &k.q_w1, k.Cin, C, 1));
    ADVB_TRY(alloc_pack(alloc, &k.p_w3, C, C, 1));
    ADVB_TRY(alloc_pack(alloc, &k.q_w3, C, C, 1));
    if (l == 0) {
      ADVB_TRY(alloc_pack(alloc, &k.p_res, C, k.Cin, 1));
      ADVB_TRY(alloc_pack(alloc, &k.q_res, k.Cin, C, 1));
      ADVB_TRY(alloc_n(alloc, &k.res, R * C));
    }
    for (int i = 0; i < 7; ++i) {
      ADVB_TRY(alloc_pack(alloc, &k.p_c[i], W, W, 3));
      ADVB_TRY(alloc_pack(alloc, &k.q_c[i], W, W, 3));
    }
    ADVB_TRY(alloc_n(alloc, &k.xin, R * k.Cin));
    ADVB_TRY(alloc_n(alloc, &k.o1, R * C));
    ADVB_TRY(alloc_n(alloc, &k.cat, R * C));
    ADVB_TRY(alloc_n(alloc, &k.y, R * C));
    for (int j = 0; j < 2; ++j) {
      ADVB_TRY(alloc_n(alloc, &k.spin[j], R * W));
      ADVB_TRY(alloc_n(alloc, &k.gpre[j], R * W));
    }
    ADVB_TRY(alloc_n(alloc, &k.m1, R * C));
    ADVB_TRY(alloc_n(alloc, &k.mc, R * C));
    ADVB_TRY(alloc_n(alloc, &k.m3, R * C));
    ADVB_TRY(alloc_n(alloc, &k.P, B * k.To * C));
    ADVB_TRY(alloc_n(alloc, &k.arg, B * k.To * C));
    ADVB_TRY(alloc_n(alloc, &k.pm, B * C));
    ADVB_TRY(alloc_n(alloc, &k.yv, B * C));
    ADVB_TRY(alloc_n(alloc, &k.gyv, B * C));
    ADVB_TRY(alloc_n(alloc, &k.gz, B * C));
    ADVB_TRY(alloc_n(alloc, &k.gm, B * C));
  }
  ADVB_CHECK(m.layer[1].To == m.layer[2].T && m.layer[0].To / 3 == m.layer[1].To, "inconsistent RawNet3 time axes");
  m.T3 = m.layer[2].T;
  ADVB_CHECK(m.T3 >= 2, "clip too short for attentive statistics pooling");
  const size_t R3 = B * m.T3, R0 = B * m.L0, R1 = B * m.layer[0].Tp;
  ADVB_TRY(alloc_n(alloc, &m.filt, (size_t)NS * SK));
  ADVB_TRY(alloc_n(alloc, &m.abn_s, W));
  ADVB_TRY(alloc_n(alloc, &m.abn_t, W));
  ADVB_TRY(alloc_n(alloc, &m.bn5_s, 2 * CA));
  ADVB_TRY(alloc_n(alloc, &m.bn5_t, 2 * CA));
  ADVB_TRY(alloc_pack(alloc, &m.p_sinc, NS, 256, 1));
  ADVB_TRY(alloc_pack(alloc, &m.q_sinc, 256, NS, 1));
  ADVB_TRY(alloc_pack(alloc, &m.p_w4, CA, 3 * C, 1));
  ADVB_TRY(alloc_pack(alloc, &m.q_w4, 3 * C, CA, 1));
  ADVB_TRY(alloc_pack(alloc, &m.p_wa, W, CA, 1));
  ADVB_TRY(alloc_pack(alloc, &m.q_wa, CA, W, 1));
  ADVB_TRY(alloc_pack(alloc, &m.p_wb, CA, W, 1));
  ADVB_TRY(alloc_pack(alloc, &m.q_wb, W, CA, 1));
  ADVB_TRY(alloc_n(alloc, &m.pre_stats, 2 * B));
  ADVB_TRY(alloc_n(alloc, &m.nsig, B * T));
  ADVB_TRY(alloc_n(alloc, &m.S, R0 * NS));
  ADVB_TRY(alloc_n(alloc, &m.cat4, R3 * 3 * C));
  ADVB_TRY(alloc_n(alloc, &m.arg1, R3 * C));
  ADVB_TRY(alloc_n(alloc, &m.H, R3 * CA));
  ADVB_TRY(alloc_n(alloc, &m.stats, B * 2 * CA));
  ADVB_TRY(alloc_n(alloc, &m.var, B * CA));
  ADVB_TRY(alloc_n(alloc, &m.cb, B * W));
  ADVB_TRY(alloc_n(alloc, &m.A1, R3 * W));
  ADVB_TRY(alloc_n(alloc, &m.ma, R3 * W));
  ADVB_TRY(alloc_n(alloc, &m.E, R3 * CA));
  ADVB_TRY(alloc_n(alloc, &m.emax, B * CA));
  ADVB_TRY(alloc_n(alloc, &m.Z, B * CA));
  ADVB_TRY(alloc_n(alloc, &m.vq, B * CA));
  ADVB_TRY(alloc_n(alloc, &m.m2, B * CA));
  ADVB_TRY(alloc_n(alloc, &m.pooled, B * 2 * CA));
  ADVB_TRY(alloc_n(alloc, &m.gmu, B * CA));
  ADVB_TRY(alloc_n(alloc, &m.gm2, B * CA));
  ADVB_TRY(alloc_n(alloc, &m.dot, B * CA));
  ADVB_TRY(alloc_n(alloc, &m.GE, R3 * CA));
  ADVB_TRY(alloc_n(alloc, &m.G1, R3 * CA));
  ADVB_TRY(alloc_n(alloc, &m.GA, R3 * W));
  ADVB_TRY(alloc_n(alloc, &m.gasum, B * W));
  ADVB_TRY(alloc_n(alloc, &m.gstat, B * 2 * CA));
  ADVB_TRY(alloc_n(alloc, &m.gcat4, R3 * 3 * C));
  ADVB_TRY(alloc_n(alloc, &m.GY, R1 * C));
  ADVB_TRY(alloc_n(alloc, &m.GC3, R1 * C));
  ADVB_TRY(alloc_n(alloc, &m.GCAT, R1 * C));
  ADVB_TRY(alloc_n(alloc, &m.GC1, R1 * C));
  ADVB_TRY(alloc_n(alloc, &m.GX, std::max(R1 * NS, B * m.layer[1].Tp * C)));
  ADVB_TRY(alloc_n(alloc, &m.GX1, B * m.layer[0].To * C));
  ADVB_TRY(alloc_n(alloc, &m.GS, R0 * NS));
  ADVB_TRY(alloc_n(alloc, &m.Zc, R0 * NS));
  ADVB_TRY(alloc_n(alloc, &m.gn, B * T));
  ADVB_TRY(alloc_n(alloc, &m.bst, 2 * B));
  return 0;
}

int rn_prepare(RnModel& m, int path, cudaStream_t st) {
  auto fold = [&](const float* const (&bn)[4], float* s, float* t, int n) {
    rn_fold_bn_kernel<<<cdiv(n, 256), 256, 0, st>>>(bn[0], bn[1], bn[2], bn[3], s, t, n);
    ADVB_KERNEL_OK("rn_fold_bn", st);
    return 0;
  };
  auto pack = [&](const WSpec& s, unsigned char* dst) { return path == 0 ? gemm_pack(s.v, s.N, s.K, s.ntap, dst, st) : 0; };
  rn_sinc_filters_kernel<<<NS, 256, 0, st>>>(m.low_hz, m.band_hz, m.window, m.n_axis, m.filt);
  ADVB_KERNEL_OK("rn_sinc_filters", st);
  ADVB_TRY(pack(spec_sinc_f(m), m.p_sinc));
  ADVB_TRY(pack(spec_sinc_b(m), m.q_sinc));
  for (int l = 0; l < 3; ++l) {
    RnLayer& k = m.layer[l];
    ADVB_TRY(fold(k.bn1, k.bn1_s, k.bn1_t, C));
    ADVB_TRY(fold(k.bn3, k.bn3_s, k.bn3_t, C));
    for (int i = 0; i < 7; ++i) ADVB_TRY(fold(k.bns[i], k.bns_s + i * W, k.bns_t + i * W, W));
    ADVB_TRY(pack(spec_1x1_f(k.w1, C, k.Cin, k.Cin), k.p_w1));
    ADVB_TRY(pack(spec_1x1_b(k.w1, C, k.Cin, k.Cin), k.q_w1));
    ADVB_TRY(pack(spec_1x1_f(k.w3, C, C, C), k.p_w3));
    ADVB_TRY(pack(spec_1x1_b(k.w3, C, C, C), k.q_w3));
    if (k.wres != nullptr) {
      ADVB_TRY(pack(spec_1x1_f(k.wres, C, k.Cin, k.Cin), k.p_res));
      ADVB_TRY(pack(spec_1x1_b(k.wres, C, k.Cin, k.Cin), k.q_res));
    }
    for (int i = 0; i < 7; ++i) {
      ADVB_TRY(pack(spec_k3_f(k.wc[i]), k.p_c[i]));
      ADVB_TRY(pack(spec_k3_b(k.wc[i]), k.q_c[i]));
    }
  }
  ADVB_TRY(fold(m.abn, m.abn_s, m.abn_t, W));
  ADVB_TRY(fold(m.bn5, m.bn5_s, m.bn5_t, 2 * CA));
  ADVB_TRY(pack(spec_1x1_f(m.w4, CA, 3 * C, 3 * C), m.p_w4));
  ADVB_TRY(pack(spec_1x1_b(m.w4, CA, 3 * C, 3 * C), m.q_w4));
  ADVB_TRY(pack(spec_1x1_f(m.wa, W, CA, 3 * CA), m.p_wa));   // only the first 1536 input channels vary over time
  ADVB_TRY(pack(spec_1x1_b(m.wa, W, CA, 3 * CA), m.q_wa));
  ADVB_TRY(pack(spec_1x1_f(m.wb, CA, W, W), m.p_wb));
  ADVB_TRY(pack(spec_1x1_b(m.wb, CA, W, W), m.q_wb));
  return 0;
}

namespace {

// Bottle2neck forward (rawnet3.py:242-274) + max-pool + AFMS (:176-182); writes the layer output to `out`
// (row r = b * Tpo + pado + t', leading dimension ldo).
int layer_forward(RnModel& m, RnLayer& k, int B, float* out, int ldo, int Tpo, int pado, int path, int passes,
                  cudaStream_t st) {
  const int M = B * k.Tp;
  const float* residual = k.xin;
  if (k.wres != nullptr) {
    GemmArgs a = base_args(k.xin, k.Cin, M, spec_1x1_f(k.wres, C, k.Cin, k.Cin), k.p_res, k.Tp, k.d, k.T, "rn_res_fwd");
    a.out = k.res, a.ldc = C;
    ADVB_TRY(gemm_run(a, path, passes, st));
    residual = k.res;
  }
  {
    GemmArgs a = base_args(k.xin, k.Cin, M, spec_1x1_f(k.w1, C, k.Cin, k.Cin), k.p_w1, k.Tp, k.d, k.T, "rn_conv1_fwd");
    a.bias = k.b1, a.relu = 1, a.mask_out = k.m1, a.ld_mask = C, a.bn_scale = k.bn1_s, a.bn_shift = k.bn1_t;
    a.out = k.o1, a.ldc = C;
    ADVB_TRY(gemm_run(a, path, passes, st));
  }
  for (int i = 0; i < 7; ++i) {
    const float* A = i == 0 ? k.o1 : k.spin[i & 1];
    GemmArgs a = base_args(A, i == 0 ? C : W, M, spec_k3_f(k.wc[i]), k.p_c[i], k.Tp, k.d, k.T, "rn_res2_fwd");
    a.shift[0] = -k.d, a.shift[1] = 0, a.shift[2] = k.d;
    a.bias = k.bc[i], a.relu = 1, a.mask_out = k.mc + i * W, a.ld_mask = C;
    a.bn_scale = k.bns_s + i * W, a.bn_shift = k.bns_t + i * W;
    a.out = k.cat + i * W, a.ldc = C;
    if (i < 6) a.out2 = k.spin[(i + 1) & 1], a.ld2 = W, a.add2 = k.o1 + (i + 1) * W, a.ld_add2 = C;
    ADVB_TRY(gemm_run(a, path, passes, st));
  }
  rn_copy_cols_kernel<<<ew_grid((size_t)M * W / 4), 256, 0, st>>>(k.o1 + 7 * W, C, k.cat + 7 * W, C, (size_t)M, W);
  ADVB_KERNEL_OK("rn_copy_cols", st);
  {
    GemmArgs a = base_args(k.cat, C, M, spec_1x1_f(k.w3, C, C, C), k.p_w3, k.Tp, k.d, k.T, "rn_conv3_fwd");
    a.bias = k.b3, a.relu = 1, a.mask_out = k.m3, a.ld_mask = C, a.bn_scale = k.bn3_s, a.bn_shift = k.bn3_t;
    a.add = residual, a.ld_add = C, a.out = k.y, a.ldc = C;
    ADVB_TRY(gemm_run(a, path, passes, st));
  }
  rn_pool_mean_kernel<<<dim3(C / 32, B), dim3(32, 32), 0, st>>>(k.y, k.Tp, k.d, k.pool, k.To, k.P, k.arg, k.pm);
  ADVB_KERNEL_OK("rn_pool_mean", st);
  rn_linear_kernel<<<C / 8, 256, 0, st>>>(k.afc_w, C, k.afc_b, k.pm, k.yv, B, C, C, 1);
  ADVB_KERNEL_OK("rn_afms_fc", st);
  rn_afms_scale_kernel<<<ew_grid((size_t)B * k.To * C / 4), 256, 0, st>>>(k.P, k.alpha, k.yv, out, ldo, Tpo, pado, B, k.To);
  ADVB_KERNEL_OK("rn_afms_scale", st);
  return 0;
}

// Backward of layer_forward: gout rows r = b * To + t' (leading dimension ldg) -> m.GX [B Tp][Cin] (valid rows).
int layer_backward(RnModel& m, RnLayer& k, int B, const float* gout, int ldg, int path, int passes, cudaStream_t st) {
  const int M = B * k.Tp;
  rn_afms_bwd_sum_kernel<<<dim3(C / 32, B), dim3(32, 32), 0, st>>>(gout, ldg, k.P, k.alpha, k.yv, k.To, k.gz);
  ADVB_KERNEL_OK("rn_afms_bwd_sum", st);
  rn_linear_t_kernel<<<dim3(C / 64, B), dim3(64, 16), 0, st>>>(k.afc_w, C, k.gz, k.gm, C, C, 1.f / (float)k.To);
  ADVB_KERNEL_OK("rn_afms_fc_bwd", st);
  rn_pool_bwd_kernel<<<ew_grid((size_t)B * k.T * C / 4), 256, 0, st>>>(gout, ldg, k.yv, k.gm, k.arg, k.bn3_s, k.m3, m.GY, m.GC3,
                                                                       B, k.T, k.Tp, k.d, k.pool, k.To);
  ADVB_KERNEL_OK("rn_pool_bwd", st);
  {
    GemmArgs a = base_args(m.GC3, C, M, spec_1x1_b(k.w3, C, C, C), k.q_w3, k.Tp, k.d, k.T, "rn_conv3_bwd");
    a.out = m.GCAT, a.ldc = C;
    ADVB_TRY(gemm_run(a, path, passes, st));
  }
  // the untouched last split goes straight to conv1's output gradient; split 6 starts the Res2 chain
  rn_gate_kernel<<<ew_grid((size_t)B * k.T * W / 4), 256, 0, st>>>(m.GCAT + 7 * W, C, k.bn1_s + 7 * W, k.m1 + 7 * W, C,
                                                                   m.GC1 + 7 * W, C, B, k.T, k.Tp, k.d, W);
  ADVB_KERNEL_OK("rn_gate", st);
  rn_gate_kernel<<<ew_grid((size_t)B * k.T * W / 4), 256, 0, st>>>(m.GCAT + 6 * W, C, k.bns_s + 6 * W, k.mc + 6 * W, C, k.gpre[0],
                                                                   W, B, k.T, k.Tp, k.d, W);
  ADVB_KERNEL_OK("rn_gate", st);
  for (int i = 6; i >= 0; --i) {
    const int cur = (6 - i) & 1;
    GemmArgs a = base_args(k.gpre[cur], W, M, spec_k3_b(k.wc[i]), k.q_c[i], k.Tp, k.d, k.T, "rn_res2_bwd");
    a.shift[0] = k.d, a.shift[1] = 0, a.shift[2] = -k.d;  // g_in[t] = sum_j W_j^T g_pre[t - (j - 1) d]
    a.out = m.GC1 + i * W, a.ldc = C, a.gate_scale = k.bn1_s + i * W, a.gate_mask = k.m1 + i * W, a.ld_gate = C;
    if (i > 0) {
      a.out2 = k.gpre[cur ^ 1], a.ld2 = W, a.add2 = m.GCAT + (i - 1) * W, a.ld_add2 = C;
      a.gate2_scale = k.bns_s + (i - 1) * W, a.gate2_mask = k.mc + (i - 1) * W, a.ld_gate2 = C;
    }
    ADVB_TRY(gemm_run(a, path, passes, st));
  }
  if (k.wres != nullptr) {
    GemmArgs a = base_args(m.GY, C, M, spec_1x1_b(k.wres, C, k.Cin, k.Cin), k.q_res, k.Tp, k.d, k.T, "rn_res_bwd");
    a.out = m.GX, a.ldc = k.Cin;
    ADVB_TRY(gemm_run(a, path, passes, st));
  }
  {
    GemmArgs a = base_args(m.GC1, C, M, spec_1x1_b(k.w1, C, k.Cin, k.Cin), k.q_w1, k.Tp, k.d, k.T, "rn_conv1_bwd");
    a.add = k.wres != nullptr ? m.GX : m.GY, a.ld_add = k.wres != nullptr ? k.Cin : C;
    a.out = m.GX, a.ldc = k.Cin;
    ADVB_TRY(gemm_run(a, path, passes, st));
  }
  return 0;
}

}  // namespace

int rn_forward(RnModel& m, const float* x, float* logits, int B, int path, int passes, cudaStream_t st) {
  const int T = m.T, L0 = m.L0, T3 = m.T3;
  RnLayer &k1 = m.layer[0], &k2 = m.layer[1], &k3 = m.layer[2];
  rn_pre_stats_kernel<<<B, 1024, 0, st>>>(x, T, m.pre_stats);
  ADVB_KERNEL_OK("rn_pre_stats", st);
  rn_pre_apply_kernel<<<ew_grid((size_t)B * T), 256, 0, st>>>(x, m.pre_stats, m.in_w, m.in_b, m.nsig, B, T);
  ADVB_KERNEL_OK("rn_pre_apply", st);
  {
    GemmArgs a = base_args(m.nsig, 0, B * L0, spec_sinc_f(m), m.p_sinc, L0, 0, L0, "rn_sinc_fwd");
    a.im2col_T = T, a.out = m.S, a.ldc = NS;
    ADVB_TRY(gemm_run(a, path, passes, st));
  }
  rn_sinc_post_kernel<<<dim3(NS / 32, B), dim3(32, 32), 0, st>>>(m.S, k1.xin, L0, k1.Tp, k1.d);
  ADVB_KERNEL_OK("rn_sinc_post", st);
  ADVB_TRY(layer_forward(m, k1, B, k2.xin, C, k2.Tp, k2.d, path, passes, st));
  rn_pool3_kernel<<<ew_grid((size_t)B * T3 * C / 4), 256, 0, st>>>(k2.xin, k2.Tp, k2.d, m.cat4, m.arg1, B, T3);
  ADVB_KERNEL_OK("rn_pool3", st);
  ADVB_TRY(layer_forward(m, k2, B, m.cat4 + C, 3 * C, T3, 0, path, passes, st));
  rn_add12_kernel<<<ew_grid((size_t)B * T3 * C / 4), 256, 0, st>>>(m.cat4, k3.xin, k3.Tp, k3.d, B, T3);
  ADVB_KERNEL_OK("rn_add12", st);
  ADVB_TRY(layer_forward(m, k3, B, m.cat4 + 2 * C, 3 * C, T3, 0, path, passes, st));
  const int M3 = B * T3;
  {
    GemmArgs a = base_args(m.cat4, 3 * C, M3, spec_1x1_f(m.w4, CA, 3 * C, 3 * C), m.p_w4, T3, 0, T3, "rn_layer4_fwd");
    a.bias = m.b4, a.relu = 1, a.out = m.H, a.ldc = CA;
    ADVB_TRY(gemm_run(a, path, passes, st));
  }
  rn_asp_stats_kernel<<<dim3(CA / 32, B), dim3(32, 32), 0, st>>>(m.H, T3, m.stats, m.var);
  ADVB_KERNEL_OK("rn_asp_stats", st);
  rn_linear_kernel<<<W / 8, 256, 0, st>>>(m.wa + CA, 3 * CA, m.ba, m.stats, m.cb, B, W, 2 * CA, 0);
  ADVB_KERNEL_OK("rn_att_bias", st);
  {
    GemmArgs a = base_args(m.H, CA, M3, spec_1x1_f(m.wa, W, CA, 3 * CA), m.p_wa, T3, 0, T3, "rn_att1_fwd");
    a.bias = m.cb, a.bias_per_clip = 1, a.relu = 1, a.mask_out = m.ma, a.ld_mask = W;
    a.bn_scale = m.abn_s, a.bn_shift = m.abn_t, a.out = m.A1, a.ldc = W;
    ADVB_TRY(gemm_run(a, path, passes, st));
  }
  {
    GemmArgs a = base_args(m.A1, W, M3, spec_1x1_f(m.wb, CA, W, W), m.p_wb, T3, 0, T3, "rn_att2_fwd");
    a.bias = m.bb, a.out = m.E, a.ldc = CA;
    ADVB_TRY(gemm_run(a, path, passes, st));
  }
  rn_asp_pool_kernel<<<dim3(CA / 32, B), dim3(32, 32), 0, st>>>(m.H, m.E, T3, m.emax, m.Z, m.vq, m.m2, m.pooled);
  ADVB_KERNEL_OK("rn_asp_pool", st);
  rn_head_kernel<<<B, 256, 0, st>>>(m.pooled, m.bn5_s, m.bn5_t, m.w6, m.b6, logits);
  ADVB_KERNEL_OK("rn_head", st);
  return 0;
}

int rn_backward(RnModel& m, const float* x, const float* logits, const long long* y, int B, int mode, int n_global,
                const float* coef, float* gx, int path, int passes, cudaStream_t st) {
  const int T = m.T, L0 = m.L0, T3 = m.T3, M3 = B * T3;
  RnLayer &k1 = m.layer[0], &k2 = m.layer[1], &k3 = m.layer[2];
  rn_head_bwd_kernel<<<cdiv(B * CA, 256), 256, 0, st>>>(logits, y, coef, mode, 1.0f / (float)n_global, m.bn5_s, m.w6, m.pooled, m.vq,
                                                        m.m2, m.gmu, m.gm2, m.dot, B);
  ADVB_KERNEL_OK("rn_head_bwd", st);
  rn_asp_ge_kernel<<<ew_grid((size_t)M3 * CA), 256, 0, st>>>(m.H, m.E, m.emax, m.Z, m.gmu, m.gm2, m.dot, m.GE, B, T3);
  ADVB_KERNEL_OK("rn_asp_ge", st);
  {
    GemmArgs a = base_args(m.GE, CA, M3, spec_1x1_b(m.wb, CA, W, W), m.q_wb, T3, 0, T3, "rn_att2_bwd");
    a.out = m.GA, a.ldc = W, a.gate_scale = m.abn_s, a.gate_mask = m.ma, a.ld_gate = W;
    ADVB_TRY(gemm_run(a, path, passes, st));
  }
  rn_colsum_kernel<<<dim3(W / 32, B), dim3(32, 32), 0, st>>>(m.GA, W, T3, m.gasum, W);
  ADVB_KERNEL_OK("rn_colsum", st);
  rn_linear_t_kernel<<<dim3(2 * CA / 64, B), dim3(64, 16), 0, st>>>(m.wa + CA, 3 * CA, m.gasum, m.gstat, W, 2 * CA, 1.f);
  ADVB_KERNEL_OK("rn_att_bias_bwd", st);
  {
    GemmArgs a = base_args(m.GA, W, M3, spec_1x1_b(m.wa, W, CA, 3 * CA), m.q_wa, T3, 0, T3, "rn_att1_bwd");
    a.out = m.G1, a.ldc = CA;
    ADVB_TRY(gemm_run(a, path, passes, st));
  }
  rn_asp_combine_kernel<<<ew_grid((size_t)M3 * CA), 256, 0, st>>>(m.H, m.E, m.emax, m.Z, m.gmu, m.gm2, m.stats, m.var, m.gstat,
                                                                  m.G1, B, T3);
  ADVB_KERNEL_OK("rn_asp_combine", st);
  {
    GemmArgs a = base_args(m.G1, CA, M3, spec_1x1_b(m.w4, CA, 3 * C, 3 * C), m.q_w4, T3, 0, T3, "rn_layer4_bwd");
    a.out = m.gcat4, a.ldc = 3 * C;
    ADVB_TRY(gemm_run(a, path, passes, st));
  }
  ADVB_TRY(layer_backward(m, k3, B, m.gcat4 + 2 * C, 3 * C, path, passes, st));
  rn_add12_bwd_kernel<<<ew_grid((size_t)M3 * C / 4), 256, 0, st>>>(m.GX, k3.Tp, k3.d, m.gcat4, B, T3);
  ADVB_KERNEL_OK("rn_add12_bwd", st);
  ADVB_TRY(layer_backward(m, k2, B, m.gcat4 + C, 3 * C, path, passes, st));
  rn_x1_grad_kernel<<<ew_grid((size_t)B * k2.T * C / 4), 256, 0, st>>>(m.GX, k2.Tp, k2.d, m.gcat4, m.arg1, m.GX1, B, k2.T, T3);
  ADVB_KERNEL_OK("rn_x1_grad", st);
  ADVB_TRY(layer_backward(m, k1, B, m.GX1, C, path, passes, st));
  rn_sinc_post_bwd_kernel<<<dim3(NS / 32, B), dim3(32, 32), 0, st>>>(m.S, m.GX, m.GS, L0, k1.Tp, k1.d);
  ADVB_KERNEL_OK("rn_sinc_post_bwd", st);
  {
    GemmArgs a = base_args(m.GS, NS, B * L0, spec_sinc_b(m), m.q_sinc, L0, 0, L0, "rn_sinc_bwd");
    a.out = m.Zc, a.ldc = NS;
    ADVB_TRY(gemm_run(a, path, passes, st));
  }
  rn_col2im_kernel<<<ew_grid((size_t)B * T), 256, 0, st>>>(m.Zc, m.gn, B, T, L0);
  ADVB_KERNEL_OK("rn_col2im", st);
  rn_pre_bwd_stats_kernel<<<B, 1024, 0, st>>>(x, m.gn, m.pre_stats, T, m.bst);
  ADVB_KERNEL_OK("rn_pre_bwd_stats", st);
  rn_pre_bwd_apply_kernel<<<ew_grid((size_t)B * T), 256, 0, st>>>(x, m.gn, m.pre_stats, m.bst, m.in_w, gx, B, T);
  ADVB_KERNEL_OK("rn_pre_bwd_apply", st);
  return 0;
}

const float* rn_stage(const RnModel& m, const std::string& name, int* rows, int* cols) {
  auto ret = [&](const float* p, int r, int c) {
    *rows = r, *cols = c;
    return p;
  };
  if (name == "rn_pre") return ret(m.nsig, m.T, 1);
  if (name == "rn_filt") return ret(m.filt, 1, NS * SK);  // (same for every clip: caller reads clip 0)
  if (name == "rn_sinc_raw") return ret(m.S, m.L0, NS);
  if (name == "rn_sinc") return ret(m.layer[0].xin, m.layer[0].Tp, NS);
  if (name == "rn_o1") return ret(m.layer[0].o1, m.layer[0].Tp, C);
  if (name == "rn_cat1") return ret(m.layer[0].cat, m.layer[0].Tp, C);
  if (name == "rn_y1") return ret(m.layer[0].y, m.layer[0].Tp, C);
  if (name == "rn_y2") return ret(m.layer[1].y, m.layer[1].Tp, C);
  if (name == "rn_y3") return ret(m.layer[2].y, m.layer[2].Tp, C);
  if (name == "rn_x1") return ret(m.layer[1].xin, m.layer[1].Tp, C);
  if (name == "rn_cat4") return ret(m.cat4, m.T3, 3 * C);
  if (name == "rn_layer4") return ret(m.H, m.T3, CA);
  if (name == "rn_pooled") return ret(m.pooled, 1, 2 * CA);
  if (name == "rn_gcat4") return ret(m.gcat4, m.T3, 3 * C);
  if (name == "rn_gx1") return ret(m.GX1, m.layer[0].To, C);
  if (name == "rn_gsinc") return ret(m.GX, m.layer[0].Tp, NS);
  if (name == "rn_gs") return ret(m.GS, m.L0, NS);
  if (name == "rn_gn") return ret(m.gn, m.T, 1);
  return nullptr;
}

}  // namespace advb
