// SpecRNet forward and input-gradient backward (sm_100a, fp32 SIMT first path).
//
// Replaces src/models/specrnet.py:73-91 (Residual_block2D; its bn1/lrelu output is discarded by the reference, so the
// effective block is conv1(x) -> bn2 -> LeakyReLU(0.3) -> conv2 (+ conv_downsample(x) | + x) -> MaxPool2d(2)),
// :139-181 (first_bn + SELU, three blocks each followed by y = sigmoid(fc(avgpool)), x*y + y, MaxPool2d(2);
// bn_before_gru + SELU, 2-layer bidirectional GRU, last time step, fc1_gru, fc2_gru) and the autograd input gradient.
//
// Layout: [frames][coeffs][channels] fp32 (the transposed cepstral image, as on the LCNN path; conv taps are
// transposed while packing), zero border of 1 pixel where a 3x3 conv consumes the tensor, channels padded to a multiple
// of 8 (20 -> 24) with zero weights so padded channels stay exactly 0.
//
// Per block, forward:   F1  h  = lrelu(bn2(conv1(x)))                                      (B,H,W,C)
//                       F2  xb = maxpool(conv2(h) + identity(x)) + 2-bit arg-max + partial channel sums
//                       ATT y  = sigmoid(fc(mean xb))                                      (B,C)
//                       SP  xn = maxpool(xb * y + y) + 2-bit arg-max
// backward (no conv input is needed; the expanded gradient g_o of the conv2 output is rebuilt on the fly
// from g_xn, the two arg-max codes, y and the attention broadcast term and never stored):
//                       ATT' gadd = fc^T(gy * y(1-y)) / (Hb Wb),  gy = sum g_u (xb + 1)
//                       B2  g_c1 = conv2^T(g_o) * lrelu'(h) * bn2 scale
//                       B1  g_x  = conv1^T(g_c1) + downsample^T(g_o) | + g_o
// Algorithmic HBM bytes per clip: every tensor above written once and read once (bench.py kernel_bytes).
#include "specrnet.cuh"

#include <stdlib.h>

#include <algorithm>

#include "conv.cuh"
#include "conv_core.cuh"
#include "tc_common.cuh"

namespace advb {

namespace {

using namespace convcore;

constexpr float SELU_ALPHA = 1.6732632423543772f;
constexpr float SELU_SCALE = 1.0507009873554805f;
__device__ __forceinline__ float selu_f(float u) { return u > 0.f ? SELU_SCALE * u : SELU_SCALE * SELU_ALPHA * expm1f(u); }
// derivative expressed with the OUTPUT v = selu(u): u > 0 <=> v > 0; for u <= 0, scale*alpha*e^u = v + scale*alpha
__device__ __forceinline__ float selu_grad_from_out(float v) { return v > 0.f ? SELU_SCALE : v + SELU_SCALE * SELU_ALPHA; }

enum { F1 = 0, F2 = 1, B2 = 2, B1 = 3 };

struct SrArgs {
  int B, H, W, CK, N;   // conv grid; contraction channels as staged (padded); output channels (padded)
  int Hb, Wb, Hn, Wn, C;  // pooled grids and block-output channels (padded) for the gradient expansion
  const float* in;      // F1: x (bordered 1), F2: h (bordered 1), B1: g_c1 (compact)
  const float* wpk;     // [tap][CK][N]
  const float* bias;    // (N) padded
  const float* scale;   // bn2 scale / shift (padded)
  const float* shift;
  float* out;
  int out_pad;
  const float* x;       // block input, bordered 1 (identity path of F2)
  int Ci;               // its channels (padded)
  const float* wd;      // F2: downsample [Ci][N] or null;  B1: downsample^T [C][N] or null
  const float* bd;
  unsigned char* code1w;
  float* psum;
  const float* g_xn;
  const unsigned char* code2;
  const unsigned char* code1;
  const float* y;
  const float* gadd;
  const float* h;
  int qpc;              // quads (2x2 output pixels) per CTA
  int CKr;              // contraction channels that are not zero padding (multiple of 4 when CK is), <= CK
  int slab_all;         // 1: all 9 weight slabs are staged once per CTA (small layers), 0: one slab per tap
  const float* c2;      // F2 only: conv2(h) (B, H, W, N) computed by the tensor-core kernel - the contraction here is skipped
};

// 4 consecutive channels (c..c+3) of the gradient at conv2's output pixel (yy, xx): un-pool of
// g_xb = unpool2(g_xn) * y + gadd through the block's own max-pool.
__device__ __forceinline__ float4 expand_go(const SrArgs& a, int b, int yy, int xx, int c) {
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (yy < 0 || yy >= a.H || xx < 0 || xx >= a.W) return v;
  const int cy = yy >> 1, cx = xx >> 1;
  if (cy >= a.Hb || cx >= a.Wb) return v;
  const size_t o1 = (((size_t)b * a.Hb + cy) * a.Wb + cx) * a.C + c;
  const uchar4 c1 = __ldg(reinterpret_cast<const uchar4*>(a.code1 + o1));
  const unsigned pos1 = (unsigned)(((yy & 1) << 1) | (xx & 1));
  if (c1.x != pos1 && c1.y != pos1 && c1.z != pos1 && c1.w != pos1) return v;
  float4 g = __ldg(reinterpret_cast<const float4*>(a.gadd + (size_t)b * a.C + c));
  const int ny = cy >> 1, nx = cx >> 1;
  if (ny < a.Hn && nx < a.Wn) {
    const size_t o2 = (((size_t)b * a.Hn + ny) * a.Wn + nx) * a.C + c;
    const uchar4 c2 = __ldg(reinterpret_cast<const uchar4*>(a.code2 + o2));
    const float4 gn = __ldg(reinterpret_cast<const float4*>(a.g_xn + o2));
    const float4 yv = __ldg(reinterpret_cast<const float4*>(a.y + (size_t)b * a.C + c));
    const unsigned pos2 = (unsigned)(((cy & 1) << 1) | (cx & 1));
    if (c2.x == pos2) g.x += gn.x * yv.x;
    if (c2.y == pos2) g.y += gn.y * yv.y;
    if (c2.z == pos2) g.z += gn.z * yv.z;
    if (c2.w == pos2) g.w += gn.w * yv.w;
  }
  v.x = c1.x == pos1 ? g.x : 0.f;
  v.y = c1.y == pos1 ? g.y : 0.f;
  v.z = c1.z == pos1 ? g.z : 0.f;
  v.w = c1.w == pos1 ? g.w : 0.f;
  return v;
}

template <int MODE>
__global__ void __launch_bounds__(256) sr_conv_kernel(SrArgs a, int band_floats) {
  extern __shared__ __align__(16) float smem[];
  const TileGeom g = tile_geom(a.H, a.W, MODE == F2, 1, a.qpc);
  const int CK = a.CK;
  const int CKp = (CK & 3) == 0 ? CK + 4 : CK;
  float* band = smem;
  float* w_s = smem + band_floats;
  // F2 only: qpc x N partial sums (alone in shared memory when the contraction comes from the tensor-core kernel)
  float* s_part = (MODE == F2 && a.c2 != nullptr) ? smem : w_s + (a.slab_all ? 9 : 1) * CK * a.N;
  const int b = blockIdx.y, tid = threadIdx.x, nt = blockDim.x;

  const bool from_c2 = MODE == F2 && a.c2 != nullptr;
  // ---- stage the band: rows 2*qy0-1 ..., columns -1 .. 2*QW ----
  if (from_c2) {
    // nothing to stage: the accumulators come from the tensor-core convolution
  } else if (MODE == F1 || MODE == F2) {
    const int Hp = a.H + 2, Wp = a.W + 2;
    const float* inb = a.in + (size_t)b * Hp * Wp * CK;
    if ((CK & 3) == 0) {
      const int c4n = a.CKr >> 2, total = g.nrows * g.BW * c4n;  // padded channels are never read
      for (int i = tid; i < total; i += nt) {
        const int c4 = i % c4n, col = (i / c4n) % g.BW, r = i / (c4n * g.BW);
        const int yp = 2 * g.qy0 + r, xp = col;  // (-1 + r) + pad 1
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (yp < Hp && xp < Wp) v = __ldg(reinterpret_cast<const float4*>(inb + ((size_t)yp * Wp + xp) * CK + 4 * c4));
        *reinterpret_cast<float4*>(band + ((size_t)r * g.BW + col) * CKp + 4 * c4) = v;
      }
    } else {
      const int total = g.nrows * g.BW * CK;
      for (int i = tid; i < total; i += nt) {
        const int c = i % CK, col = (i / CK) % g.BW, r = i / (CK * g.BW);
        const int yp = 2 * g.qy0 + r, xp = col;
        float v = 0.f;
        if (yp < Hp && xp < Wp) v = __ldg(inb + ((size_t)yp * Wp + xp) * CK + c);
        band[((size_t)r * g.BW + col) * CKp + c] = v;
      }
    }
  } else {
    const int c4n = a.CKr >> 2, total = g.nrows * g.BW * c4n;  // padded channels are never read
    for (int i = tid; i < total; i += nt) {
      const int c4 = i % c4n, col = (i / c4n) % g.BW, r = i / (c4n * g.BW);
      const int y = 2 * g.qy0 - 1 + r, x = col - 1;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (MODE == B2) {
        v = expand_go(a, b, y, x, 4 * c4);
      } else if (y >= 0 && y < a.H && x >= 0 && x < a.W) {
        v = __ldg(reinterpret_cast<const float4*>(a.in + (((size_t)b * a.H + y) * a.W + x) * CK + 4 * c4));
      }
      *reinterpret_cast<float4*>(band + ((size_t)r * g.BW + col) * CKp + 4 * c4) = v;
    }
  }

  const int N = a.N, G = N >> 3;
  const int cg = tid % G, ql = tid / G;
  const int q = g.q0 + ql;
  const bool valid = q <= g.q1;
  const int qy = q / g.QW, qx = q % g.QW;
  float acc[4][8];
#pragma unroll
  for (int p = 0; p < 4; ++p)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[p][j] = 0.f;
  if (from_c2) {
    if (valid) {
      const int nh2 = N >> 1;
#pragma unroll
      for (int p = 0; p < 4; ++p) {
        const int y = 2 * qy + (p >> 1), x = 2 * qx + (p & 1);
        const float* src = a.c2 + (((size_t)b * a.H + y) * a.W + x) * N;
        const float4 lo = __ldg(reinterpret_cast<const float4*>(src + 4 * cg));
        const float4 hi = __ldg(reinterpret_cast<const float4*>(src + nh2 + 4 * cg));
        acc[p][0] = lo.x, acc[p][1] = lo.y, acc[p][2] = lo.z, acc[p][3] = lo.w;
        acc[p][4] = hi.x, acc[p][5] = hi.y, acc[p][6] = hi.z, acc[p][7] = hi.w;
      }
    }
  } else {
    conv_core<3>(band, g.BW, CK, CKp, a.wpk, w_s, N, 2 * (qy - g.qy0), 2 * qx, 4 * cg, valid, acc, a.CKr, a.slab_all != 0);
  }

  const int nh = N >> 1, c0 = 4 * cg;
  if (MODE == F1) {
    if (!valid) return;
    const int Hp = a.H + 2, Wp = a.W + 2;
#pragma unroll
    for (int p = 0; p < 4; ++p) {
      const int y = 2 * qy + (p >> 1), x = 2 * qx + (p & 1);
      if (y >= a.H || x >= a.W) continue;
      float o[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int c = (j < 4 ? c0 : nh + c0 - 4) + j;
        float v = acc[p][j] + __ldg(a.bias + c);
        v = fmaf(v, __ldg(a.scale + c), __ldg(a.shift + c));
        o[j] = v > 0.f ? v : 0.3f * v;
      }
      float* dst = a.out + (((size_t)b * Hp + y + 1) * Wp + x + 1) * N;
      *reinterpret_cast<float4*>(dst + c0) = make_float4(o[0], o[1], o[2], o[3]);
      *reinterpret_cast<float4*>(dst + nh + c0) = make_float4(o[4], o[5], o[6], o[7]);
    }
  } else if (MODE == F2) {
    float best[8];
    unsigned char code[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      best[j] = 0.f;
      code[j] = 0;
    }
    if (valid) {
      const int Hp = a.H + 2, Wp = a.W + 2;
      float idv[4][8];
#pragma unroll
      for (int p = 0; p < 4; ++p) {
        const int y = 2 * qy + (p >> 1), x = 2 * qx + (p & 1);
        const float* xp = a.x + (((size_t)b * Hp + y + 1) * Wp + x + 1) * a.Ci;
        if (a.wd != nullptr) {
#pragma unroll
          for (int j = 0; j < 8; ++j) idv[p][j] = 0.f;
          for (int ci = 0; ci < a.Ci; ++ci) {
            const float xv = __ldg(xp + ci);
            const float4 wl = __ldg(reinterpret_cast<const float4*>(a.wd + (size_t)ci * N + c0));
            const float4 wh = __ldg(reinterpret_cast<const float4*>(a.wd + (size_t)ci * N + nh + c0));
            fma4(idv[p], 0, xv, wl);
            fma4(idv[p], 4, xv, wh);
          }
#pragma unroll
          for (int j = 0; j < 8; ++j) idv[p][j] += __ldg(a.bd + (j < 4 ? c0 : nh + c0 - 4) + j);
        } else {
          const float4 lo = __ldg(reinterpret_cast<const float4*>(xp + c0));
          const float4 hi = __ldg(reinterpret_cast<const float4*>(xp + nh + c0));
          idv[p][0] = lo.x, idv[p][1] = lo.y, idv[p][2] = lo.z, idv[p][3] = lo.w;
          idv[p][4] = hi.x, idv[p][5] = hi.y, idv[p][6] = hi.z, idv[p][7] = hi.w;
        }
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float bj = __ldg(a.bias + (j < 4 ? c0 : nh + c0 - 4) + j);
        float bv = (acc[0][j] + bj) + idv[0][j];
        unsigned cd = 0;
#pragma unroll
        for (int p = 1; p < 4; ++p) {
          const float v = (acc[p][j] + bj) + idv[p][j];
          if (v > bv) {
            bv = v;
            cd = (unsigned)p;
          }
        }
        best[j] = bv;
        code[j] = (unsigned char)cd;
      }
      const size_t o = (((size_t)b * a.Hb + qy) * a.Wb + qx) * N;
      *reinterpret_cast<float4*>(a.out + o + c0) = make_float4(best[0], best[1], best[2], best[3]);
      *reinterpret_cast<float4*>(a.out + o + nh + c0) = make_float4(best[4], best[5], best[6], best[7]);
      *reinterpret_cast<uchar4*>(a.code1w + o + c0) = make_uchar4(code[0], code[1], code[2], code[3]);
      *reinterpret_cast<uchar4*>(a.code1w + o + nh + c0) = make_uchar4(code[4], code[5], code[6], code[7]);
    }
    // deterministic partial channel sums of this CTA's quads (avgpool of the attention)
#pragma unroll
    for (int j = 0; j < 8; ++j) s_part[ql * N + (j < 4 ? c0 : nh + c0 - 4) + j] = valid ? best[j] : 0.f;
    __syncthreads();
    if (tid < N) {
      float s = 0.f;
      for (int r = 0; r < a.qpc; ++r) s += s_part[r * N + tid];
      a.psum[((size_t)b * gridDim.x + blockIdx.x) * N + tid] = s;
    }
  } else if (MODE == B2) {
    if (!valid) return;
    const int Hp = a.H + 2, Wp = a.W + 2;
#pragma unroll
    for (int p = 0; p < 4; ++p) {
      const int y = 2 * qy + (p >> 1), x = 2 * qx + (p & 1);
      if (y >= a.H || x >= a.W) continue;
      const float* hp = a.h + (((size_t)b * Hp + y + 1) * Wp + x + 1) * N;
      const float4 hl = __ldg(reinterpret_cast<const float4*>(hp + c0));
      const float4 hh = __ldg(reinterpret_cast<const float4*>(hp + nh + c0));
      const float hv[8] = {hl.x, hl.y, hl.z, hl.w, hh.x, hh.y, hh.z, hh.w};
      float o[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int c = (j < 4 ? c0 : nh + c0 - 4) + j;
        o[j] = acc[p][j] * (hv[j] > 0.f ? 1.0f : 0.3f) * __ldg(a.scale + c);
      }
      float* dst = a.out + (((size_t)b * a.H + y) * a.W + x) * N;
      *reinterpret_cast<float4*>(dst + c0) = make_float4(o[0], o[1], o[2], o[3]);
      *reinterpret_cast<float4*>(dst + nh + c0) = make_float4(o[4], o[5], o[6], o[7]);
    }
  } else {  // B1
    if (!valid) return;
#pragma unroll
    for (int p = 0; p < 4; ++p) {
      const int y = 2 * qy + (p >> 1), x = 2 * qx + (p & 1);
      if (y >= a.H || x >= a.W) continue;
      float o[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = acc[p][j];
      if (a.wd != nullptr) {  // + conv_downsample^T(g_o)
        for (int co = 0; co < a.C; co += 4) {
          const float4 go = expand_go(a, b, y, x, co);
          const float gv[4] = {go.x, go.y, go.z, go.w};
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            if (gv[u] == 0.f) continue;
            const float4 wl = __ldg(reinterpret_cast<const float4*>(a.wd + (size_t)(co + u) * N + c0));
            const float4 wh = __ldg(reinterpret_cast<const float4*>(a.wd + (size_t)(co + u) * N + nh + c0));
            fma4(o, 0, gv[u], wl);
            fma4(o, 4, gv[u], wh);
          }
        }
      } else {  // + g_o (identity shortcut)
        const float4 gl = expand_go(a, b, y, x, c0), gh = expand_go(a, b, y, x, nh + c0);
        o[0] += gl.x, o[1] += gl.y, o[2] += gl.z, o[3] += gl.w;
        o[4] += gh.x, o[5] += gh.y, o[6] += gh.z, o[7] += gh.w;
      }
      float* dst = a.out + (((size_t)b * a.H + y) * a.W + x) * N;
      *reinterpret_cast<float4*>(dst + c0) = make_float4(o[0], o[1], o[2], o[3]);
      *reinterpret_cast<float4*>(dst + nh + c0) = make_float4(o[4], o[5], o[6], o[7]);
    }
  }
}

// First block, backward of conv1 (C -> 1 channel) + downsample^T + SELU' + first_bn scale: one thread per pixel.
__global__ void __launch_bounds__(256) sr_first_bwd_kernel(SrArgs a, const float* __restrict__ w1 /*[co][3][3] ref layout*/,
                                                            const float* __restrict__ wds /*[co]*/, int Cout,
                                                            const float* __restrict__ bn4, float* __restrict__ g_feat) {
  __shared__ float s_w[9 * 64];
  __shared__ float s_d[64];
  const int C = a.C;
  for (int i = threadIdx.x; i < 9 * C; i += blockDim.x) {
    const int co = i % C, tap = i / C;  // tap = r*3 + c over (frames, coeffs); flipped for the transpose
    const int r = tap / 3, c = tap % 3;
    // reference weight w[co][0][kh = coeff][kw = frame]; correlation transpose: g_x[p] = sum_t g[p - t] w[t]
    s_w[i] = co < Cout ? w1[co * 9 + (2 - c) * 3 + (2 - r)] : 0.f;
  }
  for (int i = threadIdx.x; i < C; i += blockDim.x) s_d[i] = i < Cout ? wds[i] : 0.f;
  __syncthreads();
  const int b = blockIdx.y;
  const int px = blockIdx.x * blockDim.x + threadIdx.x;
  if (px >= a.H * a.W) return;
  const int y = px / a.W, x = px % a.W;
  float acc = 0.f;
#pragma unroll 1
  for (int tap = 0; tap < 9; ++tap) {
    const int yy = y + tap / 3 - 1, xx = x + tap % 3 - 1;
    if (yy < 0 || yy >= a.H || xx < 0 || xx >= a.W) continue;
    const float* gp = a.in + (((size_t)b * a.H + yy) * a.W + xx) * C;
    const float* wp = s_w + tap * C;
    for (int c = 0; c < C; c += 4) {
      const float4 gv = __ldg(reinterpret_cast<const float4*>(gp + c));
      acc = fmaf(gv.x, wp[c], acc);
      acc = fmaf(gv.y, wp[c + 1], acc);
      acc = fmaf(gv.z, wp[c + 2], acc);
      acc = fmaf(gv.w, wp[c + 3], acc);
    }
  }
  for (int c = 0; c < C; c += 4) {
    const float4 go = expand_go(a, b, y, x, c);
    acc = fmaf(go.x, s_d[c], acc);
    acc = fmaf(go.y, s_d[c + 1], acc);
    acc = fmaf(go.z, s_d[c + 2], acc);
    acc = fmaf(go.w, s_d[c + 3], acc);
  }
  const float x0 = __ldg(a.x + ((size_t)b * (a.H + 2) + y + 1) * (a.W + 2) + x + 1);
  const float sc = __ldg(bn4 + 0) / sqrtf(__ldg(bn4 + 3) + 1e-5f);
  g_feat[((size_t)b * a.H + y) * a.W + x] = acc * selu_grad_from_out(x0) * sc;
}

constexpr int SR_C1_NMAX = 32;  // channels (padded) the first conv1 kernel stages per pixel
// First block, conv1 (1 -> C channels) + bn2 + LeakyReLU as its own kernel (round 2): thread = (pixel, 4 output channels), nine
// scalar loads of the 1-channel image (L1) and 36 FMAs; the generic F1 path staged bands and weight slabs for a contraction of
// length 9 and ran at a quarter of the HBM rate of its 812 MB output.
__global__ void __launch_bounds__(256) sr_first_conv1_kernel(const float* __restrict__ img /*(B, H+2, W+2)*/,
                                                              const float* __restrict__ wpk /*[tap][1][N]*/,
                                                              const float* __restrict__ bias, const float* __restrict__ scale,
                                                              const float* __restrict__ shift, float* __restrict__ h, int H, int W,
                                                              int N, int64_t n_px, tc::FastDiv dW, tc::FastDiv dH, tc::FastDiv dQ) {
  // thread = pixel: the nine image samples and the index arithmetic are shared by all N channels (one thread per 4 channels was
  // instruction-bound: 68 % issue slots at 22 % of the DRAM rate); weights, bias and the BatchNorm affine sit in shared memory.
  // The N outputs of a pixel are 4 N contiguous bytes, so a warp's float4 stores went out 96 bytes apart (24 cache lines per
  // instruction); they are staged per warp (row stride N + 4 floats: conflict-free) and written pixel-major, 512 contiguous bytes
  // per store instruction.
  __shared__ __align__(16) float s_w[9 * SR_C1_NMAX], s_b[SR_C1_NMAX], s_sc[SR_C1_NMAX], s_sh[SR_C1_NMAX];
  __shared__ __align__(16) float s_o[8][32 * (SR_C1_NMAX + 4)];
  __shared__ unsigned s_off[8][32];
  for (int i = threadIdx.x; i < 9 * N; i += blockDim.x) s_w[i] = wpk[i];
  for (int i = threadIdx.x; i < N; i += blockDim.x) s_b[i] = bias[i], s_sc[i] = scale[i], s_sh[i] = shift[i];
  __syncthreads();
  const int Wp = W + 2, RS = N + 4, Q = N >> 2;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* so = s_o[warp];
  unsigned* soff = s_off[warp];
  const unsigned total = (unsigned)n_px, step = gridDim.x * blockDim.x;
  for (unsigned base = (blockIdx.x * 8 + warp) * 32; base < total; base += step) {
    const unsigned i = base + lane;
    unsigned off = 0xffffffffu;
    if (i < total) {
      const int r = tc::fdiv((int)i, dW), x = (int)i - r * W;
      const int b = tc::fdiv(r, dH), y = r - b * H;
      const float* ip = img + ((size_t)b * (H + 2) + y) * Wp + x;  // top-left of the 3x3 window in the bordered image
      float v[9];
#pragma unroll
      for (int t = 0; t < 9; ++t) v[t] = __ldg(ip + (t / 3) * Wp + t % 3);
      off = (unsigned)((((size_t)b * (H + 2) + y + 1) * Wp + x + 1) * N);  // < 2^32: checked on the host
      float* sp = so + lane * RS;
      for (int c = 0; c < N; c += 4) {
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int t = 0; t < 9; ++t) {
          const float4 w = *reinterpret_cast<const float4*>(s_w + t * N + c);
          acc.x = fmaf(v[t], w.x, acc.x), acc.y = fmaf(v[t], w.y, acc.y), acc.z = fmaf(v[t], w.z, acc.z), acc.w = fmaf(v[t], w.w, acc.w);
        }
        const float4 bi = *reinterpret_cast<const float4*>(s_b + c), sc = *reinterpret_cast<const float4*>(s_sc + c);
        const float4 sh = *reinterpret_cast<const float4*>(s_sh + c);
        float o[4] = {fmaf(acc.x + bi.x, sc.x, sh.x), fmaf(acc.y + bi.y, sc.y, sh.y), fmaf(acc.z + bi.z, sc.z, sh.z),
                      fmaf(acc.w + bi.w, sc.w, sh.w)};
#pragma unroll
        for (int j = 0; j < 4; ++j) o[j] = o[j] > 0.f ? o[j] : 0.3f * o[j];
        *reinterpret_cast<float4*>(sp + c) = make_float4(o[0], o[1], o[2], o[3]);
      }
    }
    soff[lane] = off;
    __syncwarp();
    for (int f = lane; f < 32 * Q; f += 32) {
      const int p = tc::fdiv(f, dQ), q = f - p * Q;
      const unsigned po = soff[p];
      if (po != 0xffffffffu) *reinterpret_cast<float4*>(h + po + 4 * q) = *reinterpret_cast<const float4*>(so + p * RS + 4 * q);
    }
    __syncwarp();
  }
}

// Same operator, tiled (round 2; used when the expanded gradient `go` is materialised).  The kernel above reads every 24-channel
// gradient pixel nine times (once per tap) through L1: 7 GB per launch at B = 256, which is what its 0.76 ms was.  Here a CTA
// owns SR_FB_ROWS image rows: phase 1 reads each pixel of the rows + halo ONCE and forms its nine per-tap dot products
// P[t] = sum_c g[c] w[t][c] in shared memory; phase 2 gathers out = sum_t P_t[neighbour t] (the col2im of conv0_bwd.cu) and adds the
// 1x1 shortcut term from `go`.
constexpr int SR_FB_ROWS = 8;
__global__ void __launch_bounds__(256) sr_first_bwd_tiled_kernel(SrArgs a, const float* __restrict__ w1, const float* __restrict__ wds,
                                                                  int Cout, const float* __restrict__ bn4, const float* __restrict__ go,
                                                                  float* __restrict__ g_feat) {
  extern __shared__ __align__(16) float sm[];
  const int C = a.C, Wp = a.W + 2;
  float* s_w = sm;                 // [9][C] flipped taps
  float* s_d = s_w + 9 * C;        // [C]
  float* s_p = s_d + C;            // [(SR_FB_ROWS + 2) * Wp][9]
  for (int i = threadIdx.x; i < 9 * C; i += blockDim.x) {
    const int co = i % C, tap = i / C;
    const int r = tap / 3, c = tap % 3;
    s_w[i] = co < Cout ? w1[co * 9 + (2 - c) * 3 + (2 - r)] : 0.f;
  }
  for (int i = threadIdx.x; i < C; i += blockDim.x) s_d[i] = i < Cout ? wds[i] : 0.f;
  __syncthreads();
  const int b = blockIdx.y, y0 = blockIdx.x * SR_FB_ROWS;
  const int nslots = (SR_FB_ROWS + 2) * Wp;
  for (int sl = threadIdx.x; sl < nslots; sl += blockDim.x) {
    const int yy = y0 - 1 + sl / Wp, xx = sl % Wp - 1;
    float p[9];
#pragma unroll
    for (int t = 0; t < 9; ++t) p[t] = 0.f;
    if (yy >= 0 && yy < a.H && xx >= 0 && xx < a.W) {
      const float* gp = a.in + (((size_t)b * a.H + yy) * a.W + xx) * C;
      for (int c = 0; c < C; c += 4) {
        const float4 gv = __ldg(reinterpret_cast<const float4*>(gp + c));
#pragma unroll
        for (int t = 0; t < 9; ++t) {
          // one LDS.128 per tap (C is a multiple of 4 and s_w is 16-byte aligned; with a runtime C the compiler cannot know and
          // emitted four scalar loads per tap: 216 LDS for 216 FFMA per slot)
          const float4 w4 = *reinterpret_cast<const float4*>(s_w + t * C + c);
          p[t] = fmaf(gv.w, w4.w, fmaf(gv.z, w4.z, fmaf(gv.y, w4.y, fmaf(gv.x, w4.x, p[t]))));
        }
      }
    }
#pragma unroll
    for (int t = 0; t < 9; ++t) s_p[sl * 9 + t] = p[t];
  }
  __syncthreads();
  const float sc = __ldg(bn4 + 0) / sqrtf(__ldg(bn4 + 3) + 1e-5f);
  for (int i = threadIdx.x; i < SR_FB_ROWS * a.W; i += blockDim.x) {
    const int yl = i / a.W, x = i % a.W, y = y0 + yl;
    if (y >= a.H) continue;
    float acc = 0.f;
#pragma unroll
    for (int t = 0; t < 9; ++t) acc += s_p[((yl + t / 3) * Wp + x + t % 3) * 9 + t];
    const float* gop = go + (((size_t)b * (a.H + 2) + y + 1) * Wp + x + 1) * C;
    for (int c = 0; c < C; c += 4) {
      const float4 g = __ldg(reinterpret_cast<const float4*>(gop + c));
      const float4 d4 = *reinterpret_cast<const float4*>(s_d + c);
      acc = fmaf(g.w, d4.w, fmaf(g.z, d4.z, fmaf(g.y, d4.y, fmaf(g.x, d4.x, acc))));
    }
    const float x0 = __ldg(a.x + ((size_t)b * (a.H + 2) + y + 1) * Wp + x + 1);
    g_feat[((size_t)b * a.H + y) * a.W + x] = acc * selu_grad_from_out(x0) * sc;
  }
}

__global__ void sr_input_kernel(float* __restrict__ img, const float* __restrict__ bn4, int H, int W, int64_t n) {
  const float sc = bn4[0] / sqrtf(bn4[3] + 1e-5f), sh = bn4[1] - bn4[2] * sc;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int x = (int)(i % W), y = (int)((i / W) % H);
    const int64_t b = i / ((int64_t)W * H);
    float* p = img + (b * (H + 2) + y + 1) * (W + 2) + x + 1;
    *p = selu_f(fmaf(*p, sc, sh));
  }
}

__global__ void __launch_bounds__(64) sr_attention_fwd_kernel(const float* __restrict__ psum, int n_tiles,
                                                               const float* __restrict__ att_w,
                                                               const float* __restrict__ att_b, float* __restrict__ y,
                                                               int C, int Cout, float inv_hw) {
  __shared__ float s_avg[64];
  const int b = blockIdx.x, c = threadIdx.x;
  if (c < C) {
    float s = 0.f;
    for (int t = 0; t < n_tiles; ++t) s += psum[((size_t)b * n_tiles + t) * C + c];
    s_avg[c] = s * inv_hw;
  }
  __syncthreads();
  if (c < C) {
    float v = 0.f;
    if (c < Cout) {
      float z = __ldg(att_b + c);
      for (int k = 0; k < Cout; ++k) z = fmaf(__ldg(att_w + c * Cout + k), s_avg[k], z);
      v = 1.0f / (1.0f + expf(-z));
    }
    y[(size_t)b * C + c] = v;  // padded channels: 0, so x*y + y stays 0
  }
}

__global__ void sr_scale_pool_kernel(const float* __restrict__ xb, const float* __restrict__ y, float* __restrict__ xn,
                                     unsigned char* __restrict__ code2, int Hb, int Wb, int Hn, int Wn, int C, int pad,
                                     int64_t n4) {
  const int C4 = C >> 2;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = 4 * (int)(i % C4);
    const int nx = (int)((i / C4) % Wn), ny = (int)((i / ((int64_t)C4 * Wn)) % Hn);
    const int64_t b = i / ((int64_t)C4 * Wn * Hn);
    const float4 yv = *reinterpret_cast<const float4*>(y + b * C + c);
    float best[4];
    unsigned cd[4] = {0, 0, 0, 0};
#pragma unroll
    for (int p = 0; p < 4; ++p) {
      const float4 v = *reinterpret_cast<const float4*>(xb + ((b * Hb + 2 * ny + (p >> 1)) * Wb + 2 * nx + (p & 1)) * C + c);
      const float u[4] = {__fadd_rn(__fmul_rn(v.x, yv.x), yv.x), __fadd_rn(__fmul_rn(v.y, yv.y), yv.y),
                          __fadd_rn(__fmul_rn(v.z, yv.z), yv.z), __fadd_rn(__fmul_rn(v.w, yv.w), yv.w)};
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (p == 0 || u[k] > best[k]) {
          best[k] = u[k];
          cd[k] = (unsigned)p;
        }
    }
    *reinterpret_cast<float4*>(xn + ((b * (Hn + 2 * pad) + ny + pad) * (Wn + 2 * pad) + nx + pad) * C + c) =
        make_float4(best[0], best[1], best[2], best[3]);
    *reinterpret_cast<uchar4*>(code2 + ((b * Hn + ny) * Wn + nx) * C + c) =
        make_uchar4((unsigned char)cd[0], (unsigned char)cd[1], (unsigned char)cd[2], (unsigned char)cd[3]);
  }
}

// gadd[b][k] = (1 / (Hb Wb)) sum_c att_w[c][k] * gy[c] * y[c] (1 - y[c]),  gy[c] = sum_cells g_xn * (xb[winner] + 1)
__global__ void __launch_bounds__(256) sr_attention_bwd_kernel(const float* __restrict__ g_xn,
                                                                const unsigned char* __restrict__ code2,
                                                                const float* __restrict__ xb, const float* __restrict__ y,
                                                                const float* __restrict__ att_w, float* __restrict__ gadd,
                                                                int Hb, int Wb, int Hn, int Wn, int C, int Cout,
                                                                float inv_hw) {
  __shared__ float s_acc[256];
  __shared__ float s_gz[64];
  const int b = blockIdx.x, tid = threadIdx.x;
  const int R = blockDim.x / C;
  const int c = tid % C, r = tid / C;
  float acc = 0.f;
  if (r < R) {
    // eight cells per pass, their code / gradient loads issued together and then the eight dependent gathers: one cell at a time
    // the loop was a chain of 200 DRAM round trips per thread (one CTA per clip: 0.19 ms for the first block); same summation order
    constexpr int U = 8;
    const int n_cells = Hn * Wn;
    for (int cell0 = r; cell0 < n_cells; cell0 += R * U) {
      unsigned p[U];
      float g[U], xv[U];
      int ny[U], nx[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int cell = cell0 + u * R;
        const bool ok = cell < n_cells;
        ny[u] = ok ? cell / Wn : 0, nx[u] = ok ? cell - ny[u] * Wn : 0;
        const size_t o2 = (((size_t)b * Hn + ny[u]) * Wn + nx[u]) * C + c;
        p[u] = ok ? code2[o2] : 0xffu;
        g[u] = ok ? g_xn[o2] : 0.f;
      }
#pragma unroll
      for (int u = 0; u < U; ++u)
        xv[u] = p[u] != 0xffu ? xb[(((size_t)b * Hb + 2 * ny[u] + (p[u] >> 1)) * Wb + 2 * nx[u] + (p[u] & 1)) * C + c] : 0.f;
#pragma unroll
      for (int u = 0; u < U; ++u)
        if (p[u] != 0xffu) acc = fmaf(g[u], xv[u] + 1.0f, acc);
    }
  }
  s_acc[tid] = acc;
  __syncthreads();
  if (tid < C) {
    float s = 0.f;
    for (int k = 0; k < R; ++k) s += s_acc[k * C + tid];
    const float yv = y[(size_t)b * C + tid];
    s_gz[tid] = tid < Cout ? s * yv * (1.0f - yv) : 0.f;
  }
  __syncthreads();
  if (tid < C) {
    float s = 0.f;
    if (tid < Cout)
      for (int cc = 0; cc < Cout; ++cc) s = fmaf(__ldg(att_w + cc * Cout + tid), s_gz[cc], s);
    gadd[(size_t)b * C + tid] = s * inv_hw;
  }
}

// __umulhi(n, ceil(2^32 / d)) == n / d for every n with n * (d - 1) < 2^32
inline bool fastdiv_exact(int64_t n_max, int d) { return d <= 1 || n_max * (int64_t)(d - 1) < (1LL << 32); }

// The expanded gradient at conv2's output, materialised with a zero border for the tensor-core transposed convolution.
// Thread = (clip, 2x2 pool cell, 4 channels): the four pixels of a cell share the code bytes, the attention term and the pooled
// gradient (expand_go loads them per pixel), so one set of loads feeds four stores; the border is never written - the buffer is
// zeroed when it is allocated and nothing else touches it.  Same expressions as expand_go: same bits.
__global__ void sr_expand_go_kernel(SrArgs a, float* __restrict__ go, int64_t n_items, tc::FastDiv dC4, tc::FastDiv dQW, tc::FastDiv dQH) {
  const int C4 = a.C >> 2, Hp = a.H + 2, Wp = a.W + 2, QH = (a.H + 1) >> 1, QW = (a.W + 1) >> 1;
  const unsigned total = (unsigned)n_items, step = gridDim.x * blockDim.x;  // 32-bit index arithmetic (checked on the host)
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += step) {
    const int r1 = tc::fdiv((int)i, dC4), c = 4 * ((int)i - r1 * C4);
    const int r2 = tc::fdiv(r1, dQW), cx = r1 - r2 * QW;
    const int b = tc::fdiv(r2, dQH), cy = r2 - b * QH;
    const bool live = cy < a.Hb && cx < a.Wb;
    uchar4 c1 = make_uchar4(255, 255, 255, 255);
    float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
    if (live) {
      const size_t o1 = (((size_t)b * a.Hb + cy) * a.Wb + cx) * a.C + c;
      c1 = __ldg(reinterpret_cast<const uchar4*>(a.code1 + o1));
      g = __ldg(reinterpret_cast<const float4*>(a.gadd + (size_t)b * a.C + c));
      const int ny = cy >> 1, nx = cx >> 1;
      if (ny < a.Hn && nx < a.Wn) {
        const size_t o2 = (((size_t)b * a.Hn + ny) * a.Wn + nx) * a.C + c;
        const uchar4 c2 = __ldg(reinterpret_cast<const uchar4*>(a.code2 + o2));
        const float4 gn = __ldg(reinterpret_cast<const float4*>(a.g_xn + o2));
        const float4 yv = __ldg(reinterpret_cast<const float4*>(a.y + (size_t)b * a.C + c));
        const unsigned pos2 = (unsigned)(((cy & 1) << 1) | (cx & 1));
        if (c2.x == pos2) g.x += gn.x * yv.x;
        if (c2.y == pos2) g.y += gn.y * yv.y;
        if (c2.z == pos2) g.z += gn.z * yv.z;
        if (c2.w == pos2) g.w += gn.w * yv.w;
      }
    }
#pragma unroll
    for (int p = 0; p < 4; ++p) {
      const int y = 2 * cy + (p >> 1), x = 2 * cx + (p & 1);
      if (y >= a.H || x >= a.W) continue;
      const unsigned pos1 = (unsigned)p;
      float4 v;
      v.x = c1.x == pos1 ? g.x : 0.f;
      v.y = c1.y == pos1 ? g.y : 0.f;
      v.z = c1.z == pos1 ? g.z : 0.f;
      v.w = c1.w == pos1 ? g.w : 0.f;
      *reinterpret_cast<float4*>(go + (((size_t)b * Hp + y + 1) * Wp + x + 1) * a.C + c) = v;
    }
  }
}

// Shortcut term of a block's input gradient, added to the transposed conv1's output: g_x += downsample^T(g_o) (wd: [C][Ci]) or
// g_x += g_o (identity shortcut, Ci == C).  go (B, H+2, W+2, C) bordered, g_x (B, H, W, Ci).
__global__ void sr_shortcut_bwd_kernel(const float* __restrict__ go, const float* __restrict__ wd, float* __restrict__ g_x, int H,
                                       int W, int C, int Ci, int64_t n4) {
  const int C4 = Ci >> 2;
  const unsigned total = (unsigned)n4, step = gridDim.x * blockDim.x;
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += step) {
    const int c = 4 * (int)(i % (unsigned)C4);
    unsigned r = i / (unsigned)C4;
    const int x = (int)(r % (unsigned)W);
    r /= (unsigned)W;
    const int y = (int)(r % (unsigned)H), b = (int)(r / (unsigned)H);
    const float* gp = go + (((size_t)b * (H + 2) + y + 1) * (W + 2) + x + 1) * C;
    float4 acc = reinterpret_cast<float4*>(g_x)[i];
    if (wd != nullptr) {
      for (int co = 0; co < C; co += 4) {
        const float4 g = __ldg(reinterpret_cast<const float4*>(gp + co));
        const float gv[4] = {g.x, g.y, g.z, g.w};
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const float4 w = __ldg(reinterpret_cast<const float4*>(wd + (size_t)(co + u) * Ci + c));
          acc.x = fmaf(gv[u], w.x, acc.x), acc.y = fmaf(gv[u], w.y, acc.y), acc.z = fmaf(gv[u], w.z, acc.z), acc.w = fmaf(gv[u], w.w, acc.w);
        }
      }
    } else {
      const float4 g = __ldg(reinterpret_cast<const float4*>(gp + c));
      acc.x += g.x, acc.y += g.y, acc.z += g.z, acc.w += g.w;
    }
    reinterpret_cast<float4*>(g_x)[i] = acc;
  }
}

// wt[co][ci][a][b] = w[co][ci][b][a]: the engine's image is the transpose of the reference's (frames x coeffs)
__global__ void sr_tap_transpose_kernel(const float* __restrict__ w, float* __restrict__ wt, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    const int t = i % 9, a_ = t / 3, b_ = t % 3;
    wt[i] = w[i - t + b_ * 3 + a_];
  }
}

// ---- packing (live weights -> engine layouts) ----
// fwd: dst[((r*3+c)*CK + ci)*N + n] = w[n][ci][kh=c][kw=r];  bwd: dst[((r*3+c)*CKb + co)*Nb + ci] = w[co][ci][2-c][2-r]
__global__ void sr_pack_conv_kernel(const float* __restrict__ w, float* __restrict__ dst, int Cout, int Cin, int CK, int N,
                                    int bwd) {
  const int total = 9 * CK * N;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int n = i % N, k = (i / N) % CK, tap = i / (N * CK);
    const int r = tap / 3, c = tap % 3;
    float v = 0.f;
    if (!bwd) {
      if (n < Cout && k < Cin) v = w[((size_t)(n * Cin + k) * 3 + c) * 3 + r];
    } else {
      if (k < Cout && n < Cin) v = w[((size_t)(k * Cin + n) * 3 + (2 - c)) * 3 + (2 - r)];
    }
    dst[i] = v;
  }
}
// 1x1: fwd dst[ci*N + n] = w[n][ci]; bwd dst[co*Nb + ci] = w[co][ci]
__global__ void sr_pack_1x1_kernel(const float* __restrict__ w, float* __restrict__ dst, int Cout, int Cin, int K, int N,
                                   int bwd) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < K * N; i += gridDim.x * blockDim.x) {
    const int n = i % N, k = i / N;
    float v = 0.f;
    if (!bwd) {
      if (n < Cout && k < Cin) v = w[n * Cin + k];
    } else {
      if (k < Cout && n < Cin) v = w[k * Cin + n];
    }
    dst[i] = v;
  }
}
__global__ void sr_pack_vec_kernel(const float* __restrict__ b1, const float* __restrict__ b2, const float* __restrict__ bd,
                                   const float* __restrict__ bn_w, const float* __restrict__ bn_b,
                                   const float* __restrict__ bn_rm, const float* __restrict__ bn_rv, float* b1p, float* b2p,
                                   float* bdp, float* scale, float* shift, int Cout, int C) {
  const int c = threadIdx.x;
  if (c >= C) return;
  const bool ok = c < Cout;
  b1p[c] = ok ? b1[c] : 0.f;
  b2p[c] = ok ? b2[c] : 0.f;
  bdp[c] = (ok && bd != nullptr) ? bd[c] : 0.f;
  const float sc = ok ? bn_w[c] / sqrtf(bn_rv[c] + 1e-5f) : 0.f;
  scale[c] = sc;
  shift[c] = ok ? bn_b[c] - bn_rm[c] * sc : 0.f;
}

// ---- GRU + head ----
constexpr int GH = 64, G3 = 192, GL_MAX = 16;

__global__ void sr_gru_pack_kernel(SrGru g) {
  // transposes [192][K] -> [K][192]; head vector v = fc2 . fc1; bn_before_gru scale / shift
  const int tid = blockIdx.x * blockDim.x + threadIdx.x, nthr = gridDim.x * blockDim.x;
  for (int l = 0; l < 2; ++l)
    for (int d = 0; d < 2; ++d) {
      const int K = l == 0 ? 64 : 128;
      for (int i = tid; i < G3 * K; i += nthr) g.wihT[l][d][(i % K) * G3 + i / K] = g.w_ih[l][d][i];
      for (int i = tid; i < G3 * GH; i += nthr) g.whhT[l][d][(i % GH) * G3 + i / GH] = g.w_hh[l][d][i];
    }
  for (int k = tid; k < 128; k += nthr) {
    float s = 0.f;
    for (int j = 0; j < 128; ++j) s = fmaf(g.fc2_w[j], g.fc1_w[j * 128 + k], s);
    g.v[k] = s;
  }
  for (int c = tid; c < 64; c += nthr) {
    const float sc = g.bn_w[c] / sqrtf(g.bn_rv[c] + 1e-5f);
    g.bn_scale[c] = sc;
    g.bn_shift[c] = g.bn_b[c] - g.bn_rm[c] * sc;
  }
}

__device__ __forceinline__ float sigmoid_f(float v) { return 1.0f / (1.0f + expf(-v)); }

// acc += sum_k w[k * stride] * v[k] (v in shared memory), one sequential fmaf chain in k like the plain loop (same bits), but with
// the NEXT 16 weights already in flight while 16 are consumed: the plain loop kept 2-4 loads ahead of their use and the kernels sat
// on the L2 latency of these weight reads (two layers x two directions exceed L1).
template <int N>
__device__ __forceinline__ float dot_prefetched(const float* __restrict__ w, int stride, const float* v, float acc) {
  static_assert(N % 16 == 0, "chunks of 16");
  float cur[16], nxt[16];
#pragma unroll
  for (int u = 0; u < 16; ++u) cur[u] = __ldg(w + (size_t)u * stride);
#pragma unroll 1
  for (int k0 = 0; k0 < N; k0 += 16) {
    if (k0 + 16 < N) {
#pragma unroll
      for (int u = 0; u < 16; ++u) nxt[u] = __ldg(w + (size_t)(k0 + 16 + u) * stride);
    }
#pragma unroll
    for (int u = 0; u < 16; ++u) acc = fmaf(cur[u], v[k0 + u], acc);
#pragma unroll
    for (int u = 0; u < 16; ++u) cur[u] = nxt[u];
  }
  return acc;
}

// One CTA per clip, 384 threads = 2 directions x 192 gate rows.
__global__ void __launch_bounds__(384) sr_gru_fwd_kernel(SrGru g, const float* __restrict__ xn, float* __restrict__ logits,
                                                          int L) {
  __shared__ float s_in[GL_MAX][128];   // current layer input sequence
  __shared__ float s_out[GL_MAX][128];  // current layer output sequence
  __shared__ float s_h[2][GH];
  __shared__ float s_gi[2][G3], s_gh[2][G3];
  __shared__ float s_red[4];
  const int b = blockIdx.x, tid = threadIdx.x, d = tid / G3, j = tid % G3;
  for (int i = tid; i < L * GH; i += 384) {
    const int t = i / GH, c = i % GH;
    const float v = selu_f(fmaf(xn[((size_t)b * L + t) * GH + c], g.bn_scale[c], g.bn_shift[c]));
    s_in[t][c] = v;
    g.xin[((size_t)b * L + t) * GH + c] = v;
  }
  __syncthreads();
  for (int l = 0; l < 2; ++l) {
    const int K = l == 0 ? 64 : 128;
    if (tid < 2 * GH) s_h[tid / GH][tid % GH] = 0.f;
    __syncthreads();
    const float* wih = g.wihT[l][d];
    const float* whh = g.whhT[l][d];
    const float bi = g.b_ih[l][d][j], bh = g.b_hh[l][d][j];
    for (int s = 0; s < L; ++s) {
      const int t = d == 0 ? s : L - 1 - s;
      float gi = bi, gh = bh;
      gi = l == 0 ? dot_prefetched<64>(wih + j, G3, s_in[t], gi) : dot_prefetched<128>(wih + j, G3, s_in[t], gi);
      gh = dot_prefetched<GH>(whh + j, G3, s_h[d], gh);
      s_gi[d][j] = gi;
      s_gh[d][j] = gh;
      __syncthreads();
      if (j < GH) {
        const float r = sigmoid_f(s_gi[d][j] + s_gh[d][j]);
        const float z = sigmoid_f(s_gi[d][GH + j] + s_gh[d][GH + j]);
        const float hn = s_gh[d][2 * GH + j];
        const float n = tanhf(s_gi[d][2 * GH + j] + r * hn);
        const float hnew = (1.0f - z) * n + z * s_h[d][j];
        s_h[d][j] = hnew;
        s_out[t][d * GH + j] = hnew;
        float* gs = g.gates + ((((size_t)b * 2 + l) * 2 + d) * L + t) * 4 * GH;
        gs[j] = r;
        gs[GH + j] = z;
        gs[2 * GH + j] = n;
        gs[3 * GH + j] = hn;
      }
      __syncthreads();
    }
    for (int i = tid; i < L * 128; i += 384) {
      const float v = s_out[i / 128][i % 128];
      g.outs[(((size_t)b * 2 + l) * L + i / 128) * 128 + i % 128] = v;
      s_in[i / 128][i % 128] = v;
    }
    __syncthreads();
  }
  // head: fc2(fc1(out[L-1]))
  float part = 0.f;
  if (tid < 128) {
    float v = g.fc1_b[tid];
    for (int k = 0; k < 128; ++k) v = fmaf(__ldg(g.fc1_w + tid * 128 + k), s_in[L - 1][k], v);
    part = v * g.fc2_w[tid];
  }
  part = warp_sum(part);
  if (tid < 128 && (tid & 31) == 0) s_red[tid >> 5] = part;
  __syncthreads();
  if (tid == 0) logits[b] = s_red[0] + s_red[1] + s_red[2] + s_red[3] + g.fc2_b[0];
}

// One CTA per clip, 256 threads = 2 directions x 128.
__global__ void __launch_bounds__(256) sr_gru_bwd_kernel(SrGru g, const float* __restrict__ xn,
                                                          const float* __restrict__ logits, const long long* __restrict__ y,
                                                          int mode, float inv_n, const float* __restrict__ coef,
                                                          float* __restrict__ g_xn, int L) {
  __shared__ float s_gout[GL_MAX][128];    // gradient w.r.t. the current layer's output sequence
  __shared__ float s_gin[2][GL_MAX][128];  // per-direction gradient w.r.t. the current layer's input sequence
  __shared__ float s_dh[2][GH];
  __shared__ float s_ggi[2][G3], s_ggh[2][G3];
  const int b = blockIdx.x, tid = threadIdx.x, d = tid >> 7, i = tid & 127;
  float go = 1.0f;
  if (mode == 2) go = coef[b];
  if (mode == 0) go = 2.0f * (sigmoid_f(2.0f * logits[b]) - (float)y[b]) * inv_n;
  for (int k = tid; k < L * 128; k += 256) s_gout[k / 128][k % 128] = 0.f;
  __syncthreads();
  if (tid < 128) s_gout[L - 1][tid] = go * g.v[tid];
  __syncthreads();
  for (int l = 1; l >= 0; --l) {
    const int K = l == 0 ? 64 : 128;
    if (i < GH) s_dh[d][i] = 0.f;
    for (int k = i; k < L * 128; k += 128) s_gin[d][k / 128][k % 128] = 0.f;
    __syncthreads();
    const float* wih = g.w_ih[l][d];  // [192][K]
    const float* whh = g.w_hh[l][d];  // [192][64]
    for (int s = L - 1; s >= 0; --s) {
      const int t = d == 0 ? s : L - 1 - s;          // time index processed at step s
      const int tp = d == 0 ? t - 1 : t + 1;         // time index of the previous hidden state
      if (i < GH) {
        const float* gs = g.gates + ((((size_t)b * 2 + l) * 2 + d) * L + t) * 4 * GH;
        const float r = gs[i], z = gs[GH + i], n = gs[2 * GH + i], hn = gs[3 * GH + i];
        const float hprev = s > 0 ? g.outs[(((size_t)b * 2 + l) * L + tp) * 128 + d * GH + i] : 0.f;
        const float dh = s_dh[d][i] + s_gout[t][d * GH + i];
        const float dn = dh * (1.0f - z);
        const float dz = dh * (hprev - n);
        const float dnp = dn * (1.0f - n * n);
        const float drp = dnp * hn * r * (1.0f - r);
        const float dzp = dz * z * (1.0f - z);
        s_ggi[d][i] = drp;
        s_ggi[d][GH + i] = dzp;
        s_ggi[d][2 * GH + i] = dnp;
        s_ggh[d][i] = drp;
        s_ggh[d][GH + i] = dzp;
        s_ggh[d][2 * GH + i] = dnp * r;
        s_dh[d][i] = dh * z;  // direct path; the W_hh^T term is added below
      }
      __syncthreads();
      if (i < K) s_gin[d][t][i] = dot_prefetched<G3>(wih + i, K, s_ggi[d], 0.f);
      float accH = 0.f;
      if (i < GH) accH = dot_prefetched<G3>(whh + i, GH, s_ggh[d], 0.f);
      __syncthreads();
      if (i < GH) s_dh[d][i] += accH;
      __syncthreads();
    }
    for (int k = tid; k < L * 128; k += 256) s_gout[k / 128][k % 128] = s_gin[0][k / 128][k % 128] + s_gin[1][k / 128][k % 128];
    __syncthreads();
  }
  // bn_before_gru + SELU backward (from the saved output)
  for (int k = tid; k < L * GH; k += 256) {
    const int t = k / GH, c = k % GH;
    const float v = g.xin[((size_t)b * L + t) * GH + c];
    g_xn[((size_t)b * L + t) * GH + c] = s_gout[t][c] * selu_grad_from_out(v) * g.bn_scale[c];
  }
}

__global__ void sr_pack4_kernel(const float* w, const float* b, const float* rm, const float* rv, float* bn4) {
  bn4[0] = w[0], bn4[1] = b[0], bn4[2] = rm[0], bn4[3] = rv[0];
}

inline int ew_blocks(int64_t n) {
  int64_t b = (n + 255) / 256;
  return (int)(b > 148 * 16 ? 148 * 16 : b);
}

// Quads per CTA: whole quad rows (so that a tile's halo band is its rows + 2, not up to 3x its rows as with 32 quads of a
// 40-quad row), as many as 512 threads (qpc x N/8) and half an SM's shared memory allow; 32 when a single row does not fit.
int sr_band_rows(int QH, int QW, int qpc) {  // rows of the halo band of a tile (row-aligned tiles need exactly theirs + 2)
  if (qpc % QW == 0) return 2 * std::min(qpc / QW, QH) + 2;
  return band_rows_max(QH, QW, 1, qpc);
}
bool sr_slab_all(int N, int CK) {
  static const int on = [] {
    const char* e = getenv("ADVB_SR_SLAB_ALL");
    return e != nullptr ? atoi(e) : 1;
  }();
  // 24 KB = block 0's 24 x 24 layers (conv2 forward 107 -> 99 ms, backward 103 -> 94 ms per PGD-40 call at B = 256); with a
  // 64 KB limit block 2's 64 x 24 backward also qualifies and got slower (36 -> 54 ms: it loses a CTA per SM)
  static const int limit_kb = [] {
    const char* e = getenv("ADVB_SR_SLAB_KB");
    return e != nullptr ? atoi(e) : 24;
  }();
  return on != 0 && 9 * CK * N * (int)sizeof(float) <= limit_kb * 1024;
}
int sr_qpc(int N, int CK, int QH, int QW) {
  const int CKp = (CK % 4 == 0) ? CK + 4 : CK;
  const int wfl = (sr_slab_all(N, CK) ? 9 : 1) * CK * N;
  const int G = N / 8;
  int best = 32;
  // measured (B = 256, ms per PGD-40 call): row-aligned tiles of <= 256 threads pay where the band is deep and the thread
  // tile narrow - block 0 conv2 forward 160 -> 118, backward 231 -> 136, block 2 conv1 backward 46 -> 37; the N = 64 layers
  // (ADVB_SR_WIDE=1) and the 1-channel first conv got slower and keep 32 quads; 480-thread tiles (1 CTA per SM) lost to
  // 240-thread ones everywhere
  static const int wide = [] {
    const char* e = getenv("ADVB_SR_WIDE");
    return e != nullptr ? atoi(e) : 0;
  }();
  if ((N > 32 && !wide) || CK < 8) return best;
  static const int max_threads = [] {
    const char* e = getenv("ADVB_SR_THREADS");
    const int v = e != nullptr ? atoi(e) : 256;
    return v > 256 ? 256 : v;  // the kernel's __launch_bounds__
  }();
  for (int rows = 1; rows <= QH && rows * QW * G <= max_threads; ++rows) {
    const int qpc = rows * QW;
    const size_t smem = (size_t)(sr_band_rows(QH, QW, qpc) * (2 * QW + 2) * CKp + wfl + qpc * N) * sizeof(float);
    if (smem > 110 * 1024) break;
    if (qpc >= 24) best = qpc;
  }
  return best;
}

template <int MODE>
int launch_conv(SrArgs a, bool floor_quads, const char* tag, cudaStream_t stream) {
  const int QW = floor_quads ? a.W / 2 : (a.W + 1) / 2, QH = floor_quads ? a.H / 2 : (a.H + 1) / 2;
  ADVB_CHECK(QW > 0 && QH > 0, "empty SpecRNet conv output");
  const int CKp = (a.CK % 4 == 0) ? a.CK + 4 : a.CK;
  a.qpc = sr_qpc(a.N, a.CK, QH, QW);
  a.slab_all = sr_slab_all(a.N, a.CK) ? 1 : 0;
  int band = sr_band_rows(QH, QW, a.qpc) * (2 * QW + 2) * CKp;
  band = (band + 3) & ~3;
  size_t smem = (size_t)(band + (a.slab_all ? 9 : 1) * a.CK * a.N + (MODE == F2 ? a.qpc * a.N : 0)) * sizeof(float);
  if (MODE == F2 && a.c2 != nullptr) smem = (size_t)a.qpc * a.N * sizeof(float);  // no band, no weights: only the partial sums
  ADVB_CHECK(smem <= 227 * 1024, "SpecRNet conv tile does not fit shared memory");
  ADVB_CHECK(a.N % 8 == 0 && a.N <= 64, "SpecRNet conv: N must be a multiple of 8, <= 64");
  ADVB_CUDA_OK(cudaFuncSetAttribute(sr_conv_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid(cdiv(QH * QW, a.qpc), a.B);
  sr_conv_kernel<MODE><<<grid, a.qpc * (a.N / 8), smem, stream>>>(a, band);
  ADVB_KERNEL_OK(tag, stream);
  return 0;
}

SrArgs base_args(const SrBlock& k, int B) {
  SrArgs a{};
  a.B = B, a.H = k.H, a.W = k.W, a.Hb = k.Hb, a.Wb = k.Wb, a.Hn = k.Hn, a.Wn = k.Wn, a.C = k.C;
  a.g_xn = k.g_xn, a.code2 = k.code2, a.code1 = k.code1, a.y = k.y, a.gadd = k.gadd, a.h = k.h;
  a.scale = k.bn_scale, a.shift = k.bn_shift;
  return a;
}

}  // namespace

int sr_conv2_tiles(int H, int W, int C) { return cdiv((H / 2) * (W / 2), sr_qpc(C, C, H / 2, W / 2)); }

int sr_pack_first_bn(const float* w, const float* b, const float* rm, const float* rv, float* bn4, cudaStream_t stream) {
  sr_pack4_kernel<<<1, 1, 0, stream>>>(w, b, rm, rv, bn4);
  ADVB_KERNEL_OK("sr_pack", stream);
  return 0;
}

int sr_pack_block(SrBlock& k, cudaStream_t stream) {
  sr_pack_conv_kernel<<<cdiv(9 * k.Ci * k.C, 256), 256, 0, stream>>>(k.w1, k.w1f, k.Cout, k.Cin, k.Ci, k.C, 0);
  ADVB_KERNEL_OK("sr_pack", stream);
  sr_pack_conv_kernel<<<cdiv(9 * k.C * k.C, 256), 256, 0, stream>>>(k.w2, k.w2f, k.Cout, k.Cout, k.C, k.C, 0);
  ADVB_KERNEL_OK("sr_pack", stream);
  sr_pack_conv_kernel<<<cdiv(9 * k.C * k.C, 256), 256, 0, stream>>>(k.w2, k.w2d, k.Cout, k.Cout, k.C, k.C, 1);
  ADVB_KERNEL_OK("sr_pack", stream);
  if (k.Cin > 1) {
    sr_pack_conv_kernel<<<cdiv(9 * k.C * k.Ci, 256), 256, 0, stream>>>(k.w1, k.w1d, k.Cout, k.Cin, k.C, k.Ci, 1);
    ADVB_KERNEL_OK("sr_pack", stream);
  }
  if (k.downsample) {
    sr_pack_1x1_kernel<<<cdiv(k.Ci * k.C, 256), 256, 0, stream>>>(k.wds, k.wdf, k.Cout, k.Cin, k.Ci, k.C, 0);
    ADVB_KERNEL_OK("sr_pack", stream);
    if (k.Cin > 1) {
      sr_pack_1x1_kernel<<<cdiv(k.C * k.Ci, 256), 256, 0, stream>>>(k.wds, k.wdd, k.Cout, k.Cin, k.C, k.Ci, 1);
      ADVB_KERNEL_OK("sr_pack", stream);
    }
  }
  if (k.tcf1 != nullptr) {
    sr_tap_transpose_kernel<<<cdiv(k.Cout * k.Cin * 9, 256), 256, 0, stream>>>(k.w1, k.w1t, k.Cout * k.Cin * 9);
    ADVB_KERNEL_OK("sr_pack", stream);
    ADVB_TRY(conv_tc_pack_padded(k.w1t, k.tcf1, k.tcd1, k.Cout, k.Cin, 3, k.C == 64 ? 64 : 32, k.Ci == 64 ? 64 : 32, stream));
  }
  if (k.tcf2 != nullptr) {
    sr_tap_transpose_kernel<<<cdiv(k.Cout * k.Cout * 9, 256), 256, 0, stream>>>(k.w2, k.w2t, k.Cout * k.Cout * 9);
    ADVB_KERNEL_OK("sr_pack", stream);
    if (k.C == 64) ADVB_TRY(conv_tc_pack(k.w2t, k.tcf2, k.tcd2, k.Cout, k.Cout, 3, stream));
    else ADVB_TRY(conv_tc_pack_padded(k.w2t, k.tcf2, k.tcd2, k.Cout, k.Cout, 3, 32, 32, stream));  // 20 x 20 -> K = 24, N = 32 (zero rows / columns)
  }
  sr_pack_vec_kernel<<<1, 64, 0, stream>>>(k.b1, k.b2, k.downsample ? k.bds : nullptr, k.bn_w, k.bn_b, k.bn_rm, k.bn_rv,
                                          k.b1p, k.b2p, k.bdp, k.bn_scale, k.bn_shift, k.Cout, k.C);
  ADVB_KERNEL_OK("sr_pack", stream);
  return 0;
}

int sr_pack_gru(SrGru& g, cudaStream_t stream) {
  sr_gru_pack_kernel<<<32, 256, 0, stream>>>(g);
  ADVB_KERNEL_OK("sr_gru_pack", stream);
  return 0;
}

int sr_input_forward(float* img, const float* bn4, int B, int H, int W, cudaStream_t stream) {
  const int64_t n = (int64_t)B * H * W;
  sr_input_kernel<<<ew_blocks(n), 256, 0, stream>>>(img, bn4, H, W, n);
  ADVB_KERNEL_OK("sr_input", stream);
  return 0;
}

// profiler labels must outlive the call (prof_mark keeps the pointer): string literals per block
struct SrTags {
  const char *conv1, *conv2, *conv2_bwd, *conv1_bwd, *conv2_tc, *expand_go, *shortcut_bwd;
};
const SrTags& sr_tags(const char* tag) {
  static const SrTags t0{"sr_b0_conv1", "sr_b0_conv2", "sr_b0_conv2_bwd", "sr_b0_conv1_bwd", "sr_b0_conv2_tc", "sr_b0_expand_go", "sr_b0_shortcut_bwd"};
  static const SrTags t2{"sr_b2_conv1", "sr_b2_conv2", "sr_b2_conv2_bwd", "sr_b2_conv1_bwd", "sr_b2_conv2_tc", "sr_b2_expand_go", "sr_b2_shortcut_bwd"};
  static const SrTags t4{"sr_b4_conv1", "sr_b4_conv2", "sr_b4_conv2_bwd", "sr_b4_conv1_bwd", "sr_b4_conv2_tc", "sr_b4_expand_go", "sr_b4_shortcut_bwd"};
  return tag[4] == '0' ? t0 : (tag[4] == '2' ? t2 : t4);
}

int sr_block_forward(const SrBlock& k, const float* x, int B, const char* tag, cudaStream_t stream) {
  const SrTags& t = sr_tags(tag);
  SrArgs a = base_args(k, B);
  // CKr: the padded channels (20 -> 24) carry zero weights; the contraction skips them (bit-identical up to the sign of zero)
  a.CK = k.Ci, a.N = k.C, a.CKr = k.Ci % 4 == 0 ? std::min(k.Ci, (k.Cin + 3) / 4 * 4) : k.Ci;
  a.in = x, a.wpk = k.w1f, a.bias = k.b1p, a.out = k.h;
  if (k.Ci == 1 && k.tc2) {  // first block: 1-channel input, a 9-term contraction
    const int64_t n_px = (int64_t)B * k.H * k.W;
    ADVB_CHECK(k.C <= 64 && n_px < (1LL << 31), "SpecRNet first conv1: at most 64 channels, 2^31 pixels");
    ADVB_CHECK(k.C <= SR_C1_NMAX && k.C % 4 == 0 && (int64_t)B * (k.H + 2) * (k.W + 2) * k.C < (1LL << 32) && fastdiv_exact(n_px, k.W),
               "SpecRNet first conv1: channel count / index range");
    sr_first_conv1_kernel<<<ew_blocks(n_px), 256, 0, stream>>>(x, k.w1f, k.b1p, k.bn_scale, k.bn_shift, k.h, k.H, k.W, k.C, n_px,
                                                               tc::make_fastdiv(k.W), tc::make_fastdiv(k.H), tc::make_fastdiv(k.C / 4));
    ADVB_KERNEL_OK(t.conv1, stream);
  } else if (k.tc1) {  // conv1 -> bn2 -> LeakyReLU on the tensor cores (affine + activation in the epilogue)
    P3Plain p;
    p.in = x, p.out = k.h, p.out_pad = 1, p.wpack = k.tcf1, p.bias = k.b1p;
    p.aff_scale = k.bn_scale, p.aff_shift = k.bn_shift, p.act_slope = 0.3f;
    p.B = B, p.H = k.H, p.W = k.W, p.Cin = k.Ci, p.Cout = k.C, p.tag = t.conv1;
    ADVB_TRY(conv_p3_plain_forward(p, stream));
  } else {
    ADVB_TRY(launch_conv<F1>(a, false, t.conv1, stream));
  }
  a.CK = k.C, a.CKr = std::min(k.C, (k.Cout + 3) / 4 * 4), a.in = k.h, a.wpk = k.w2f, a.bias = k.b2p, a.out = k.xb;
  a.x = x, a.Ci = k.Ci, a.wd = k.downsample ? k.wdf : nullptr, a.bd = k.bdp, a.code1w = k.code1, a.psum = k.psum;
  if (k.tc2) {  // conv2(h) on the tensor cores; the fused kernel below keeps bias + identity + max-pool + arg-max + channel sums
    P3Plain p;
    p.in = k.h, p.out = k.c2, p.out_pad = 0, p.wpack = k.tcf2;
    p.B = B, p.H = k.H, p.W = k.W, p.Cin = k.C, p.Cout = k.C, p.tag = t.conv2_tc;
    ADVB_TRY(conv_p3_plain_forward(p, stream));
    a.c2 = k.c2;
  }
  ADVB_TRY(launch_conv<F2>(a, true, t.conv2, stream));
  sr_attention_fwd_kernel<<<B, 64, 0, stream>>>(k.psum, k.n_tiles, k.att_w, k.att_b, k.y, k.C, k.Cout,
                                               1.0f / (float)(k.Hb * k.Wb));
  ADVB_KERNEL_OK("sr_attention", stream);
  const int64_t n4 = (int64_t)B * k.Hn * k.Wn * (k.C / 4);
  sr_scale_pool_kernel<<<ew_blocks(n4), 256, 0, stream>>>(k.xb, k.y, k.xn, k.code2, k.Hb, k.Wb, k.Hn, k.Wn, k.C, k.xn_pad,
                                                         n4);
  ADVB_KERNEL_OK("sr_scale_pool", stream);
  return 0;
}

int sr_block_backward(const SrBlock& k, const float* x, float* g_x, int B, bool first, const float* bn4, const char* tag,
                      cudaStream_t stream) {
  const SrTags& t = sr_tags(tag);
  const int threads = (256 / k.C) * k.C;
  sr_attention_bwd_kernel<<<B, threads, 0, stream>>>(k.g_xn, k.code2, k.xb, k.y, k.att_w, k.gadd, k.Hb, k.Wb, k.Hn, k.Wn,
                                                    k.C, k.Cout, 1.0f / (float)(k.Hb * k.Wb));
  ADVB_KERNEL_OK("sr_attention_bwd", stream);
  SrArgs a = base_args(k, B);
  a.CK = k.C, a.CKr = std::min(k.C, (k.Cout + 3) / 4 * 4), a.N = k.C, a.wpk = k.w2d, a.out = k.g_c1;
  if (k.tc2) {  // g_o materialised once (zero border), conv2^T on the tensor cores with the LeakyReLU' * bn2-scale factor in its epilogue
    const int QH = (k.H + 1) / 2, QW = (k.W + 1) / 2;
    const int64_t n4 = (int64_t)B * QH * QW * (k.C / 4);  // one thread per (2x2 pool cell, 4 channels)
    ADVB_CHECK((int64_t)B * (k.H + 2) * (k.W + 2) * k.C < (1LL << 32) && n4 < (1LL << 31),
               "SpecRNet: batch too large for the 32-bit element index of the expanded gradient");
    ADVB_CHECK(fastdiv_exact(n4, k.C / 4) && fastdiv_exact(n4 / (k.C / 4), QW), "SpecRNet: index range of the multiply-high division");
    sr_expand_go_kernel<<<ew_blocks(n4), 256, 0, stream>>>(a, k.go, n4, tc::make_fastdiv(k.C / 4), tc::make_fastdiv(QW),
                                                           tc::make_fastdiv(QH));
    ADVB_KERNEL_OK(t.expand_go, stream);
    P3Plain p;
    p.in = k.go, p.wpack = k.tcd2, p.mul_h = k.h, p.mul_scale = k.bn_scale, p.mul_slope = 0.3f;
    p.out = k.tc1 ? k.g_c1b : k.g_c1, p.out_pad = k.tc1 ? 1 : 0;  // bordered when the transposed conv1 below is on this kernel too
    p.B = B, p.H = k.H, p.W = k.W, p.Cin = k.C, p.Cout = k.C, p.tag = t.conv2_bwd;
    ADVB_TRY(conv_p3_plain_forward(p, stream));
  } else {
    ADVB_TRY(launch_conv<B2>(a, false, t.conv2_bwd, stream));
  }
  a.in = k.g_c1, a.x = x;
  if (first && k.tc2) {
    ADVB_CHECK(k.C % 4 == 0, "SpecRNet first block: channel count padded to a multiple of 4");
    dim3 grid(cdiv(k.H, SR_FB_ROWS), B);
    const size_t smem = (size_t)(10 * k.C + (SR_FB_ROWS + 2) * (k.W + 2) * 9) * sizeof(float);
    sr_first_bwd_tiled_kernel<<<grid, 256, smem, stream>>>(a, k.w1, k.wds, k.Cout, bn4, k.go, g_x);
    ADVB_KERNEL_OK(t.conv1_bwd, stream);
    return 0;
  }
  if (first) {
    dim3 grid(cdiv(k.H * k.W, 256), B);
    sr_first_bwd_kernel<<<grid, 256, 0, stream>>>(a, k.w1, k.wds, k.Cout, bn4, g_x);
    ADVB_KERNEL_OK(t.conv1_bwd, stream);
    return 0;
  }
  a.CK = k.C, a.CKr = std::min(k.C, (k.Cout + 3) / 4 * 4), a.N = k.Ci, a.wpk = k.w1d, a.out = g_x, a.wd = k.downsample ? k.wdd : nullptr;
  if (k.tc1) {  // g_x = conv1^T(g_c1) on the tensor cores, then + downsample^T(g_o) | + g_o from the materialised go
    P3Plain p;
    p.in = k.g_c1b, p.out = g_x, p.out_pad = 0, p.wpack = k.tcd1;
    p.B = B, p.H = k.H, p.W = k.W, p.Cin = k.C, p.Cout = k.Ci, p.tag = t.conv1_bwd;
    ADVB_TRY(conv_p3_plain_forward(p, stream));
    const int64_t n4 = (int64_t)B * k.H * k.W * (k.Ci / 4);
    sr_shortcut_bwd_kernel<<<ew_blocks(n4), 256, 0, stream>>>(k.go, k.downsample ? k.wdd : nullptr, g_x, k.H, k.W, k.C, k.Ci, n4);
    ADVB_KERNEL_OK(t.shortcut_bwd, stream);
    return 0;
  }
  ADVB_TRY(launch_conv<B1>(a, false, t.conv1_bwd, stream));
  return 0;
}

int sr_gru_forward(const SrGru& g, const float* xn, float* logits, int B, int L, cudaStream_t stream) {
  ADVB_CHECK(L >= 1 && L <= GL_MAX, "SpecRNet GRU sequence length out of range (clip too short or too long)");
  sr_gru_fwd_kernel<<<B, 384, 0, stream>>>(g, xn, logits, L);
  ADVB_KERNEL_OK("sr_gru_fwd", stream);
  return 0;
}

int sr_gru_backward(const SrGru& g, const float* xn, const float* logits, const long long* y, int mode, int n_global,
                    const float* coef, float* g_xn, int B, int L, cudaStream_t stream) {
  sr_gru_bwd_kernel<<<B, 256, 0, stream>>>(g, xn, logits, y, mode, 1.0f / (float)n_global, coef, g_xn, L);
  ADVB_KERNEL_OK("sr_gru_bwd", stream);
  return 0;
}

}  // namespace advb
