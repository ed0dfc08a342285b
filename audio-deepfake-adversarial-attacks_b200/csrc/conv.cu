// LCNN convolution blocks: conv -> +bias -> Max-Feature-Map -> [2x2 max-pool] -> [BatchNorm(eval)], forward and
// input-gradient backward, fp32 SIMT path (sm_100a).
//
// Replaces the nn.Sequential of src/models/lcnn.py:120-157 (Conv2d, MaxFeatureMap2D :89-95, MaxPool2d,
// BatchNorm2d(affine=False)) and its autograd backward w.r.t. the input (no weight gradients, so no
// conv input is saved; only a 3-bit code per stage output: pool arg-max position | MFM half << 2).
//
// Layout: activations NHWC fp32 with a zero spatial border equal to the consuming conv's padding (the border is
// zeroed once at workspace creation and never written).  Stage gradients are compact NHWC (no border).
//
// Work decomposition: a CTA owns 32 consecutive 2x2 "quads" of pre-pool output pixels (row-major over the quad
// grid) of one clip and all output channels; it stages the full-width band of input rows it needs in shared
// memory once, then streams one tap's weight slab at a time.  A thread owns one quad and 8 output channels
// {4g..4g+3} u {C/2+4g..C/2+4g+3}, so MFM (channel c vs c+C/2) and the 2x2 pool are register-local.
// Backward is the same contraction over an on-the-fly expanded gradient band (BN scale, un-pool, un-MFM applied
// while staging), with tap-flipped / transposed weights; the expanded gradient never touches HBM.
#include "conv.cuh"
#include "conv_core.cuh"

namespace advb {

namespace {

using namespace convcore;

template <int KS, bool POOL>
__global__ void __launch_bounds__(512) conv_fwd_kernel(ConvFwdArgs a) {
  extern __shared__ __align__(16) float smem[];
  constexpr int pc = KS / 2;
  const TileGeom g = tile_geom(a.H, a.W, POOL, pc);
  const int CK = a.Cin;
  const int CKp = (CK & 3) == 0 ? CK + 4 : CK;
  float* band = smem;
  float* w_s = smem + a.band_floats;
  const int b = blockIdx.y, tid = threadIdx.x, nt = blockDim.x;

  // stage the input band (rows 2*qy0-pc ..., all columns -pc .. 2*QW-1+pc)
  {
    const int Hp = a.H + 2 * a.in_pad, Wp = a.W + 2 * a.in_pad;
    const float* inb = a.in + (size_t)b * Hp * Wp * CK;
    if ((CK & 3) == 0) {
      const int c4n = CK >> 2;
      const int total = g.nrows * g.BW * c4n;
      for (int i = tid; i < total; i += nt) {
        const int c4 = i % c4n, col = (i / c4n) % g.BW, r = i / (c4n * g.BW);
        const int yp = 2 * g.qy0 - pc + r + a.in_pad, xp = col - pc + a.in_pad;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (yp >= 0 && yp < Hp && xp >= 0 && xp < Wp)
          v = __ldg(reinterpret_cast<const float4*>(inb + ((size_t)yp * Wp + xp) * CK + 4 * c4));
        *reinterpret_cast<float4*>(band + ((size_t)r * g.BW + col) * CKp + 4 * c4) = v;
      }
    } else {
      const int total = g.nrows * g.BW * CK;
      for (int i = tid; i < total; i += nt) {
        const int c = i % CK, col = (i / CK) % g.BW, r = i / (CK * g.BW);
        const int yp = 2 * g.qy0 - pc + r + a.in_pad, xp = col - pc + a.in_pad;
        float v = 0.f;
        if (yp >= 0 && yp < Hp && xp >= 0 && xp < Wp) v = __ldg(inb + ((size_t)yp * Wp + xp) * CK + c);
        band[((size_t)r * g.BW + col) * CKp + c] = v;
      }
    }
  }

  const int G = a.Cout >> 3;
  const int cg = tid % G, ql = tid / G;
  const int q = g.q0 + ql;
  const bool valid = q <= g.q1;
  const int qy = q / g.QW, qx = q % g.QW;
  float acc[4][8];
#pragma unroll
  for (int p = 0; p < 4; ++p)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[p][j] = 0.f;

  conv_core<KS>(band, g.BW, CK, CKp, a.wf, w_s, a.Cout, 2 * (qy - g.qy0), 2 * qx, 4 * cg, valid, acc);
  if (!valid) return;

  const int Ch = a.Cout >> 1, c0 = 4 * cg;
  float m[4][4];
  unsigned hf[4][4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float bl = __ldg(a.bias + c0 + j), bh = __ldg(a.bias + Ch + c0 + j);
#pragma unroll
    for (int p = 0; p < 4; ++p) {
      const float lo = acc[p][j] + bl, hi = acc[p][4 + j] + bh;
      const bool h = hi > lo;
      m[p][j] = h ? hi : lo;
      hf[p][j] = h ? 4u : 0u;
    }
  }
  float mean[4] = {0.f, 0.f, 0.f, 0.f}, inv[4] = {1.f, 1.f, 1.f, 1.f};
  if (a.bn_mean != nullptr) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      mean[j] = __ldg(a.bn_mean + c0 + j);
      inv[j] = __ldg(a.bn_invstd + c0 + j);
    }
  }
  const int Hop = a.Ho + 2 * a.out_pad, Wop = a.Wo + 2 * a.out_pad;
  if (POOL) {
    float v[4];
    unsigned char cd[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float best = m[0][j];
      unsigned code = hf[0][j];
#pragma unroll
      for (int p = 1; p < 4; ++p)
        if (m[p][j] > best) {
          best = m[p][j];
          code = hf[p][j] | (unsigned)p;
        }
      v[j] = (best - mean[j]) * inv[j];
      cd[j] = (unsigned char)code;
    }
    float* o = a.out + (((size_t)b * Hop + qy + a.out_pad) * Wop + qx + a.out_pad) * Ch + c0;
    *reinterpret_cast<float4*>(o) = make_float4(v[0], v[1], v[2], v[3]);
    *reinterpret_cast<uchar4*>(a.codes + (((size_t)b * a.Ho + qy) * a.Wo + qx) * Ch + c0) =
        make_uchar4(cd[0], cd[1], cd[2], cd[3]);
  } else {
#pragma unroll
    for (int p = 0; p < 4; ++p) {
      const int y = 2 * qy + (p >> 1), x = 2 * qx + (p & 1);
      if (y >= a.H || x >= a.W) continue;
      float* o = a.out + (((size_t)b * Hop + y + a.out_pad) * Wop + x + a.out_pad) * Ch + c0;
      *reinterpret_cast<float4*>(o) = make_float4((m[p][0] - mean[0]) * inv[0], (m[p][1] - mean[1]) * inv[1],
                                                  (m[p][2] - mean[2]) * inv[2], (m[p][3] - mean[3]) * inv[3]);
      *reinterpret_cast<uchar4*>(a.codes + (((size_t)b * a.Ho + y) * a.Wo + x) * Ch + c0) =
          make_uchar4((unsigned char)hf[p][0], (unsigned char)hf[p][1], (unsigned char)hf[p][2],
                      (unsigned char)hf[p][3]);
    }
  }
}

template <int KS, bool POOL>
__global__ void __launch_bounds__(256) conv_bwd_kernel(ConvBwdArgs a) {
  extern __shared__ __align__(16) float smem[];
  constexpr int pc = KS / 2;
  const TileGeom g = tile_geom(a.H, a.W, false, pc);
  const int CK = a.Cout, CKp = CK + 4, Ch = CK >> 1;
  float* band = smem;
  float* w_s = smem + a.band_floats;
  const int b = blockIdx.y, tid = threadIdx.x, nt = blockDim.x;

  // stage the expanded gradient of the conv output: BN scale, un-pool and un-MFM applied on the fly
  {
    const int c4n = CK >> 2;
    const int total = g.nrows * g.BW * c4n;
    for (int i = tid; i < total; i += nt) {
      const int c4 = i % c4n, col = (i / c4n) % g.BW, r = i / (c4n * g.BW);
      const int y = 2 * g.qy0 - pc + r, x = col - pc;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (y >= 0 && y < a.H && x >= 0 && x < a.W) {
        const int py = POOL ? (y >> 1) : y, px = POOL ? (x >> 1) : x;
        if (py < a.Ho && px < a.Wo) {
          const int co = 4 * c4;
          const int half = co >= Ch ? 1 : 0;
          const int c = co - half * Ch;
          const size_t o = (((size_t)b * a.Ho + py) * a.Wo + px) * Ch + c;
          float4 gv = __ldg(reinterpret_cast<const float4*>(a.gout + o));
          const uchar4 cd = *reinterpret_cast<const uchar4*>(a.codes + o);
          const unsigned want = (POOL ? (unsigned)(((y & 1) << 1) | (x & 1)) : 0u) | (unsigned)(half << 2);
          if (a.bn_invstd != nullptr) {
            gv.x *= __ldg(a.bn_invstd + c + 0);
            gv.y *= __ldg(a.bn_invstd + c + 1);
            gv.z *= __ldg(a.bn_invstd + c + 2);
            gv.w *= __ldg(a.bn_invstd + c + 3);
          }
          v.x = cd.x == want ? gv.x : 0.f;
          v.y = cd.y == want ? gv.y : 0.f;
          v.z = cd.z == want ? gv.z : 0.f;
          v.w = cd.w == want ? gv.w : 0.f;
        }
      }
      *reinterpret_cast<float4*>(band + ((size_t)r * g.BW + col) * CKp + 4 * c4) = v;
    }
  }

  const int N = a.Cin;
  const int G = N >> 3;
  const int cg = tid % G, ql = tid / G;
  const int q = g.q0 + ql;
  const bool valid = q <= g.q1;
  const int qy = q / g.QW, qx = q % g.QW;
  float acc[4][8];
#pragma unroll
  for (int p = 0; p < 4; ++p)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[p][j] = 0.f;

  conv_core<KS>(band, g.BW, CK, CKp, a.wd, w_s, N, 2 * (qy - g.qy0), 2 * qx, 4 * cg, valid, acc);
  if (!valid) return;
  const int nh = N >> 1, c0 = 4 * cg;
#pragma unroll
  for (int p = 0; p < 4; ++p) {
    const int y = 2 * qy + (p >> 1), x = 2 * qx + (p & 1);
    if (y >= a.H || x >= a.W) continue;
    float* o = a.gin + (((size_t)b * a.H + y) * a.W + x) * N;
    *reinterpret_cast<float4*>(o + c0) = make_float4(acc[p][0], acc[p][1], acc[p][2], acc[p][3]);
    *reinterpret_cast<float4*>(o + nh + c0) = make_float4(acc[p][4], acc[p][5], acc[p][6], acc[p][7]);
  }
}

// Backward of the first block (1 input channel, 5x5, MFM, pool): gather over the <= 3x3 pooled cells whose winner
// can fall inside the 5x5 window of this input pixel.  Deterministic (no atomics).
constexpr int C0_TILE = 16;
constexpr int C0_CELLS = C0_TILE / 2 + 3;  // pooled cells per side needed by a 16-pixel tile with a 2-pixel reach
__global__ void __launch_bounds__(C0_TILE* C0_TILE) conv0_bwd_kernel(const float* __restrict__ gout,
                                                                      const unsigned char* __restrict__ codes,
                                                                      const float* __restrict__ w0, float* gin, int H,
                                                                      int W, int Ho, int Wo) {
  __shared__ float s_g[C0_CELLS][C0_CELLS][32];
  __shared__ unsigned char s_cd[C0_CELLS][C0_CELLS][32];
  __shared__ float s_w[64 * 25];
  const int tid = threadIdx.y * C0_TILE + threadIdx.x;
  const int b = blockIdx.z, y0 = blockIdx.y * C0_TILE, x0 = blockIdx.x * C0_TILE;
  const int py0 = (y0 - 2) >> 1, px0 = (x0 - 2) >> 1;  // arithmetic shift: floor for negatives
  for (int i = tid; i < 64 * 25; i += C0_TILE * C0_TILE) s_w[i] = w0[i];
  for (int i = tid; i < C0_CELLS * C0_CELLS * 32; i += C0_TILE * C0_TILE) {
    const int c = i & 31, cx = (i >> 5) % C0_CELLS, cy = (i >> 5) / C0_CELLS;
    const int py = py0 + cy, px = px0 + cx;
    float gv = 0.f;
    unsigned char cd = 0;
    if (py >= 0 && py < Ho && px >= 0 && px < Wo) {
      const size_t o = (((size_t)b * Ho + py) * Wo + px) * 32 + c;
      gv = gout[o];
      cd = codes[o];
    }
    s_g[cy][cx][c] = gv;
    s_cd[cy][cx][c] = cd;
  }
  __syncthreads();
  const int y = y0 + threadIdx.y, x = x0 + threadIdx.x;
  if (y >= H || x >= W) return;
  float acc = 0.f;
  const int cy_lo = ((y - 2) >> 1) - py0, cy_hi = ((y + 2) >> 1) - py0;
  const int cx_lo = ((x - 2) >> 1) - px0, cx_hi = ((x + 2) >> 1) - px0;
  for (int cy = cy_lo; cy <= cy_hi; ++cy)
    for (int cx = cx_lo; cx <= cx_hi; ++cx) {
      const int wy0 = 2 * (py0 + cy), wx0 = 2 * (px0 + cx);
#pragma unroll 4
      for (int c = 0; c < 32; ++c) {
        const unsigned cd = s_cd[cy][cx][c];
        const int dy = y - (wy0 + (int)((cd >> 1) & 1u)) + 2;
        const int dx = x - (wx0 + (int)(cd & 1u)) + 2;
        if (dy >= 0 && dy < 5 && dx >= 0 && dx < 5) {
          const int co = c + ((cd >> 2) & 1u) * 32;
          acc = fmaf(s_g[cy][cx][c], s_w[co * 25 + dy * 5 + dx], acc);
        }
      }
    }
  gin[((size_t)b * H + y) * W + x] = acc;
}

__global__ void pack_conv_kernel(const float* __restrict__ w, float* wf, float* wd, int Cout, int Cin, int KS) {
  const int n = Cout * Cin * KS * KS;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int dx = i % KS, dy = (i / KS) % KS, ci = (i / (KS * KS)) % Cin, co = i / (KS * KS * Cin);
    const float v = w[i];
    wf[((size_t)(dy * KS + dx) * Cin + ci) * Cout + co] = v;
    if (wd != nullptr) {
      const int ty = KS - 1 - dy, tx = KS - 1 - dx;
      wd[((size_t)(ty * KS + tx) * Cout + co) * Cin + ci] = v;
    }
  }
}

__global__ void bn_prepare_kernel(const float* __restrict__ var, float* invstd, int C, float eps) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < C) invstd[i] = 1.0f / sqrtf(var[i] + eps);
}

template <typename K>
int set_smem(K kernel, size_t bytes) {
  ADVB_CHECK(bytes <= 227 * 1024, "conv tile does not fit shared memory");
  ADVB_CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  return 0;
}

}  // namespace

int conv_pack_weights(const float* w, float* wf, float* wd, int Cout, int Cin, int KS, cudaStream_t stream) {
  const int n = Cout * Cin * KS * KS;
  pack_conv_kernel<<<cdiv(n, 256), 256, 0, stream>>>(w, wf, wd, Cout, Cin, KS);
  ADVB_KERNEL_OK("pack_conv", stream);
  return 0;
}

int bn_prepare(const float* var, float* invstd, int C, cudaStream_t stream) {
  bn_prepare_kernel<<<cdiv(C, 128), 128, 0, stream>>>(var, invstd, C, 1e-5f);
  ADVB_KERNEL_OK("bn_prepare", stream);
  return 0;
}

int conv_mfm_forward(ConvFwdArgs a, cudaStream_t stream) {
  ADVB_CHECK(a.Cout % 8 == 0 && a.Cout <= 128, "Cout must be a multiple of 8, <= 128");
  ADVB_CHECK(a.Cin == 1 || a.Cin % 4 == 0, "Cin must be 1 or a multiple of 4");
  const int pc = a.KS / 2;
  const int QW = a.pool ? a.W / 2 : (a.W + 1) / 2, QH = a.pool ? a.H / 2 : (a.H + 1) / 2;
  ADVB_CHECK(QW > 0 && QH > 0, "empty conv output");
  const int CKp = (a.Cin % 4 == 0) ? a.Cin + 4 : a.Cin;
  int band = band_rows_max(QH, QW, pc) * (2 * QW + 2 * pc) * CKp;
  band = (band + 3) & ~3;
  a.band_floats = band;
  const size_t smem = (size_t)(band + a.Cin * a.Cout) * sizeof(float);
  dim3 grid(cdiv(QH * QW, 32), a.B);
  const int threads = 32 * (a.Cout / 8);
#define ADVB_FWD(KS_, POOL_)                                                        \
  do {                                                                              \
    ADVB_TRY(set_smem(conv_fwd_kernel<KS_, POOL_>, smem));                          \
    conv_fwd_kernel<KS_, POOL_><<<grid, threads, smem, stream>>>(a);                \
  } while (0)
  if (a.KS == 1 && !a.pool) ADVB_FWD(1, false);
  else if (a.KS == 1 && a.pool) ADVB_FWD(1, true);
  else if (a.KS == 3 && !a.pool) ADVB_FWD(3, false);
  else if (a.KS == 3 && a.pool) ADVB_FWD(3, true);
  else if (a.KS == 5 && a.pool) ADVB_FWD(5, true);
  else if (a.KS == 5 && !a.pool) ADVB_FWD(5, false);
  else {
    set_error("unsupported conv kernel size");
    return 1;
  }
#undef ADVB_FWD
  ADVB_KERNEL_OK(a.tag, stream);
  return 0;
}

int conv_mfm_backward(ConvBwdArgs a, cudaStream_t stream) {
  ADVB_CHECK(a.Cin % 8 == 0 && a.Cin <= 64, "backward needs Cin multiple of 8, <= 64");
  ADVB_CHECK(a.Cout % 8 == 0, "Cout must be a multiple of 8");
  const int pc = a.KS / 2;
  const int QW = (a.W + 1) / 2, QH = (a.H + 1) / 2;
  int band = band_rows_max(QH, QW, pc) * (2 * QW + 2 * pc) * (a.Cout + 4);
  band = (band + 3) & ~3;
  a.band_floats = band;
  const size_t smem = (size_t)(band + a.Cin * a.Cout) * sizeof(float);
  dim3 grid(cdiv(QH * QW, 32), a.B);
  const int threads = 32 * (a.Cin / 8);
#define ADVB_BWD(KS_, POOL_)                                                        \
  do {                                                                              \
    ADVB_TRY(set_smem(conv_bwd_kernel<KS_, POOL_>, smem));                          \
    conv_bwd_kernel<KS_, POOL_><<<grid, threads, smem, stream>>>(a);                \
  } while (0)
  if (a.KS == 1 && !a.pool) ADVB_BWD(1, false);
  else if (a.KS == 1 && a.pool) ADVB_BWD(1, true);
  else if (a.KS == 3 && !a.pool) ADVB_BWD(3, false);
  else if (a.KS == 3 && a.pool) ADVB_BWD(3, true);
  else {
    set_error("unsupported conv kernel size (backward)");
    return 1;
  }
#undef ADVB_BWD
  ADVB_KERNEL_OK(a.tag, stream);
  return 0;
}

int conv0_backward(const float* gout, const unsigned char* codes, const float* w0, float* gin, int B, int H, int W,
                   int Ho, int Wo, cudaStream_t stream) {
  dim3 grid(cdiv(W, C0_TILE), cdiv(H, C0_TILE), B);
  dim3 block(C0_TILE, C0_TILE);
  conv0_bwd_kernel<<<grid, block, 0, stream>>>(gout, codes, w0, gin, H, W, Ho, Wo);
  ADVB_KERNEL_OK("conv0_bwd", stream);
  return 0;
}

}  // namespace advb
