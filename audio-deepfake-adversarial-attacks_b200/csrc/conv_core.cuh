// Shared SIMT convolution core (fp32): a CTA owns 32 consecutive 2x2 "quads" of output pixels of one clip and all
// output channels; the full-width band of input rows is staged in shared memory once and one tap's weight slab is
// streamed at a time.  A thread owns one quad and 8 output channels {4g..4g+3} u {N/2+4g..N/2+4g+3}.  Used by the LCNN
// fp32 cross-check path (conv.cu) and by the SpecRNet blocks (specrnet.cu).
#pragma once
#include "common.cuh"

namespace advb {
namespace convcore {

__device__ __forceinline__ void fma4(float (&acc)[8], int off, float a, const float4& w) {
  acc[off + 0] = fmaf(a, w.x, acc[off + 0]);
  acc[off + 1] = fmaf(a, w.y, acc[off + 1]);
  acc[off + 2] = fmaf(a, w.z, acc[off + 2]);
  acc[off + 3] = fmaf(a, w.w, acc[off + 3]);
}

// acc[p][0..3] -> channels g4..g4+3, acc[p][4..7] -> channels nh+g4..nh+g4+3, p = 2x2 pixel of the quad.
// ck_loop (optional): contraction channels actually visited, <= CK, when the staged layout is zero-padded beyond them.
// slab_all: w_s holds all KS*KS weight slabs (the caller sized it so): they are staged once, before the tap loop, instead of
// one slab per tap between two block-wide barriers - for small layers the 9 exposed L2 round trips and 18 barriers per
// tile cost about as much as the FMAs between them.
template <int KS>
__device__ __forceinline__ void conv_core(const float* band, int BW, int CK, int CKp, const float* __restrict__ wg,
                                          float* w_s, int N, int rb, int cb, int g4, bool valid, float (&acc)[4][8],
                                          int ck_loop = 0, bool slab_all = false) {
  const int nh = N >> 1;
  const int CKl = ck_loop > 0 ? ck_loop : CK;
  const int slab4 = (CK * N) >> 2;
  if (slab_all) {
    const float4* src = reinterpret_cast<const float4*>(wg);
    float4* dst = reinterpret_cast<float4*>(w_s);
    for (int i = threadIdx.x; i < KS * KS * slab4; i += blockDim.x) dst[i] = __ldg(src + i);
    __syncthreads();  // also orders the caller's band stores before the first read
  }
#pragma unroll 1
  for (int tap = 0; tap < KS * KS; ++tap) {
    const int dy = tap / KS, dx = tap % KS;
    const float* w_t = w_s;
    if (slab_all) {
      w_t = w_s + (size_t)tap * CK * N;
    } else {
      __syncthreads();
      {
        const float4* src = reinterpret_cast<const float4*>(wg + (size_t)tap * CK * N);
        float4* dst = reinterpret_cast<float4*>(w_s);
        for (int i = threadIdx.x; i < slab4; i += blockDim.x) dst[i] = __ldg(src + i);
      }
      __syncthreads();
    }
    if (!valid) continue;
    const float* p00 = band + ((size_t)(rb + dy) * BW + cb + dx) * CKp;
    const float* p01 = p00 + CKp;
    const float* p10 = p00 + (size_t)BW * CKp;
    const float* p11 = p10 + CKp;
    if ((CK & 3) == 0) {
#pragma unroll 1
      for (int ci = 0; ci < CKl; ci += 4) {
        const float4 a0 = *reinterpret_cast<const float4*>(p00 + ci);
        const float4 a1 = *reinterpret_cast<const float4*>(p01 + ci);
        const float4 a2 = *reinterpret_cast<const float4*>(p10 + ci);
        const float4 a3 = *reinterpret_cast<const float4*>(p11 + ci);
        const float av[4][4] = {{a0.x, a0.y, a0.z, a0.w}, {a1.x, a1.y, a1.z, a1.w}, {a2.x, a2.y, a2.z, a2.w},
                                {a3.x, a3.y, a3.z, a3.w}};
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const float4 wl = *reinterpret_cast<const float4*>(w_t + (ci + u) * N + g4);
          const float4 wh = *reinterpret_cast<const float4*>(w_t + (ci + u) * N + nh + g4);
#pragma unroll
          for (int p = 0; p < 4; ++p) {
            fma4(acc[p], 0, av[p][u], wl);
            fma4(acc[p], 4, av[p][u], wh);
          }
        }
      }
    } else {
#pragma unroll 1
      for (int ci = 0; ci < CKl; ++ci) {
        const float av[4] = {p00[ci], p01[ci], p10[ci], p11[ci]};
        const float4 wl = *reinterpret_cast<const float4*>(w_t + ci * N + g4);
        const float4 wh = *reinterpret_cast<const float4*>(w_t + ci * N + nh + g4);
#pragma unroll
        for (int p = 0; p < 4; ++p) {
          fma4(acc[p], 0, av[p], wl);
          fma4(acc[p], 4, av[p], wh);
        }
      }
    }
  }
}

struct TileGeom {
  int QW, QH, nQ, q0, q1, qy0, nrows, BW;
};

// qpc = quads per CTA (32 on the LCNN cross-check path; the SpecRNet launcher picks whole quad rows, specrnet.cu)
__device__ __forceinline__ TileGeom tile_geom(int H, int W, bool floor_quads, int pc, int qpc = 32) {
  TileGeom g;
  g.QW = floor_quads ? W / 2 : (W + 1) / 2;
  g.QH = floor_quads ? H / 2 : (H + 1) / 2;
  g.nQ = g.QH * g.QW;
  g.q0 = blockIdx.x * qpc;
  g.q1 = min(g.q0 + qpc, g.nQ) - 1;
  g.qy0 = g.q0 / g.QW;
  const int qy1 = g.q1 / g.QW;
  g.nrows = 2 * (qy1 - g.qy0 + 1) + 2 * pc;
  g.BW = 2 * g.QW + 2 * pc;
  return g;
}


// rows of the band a CTA of qpc quads may need (host side)
inline int band_rows_max(int QH, int QW, int pc, int qpc = 32) {
  int span = (qpc + QW - 2) / QW + 1;
  if (span > QH) span = QH;
  return 2 * span + 2 * pc;
}

}  // namespace convcore
}  // namespace advb
