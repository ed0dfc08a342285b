// Backward of LCNN's first block (Conv2d 1 -> 64, 5x5 | Max-Feature-Map | 2x2 max-pool; src/models/lcnn.py:120-126)
// as an fp32 SIMT kernel that uses the structure of the gradient (sm_100a).
//
// The stage gradient lives on the pooled grid (Ho x Wo x 32).  Un-pooling and un-MFM route every value to ONE of the
// 4 pixels of its cell and ONE of the two MFM halves, so the 64-channel pixel gradient the transposed convolution
// contracts over is 7/8 zeros.  The tensor-core version (conv_light.cu, K = 64 dense per pixel, then a two-stage
// col2im through a (B,H,W,5) scratch) spent its time materialising those zeros: 5 200 worker cycles per 124-pixel
// tile, 0.42 + 0.04 ms per launch at B = 128.  Here
//   phase 1  thread = pooled cell: Z[p][tap] += (code_c.p == p ? g_c : 0) * W[code_c.h * 32 + c][tap] for its 4 pixels
//            p and 25 taps - 100 independent FMA chains in registers, weights read as 7 LDS.128 per channel (a warp
//            touches the two rows h = 0 / 1, kept 16 banks apart), no branches;
//   phase 2  Z goes to shared memory (pixel-major, 25 floats per pixel: an odd stride, conflict-free for phase 3);
//   phase 3  thread = input pixel: gin[y][x] = sum_{dy,dx} Z[y - dy + 2][x - dx + 2][5 dy + dx] in a fixed order
//            (deterministic, no atomics), coalesced store.
// A CTA owns a strip of 7 cell rows (14 pixel rows, full width) of one clip and recomputes one cell row above and
// below it (the 2-pixel reach of the 5x5 filter), persistent over (clip, strip).  3 200 FMAs per cell: 4.3 GFMA per
// launch at 128 clips including the halo rows, i.e. 0.12 ms at the fp32 peak; measured 0.29 ms (the phases of a CTA
// do not overlap and 12 warps per SM leave LDS latency exposed) against 0.46 ms for the tensor-core version.
#include "conv.cuh"

#include <stdlib.h>

namespace advb {

namespace {

constexpr int Z_WO = 40, Z_W = 80; // LCNN: 80 cepstral coefficients -> 40 pooled columns
constexpr int Z_LDW = 28;                    // padded weight row (25 taps)
constexpr int Z_WH = 32 * Z_LDW + 16;        // offset of the h = 1 rows: 16 banks away from the h = 0 row of the same channel

// Z_R = cell rows per strip.  Registers: an SM sub-partition holds 16 384, so 13 warps (Z_R = 8: 400 threads, 4 warps on one
// sub-partition) cap a thread at 128 registers, 12 warps (Z_R = 7: 360 threads) at 168.
template <int Z_R>
__global__ void __launch_bounds__((Z_R + 2) * Z_WO, 1) conv0_bwd_cells_kernel(const float* __restrict__ gout,
                                                                        const unsigned char* __restrict__ codes,
                                                                        const float* __restrict__ w0,
                                                                        float* __restrict__ gin, int B, int H, int Ho,
                                                                        int n_strips) {
  constexpr int Z_THREADS = (Z_R + 2) * Z_WO;  // one thread per cell incl. the two halo rows
  constexpr int Z_PROWS = 2 * (Z_R + 2);       // pixel rows of Z in shared memory
  extern __shared__ __align__(16) float smem[];
  float* s_w = smem;                 // s_w[h * Z_WH + c * Z_LDW + t] = W[32 h + c][t]
  float* s_z = s_w + 2 * Z_WH;       // Z_PROWS x Z_W x 25
  const int tid = threadIdx.x;
  for (int i = tid; i < 64 * Z_LDW; i += Z_THREADS) {
    const int co = i / Z_LDW, t = i - co * Z_LDW;
    s_w[(co >> 5) * Z_WH + (co & 31) * Z_LDW + t] = t < 25 ? __ldg(w0 + co * 25 + t) : 0.f;
  }
  const int cyl = tid / Z_WO, cx = tid - cyl * Z_WO;

  for (int work = blockIdx.x; work < B * n_strips; work += gridDim.x) {
    const int b = work / n_strips, strip = work - b * n_strips;
    const int py = strip * Z_R - 1 + cyl;  // this thread's cell row; rows outside [0, Ho) contribute zeros
    const bool cell_ok = py >= 0 && py < Ho;
    const size_t cell = (((size_t)b * Ho + (cell_ok ? py : 0)) * Z_WO + cx) * 32;
    __syncthreads();  // phase 3 of the previous item has finished reading s_z (and s_w is complete on the first pass)

    // ---- phase 1 ----
    float acc[4][25];
#pragma unroll
    for (int q = 0; q < 4; ++q)
#pragma unroll
      for (int t = 0; t < 25; ++t) acc[q][t] = 0.f;
    if (cell_ok) {
#pragma unroll 2
      for (int c4 = 0; c4 < 8; ++c4) {
        const float4 g4 = __ldg(reinterpret_cast<const float4*>(gout + cell) + c4);
        const unsigned cd4 = __ldg(reinterpret_cast<const unsigned*>(codes + cell) + c4);
        const float gv[4] = {g4.x, g4.y, g4.z, g4.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const unsigned cd = (cd4 >> (8 * j)) & 0xffu;  // (h << 2) | (dy << 1) | dx
          // the lanes of a warp read one of two rows (h = 0 / 1 of this channel); the h = 1 block is shifted by 16 banks
          // so that the two 16-byte accesses never collide (with both blocks bank-aligned every LDS.128 was a 2-way
          // conflict and the shared-memory pipe, not the FMAs, set the pace)
          const float* wr = s_w + ((cd >> 2) & 1u) * Z_WH + (4 * c4 + j) * Z_LDW;
          float gp[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) gp[q] = (cd & 3u) == (unsigned)q ? gv[j] : 0.f;
#pragma unroll
          for (int t4 = 0; t4 < 7; ++t4) {
            const float4 w = *reinterpret_cast<const float4*>(wr + 4 * t4);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              acc[q][4 * t4] = fmaf(gp[q], w.x, acc[q][4 * t4]);
              if (t4 < 6) {
                acc[q][4 * t4 + 1] = fmaf(gp[q], w.y, acc[q][4 * t4 + 1]);
                acc[q][4 * t4 + 2] = fmaf(gp[q], w.z, acc[q][4 * t4 + 2]);
                acc[q][4 * t4 + 3] = fmaf(gp[q], w.w, acc[q][4 * t4 + 3]);
              }
            }
          }
        }
      }
    }
    // ---- phase 2 ----
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      float* z = s_z + ((2 * cyl + (q >> 1)) * Z_W + 2 * cx + (q & 1)) * 25;
#pragma unroll
      for (int t = 0; t < 25; ++t) z[t] = acc[q][t];
    }
    __syncthreads();

    // ---- phase 3 ----  pixel rows [16 strip, 16 strip + 16); the last strip also owns the odd last row of the image
    const int y0 = 2 * Z_R * strip;
    const int y1 = strip == n_strips - 1 ? H : min(H, y0 + 2 * Z_R);
    for (int i = tid; i < (y1 - y0) * Z_W; i += Z_THREADS) {
      const int yl = i / Z_W, x = i - yl * Z_W;
      float a = 0.f;
#pragma unroll
      for (int dy = 0; dy < 5; ++dy) {
        const int pr = yl + 4 - dy;  // local Z row of pixel row y - dy + 2 (local row 0 = pixel row y0 - 2)
        if (pr < Z_PROWS) {
#pragma unroll
          for (int dx = 0; dx < 5; ++dx) {
            const int xz = x - dx + 2;
            if (xz >= 0 && xz < Z_W) a += s_z[(pr * Z_W + xz) * 25 + 5 * dy + dx];
          }
        }
      }
      gin[((size_t)b * H + y0 + yl) * Z_W + x] = a;
    }
  }
}

size_t conv0_cells_smem(int rows) { return (size_t)(2 * Z_WH + 2 * (rows + 2) * Z_W * 25) * sizeof(float); }

}  // namespace

bool conv0_cells_supported(int H, int W, int Ho, int Wo) {
  // rows 2 Ho .. H-1 (at most one: H odd) are owned by the last strip, whose halo covers them
  return W == Z_W && Wo == Z_WO && H >= 2 * Ho && H <= 2 * Ho + 1;
}

int conv0_cells_backward(const float* gout, const unsigned char* codes, const float* w0, float* gin, int B, int H, int W,
                         int Ho, int Wo, cudaStream_t stream) {
  ADVB_CHECK(conv0_cells_supported(H, W, Ho, Wo), "conv0_cells_backward: unsupported first-block geometry");
  static const int rows = [] {
    const char* e = getenv("ADVB_C0_ROWS");
    const int v = e != nullptr ? atoi(e) : 7;  // measured at B = 128: 291 us (7 rows, 165 registers) vs 340 us (8 rows, 128)
    return v == 8 ? 8 : 7;
  }();
  int n_sm = 148, dev = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
  const int n_strips = cdiv(Ho, rows);
  const int grid = B * n_strips < n_sm ? B * n_strips : n_sm;
  if (rows == 7) {
    ADVB_CUDA_OK(cudaFuncSetAttribute(conv0_bwd_cells_kernel<7>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)conv0_cells_smem(7)));
    conv0_bwd_cells_kernel<7><<<grid, 9 * Z_WO, conv0_cells_smem(7), stream>>>(gout, codes, w0, gin, B, H, Ho, n_strips);
  } else {
    ADVB_CUDA_OK(cudaFuncSetAttribute(conv0_bwd_cells_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)conv0_cells_smem(8)));
    conv0_bwd_cells_kernel<8><<<grid, 10 * Z_WO, conv0_cells_smem(8), stream>>>(gout, codes, w0, gin, B, H, Ho, n_strips);
  }
  ADVB_KERNEL_OK("conv0_bwd_cells", stream);
  return 0;
}

}  // namespace advb
