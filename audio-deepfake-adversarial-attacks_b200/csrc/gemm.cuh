// Row-major GEMM with a fused per-element epilogue for RawNet3's Conv1d contractions (src/models/rawnet3.py:73-137,
// 242-274): activations are time-major [clip][time][channel], so every 1x1 Conv1d is out[r][n] = sum_k A[r][k] W[n][k]
// over the flattened (clip, time) rows r, and a dilated k=3 Conv1d is the same GEMM with three row-shifted copies of A
// ("taps"; rows are stored with `pad` zero rows on each side of every clip so that a shifted read never crosses clips).
//
// Two implementations of one contract: gemm_tc.cu (tcgen05 / TMEM, 3xTF32 error-compensated products = fp32-class
// accuracy; the product path) and the fp32 SIMT kernel in the same file (engine option conv_path = 1: the in-process
// cross-check, never the default).  There is no CPU path.
#pragma once
#include "common.cuh"

namespace advb {

// Strided view of a weight tensor: W(n, tap, k) = w[n * s_n + tap * s_tap + k * s_k]; 0 where n >= n_valid or k >= k_valid.
struct GemmW {
  const float* w = nullptr;
  long long s_n = 0, s_k = 0, s_tap = 0;
  int n_valid = 0, k_valid = 0;
};

struct GemmArgs {
  // ---- A operand: A(r, tap, k) = A[(r + shift[tap]) * lda + k]   (0 when the shifted row is outside [0, M)) ----
  const float* A = nullptr;
  int lda = 0;
  int M = 0;     // rows (clips * Tp)
  int K = 0;     // contraction length per tap (multiple of 32)
  int ntap = 1;
  int shift[3] = {0, 0, 0};
  // im2col_T > 0 (sinc layer): A(r, k) = A[clip * im2col_T + 10 * l + k] for r = clip * Tp + l, k < 251, 10 l + k < im2col_T
  int im2col_T = 0;
  // ---- B operand ----
  GemmW w;                                // SIMT path reads the live weights through this view
  const unsigned char* wpack = nullptr;   // tcgen05 path: image produced by gemm_pack() from the same view
  int N = 0;                              // multiple of 128
  int n_store = 0;                        // > 0: only columns n < n_store (a multiple of 16) are stored (N padded up from n_store)
  // ---- row space: r = clip * Tp + pad + t, t in [0, Tv) is a valid row; other rows are never stored ----
  int Tp = 1, pad = 0, Tv = 1;
  // ---- epilogue, in this order (every per-column vector is indexed by n in [0, N); every matrix pointer already points
  //      at column 0 of the GEMM's N range and is addressed [r * ld + n]) ----
  const float* bias = nullptr;            // v += bias[n]   (bias_per_clip: bias[clip * N + n])
  int bias_per_clip = 0;
  int relu = 0;                           // v = max(v, 0); mask_out[r][n] = v > 0
  unsigned char* mask_out = nullptr;
  int ld_mask = 0;
  const float* bn_scale = nullptr;        // v = v * bn_scale[n] + bn_shift[n]
  const float* bn_shift = nullptr;
  const float* add = nullptr;             // v += add[r][n]   (may alias out)
  int ld_add = 0;
  float* out = nullptr;                   // out[r][n] = gate ? v * gate_scale[n] * (gate_mask[r][n] != 0) : v
  int ldc = 0;
  const float* gate_scale = nullptr;
  const unsigned char* gate_mask = nullptr;
  int ld_gate = 0;
  float* out2 = nullptr;                  // out2[r][n] = (v + add2[r][n]) [* gate2_scale[n] * (gate2_mask[r][n] != 0)]
  int ld2 = 0;
  const float* add2 = nullptr;
  int ld_add2 = 0;
  const float* gate2_scale = nullptr;
  const unsigned char* gate2_mask = nullptr;
  int ld_gate2 = 0;
  const char* tag = "gemm";
  long long* prof = nullptr;  // diagnostics (ADVB_GEMM_PROF=1): clock64 phase counters of CTA 0, see gemm_tc.cu
};

// Bytes of the packed tcgen05 weight image of an (N, ntap, K) view.
size_t gemm_pack_bytes(int N, int K, int ntap);
// Pack the live weights (tf32 hi / lo split, SWIZZLE_128B rows, one contiguous slice per (N-tile, 32-wide K chunk)).
int gemm_pack(const GemmW& w, int N, int K, int ntap, unsigned char* dst, cudaStream_t stream);
// path: 0 = tcgen05 (needs a.wpack), 1 = fp32 SIMT.  passes: 3 = 3xTF32, 1 = single-pass tf32.
int gemm_run(const GemmArgs& a, int path, int passes, cudaStream_t stream);

}  // namespace advb
