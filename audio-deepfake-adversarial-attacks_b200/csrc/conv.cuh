// LCNN conv-block host entry points — see conv.cu.
#pragma once
#include "common.cuh"

namespace advb {

struct ConvFwdArgs {
  const float* in;   // (B, H+2*in_pad, W+2*in_pad, Cin) zero-bordered NHWC
  int in_pad;
  const float* wf;   // packed [tap][ci][co]
  const float* bias; // (Cout)
  const float* bn_mean;    // (Cout/2) or null
  const float* bn_invstd;  // (Cout/2) or null
  float* out;        // (B, Ho+2*out_pad, Wo+2*out_pad, Cout/2)
  int out_pad;
  unsigned char* codes;  // (B, Ho, Wo, Cout/2): pool arg-max position (bits 0-1) | MFM half (bit 2)
  int B, H, W, Cin, Cout, KS;
  int Ho, Wo;
  bool pool;
  int band_floats;  // filled by the launcher
  const char* tag;  // profiler label
};

struct ConvBwdArgs {
  const float* gout;           // (B, Ho, Wo, Cout/2) compact gradient of the block output
  const unsigned char* codes;  // as written by the forward
  const float* bn_invstd;      // (Cout/2) or null
  const float* wd;             // packed [flipped tap][co][ci]
  float* gin;                  // (B, H, W, Cin) compact gradient of the block input
  int B, H, W, Cin, Cout, KS;
  int Ho, Wo;
  bool pool;
  int band_floats;
  const char* tag;
};

int conv_pack_weights(const float* w, float* wf, float* wd, int Cout, int Cin, int KS, cudaStream_t stream);
int bn_prepare(const float* var, float* invstd, int C, cudaStream_t stream);
int conv_mfm_forward(ConvFwdArgs a, cudaStream_t stream);
int conv_mfm_backward(ConvBwdArgs a, cudaStream_t stream);
// first block (Cin = 1, 5x5, pool, no BN): w0 is the original (64,1,5,5) weight
int conv0_backward(const float* gout, const unsigned char* codes, const float* w0, float* gin, int B, int H, int W,
                   int Ho, int Wo, cudaStream_t stream);

// first block backward, production kernel (conv0_bwd.cu): fp32, thread per pooled cell, fused col2im; LCNN geometry only
bool conv0_cells_supported(int H, int W, int Ho, int Wo);
int conv0_cells_backward(const float* gout, const unsigned char* codes, const float* w0, float* gin, int B, int H, int W,
                         int Ho, int Wo, cudaStream_t stream);

// ---- tensor-core (tcgen05) path, conv_tc.cu ----
bool conv_tc_supported(int Cin, int Cout, int KS, bool pool);
size_t conv_tc_pack_bytes(int Cout, int Cin, int KS, bool bwd);
// wf / wd: packed forward / backward weight slices (wd may be null)
int conv_tc_pack(const float* w, unsigned char* wf, unsigned char* wd, int Cout, int Cin, int KS, cudaStream_t stream);
// same images with N padded to npad_f / npad_b rows (zero rows), e.g. (Cout, Cin) = (20, 20) -> MMA N = 32 both ways
int conv_tc_pack_padded(const float* w, unsigned char* wf, unsigned char* wd, int Cout, int Cin, int KS, int npad_f, int npad_b,
                        cudaStream_t stream);
size_t conv_tc_pack_bytes_padded(int Cout, int Cin, int KS, bool bwd, int npad);
// passes: 3 = 3xTF32 (fp32-class accuracy, default), 1 = single-pass tf32
int conv_tc_forward(const ConvFwdArgs& a, const unsigned char* wpack, int passes, cudaStream_t stream);
int conv_tc_backward(const ConvBwdArgs& a, const unsigned char* wpack, int passes, cudaStream_t stream);
// first block backward on tensor cores: T (B,H,W,5) scratch, gin (B,H,W) = d loss / d cepstral image
int conv0_tc_backward(const float* gout, const unsigned char* codes, const unsigned char* wpack, float* T, float* gin,
                      int B, int H, int W, int Ho, int Wo, int passes, cudaStream_t stream);

// ---- persistent pipelined kernels for the first block and the 1x1 blocks, conv_light.cu (same packed weights) ----
bool conv_light_supported(int Cin, int Cout, int KS, bool pool, bool bwd);
int conv_light_forward(const ConvFwdArgs& a, const unsigned char* wpack, int passes, cudaStream_t stream);
int conv_light_backward(const ConvBwdArgs& a, const unsigned char* wpack, int passes, cudaStream_t stream);
int conv0_light_backward_gemm(const float* gout, const unsigned char* codes, const unsigned char* wpack, float* T, int B,
                              int H, int W, int Ho, int Wo, int passes, cudaStream_t stream);

// ---- first block forward as a Toeplitz GEMM without im2col, conv0_toeplitz.cu (its own weight image) ----
bool conv0t_supported(int H, int W, int Ho, int Wo);
size_t conv0t_pack_bytes();
int conv0t_pack(const float* w0, const float* bias, unsigned char* wpack, cudaStream_t stream);
int conv0t_forward(const ConvFwdArgs& a, const unsigned char* wpack, int passes, cudaStream_t stream);

// ---- persistent warp-specialised kernels for the 3x3 blocks, conv_p3.cu (same packed weights) ----
bool conv_p3_supported(int Cin, int Cout, int KS, bool pool, int W);
int conv_p3_forward(const ConvFwdArgs& a, const unsigned char* wpack, int passes, cudaStream_t stream);
int conv_p3_backward(const ConvBwdArgs& a, const unsigned char* wpack, int passes, cudaStream_t stream);
// Plain 3x3 convolution (pad 1, stride 1) on the same persistent tcgen05 kernel (conv_p3.cu, PLAIN variant), 3xTF32.
//   in  (B, H+2, W+2, Cin) with a zero border;  wpack = a forward image of conv_tc_pack(_padded) - or its backward image, which
//   makes this the transposed convolution;  out (B, H+2p, W+2p, Cout), interior only:
//   v = acc + bias;  [v = lrelu(v * aff_scale[c] + aff_shift[c], act_slope)];  [v = v * (mul_h > 0 ? 1 : mul_slope) * mul_scale[c]]
//   with mul_h a (B, H+2, W+2, Cout) bordered tensor.  Shapes: see conv_p3_plain_supported.
struct P3Plain {
  const float* in = nullptr;
  float* out = nullptr;
  int out_pad = 0;
  const unsigned char* wpack = nullptr;
  const float* bias = nullptr;
  const float* aff_scale = nullptr;
  const float* aff_shift = nullptr;
  float act_slope = 1.f;
  const float* mul_h = nullptr;
  const float* mul_scale = nullptr;
  float mul_slope = 1.f;
  int B = 0, H = 0, W = 0, Cin = 0, Cout = 0, passes = 3;
  const char* tag = "conv_p3_plain";
};
int conv_p3_plain_forward(const P3Plain& p, cudaStream_t stream);
bool conv_p3_plain_supported(int Cin, int Cout, int W);
// mixed mode (passes = 2, forward only): tf32 main term + both cross terms as ONE bf16 MMA; its own weight image (same bytes)
int conv_p3_pack_mix(const float* w, unsigned char* wf, int Cout, int Cin, cudaStream_t stream);
// "horizontal scatter" backward (N = 3 C_in): needs its own weight image in the layer's backward pack buffer
bool conv_p3_bwd_hs(int Cin, int Cout, int KS, bool pool, int W);
int conv_p3_pack_hs(const float* w, unsigned char* wd, int Cout, int Cin, cudaStream_t stream);

}  // namespace advb
