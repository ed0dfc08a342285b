// SpecRNet host entry points — see specrnet.cu.
#pragma once
#include "common.cuh"

namespace advb {

// One residual block + its channel attention (src/models/specrnet.py:73-91,145-170).
// Internal layout: [frames][coeffs][channels] (the frontend's image transposed, like the LCNN path), channels padded
// to a multiple of 8 (20 -> 24) with zero weights.
struct SrBlock {
  int Cin, Cout;       // logical channels
  int Ci, C;           // padded channels of the block input / output
  int H, W;            // conv grid (frames, coeffs)
  int Hb, Wb;          // after the block's max-pool
  int Hn, Wn;          // after the attention max-pool
  bool downsample;
  int n_tiles;         // CTAs per clip of the conv2 kernel (partial channel sums)
  // live parameters (borrowed)
  const float *w1, *b1, *w2, *b2, *wds, *bds;
  const float *bn_w, *bn_b, *bn_rm, *bn_rv;
  const float *att_w, *att_b;
  // packed per call
  float *w1f, *w1d, *w2f, *w2d, *wdf, *wdd, *b1p, *b2p, *bdp, *bn_scale, *bn_shift;
  // activations / side state
  float* h;               // conv1 -> bn2 -> lrelu   (B, H+2, W+2, C)
  float* xb;              // block output after pool (B, Hb, Wb, C)
  unsigned char* code1;   // pool arg-max of xb
  float* psum;            // (B, n_tiles, C) partial channel sums of xb
  float* y;               // (B, C) attention
  float* xn;              // max-pool(xb * y + y)    (B, Hn+2p, Wn+2p, C), p = 1 except after the last block
  int xn_pad;
  unsigned char* code2;
  // gradients
  float* g_xn;            // (B, Hn, Wn, C)
  float* gadd;            // (B, C) broadcast term of the attention backward
  float* g_c1;            // (B, H, W, C) gradient at conv1's output (pre-BN)
  // conv2 on the tensor cores (round 2; 64 -> 64 channels at W <= 40 and 20 -> 20 (padded 24) at W <= 80, engine option conv_path = 0): the persistent
  // tcgen05 kernel of conv_p3.cu computes the plain convolution / its transpose, the fused SIMT kernel keeps only its epilogue
  bool tc2 = false;
  float* w2t = nullptr;            // conv2 weights with the 3x3 taps transposed (the engine's image is the reference's transposed)
  unsigned char* tcf2 = nullptr;   // conv_tc_pack images of w2t: forward ...
  unsigned char* tcd2 = nullptr;   // ... and flipped / channel-transposed (the transposed convolution as a forward one)
  float* c2 = nullptr;             // (B, H, W, C) conv2(h) without bias
  float* go = nullptr;             // (B, H+2, W+2, C) gradient at conv2's output, zero border
  // conv1 of the blocks after the first one (24 -> 64 and 64 -> 64) and its transpose, same kernel (needs tc2: the shortcut term of
  // the backward reads the materialised go)
  bool tc1 = false;
  float* w1t = nullptr;
  unsigned char* tcf1 = nullptr;
  unsigned char* tcd1 = nullptr;
  float* g_c1b = nullptr;          // (B, H+2, W+2, C) gradient at conv1's output with a zero border (the transposed conv's input)
};

struct SrGru {
  const float *w_ih[2][2], *w_hh[2][2], *b_ih[2][2], *b_hh[2][2];  // [layer][direction], live
  const float *fc1_w, *fc1_b, *fc2_w, *fc2_b, *bn_w, *bn_b, *bn_rm, *bn_rv;
  float *wihT[2][2], *whhT[2][2];  // packed transposes [K][192]
  float* v;                        // (128) fc2 . fc1 (the head is linear)
  float* bn_scale;                 // 64
  float* bn_shift;                 // 64
  float* gates;                    // (B, 2, 2, L, 4, 64): r, z, n, W_hn h + b_hn
  float* outs;                     // (B, 2, L, 128): layer outputs
  float* xin;                      // (B, L, 64): selu(bn(x)) fed to the GRU
};

int sr_conv2_tiles(int H, int W, int C);  // CTAs per clip of the conv2 kernel = rows of SrBlock::psum
int sr_pack_first_bn(const float* w, const float* b, const float* rm, const float* rv, float* bn4, cudaStream_t stream);
int sr_pack_block(SrBlock& k, cudaStream_t stream);
int sr_pack_gru(SrGru& g, cudaStream_t stream);
// x0 = selu(first_bn(feat)) in place on the bordered image (B, F+2, 82, 1)
int sr_input_forward(float* img, const float* bn4 /*w,b,rm,rv*/, int B, int H, int W, cudaStream_t stream);
int sr_block_forward(const SrBlock& k, const float* x, int B, const char* tag, cudaStream_t stream);
// g_xn (k.g_xn) -> g_x (compact (B,H,W,Ci)); first: additionally multiply by selu'(x0) * first_bn scale, Ci = 8 padded,
// only channel 0 is written to g_x (B,H,W)
int sr_block_backward(const SrBlock& k, const float* x, float* g_x, int B, bool first, const float* bn4,
                      const char* tag, cudaStream_t stream);
int sr_gru_forward(const SrGru& g, const float* xn /*(B,L,1,64)*/, float* logits, int B, int L, cudaStream_t stream);
// seed: mode 0 CE mean, 1 logit, 2 coef[b]; writes g wrt xn (B,L,64)
int sr_gru_backward(const SrGru& g, const float* xn, const float* logits, const long long* y, int mode, int n_global,
                    const float* coef, float* g_xn, int B, int L, cudaStream_t stream);

}  // namespace advb
