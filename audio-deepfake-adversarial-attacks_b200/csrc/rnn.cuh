// BLSTM / head host entry points — see rnn.cu.
#pragma once
#include "common.cuh"

namespace advb {

struct LstmWeights {  // live tensors of one nn.LSTM(160, 80, bidirectional=True): [0] forward, [1] reverse
  const float* w_ih[2];  // (320,160)
  const float* w_hh[2];  // (320,80)
  const float* b_ih[2];  // (320)
  const float* b_hh[2];  // (320)
};
struct LstmPacked {
  float* wihT;     // [160][640]
  float* bias;     // [640]
  float* whhT;     // [2][80][320]
  float* wih_cat;  // [640][160]
  float* whh;      // [2][320][80]
  // tcgen05 images of wih_cat for the two input projections (gemm.cuh; null = fp32 SIMT gemm_kernel):
  unsigned char* tc_fwd;  // gates[r][n] = sum_k x[r][k] wih_cat[n][k]:        N = 640, K = 160
  unsigned char* tc_bwd;  // dx[r][k]    = sum_n dgates[r][n] wih_cat[n][k]:   N = 160 (padded to 256), K = 640
};
size_t lstm_tc_fwd_bytes();
size_t lstm_tc_bwd_bytes();

int rnn_init();
int gemm(const float* A, const float* Bm, const float* bias, const float* Cadd, float* C, int M, int N, int K,
         cudaStream_t stream, const char* tag = "gemm");
// perm_wf > 0: the layer input is the NHWC block output (index w * C + c) instead of the reference's c * Wf + w order
int lstm_pack(const LstmWeights& w, const LstmPacked& p, cudaStream_t stream, int perm_wf = 0);
// x (B,L,160) -> hout (B,L,160); gates (B,L,640) and cs (B,L,2,80) are saved for backward
int blstm_forward(const LstmPacked& p, const float* x, float* gates, float* hout, float* cs, int B, int L,
                  cudaStream_t stream);
// dout (B,L,160) -> dx (B,L,160) (+ dx_add if non-null); gates is overwritten with pre-activation gradients
int blstm_backward(const LstmPacked& p, float* gates, const float* dout, const float* cs, const float* dx_add,
                   float* dx, int B, int L, cudaStream_t stream);
int feats_gather(const float* act, float* feats, int B, int L, int Wf, int C, cudaStream_t stream);
int feats_scatter(const float* gfeats, float* gact, int B, int L, int Wf, int C, cudaStream_t stream);
int head_forward(const float* l2, const float* feats, const float* w, const float* bias, float* logits, int B, int L,
                 cudaStream_t stream, int perm_wf = 0);
// mode 0: d mean-CE / d logit = 2 (sigmoid(2 o) - y) / n_global ; mode 1: 1 ; mode 2: coef[b] (caller-supplied seed)
// dfeat_add (nullable): the same gradient for the residual feature branch, written in the feature buffer's order (perm_wf)
int head_backward(const float* logits, const long long* y, const float* w, float* dl2, int B, int L, int mode,
                  int n_global, cudaStream_t stream, const float* coef = nullptr, float* dfeat_add = nullptr, int perm_wf = 0);

}  // namespace advb
