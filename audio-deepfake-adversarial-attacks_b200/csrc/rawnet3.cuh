// RawNet3 (src/models/rawnet3.py:11-291, prepare_model() configuration) forward and input-gradient backward.
// Layout: every activation is time-major [clip][time row][channel]; inside a Bottle2neck layer the rows of a clip carry
// `d` (= dilation) zero rows on each side so that the dilated k=3 convolutions are row-shifted GEMMs (gemm.cuh).
#pragma once
#include <functional>
#include <string>

#include "gemm.cuh"

namespace advb {

struct RnLayer {
  std::string name;
  int Cin = 0, d = 1, pool = 1;
  int T = 0, Tp = 0, To = 0;  // input time steps, padded rows per clip (T + 2 d), output time steps (T / pool)
  // live weights
  const float *w1 = nullptr, *b1 = nullptr, *w3 = nullptr, *b3 = nullptr, *wres = nullptr;
  const float *wc[7] = {}, *bc[7] = {};
  const float *bn1[4] = {}, *bn3[4] = {}, *bns[7][4] = {};  // weight, bias, running_mean, running_var
  const float *alpha = nullptr, *afc_w = nullptr, *afc_b = nullptr;
  // folded BatchNorm (per call)
  float *bn1_s = nullptr, *bn1_t = nullptr, *bn3_s = nullptr, *bn3_t = nullptr, *bns_s = nullptr, *bns_t = nullptr;
  // packed tcgen05 weight images: forward / backward (transposed)
  unsigned char *p_w1 = nullptr, *p_w3 = nullptr, *p_res = nullptr, *p_c[7] = {};
  unsigned char *q_w1 = nullptr, *q_w3 = nullptr, *q_res = nullptr, *q_c[7] = {};
  // forward state
  float *xin = nullptr;  // [B Tp][Cin]
  float *res = nullptr;  // [B Tp][1024]  (layer1 only)
  float *o1 = nullptr;   // [B Tp][1024]  bn1(relu(conv1))
  float *cat = nullptr;  // [B Tp][1024]
  float *y = nullptr;    // [B Tp][1024]  bn3(relu(conv3)) + residual
  float *spin[2] = {};   // [B Tp][128]   Res2 chain inputs
  float *gpre[2] = {};   // [B Tp][128]   Res2 chain backward inputs
  unsigned char *m1 = nullptr, *mc = nullptr, *m3 = nullptr;  // ReLU masks [B Tp][1024]
  float *P = nullptr;    // [B To][1024] max-pooled
  unsigned char* arg = nullptr;
  float *pm = nullptr, *yv = nullptr, *gyv = nullptr, *gz = nullptr, *gm = nullptr;  // [B][1024]
};

struct RnModel {
  int Bmax = 0, T = 0, L0 = 0, T3 = 0;
  RnLayer layer[3];
  // live weights
  const float *in_w = nullptr, *in_b = nullptr;  // preprocess.1 (InstanceNorm affine)
  const float *low_hz = nullptr, *band_hz = nullptr, *window = nullptr, *n_axis = nullptr;
  const float *w4 = nullptr, *b4 = nullptr, *wa = nullptr, *ba = nullptr, *wb = nullptr, *bb = nullptr;
  const float *abn[4] = {}, *bn5[4] = {}, *w6 = nullptr, *b6 = nullptr;
  // per call
  float *filt = nullptr;  // [256][251]
  float *abn_s = nullptr, *abn_t = nullptr, *bn5_s = nullptr, *bn5_t = nullptr;
  unsigned char *p_sinc = nullptr, *q_sinc = nullptr, *p_w4 = nullptr, *q_w4 = nullptr, *p_wa = nullptr, *q_wa = nullptr,
                *p_wb = nullptr, *q_wb = nullptr;
  // forward state
  double* pre_stats = nullptr;  // [B][2] mean, 1/sqrt(var + eps) of the pre-emphasised clip
  float *nsig = nullptr;        // [B][T] normalised signal
  float *S = nullptr;           // [B L0][256] raw sinc outputs
  float *cat4 = nullptr;        // [B T3][3072] mp3(x1) | x2 | x3
  unsigned char* arg1 = nullptr;  // [B T3][1024] arg-max of mp3(x1)
  float *H = nullptr;           // [B T3][1536]
  float *stats = nullptr;       // [B][2][1536] mean | sd
  float *var = nullptr;         // [B][1536] unbiased variance (clamp membership)
  float *cb = nullptr;          // [B][128] per-clip attention bias
  float *A1 = nullptr;          // [B T3][128]
  unsigned char* ma = nullptr;
  float *E = nullptr;           // [B T3][1536]
  float *emax = nullptr, *Z = nullptr, *vq = nullptr, *m2 = nullptr;  // [B][1536]
  float *pooled = nullptr;      // [B][3072] mu | sg
  // backward scratch
  float *gmu = nullptr, *gm2 = nullptr, *dot = nullptr;  // [B][1536]
  float *GE = nullptr, *G1 = nullptr;                    // [B T3][1536]
  float *GA = nullptr;                                   // [B T3][128]
  float *gasum = nullptr;                                // [B][128]
  float *gstat = nullptr;                                // [B][3072]
  float *gcat4 = nullptr;                                // [B T3][3072]
  float *GY = nullptr, *GC3 = nullptr, *GCAT = nullptr, *GC1 = nullptr;  // [B Tp1][1024]
  float *GX = nullptr;                                   // max(B Tp1 * 256, B Tp2 * 1024)
  float *GX1 = nullptr;                                  // [B T2][1024]
  float *GS = nullptr, *Zc = nullptr;                    // [B L0][256]
  float *gn = nullptr;                                   // [B][T]
  double* bst = nullptr;                                 // [B][2]
};

using RnAllocFn = std::function<int(void**, size_t)>;            // zero-initialised device memory, 0 on success
using RnLookupFn = std::function<const float*(const std::string&)>;

int rn_check_tensors(const std::function<int(const std::string&, long long)>& require);
int rn_build(RnModel& m, int Bmax, int T, const RnAllocFn& alloc);
void rn_bind(RnModel& m, const RnLookupFn& t);
int rn_prepare(RnModel& m, int path, cudaStream_t st);
int rn_forward(RnModel& m, const float* x, float* logits, int B, int path, int passes, cudaStream_t st);
// mode: 0 = CE (needs y, n_global), 1 = d logit, 2 = seeded (coef[b])
int rn_backward(RnModel& m, const float* x, const float* logits, const long long* y, int B, int mode, int n_global,
                const float* coef, float* gx, int path, int passes, cudaStream_t st);
// test introspection: stage name -> (pointer, rows per clip, channels); nullptr if unknown
const float* rn_stage(const RnModel& m, const std::string& name, int* rows, int* cols);

}  // namespace advb
