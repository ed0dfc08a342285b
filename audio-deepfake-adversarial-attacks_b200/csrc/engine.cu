// libadvb200: handle, workspace, LCNN execution plan and the C ABI of include/advb200.h.
//
// One handle = one (model instance, device).  The workspace is allocated once for (max_batch, n_samples); every
// API call re-reads the borrowed weight tensors (repack kernels at the head of the call), enqueues the whole
// attack on the caller's stream and returns without synchronising.
#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h>

#include <string.h>

#include <map>
#include <string>
#include <vector>

#include "../../include/advb200.h"
#include "common.cuh"
#include "conv.cuh"
#include "fabcw.cuh"
#include "frontend.cuh"
#include "rawnet3.cuh"
#include "rnn.cuh"
#include "specrnet.cuh"
#include "update.cuh"

namespace advb {

static thread_local std::string g_error;
thread_local LaunchCounter* g_counter = nullptr;
thread_local int g_conv_sched = 0;
void set_error(const std::string& msg) { g_error = msg; }

struct Profiler {
  bool on = false;
  std::vector<std::pair<const char*, cudaEvent_t>> marks;
};
static thread_local Profiler* g_prof = nullptr;
void prof_mark(const char* name, cudaStream_t stream) {
  if (g_prof == nullptr || !g_prof->on) return;
  cudaEvent_t e;
  if (cudaEventCreate(&e) != cudaSuccess) return;
  cudaEventRecord(e, stream);
  g_prof->marks.emplace_back(name, e);
}

struct TensorRef {
  const float* p;
  int64_t n;
};

struct NvtxRange {  // NVTX range around each API phase (nsys / ncu --nvtx): header-only nvtx3, a no-op without a tool attached
  explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
  ~NvtxRange() { nvtxRangePop(); }
};

struct LcnnBlock {
  int idx, Cin, Cout, KS, bn_idx;
  bool pool;
  int H, W, Ho, Wo;    // conv (pre-pool) size, block output size
  float* wf = nullptr;  // packed forward weights
  float* wd = nullptr;  // packed backward weights
  unsigned char* tcf = nullptr;  // tensor-core weight slices (forward / backward)
  unsigned char* tcd = nullptr;
  unsigned char* tcf_mix = nullptr;  // forward slices of the mixed mode (tf32_passes = 2): tf32 hi + bf16 cross image
  bool tc = false;
  float* invstd = nullptr;
  Act out{};             // block output (zero-bordered for the next conv)
  unsigned char* codes = nullptr;
  float* gout = nullptr;  // compact gradient of the block output
  std::string tag_f, tag_b;
};

}  // namespace advb

using namespace advb;

struct advb_handle {
  int device = 0, model_kind = 0, frontend_kind = 0, Bmax = 0, T = 0, F = 0;
  std::map<std::string, TensorRef> tensors;
  LaunchCounter counter;
  Profiler prof;
  std::vector<void*> allocs;
  size_t ws_bytes = 0;
  int conv_path = 0;    // 0 = tcgen05 tensor cores (default), 1 = fp32 SIMT cross-check path
  int tf32_passes = 3;  // 3 = 3xTF32 (fp32-class accuracy), 1 = single-pass tf32
  int conv_sched = 0;   // 0 = persistent warp-specialised conv kernels, 1 = one-tile-per-CTA kernels only
  int conv0_bwd = 0;    // first block backward: 0 = fp32 cell kernel (conv0_bwd.cu), 1 = tcgen05 GEMM + col2im
  int conv0_fwd = 0;    // first block forward: 0 = Toeplitz GEMM without im2col (conv0_toeplitz.cu), 1 = im2col GEMM (conv_light.cu)
  int use_graph = 1;    // 1 = the PGD / PGDL2 iteration is captured once into a CUDA graph and replayed
  int fuse_update = 1;  // 1 = FGSM / PGD update rule applied in the frontend backward's epilogue (no gradient in HBM)
  int weight_cache = 0; // 1 = the caller promises advb_invalidate_weights() after every weight change: skip the repack otherwise
  bool packed_valid = false;
  int64_t bind_epoch = 0;  // bumped whenever the borrowed tensor table changes (graph cache key)

  // cached CUDA graph of one attack iteration (all kernel arguments are engine-owned buffers, see run_attack)
  struct GraphCache {
    cudaGraphExec_t exec = nullptr;
    cudaGraph_t graph = nullptr;
    cudaEvent_t done = nullptr;  // recorded after the last replay: waited on before the exec is destroyed
    int64_t launches = 0;        // kernel launches per replay
    std::string key;
  } gc;
  cudaStream_t cap_stream = nullptr;
  cudaEvent_t last_done = nullptr;     // recorded at the end of every call: a call on another stream waits for it (CallScope)
  cudaStream_t last_stream = nullptr;
  float *x_in = nullptr, *adv2 = nullptr;  // engine-owned copies: clean clips, second ping-pong iterate
  int64_t* y_in = nullptr;

  // frontend
  float2* tw = nullptr;
  float *dB = nullptr, *g_dB = nullptr, *mass_partial = nullptr, *g_coef = nullptr;
  int *klo = nullptr, *kcnt = nullptr, *mlo = nullptr, *mcnt = nullptr;
  float* dctT = nullptr;
  float2* spec = nullptr;  // packed spectra of the last forward (option "fe_spec"): the backward reads them instead of recomputing
  int fe_spec = 1;
  FrontendState fst{};
  FrontendHost fe_host{};
  // strict multi-GPU mode (advb_xrank_*): own mailbox + peers' mailboxes mapped with CUDA IPC
  unsigned long long* xr_mailbox = nullptr;
  std::vector<void*> xr_ipc_opened;
  int xr_gen = 0;  // graph cache key: the captured exchange nodes hold the peers' pointers
  FrontendTables ftb{};

  // LCNN
  Act act0{};
  LcnnBlock blk[9];
  int L = 0, Wf = 0;
  float *feats = nullptr, *l1 = nullptr, *l2 = nullptr, *gates1 = nullptr, *gates2 = nullptr, *cs1 = nullptr,
        *cs2 = nullptr, *dl2 = nullptr, *dl1 = nullptr, *dfeats = nullptr, *logits = nullptr;
  LstmPacked lp[2]{};
  unsigned char* lp_tc[2][2] = {{nullptr, nullptr}, {nullptr, nullptr}};  // tcgen05 images of the BLSTM input projections
  int sr_tc = 1;    // SpecRNet: 1 = the 64 -> 64 convolutions (conv2 of blocks 2 and 4) on the tcgen05 kernel (round 2), 0 = fp32 SIMT
  int lstm_tc = 0;  // 1 = BLSTM input projections on the tcgen05 GEMM (3xTF32; measured 21 / 40 us vs 30 / 37 us: the persistent
                    // GEMM's fixed costs eat the gain at 0.66 GFLOP), 0 = fp32 SIMT GEMM (default)

  // SpecRNet
  SrBlock sr[3]{};
  SrGru gru{};
  float *sr_img = nullptr, *sr_bn4 = nullptr;
  int sr_L = 0;

  // RawNet3
  RnModel rn{};

  // attack scratch
  unsigned char* c0t_w = nullptr;  // Toeplitz weight image of the first block (conv0_toeplitz.cu)
  float* conv0_T = nullptr;  // (B,F,80,5) horizontal col2im partial sums of the first block's backward
  float *grad = nullptr, *partial_g = nullptr, *partial_d = nullptr, *coef_tmp = nullptr;
  FabScratch fab{};  // allocated on the first FAB / CW call
  CwScratch cw{};
  float *mm_mn = nullptr, *mm_mx = nullptr;  // advb_attack_minmax: per-clip min / max (first use)
  float* host_cost = nullptr;  // pinned: CW's batch-wide early-stop scalar (cw.py:107-110)

  template <typename Tp>
  int alloc(Tp** out, size_t count) {
    void* p = nullptr;
    size_t bytes = count * sizeof(Tp);
    if (bytes == 0) bytes = 16;
    ADVB_CUDA_OK(cudaMalloc(&p, bytes));
    ADVB_CUDA_OK(cudaMemset(p, 0, bytes));
    allocs.push_back(p);
    ws_bytes += bytes;
    *out = reinterpret_cast<Tp*>(p);
    return 0;
  }
  const float* t(const std::string& name) const {
    auto it = tensors.find(name);
    return it == tensors.end() ? nullptr : it->second.p;
  }
};

namespace {

struct DeviceGuard {
  int prev = -1;
  explicit DeviceGuard(int dev) {
    cudaGetDevice(&prev);
    if (prev != dev) cudaSetDevice(dev);
    else prev = -1;
  }
  ~DeviceGuard() {
    if (prev >= 0) cudaSetDevice(prev);
  }
};

const int kLcnnSpec[9][6] = {
    // idx, Cin, Cout, KS, pool, bn_idx   (src/models/lcnn.py:121-153)
    {0, 1, 64, 5, 1, -1},  {3, 32, 64, 1, 0, 5},    {6, 32, 96, 3, 1, 9},   {10, 48, 96, 1, 0, 12}, {13, 48, 128, 3, 1, -1},
    {16, 64, 128, 1, 0, 18}, {19, 64, 64, 3, 0, 21}, {22, 32, 64, 1, 0, 24}, {25, 32, 64, 3, 1, -1},
};

int bind_tensors(advb_handle* h, int n, const advb_tensor_ref* refs) {
  h->bind_epoch++;
  h->packed_valid = false;
  for (int i = 0; i < n; ++i) {
    ADVB_CHECK(refs[i].name != nullptr && refs[i].ptr != nullptr, "null tensor reference");
    h->tensors[refs[i].name] = TensorRef{refs[i].ptr, refs[i].numel};
  }
  return 0;
}

int require(const advb_handle* h, const std::string& name, int64_t numel) {
  auto it = h->tensors.find(name);
  if (it == h->tensors.end()) {
    set_error("missing tensor '" + name + "'");
    return 1;
  }
  if (it->second.n != numel) {
    set_error("tensor '" + name + "' has " + std::to_string(it->second.n) + " elements, expected " +
              std::to_string(numel));
    return 1;
  }
  return 0;
}

int check_frontend_tensors(advb_handle* h) {
  if (h->frontend_kind == ADVB_FRONTEND_LFCC) {
    ADVB_TRY(require(h, "frontend.filter_mat", 257 * 128));
    ADVB_TRY(require(h, "frontend.dct_mat", 128 * 80));
    ADVB_TRY(require(h, "frontend.Spectrogram.window", 400));
  } else if (h->frontend_kind == ADVB_FRONTEND_MFCC) {
    ADVB_TRY(require(h, "frontend.MelSpectrogram.mel_scale.fb", 257 * 128));
    ADVB_TRY(require(h, "frontend.dct_mat", 128 * 80));
    ADVB_TRY(require(h, "frontend.MelSpectrogram.spectrogram.window", 400));
  } else {
    set_error("model needs an LFCC or MFCC frontend");
    return 1;
  }
  return 0;
}

void refresh_frontend_tables(advb_handle* h) {
  FrontendTables& tb = h->ftb;
  if (h->frontend_kind == ADVB_FRONTEND_LFCC) {
    tb.fb = h->t("frontend.filter_mat");
    tb.window = h->t("frontend.Spectrogram.window");
  } else {
    tb.fb = h->t("frontend.MelSpectrogram.mel_scale.fb");
    tb.window = h->t("frontend.MelSpectrogram.spectrogram.window");
  }
  tb.dct = h->t("frontend.dct_mat");
  tb.tw = h->tw;
  tb.dctT = h->dctT;
  tb.klo = h->klo;
  tb.kcnt = h->kcnt;
  tb.mlo = h->mlo;
  tb.mcnt = h->mcnt;
}

int check_lcnn_tensors(advb_handle* h) {
  for (int i = 0; i < 9; ++i) {
    const int* s = kLcnnSpec[i];
    const std::string p = "m_transform." + std::to_string(s[0]);
    ADVB_TRY(require(h, p + ".weight", (int64_t)s[2] * s[1] * s[3] * s[3]));
    ADVB_TRY(require(h, p + ".bias", s[2]));
    if (s[5] >= 0) {
      const std::string q = "m_transform." + std::to_string(s[5]);
      ADVB_TRY(require(h, q + ".running_mean", s[2] / 2));
      ADVB_TRY(require(h, q + ".running_var", s[2] / 2));
    }
  }
  for (int l = 0; l < 2; ++l)
    for (int d = 0; d < 2; ++d) {
      const std::string p = "m_before_pooling." + std::to_string(l) + ".l_blstm.";
      const std::string sfx = d == 0 ? "_l0" : "_l0_reverse";
      ADVB_TRY(require(h, p + "weight_ih" + sfx, 320 * 160));
      ADVB_TRY(require(h, p + "weight_hh" + sfx, 320 * 80));
      ADVB_TRY(require(h, p + "bias_ih" + sfx, 320));
      ADVB_TRY(require(h, p + "bias_hh" + sfx, 320));
    }
  ADVB_TRY(require(h, "m_output_act.weight", 160));
  ADVB_TRY(require(h, "m_output_act.bias", 1));
  return 0;
}

int build_lcnn(advb_handle* h) {
  const int B = h->Bmax, F = h->F;
  ADVB_TRY(check_lcnn_tensors(h));
  // cepstral image (B, F, 80, 1) with the 2-pixel border of the 5x5 conv
  h->act0 = Act{nullptr, F, 80, 1, 2};
  ADVB_TRY(h->alloc(&h->act0.p, (size_t)B * h->act0.per_clip()));
  int H = F, W = 80;
  for (int i = 0; i < 9; ++i) {
    LcnnBlock& k = h->blk[i];
    const int* s = kLcnnSpec[i];
    k.idx = s[0];
    k.Cin = s[1];
    k.Cout = s[2];
    k.KS = s[3];
    k.pool = s[4] != 0;
    k.bn_idx = s[5];
    k.tag_f = "conv_fwd_b" + std::to_string(i);
    k.tag_b = "conv_bwd_b" + std::to_string(i);
    k.H = H;
    k.W = W;
    k.Ho = k.pool ? H / 2 : H;
    k.Wo = k.pool ? W / 2 : W;
    ADVB_CHECK(k.Ho > 0 && k.Wo > 0, "clip too short for the LCNN pooling stack");
    const int next_pad = i < 8 ? kLcnnSpec[i + 1][3] / 2 : 0;
    k.out = Act{nullptr, k.Ho, k.Wo, k.Cout / 2, next_pad};
    ADVB_TRY(h->alloc(&k.out.p, (size_t)B * k.out.per_clip()));
    ADVB_TRY(h->alloc(&k.codes, (size_t)B * k.Ho * k.Wo * (k.Cout / 2)));
    ADVB_TRY(h->alloc(&k.gout, (size_t)B * k.Ho * k.Wo * (k.Cout / 2)));
    const size_t wn = (size_t)k.Cout * k.Cin * k.KS * k.KS;
    ADVB_TRY(h->alloc(&k.wf, wn));
    if (i > 0) ADVB_TRY(h->alloc(&k.wd, wn));
    if (k.bn_idx >= 0) ADVB_TRY(h->alloc(&k.invstd, k.Cout / 2));
    k.tc = conv_tc_supported(k.Cin, k.Cout, k.KS, k.pool) &&
           (128 * 2) / (k.W + 2 * (k.KS / 2 == 2 ? 0 : k.KS / 2)) >= 2;  // the smallest CTA tile (2 M-tiles) must hold >= 2 image rows
    if (k.tc) {
      ADVB_TRY(h->alloc(&k.tcf, conv_tc_pack_bytes(k.Cout, k.Cin, k.KS, false)));
      ADVB_TRY(h->alloc(&k.tcd, conv_tc_pack_bytes(k.Cout, k.Cin, k.KS, true)));
      if (k.KS == 3) ADVB_TRY(h->alloc(&k.tcf_mix, conv_tc_pack_bytes(k.Cout, k.Cin, k.KS, false)));
    }
    H = k.Ho;
    W = k.Wo;
  }
  ADVB_TRY(h->alloc(&h->conv0_T, (size_t)B * F * 80 * 5));
  ADVB_TRY(h->alloc(&h->c0t_w, conv0t_pack_bytes()));
  h->L = H;
  h->Wf = W;
  ADVB_CHECK(h->Wf * 32 == 160, "LCNN feature width must be 160");
  ADVB_CHECK(h->blk[8].out.pad == 0, "the last block's output feeds the BLSTM without a border");
  const size_t bl = (size_t)B * h->L;
  ADVB_TRY(h->alloc(&h->feats, bl * 160));
  ADVB_TRY(h->alloc(&h->l1, bl * 160));
  ADVB_TRY(h->alloc(&h->l2, bl * 160));
  ADVB_TRY(h->alloc(&h->gates1, bl * 640));
  ADVB_TRY(h->alloc(&h->gates2, bl * 640));
  ADVB_TRY(h->alloc(&h->cs1, bl * 160));
  ADVB_TRY(h->alloc(&h->cs2, bl * 160));
  ADVB_TRY(h->alloc(&h->dl2, bl * 160));
  ADVB_TRY(h->alloc(&h->dl1, bl * 160));
  ADVB_TRY(h->alloc(&h->dfeats, bl * 160));
  for (int l = 0; l < 2; ++l) {
    ADVB_TRY(h->alloc(&h->lp[l].wihT, 160 * 640));
    ADVB_TRY(h->alloc(&h->lp[l].bias, 640));
    ADVB_TRY(h->alloc(&h->lp[l].whhT, 2 * 80 * 320));
    ADVB_TRY(h->alloc(&h->lp[l].wih_cat, 640 * 160));
    ADVB_TRY(h->alloc(&h->lp[l].whh, 2 * 320 * 80));
    ADVB_TRY(h->alloc(&h->lp_tc[l][0], lstm_tc_fwd_bytes()));
    ADVB_TRY(h->alloc(&h->lp_tc[l][1], lstm_tc_bwd_bytes()));
  }
  ADVB_TRY(rnn_init());
  return 0;
}

// Re-read the live weights: repack convolution / LSTM weights, BatchNorm inverse std, filterbank extents.
int prepare_lcnn(advb_handle* h, cudaStream_t st) {
  refresh_frontend_tables(h);
  ADVB_TRY(frontend_prepare(h->ftb, st));
  for (int i = 0; i < 9; ++i) {
    LcnnBlock& k = h->blk[i];
    const std::string p = "m_transform." + std::to_string(k.idx);
    if (k.tc && h->conv_path == 0) {
      ADVB_TRY(conv_tc_pack(h->t(p + ".weight"), k.tcf, k.tcd, k.Cout, k.Cin, k.KS, st));
      if (conv_p3_bwd_hs(k.Cin, k.Cout, k.KS, k.pool, k.W))  // same bytes, horizontal-scatter layout
        ADVB_TRY(conv_p3_pack_hs(h->t(p + ".weight"), k.tcd, k.Cout, k.Cin, st));
      if (h->tf32_passes == 2 && k.tcf_mix != nullptr) ADVB_TRY(conv_p3_pack_mix(h->t(p + ".weight"), k.tcf_mix, k.Cout, k.Cin, st));
    }
    if (!(k.tc && h->conv_path == 0))
      ADVB_TRY(conv_pack_weights(h->t(p + ".weight"), k.wf, k.wd, k.Cout, k.Cin, k.KS, st));
    if (k.bn_idx >= 0)
      ADVB_TRY(bn_prepare(h->t("m_transform." + std::to_string(k.bn_idx) + ".running_var"), k.invstd, k.Cout / 2, st));
  }
  ADVB_TRY(conv0t_pack(h->t("m_transform.0.weight"), h->t("m_transform.0.bias"), h->c0t_w, st));
  for (int l = 0; l < 2; ++l) {
    const std::string p = "m_before_pooling." + std::to_string(l) + ".l_blstm.";
    LstmWeights w;
    for (int d = 0; d < 2; ++d) {
      const std::string sfx = d == 0 ? "_l0" : "_l0_reverse";
      w.w_ih[d] = h->t(p + "weight_ih" + sfx);
      w.w_hh[d] = h->t(p + "weight_hh" + sfx);
      w.b_ih[d] = h->t(p + "bias_ih" + sfx);
      w.b_hh[d] = h->t(p + "bias_hh" + sfx);
    }
    h->lp[l].tc_fwd = (h->lstm_tc && h->conv_path == 0) ? h->lp_tc[l][0] : nullptr;
    h->lp[l].tc_bwd = (h->lstm_tc && h->conv_path == 0) ? h->lp_tc[l][1] : nullptr;
    ADVB_TRY(lstm_pack(w, h->lp[l], st, l == 0 ? h->Wf : 0));  // layer 0 reads the last block's NHWC output as it is
  }
  return 0;
}

int lcnn_forward(advb_handle* h, const float* x, int B, cudaStream_t st) {
  const Act& a0 = h->act0;
  ADVB_TRY(frontend_forward(h->ftb, h->fst, x, B, h->T, h->dB, a0.p, (long long)a0.per_clip(), (long long)a0.Wp(), 1,
                            (long long)a0.pad * a0.Wp() + a0.pad, st, h->fe_spec ? h->spec : nullptr, &h->fe_host));
  const float* in = a0.p;
  int in_pad = a0.pad;
  for (int i = 0; i < 9; ++i) {
    LcnnBlock& k = h->blk[i];
    ConvFwdArgs a{};
    a.in = in;
    a.in_pad = in_pad;
    a.wf = k.wf;
    a.bias = h->t("m_transform." + std::to_string(k.idx) + ".bias");
    a.bn_mean = k.bn_idx >= 0 ? h->t("m_transform." + std::to_string(k.bn_idx) + ".running_mean") : nullptr;
    a.bn_invstd = k.bn_idx >= 0 ? k.invstd : nullptr;
    a.out = k.out.p;
    a.out_pad = k.out.pad;
    a.codes = k.codes;
    a.B = B;
    a.H = k.H;
    a.W = k.W;
    a.Cin = k.Cin;
    a.Cout = k.Cout;
    a.KS = k.KS;
    a.Ho = k.Ho;
    a.Wo = k.Wo;
    a.pool = k.pool;
    a.tag = k.tag_f.c_str();
    // tf32_passes = 2 (mixed mode) exists for the forward 3x3 kernels only; everything else keeps 3xTF32
    const int passes = h->tf32_passes == 2 ? 3 : h->tf32_passes;
    if (i == 0 && k.tc && h->conv_path == 0 && h->conv0_fwd == 0 && conv0t_supported(k.H, k.W, k.Ho, k.Wo))
      ADVB_TRY(conv0t_forward(a, h->c0t_w, passes, st));
    else if (k.tc && h->conv_path == 0 && h->tf32_passes == 2 && k.tcf_mix != nullptr &&
             conv_p3_supported(k.Cin, k.Cout, k.KS, k.pool, k.W))
      ADVB_TRY(conv_p3_forward(a, k.tcf_mix, 2, st));
    else if (k.tc && h->conv_path == 0) ADVB_TRY(conv_tc_forward(a, k.tcf, passes, st));
    else ADVB_TRY(conv_mfm_forward(a, st));
    in = k.out.p;
    in_pad = k.out.pad;
  }
  // The (B, L, Wf, 32) output of the last block IS the BLSTM's input: the permutation of lcnn.py:196-199 lives in the packed
  // input-projection weights of layer 0 and in the head's index map (no gather kernel)
  const float* feats = h->blk[8].out.p;
  ADVB_TRY(blstm_forward(h->lp[0], feats, h->gates1, h->l1, h->cs1, B, h->L, st));
  ADVB_TRY(blstm_forward(h->lp[1], h->l1, h->gates2, h->l2, h->cs2, B, h->L, st));
  ADVB_TRY(head_forward(h->l2, feats, h->t("m_output_act.weight"), h->t("m_output_act.bias"), h->logits, B, h->L, st,
                        h->Wf));
  return 0;
}

// Gradient of (mode 0) the mean 2-class CE or (mode 1) the logit, w.r.t. the waveform of the last forward.
int lcnn_backward(advb_handle* h, const float* x, const int64_t* y, int B, int mode, int n_global, float* gx,
                  cudaStream_t st, const float* coef = nullptr, const FusedUpdate* upd = nullptr) {
  // dfeats = the head's residual contribution, already in the block's NHWC order; the input gradient of layer 0 then lands
  // in the last block's gradient buffer directly (no scatter kernel)
  ADVB_TRY(head_backward(h->logits, reinterpret_cast<const long long*>(y), h->t("m_output_act.weight"), h->dl2, B,
                         h->L, mode, n_global, st, coef, h->dfeats, h->Wf));
  ADVB_TRY(blstm_backward(h->lp[1], h->gates2, h->dl2, h->cs2, nullptr, h->dl1, B, h->L, st));
  ADVB_TRY(blstm_backward(h->lp[0], h->gates1, h->dl1, h->cs1, h->dfeats, h->blk[8].gout, B, h->L, st));
  for (int i = 8; i >= 1; --i) {
    LcnnBlock& k = h->blk[i];
    ConvBwdArgs a{};
    a.gout = k.gout;
    a.codes = k.codes;
    a.bn_invstd = k.bn_idx >= 0 ? k.invstd : nullptr;
    a.wd = k.wd;
    a.gin = h->blk[i - 1].gout;
    a.B = B;
    a.H = k.H;
    a.W = k.W;
    a.Cin = k.Cin;
    a.Cout = k.Cout;
    a.KS = k.KS;
    a.Ho = k.Ho;
    a.Wo = k.Wo;
    a.pool = k.pool;
    a.tag = k.tag_b.c_str();
    if (k.tc && h->conv_path == 0) ADVB_TRY(conv_tc_backward(a, k.tcd, h->tf32_passes == 2 ? 3 : h->tf32_passes, st));
    else ADVB_TRY(conv_mfm_backward(a, st));
  }
  const LcnnBlock& k0 = h->blk[0];
  if (k0.tc && h->conv_path == 0 && h->conv0_bwd == 0 && conv0_cells_supported(k0.H, k0.W, k0.Ho, k0.Wo))
    ADVB_TRY(conv0_cells_backward(k0.gout, k0.codes, h->t("m_transform.0.weight"), h->g_coef, B, k0.H, k0.W, k0.Ho, k0.Wo, st));
  else if (k0.tc && h->conv_path == 0)
    ADVB_TRY(conv0_tc_backward(k0.gout, k0.codes, k0.tcd, h->conv0_T, h->g_coef, B, k0.H, k0.W, k0.Ho, k0.Wo,
                               h->tf32_passes == 2 ? 3 : h->tf32_passes, st));
  else
    ADVB_TRY(conv0_backward(k0.gout, k0.codes, h->t("m_transform.0.weight"), h->g_coef, B, k0.H, k0.W, k0.Ho, k0.Wo, st));
  ADVB_TRY(frontend_backward(h->ftb, h->fst, x, B, h->T, h->dB, h->g_coef, (long long)h->F * 80, 80, 1,
                             h->mass_partial, h->g_dB, gx, st, upd, h->fe_spec ? h->spec : nullptr, &h->fe_host));
  return 0;
}

// ---- SpecRNet (src/models/specrnet.py) ----
const char* kSrNames[3] = {"0", "2", "4"};

int check_specrnet_tensors(advb_handle* h) {
  const int cin[3] = {1, 20, 64}, cout[3] = {20, 64, 64};
  ADVB_TRY(require(h, "first_bn.weight", 1));
  ADVB_TRY(require(h, "first_bn.bias", 1));
  ADVB_TRY(require(h, "first_bn.running_mean", 1));
  ADVB_TRY(require(h, "first_bn.running_var", 1));
  for (int i = 0; i < 3; ++i) {
    const std::string p = std::string("block") + kSrNames[i] + ".0.";
    ADVB_TRY(require(h, p + "conv1.weight", (int64_t)cout[i] * cin[i] * 9));
    ADVB_TRY(require(h, p + "conv1.bias", cout[i]));
    ADVB_TRY(require(h, p + "conv2.weight", (int64_t)cout[i] * cout[i] * 9));
    ADVB_TRY(require(h, p + "conv2.bias", cout[i]));
    for (const char* q : {"bn2.weight", "bn2.bias", "bn2.running_mean", "bn2.running_var"}) ADVB_TRY(require(h, p + q, cout[i]));
    if (cin[i] != cout[i]) {
      ADVB_TRY(require(h, p + "conv_downsample.weight", (int64_t)cout[i] * cin[i]));
      ADVB_TRY(require(h, p + "conv_downsample.bias", cout[i]));
    }
    const std::string a = std::string("fc_attention") + kSrNames[i] + ".0.";
    ADVB_TRY(require(h, a + "weight", (int64_t)cout[i] * cout[i]));
    ADVB_TRY(require(h, a + "bias", cout[i]));
  }
  for (const char* q : {"weight", "bias", "running_mean", "running_var"}) ADVB_TRY(require(h, std::string("bn_before_gru.") + q, 64));
  for (int l = 0; l < 2; ++l)
    for (int d = 0; d < 2; ++d) {
      const std::string sfx = "_l" + std::to_string(l) + (d ? "_reverse" : "");
      ADVB_TRY(require(h, "gru.weight_ih" + sfx, 192 * (l == 0 ? 64 : 128)));
      ADVB_TRY(require(h, "gru.weight_hh" + sfx, 192 * 64));
      ADVB_TRY(require(h, "gru.bias_ih" + sfx, 192));
      ADVB_TRY(require(h, "gru.bias_hh" + sfx, 192));
    }
  ADVB_TRY(require(h, "fc1_gru.weight", 128 * 128));
  ADVB_TRY(require(h, "fc1_gru.bias", 128));
  ADVB_TRY(require(h, "fc2_gru.weight", 128));
  ADVB_TRY(require(h, "fc2_gru.bias", 1));
  return 0;
}

void bind_specrnet(advb_handle* h) {
  for (int i = 0; i < 3; ++i) {
    SrBlock& k = h->sr[i];
    const std::string p = std::string("block") + kSrNames[i] + ".0.";
    k.w1 = h->t(p + "conv1.weight"), k.b1 = h->t(p + "conv1.bias");
    k.w2 = h->t(p + "conv2.weight"), k.b2 = h->t(p + "conv2.bias");
    k.wds = h->t(p + "conv_downsample.weight"), k.bds = h->t(p + "conv_downsample.bias");
    k.bn_w = h->t(p + "bn2.weight"), k.bn_b = h->t(p + "bn2.bias");
    k.bn_rm = h->t(p + "bn2.running_mean"), k.bn_rv = h->t(p + "bn2.running_var");
    const std::string a = std::string("fc_attention") + kSrNames[i] + ".0.";
    k.att_w = h->t(a + "weight"), k.att_b = h->t(a + "bias");
  }
  SrGru& g = h->gru;
  for (int l = 0; l < 2; ++l)
    for (int d = 0; d < 2; ++d) {
      const std::string sfx = "_l" + std::to_string(l) + (d ? "_reverse" : "");
      g.w_ih[l][d] = h->t("gru.weight_ih" + sfx), g.w_hh[l][d] = h->t("gru.weight_hh" + sfx);
      g.b_ih[l][d] = h->t("gru.bias_ih" + sfx), g.b_hh[l][d] = h->t("gru.bias_hh" + sfx);
    }
  g.fc1_w = h->t("fc1_gru.weight"), g.fc1_b = h->t("fc1_gru.bias");
  g.fc2_w = h->t("fc2_gru.weight"), g.fc2_b = h->t("fc2_gru.bias");
  g.bn_w = h->t("bn_before_gru.weight"), g.bn_b = h->t("bn_before_gru.bias");
  g.bn_rm = h->t("bn_before_gru.running_mean"), g.bn_rv = h->t("bn_before_gru.running_var");
}

int build_specrnet(advb_handle* h) {
  const size_t B = h->Bmax;
  ADVB_TRY(check_specrnet_tensors(h));
  const int cin[3] = {1, 20, 64}, cout[3] = {20, 64, 64};
  int H = h->F, W = 80;
  ADVB_TRY(h->alloc(&h->sr_img, B * (H + 2) * (W + 2)));
  ADVB_TRY(h->alloc(&h->sr_bn4, 4));
  for (int i = 0; i < 3; ++i) {
    SrBlock& k = h->sr[i];
    k.Cin = cin[i], k.Cout = cout[i];
    k.Ci = cin[i] == 1 ? 1 : (cin[i] + 7) / 8 * 8;
    k.C = (cout[i] + 7) / 8 * 8;
    k.H = H, k.W = W, k.Hb = H / 2, k.Wb = W / 2, k.Hn = k.Hb / 2, k.Wn = k.Wb / 2;
    ADVB_CHECK(k.Hn > 0 && k.Wn > 0, "clip too short for the SpecRNet pooling stack");
    k.downsample = cin[i] != cout[i];
    k.n_tiles = sr_conv2_tiles(H, W, k.C);
    k.xn_pad = i < 2 ? 1 : 0;
    const size_t C = k.C;
    ADVB_TRY(h->alloc(&k.w1f, 9 * (size_t)k.Ci * C));
    ADVB_TRY(h->alloc(&k.w1d, 9 * C * (size_t)(k.Ci < 8 ? 8 : k.Ci)));
    ADVB_TRY(h->alloc(&k.w2f, 9 * C * C));
    ADVB_TRY(h->alloc(&k.w2d, 9 * C * C));
    ADVB_TRY(h->alloc(&k.wdf, (size_t)k.Ci * C));
    ADVB_TRY(h->alloc(&k.wdd, C * (size_t)(k.Ci < 8 ? 8 : k.Ci)));
    ADVB_TRY(h->alloc(&k.b1p, C));
    ADVB_TRY(h->alloc(&k.b2p, C));
    ADVB_TRY(h->alloc(&k.bdp, C));
    ADVB_TRY(h->alloc(&k.bn_scale, C));
    ADVB_TRY(h->alloc(&k.bn_shift, C));
    ADVB_TRY(h->alloc(&k.h, B * (H + 2) * (W + 2) * C));
    ADVB_TRY(h->alloc(&k.xb, B * k.Hb * k.Wb * C));
    ADVB_TRY(h->alloc(&k.code1, B * k.Hb * k.Wb * C));
    ADVB_TRY(h->alloc(&k.psum, B * k.n_tiles * C));
    ADVB_TRY(h->alloc(&k.y, B * C));
    ADVB_TRY(h->alloc(&k.xn, B * (k.Hn + 2 * k.xn_pad) * (k.Wn + 2 * k.xn_pad) * C));
    ADVB_TRY(h->alloc(&k.code2, B * k.Hn * k.Wn * C));
    ADVB_TRY(h->alloc(&k.g_xn, B * k.Hn * k.Wn * C));
    ADVB_TRY(h->alloc(&k.gadd, B * C));
    ADVB_TRY(h->alloc(&k.g_c1, B * H * W * C));
    if ((k.C == 64 && W <= 40) || (k.C == 24 && W <= 80)) {  // conv2 on the persistent tcgen05 kernel (conv_path = 0)
      ADVB_TRY(h->alloc(&k.w2t, 9 * C * C));
      const int npad = k.C == 64 ? 64 : 32;
      ADVB_TRY(h->alloc(&k.tcf2, conv_tc_pack_bytes_padded(k.Cout, k.Cout, 3, false, npad)));
      ADVB_TRY(h->alloc(&k.tcd2, conv_tc_pack_bytes_padded(k.Cout, k.Cout, 3, true, npad)));
      ADVB_TRY(h->alloc(&k.c2, B * H * W * C));
      ADVB_TRY(h->alloc(&k.go, B * (H + 2) * (W + 2) * C));
      if (i > 0 && conv_p3_plain_supported(k.Ci, k.C, W) && conv_p3_plain_supported(k.C, k.Ci, W)) {  // conv1 and its transpose too
        ADVB_TRY(h->alloc(&k.w1t, 9 * (size_t)k.Cout * k.Cin));
        ADVB_TRY(h->alloc(&k.tcf1, conv_tc_pack_bytes_padded(k.Cout, k.Cin, 3, false, k.C == 64 ? 64 : 32)));
        ADVB_TRY(h->alloc(&k.tcd1, conv_tc_pack_bytes_padded(k.Cout, k.Cin, 3, true, k.Ci == 64 ? 64 : 32)));
        ADVB_TRY(h->alloc(&k.g_c1b, B * (H + 2) * (W + 2) * C));
      }
    }
    H = k.Hn, W = k.Wn;
  }
  ADVB_CHECK(W == 1, "SpecRNet expects the coefficient axis to pool down to 1");
  h->sr_L = H;
  SrGru& g = h->gru;
  for (int l = 0; l < 2; ++l)
    for (int d = 0; d < 2; ++d) {
      ADVB_TRY(h->alloc(&g.wihT[l][d], 192 * (size_t)(l == 0 ? 64 : 128)));
      ADVB_TRY(h->alloc(&g.whhT[l][d], 192 * 64));
    }
  ADVB_TRY(h->alloc(&g.v, 128));
  ADVB_TRY(h->alloc(&g.bn_scale, 64));
  ADVB_TRY(h->alloc(&g.bn_shift, 64));
  ADVB_TRY(h->alloc(&g.gates, B * 2 * 2 * h->sr_L * 256));
  ADVB_TRY(h->alloc(&g.outs, B * 2 * h->sr_L * 128));
  ADVB_TRY(h->alloc(&g.xin, B * h->sr_L * 64));
  return 0;
}

int prepare_specrnet(advb_handle* h, cudaStream_t st) {
  refresh_frontend_tables(h);
  ADVB_TRY(frontend_prepare(h->ftb, st));
  bind_specrnet(h);
  ADVB_TRY(sr_pack_first_bn(h->t("first_bn.weight"), h->t("first_bn.bias"), h->t("first_bn.running_mean"),
                            h->t("first_bn.running_var"), h->sr_bn4, st));
  for (int i = 0; i < 3; ++i) {
    h->sr[i].tc2 = h->sr[i].tcf2 != nullptr && h->conv_path == 0 && h->sr_tc != 0;
    h->sr[i].tc1 = h->sr[i].tc2 && h->sr[i].tcf1 != nullptr;
    ADVB_TRY(sr_pack_block(h->sr[i], st));
  }
  ADVB_TRY(sr_pack_gru(h->gru, st));
  return 0;
}

int specrnet_forward(advb_handle* h, const float* x, int B, cudaStream_t st) {
  const int F = h->F;
  ADVB_TRY(frontend_forward(h->ftb, h->fst, x, B, h->T, h->dB, h->sr_img, (long long)(F + 2) * 82, 82, 1, 82 + 1, st,
                            h->fe_spec ? h->spec : nullptr, &h->fe_host));
  ADVB_TRY(sr_input_forward(h->sr_img, h->sr_bn4, B, F, 80, st));
  const float* in = h->sr_img;
  const char* tags[3] = {"sr_b0", "sr_b2", "sr_b4"};
  for (int i = 0; i < 3; ++i) {
    ADVB_TRY(sr_block_forward(h->sr[i], in, B, tags[i], st));
    in = h->sr[i].xn;
  }
  ADVB_TRY(sr_gru_forward(h->gru, h->sr[2].xn, h->logits, B, h->sr_L, st));
  return 0;
}

int specrnet_backward(advb_handle* h, const float* x, const int64_t* y, int B, int mode, int n_global, float* gx,
                      cudaStream_t st, const float* coef, const FusedUpdate* upd = nullptr) {
  ADVB_TRY(sr_gru_backward(h->gru, h->sr[2].xn, h->logits, reinterpret_cast<const long long*>(y), mode, n_global, coef,
                           h->sr[2].g_xn, B, h->sr_L, st));
  const char* tags[3] = {"sr_b0", "sr_b2", "sr_b4"};
  for (int i = 2; i >= 0; --i) {
    const float* in = i == 0 ? h->sr_img : h->sr[i - 1].xn;
    float* gin = i == 0 ? h->g_coef : h->sr[i - 1].g_xn;
    ADVB_TRY(sr_block_backward(h->sr[i], in, gin, B, i == 0, h->sr_bn4, tags[i], st));
  }
  ADVB_TRY(frontend_backward(h->ftb, h->fst, x, B, h->T, h->dB, h->g_coef, (long long)h->F * 80, 80, 1, h->mass_partial,
                             h->g_dB, gx, st, upd, h->fe_spec ? h->spec : nullptr, &h->fe_host));
  return 0;
}

// ---- RawNet3 (src/models/rawnet3.py) ----
int check_rawnet3_tensors(advb_handle* h) {
  return rn_check_tensors([h](const std::string& name, long long numel) { return require(h, name, numel); });
}
int build_rawnet3(advb_handle* h) {
  ADVB_TRY(check_rawnet3_tensors(h));
  return rn_build(h->rn, h->Bmax, h->T, [h](void** p, size_t bytes) {
    unsigned char* q = nullptr;
    const int r = h->alloc(&q, bytes);
    *p = q;
    return r;
  });
}
int prepare_rawnet3(advb_handle* h, cudaStream_t st) {
  rn_bind(h->rn, [h](const std::string& name) { return h->t(name); });
  return rn_prepare(h->rn, h->conv_path, st);
}

int model_prepare(advb_handle* h, cudaStream_t st) {
  // Weights are live (adversarial training mutates them between calls): by default every API call repacks them.  A host
  // that tracks weight versions (advb200/engine.py: data_ptr + tensor._version stamps) opts into "weight_cache" and
  // calls advb_invalidate_weights() when a stamp changes; unchanged weights then skip the repack kernels.
  if (h->weight_cache && h->packed_valid) return 0;
  NvtxRange nvtx("advb.prepare");
  int rc = 1;
  if (h->model_kind == ADVB_MODEL_RAWNET3) rc = prepare_rawnet3(h, st);
  else if (h->model_kind == ADVB_MODEL_LCNN) rc = prepare_lcnn(h, st);
  else if (h->model_kind == ADVB_MODEL_SPECRNET) rc = prepare_specrnet(h, st);
  else set_error("model kind not implemented");
  h->packed_valid = rc == 0;
  return rc;
}
int model_forward(advb_handle* h, const float* x, int B, cudaStream_t st) {
  if (h->model_kind == ADVB_MODEL_RAWNET3) return rn_forward(h->rn, x, h->logits, B, h->conv_path, h->tf32_passes == 2 ? 3 : h->tf32_passes, st);
  if (h->model_kind == ADVB_MODEL_LCNN) return lcnn_forward(h, x, B, st);
  if (h->model_kind == ADVB_MODEL_SPECRNET) return specrnet_forward(h, x, B, st);
  set_error("model kind not implemented");
  return 1;
}
int model_backward(advb_handle* h, const float* x, const int64_t* y, int B, int mode, int n_global, float* gx,
                   cudaStream_t st, const float* coef = nullptr, const FusedUpdate* upd = nullptr) {
  if (h->model_kind == ADVB_MODEL_RAWNET3) {
    ADVB_CHECK(upd == nullptr, "RawNet3 has no frontend backward to fuse the update into");
    return rn_backward(h->rn, x, h->logits, reinterpret_cast<const long long*>(y), B, mode, n_global, coef, gx, h->conv_path,
                       h->tf32_passes == 2 ? 3 : h->tf32_passes, st);
  }
  if (h->model_kind == ADVB_MODEL_LCNN) return lcnn_backward(h, x, y, B, mode, n_global, gx, st, coef, upd);
  if (h->model_kind == ADVB_MODEL_SPECRNET) return specrnet_backward(h, x, y, B, mode, n_global, gx, st, coef, upd);
  set_error("model kind not implemented");
  return 1;
}

int check_call(advb_handle* h, int B, int T, bool needs_model = true) {
  ADVB_CHECK(h != nullptr, "null handle");
  ADVB_CHECK(!needs_model || h->model_kind != ADVB_MODEL_FRONTEND_ONLY, "this handle holds a frontend only (no model to run)");
  ADVB_CHECK(B > 0 && B <= h->Bmax, "batch exceeds the handle's max_batch");
  ADVB_CHECK(T == h->T, "clip length differs from the handle's n_samples");
  return 0;
}

constexpr int ADVB_GRAD_SEEDED = 2;  // internal: d (sum_b coef[b] o_b) / d x

// Lazily allocated scratch: every pointer is tested on its own, so a group whose allocation failed half-way (out of
// memory) is completed - not skipped - by the next call.
#define ADVB_ENSURE(ptr, count)                         \
  do {                                                  \
    if ((ptr) == nullptr) ADVB_TRY(h->alloc(&(ptr), (count))); \
  } while (0)

int ensure_fab_scratch(advb_handle* h) {
  const size_t n = (size_t)h->Bmax * h->T;
  ADVB_ENSURE(h->fab.x1, n);
  ADVB_ENSURE(h->fab.d3, 2 * n);
  ADVB_ENSURE(h->fab.w, n);
  ADVB_ENSURE(h->fab.bh, h->Bmax);
  ADVB_ENSURE(h->fab.a0, 2 * (size_t)h->Bmax);
  ADVB_ENSURE(h->fab.res2, h->Bmax);
  return 0;
}

int ensure_cw_scratch(advb_handle* h) {
  const size_t n = (size_t)h->Bmax * h->T;
  ADVB_ENSURE(h->cw.w, n);
  ADVB_ENSURE(h->cw.m, n);
  ADVB_ENSURE(h->cw.v, n);
  ADVB_ENSURE(h->cw.adv, n);
  ADVB_ENSURE(h->cw.l2_partial, (size_t)h->Bmax * 8);
  ADVB_ENSURE(h->cw.cur_l2, h->Bmax);
  ADVB_ENSURE(h->cw.best_l2, h->Bmax);
  ADVB_ENSURE(h->cw.coef, h->Bmax);
  ADVB_ENSURE(h->cw.mask, h->Bmax);
  ADVB_ENSURE(h->cw.cost, 1);
  if (h->host_cost == nullptr) ADVB_CUDA_OK(cudaMallocHost(reinterpret_cast<void**>(&h->host_cost), sizeof(float)));
  return 0;
}

int ensure_loop_buffers(advb_handle* h) {
  const size_t n = (size_t)h->Bmax * h->T;
  ADVB_ENSURE(h->x_in, n);
  ADVB_ENSURE(h->adv2, n);
  ADVB_ENSURE(h->y_in, h->Bmax);
  if (h->cap_stream == nullptr) ADVB_CUDA_OK(cudaStreamCreateWithFlags(&h->cap_stream, cudaStreamNonBlocking));
  return 0;
}

void drop_graph(advb_handle* h) {
  auto& g = h->gc;
  if (g.done != nullptr) {
    cudaEventSynchronize(g.done);  // replays of the cached exec may still be in flight
    cudaEventDestroy(g.done);
  }
  if (g.exec != nullptr) cudaGraphExecDestroy(g.exec);
  if (g.graph != nullptr) cudaGraphDestroy(g.graph);
  g = advb_handle::GraphCache{};
}

// Enqueue `body` `reps` times on `st`.  With graphs on, `body` is captured ONCE on the engine's capture stream (PyTorch's
// current stream is usually the legacy default stream, which cannot be captured) and the instantiated graph is replayed;
// the exec is cached under `key`, which must name everything the captured launches depend on besides the engine-owned
// buffers they read and write.  Per-kernel profiling needs one event per launch, so it takes the direct path.
template <typename Body>
int replay(advb_handle* h, cudaStream_t st, int reps, const std::string& key, Body&& body) {
  if (reps <= 0) return 0;
  if (!h->use_graph || h->prof.on || reps < 2) {
    for (int i = 0; i < reps; ++i) ADVB_TRY(body(st));
    return 0;
  }
  auto& g = h->gc;
  if (g.exec == nullptr || g.key != key) {
    drop_graph(h);
    const int64_t n0 = h->counter.n;
    ADVB_CUDA_OK(cudaStreamBeginCapture(h->cap_stream, cudaStreamCaptureModeThreadLocal));
    const int rc = body(h->cap_stream);
    cudaGraph_t graph = nullptr;
    const cudaError_t e = cudaStreamEndCapture(h->cap_stream, &graph);
    g.launches = h->counter.n - n0;
    h->counter.n = n0;  // nothing ran yet
    if (rc != 0) {
      if (graph != nullptr) cudaGraphDestroy(graph);
      return rc;
    }
    ADVB_CUDA_OK(e);
    g.graph = graph;
    ADVB_CUDA_OK(cudaGraphInstantiate(&g.exec, graph, 0));
    ADVB_CUDA_OK(cudaEventCreateWithFlags(&g.done, cudaEventDisableTiming));
    g.key = key;
  }
  for (int i = 0; i < reps; ++i) ADVB_CUDA_OK(cudaGraphLaunch(g.exec, st));
  ADVB_CUDA_OK(cudaEventRecord(g.done, st));
  h->counter.n += g.launches * reps;
  return 0;
}

// One call per handle in flight: every activation / scratch buffer belongs to the handle, so a call enqueued on another
// stream than the previous one first waits for that call's completion event (calls on one stream are ordered anyway).
struct CallScope {
  DeviceGuard guard;
  advb_handle* h;
  cudaStream_t st = nullptr;
  bool ordered = false;
  explicit CallScope(advb_handle* handle) : guard(handle->device), h(handle) {
    g_counter = &h->counter;
    g_prof = &h->prof;
    g_conv_sched = h->conv_sched;
  }
  void order(cudaStream_t stream) {
    st = stream;
    ordered = true;
    if (h->last_done != nullptr && h->last_stream != stream) cudaStreamWaitEvent(stream, h->last_done, 0);
  }
  ~CallScope() {
    if (ordered) {
      if (h->last_done == nullptr) cudaEventCreateWithFlags(&h->last_done, cudaEventDisableTiming);
      if (h->last_done != nullptr) cudaEventRecord(h->last_done, st);
      h->last_stream = st;
    }
    g_counter = nullptr;
    g_prof = nullptr;
  }
};

}  // namespace

namespace advb {
namespace {
__global__ void u8_to_f32_kernel(const unsigned char* __restrict__ src, float* __restrict__ dst, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) dst[i] = (float)src[i];
}
}  // namespace
}  // namespace advb

extern "C" {

int advb_version(void) { return ADVB_VERSION; }
const char* advb_last_error(void) { return g_error.c_str(); }

int advb_create(advb_handle** out, const advb_model_desc* d) {
  ADVB_CHECK(out != nullptr && d != nullptr, "null argument");
  *out = nullptr;
  ADVB_CHECK(d->max_batch > 0 && d->n_samples >= 512, "bad max_batch / n_samples");
  int ndev = 0;
  ADVB_CUDA_OK(cudaGetDeviceCount(&ndev));
  ADVB_CHECK(d->device >= 0 && d->device < ndev, "no such CUDA device (libadvb200 has no CPU path)");
  advb_handle* h = new advb_handle();
  h->device = d->device;
  h->model_kind = d->model_kind;
  h->frontend_kind = d->frontend_kind;
  h->Bmax = d->max_batch;
  h->T = d->n_samples;
  h->F = frontend_frames(d->n_samples);
  CallScope scope(h);
  auto fail = [&]() {
    advb_destroy(h);
    return 1;
  };
  if (bind_tensors(h, d->n_tensors, d->tensors)) return fail();
  if (h->model_kind != ADVB_MODEL_LCNN && h->model_kind != ADVB_MODEL_SPECRNET && h->model_kind != ADVB_MODEL_RAWNET3 &&
      h->model_kind != ADVB_MODEL_FRONTEND_ONLY) {
    set_error("unknown model kind");
    return fail();
  }
  const size_t B = h->Bmax;
  if (h->model_kind == ADVB_MODEL_RAWNET3) {  // raw-waveform model: no spectral frontend (rawnet3.py:73-137)
    if (h->frontend_kind != ADVB_FRONTEND_NONE) {
      set_error("RawNet3 takes the raw waveform (frontend_kind must be ADVB_FRONTEND_NONE)");
      return fail();
    }
    if (h->alloc(&h->logits, B) || h->alloc(&h->grad, B * h->T) || h->alloc(&h->partial_g, B * ROW_CHUNKS) ||
        h->alloc(&h->partial_d, B * ROW_CHUNKS) || build_rawnet3(h))
      return fail();
    if (cudaDeviceSynchronize() != cudaSuccess) {
      set_error("device error during advb_create");
      return fail();
    }
    *out = h;
    return 0;
  }
  if (check_frontend_tensors(h)) return fail();
  if (h->alloc(&h->tw, 512) || h->alloc(&h->dctT, 128 * 80) || h->alloc(&h->klo, 128) || h->alloc(&h->kcnt, 128) ||
      h->alloc(&h->mlo, 257) || h->alloc(&h->mcnt, 257) || h->alloc(&h->fst.gmax_packed, 1) ||
      h->alloc(&h->fst.n_clamped, 1) || h->alloc(&h->fst.mass_total, 1) || h->alloc(&h->fst.done, 2) ||
      h->alloc(&h->dB, B * h->F * 128) || h->alloc(&h->g_dB, B * h->F * 128) || h->alloc(&h->g_coef, B * h->F * 80) ||
      h->alloc(&h->coef_tmp, B * h->F * 80) || h->alloc(reinterpret_cast<float**>(&h->spec), frontend_spec_floats(h->Bmax, h->T)) ||
      h->alloc(&h->mass_partial, (size_t)frontend_mass_blocks(h->Bmax, h->T)) || h->alloc(&h->logits, B) ||
      h->alloc(&h->grad, B * h->T) || h->alloc(&h->partial_g, B * ROW_CHUNKS) ||
      h->alloc(&h->partial_d, B * ROW_CHUNKS))
    return fail();
  if (frontend_init_constants(h->tw, 0)) return fail();
  if (h->model_kind == ADVB_MODEL_LCNN ? build_lcnn(h) : h->model_kind == ADVB_MODEL_SPECRNET ? build_specrnet(h) : 0) return fail();
  if (cudaDeviceSynchronize() != cudaSuccess) {
    set_error("device error during advb_create");
    return fail();
  }
  *out = h;
  return 0;
}

void advb_destroy(advb_handle* h) {
  if (h == nullptr) return;
  DeviceGuard guard(h->device);
  drop_graph(h);
  for (void* p : h->xr_ipc_opened) cudaIpcCloseMemHandle(p);
  if (h->cap_stream != nullptr) cudaStreamDestroy(h->cap_stream);
  if (h->last_done != nullptr) cudaEventDestroy(h->last_done);
  for (void* p : h->allocs) cudaFree(p);
  if (h->host_cost != nullptr) cudaFreeHost(h->host_cost);
  delete h;
}

size_t advb_workspace_bytes(const advb_handle* h) { return h ? h->ws_bytes : 0; }
int64_t advb_launch_count(const advb_handle* h) { return h ? h->counter.n : 0; }

int advb_set_option(advb_handle* h, const char* key, int value) {
  ADVB_CHECK(h != nullptr && key != nullptr, "null argument");
  const std::string k(key);
  h->packed_valid = false;  // conv_path / conv_sched choose which packed images are built
  if (k == "conv_path") {
    ADVB_CHECK(value == 0 || value == 1, "conv_path: 0 = tcgen05, 1 = fp32 SIMT");
    h->conv_path = value;
  } else if (k == "tf32_passes") {
    ADVB_CHECK(value == 1 || value == 2 || value == 3,
               "tf32_passes: 3 = 3xTF32, 2 = tf32 main term + bf16 cross terms in the forward 3x3 blocks (experiment), 1 = single pass");
    h->tf32_passes = value;
  } else if (k == "conv_sched") {
    ADVB_CHECK(value == 0 || value == 1, "conv_sched: 0 = persistent kernels, 1 = one-tile-per-CTA kernels");
    h->conv_sched = value;
  } else if (k == "conv0_bwd") {
    ADVB_CHECK(value == 0 || value == 1, "conv0_bwd: 0 = fp32 cell kernel, 1 = tcgen05 GEMM + col2im");
    h->conv0_bwd = value;
  } else if (k == "conv0_fwd") {
    ADVB_CHECK(value == 0 || value == 1, "conv0_fwd: 0 = Toeplitz GEMM without im2col, 1 = im2col GEMM");
    h->conv0_fwd = value;
  } else if (k == "sr_tc") {
    ADVB_CHECK(value == 0 || value == 1, "sr_tc: 1 = SpecRNet 64 -> 64 convolutions on tcgen05 (3xTF32), 0 = fp32 SIMT");
    h->sr_tc = value;
  } else if (k == "lstm_tc") {
    ADVB_CHECK(value == 0 || value == 1, "lstm_tc: 1 = BLSTM input projections on the tcgen05 GEMM, 0 = fp32 SIMT GEMM");
    h->lstm_tc = value;
  } else if (k == "fe_spec") {
    ADVB_CHECK(value == 0 || value == 1, "fe_spec: 1 = the frontend backward reads the forward's stored spectra, 0 = recomputes the STFT");
    h->fe_spec = value;
  } else if (k == "graph") {
    ADVB_CHECK(value == 0 || value == 1, "graph: 1 = replay the attack iteration as a CUDA graph, 0 = enqueue every kernel");
    h->use_graph = value;
  } else if (k == "fuse_update") {
    ADVB_CHECK(value == 0 || value == 1, "fuse_update: 1 = FGSM / PGD update in the frontend backward's epilogue, 0 = own kernel");
    h->fuse_update = value;
  } else if (k == "weight_cache") {
    ADVB_CHECK(value == 0 || value == 1, "weight_cache: 1 = skip the weight repack until advb_invalidate_weights()");
    h->weight_cache = value;
  } else {
    set_error("unknown option '" + k + "'");
    return 1;
  }
  return 0;
}

// ---- strict multi-GPU mode: the batch-wide dB floor spans every rank's clips (frontend.cuh, FrontendXRank) -------------
int advb_xrank_export(advb_handle* h, unsigned char* ipc_handle, void** local_ptr) {
  ADVB_CHECK(h != nullptr, "null handle");
  ADVB_CHECK(h->frontend_kind != ADVB_FRONTEND_NONE, "strict mode couples the dB floor of a spectral frontend; this model has none");
  static_assert(sizeof(cudaIpcMemHandle_t) == ADVB_XRANK_HANDLE_BYTES, "cudaIpcMemHandle_t is 64 bytes");
  DeviceGuard guard(h->device);
  CallScope scope(h);
  ADVB_CUDA_OK(cudaDeviceSynchronize());
  if (h->xr_mailbox == nullptr) {
    ADVB_TRY(h->alloc(&h->xr_mailbox, 2 * XR_MAX_RANKS));
    ADVB_TRY(h->alloc(&h->fst.xr.epoch, 1));
    ADVB_TRY(h->alloc(&h->fst.xr.timed_out, 1));
  }
  // every rank clears its mailbox and exchange counter BEFORE the handles are exchanged (the exchange is the barrier), so no
  // peer can have written into it yet
  ADVB_CUDA_OK(cudaMemset(h->xr_mailbox, 0, 2 * XR_MAX_RANKS * sizeof(unsigned long long)));
  ADVB_CUDA_OK(cudaMemset(h->fst.xr.epoch, 0, sizeof(unsigned)));
  ADVB_CUDA_OK(cudaMemset(h->fst.xr.timed_out, 0, sizeof(int)));
  ADVB_CUDA_OK(cudaDeviceSynchronize());
  if (ipc_handle != nullptr) {
    cudaIpcMemHandle_t ipc;
    ADVB_CUDA_OK(cudaIpcGetMemHandle(&ipc, h->xr_mailbox));
    memcpy(ipc_handle, &ipc, sizeof(ipc));
  }
  if (local_ptr != nullptr) *local_ptr = h->xr_mailbox;
  return 0;
}

int advb_xrank_connect(advb_handle* h, int rank, int world, const unsigned char* ipc_handles, void* const* local_ptrs) {
  ADVB_CHECK(h != nullptr, "null handle");
  DeviceGuard guard(h->device);
  CallScope scope(h);
  ADVB_CUDA_OK(cudaDeviceSynchronize());
  drop_graph(h);
  h->xr_gen++;
  for (void* p : h->xr_ipc_opened) cudaIpcCloseMemHandle(p);
  h->xr_ipc_opened.clear();
  h->fst.xr.world = 1;
  h->fst.xr.rank = 0;
  if (world <= 1) return 0;  // back to per-shard floors
  ADVB_CHECK(world <= XR_MAX_RANKS && rank >= 0 && rank < world, "strict mode: 2..8 ranks, 0 <= rank < world");
  ADVB_CHECK(h->xr_mailbox != nullptr, "advb_xrank_export first");
  ADVB_CHECK(ipc_handles != nullptr || local_ptrs != nullptr, "peer mailboxes: IPC handles (other processes) or device pointers (this process)");
  for (int r = 0; r < world; ++r) {
    void* p = nullptr;
    if (r == rank) {
      p = h->xr_mailbox;
    } else if (local_ptrs != nullptr && local_ptrs[r] != nullptr) {
      p = local_ptrs[r];
      cudaPointerAttributes attr{};
      ADVB_CUDA_OK(cudaPointerGetAttributes(&attr, p));
      if (attr.device != h->device) {
        int can = 0;
        ADVB_CUDA_OK(cudaDeviceCanAccessPeer(&can, h->device, attr.device));
        ADVB_CHECK(can != 0, "strict mode needs peer access between the ranks' devices");
        cudaError_t e = cudaDeviceEnablePeerAccess(attr.device, 0);
        if (e == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();
        else ADVB_CUDA_OK(e);
      }
    } else {
      ADVB_CHECK(ipc_handles != nullptr, "no mailbox given for a peer rank");
      cudaIpcMemHandle_t ipc;
      memcpy(&ipc, ipc_handles + (size_t)r * ADVB_XRANK_HANDLE_BYTES, sizeof(ipc));
      ADVB_CUDA_OK(cudaIpcOpenMemHandle(&p, ipc, cudaIpcMemLazyEnablePeerAccess));
      h->xr_ipc_opened.push_back(p);
    }
    h->fst.xr.mailbox[r] = static_cast<unsigned long long*>(p);
  }
  h->fst.xr.world = world;
  h->fst.xr.rank = rank;
  return 0;
}

int advb_xrank_status(advb_handle* h, int* timed_out) {
  ADVB_CHECK(h != nullptr && timed_out != nullptr, "null argument");
  *timed_out = 0;
  if (h->fst.xr.timed_out == nullptr) return 0;
  DeviceGuard guard(h->device);
  ADVB_CUDA_OK(cudaMemcpy(timed_out, h->fst.xr.timed_out, sizeof(int), cudaMemcpyDeviceToHost));
  return 0;
}

int advb_rebind(advb_handle* h, int n_tensors, const advb_tensor_ref* tensors) {
  ADVB_CHECK(h != nullptr, "null handle");
  ADVB_TRY(bind_tensors(h, n_tensors, tensors));
  if (h->model_kind == ADVB_MODEL_RAWNET3) return check_rawnet3_tensors(h);
  ADVB_TRY(check_frontend_tensors(h));
  if (h->model_kind == ADVB_MODEL_LCNN) ADVB_TRY(check_lcnn_tensors(h));
  if (h->model_kind == ADVB_MODEL_SPECRNET) ADVB_TRY(check_specrnet_tensors(h));
  return 0;
}

int advb_forward(advb_handle* h, const float* x, float* logits, int B, int T, void* cuda_stream) {
  ADVB_TRY(check_call(h, B, T));
  CallScope scope(h);
  cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
  scope.order(st);
  ADVB_TRY(model_prepare(h, st));
  ADVB_TRY(model_forward(h, x, B, st));
  ADVB_CUDA_OK(cudaMemcpyAsync(logits, h->logits, (size_t)B * sizeof(float), cudaMemcpyDeviceToDevice, st));
  return 0;
}

int advb_grad(advb_handle* h, int what, const float* x, const int64_t* y, float* grad, float* logits, int B, int T,
              int n_global_batch, void* cuda_stream) {
  ADVB_TRY(check_call(h, B, T));
  ADVB_CHECK(what == ADVB_GRAD_CE || what == ADVB_GRAD_LOGIT, "bad gradient kind");
  ADVB_CHECK(what == ADVB_GRAD_LOGIT || y != nullptr, "labels required");
  CallScope scope(h);
  cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
  scope.order(st);
  ADVB_TRY(model_prepare(h, st));
  ADVB_TRY(model_forward(h, x, B, st));
  ADVB_TRY(model_backward(h, x, y, B, what, n_global_batch > 0 ? n_global_batch : B, grad, st));
  if (logits != nullptr)
    ADVB_CUDA_OK(cudaMemcpyAsync(logits, h->logits, (size_t)B * sizeof(float), cudaMemcpyDeviceToDevice, st));
  return 0;
}

}  // extern "C"

namespace {

std::string graph_key(const advb_handle* h, const advb_attack_desc* atk, int B, int n_global, int fused) {
  auto bits = [](float f) {
    uint32_t u;
    memcpy(&u, &f, sizeof(u));
    return std::to_string(u);
  };
  return std::to_string(atk->kind) + "|" + std::to_string(B) + "|" + bits(atk->eps) + "|" + bits(atk->alpha) + "|" +
         bits(atk->eps_div) + "|" + std::to_string(n_global) + "|" + std::to_string(fused) + "|" +
         std::to_string(h->conv_path) + "|" + std::to_string(h->tf32_passes) + "|" + std::to_string(h->conv_sched) + "|" +
         std::to_string(h->conv0_bwd) + "|" + std::to_string(h->conv0_fwd) + "|" + std::to_string(h->fe_spec) + "|" + std::to_string(h->lstm_tc) + "|" + std::to_string(h->sr_tc) + "|" +
         std::to_string(h->bind_epoch) + "|" + std::to_string(h->xr_gen);
}

// The attack loops (caller holds a CallScope and has validated the arguments).  `minmax`: x / x_adv are raw waveforms and
// to_minmax / revert_minmax (src/aa/utils.py:4-14) replace the copy-in / copy-out of the engine-owned loop buffers.
int run_attack(advb_handle* h, const advb_attack_desc* atk, const float* x, const int64_t* y, const float* start, float* x_adv,
               int B, int T, cudaStream_t st, bool minmax) {
  const int n_global = atk->n_global_batch > 0 ? atk->n_global_batch : B;
  const int64_t n = (int64_t)B * T;
  const size_t bytes = (size_t)n * sizeof(float);
  const bool has_frontend = h->model_kind != ADVB_MODEL_RAWNET3;
  const bool fused = h->fuse_update && has_frontend;
  // targeted mode (fgsm.py:49-50, pgd.py:64-65, pgdl2.py:69-70): cost = -loss(outputs, target_labels), i.e. the same
  // gradient negated: the ascent step becomes a descent step on the target labels' loss
  const float dir = atk->targeted ? -1.0f : 1.0f;
  if (atk->targeted && atk->kind != ADVB_ATTACK_CW) y = atk->target_labels;
  ADVB_TRY(model_prepare(h, st));
  if (minmax || atk->kind == ADVB_ATTACK_PGD || atk->kind == ADVB_ATTACK_PGDL2) ADVB_TRY(ensure_loop_buffers(h));
  if (minmax) {
    ADVB_ENSURE(h->mm_mn, h->Bmax);
    ADVB_ENSURE(h->mm_mx, h->Bmax);
  }
  // single-evaluation / host-driven attacks under minmax: scale into x_in, attack into adv2, revert into x_adv
  const float* xc = x;
  float* out = x_adv;
  if (minmax && atk->kind != ADVB_ATTACK_PGD && atk->kind != ADVB_ATTACK_PGDL2) {
    ADVB_TRY(minmax_scale(x, h->x_in, h->mm_mn, h->mm_mx, B, T, st));
    xc = h->x_in;
    out = h->adv2;
  }
  switch (atk->kind) {
    case ADVB_ATTACK_FGSM: {
      NvtxRange nvtx("advb.fgsm");
      ADVB_TRY(model_forward(h, xc, B, st));
      if (fused) {
        FusedUpdate u;
        u.kind = 1, u.x_clean = xc, u.adv_out = out, u.eps = dir * atk->eps;
        ADVB_TRY(model_backward(h, xc, y, B, ADVB_GRAD_CE, n_global, h->grad, st, nullptr, &u));
      } else {
        ADVB_TRY(model_backward(h, xc, y, B, ADVB_GRAD_CE, n_global, h->grad, st));
        ADVB_TRY(fgsm_step(xc, h->grad, out, dir * atk->eps, n, st));
      }
      break;
    }
    case ADVB_ATTACK_PGD:
    case ADVB_ATTACK_PGDL2: {
      // pgd.py:59-76 / pgdl2.py:64-88.  The loop runs on engine-owned buffers only (clean clips x_in, labels y_in, iterates
      // grad / adv2), so one captured iteration is valid for every later call with the same parameters.
      ADVB_CHECK(atk->steps >= 0, "bad step count");
      NvtxRange nvtx(atk->kind == ADVB_ATTACK_PGD ? "advb.pgd" : "advb.pgdl2");
      if (minmax) ADVB_TRY(minmax_scale(x, h->x_in, h->mm_mn, h->mm_mx, B, T, st));
      else ADVB_CUDA_OK(cudaMemcpyAsync(h->x_in, x, bytes, cudaMemcpyDeviceToDevice, st));
      ADVB_CUDA_OK(cudaMemcpyAsync(h->y_in, y, (size_t)B * sizeof(int64_t), cudaMemcpyDeviceToDevice, st));
      if (has_frontend) ADVB_TRY(frontend_clean(h->fst, &h->fe_host, st));  // the captured body assumes (and leaves) a clean state
      const float* xin = h->x_in;
      const int64_t* yin = h->y_in;
      const float eps = atk->eps, alpha = dir * atk->alpha, eps_div = atk->eps_div;
      const bool pgd_fused = fused && atk->kind == ADVB_ATTACK_PGD;
      const std::string key = graph_key(h, atk, B, n_global, pgd_fused);
      float* result = nullptr;
      if (pgd_fused) {
        // ping-pong iterates pp[0] = grad (no gradient is ever stored), pp[1] = adv2; the update rule runs in the
        // epilogue of the frontend backward.  The graph holds two iterations pp[0] -> pp[1] -> pp[0].
        float* pp[2] = {h->grad, h->adv2};
        auto iter = [&](int from, cudaStream_t s) -> int {
          FusedUpdate u;
          u.kind = 2, u.x_clean = xin, u.adv_out = pp[from ^ 1], u.eps = eps, u.alpha = alpha;
          ADVB_TRY(model_forward(h, pp[from], B, s));
          ADVB_TRY(model_backward(h, pp[from], yin, B, ADVB_GRAD_CE, n_global, pp[from ^ 1], s, nullptr, &u));
          return 0;
        };
        const int odd = atk->steps & 1;
        ADVB_TRY(pgd_start(xin, start, pp[odd], n, st));
        if (odd) ADVB_TRY(iter(1, st));
        ADVB_TRY(replay(h, st, atk->steps / 2, key, [&](cudaStream_t s) -> int {
          ADVB_TRY(iter(0, s));
          ADVB_TRY(iter(1, s));
          return 0;
        }));
        result = pp[0];
      } else {
        float* adv = h->adv2;
        ADVB_TRY(pgd_start(xin, start, adv, n, st));
        ADVB_TRY(replay(h, st, atk->steps, key, [&](cudaStream_t s) -> int {
          ADVB_TRY(model_forward(h, adv, B, s));
          ADVB_TRY(model_backward(h, adv, yin, B, ADVB_GRAD_CE, n_global, h->grad, s));
          if (atk->kind == ADVB_ATTACK_PGD) ADVB_TRY(pgd_step(xin, h->grad, adv, eps, alpha, n, s));
          else ADVB_TRY(pgdl2_step(xin, h->grad, adv, eps, alpha, eps_div, B, T, h->partial_g, h->partial_d, s));
          return 0;
        }));
        result = adv;
      }
      if (minmax) ADVB_TRY(minmax_revert(result, h->mm_mn, h->mm_mx, x_adv, B, T, st));
      else ADVB_CUDA_OK(cudaMemcpyAsync(x_adv, result, bytes, cudaMemcpyDeviceToDevice, st));
      return 0;
    }
    case ADVB_ATTACK_FAB: {
      // attack_single_run (fab.py:131-307) on a batch of correctly classified clips: L-inf or L2 (atk->norm), untargeted.  `start` = the
      // random restart point x1 of fab.py:176-206 (nullable: x1 = x).  One backward per step (the class-0 gradient of
      // z = [-o, o] is the exact negation of class 1's).
      ADVB_CHECK(atk->steps >= 0, "bad step count");
      ADVB_CHECK(atk->norm == ADVB_NORM_LINF || atk->norm == ADVB_NORM_L2, "FAB norm must be ADVB_NORM_LINF or ADVB_NORM_L2");
      const int norm_l2 = atk->norm == ADVB_NORM_L2 ? 1 : 0;
      ADVB_CHECK(!minmax || start == nullptr, "FAB random restarts are not available through the min-max entry point");
      ADVB_TRY(ensure_fab_scratch(h));
      const long long* yl = reinterpret_cast<const long long*>(y);
      NvtxRange nvtx("advb.fab");
      ADVB_TRY(fab_init(xc, out, h->fab, B, T, st));
      if (start != nullptr) ADVB_CUDA_OK(cudaMemcpyAsync(h->fab.x1, start, bytes, cudaMemcpyDeviceToDevice, st));
      for (int it = 0; it < atk->steps; ++it) {
        ADVB_TRY(model_forward(h, h->fab.x1, B, st));
        ADVB_TRY(model_backward(h, h->fab.x1, y, B, ADVB_GRAD_LOGIT, n_global, h->grad, st));
        ADVB_TRY(fab_hyperplane(h->grad, h->logits, yl, h->fab, B, T, st));
        ADVB_TRY(fab_project(xc, h->fab, B, T, norm_l2, st));
        ADVB_TRY(fab_combine(xc, h->fab, atk->eta, atk->alpha_max, B, T, st));
        ADVB_TRY(model_forward(h, h->fab.x1, B, st));
        ADVB_TRY(fab_bookkeep(xc, h->logits, yl, out, h->fab, atk->beta, B, T, norm_l2, st));
      }
      break;
    }
    case ADVB_ATTACK_CW: {
      ADVB_CHECK(atk->steps >= 0, "bad step count");
      ADVB_TRY(ensure_cw_scratch(h));
      const long long* yl = reinterpret_cast<const long long*>(y);
      const long long* yt = atk->targeted ? reinterpret_cast<const long long*>(atk->target_labels) : nullptr;
      NvtxRange nvtx("advb.cw");
      ADVB_TRY(cw_init(xc, out, h->cw, B, T, st));
      float prev_cost = 1e10f;
      const int every = atk->steps / 10 > 1 ? atk->steps / 10 : 1;
      for (int step = 0; step < atk->steps; ++step) {
        ADVB_TRY(cw_forward_image(xc, h->cw, B, T, st));
        ADVB_TRY(model_forward(h, h->cw.adv, B, st));
        ADVB_TRY(cw_head(h->logits, yl, yt, h->cw, atk->c, atk->kappa, B, st));
        ADVB_TRY(model_backward(h, h->cw.adv, y, B, ADVB_GRAD_SEEDED, n_global, h->grad, st, h->cw.coef));
        ADVB_TRY(cw_adam(xc, h->grad, out, h->cw, atk->lr, step + 1, B, T, st));
        if (step % every == 0) {  // batch-wide early stop: the one host sync the reference has too (cw.py:107-110)
          ADVB_CUDA_OK(cudaMemcpyAsync(h->host_cost, h->cw.cost, sizeof(float), cudaMemcpyDeviceToHost, st));
          ADVB_CUDA_OK(cudaStreamSynchronize(st));
          if (*h->host_cost > prev_cost) break;
          prev_cost = *h->host_cost;
        }
      }
      break;
    }
    default:
      set_error("attack kind not implemented in the native loop");
      return 1;
  }
  if (minmax) ADVB_TRY(minmax_revert(out, h->mm_mn, h->mm_mx, x_adv, B, T, st));
  return 0;
}

int check_attack_args(advb_handle* h, const advb_attack_desc* atk, const float* x, const int64_t* y, const float* x_adv, int B,
                      int T) {
  ADVB_TRY(check_call(h, B, T));
  ADVB_CHECK(atk != nullptr && x != nullptr && y != nullptr && x_adv != nullptr, "null argument");
  ADVB_CHECK(!atk->targeted || atk->target_labels != nullptr, "targeted mode needs target_labels");
  ADVB_CHECK(!atk->targeted || atk->kind != ADVB_ATTACK_FAB, "FAB has no targeted mode in the reference's patched copy (fab.py:63)");
  ADVB_CHECK(h->fst.xr.world <= 1 || (atk->kind != ADVB_ATTACK_FAB && atk->kind != ADVB_ATTACK_CW),
             "strict multi-GPU mode needs the same call sequence on every rank: FAB (per-rank clip selection) and CW (per-rank early "
             "stop) are not supported under it");
  return 0;
}
}  // namespace

extern "C" {

int advb_attack(advb_handle* h, const advb_attack_desc* atk, const float* x, const int64_t* y, const float* start,
                float* x_adv, int B, int T, void* cuda_stream) {
  ADVB_TRY(check_attack_args(h, atk, x, y, x_adv, B, T));
  ADVB_CHECK(x != x_adv, "x_adv must not alias x");
  CallScope scope(h);
  scope.order(static_cast<cudaStream_t>(cuda_stream));
  return run_attack(h, atk, x, y, start, x_adv, B, T, static_cast<cudaStream_t>(cuda_stream), false);
}

int advb_attack_minmax(advb_handle* h, const advb_attack_desc* atk, const float* x_raw, const int64_t* y, const float* start,
                       float* x_adv_raw, int B, int T, void* cuda_stream) {
  ADVB_TRY(check_attack_args(h, atk, x_raw, y, x_adv_raw, B, T));
  CallScope scope(h);
  scope.order(static_cast<cudaStream_t>(cuda_stream));
  return run_attack(h, atk, x_raw, y, start, x_adv_raw, B, T, static_cast<cudaStream_t>(cuda_stream), true);
}

int advb_invalidate_weights(advb_handle* h) {
  ADVB_CHECK(h != nullptr, "null handle");
  h->packed_valid = false;
  return 0;
}

static int projection_rows(const float* t, const float* w, const float* b, float* d, int R, int T, int norm_l2, void* cuda_stream);

int advb_projection_linf(const float* t, const float* w, const float* b, float* d, int R, int T, void* cuda_stream) {
  return projection_rows(t, w, b, d, R, T, 0, cuda_stream);
}

int advb_projection_l2(const float* t, const float* w, const float* b, float* d, int R, int T, void* cuda_stream) {
  return projection_rows(t, w, b, d, R, T, 1, cuda_stream);
}

static int projection_rows(const float* t, const float* w, const float* b, float* d, int R, int T, int norm_l2, void* cuda_stream) {
  ADVB_CHECK(t && w && b && d && R > 0 && T > 0, "bad argument");
  cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
  // rows are independent: run them as the "x1" half of a FAB step whose second half is empty
  float* a0 = nullptr;
  ADVB_CUDA_OK(cudaMallocAsync(reinterpret_cast<void**>(&a0), sizeof(float) * R, st));
  FabScratch s{};
  s.x1 = const_cast<float*>(t);
  s.w = const_cast<float*>(w);
  s.bh = const_cast<float*>(b);
  s.d3 = d;
  s.a0 = a0;
  const int rc = fab_project_rows(s, R, T, norm_l2, st);
  cudaFreeAsync(a0, st);
  return rc;
}

int advb_mel_spec_fwd(const float* x, const float* fb, int n_mels, float* out, int B, int T, void* cuda_stream) {
  ADVB_CHECK(x && fb && out && B > 0 && T > 0, "bad argument");
  return frontend_mel_spec(x, fb, n_mels, out, B, T, static_cast<cudaStream_t>(cuda_stream));
}

int advb_row_diff_norms(const float* a, const float* b, float* linf, float* l2, int B, int T, void* cuda_stream) {
  ADVB_CHECK(a && b && (linf || l2) && B > 0 && T > 0, "bad argument");
  return row_diff_norms(a, b, linf, l2, B, T, static_cast<cudaStream_t>(cuda_stream));
}

int advb_frontend_fwd(advb_handle* h, const float* x, float* coeff, int B, int T, void* cuda_stream) {
  ADVB_TRY(check_call(h, B, T, false));
  ADVB_CHECK(h->frontend_kind != ADVB_FRONTEND_NONE, "this model takes the raw waveform: it has no spectral frontend");
  CallScope scope(h);
  cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
  scope.order(st);
  refresh_frontend_tables(h);
  ADVB_TRY(frontend_prepare(h->ftb, st));
  ADVB_TRY(frontend_forward(h->ftb, h->fst, x, B, T, h->dB, coeff, (long long)80 * h->F, 1, h->F, 0, st, nullptr, &h->fe_host));
  return 0;
}

int advb_frontend_bwd(advb_handle* h, const float* x, const float* g_coeff, float* g_x, int B, int T,
                      void* cuda_stream) {
  ADVB_TRY(check_call(h, B, T, false));
  ADVB_CHECK(h->frontend_kind != ADVB_FRONTEND_NONE, "this model takes the raw waveform: it has no spectral frontend");
  CallScope scope(h);
  cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
  scope.order(st);
  refresh_frontend_tables(h);
  ADVB_TRY(frontend_prepare(h->ftb, st));
  // forward first: backward needs the batch arg-max / floor state of this input
  ADVB_TRY(frontend_forward(h->ftb, h->fst, x, B, T, h->dB, h->coef_tmp, (long long)80 * h->F, 1, h->F, 0, st, nullptr, &h->fe_host));
  ADVB_TRY(frontend_backward(h->ftb, h->fst, x, B, T, h->dB, g_coeff, (long long)80 * h->F, 1, h->F, h->mass_partial,
                             h->g_dB, g_x, st, nullptr, nullptr, &h->fe_host));
  return 0;
}

int advb_minmax(const float* x, float* x01, float* mn, float* mx, int B, int T, void* cuda_stream) {
  ADVB_CHECK(x && x01 && mn && mx && B > 0 && T > 0, "bad argument");
  return minmax_scale(x, x01, mn, mx, B, T, static_cast<cudaStream_t>(cuda_stream));
}
int advb_revert_minmax(const float* x01, const float* mn, const float* mx, float* x, int B, int T, void* cuda_stream) {
  ADVB_CHECK(x && x01 && mn && mx && B > 0 && T > 0, "bad argument");
  return minmax_revert(x01, mn, mx, x, B, T, static_cast<cudaStream_t>(cuda_stream));
}

int advb_profile_begin(advb_handle* h, void* cuda_stream) {
  ADVB_CHECK(h != nullptr, "null handle");
  CallScope scope(h);
  for (auto& m : h->prof.marks) cudaEventDestroy(m.second);
  h->prof.marks.clear();
  h->prof.on = true;
  prof_mark("begin", static_cast<cudaStream_t>(cuda_stream));
  return 0;
}

int64_t advb_profile_end(advb_handle* h, char* buf, int64_t capacity) {
  if (h == nullptr) return -1;
  DeviceGuard guard(h->device);
  h->prof.on = false;
  auto& mk = h->prof.marks;
  std::map<std::string, std::pair<double, int64_t>> agg;
  if (!mk.empty()) cudaEventSynchronize(mk.back().second);
  for (size_t i = 1; i < mk.size(); ++i) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, mk[i - 1].second, mk[i].second) != cudaSuccess) continue;
    auto& a = agg[mk[i].first];
    a.first += ms;
    a.second += 1;
  }
  for (auto& m : mk) cudaEventDestroy(m.second);
  mk.clear();
  std::string js = "[";
  bool first = true;
  for (auto& kv : agg) {
    if (!first) js += ",";
    first = false;
    js += "{\"name\":\"" + kv.first + "\",\"count\":" + std::to_string(kv.second.second) + ",\"total_ms\":" +
          std::to_string(kv.second.first) + "}";
  }
  js += "]";
  if (buf != nullptr && capacity > 0) {
    const size_t n = js.size() < (size_t)capacity - 1 ? js.size() : (size_t)capacity - 1;
    memcpy(buf, js.data(), n);
    buf[n] = 0;
  }
  return (int64_t)js.size() + 1;
}

int64_t advb_debug_stage(advb_handle* h, const char* stage, float* dst, int64_t capacity, int64_t dims[5],
                         void* cuda_stream) {
  if (h == nullptr || stage == nullptr) {
    set_error("null argument");
    return -1;
  }
  CallScope scope(h);
  cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
  scope.order(st);
  const std::string s(stage);
  const float* src = nullptr;
  const unsigned char* src_u8 = nullptr;  // winner codes: converted to float on the way out
  int64_t d[5] = {h->Bmax, 1, 1, 1, 0};
  auto from_act = [&](const Act& a) {
    src = a.p;
    d[1] = a.H;
    d[2] = a.W;
    d[3] = a.C;
    d[4] = a.pad;
  };
  const bool is_sr = h->model_kind == ADVB_MODEL_SPECRNET;
  auto sr_idx = [&](char c) { return c == '0' ? 0 : c == '2' ? 1 : c == '4' ? 2 : -1; };
  if (h->model_kind == ADVB_MODEL_RAWNET3) {
    int rows = 0, cols = 0;
    src = rn_stage(h->rn, s, &rows, &cols);
    if (src == nullptr) {
      set_error("unknown stage '" + s + "'");
      return -1;
    }
    if (s == "rn_filt") d[0] = 1;                  // one filter bank, shared by every clip
    d[1] = rows, d[2] = 1, d[3] = cols, d[4] = 0;  // raw rows (zero border rows included); W = 1, no reported border
  } else if (is_sr && s == "frontend") {
    src = h->sr_img;
    d[1] = h->F, d[2] = 80, d[3] = 1, d[4] = 1;
  } else if (is_sr && s.size() == 6 && (s.rfind("sr_xb", 0) == 0 || s.rfind("sr_xn", 0) == 0 || s.rfind("sr_gn", 0) == 0) &&
             sr_idx(s[5]) >= 0) {
    const SrBlock& k = h->sr[sr_idx(s[5])];
    const bool xb = s[4] == 'b', xn = s[3] == 'x' && s[4] == 'n';
    src = xb ? k.xb : xn ? k.xn : k.g_xn;
    d[1] = xb ? k.Hb : k.Hn, d[2] = xb ? k.Wb : k.Wn, d[3] = k.C, d[4] = xn ? k.xn_pad : 0;
  } else if (s == "frontend") from_act(h->act0);
  else if (s.rfind("block", 0) == 0 && s.size() == 6 && s[5] >= '0' && s[5] <= '8') from_act(h->blk[s[5] - '0'].out);
  else if (s.rfind("gblock", 0) == 0 && s.size() == 7 && s[6] >= '0' && s[6] <= '8') {
    const LcnnBlock& k = h->blk[s[6] - '0'];
    src = k.gout;
    d[1] = k.Ho;
    d[2] = k.Wo;
    d[3] = k.Cout / 2;
  } else if (!is_sr && s.rfind("codes", 0) == 0 && s.size() == 6 && s[5] >= '0' && s[5] <= '8') {
    // per output element of block N: (Max-Feature-Map half of the winning pixel) << 2 | (2x2 pool position dy << 1 | dx)
    const LcnnBlock& k = h->blk[s[5] - '0'];
    src_u8 = k.codes;
    d[1] = k.Ho;
    d[2] = k.Wo;
    d[3] = k.Cout / 2;
  } else if (s == "feats" || s == "lstm1" || s == "lstm2" || s == "dfeats") {
    // features / their gradient in the reference's (c * Wf + w) order: gathered on demand from the NHWC block buffers
    if (s == "feats" && feats_gather(h->blk[8].out.p, h->feats, h->Bmax, h->L, h->Wf, 32, st)) return -1;
    if (s == "dfeats" && feats_gather(h->blk[8].gout, h->feats, h->Bmax, h->L, h->Wf, 32, st)) return -1;
    src = (s == "feats" || s == "dfeats") ? h->feats : s == "lstm1" ? h->l1 : h->l2;
    d[1] = h->L;
    d[2] = 1;
    d[3] = 160;
  } else if (s == "gcoef") {
    src = h->g_coef;
    d[1] = h->F;
    d[2] = 80;
    d[3] = 1;
  } else if (s == "dB") {
    src = h->dB;
    d[1] = h->F;
    d[2] = 1;
    d[3] = 128;
  } else {
    set_error("unknown stage '" + s + "'");
    return -1;
  }
  const int64_t n = d[0] * (d[1] + 2 * d[4]) * (d[2] + 2 * d[4]) * d[3];
  if (dims != nullptr)
    for (int i = 0; i < 5; ++i) dims[i] = d[i];
  if (dst == nullptr) return n;
  if (capacity < n) {
    set_error("debug buffer too small");
    return -1;
  }
  if (src_u8 != nullptr) {
    advb::u8_to_f32_kernel<<<1024, 256, 0, st>>>(src_u8, dst, n);
    if (cudaGetLastError() != cudaSuccess) {
      set_error("debug copy failed");
      return -1;
    }
    return n;
  }
  if (cudaMemcpyAsync(dst, src, (size_t)n * sizeof(float), cudaMemcpyDeviceToDevice, st) != cudaSuccess) {
    set_error("debug copy failed");
    return -1;
  }
  return n;
}

}  // extern "C"
