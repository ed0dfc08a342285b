// Frontend (LFCC/MFCC) host entry points — see frontend.cu.
#pragma once
#include "common.cuh"

namespace advb {

struct FrontendTables {
  const float* fb;      // (257,128) live buffer of the model (frontend.filter_mat / ...mel_scale.fb)
  const float* dct;     // (128,80)  live buffer (frontend.dct_mat)
  const float* window;  // (400)     live buffer (Hann)
  const float2* tw;     // 512  (cos, -sin)(2 pi n / 512)   (engine-owned constants)
  float* dctT;          // (80,128) transpose of dct            (rebuilt from dct on every call)
  int* klo;             // 128: first non-zero bin of each filter   (rebuilt from fb on every call)
  int* kcnt;            // 128: number of bins spanned
  int* mlo;             // 257: first filter touching each bin
  int* mcnt;            // 257
};

// Strict multi-GPU mode (SURVEY.md §8(e) ii): the top_db floor of amplitude_to_DB is relative to the maximum of the WHOLE
// batch (F5), so a clip-sharded run reproduces the single-device result only if the ranks agree on that maximum (forward) and
// on the summed gradient of the clamped elements, which autograd routes to the arg-max element (backward).  Both are one float
// per rank: each rank stores (value << 32 | exchange number) straight into every peer's mailbox over NVLink (peer memory mapped
// with CUDA IPC; one 8-byte store is atomic) and polls its own mailbox until every rank's word carries the current exchange
// number.  No NCCL call, no host round trip: the exchange is a 32-thread kernel node of the replayed CUDA graph.
constexpr int XR_MAX_RANKS = 8;
struct FrontendXRank {
  int world = 1, rank = 0;                               // world == 1: exchange disabled
  unsigned long long* mailbox[XR_MAX_RANKS] = {};        // mailbox[r]: rank r's [2 parities][XR_MAX_RANKS] words (mailbox[rank] is local)
  unsigned* epoch = nullptr;                             // local: exchanges done so far (all ranks make the same sequence of calls)
  int* timed_out = nullptr;                              // local: set when a peer did not arrive within XR_TIMEOUT_NS
};

struct FrontendState {
  unsigned long long* gmax_packed;  // batch arg-max of the dB tensor: (ordered float key << 32) | ~index
  int* n_clamped;                   // > 0 iff the top_db floor clamped anything in the last forward
  float* mass_total;                // backward: summed gradient of clamped elements
  unsigned* done;                   // [2] "last block" counters: fe_dct_t (mass reduction), fe_bwd (state reset)
  FrontendXRank xr;                 // strict multi-GPU mode; world == 1 otherwise
};

// Optional epilogue of the backward: the attack's element-wise update rule (fgsm.py:59-60, pgd.py:74-76) applied to each
// gradient sample while it is still in a register, so that the waveform gradient never reaches HBM.  `adv_out` must not
// alias the waveform the backward reads (other tiles re-read its halo): the PGD loop ping-pongs two buffers.
struct FusedUpdate {
  int kind = 0;                 // 0 = none (store the gradient), 1 = FGSM sign step, 2 = PGD L-inf step
  const float* x_clean = nullptr;  // (B,T) clean clips
  float* adv_out = nullptr;        // (B,T) next iterate
  float eps = 0.f, alpha = 0.f;
};

int frontend_frames(int T);
int frontend_mass_blocks(int B, int T);
// Host-side bookkeeping of the device state: the backward's last block resets it, so only a forward that follows another
// forward needs the reset kernel.
struct FrontendHost {
  bool dirty = false;  // a forward has accumulated into the state and no backward has reset it yet
};
int frontend_init_constants(float2* tw, cudaStream_t stream);
int frontend_prepare(const FrontendTables& tb, cudaStream_t stream);

// out[b*clip_stride + offset + f*stride_f + c*stride_c] = coefficient c of frame f
int frontend_forward(const FrontendTables& tb, const FrontendState& st, const float* x, int B, int T, float* dB,
                     float* out, long long clip_stride, long long stride_f, long long stride_c, long long offset,
                     cudaStream_t stream, float2* spec = nullptr, FrontendHost* host = nullptr);
// gcoef read through the same kind of strides (no offset: pass the shifted pointer); gd: (B,F,128) scratch for the
// gradient of the dB tensor; gx (B,T)
int frontend_backward(const FrontendTables& tb, const FrontendState& st, const float* x, int B, int T,
                      const float* dB, const float* gcoef, long long g_clip_stride, long long g_stride_f,
                      long long g_stride_c, float* mass_partial, float* gd, float* gx, cudaStream_t stream,
                      const FusedUpdate* upd = nullptr, const float2* spec = nullptr, FrontendHost* host = nullptr);
// src/frontends.py:53-79 (`mel_spec`): out (B, 2, n_mels, F) = |.| and angle of (fb^T Re STFT, fb^T Im STFT); fb (257, n_mels) live
// buffer (MEL_SCALE_FN.fb).  Stateless; forward only (no model of the reference consumes this 2-channel frontend).
int frontend_mel_spec(const float* x, const float* fb, int n_mels, float* out, int B, int T, cudaStream_t stream);
// reset the state now if a forward left it dirty (before a captured loop whose body assumes a clean state)
int frontend_clean(const FrontendState& st, FrontendHost* host, cudaStream_t stream);
// floats of the optional packed-spectrum buffer (B clips): `spec` of frontend_forward (written) / frontend_backward (read
// instead of recomputing the STFT); both calls must then see the same waveform
size_t frontend_spec_floats(int B, int T);

}  // namespace advb
