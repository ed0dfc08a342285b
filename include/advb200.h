/*
 * advb200 — C ABI of the B200-native adversarial-perturbation engine (libadvb200.so).
 *
 * The reference (piotrkawa/audio-deepfake-adversarial-attacks) is pure Python; its "FFI" for this path is the
 * Python call  atk(images, labels)  on a torchattacks object wrapping an nn.Module.  Each entry point below names
 * the reference interface it replaces (file:line under /root/reference).  Plain pointers and sizes only: all
 * tensor memory is owned by the caller (device pointers from the PyTorch allocator), the engine owns only its
 * workspace.  Every function returns 0 on success, non-zero on error (message via advb_last_error()); no C++
 * exception crosses this boundary.  A handle is bound to one CUDA device and is not thread-safe.
 */
#ifndef ADVB200_H
#define ADVB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ADVB_VERSION 202
#if defined(__GNUC__)
#define ADVB_API __attribute__((visibility("default")))
#else
#define ADVB_API
#endif

/* model kinds: src/models/models.py:6-18 (get_model names) */
enum { ADVB_MODEL_LCNN = 1, ADVB_MODEL_SPECRNET = 2, ADVB_MODEL_RAWNET3 = 3,
       ADVB_MODEL_FRONTEND_ONLY = 4 /* LFCC_FN(x) / MFCC_FN(x) on their own (src/frontends.py:13-50): only
                                       advb_frontend_fwd / advb_frontend_bwd are valid on such a handle */ };
/* frontend kinds: src/frontends.py:13-50 (0 = raw waveform, RawNet3) */
enum { ADVB_FRONTEND_NONE = 0, ADVB_FRONTEND_LFCC = 1, ADVB_FRONTEND_MFCC = 2 };
/* attack kinds: adversarial_attacks/torchattacks/attacks/{fgsm,pgd,pgdl2,fab,cw}.py */
enum { ADVB_ATTACK_FGSM = 1, ADVB_ATTACK_PGD = 2, ADVB_ATTACK_PGDL2 = 3, ADVB_ATTACK_FAB = 4, ADVB_ATTACK_CW = 5 };
enum { ADVB_NORM_LINF = 0, ADVB_NORM_L2 = 1 }; /* advb_attack_desc.norm (FAB) */
/* what advb_grad differentiates */
enum { ADVB_GRAD_CE = 0 /* mean 2-class cross-entropy, fgsm.py:43-57 */, ADVB_GRAD_LOGIT = 1 /* d o_i / d x_i, fab.py:90-105 */ };

typedef struct advb_handle advb_handle;

/* One named weight/buffer tensor borrowed from the caller (a state_dict entry; names as in the reference's
 * state_dict, e.g. "m_transform.6.weight", "frontend.filter_mat").  The engine reads the storage on every call
 * and never caches values across calls (adversarial training mutates weights, src/trainer.py:309-331). */
typedef struct {
  const char* name;
  const float* ptr; /* device pointer, contiguous fp32 */
  int64_t numel;
} advb_tensor_ref;

/* Replaces: models.get_model(model_name, config, device) + load_model (src/models/models.py:6-18,
 * src/utils.py:47-71) as seen from the attack: Attack.__init__(name, model) (attack.py:14-35). */
typedef struct {
  int model_kind;     /* ADVB_MODEL_* */
  int frontend_kind;  /* ADVB_FRONTEND_* */
  int device;         /* CUDA ordinal */
  int max_batch;      /* workspace is sized for this many clips */
  int n_samples;      /* T: samples per clip (64000 in BASELINE.json, 64600 native) */
  int n_tensors;
  const advb_tensor_ref* tensors;
} advb_model_desc;

/* Replaces the constructor arguments of torchattacks.FGSM/PGD/PGDL2/FAB/CW
 * (fgsm.py:28, pgd.py:31-32, pgdl2.py:31, fab.py:51-53, cw.py:38). */
typedef struct {
  int kind;           /* ADVB_ATTACK_* */
  float eps;          /* FGSM/PGD/PGDL2/FAB */
  float alpha;        /* PGD / PGDL2 step */
  int steps;
  float eps_div;      /* PGDL2 eps_for_division */
  float alpha_max;    /* FAB */
  float eta;          /* FAB */
  float beta;         /* FAB */
  float c;            /* CW */
  float kappa;        /* CW */
  float lr;           /* CW */
  int n_global_batch; /* N of the CE mean when the batch is sharded over ranks (0 = use B) */
  int targeted;       /* attack.py:60-108 targeted modes: cost = -loss(outputs, target_labels) (fgsm.py:49-50, pgd.py:64-65,
                         pgdl2.py:69-70); CW: f = clamp(i - j, -kappa) on the target one-hot (cw.py:82-83,131-132) */
  int norm;           /* FAB: ADVB_NORM_LINF (0, the AttackEnum presets) or ADVB_NORM_L2 (fab.py:55,184-194,216-219,236-240,
                         251-253,277-279).  L1 cannot run in the reference (FAB.perturb never defines `res` for it, fab.py:515-521)
                         and is not provided.  Occupies what used to be padding: the layout of every other field is unchanged */
  const int64_t* target_labels; /* [B] int64 device pointer, required when targeted != 0 (the host evaluates the
                         target_map_function, attack.py:258-270) */
} advb_attack_desc;

ADVB_API int advb_version(void);
ADVB_API const char* advb_last_error(void);

ADVB_API int advb_create(advb_handle** out, const advb_model_desc* desc);
ADVB_API void advb_destroy(advb_handle* h);
ADVB_API size_t advb_workspace_bytes(const advb_handle* h);
/* Re-point the borrowed tensors (same names/sizes) without reallocating the workspace. */
ADVB_API int advb_rebind(advb_handle* h, int n_tensors, const advb_tensor_ref* tensors);

/* The engine repacks the borrowed weights at the head of every call because they are live.  A host that tracks weight
 * versions itself sets option "weight_cache" = 1 and calls this after every change; unchanged weights then skip the repack. */
ADVB_API int advb_invalidate_weights(advb_handle* h);

/* Engine options (no reference counterpart; the reference's knobs are torch-global):
 *   "conv_path"   0 = tcgen05 tensor-core convolutions / GEMMs (default), 1 = fp32 SIMT convolutions / GEMMs (cross-check)
 *   "tf32_passes" 3 = 3xTF32 error-compensated products, fp32-class accuracy (default); 2 = LCNN forward 3x3 blocks with the tf32
 *                 main term + ONE bf16 MMA for both cross terms, 3xTF32 everywhere else (opt-in experiment: same logits and
 *                 gradients to the reference's tolerance, but the strict element-wise CW gate fails under it); 1 = single-pass tf32
 *   "conv_sched"  0 = persistent warp-specialised convolution / GEMM kernels (default), 1 = one-tile-per-CTA kernels only
 *                 (the first tcgen05 version; same arithmetic, kept as an in-process cross-check)
 *   "conv0_bwd"   LCNN first block backward: 0 = fp32 cell kernel (default), 1 = tcgen05 GEMM + col2im (cross-check)
 *   "conv0_fwd"   LCNN first block forward: 0 = Toeplitz GEMM without im2col (default), 1 = im2col GEMM (cross-check)
 *   "graph"       1 = one PGD / PGDL2 iteration is captured into a CUDA graph and replayed `steps` times (default), 0 = every
 *                 kernel enqueued by the host loop (same kernels, same results)
 *   "fuse_update" 1 = the FGSM / PGD L-inf update rule runs in the epilogue of the frontend backward and the waveform gradient
 *                 never reaches HBM (default; LFCC / MFCC models), 0 = separate update kernel (bit-identical iterates)
 *   "fe_spec"     1 = the frontend backward reads the packed spectra the forward stored (default), 0 = recomputes the STFT
 *   "lstm_tc"     1 = BLSTM input projections on the tcgen05 3xTF32 GEMM, 0 = fp32 SIMT GEMM (default)
 *   "sr_tc"       SpecRNet: 1 = the 64 -> 64 convolutions (conv2 of blocks 2 and 4) and their transposes on the persistent tcgen05
 *                 kernel, 3xTF32 (default), 0 = fp32 SIMT like the rest of SpecRNet
 *   "weight_cache" see advb_invalidate_weights */
ADVB_API int advb_set_option(advb_handle* h, const char* key, int value);

/* Strict multi-GPU mode (SURVEY.md §8(e) ii; no reference counterpart: the reference's nn.DataParallel computes one dB floor per
 * replica, evaluate_models_on_adversarial_attacks.py:163-167, which is also this engine's default).  amplitude_to_DB's top_db floor
 * is relative to the maximum of the whole batch, so a clip-sharded LFCC / MFCC run equals the single-device run only if the ranks
 * share that maximum (forward) and the summed gradient of the clamped elements (backward).  Under strict mode every frontend
 * forward / backward exchanges one float per rank through peer memory (8-byte stores into the peers' mailboxes over NVLink, polled
 * by a 32-thread kernel; no NCCL call, no host sync, part of the replayed CUDA graph).  Protocol, identical on every rank:
 *   1. advb_xrank_export   clears this rank's mailbox and returns its CUDA IPC handle (64 bytes) and device pointer
 *   2. the host exchanges the handles (torch.distributed.all_gather_object; this exchange is the barrier the protocol needs)
 *   3. advb_xrank_connect  maps the peers' mailboxes: ipc_handles = world x 64 bytes in rank order (other processes), and / or
 *      local_ptrs[r] = the device pointer of a handle living in THIS process (tests; peer access is enabled when the devices
 *      differ).  world <= 1 disconnects.
 * From then on every rank must make the same sequence of forward / gradient / FGSM / PGD / PGDL2 calls (any batch size >= 1 per
 * rank; pass n_global_batch); FAB and CW return an error.  A peer that does not arrive within 5 s makes the exchange give up
 * (results of that call are then per-shard floors) and sets the flag advb_xrank_status reads (call it after a stream sync).
 * Disconnect (advb_xrank_connect with world <= 1) on every rank before any rank destroys its handle: peers store into its mailbox. */
#define ADVB_XRANK_HANDLE_BYTES 64
ADVB_API int advb_xrank_export(advb_handle* h, unsigned char* ipc_handle, void** local_ptr);
ADVB_API int advb_xrank_connect(advb_handle* h, int rank, int world, const unsigned char* ipc_handles, void* const* local_ptrs);
ADVB_API int advb_xrank_status(advb_handle* h, int* timed_out);

/* Replaces  atk(images, labels)  = Attack.__call__ -> {FGSM,PGD,PGDL2,FAB,CW}.forward
 * (attack.py:308-331; fgsm.py:33-62; pgd.py:40-78; pgdl2.py:40-90; fab.py:70-78; cw.py:46-112), called from
 * evaluate_models_on_adversarial_attacks.py:220 and src/trainer.py:426,470,492,511,539.
 *   x        [B,T] fp32 in [0,1], device;  y [B] int64 labels (1 = bonafide), device
 *   start    nullable [B,T]: the random start drawn by the host with torch:
 *            PGD: the U(-eps,eps) noise added at pgd.py:56;  PGDL2: the already scaled delta of pgdl2.py:57-61;
 *            FAB: the restart point x1 of fab.py:176-206 (random restarts, n_restarts > 1)
 *   FAB      runs attack_single_run (fab.py:131-307; L-inf, untargeted, n_restarts = 1) on the given clips, which
 *            the caller has already restricted to the correctly classified ones as FAB.perturb does (fab.py:506-513)
 *   CW       cw.py:46-112, including the batch-wide early stop (one host sync every steps/10 iterations)
 *   x_adv    [B,T] fp32 out (must not alias x)
 * Work is enqueued on `cuda_stream` (a cudaStream_t); the call does not synchronise. */
ADVB_API int advb_attack(advb_handle* h, const advb_attack_desc* atk, const float* x, const int64_t* y, const float* start,
                float* x_adv, int B, int T, void* cuda_stream);

/* the three steps every call site wraps around the attack, in ONE call -
 *   x01, mn, mx = to_minmax(x_raw);  adv01 = atk(x01, y);  x_adv_raw = revert_minmax(adv01, mn, mx)
 * (evaluate_models_on_adversarial_attacks.py:219-221, src/trainer.py:425-427,469-471,491-493,510-512,538-540;
 * src/aa/utils.py:4-14).  Same arguments as advb_attack, raw (unscaled) waveforms in and out; x_adv_raw may alias x_raw.
 * A constant clip yields NaN, as in the reference (division by max - min = 0). */
ADVB_API int advb_attack_minmax(advb_handle* h, const advb_attack_desc* atk, const float* x_raw, const int64_t* y,
                       const float* start, float* x_adv_raw, int B, int T, void* cuda_stream);

/* Replaces  model(x)  (lcnn.py:239-243, specrnet.py:211-214, rawnet3.py:73-137): logits [B] (the (B,1) column).
 * Used for clean inference on the attacked batch (evaluate_models_on_adversarial_attacks.py:236-238) and
 * FAB's _get_predicted_label (fab.py:80-85). */
ADVB_API int advb_forward(advb_handle* h, const float* x, float* logits, int B, int T, void* cuda_stream);

/* Replaces  torch.autograd.grad(cost, adv_images)  (fgsm.py:56-57, pgd.py:71-72, pgdl2.py:76-77) and FAB's
 * get_diff_logits_grads_batch (fab.py:90-112).  grad [B,T], logits [B] (nullable). */
ADVB_API int advb_grad(advb_handle* h, int what, const float* x, const int64_t* y, float* grad, float* logits, int B, int T,
              int n_global_batch, void* cuda_stream);

/* Replaces LFCC_FN(x) / MFCC_FN(x) (src/frontends.py:13-32): coefficients [B,80,F] (torchaudio layout),
 * F = 1 + T/160, and their vector-Jacobian product.  Test / f4 entry points.  Error for a handle created with
 * ADVB_FRONTEND_NONE (RawNet3 consumes the raw waveform, rawnet3.py:73-137). */
ADVB_API int advb_frontend_fwd(advb_handle* h, const float* x, float* coeff, int B, int T, void* cuda_stream);
ADVB_API int advb_frontend_bwd(advb_handle* h, const float* x, const float* g_coeff, float* g_x, int B, int T,
                      void* cuda_stream);

/* Replaces prepare_mel_scale_vector(audio) (src/frontends.py:53-79, the `mel_spec` frontend of get_frontend, :48-49):
 * torch.stft(n_fft=512, hop=160, win_length=400, window=None) -> MEL_SCALE_FN on the real and on the imaginary part ->
 * stack([abs, angle], dim=1).  x [B,T]; fb [257,n_mels] = MEL_SCALE_FN.fb (live buffer, n_mels <= 128); out [B,2,n_mels,F],
 * F = 1 + T / 160.  Forward only: nothing in the reference differentiates through it (no model takes its 2 channels). */
ADVB_API int advb_mel_spec_fwd(const float* x, const float* fb, int n_mels, float* out, int B, int T, void* cuda_stream);

/* Replaces to_minmax / revert_minmax (src/aa/utils.py:4-14).  mn, mx: [B]. */
ADVB_API int advb_minmax(const float* x, float* x01, float* mn, float* mx, int B, int T, void* cuda_stream);
ADVB_API int advb_revert_minmax(const float* x01, const float* mn, const float* mx, float* x, int B, int T,
                       void* cuda_stream);

/* Replaces projection_linf(points_to_project, w_hyperplane, b_hyperplane) (fab.py:562-614): for each of R rows the
 * minimal-L-inf move d with <w, t + d> = b, t + d in [0,1]^T.  t, w, d: [R,T]; b: [R].  Test / f4 entry point. */
ADVB_API int advb_projection_linf(const float* t, const float* w, const float* b, float* d, int R, int T, void* cuda_stream);
/* Same contract for fab.py:617-665 (projection_l2): the minimum-L2 step onto the hyperplane inside the box. */
ADVB_API int advb_projection_l2(const float* t, const float* w, const float* b, float* d, int R, int T, void* cuda_stream);

/* Per-clip perturbation norms  ||a_i - b_i||_inf  and  ||a_i - b_i||_2  (either output nullable): the row reductions of
 * FAB.perturb (fab.py:515-521) and of the L-inf / L2 parity gates. */
ADVB_API int advb_row_diff_norms(const float* a, const float* b, float* linf, float* l2, int B, int T, void* cuda_stream);

/* Test introspection: copy one internal stage buffer of the last forward (raw engine layout) to `dst` (device).
 * dims[0..3] receive the logical (B, H, W, C) of the stage, dims[4] the spatial zero-border width of the
 * stored layout (rows/cols of padding on each side).  Returns the number of floats copied (<0 on error). */
ADVB_API int64_t advb_debug_stage(advb_handle* h, const char* stage, float* dst, int64_t capacity, int64_t dims[5],
                         void* cuda_stream);

/* Live per-kernel timing for bench.py's roofline: between begin and end an event is recorded after every launch
 * on the call's stream; end() synchronises and writes a JSON array [{"name","count","total_ms"}, ...] into buf
 * (NUL-terminated, truncated to capacity) and returns the size needed. */
ADVB_API int advb_profile_begin(advb_handle* h, void* cuda_stream);
ADVB_API int64_t advb_profile_end(advb_handle* h, char* buf, int64_t capacity);

/* Number of kernel launches issued by this handle since creation (bench.py's "gpu_launches"). */
ADVB_API int64_t advb_launch_count(const advb_handle* h);

#ifdef __cplusplus
}
#endif
#endif /* ADVB200_H */
