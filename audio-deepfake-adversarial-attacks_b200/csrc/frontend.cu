// LFCC / MFCC frontend, forward and input-gradient backward (sm_100a).
//
// Replaces torchaudio.transforms.LFCC / MFCC as instantiated by the reference's src/frontends.py:13-32
// (n_fft 512, Hann-400 centred, hop 160, reflect padding, power 2, 128 triangular filters, dB with the
// batch-wide top_db=80 floor, DCT-II ortho 128->80).  Math: SURVEY.md App. A.1.
//
// Data layout: waveform (B,T) fp32 row-major; dB energies (B,F,128); coefficients are written through
// caller-supplied strides so the same kernel emits torchaudio's (B,80,F), LCNN's zero-bordered (B,F+4,84,1)
// image or SpecRNet's (B,82,F+2,1) image.  The spectrum (824 KB/clip) is never stored: backward recomputes the
// STFT from the waveform.
//
// Kernels
//   fe_power_db   : one warp = two frames packed in one 512-point complex FFT held in shared memory;
//                   power -> sparse triangular filterbank -> 10 log10 -> dB store + batch arg-max (64-bit atomicMax)
//   fe_floor_dct  : floor at (batch max - 80), DCT by shared-memory matrix, strided store, clamped-element count
//   fe_dct_t      : d dB = d coefficients x dct^T for every frame (register-tiled fp32 SIMT product, (B,F,128) out); when
//                   the floor is active also the summed gradient of the clamped elements, which autograd routes to the batch
//                   arg-max element (per-block sums, added in block order by the last block to finish: deterministic)
//   fe_bwd        : per tile of 40 hops, one warp per frame pair: recompute FFT/power/energies, dB+floor backward,
//                   filterbank^T, one-sided inverse DFT as a packed complex FFT, window, overlap-add and reflect-pad
//                   fold in shared memory in a fixed order (deterministic, no atomics), gradient store
#include "frontend.cuh"
#include "tc_common.cuh"

#include <math.h>

#include <algorithm>

namespace advb {

namespace {

constexpr int NFFT = 512;
constexpr int WIN = 400;
constexpr int WOFF = 56;  // (512-400)/2
constexpr int HOP = 160;
constexpr int NBIN = 257;
constexpr int NFILT = 128;
constexpr int NCOEF = 80;
constexpr int FE_WARPS = 8;
constexpr int FE_THREADS = FE_WARPS * 32;
constexpr int FB_WARPS = 24;     // backward: a tile of 40 hops needs 43-45 frames = at most 23 frame pairs = one warp each
constexpr int FB_THREADS = FB_WARPS * 32;
constexpr int PSTRIDE = 260;     // padded 257
constexpr int TILE_HOPS = 40;    // backward tile = 40 hops = 6400 samples (10 % of the frames are recomputed by a neighbour)
constexpr int TILE_S = TILE_HOPS * HOP;
constexpr int NF_MAX = 2 * FB_WARPS;  // frames a backward tile may need (checked on the host: one frame pair per warp)
constexpr int DT_FR = 64;        // fe_dct_t: frames per work item
constexpr int DT_GLD = 84;       // fe_dct_t: padded row of the staged coefficient gradients (16-byte aligned rows)

__device__ __forceinline__ int reflect_index(int j, int T) {
  if (j < 0) j = -j;
  if (j >= T) j = 2 * T - 2 - j;
  return j;
}

// ---- 512-point complex FFT of one warp: radix-8 Stockham autosort, 3 out-of-place passes A -> B -> A -> B ----
// Each lane owns 2 x 8 points per pass (16 independent shared-memory loads in flight, one __syncwarp per pass); input
// and output are in natural order (no bit reversal).  A: 512 float2; B: 576 float2 - the intermediate after pass 1 is
// stored at i + (i >> 3) so that its stride-8 scatter is bank-conflict free.
constexpr int FFT_A = 512;   // float2 elements of buffer A
constexpr int FFT_B = 576;   // float2 elements of buffer B

__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
  return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ float2 mul_mi(float2 a) { return make_float2(a.y, -a.x); }  // a * (-i)

// forward DFT-8 (kernel e^{-2 pi i nk/8}) in registers, decimation in frequency
__device__ __forceinline__ void dft8(float2 (&v)[8]) {
  const float h = 0.70710678118654752f;
  const float2 a0 = cadd(v[0], v[4]), a4 = csub(v[0], v[4]);
  const float2 a1 = cadd(v[1], v[5]);
  float2 a5 = csub(v[1], v[5]);
  const float2 a2 = cadd(v[2], v[6]);
  float2 a6 = csub(v[2], v[6]);
  const float2 a3 = cadd(v[3], v[7]);
  float2 a7 = csub(v[3], v[7]);
  a5 = make_float2(h * (a5.x + a5.y), h * (a5.y - a5.x));    // * (1 - i)/sqrt2
  a6 = mul_mi(a6);                                           // * (-i)
  a7 = make_float2(h * (a7.y - a7.x), -h * (a7.x + a7.y));   // * (-1 - i)/sqrt2
  const float2 b0 = cadd(a0, a2), b2 = csub(a0, a2), b1 = cadd(a1, a3), b3 = mul_mi(csub(a1, a3));
  const float2 b4 = cadd(a4, a6), b6 = csub(a4, a6), b5 = cadd(a5, a7), b7 = mul_mi(csub(a5, a7));
  v[0] = cadd(b0, b1);
  v[4] = csub(b0, b1);
  v[2] = cadd(b2, b3);
  v[6] = csub(b2, b3);
  v[1] = cadd(b4, b5);
  v[5] = csub(b4, b5);
  v[3] = cadd(b6, b7);
  v[7] = csub(b6, b7);
}

// PASS 0: Ns = 1 (A -> B padded), PASS 1: Ns = 8 (B padded -> A), PASS 2: Ns = 64 (A -> B natural)
template <int PASS>
__device__ __forceinline__ void fft_pass(const float2* __restrict__ in, float2* __restrict__ out,
                                         const float2* __restrict__ s_tw, int lane) {
  constexpr int Ns = PASS == 0 ? 1 : (PASS == 1 ? 8 : 64);
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int j = lane + 32 * h;
    const int k = j & (Ns - 1);
    float2 v[8];
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      const int i = j + 64 * r;
      v[r] = in[PASS == 1 ? i + (i >> 3) : i];
    }
    if (PASS > 0) {
#pragma unroll
      for (int r = 1; r < 8; ++r) v[r] = cmul(v[r], s_tw[(r * k * (PASS == 1 ? 8 : 1)) & 511]);
    }
    dft8(v);
    const int j0 = ((j - k) << 3) + k;
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      const int i = j0 + r * Ns;
      out[PASS == 0 ? i + (i >> 3) : i] = v[r];
    }
  }
  __syncwarp();
}

// Z = FFT(bufA) -> bufB (natural order); bufA is clobbered.
__device__ __forceinline__ void warp_fft512(float2* bufA, float2* bufB, const float2* s_tw, int lane) {
  fft_pass<0>(bufA, bufB, s_tw, lane);
  fft_pass<1>(bufB, bufA, s_tw, lane);
  fft_pass<2>(bufA, bufB, s_tw, lane);
}

// Load two windowed frames (ta, ta+1) of clip `xb` packed as re = frame a, im = frame b, natural order.
__device__ __forceinline__ void load_frame_pair(const float* __restrict__ xb, int T, int F, int ta, const float* s_win,
                                                float2* buf, int lane) {
  const bool has_b = (ta + 1) < F;
  for (int n = lane; n < NFFT; n += 32) {
    float a = 0.f, b = 0.f;
    if (n >= WOFF && n < WOFF + WIN) {
      const float w = s_win != nullptr ? s_win[n - WOFF] : 1.0f;  // no table: torch.stft's default rectangular window
      const int ja = HOP * ta + n - NFFT / 2;
      a = w * __ldg(xb + reflect_index(ja, T));
      if (has_b) b = w * __ldg(xb + reflect_index(ja + HOP, T));
    }
    buf[n] = make_float2(a, b);
  }
  __syncwarp();
}

// Same frame pair out of `raw` = the 560 consecutive samples x[160 ta - 200 ..] (interior pairs: no reflection), already in
// shared memory (TMA prefetch of fe_power_db_kernel).
__device__ __forceinline__ void load_frame_pair_raw(const float* raw, const float* s_win, float2* buf, int lane) {
  for (int n = lane; n < NFFT; n += 32) {
    float a = 0.f, b = 0.f;
    if (n >= WOFF && n < WOFF + WIN) {
      const float w = s_win[n - WOFF];
      a = w * raw[n - WOFF];
      b = w * raw[n - WOFF + HOP];
    }
    buf[n] = make_float2(a, b);
  }
  __syncwarp();
}

// From the packed spectrum Z: one-sided spectra of both frames at bin k.
__device__ __forceinline__ void unpack_bin(const float2* z, int k, float& xar, float& xai, float& xbr, float& xbi) {
  const float2 p = z[k], q = z[(NFFT - k) & (NFFT - 1)];
  xar = 0.5f * (p.x + q.x);
  xai = 0.5f * (p.y - q.y);
  xbr = 0.5f * (p.y + q.y);
  xbi = 0.5f * (q.x - p.x);
}

// Energies of filters lane, lane+32, lane+64, lane+96 for both frames.  The four sparse dot products advance together: each
// is a chain of dependent L1 loads of run-time length, and run one after the other they were ~4 x the latency.  Per filter
// the bins are still added in ascending order (bit-identical to the sequential form).
__device__ __forceinline__ void filter_energy4(const float* __restrict__ fb, const int* __restrict__ klo,
                                               const int* __restrict__ kcnt, const float* pa, const float* pb, int lane,
                                               float (&ea)[4], float (&eb)[4]) {
  int k0[4], n[4], nmax = 0;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    k0[q] = klo[lane + 32 * q];
    n[q] = kcnt[lane + 32 * q];
    nmax = max(nmax, n[q]);
    ea[q] = 0.f;
    eb[q] = 0.f;
  }
  for (int i = 0; i < nmax; ++i) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      if (i < n[q]) {
        const float w = __ldg(fb + (size_t)(k0[q] + i) * NFILT + lane + 32 * q);
        ea[q] += pa[k0[q] + i] * w;
        eb[q] += pb[k0[q] + i] * w;
      }
    }
  }
}

__device__ __forceinline__ float to_db(float e) { return 10.0f * log10f(fmaxf(e, 1e-10f)); }

// Strict multi-GPU mode: all ranks agree on the batch-wide dB maximum (kind 0, between fe_power_db and fe_floor_dct) or on
// the summed gradient of the clamped elements (kind 1, between fe_dct_t and fe_bwd).  Lane r talks to rank r.
//   kind 0: the rank holding the largest key keeps its packed arg-max (ties: lowest rank = earliest clips, as the packed
//           ~index order decides inside one device); every other rank takes the key with an index no element matches, so the
//           clamped mass is applied exactly once, by the owner.
//   kind 1: partial sums are added in rank order on every rank (same bits everywhere).
constexpr unsigned long long XR_TIMEOUT_NS = 5ull * 1000 * 1000 * 1000;  // then sticky: later exchanges do not wait again
__device__ __forceinline__ unsigned long long xr_globaltimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
__global__ void __launch_bounds__(32) fe_xrank_kernel(FrontendState st, int kind) {
  const FrontendXRank& xr = st.xr;
  const int lane = threadIdx.x;
  const unsigned e = *xr.epoch + 1u;
  __syncwarp();
  const unsigned mine = kind == 0 ? (unsigned)(*st.gmax_packed >> 32) : __float_as_uint(*st.mass_total);
  const int par = (int)(e & 1u);
  unsigned got = 0u;
  bool ok = true;
  const bool dead = *(volatile int*)xr.timed_out != 0;
  if (lane < xr.world) {
    // a rank can run at most one exchange ahead of a peer that is still reading, hence the two parities
    unsigned long long* dst = xr.mailbox[lane] + par * XR_MAX_RANKS + xr.rank;
    const unsigned long long word = ((unsigned long long)mine << 32) | e;
    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(dst), "l"(word) : "memory");
    const unsigned long long* src = xr.mailbox[xr.rank] + par * XR_MAX_RANKS + lane;
    const unsigned long long t0 = xr_globaltimer();
    unsigned long long w;
    for (;;) {
      asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(w) : "l"(src) : "memory");
      if ((unsigned)w == e) break;
      if (dead || xr_globaltimer() - t0 > XR_TIMEOUT_NS) {
        ok = false;
        break;
      }
      __nanosleep(200);
    }
    got = (unsigned)(w >> 32);
  }
  if (__any_sync(0xffffffffu, !ok) && lane == 0) *xr.timed_out = 1;
  if (kind == 0) {
    // the packed key (float_to_key ^ 0x80000000) orders like the float under an unsigned compare; 0 = nothing seen
    unsigned best = lane < xr.world ? got : 0u;
    int owner = lane < xr.world ? lane : XR_MAX_RANKS;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const unsigned b2 = __shfl_xor_sync(0xffffffffu, best, o);
      const int o2 = __shfl_xor_sync(0xffffffffu, owner, o);
      if (b2 > best || (b2 == best && o2 < owner)) best = b2, owner = o2;
    }
    if (lane == 0 && owner != xr.rank) *st.gmax_packed = (unsigned long long)best << 32;  // ~index = 0: no element matches
  } else {
    float s = 0.f;
    for (int r = 0; r < xr.world; ++r) s += __uint_as_float(__shfl_sync(0xffffffffu, got, r));
    if (lane == 0) *st.mass_total = s;
  }
  if (lane == 0) *xr.epoch = e;
}

// ---------------------------------------------------------------------------------------------------
__global__ void fe_tables_kernel(const float* __restrict__ fb, int* klo, int* kcnt, int* mlo, int* mcnt) {
  const int i = threadIdx.x;
  if (i < NFILT) {  // per filter: first / last non-zero bin
    int lo = NBIN, hi = -1;
    for (int k = 0; k < NBIN; ++k)
      if (fb[(size_t)k * NFILT + i] != 0.f) {
        if (k < lo) lo = k;
        hi = k;
      }
    klo[i] = hi < 0 ? 0 : lo;
    kcnt[i] = hi < 0 ? 0 : hi - lo + 1;
  }
  if (i < NBIN) {  // per bin: first / last filter with a non-zero weight
    int lo = NFILT, hi = -1;
    for (int m = 0; m < NFILT; ++m)
      if (fb[(size_t)i * NFILT + m] != 0.f) {
        if (m < lo) lo = m;
        hi = m;
      }
    mlo[i] = hi < 0 ? 0 : lo;
    mcnt[i] = hi < 0 ? 0 : hi - lo + 1;
  }
}

__global__ void fe_dct_transpose_kernel(const float* __restrict__ dct, float* __restrict__ dctT) {
  const int i = threadIdx.x + blockIdx.x * blockDim.x;  // i = c * 128 + m
  if (i < NFILT * NCOEF) dctT[i] = dct[(i % NFILT) * NCOEF + i / NFILT];
}

__global__ void fe_twiddle_kernel(float2* tw) {
  const int k = threadIdx.x + blockIdx.x * blockDim.x;
  if (k < 512) {
    double s, c;
    sincospi(2.0 * (double)k / 512.0, &s, &c);
    tw[k] = make_float2((float)c, (float)(-s));
  }
}

__global__ void fe_reset_kernel(FrontendState st) {
  *st.gmax_packed = 0ull;
  *st.n_clamped = 0;
  *st.mass_total = 0.f;
  st.done[0] = 0u;
  st.done[1] = 0u;
}

// ---------------------------------------------------------------------------------------------------
// Forward 1: waveform -> dB filterbank energies (B,F,128) + batch arg-max.
__global__ void __launch_bounds__(FE_THREADS, 3) fe_power_db_kernel(const float* __restrict__ x, int T, int F,
                                                                      FrontendTables tb, FrontendState st,
                                                                      float* __restrict__ dB, float2* __restrict__ spec,
                                                                      int n_blocks, int n_clips) {
  extern __shared__ __align__(16) float smem[];
  float2* s_tw = reinterpret_cast<float2*>(smem);                    // 512 float2
  float* s_win = smem + 1024;                                        // 400
  float2* s_fft = reinterpret_cast<float2*>(s_win + 400);            // FE_WARPS * (FFT_A + FFT_B) float2
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int i = tid; i < 512; i += FE_THREADS) s_tw[i] = tb.tw[i];
  for (int i = tid; i < WIN; i += FE_THREADS) s_win[i] = tb.window[i];
  __syncthreads();
  float2* bufA = s_fft + warp * (FFT_A + FFT_B);
  float2* bufB = bufA + FFT_A;
  float* pa = reinterpret_cast<float*>(bufA);  // power vectors live in buffer A, which is scratch once Z sits in buffer B
  float* pb = pa + PSTRIDE;
  // Sample prefetch (round 2): an interior frame pair is 560 CONSECUTIVE, 16-byte aligned samples (160 ta - 200 is a multiple of
  // 4), and buffer B is free from the moment the power vectors are built until the next item's FFT writes it - so the warp's
  // elected lane posts ONE 2 240-byte TMA bulk copy of the NEXT item's samples into B there, and the filterbank / dB / arg-max
  // phase of this item hides its latency (the per-lane scalar __ldg's of the frame load were the kernel's top stall:
  // long_scoreboard 5.9 warps per issue).  The first / last pair of a clip (reflection) keep the direct loads.
  __shared__ uint64_t bar_pf[FE_WARPS];
  if (lane == 0) tc::mbar_init(&bar_pf[warp], 1);
  tc::fence_barrier_init();
  __syncwarp();
  bool pf_valid = false;
  unsigned pf_count = 0;
  const bool x_al16 = (reinterpret_cast<uintptr_t>(x) & 15) == 0;
  // interior pair whose first sample is 16-byte aligned in global memory (clip starts are not when T is not a multiple of 4)
  auto pair_interior = [&](int b_, int ta_) {
    const int s0 = HOP * ta_ - (NFFT / 2 - WOFF);
    return x_al16 && (ta_ + 1) < F && s0 >= 0 && s0 + HOP + WIN <= T && ((((long long)b_ * T + s0) & 3) == 0);
  };

  float best = -INFINITY;
  unsigned best_idx = 0xffffffffu;
  // persistent over (clip, block of 16 frames): tables staged once per CTA, one atomicMax per warp per launch
  for (int work = blockIdx.x; work < n_blocks * n_clips; work += gridDim.x) {
    const int b = work / n_blocks;
    const int ta = (work - b * n_blocks) * (2 * FE_WARPS) + 2 * warp;
    if (ta >= F) {  // warp-uniform; the loop has no block-wide barrier
      pf_valid = false;
      continue;
    }
    const bool has_b = (ta + 1) < F;
    const float* xb = x + (size_t)b * T;

    if (pf_valid) {
      tc::mbar_wait(&bar_pf[warp], (pf_count - 1u) & 1u);
      load_frame_pair_raw(reinterpret_cast<const float*>(bufB), s_win, bufA, lane);
    } else {
      load_frame_pair(xb, T, F, ta, s_win, bufA, lane);
    }
    warp_fft512(bufA, bufB, s_tw, lane);
    if (spec != nullptr) {
      // keep the packed spectrum of the frame pair (ta, ta + 1) for the backward: 4 KB per pair, coalesced float2 rows.  The
      // backward used to recompute load + FFT per pair; every frontend kernel is latency-bound at 3 % of the DRAM bandwidth,
      // so 105 MB of extra traffic per pass (B = 128) is cheaper than the second FFT.
      float2* dst = spec + ((size_t)b * ((F + 1) >> 1) + (ta >> 1)) * NFFT;
#pragma unroll
      for (int i = 0; i < NFFT / 32; ++i) dst[lane + 32 * i] = bufB[lane + 32 * i];
    }
    for (int k = lane; k < NBIN; k += 32) {
      float xar, xai, xbr, xbi;
      unpack_bin(bufB, k, xar, xai, xbr, xbi);
      pa[k] = xar * xar + xai * xai;
      pb[k] = xbr * xbr + xbi * xbi;
    }
    __syncwarp();
    {  // buffer B is free: prefetch the next item's samples into it
      const int nwork = work + gridDim.x;
      pf_valid = false;
      if (nwork < n_blocks * n_clips) {
        const int nb = nwork / n_blocks;
        const int nta = (nwork - nb * n_blocks) * (2 * FE_WARPS) + 2 * warp;
        if (nta < F && pair_interior(nb, nta)) {
          pf_valid = true;
          ++pf_count;
          if (lane == 0) {
            tc::fence_proxy_async();
            tc::mbar_expect_tx(&bar_pf[warp], (uint32_t)((HOP + WIN) * sizeof(float)));
            tc::bulk_g2s(bufB, x + (size_t)nb * T + HOP * nta - (NFFT / 2 - WOFF), (uint32_t)((HOP + WIN) * sizeof(float)), &bar_pf[warp]);
          }
        }
      }
    }

    const size_t rowa = ((size_t)b * F + ta) * NFILT;
    float ea[4], eb[4];
    filter_energy4(tb.fb, tb.klo, tb.kcnt, pa, pb, lane, ea, eb);
    __syncwarp();  // the next item's frame load overwrites the power vectors (buffer A)
#pragma unroll
    for (int i = 0; i < NFILT / 32; ++i) {
      const int m = lane + 32 * i;
      const float da = to_db(ea[i]);
      dB[rowa + m] = da;
      if (da > best || (da == best && (unsigned)(rowa + m) < best_idx)) {
        best = da;
        best_idx = (unsigned)(rowa + m);
      }
      if (has_b) {
        const float db = to_db(eb[i]);
        dB[rowa + NFILT + m] = db;
        if (db > best || (db == best && (unsigned)(rowa + NFILT + m) < best_idx)) {
          best = db;
          best_idx = (unsigned)(rowa + NFILT + m);
        }
      }
    }
  }
  // warp arg-max (largest value, smallest index on ties), then one 64-bit atomicMax per warp
  unsigned long long packed =
      ((unsigned long long)((unsigned)float_to_key(best) ^ 0x80000000u) << 32) | (unsigned long long)(~best_idx);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const unsigned long long other = __shfl_xor_sync(0xffffffffu, packed, o);
    packed = other > packed ? other : packed;
  }
  if (lane == 0 && best_idx != 0xffffffffu) atomicMax(st.gmax_packed, packed);
}

// ---------------------------------------------------------------------------------------------------
// `mel_spec` frontend (src/frontends.py:53-79, prepare_mel_scale_vector): torch.stft(n_fft 512, hop 160, win_length 400, no
// window = ones(400) centred in the 512-sample frame, reflect-padded, one-sided), MelScale applied to the REAL and to the
// IMAGINARY part separately (a linear map: fb^T re, fb^T im), then |.| and angle of the resulting complex number.
// out (B, 2, n_mels, F): channel 0 = abs, channel 1 = angle.  Same two-frames-per-warp FFT as the cepstral frontends; the
// filter sums read the packed spectrum directly (each bin feeds at most two triangular filters).
constexpr int MS_MAX_MELS = 128;
__global__ void __launch_bounds__(FE_THREADS, 3) fe_melspec_kernel(const float* __restrict__ x, int T, int F,
                                                                     const float* __restrict__ fb, int n_mels,
                                                                     float* __restrict__ out, int n_blocks, int n_clips) {
  extern __shared__ __align__(16) float smem[];
  float2* s_tw = reinterpret_cast<float2*>(smem);                    // 512 float2
  int* s_klo = reinterpret_cast<int*>(smem + 1024);                  // MS_MAX_MELS
  int* s_kcnt = s_klo + MS_MAX_MELS;
  float2* s_fft = reinterpret_cast<float2*>(s_kcnt + MS_MAX_MELS);   // FE_WARPS * (FFT_A + FFT_B) float2
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int k = tid; k < 512; k += FE_THREADS) {
    float sn, cs;
    sincospif(2.0f * (float)k / 512.0f, &sn, &cs);
    s_tw[k] = make_float2(cs, -sn);
  }
  for (int m = tid; m < n_mels; m += FE_THREADS) {  // extent of each (triangular) filter in the live table
    int lo = NBIN, hi = -1;
    for (int k = 0; k < NBIN; ++k)
      if (fb[(size_t)k * n_mels + m] != 0.f) {
        if (k < lo) lo = k;
        hi = k;
      }
    s_klo[m] = hi < 0 ? 0 : lo;
    s_kcnt[m] = hi < 0 ? 0 : hi - lo + 1;
  }
  __syncthreads();
  float2* bufA = s_fft + warp * (FFT_A + FFT_B);
  float2* bufB = bufA + FFT_A;
  for (int work = blockIdx.x; work < n_blocks * n_clips; work += gridDim.x) {
    const int b = work / n_blocks;
    const int ta = (work - b * n_blocks) * (2 * FE_WARPS) + 2 * warp;
    if (ta >= F) continue;  // warp-uniform
    const bool has_b = (ta + 1) < F;
    load_frame_pair(x + (size_t)b * T, T, F, ta, nullptr, bufA, lane);
    warp_fft512(bufA, bufB, s_tw, lane);
    for (int m = lane; m < n_mels; m += 32) {
      float ar = 0.f, ai = 0.f, br = 0.f, bi = 0.f;
      const int k0 = s_klo[m], n = s_kcnt[m];
      for (int i = 0; i < n; ++i) {
        const float w = __ldg(fb + (size_t)(k0 + i) * n_mels + m);
        float xar, xai, xbr, xbi;
        unpack_bin(bufB, k0 + i, xar, xai, xbr, xbi);
        ar = fmaf(xar, w, ar), ai = fmaf(xai, w, ai);
        br = fmaf(xbr, w, br), bi = fmaf(xbi, w, bi);
      }
      float* o_abs = out + (((size_t)b * 2 + 0) * n_mels + m) * F + ta;
      float* o_ang = out + (((size_t)b * 2 + 1) * n_mels + m) * F + ta;
      o_abs[0] = hypotf(ar, ai);
      o_ang[0] = atan2f(ai, ar);
      if (has_b) {
        o_abs[1] = hypotf(br, bi);
        o_ang[1] = atan2f(bi, br);
      }
    }
    __syncwarp();  // the next item's frame load overwrites buffer A / the spectrum in buffer B is re-used
  }
}

__device__ __forceinline__ void decode_gmax(unsigned long long packed, float& vmax, unsigned& idx) {
  vmax = key_to_float((int)((unsigned)(packed >> 32) ^ 0x80000000u));
  idx = ~(unsigned)(packed & 0xffffffffull);
}

// Forward 2: floor + DCT, (B F, 128) x (128, 80) fp32 SIMT.  60 frames per work item; thread = (coefficient quad, group of
// 5 frames): per 4 filters 4 LDS.128 of dct rows + 5 LDS.128 of dB values feed 80 FMAs (the first version, thread =
// (frame, 5 coefficients), issued 6 LDS per 5 FMAs).  Persistent: the dct matrix (40 KB) is staged once per CTA.
constexpr int FD_FR = 60;    // frames per work item = 12 groups of 5
constexpr int FD_LD = 132;   // padded dB row in shared memory (16-byte aligned; the two frame groups of a warp hit different banks)
__global__ void __launch_bounds__(256, 3) fe_floor_dct_kernel(const float* __restrict__ dB, int F, FrontendTables tb,
                                                            FrontendState st, float top_db, float* __restrict__ out,
                                                            long long clip_stride, long long stride_f,
                                                            long long stride_c, long long offset, int n_rows) {
  extern __shared__ __align__(16) float smem[];
  float* s_dct = smem;                     // 128 x 80
  float* s_d = s_dct + NFILT * NCOEF;      // FD_FR x FD_LD
  const int tid = threadIdx.x;
  float vmax;
  unsigned amax_idx;
  decode_gmax(*st.gmax_packed, vmax, amax_idx);
  const float floor_v = vmax - top_db;
  for (int i = tid; i < NFILT * NCOEF / 4; i += 256)  // staged once per (persistent) CTA
    reinterpret_cast<float4*>(s_dct)[i] = __ldg(reinterpret_cast<const float4*>(tb.dct) + i);
  const int cq = tid % 20, fg = tid / 20;  // fg 0..11 active, 12 idle (threads 240..255)
  int clamped_any = 0;
  // rows = (clip, frame) pairs, flat: dB is (B F, 128) contiguous
  for (int r0 = blockIdx.x * FD_FR; r0 < n_rows; r0 += gridDim.x * FD_FR) {
    __syncthreads();  // previous iteration's reads of s_d are done (and s_dct is complete on the first pass)
    int clamped = 0;
    for (int i = tid; i < FD_FR * (NFILT / 4); i += 256) {
      const int fl = i >> 5, m4 = i & 31;
      float4 d = make_float4(0.f, 0.f, 0.f, 0.f);
      if (r0 + fl < n_rows) {
        d = __ldg(reinterpret_cast<const float4*>(dB + (size_t)(r0 + fl) * NFILT) + m4);
        if (d.x < floor_v) d.x = floor_v, ++clamped;
        if (d.y < floor_v) d.y = floor_v, ++clamped;
        if (d.z < floor_v) d.z = floor_v, ++clamped;
        if (d.w < floor_v) d.w = floor_v, ++clamped;
      }
      *reinterpret_cast<float4*>(s_d + fl * FD_LD + 4 * m4) = d;
    }
    clamped_any += __syncthreads_count(clamped);  // threads that clamped something (enough for an "active" flag)

    if (fg < FD_FR / 5) {
      float acc[5][4];
#pragma unroll
      for (int j = 0; j < 5; ++j) acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.f;
      const float* drow = s_d + (fg * 5) * FD_LD;
#pragma unroll 2
      for (int m = 0; m < NFILT; m += 4) {
        float4 w[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) w[i] = *reinterpret_cast<const float4*>(s_dct + (m + i) * NCOEF + 4 * cq);
#pragma unroll
        for (int j = 0; j < 5; ++j) {
          const float4 d = *reinterpret_cast<const float4*>(drow + j * FD_LD + m);
          acc[j][0] = fmaf(d.w, w[3].x, fmaf(d.z, w[2].x, fmaf(d.y, w[1].x, fmaf(d.x, w[0].x, acc[j][0]))));
          acc[j][1] = fmaf(d.w, w[3].y, fmaf(d.z, w[2].y, fmaf(d.y, w[1].y, fmaf(d.x, w[0].y, acc[j][1]))));
          acc[j][2] = fmaf(d.w, w[3].z, fmaf(d.z, w[2].z, fmaf(d.y, w[1].z, fmaf(d.x, w[0].z, acc[j][2]))));
          acc[j][3] = fmaf(d.w, w[3].w, fmaf(d.z, w[2].w, fmaf(d.y, w[1].w, fmaf(d.x, w[0].w, acc[j][3]))));
        }
      }
#pragma unroll
      for (int j = 0; j < 5; ++j) {
        const int r = r0 + fg * 5 + j;
        if (r < n_rows) {
          const int b = r / F, f = r - b * F;
          float* o = out + (size_t)b * clip_stride + offset + (long long)f * stride_f + (long long)(4 * cq) * stride_c;
#pragma unroll
          for (int i = 0; i < 4; ++i) o[(long long)i * stride_c] = acc[j][i];
        }
      }
    }
  }
  if (tid == 0 && clamped_any > 0) atomicAdd(st.n_clamped, clamped_any);
}

// ---------------------------------------------------------------------------------------------------
// Backward pre-pass: d dB (before the floor) = d coefficients x dct^T, one (B F, 80) x (80, 128) fp32 SIMT product.
// Thread = 8 frames x 4 filters (32 accumulators): per 4 coefficients 4 conflict-free LDS.128 of dct^T and 8 broadcast
// LDS.128 of the gradients feed 128 FMAs.  Persistent: dct^T (40 KB) is staged once per CTA.
// It also produces what the two mass kernels used to: the summed gradient of the dB elements the top_db floor clamped (the
// product row IS their gradient), per block in a fixed order, reduced by the last block to finish in block order.
__global__ void __launch_bounds__(256, 3) fe_dct_t_kernel(const float* __restrict__ gcoef, long long g_clip_stride,
                                                        long long g_stride_f, long long g_stride_c, int F,
                                                        FrontendTables tb, float* __restrict__ gd, int n_rows,
                                                        const float* __restrict__ dB, FrontendState st, float top_db,
                                                        float* __restrict__ partial) {
  extern __shared__ __align__(16) float smem[];
  float* s_w = smem;                  // 80 x 128: s_w[c][m] = dct[m][c]
  float* s_g = s_w + NCOEF * NFILT;   // DT_FR x DT_GLD
  __shared__ float s_mass[8];
  __shared__ unsigned s_last;
  const int tid = threadIdx.x, mq = tid & 31, fg = tid >> 5;
  const bool floor_active = *st.n_clamped != 0;
  float floor_v = 0.f, mass = 0.f;
  if (floor_active) {
    float vmax;
    unsigned amax_idx;
    decode_gmax(*st.gmax_packed, vmax, amax_idx);
    floor_v = vmax - top_db;
  }
  for (int i = tid; i < NFILT * NCOEF / 4; i += 256)
    reinterpret_cast<float4*>(s_w)[i] = __ldg(reinterpret_cast<const float4*>(tb.dctT) + i);
  // rows = (clip, frame) pairs, flat
  for (int r0 = blockIdx.x * DT_FR; r0 < n_rows; r0 += gridDim.x * DT_FR) {
    __syncthreads();  // the previous item's reads of s_g are done
    for (int i = tid; i < DT_FR * NCOEF; i += 256) {
      int fl, c;
      if (g_stride_c == 1) fl = i / NCOEF, c = i - fl * NCOEF;   // coefficient-contiguous gradients (the engine's layout)
      else c = i / DT_FR, fl = i - c * DT_FR;                   // frame-contiguous (torchaudio's (B,80,F))
      const int r = r0 + fl;
      float v = 0.f;
      if (r < n_rows) {
        const int b = r / F, f = r - b * F;
        v = gcoef[(size_t)b * g_clip_stride + (long long)f * g_stride_f + (long long)c * g_stride_c];
      }
      s_g[fl * DT_GLD + c] = v;
    }
    __syncthreads();
    float acc[8][4];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.f;
#pragma unroll 2
    for (int c = 0; c < NCOEF; c += 4) {
      float4 w[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) w[i] = *reinterpret_cast<const float4*>(s_w + (c + i) * NFILT + 4 * mq);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float4 g = *reinterpret_cast<const float4*>(s_g + (fg * 8 + j) * DT_GLD + c);
        acc[j][0] = fmaf(g.w, w[3].x, fmaf(g.z, w[2].x, fmaf(g.y, w[1].x, fmaf(g.x, w[0].x, acc[j][0]))));
        acc[j][1] = fmaf(g.w, w[3].y, fmaf(g.z, w[2].y, fmaf(g.y, w[1].y, fmaf(g.x, w[0].y, acc[j][1]))));
        acc[j][2] = fmaf(g.w, w[3].z, fmaf(g.z, w[2].z, fmaf(g.y, w[1].z, fmaf(g.x, w[0].z, acc[j][2]))));
        acc[j][3] = fmaf(g.w, w[3].w, fmaf(g.z, w[2].w, fmaf(g.y, w[1].w, fmaf(g.x, w[0].w, acc[j][3]))));
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int r = r0 + fg * 8 + j;
      if (r < n_rows) {
        *reinterpret_cast<float4*>(gd + (size_t)r * NFILT + 4 * mq) =
            make_float4(acc[j][0], acc[j][1], acc[j][2], acc[j][3]);
        if (floor_active) {
          const float4 d = __ldg(reinterpret_cast<const float4*>(dB + (size_t)r * NFILT) + mq);
          mass += (d.x < floor_v ? acc[j][0] : 0.f) + (d.y < floor_v ? acc[j][1] : 0.f) +
                  (d.z < floor_v ? acc[j][2] : 0.f) + (d.w < floor_v ? acc[j][3] : 0.f);
        }
      }
    }
  }
  // block sum (fixed order), then the last block to arrive adds the per-block sums in block order: deterministic
  mass = warp_sum(mass);
  if (mq == 0) s_mass[fg] = mass;
  __syncthreads();
  if (tid == 0) {
    float s = 0.f;
    for (int w = 0; w < 8; ++w) s += s_mass[w];
    partial[blockIdx.x] = s;
    __threadfence();
    s_last = atomicAdd(&st.done[0], 1u) == gridDim.x - 1 ? 1u : 0u;
  }
  __syncthreads();
  if (s_last != 0u && tid < 32) {
    __threadfence();
    float s = 0.f;
    for (int i = tid; i < (int)gridDim.x; i += 32) s += __ldcg(partial + i);
    s = warp_sum(s);
    if (tid == 0) {
      *st.mass_total = floor_active ? s : 0.f;
      st.done[0] = 0u;
    }
  }
}

// Frames whose (reflected) support can touch the samples [s0, s1) of a clip of T samples / F frames.
__host__ __device__ __forceinline__ void tile_frames(int s0, int s1, int T, int F, int& t_lo, int& t_hi) {
  t_lo = (s0 - 199 + 159 + 160 * 4) / 160 - 4;
  if (t_lo < 0) t_lo = 0;
  t_hi = (s1 - 1 + 200) / 160;
  if (s1 >= T - 202) t_hi = F - 1;
  if (t_hi > F - 1) t_hi = F - 1;
}

// Backward: d dB -> d waveform for one tile of TILE_S samples of one clip.  One warp per frame pair, 24 warps per SM.
// Shared memory per warp is the two FFT buffers only (8.5 KB): after the forward FFT buffer A is scratch and holds the
// power / d power / d energy vectors, and the windowless frame gradients stay in buffer B until the overlap-add reads
// them there.  (With a private dct^T, power and frame-gradient buffers the kernel fitted 10 warps per SM and sat on
// shared-memory latency: 15 % of the warp slots, 29 % of the issue slots, profiles/r01_ncu_full_final.md.)
__global__ void __launch_bounds__(FB_THREADS, 1) fe_bwd_kernel(const float* __restrict__ x, int T, int F,
                                                                FrontendTables tb, FrontendState st, float top_db,
                                                                const float* __restrict__ gd, float* __restrict__ gx,
                                                                const float2* __restrict__ spec, int n_tiles, int n_clips,
                                                                FusedUpdate upd) {
  extern __shared__ __align__(16) float smem[];
  float2* s_tw = reinterpret_cast<float2*>(smem);                  // 512 float2
  float* s_win = smem + 1024;                                      // 400
  float2* s_fft = reinterpret_cast<float2*>(s_win + 400);          // FB_WARPS * (FFT_A + FFT_B) float2
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int i = tid; i < 512; i += FB_THREADS) s_tw[i] = tb.tw[i];
  for (int i = tid; i < WIN; i += FB_THREADS) s_win[i] = tb.window[i];
  __syncthreads();

  float vmax;
  unsigned amax_idx;
  decode_gmax(*st.gmax_packed, vmax, amax_idx);
  const float floor_v = vmax - top_db;
  const float mass_total = *st.mass_total;

  float2* bufA = s_fft + warp * (FFT_A + FFT_B);
  float2* bufB = bufA + FFT_A;
  float* pa = reinterpret_cast<float*>(bufA);  // power, then d power        (buffer A is free once Z sits in buffer B)
  float* pb = pa + PSTRIDE;
  float* gea = pb + PSTRIDE;                   // d energy
  float* geb = gea + NFILT;                    // 2 * 260 + 2 * 128 = 776 floats <= 1024

  for (int work = blockIdx.x; work < n_tiles * n_clips; work += gridDim.x) {
    const int b = work / n_tiles, tile = work - b * n_tiles;
    const int s0 = tile * TILE_S;
    const int s1 = min(T, s0 + TILE_S);
    int t_lo, t_hi;
    tile_frames(s0, s1, T, F, t_lo, t_hi);
    if (spec != nullptr) t_lo &= ~1;  // pairs aligned with the forward's (even, even + 1) pairs, whose packed spectra are stored
    const int nf = t_hi - t_lo + 1;  // host guarantees nf <= NF_MAX: warp w owns frames t_lo + 2 w, t_lo + 2 w + 1
    const float* xb = x + (size_t)b * T;

    if (2 * warp < nf) {
      const int ta = t_lo + 2 * warp;
      const bool has_b = (2 * warp + 1 < nf);  // frame ta+1 is inside [t_lo, t_hi] (hence < F)
      // 1. d dB of both frames (4 filters per lane), in flight during the FFT
      const size_t rowa = ((size_t)b * F + ta) * NFILT;
      float gda[4], gdb[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        gda[i] = __ldg(gd + rowa + lane + 32 * i);
        gdb[i] = has_b ? __ldg(gd + rowa + NFILT + lane + 32 * i) : 0.f;
      }
      // 2. packed FFT of both frames, Z in bufB: read back from the forward's store, or recomputed
      if (spec != nullptr) {
        const float2* src = spec + ((size_t)b * ((F + 1) >> 1) + (ta >> 1)) * NFFT;
#pragma unroll
        for (int i = 0; i < NFFT / 32; ++i) bufB[lane + 32 * i] = __ldg(src + lane + 32 * i);
        __syncwarp();
      } else {
        load_frame_pair(xb, T, has_b ? F : ta + 1, ta, s_win, bufA, lane);
        warp_fft512(bufA, bufB, s_tw, lane);
      }
      // 3. power
      for (int k = lane; k < NBIN; k += 32) {
        float xar, xai, xbr, xbi;
        unpack_bin(bufB, k, xar, xai, xbr, xbi);
        pa[k] = xar * xar + xai * xai;
        pb[k] = xbr * xbr + xbi * xbi;
      }
      __syncwarp();
      // 4. energies, dB / floor backward, d energy
      {
        const float k10 = 4.342944819032518f;  // 10 / ln 10
        float ea[4], eb[4];
        filter_energy4(tb.fb, tb.klo, tb.kcnt, pa, pb, lane, ea, eb);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int m = lane + 32 * i;
          if ((unsigned)(rowa + m) == amax_idx) gda[i] += mass_total;
          if ((unsigned)(rowa + NFILT + m) == amax_idx) gdb[i] += mass_total;
          const float da = to_db(ea[i]), db = to_db(eb[i]);
          gea[m] = (da > floor_v && ea[i] >= 1e-10f) ? gda[i] * k10 / ea[i] : 0.f;
          geb[m] = (has_b && db > floor_v && eb[i] >= 1e-10f) ? gdb[i] * k10 / eb[i] : 0.f;
        }
      }
      __syncwarp();
      // 5. d power (overwrites the power vectors), kept in registers: step 6 overwrites buffer A
      float ga[9], gb[9];
      {
        // the 9 bins of a lane advance together (same reason as filter_energy4); per bin ascending filter order
        int mlo9[9], n9[9], nmax = 0;
#pragma unroll
        for (int i = 0; i < 9; ++i) {
          const int k = lane + 32 * i;
          mlo9[i] = k < NBIN ? tb.mlo[k] : 0;
          n9[i] = k < NBIN ? tb.mcnt[k] : 0;
          nmax = max(nmax, n9[i]);
          ga[i] = 0.f;
          gb[i] = 0.f;
        }
        for (int q = 0; q < nmax; ++q) {
#pragma unroll
          for (int i = 0; i < 9; ++i) {
            if (q < n9[i]) {
              const float w = __ldg(tb.fb + (size_t)(lane + 32 * i) * NFILT + mlo9[i] + q);
              ga[i] += w * gea[mlo9[i] + q];
              gb[i] += w * geb[mlo9[i] + q];
            }
          }
        }
      }
      __syncwarp();
      // 6. conj(H) into bufA, H = Ha + i Hb Hermitian-extended one-sided gradients (no doubling of interior bins)
#pragma unroll
      for (int i = 0; i < 9; ++i) {
        const int k = lane + 32 * i;
        if (k < NBIN) {
          float xar, xai, xbr, xbi;
          unpack_bin(bufB, k, xar, xai, xbr, xbi);
          if (k == 0 || k == 256) {
            bufA[k] = make_float2(2.f * ga[i] * xar, -(2.f * gb[i] * xbr));
          } else {
            const float har = ga[i] * xar, hai = ga[i] * xai, hbr = gb[i] * xbr, hbi = gb[i] * xbi;
            bufA[k] = make_float2(har - hbi, -(hai + hbr));
            bufA[NFFT - k] = make_float2(har + hbi, -(hbr - hai));
          }
        }
      }
      __syncwarp();
      // 7. W = conj(FFT(conj H)): ya = Re, yb = -Im; stays in bufB (natural order) for the overlap-add below
      warp_fft512(bufA, bufB, s_tw, lane);
    }
    __syncthreads();

    // window + overlap-add + reflect-pad fold, fixed order
    for (int sl = tid; sl < s1 - s0; sl += FB_THREADS) {
      const int s = s0 + sl;
      float acc = 0.f;
#pragma unroll 1
      for (int v = 0; v < 3; ++v) {
        int j;
        if (v == 0) j = s;
        else if (v == 1) {
          if (s < 1 || s > 256) continue;
          j = -s;
        } else {
          if (s > T - 2 || s < T - 2 - 255) continue;
          j = 2 * T - 2 - s;
        }
        int tf = (j - 199 + 159 + 160 * 4) / 160 - 4;
        int tl = (j + 200 + 160 * 4) / 160 - 4;
        if (tf < t_lo) tf = t_lo;
        if (tl > t_hi) tl = t_hi;
        for (int t = tf; t <= tl; ++t) {
          const int fi = t - t_lo, n = j - HOP * t + 200;
          const float2 wv = s_fft[(fi >> 1) * (FFT_A + FFT_B) + FFT_A + WOFF + n];
          const float w = s_win[n];
          acc += (fi & 1) ? -__fmul_rn(w, wv.y) : __fmul_rn(w, wv.x);
        }
      }
      const size_t o = (size_t)b * T + s;
      if (upd.kind == 0) {
        gx[o] = acc;
      } else {
        // same op-by-op fp32 rounding as update.cu (fgsm_step_kernel / pgd_step_kernel): bit-identical iterates
        const float sg = acc > 0.f ? 1.f : (acc < 0.f ? -1.f : 0.f);
        const float xi = __ldg(upd.x_clean + o);
        if (upd.kind == 1) {
          upd.adv_out[o] = fminf(fmaxf(__fadd_rn(xi, __fmul_rn(upd.eps, sg)), 0.f), 1.f);
        } else {
          const float a1 = __fadd_rn(__ldg(xb + s), __fmul_rn(upd.alpha, sg));
          const float d = fminf(fmaxf(__fsub_rn(a1, xi), -upd.eps), upd.eps);
          upd.adv_out[o] = fminf(fmaxf(__fadd_rn(xi, d), 0.f), 1.f);
        }
      }
    }
    __syncthreads();  // the FFT buffers are reused by the next tile
  }
  // Every block read the batch arg-max / floor state at its start; the last one to finish resets it for the next forward
  // (the separate reset kernel of every iteration is gone; a forward that follows a forward still launches it).
  if (tid == 0) {
    __threadfence();
    if (atomicAdd(&st.done[1], 1u) == gridDim.x - 1) {
      *st.gmax_packed = 0ull;
      *st.n_clamped = 0;
      *st.mass_total = 0.f;
      st.done[1] = 0u;
    }
  }
}

size_t fe_fwd_smem() { return (size_t)(1024 + 400 + FE_WARPS * 2 * (FFT_A + FFT_B)) * sizeof(float); }  // 74 KB: 3 CTAs / SM
size_t fe_dct_smem() { return (size_t)(NFILT * NCOEF + FD_FR * FD_LD) * sizeof(float); }
size_t fe_dct_t_smem() { return (size_t)(NCOEF * NFILT + DT_FR * DT_GLD) * sizeof(float); }
size_t fe_bwd_smem() { return (size_t)(1024 + 400 + FB_WARPS * 2 * (FFT_A + FFT_B)) * sizeof(float); }

}  // namespace

int frontend_frames(int T) { return 1 + T / HOP; }
size_t frontend_spec_floats(int B, int T) { return (size_t)B * ((frontend_frames(T) + 1) / 2) * NFFT * 2; }
int frontend_mass_blocks(int B, int T) { return std::max(B * cdiv(frontend_frames(T), 2 * FE_WARPS), 148 * 3 + 8); }

int frontend_init_constants(float2* tw, cudaStream_t stream) {
  fe_twiddle_kernel<<<2, 256, 0, stream>>>(tw);
  ADVB_KERNEL_OK("fe_twiddle", stream);
  ADVB_CUDA_OK(cudaFuncSetAttribute(fe_power_db_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fe_fwd_smem()));
  ADVB_CUDA_OK(cudaFuncSetAttribute(fe_floor_dct_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fe_dct_smem()));
  ADVB_CUDA_OK(cudaFuncSetAttribute(fe_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fe_bwd_smem()));
  ADVB_CUDA_OK(cudaFuncSetAttribute(fe_dct_t_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fe_dct_t_smem()));
  return 0;
}

int frontend_prepare(const FrontendTables& tb, cudaStream_t stream) {
  fe_tables_kernel<<<1, 288, 0, stream>>>(tb.fb, tb.klo, tb.kcnt, tb.mlo, tb.mcnt);
  ADVB_KERNEL_OK("fe_tables", stream);
  fe_dct_transpose_kernel<<<cdiv(NFILT * NCOEF, 256), 256, 0, stream>>>(tb.dct, tb.dctT);
  ADVB_KERNEL_OK("fe_dct_transpose", stream);
  return 0;
}

int frontend_forward(const FrontendTables& tb, const FrontendState& st, const float* x, int B, int T, float* dB,
                     float* out, long long clip_stride, long long stride_f, long long stride_c, long long offset,
                     cudaStream_t stream, float2* spec, FrontendHost* host) {
  const int F = frontend_frames(T);
  ADVB_CHECK(T >= 512, "clip shorter than one FFT frame");
  if (host == nullptr || host->dirty) {
    fe_reset_kernel<<<1, 1, 0, stream>>>(st);
    ADVB_KERNEL_OK("fe_reset", stream);
  }
  if (host != nullptr) host->dirty = true;
  const int n_fb = cdiv(F, 2 * FE_WARPS);
  const int g1 = n_fb * B < 148 * 3 ? n_fb * B : 148 * 3;  // persistent: 74 KB of shared memory -> 3 CTAs / SM
  fe_power_db_kernel<<<g1, FE_THREADS, fe_fwd_smem(), stream>>>(x, T, F, tb, st, dB, spec, n_fb, B);
  ADVB_KERNEL_OK("fe_power_db", stream);
  if (st.xr.world > 1) {
    fe_xrank_kernel<<<1, 32, 0, stream>>>(st, 0);
    ADVB_KERNEL_OK("fe_xrank_max", stream);
  }
  const int n_items = cdiv(B * F, FD_FR);
  const int g2 = n_items < 148 * 3 ? n_items : 148 * 3;  // persistent: 71 KB of shared memory -> 3 CTAs / SM
  fe_floor_dct_kernel<<<g2, 256, fe_dct_smem(), stream>>>(dB, F, tb, st, 80.0f, out, clip_stride, stride_f, stride_c,
                                                         offset, B * F);
  ADVB_KERNEL_OK("fe_floor_dct", stream);
  return 0;
}

int frontend_mel_spec(const float* x, const float* fb, int n_mels, float* out, int B, int T, cudaStream_t stream) {
  ADVB_CHECK(T >= 512, "clip shorter than one FFT frame");
  ADVB_CHECK(n_mels >= 1 && n_mels <= MS_MAX_MELS, "mel_spec: 1..128 mel filters");
  const int F = frontend_frames(T);
  const int n_fb = cdiv(F, 2 * FE_WARPS);
  const size_t smem = (size_t)(1024 + 2 * MS_MAX_MELS) * sizeof(float) + (size_t)FE_WARPS * (FFT_A + FFT_B) * sizeof(float2);
  ADVB_CUDA_OK(cudaFuncSetAttribute(fe_melspec_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));  // per device
  const int grid = n_fb * B < 148 * 3 ? n_fb * B : 148 * 3;
  fe_melspec_kernel<<<grid, FE_THREADS, smem, stream>>>(x, T, F, fb, n_mels, out, n_fb, B);
  ADVB_KERNEL_OK("fe_melspec", stream);
  return 0;
}

int frontend_backward(const FrontendTables& tb, const FrontendState& st, const float* x, int B, int T,
                      const float* dB, const float* gcoef, long long g_clip_stride, long long g_stride_f,
                      long long g_stride_c, float* mass_partial, float* gd, float* gx, cudaStream_t stream,
                      const FusedUpdate* upd, const float2* spec, FrontendHost* host) {
  const int F = frontend_frames(T);
  const int n_items = cdiv(B * F, DT_FR);
  const int g2 = n_items < 148 * 3 ? n_items : 148 * 3;  // persistent: 61 KB of shared memory -> 3 CTAs / SM
  fe_dct_t_kernel<<<g2, 256, fe_dct_t_smem(), stream>>>(gcoef, g_clip_stride, g_stride_f, g_stride_c, F, tb, gd, B * F, dB, st,
                                                       80.0f, mass_partial);
  ADVB_KERNEL_OK("fe_dct_t", stream);
  if (st.xr.world > 1) {
    fe_xrank_kernel<<<1, 32, 0, stream>>>(st, 1);
    ADVB_KERNEL_OK("fe_xrank_sum", stream);
  }
  const int n_tiles = cdiv(T, TILE_S);
  for (int tile = 0; tile < n_tiles; ++tile) {
    int t_lo, t_hi;
    tile_frames(tile * TILE_S, std::min(T, (tile + 1) * TILE_S), T, F, t_lo, t_hi);
    if (spec != nullptr) t_lo &= ~1;
    ADVB_CHECK(t_hi - t_lo + 1 <= NF_MAX, "frontend backward: a tile needs more frames than the kernel has warps for");
  }
  const int grid = n_tiles * B < 148 ? n_tiles * B : 148;  // one persistent CTA per SM (210 KB of shared memory each)
  FusedUpdate u{};
  if (upd != nullptr) {
    u = *upd;
    ADVB_CHECK(u.kind == 0 || (u.x_clean != nullptr && u.adv_out != nullptr && u.adv_out != x),
               "fused update needs the clean clips and an output buffer that does not alias the waveform");
  }
  fe_bwd_kernel<<<grid, FB_THREADS, fe_bwd_smem(), stream>>>(x, T, F, tb, st, 80.0f, gd, gx, spec, n_tiles, B, u);
  ADVB_KERNEL_OK("fe_bwd", stream);
  if (host != nullptr) host->dirty = false;  // the kernel's last block has reset the state
  return 0;
}

int frontend_clean(const FrontendState& st, FrontendHost* host, cudaStream_t stream) {
  if (host != nullptr && !host->dirty) return 0;
  fe_reset_kernel<<<1, 1, 0, stream>>>(st);
  ADVB_KERNEL_OK("fe_reset", stream);
  if (host != nullptr) host->dirty = false;
  return 0;
}

}  // namespace advb
