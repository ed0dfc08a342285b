// LCNN 3x3 convolution blocks on tcgen05 as PERSISTENT warp-specialised kernels (sm_100a): forward and input-gradient
// backward of blocks 2, 4, 6, 8 (conv -> +bias -> Max-Feature-Map -> [2x2 max-pool] -> [BatchNorm(eval)]).
//
// Same math, weight images, MMA order and epilogue arithmetic as conv_tc.cu (implicit GEMM over flattened padded
// pixels, 3xTF32, fp32 accumulators in TMEM): outputs are bit-identical.  What changes is the schedule.  conv_tc.cu runs
// fill -> MMA -> epilogue serially inside a one-tile CTA (tensor pipe 27-59 % busy, profiles/r01_ncu_full_persistent.md);
// here one CTA per SM loops over tiles with three roles:
//   * 8 worker warps: issue the global loads of unit u+1 (unit = one 32-channel chunk of one tile's halo band) into
//     registers, convert unit u (BN scale / un-pool / un-MFM for the backward, tf32 hi/lo split, SWIZZLE_128B store),
//     and run the epilogue of tile t-1 (TMEM -> registers -> staging -> pool / BN / coalesced stores);
//   * 1 MMA warp: waits for the band of unit u and for each weight slice, issues the (tap, k-step, pass, M-tile) nest
//     into accumulator t & 1, releases ring slots and the band with tcgen05.commit;
//   * 1 weight warp: streams the (chunk, tap) weight slices through a ring of 1-D TMA bulk copies, tile after tile.
// The MMAs of unit u therefore overlap the loads of u+1 and the epilogue of the previous tile; only the convert of u+1
// waits for them (single band buffer: 2 x 88 KB does not fit beside the staging tile).
#include <stdlib.h>

#include "conv.cuh"
#include "tc_common.cuh"

namespace advb {

namespace {

using namespace tc;

// worker threads per CTA: 8 warps forward, 16 warps backward (the backward band fill - un-pool / un-MFM / BN scale per
// element - is instruction-heavy; with 8 workers it, not the MMAs, set the pace); + 1 MMA warp + 1 weight warp
// M-tiles (128 pixels) per tile and band buffering are chosen per direction:
//   forward  : 2 M-tiles (pooled layers need an even number of image rows per tile), one band buffer;
//   backward : 1 M-tile and TWO band buffers, so the convert of unit u+1 overlaps the MMAs of unit u (3-4 chunks per
//              tile and N = 32/48 operand-read-bound MMAs made the single-buffer version convert-then-multiply serial).

#ifndef P3_NMW_ALL_FWD
#define P3_NMW_ALL_FWD 1  // 1: every forward variant issues its two M-tiles from two warps (0: PLAIN only)
#endif
__host__ __device__ constexpr int p3_pow2(int v) {
  int p = 32;
  while (p < v) p <<= 1;
  return p;
}

struct P3Args {
  int B, H, W, Ho, Wo;
  int R, tiles_per_clip, n_tiles, band_rows;
  const unsigned char* wpack;
  const float* in;
  float* out;
  int out_pad;
  unsigned char* codes;
  const float* bias;
  const float* bn_mean;
  const float* bn_invstd;
  const float* gout;
  const unsigned char* codes_in;
  float* gin;
  int passes;
  const float* mul_h;      // PLAIN: optional (B, H+2, W+2, NOUT) bordered tensor whose sign gates the output (LeakyReLU')
  const float* mul_scale;  // PLAIN: per-channel factor applied together with mul_h
  float mul_slope;
  int out_ch;              // PLAIN: channels stored per pixel (= the channel stride of out and mul_h); 0 = NOUT
  const float* aff_scale;  // PLAIN: optional per-channel affine (BatchNorm eval) ...
  const float* aff_shift;
  float act_slope;         // ... followed by LeakyReLU(act_slope), applied when aff_scale != nullptr
  FastDiv dTiles, dWp, dW, dWo;
  long long* prof;  // optional (ADVB_P3_PROF=1): per-phase cycle counts of CTA 0 (worker thread 0, MMA warp leader)
};

// HS ("horizontal scatter", backward only): the MMA computes Z[q][dx][ci] = sum_dy sum_co g[q + (dy-1) Wp][co] Wt[dy][dx][co][ci]
// with N = 3 C_in (one MMA per vertical tap instead of nine per 3x3 tap) and the epilogue gathers
// gin[q][ci] = sum_dx Z[q + dx - 1][dx][ci].  The plain backward has N = C_in = 32 / 48, and tcgen05.mma costs ~64 cycles
// per M=128, K=8 instruction however small N is (measured: block 2 backward ran at exactly that issue floor whatever the
// schedule), so 3x fewer, 3x wider MMAs cut its tensor time 3x.
// PLAIN (forward only): an ordinary 3x3 convolution - every output channel kept (no Max-Feature-Map pairing, no pooling, no code
// bytes), out = (acc + bias) [* (mul_h > 0 ? 1 : mul_slope) * mul_scale[c]].  SpecRNet's 64 -> 64 convolutions and, fed with the
// flipped / transposed weight image, their transposes (specrnet.cu).
// WMAX: widest image row the band is sized for (40: every LCNN block; 80: SpecRNet's first block, with 12 worker warps so that the
// larger band still is <= 10 prefetch items per thread).
template <int KTOT, int NOUT, bool POOL, bool BWD, bool HS, bool PLAIN = false, int WMAX = 40>
struct P3Cfg {
  static_assert(!HS || BWD, "horizontal scatter is a backward formulation");
  static_assert(!PLAIN || (!BWD && !POOL), "the plain variant is a forward convolution without pooling");
  static constexpr int PW = WMAX > 40 ? 384 : 256;
  static constexpr int NM3 = BWD ? 1 : 2;
  // MMA-issuing warps.  One thread issues an MMA every ~140-150 cycles whatever N is (descriptor arithmetic on the uniform datapath:
  // tools/tc_probe.cu section 5; ADVB_P3_PROF: 8.0 k cycles of issue per 2-M-tile unit of SpecRNet's first block with three OR two
  // MMAs per k-step), which at N <= 64 is longer than the MMAs take.  PLAIN: one issuing warp per M-tile - both wait on the same
  // full barriers and each commits to the empty / done barriers (arrival count NMW).
  static constexpr int NMW = P3_NMW_ALL_FWD ? NM3 : (PLAIN ? NM3 : 1);
  static constexpr int PT = PW + 32 * (1 + NMW);
  static constexpr int NTAP = HS ? 3 : 9;
  static constexpr int NMMA = HS ? 3 * NOUT : NOUT;  // MMA N = accumulator columns per M-tile
  static constexpr int BR = (NM3 * 128 + 2 * (WMAX + 2 + 1) + 7) & ~7;  // band rows per buffer (widest row: W = WMAX)
  static constexpr int NI_MAX = (BR * 8 + PW - 1) / PW;            // prefetch items per worker thread
  static constexpr int NKC = (KTOT + 31) / 32;
  static constexpr int NSLICE = NKC * NTAP;
  // WIDE (PLAIN only, 3xTF32): the weight slice holds its hi rows then its lo rows, NMMA x 128 B each, so ONE B descriptor with
  // N = 2 NMMA covers [W_hi | W_lo]: A_hi x [W_hi | W_lo] is one MMA into 2 NMMA accumulator columns and A_lo x W_hi a second one
  // into the first NMMA; the epilogue adds the two halves.  tcgen05.mma costs ~(128 + N) / 4 cycles per K = 8 step however small N
  // is (profiles/r01_tc_probe.log), so two MMAs (N = 64 + 32: 48 + 46.5 cycles) replace three (3 x 46.5) at SpecRNet's N = 32, and
  // 64 + 48 replace 3 x 48 at N = 64.  LCNN's forward blocks 2 / 4 (N = 96 / 128) cannot do this: two double-buffered M-tiles of
  // 2 N accumulator columns exceed the 512 TMEM columns.
  static constexpr bool WIDE = PLAIN;
  static constexpr int NACC = WIDE ? 2 * NMMA : NMMA;  // accumulator columns per M-tile
  static constexpr int NSTRIDE = p3_pow2(NACC);
  static constexpr int TMEM_COLS = p3_pow2(2 * NM3 * NSTRIDE);
  static constexpr int SLICE_BYTES = 2 * NMMA * 128;
  static constexpr int CS = (BWD || PLAIN) ? NMMA : NOUT / 2;  // staged floats per pixel
  static constexpr int SS = CS + 4;
  static constexpr int STAGE_BYTES = NM3 * 128 * SS * 4 + NM3 * 128 * 8;
  static constexpr int BUF_BYTES = 2 * BR * 128;                    // hi rows then lo rows
  // second band buffer (convert of unit u+1 overlaps the MMAs of unit u) wherever it fits beside staging + a 2-slice ring
  static constexpr bool DB = BWD && (227 * 1024 - 1024 - 2 * BUF_BYTES - STAGE_BYTES >= 2 * SLICE_BYTES);
  static constexpr int BAND_BYTES = (DB ? 2 : 1) * BUF_BYTES;
  static constexpr int ROOM = 227 * 1024 - 1024 - BAND_BYTES - STAGE_BYTES;
  // RESIDENT: every weight slice fits beside the band and the staging tile (SpecRNet's first block: 9 x 8 KB) - loaded once per CTA
  // instead of once per tile; the MMA warps waited ~1 k of 8.3 k cycles per tile for the ring's refills (ADVB_P3_PROF "wait_w")
  static constexpr bool RESIDENT = ROOM / SLICE_BYTES >= NSLICE;
  static constexpr int NST = RESIDENT ? NSLICE : (ROOM / SLICE_BYTES >= 4 ? 4 : (ROOM / SLICE_BYTES >= 3 ? 3 : 2));
  static constexpr size_t SMEM = (size_t)NST * SLICE_BYTES + BAND_BYTES + STAGE_BYTES + 1024;
  static_assert(TMEM_COLS <= 512, "two accumulator sets must fit the 512 TMEM columns");
  static_assert(ROOM / SLICE_BYTES >= 2, "weight ring does not fit");
};

template <int KTOT, int NOUT, bool POOL, bool BWD, bool HS, bool PLAIN = false, int WMAX = 40>
__global__ void __launch_bounds__(P3Cfg<KTOT, NOUT, POOL, BWD, HS, PLAIN, WMAX>::PT, 1) conv_p3_kernel(P3Args a) {
  using Cfg = P3Cfg<KTOT, NOUT, POOL, BWD, HS, PLAIN, WMAX>;
  constexpr int PW = Cfg::PW, PT = Cfg::PT;
  constexpr int NKC = Cfg::NKC, NSLICE = Cfg::NSLICE, NST = Cfg::NST, CS = Cfg::CS, SS = Cfg::SS;
  constexpr int NM3 = Cfg::NM3, NI_MAX = Cfg::NI_MAX, BR = Cfg::BR, NTAP = Cfg::NTAP, NMMA = Cfg::NMMA;
  constexpr bool DB = Cfg::DB;
  constexpr int KSU = BWD ? 1 : 4;
  constexpr uint32_t IDESC = idesc_tf32(128, Cfg::NMMA);
  constexpr uint32_t IDESC_WIDE = idesc_tf32(128, Cfg::NACC);

  extern __shared__ unsigned char smem_raw[];
  unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  unsigned char* band = base;                        // [buffer][hi BR x 128 B | lo BR x 128 B]
  unsigned char* wring = band + Cfg::BAND_BYTES;     // BUF_BYTES is a multiple of 1024: stays 1024-byte aligned
  float* stage = reinterpret_cast<float*>(wring + (size_t)NST * Cfg::SLICE_BYTES);
  unsigned long long* flags = reinterpret_cast<unsigned long long*>(stage + (size_t)NM3 * 128 * SS);
  // bar_band_full[b]: one per band buffer.  With a single barrier and two buffers the workers could complete the phase of unit
  // u + 1 before the MMA warp had looked at the phase of unit u (parity waits cannot tell two completed phases from none:
  // compute-sanitizer synccheck, "missing wait"); per buffer, unit u + 2 is staged only after the MMAs of unit u completed.
  __shared__ uint64_t bar_wfull[NST], bar_wempty[NST], bar_band_full[2], bar_unit_done[2];
  __shared__ uint32_t tmem_base_s;
  __shared__ __align__(16) float s_bias[BWD ? 4 : NOUT];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int Wp = a.W + 2, Hp = a.H + 2;
  const int Heff = (!BWD && POOL) ? 2 * a.Ho : a.H;
  const int npix = Hp * Wp;

  if (tid == 0) {
    for (int s = 0; s < NST; ++s) {
      mbar_init(&bar_wfull[s], 1);
      mbar_init(&bar_wempty[s], Cfg::NMW);
    }
    mbar_init(&bar_band_full[0], PW / 32);
    mbar_init(&bar_band_full[1], PW / 32);
    mbar_init(&bar_unit_done[0], Cfg::NMW);
    mbar_init(&bar_unit_done[1], Cfg::NMW);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc<Cfg::TMEM_COLS>(&tmem_base_s);
  if (!BWD)
    for (int i = tid; i < NOUT; i += PT) s_bias[i] = a.bias != nullptr ? __ldg(a.bias + i) : 0.f;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;

  // geometry of a tile (rows y0 .. y0+rows-1 of clip b)
  auto tile_geom = [&](int tile, int& b, int& y0, int& rows, int& nM) {
    b = fdiv(tile, a.dTiles);
    const int tr = tile - b * a.tiles_per_clip;
    y0 = tr * a.R;
    rows = min(a.R, Heff - y0);
    nM = (rows * Wp + 127) >> 7;
  };

  if (warp == PW / 32 + Cfg::NMW) {
    // ================= weight warp: TMA ring =================
    int s_glob = 0;
    if (Cfg::RESIDENT) {
      if (blockIdx.x < a.n_tiles && lane == 0)
        for (int sl = 0; sl < NSLICE; ++sl) {
          mbar_expect_tx(&bar_wfull[sl], Cfg::SLICE_BYTES);
          bulk_g2s(wring + (size_t)sl * Cfg::SLICE_BYTES, a.wpack + (size_t)sl * Cfg::SLICE_BYTES, Cfg::SLICE_BYTES, &bar_wfull[sl]);
        }
      __syncwarp();
    }
    for (int tile = blockIdx.x; !Cfg::RESIDENT && tile < a.n_tiles; tile += gridDim.x) {
      for (int sl = 0; sl < NSLICE; ++sl, ++s_glob) {
        const int slot = s_glob % NST, use = s_glob / NST;
        if (use > 0) mbar_wait(&bar_wempty[slot], (uint32_t)((use - 1) & 1));
        if (lane == 0) {
          mbar_expect_tx(&bar_wfull[slot], Cfg::SLICE_BYTES);
          bulk_g2s(wring + (size_t)slot * Cfg::SLICE_BYTES, a.wpack + (size_t)sl * Cfg::SLICE_BYTES, Cfg::SLICE_BYTES,
                   &bar_wfull[slot]);
        }
        __syncwarp();
      }
    }
  } else if (warp >= PW / 32) {
    // ================= MMA warp(s) =================
    const bool leader = elect_one();
    const int mw = warp - PW / 32;  // this warp issues M-tiles mw, mw + NMW, ...
    int u_glob = 0, s_glob = 0, it = 0;
    const bool mprof = a.prof != nullptr && blockIdx.x == 0 && leader && mw == 0;
    long long mc[3] = {0, 0, 0}, m_last = clock64();
#define P3MPROF(k)                    \
  if (mprof) {                        \
    const long long now_ = clock64(); \
    mc[k] += now_ - m_last;           \
    m_last = now_;                    \
  }
    for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x, ++it) {
      int b, y0, rows, nM;
      tile_geom(tile, b, y0, rows, nM);
      const uint32_t dbase = tmem + (uint32_t)(it & 1) * NM3 * Cfg::NSTRIDE;
#pragma unroll 1
      for (int kc = 0; kc < NKC; ++kc, ++u_glob) {
        P3MPROF(2)
        mbar_wait(&bar_band_full[DB ? (u_glob & 1) : 0], (uint32_t)(DB ? ((u_glob >> 1) & 1) : (u_glob & 1)));
        tc_fence_after();
        P3MPROF(0)
        const uint32_t a_hi_addr = smem_u32(band + (DB ? (size_t)(u_glob & 1) * Cfg::BUF_BYTES : 0));
        const uint32_t a_lo_addr = a_hi_addr + BR * 128;
        const int kvalid = (KTOT - 32 * kc) >= 32 ? 32 : (KTOT - 32 * kc);
#pragma unroll 1
        for (int tap = 0; tap < NTAP; ++tap, ++s_glob) {
          const int slot = Cfg::RESIDENT ? kc * NTAP + tap : s_glob % NST;
          P3MPROF(2)
          mbar_wait(&bar_wfull[slot], Cfg::RESIDENT ? 0u : (uint32_t)((s_glob / NST) & 1));  // resident: phase 0 completes once
          tc_fence_after();
          P3MPROF(1)
          const uint32_t w_hi = smem_u32(wring + (size_t)slot * Cfg::SLICE_BYTES), w_lo = w_hi + NMMA * 128;
          const int dy = HS ? tap : tap / 3, dx = HS ? 1 : tap - dy * 3;  // HS: vertical taps only, centre column
          const uint32_t row_off = (uint32_t)(dy * Wp + dx) * 128u;
          // Forward: k-steps unrolled - the descriptor arithmetic of a step is a chain of uniform-datapath instructions (~140 cycles
          // when the steps run one after the other: tools/tc_probe.cu section 5).  Backward: NOT unrolled - denser MMA issue slowed the
          // worker warps' fill by more than it saved (block 2: 388 -> 445 us, profiles/r02_p3_phases.md).
#pragma unroll KSU
          for (int ks = 0; ks < 4; ++ks) {
            if (ks >= kvalid / 8) break;
            const uint64_t bh = desc_sw128(w_hi + ks * 32), bl = desc_sw128(w_lo + ks * 32);
            const uint64_t ah = desc_sw128(a_hi_addr + row_off + ks * 32), al = desc_sw128(a_lo_addr + row_off + ks * 32);
            const uint32_t acc0 = (kc > 0 || tap > 0 || ks > 0) ? 1u : 0u;
#pragma unroll
            for (int m0 = 0; m0 < NM3; m0 += Cfg::NMW) {
              const int m = m0 + mw;
              if (m < nM && leader) {
                const uint32_t dcol = dbase + m * Cfg::NSTRIDE;
                const uint64_t moff = (uint64_t)(m * ((128 * 128) >> 4));
                if (Cfg::WIDE && a.passes == 3) {
                  mma_tf32(dcol, ah + moff, bh, IDESC_WIDE, acc0);  // columns [0, N): a_hi w_hi, [N, 2N): a_hi w_lo
                  mma_tf32(dcol, al + moff, bh, IDESC, 1u);         // columns [0, N) += a_lo w_hi
                  continue;
                }
                mma_tf32(dcol, ah + moff, bh, IDESC, acc0);
                if (a.passes == 3) {
                  mma_tf32(dcol, ah + moff, bl, IDESC, 1u);
                  mma_tf32(dcol, al + moff, bh, IDESC, 1u);
                } else if (a.passes == 2) {  // both cross terms as one bf16 MMA over the interleaved second planes
                  mma_bf16(dcol, al + moff, bl, idesc_bf16(128, Cfg::NMMA), 1u);
                }
              }
            }
          }
          if (leader && !Cfg::RESIDENT) mma_commit(&bar_wempty[slot]);  // slot free once these MMAs have read it
          __syncwarp();
        }
        if (leader) mma_commit(&bar_unit_done[DB ? (u_glob & 1) : 0]);  // band buffer free; last chunk: accumulator complete
        __syncwarp();
      }
    }
    P3MPROF(2)
    if (mprof) {
      a.prof[8] = mc[0];
      a.prof[9] = mc[1];
      a.prof[10] = mc[2];
    }
  } else {
    // ================= worker warps =================
    const bool prof = a.prof != nullptr && blockIdx.x == 0 && tid == 0;
    long long pc[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0}, t_last = clock64();
#define P3PROF(k)                     \
  if (prof) {                         \
    const long long now_ = clock64(); \
    pc[k] += now_ - t_last;           \
    t_last = now_;                    \
  }
    const int c4 = tid & 7, r0 = tid >> 3;  // channel group; first band row (rows r0 + (PW / 8) u)
    constexpr int RSTEP = PW / 8;
    float4 rv[NI_MAX];
    uchar4 rc[NI_MAX];
    unsigned rok = 0;  // bit u: item u is a real element (loads stay RAW in registers until convert)
    unsigned char rwant[NI_MAX];
    // Backward: the pixel a band row holds does not depend on the tile beyond its first image row y0 - band row r >= 1 is padded
    // pixel (y0 + (r - 1) / Wp, (r - 1) % Wp) - so each item's column, pooled-column offset and row offset are computed ONCE per
    // kernel (the per-unit index math - a division, 64-bit address arithmetic and five range checks per item - was 46 % of the
    // worker time of the block-2 backward: ADVB_P3_PROF phase counters, profiles/r02_p3_phases.md).
    int geo[BWD ? NI_MAX : 1];  // dy + 2 (bits 0-4) | x valid (bit 5) | x parity (bit 6) | pooled-column element offset (bits 8..)
    if (BWD) {
      constexpr int Ch = KTOT / 2;
#pragma unroll
      for (int u = 0; u < NI_MAX; ++u) {
        const int r = r0 + RSTEP * u;
        const int rr = r > 0 ? r - 1 : 0;
        const int dyp = fdiv(rr, a.dWp), xp = rr - dyp * Wp;
        const int x = xp - 1;
        const int px = POOL ? (x >> 1) : x;
        const bool xok = r > 0 && x >= 0 && x < a.W && px < a.Wo;
        geo[u] = ((dyp + 1) & 31) | (xok ? 32 : 0) | ((x & 1) << 6) | ((xok ? px * Ch : 0) << 8);
      }
    }

    // BatchNorm(eval) scale of this thread's channel group, per chunk (was a dependent global load at the head of every convert)
    float4 sc_kc[BWD ? NKC : 1];
    if (BWD) {
      constexpr int Ch = KTOT / 2;
#pragma unroll
      for (int k = 0; k < NKC; ++k) {
        const int ch = 32 * k + 4 * c4;
        sc_kc[k] = make_float4(1.f, 1.f, 1.f, 1.f);
        if (a.bn_invstd != nullptr && ch < KTOT) sc_kc[k] = __ldg(reinterpret_cast<const float4*>(a.bn_invstd + (ch >= Ch ? ch - Ch : ch)));
      }
    }
    const uint32_t off0 = sw128_chunk(r0, c4);  // band row r0 + 32 u keeps r0's swizzle phase: its chunk sits u * 4096 bytes further

    auto issue_loads = [&](int tile, int kc) {
      int b, y0, rows, nM;
      tile_geom(tile, b, y0, rows, nM);
      const int q_lo = y0 * Wp - 1;
      const int band_used = nM * 128 + 2 * (Wp + 1);
      const int ch = 32 * kc + 4 * c4;
      const bool ch_ok = ch < KTOT;
      rok = 0;
      if (BWD) {
        constexpr int Ch = KTOT / 2;
        const int half = ch >= Ch ? 1 : 0;
        const int c = ch - half * Ch;
        const float* gbase = a.gout + (size_t)b * a.Ho * a.Wo * Ch + c;
        const unsigned char* cbase = a.codes_in + (size_t)b * a.Ho * a.Wo * Ch + c;
        const int row_elems = a.Wo * Ch;
#pragma unroll
        for (int u = 0; u < NI_MAX; ++u) {
          const int r = r0 + RSTEP * u;
          const int g = geo[u];
          const int y = y0 + (g & 31) - 2;
          const int py = POOL ? (y >> 1) : y;
          const bool ok = ch_ok && (g & 32) != 0 && r < band_used && y >= 0 && y < a.H && py < a.Ho;
          const int o = ok ? py * row_elems + (g >> 8) : 0;
          rv[u] = __ldg(reinterpret_cast<const float4*>(gbase + o));
          rc[u] = __ldg(reinterpret_cast<const uchar4*>(cbase + o));
          rwant[u] = (unsigned char)(ok ? ((POOL ? (((y & 1) << 1) | ((g >> 6) & 1)) : 0) | (half << 2)) : 0xff);
        }
      } else {
        const float* inb = a.in + (size_t)b * npix * KTOT + ch;
#pragma unroll
        for (int u = 0; u < NI_MAX; ++u) {
          const int r = r0 + RSTEP * u;
          const int q = q_lo + r;
          const bool ok = ch_ok && r < band_used && q >= 0 && q < npix;
          rv[u] = __ldg(reinterpret_cast<const float4*>(ok ? inb + (size_t)q * KTOT : a.in));
          rok |= (ok ? 1u : 0u) << u;
        }
      }
    };

    auto convert_store = [&](int tile, int kc, int bufsel) {
      int b, y0, rows, nM;
      tile_geom(tile, b, y0, rows, nM);
      unsigned char* a_hi = band + (size_t)bufsel * Cfg::BUF_BYTES;
      unsigned char* a_lo = a_hi + BR * 128;
      const int band_used = nM * 128 + 2 * (Wp + 1);
      float4 sc = make_float4(1.f, 1.f, 1.f, 1.f);
      if (BWD) {
#pragma unroll
        for (int k = 0; k < NKC; ++k)
          if (k == kc) sc = sc_kc[k];
      }
#pragma unroll
      for (int u = 0; u < NI_MAX; ++u) {
        const int r = r0 + RSTEP * u;
        if (r < band_used) {
          float4 v;
          if (BWD) {
            const unsigned want = rwant[u];
            v.x = rc[u].x == want ? rv[u].x * sc.x : 0.f;
            v.y = rc[u].y == want ? rv[u].y * sc.y : 0.f;
            v.z = rc[u].z == want ? rv[u].z * sc.z : 0.f;
            v.w = rc[u].w == want ? rv[u].w * sc.w : 0.f;
          } else {
            v = ((rok >> u) & 1u) ? rv[u] : make_float4(0.f, 0.f, 0.f, 0.f);
          }
          float4 hi, lo;
          split_tf32(v.x, hi.x, lo.x);
          split_tf32(v.y, hi.y, lo.y);
          split_tf32(v.z, hi.z, lo.z);
          split_tf32(v.w, hi.w, lo.w);
          const uint32_t off = off0 + (uint32_t)u * (RSTEP * 128);
          *reinterpret_cast<float4*>(a_hi + off) = hi;
          if (!BWD && a.passes == 2) {
            // mixed mode: the second plane holds, per 16-byte chunk, bf16(a) x 4 then bf16(a - tf32(a)) x 4: ONE bf16 MMA
            // (K = 16) against [bf16(w_lo) x 4 | bf16(w) x 4] adds both cross terms a_hi w_lo + a_lo w_hi
            uint4 xw;
            xw.x = pack_bf16x2(v.x, v.y), xw.y = pack_bf16x2(v.z, v.w);
            xw.z = pack_bf16x2(lo.x, lo.y), xw.w = pack_bf16x2(lo.z, lo.w);
            *reinterpret_cast<uint4*>(a_lo + off) = xw;
          } else {
            *reinterpret_cast<float4*>(a_lo + off) = lo;
          }
        }
      }
    };

    auto epilogue = [&](int tile, int buf) {
      int b, y0, rows, nM;
      tile_geom(tile, b, y0, rows, nM);
      P3PROF(4)
      if (PLAIN && a.mul_h != nullptr) {
        // the gating tensor's lines of this tile into L2 / L1 now: the TMEM -> stage phase below (~2 k cycles) hides their DRAM
        // latency, which the store loop otherwise pays once per pass
        const int och = a.out_ch > 0 ? a.out_ch : NOUT;
        const float* mul_tile = a.mul_h + (((size_t)b * Hp + y0 + 1) * Wp + 1) * och;
        const int lines = (rows * Wp * och + 31) >> 5;  // 128-byte lines of the rows' contiguous span (borders included)
        for (int i = tid; i < lines; i += PW) asm volatile("prefetch.global.L2 [%0];" ::"l"(mul_tile + (size_t)i * 32));
      }
      asm volatile("bar.sync 1, %0;" ::"n"(PW) : "memory");  // previous staging tile fully consumed
      P3PROF(6)
      {
        const int wq = warp & 3, wg = warp >> 2;
        for (int m = wg; m < nM; m += 2) {
          const int r = m * 128 + wq * 32 + lane;
          const uint32_t taddr = tmem + ((uint32_t)(wq * 32) << 16) + (uint32_t)(buf * NM3 + m) * Cfg::NSTRIDE;
          float* srow = stage + (size_t)r * SS;
          if (BWD) {
#pragma unroll 1
            for (int c0 = 0; c0 < NMMA; c0 += 16) {
              uint32_t v[16];
              tmem_ld16_issue(taddr + c0, v);
              tmem_ld_wait();
#pragma unroll
              for (int j = 0; j < 16; j += 4)
                *reinterpret_cast<float4*>(srow + c0 + j) = make_float4(
                    __uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
            }
          } else if (PLAIN) {
            const uint32_t srow_s = smem_u32(srow);
#pragma unroll
            for (int c0 = 0; c0 < NOUT; c0 += 16) {
              uint32_t v[16];
              tmem_ld16_issue(taddr + c0, v);
              if (Cfg::WIDE && a.passes == 3) {  // second accumulator half: the a_hi w_lo cross term
                uint32_t v2[16];
                tmem_ld16_issue(taddr + NMMA + c0, v2);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 16; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) + __uint_as_float(v2[j]));
              } else {
                tmem_ld_wait();
              }
#pragma unroll
              for (int j = 0; j < 16; j += 4) {
                const float4 b4 = *reinterpret_cast<const float4*>(s_bias + c0 + j);
                asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(srow_s + (uint32_t)(c0 + j) * 4u),
                             "f"(__uint_as_float(v[j]) + b4.x), "f"(__uint_as_float(v[j + 1]) + b4.y),
                             "f"(__uint_as_float(v[j + 2]) + b4.z), "f"(__uint_as_float(v[j + 3]) + b4.w)
                             : "memory");
              }
            }
          } else {
            // c0 unrolled: flag bits become immediates, the bias comes by LDS.128 and the row goes out by st.shared (the generic
            // stores emitted before cost an address-space check each): 92 instructions per 16 channels instead of 210; this phase
            // was 4.4-6 k cycles per tile (ADVB_P3_PROF "TMEM->stage"), issue-bound, not TMEM-latency-bound
            uint32_t fl0 = 0u, fl1 = 0u;
            const uint32_t srow_s = smem_u32(srow);
#pragma unroll
            for (int c0 = 0; c0 < CS; c0 += 16) {
              uint32_t lo[16], hi[16];
              tmem_ld16_issue(taddr + c0, lo);
              tmem_ld16_issue(taddr + CS + c0, hi);
              tmem_ld_wait();
#pragma unroll
              for (int j = 0; j < 16; j += 4) {
                const float4 bl = *reinterpret_cast<const float4*>(s_bias + c0 + j);
                const float4 bh = *reinterpret_cast<const float4*>(s_bias + CS + c0 + j);
                const float bla[4] = {bl.x, bl.y, bl.z, bl.w}, bha[4] = {bh.x, bh.y, bh.z, bh.w};
                float m4[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                  const float l = __uint_as_float(lo[j + k]) + bla[k];
                  const float h = __uint_as_float(hi[j + k]) + bha[k];
                  const bool sel = h > l;
                  m4[k] = sel ? h : l;
                  const int bit = c0 + j + k;
                  if (bit < 32) fl0 |= sel ? (1u << bit) : 0u;
                  else fl1 |= sel ? (1u << (bit - 32)) : 0u;
                }
                asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(srow_s + (uint32_t)(c0 + j) * 4u), "f"(m4[0]), "f"(m4[1]),
                             "f"(m4[2]), "f"(m4[3])
                             : "memory");
              }
            }
            flags[r] = ((unsigned long long)fl1 << 32) | fl0;
          }
        }
      }
      P3PROF(7)
      tc_fence_before();
      asm volatile("bar.sync 1, %0;" ::"n"(PW) : "memory");
      P3PROF(8)
      constexpr int C4 = CS / 4;
      if (BWD) {
        constexpr int O4 = NOUT / 4;  // float4 groups of the C_in output channels
        const int items = rows * a.W * O4;
        for (int i = tid; i < items; i += PW) {
          const int ic = i / O4, c4i = i - O4 * ic, yl = fdiv(ic, a.dW), x = ic - yl * a.W;
          const float* sp = stage + (size_t)(yl * Wp + x + 1) * SS + 4 * c4i;
          float4 v;
          if (HS) {  // gin[q] = Z[q-1][dx=0] + Z[q][dx=1] + Z[q+1][dx=2]
            const float4 z0 = *reinterpret_cast<const float4*>(sp - SS);
            const float4 z1 = *reinterpret_cast<const float4*>(sp + NOUT);
            const float4 z2 = *reinterpret_cast<const float4*>(sp + SS + 2 * NOUT);
            v = make_float4(z0.x + z1.x + z2.x, z0.y + z1.y + z2.y, z0.z + z1.z + z2.z, z0.w + z1.w + z2.w);
          } else {
            v = *reinterpret_cast<const float4*>(sp);
          }
          *reinterpret_cast<float4*>(a.gin + (((size_t)b * a.H + y0 + yl) * a.W + x) * NOUT + 4 * c4i) = v;
        }
      } else if (PLAIN) {
        const int Hop = a.Ho + 2 * a.out_pad, Wop = a.Wo + 2 * a.out_pad;
        const int items = rows * a.Wo * C4;
        const int och = a.out_ch > 0 ? a.out_ch : NOUT;  // NOUT may be padded beyond the stored channels (24 -> 32)
        float* out_tile = a.out + (((size_t)b * Hop + y0 + a.out_pad) * Wop + a.out_pad) * och;
        const bool has_mul = a.mul_h != nullptr;
        const float* mul_tile = has_mul ? a.mul_h + (((size_t)b * Hp + y0 + 1) * Wp + 1) * och : nullptr;
        // EU items per thread and pass, the gating tensor's loads of all of them issued before the first is used: one item at a time,
        // each (dependent, DRAM-latency) load of mul_h stood alone and this loop was 5.4 k of the transposed convolution's 10.9 k
        // cycles per tile in SpecRNet's first block (ADVB_P3_PROF)
        constexpr int EU = PW > 256 ? 3 : 5;
        for (int i0 = tid; i0 < items; i0 += EU * PW) {
          float4 hv[EU];
          int sidx[EU], oidx[EU], cc[EU];
#pragma unroll
          for (int u = 0; u < EU; ++u) {
            const int i = i0 + u * PW;
            const int ic = i / C4, c = 4 * (i - C4 * ic), yl = fdiv(ic, a.dWo), x = ic - yl * a.Wo;
            cc[u] = (i < items && c < och) ? c : -1;
            sidx[u] = (yl * Wp + x + 1) * SS + c;
            oidx[u] = (yl * Wop + x) * och + c;
            hv[u] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (has_mul && cc[u] >= 0) hv[u] = __ldg(reinterpret_cast<const float4*>(mul_tile + (yl * Wp + x) * och + c));
          }
#pragma unroll
          for (int u = 0; u < EU; ++u) {
            const int c = cc[u];
            if (c < 0) continue;
            float4 v = *reinterpret_cast<const float4*>(stage + sidx[u]);
            if (a.aff_scale != nullptr) {  // BatchNorm(eval) + LeakyReLU of the reference's conv1 -> bn2 -> lrelu
              const float4 sc4 = __ldg(reinterpret_cast<const float4*>(a.aff_scale + c));
              const float4 sh4 = __ldg(reinterpret_cast<const float4*>(a.aff_shift + c));
              v.x = fmaf(v.x, sc4.x, sh4.x), v.y = fmaf(v.y, sc4.y, sh4.y), v.z = fmaf(v.z, sc4.z, sh4.z), v.w = fmaf(v.w, sc4.w, sh4.w);
              v.x = v.x > 0.f ? v.x : a.act_slope * v.x;
              v.y = v.y > 0.f ? v.y : a.act_slope * v.y;
              v.z = v.z > 0.f ? v.z : a.act_slope * v.z;
              v.w = v.w > 0.f ? v.w : a.act_slope * v.w;
            }
            if (has_mul) {  // * LeakyReLU'(h) * per-channel scale: the transposed convolution's chain-rule factor
              const float4 sv = __ldg(reinterpret_cast<const float4*>(a.mul_scale + c));
              v.x = v.x * (hv[u].x > 0.f ? 1.0f : a.mul_slope) * sv.x;
              v.y = v.y * (hv[u].y > 0.f ? 1.0f : a.mul_slope) * sv.y;
              v.z = v.z * (hv[u].z > 0.f ? 1.0f : a.mul_slope) * sv.z;
              v.w = v.w * (hv[u].w > 0.f ? 1.0f : a.mul_slope) * sv.w;
            }
            *reinterpret_cast<float4*>(out_tile + oidx[u]) = v;
          }
        }
      } else {
        const int Hop = a.Ho + 2 * a.out_pad, Wop = a.Wo + 2 * a.out_pad;
        const int orows = POOL ? rows / 2 : rows, oy0 = POOL ? y0 / 2 : y0;
        const int items = orows * a.Wo * C4;
        for (int i = tid; i < items; i += PW) {
          const int ic = i / C4, c4i = i - C4 * ic, yl = fdiv(ic, a.dWo), x = ic - yl * a.Wo;
          const int c = 4 * c4i;
          float4 v;
          uchar4 cd;
          if (POOL) {
            const int r00 = (2 * yl) * Wp + 2 * x + 1;
            const int rr[4] = {r00, r00 + 1, r00 + Wp, r00 + Wp + 1};
            float best[4];
            unsigned code[4];
#pragma unroll
            for (int p = 0; p < 4; ++p) {
              const float4 t = *reinterpret_cast<const float4*>(stage + (size_t)rr[p] * SS + c);
              const unsigned f = (unsigned)(flags[rr[p]] >> c) & 15u;
              const float tv[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
              for (int k = 0; k < 4; ++k)
                if (p == 0 || tv[k] > best[k]) {
                  best[k] = tv[k];
                  code[k] = (((f >> k) & 1u) << 2) | (unsigned)p;
                }
            }
            v = make_float4(best[0], best[1], best[2], best[3]);
            cd = make_uchar4((unsigned char)code[0], (unsigned char)code[1], (unsigned char)code[2], (unsigned char)code[3]);
          } else {
            const int r = yl * Wp + x + 1;
            v = *reinterpret_cast<const float4*>(stage + (size_t)r * SS + c);
            const unsigned f = (unsigned)(flags[r] >> c) & 15u;
            cd = make_uchar4((unsigned char)((f & 1u) << 2), (unsigned char)(((f >> 1) & 1u) << 2),
                             (unsigned char)(((f >> 2) & 1u) << 2), (unsigned char)(((f >> 3) & 1u) << 2));
          }
          if (a.bn_mean != nullptr) {
            const float4 is = __ldg(reinterpret_cast<const float4*>(a.bn_invstd + c));
            v.x = (v.x - __ldg(a.bn_mean + c)) * is.x;
            v.y = (v.y - __ldg(a.bn_mean + c + 1)) * is.y;
            v.z = (v.z - __ldg(a.bn_mean + c + 2)) * is.z;
            v.w = (v.w - __ldg(a.bn_mean + c + 3)) * is.w;
          }
          const int oy = oy0 + yl;
          *reinterpret_cast<float4*>(a.out + (((size_t)b * Hop + oy + a.out_pad) * Wop + x + a.out_pad) * CS + c) = v;
          *reinterpret_cast<uchar4*>(a.codes + (((size_t)b * a.Ho + oy) * a.Wo + x) * CS + c) = cd;
        }
      }
    };

    // ---- persistent loop over units (tile, chunk) ----
    int tile = blockIdx.x, kc = 0, u_glob = 0, it = 0, prev_tile = -1;
    if (tile < a.n_tiles) issue_loads(tile, 0);
    // unit u commits bar_unit_done[DB ? u & 1 : 0]; its k-th completion there has parity k & 1
    auto wait_unit = [&](int u) {
      if (DB) mbar_wait(&bar_unit_done[u & 1], (uint32_t)((u >> 1) & 1));
      else mbar_wait(&bar_unit_done[0], (uint32_t)(u & 1));
      tc_fence_after();
    };
    while (tile < a.n_tiles) {
      // the band buffer this unit overwrites must have been read completely: unit u-1 (single buffer) / u-2 (double)
      P3PROF(5)
      if (u_glob >= (DB ? 2 : 1)) wait_unit(u_glob - (DB ? 2 : 1));
      P3PROF(0)
      convert_store(tile, kc, DB ? (u_glob & 1) : 0);
      P3PROF(1)
      fence_proxy_async();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bar_band_full[DB ? (u_glob & 1) : 0]);
      P3PROF(2)
      // next unit
      int ntile = tile, nkc = kc + 1;
      if (nkc == NKC) {
        nkc = 0;
        ntile = tile + gridDim.x;
      }
      if (ntile < a.n_tiles) issue_loads(ntile, nkc);
      P3PROF(3)
      if (kc == 0 && it > 0) {  // overlaps the MMAs of this tile
        if (DB) wait_unit(u_glob - 1);  // last chunk of the previous tile (already awaited when single-buffered)
        P3PROF(0)
        epilogue(prev_tile, (it - 1) & 1);
        P3PROF(4)
      }
      if (nkc == 0) {
        prev_tile = tile;
        ++it;
      }
      tile = ntile;
      kc = nkc;
      ++u_glob;
    }
    if (it > 0) {
      wait_unit(u_glob - 1);
      epilogue(prev_tile, (it - 1) & 1);
    }
    if (prof) {
      for (int k = 0; k < 6; ++k) a.prof[k] = pc[k];
      a.prof[6] = it;
      a.prof[12] = pc[6], a.prof[13] = pc[7], a.prof[14] = pc[8];
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<Cfg::TMEM_COLS>(tmem);
}

int tune_p3() {
  static const int v = [] {
    const char* e = getenv("ADVB_P3");
    return e != nullptr ? atoi(e) : 1;
  }();
  return v;
}

template <int KTOT, int NOUT, bool POOL, bool BWD, bool HS = false, bool PLAIN = false, int WMAX = 40>
int launch_p3(P3Args a, const char* tag, cudaStream_t stream) {
  using Cfg = P3Cfg<KTOT, NOUT, POOL, BWD, HS, PLAIN, WMAX>;
  const int Wp = a.W + 2;
  const int Heff = (!BWD && POOL) ? 2 * a.Ho : a.H;
  const bool even = !BWD && POOL;
  constexpr int NM3 = Cfg::NM3;
  int Rmax = (NM3 * 128) / Wp;
  if (even) Rmax &= ~1;
  ADVB_CHECK(Rmax >= (even ? 2 : 1) && Wp <= WMAX + 2, "persistent 3x3 conv: image too wide for the tile");
  int tiles = cdiv(Heff, Rmax);
  int R = cdiv(Heff, tiles);
  if (even && (R & 1)) ++R;
  if (R > Rmax) R = Rmax;
  tiles = cdiv(Heff, R);
  a.R = R;
  a.tiles_per_clip = tiles;
  a.n_tiles = a.B * tiles;
  a.dTiles = make_fastdiv(tiles);
  a.dWp = make_fastdiv(Wp);
  a.dW = make_fastdiv(a.W);
  a.dWo = make_fastdiv(a.Wo);
  ADVB_CHECK(cdiv(R * Wp, 128) * 128 + 2 * (Wp + 1) <= Cfg::BR, "persistent 3x3 conv: band exceeds its buffer");
  auto kern = conv_p3_kernel<KTOT, NOUT, POOL, BWD, HS, PLAIN, WMAX>;
  ADVB_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM));
  int n_sm = 148, dev = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
  const int grid = a.n_tiles < n_sm ? a.n_tiles : n_sm;
  static long long* prof_buf = nullptr;
  static const bool want_prof = getenv("ADVB_P3_PROF") != nullptr;
  if (want_prof && prof_buf == nullptr) cudaMalloc(reinterpret_cast<void**>(&prof_buf), 16 * sizeof(long long));
  a.prof = want_prof ? prof_buf : nullptr;
  kern<<<grid, Cfg::PT, Cfg::SMEM, stream>>>(a);
  ADVB_KERNEL_OK(tag, stream);
  if (want_prof) {  // diagnostic only: synchronises
    long long hp[16];
    cudaStreamSynchronize(stream);
    cudaMemcpy(hp, prof_buf, sizeof(hp), cudaMemcpyDeviceToHost);
    const long long n = hp[6] > 0 ? hp[6] : 1;
    fprintf(stderr, "[p3 %s] NST %d DB %d tiles/cta %lld | worker cycles/tile: wait_unit %lld convert %lld fence+arrive %lld issue_loads %lld "
                    "epilogue %lld other %lld | mma warp: wait_band %lld wait_w %lld issue %lld\n",
            tag, Cfg::NST, (int)Cfg::DB, hp[6], hp[0] / n, hp[1] / n, hp[2] / n, hp[3] / n, hp[4] / n, hp[5] / n, hp[8] / n, hp[9] / n, hp[10] / n);
    fprintf(stderr, "      epilogue split: entry barrier %lld, TMEM->stage %lld, barrier %lld, pool/gather+store (rest of 'epilogue') \n", hp[12] / n,
            hp[13] / n, hp[14] / n);
  }
  return 0;
}

// HS weight slices in consumption order [chunk kc][dy]: K-major SWIZZLE_128B matrix [3 C_in rows][32 k], hi image then
// lo image; row n = dx * C_in + ci, column k <-> co = 32 kc + k; value = W[co][ci][2 - dy][2 - dx] (transposed conv).
__global__ void pack_hs_kernel(const float* __restrict__ w, unsigned char* __restrict__ dst, int Cout, int Cin) {
  const int N = 3 * Cin, NKC = (Cout + 31) / 32;
  const int total = NKC * 3 * N * 32;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int k = i & 31, n = (i >> 5) % N, sl = i / (32 * N);
    const int dy = sl % 3, kc = sl / 3;
    const int dx = n / Cin, ci = n - dx * Cin, co = 32 * kc + k;
    float v = 0.f;
    if (co < Cout) v = w[((size_t)co * Cin + ci) * 9 + (2 - dy) * 3 + (2 - dx)];
    float hi, lo;
    tc::split_tf32(v, hi, lo);
    const size_t slice = (size_t)sl * 2 * N * 128;
    const uint32_t off = (uint32_t)(n * 128 + ((((k >> 2) ^ (n & 7)) << 4) | ((k & 3) << 2)));
    *reinterpret_cast<float*>(dst + slice + off) = hi;
    *reinterpret_cast<float*>(dst + slice + (size_t)N * 128 + off) = lo;
  }
}

// Mixed-mode forward slices in consumption order [chunk kc][tap]: tf32-hi image (K-major SWIZZLE_128B [Cout rows][32 k]) then
// the CROSS image: row n, 16-byte chunk c = [bf16(w_lo[4c..4c+3]) | bf16(w[4c..4c+3])], same swizzle.
__global__ void pack_mix_kernel(const float* __restrict__ w, unsigned char* __restrict__ dst, int Cout, int Cin) {
  const int NKC = (Cin + 31) / 32;
  const int total = NKC * 9 * Cout * 8;  // one thread per (slice, row, chunk of 4 channels)
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int c = i & 7, n = (i >> 3) % Cout, sl = i / (8 * Cout);
    const int tap = sl % 9, kc = sl / 9;
    float v[4], hi[4], lo[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int ci = 32 * kc + 4 * c + j;
      v[j] = ci < Cin ? w[((size_t)n * Cin + ci) * 9 + tap] : 0.f;
      tc::split_tf32(v[j], hi[j], lo[j]);
    }
    const size_t slice = (size_t)sl * 2 * Cout * 128;
    const uint32_t off = (uint32_t)(n * 128 + ((c ^ (n & 7)) << 4));
    *reinterpret_cast<float4*>(dst + slice + off) = make_float4(hi[0], hi[1], hi[2], hi[3]);
    uint4 x;
    x.x = tc::pack_bf16x2(lo[0], lo[1]), x.y = tc::pack_bf16x2(lo[2], lo[3]);
    x.z = tc::pack_bf16x2(v[0], v[1]), x.w = tc::pack_bf16x2(v[2], v[3]);
    *reinterpret_cast<uint4*>(dst + slice + (size_t)Cout * 128 + off) = x;
  }
}

int tune_hs() {
  static const int v = [] {
    const char* e = getenv("ADVB_P3_HS");
    return e != nullptr ? atoi(e) : 1;
  }();
  return v;
}

}  // namespace

bool conv_p3_bwd_hs(int Cin, int Cout, int KS, bool pool, int W) {
  return tune_hs() != 0 && conv_p3_supported(Cin, Cout, KS, pool, W) && pool && Cin <= 48;
}

int conv_p3_pack_hs(const float* w, unsigned char* wd, int Cout, int Cin, cudaStream_t stream) {
  const int n = ((Cout + 31) / 32) * 3 * 3 * Cin * 32;
  pack_hs_kernel<<<cdiv(n, 256), 256, 0, stream>>>(w, wd, Cout, Cin);
  ADVB_KERNEL_OK("pack_tc_bwd_hs", stream);
  return 0;
}

int conv_p3_pack_mix(const float* w, unsigned char* wf, int Cout, int Cin, cudaStream_t stream) {
  const int total = ((Cin + 31) / 32) * 9 * Cout * 8;
  pack_mix_kernel<<<cdiv(total, 256), 256, 0, stream>>>(w, wf, Cout, Cin);
  ADVB_KERNEL_OK("pack_mix", stream);
  return 0;
}

bool conv_p3_supported(int Cin, int Cout, int KS, bool pool, int W) {
  if (tune_p3() == 0 || g_conv_sched != 0 || KS != 3 || W > 40) return false;
  if (pool) return (Cin == 32 && Cout == 96) || (Cin == 48 && Cout == 128) || (Cin == 32 && Cout == 64);
  return Cin == 64 && Cout == 64;
}

int conv_p3_forward(const ConvFwdArgs& f, const unsigned char* wpack, int passes, cudaStream_t stream) {
  ADVB_CHECK(f.in_pad == 1, "3x3 conv expects a 1-pixel input border");
  P3Args a{};
  a.B = f.B, a.H = f.H, a.W = f.W, a.Ho = f.Ho, a.Wo = f.Wo;
  a.wpack = wpack;
  a.in = f.in, a.out = f.out, a.out_pad = f.out_pad, a.codes = f.codes, a.bias = f.bias;
  a.bn_mean = f.bn_mean, a.bn_invstd = f.bn_invstd;
  a.passes = passes;
  if (f.Cin == 32 && f.Cout == 96 && f.pool) return launch_p3<32, 96, true, false>(a, f.tag, stream);
  if (f.Cin == 48 && f.Cout == 128 && f.pool) return launch_p3<48, 128, true, false>(a, f.tag, stream);
  if (f.Cin == 32 && f.Cout == 64 && f.pool) return launch_p3<32, 64, true, false>(a, f.tag, stream);
  if (f.Cin == 64 && f.Cout == 64 && !f.pool) return launch_p3<64, 64, false, false>(a, f.tag, stream);
  set_error("conv shape has no persistent 3x3 instantiation");
  return 1;
}

int conv_p3_plain_forward(const P3Plain& p, cudaStream_t stream) {
  P3Args a{};
  a.B = p.B, a.H = p.H, a.W = p.W, a.Ho = p.H, a.Wo = p.W;
  a.wpack = p.wpack;
  a.in = p.in, a.out = p.out, a.out_pad = p.out_pad, a.bias = p.bias;
  a.mul_h = p.mul_h, a.mul_scale = p.mul_scale, a.mul_slope = p.mul_slope;
  a.aff_scale = p.aff_scale, a.aff_shift = p.aff_shift, a.act_slope = p.act_slope;
  a.passes = p.passes;
  a.out_ch = p.Cout;
  // (channels in, channels out) as stored: K / N are padded to the next multiple of 8 / to 32 or 64 inside the kernel
  if (p.Cin == 64 && p.Cout == 64 && p.W <= 40) return launch_p3<64, 64, false, false, false, true>(a, p.tag, stream);
  if (p.Cin == 24 && p.Cout == 24 && p.W <= 80) return launch_p3<24, 32, false, false, false, true, 80>(a, p.tag, stream);
  if (p.Cin == 64 && p.Cout == 24 && p.W <= 40) return launch_p3<64, 32, false, false, false, true>(a, p.tag, stream);
  if (p.Cin == 24 && p.Cout == 64 && p.W <= 40) return launch_p3<24, 64, false, false, false, true>(a, p.tag, stream);
  set_error("plain persistent 3x3 conv: no instantiation for this shape");
  return 1;
}

bool conv_p3_plain_supported(int Cin, int Cout, int W) {
  return (Cin == 64 && Cout == 64 && W <= 40) || (Cin == 24 && Cout == 24 && W <= 80) || (Cin == 64 && Cout == 24 && W <= 40) ||
         (Cin == 24 && Cout == 64 && W <= 40);
}

int conv_p3_backward(const ConvBwdArgs& g, const unsigned char* wpack, int passes, cudaStream_t stream) {
  P3Args a{};
  a.B = g.B, a.H = g.H, a.W = g.W, a.Ho = g.Ho, a.Wo = g.Wo;
  a.wpack = wpack;
  a.gout = g.gout, a.codes_in = g.codes, a.gin = g.gin, a.bn_invstd = g.bn_invstd;
  a.passes = passes;
  if (conv_p3_bwd_hs(g.Cin, g.Cout, g.KS, g.pool, g.W)) {  // wpack is then in the HS layout (conv_p3_pack_hs)
    if (g.Cin == 32 && g.Cout == 96) return launch_p3<96, 32, true, true, true>(a, g.tag, stream);
    if (g.Cin == 48 && g.Cout == 128) return launch_p3<128, 48, true, true, true>(a, g.tag, stream);
    if (g.Cin == 32 && g.Cout == 64) return launch_p3<64, 32, true, true, true>(a, g.tag, stream);
  }
  if (g.Cin == 32 && g.Cout == 96 && g.pool) return launch_p3<96, 32, true, true>(a, g.tag, stream);
  if (g.Cin == 48 && g.Cout == 128 && g.pool) return launch_p3<128, 48, true, true>(a, g.tag, stream);
  if (g.Cin == 32 && g.Cout == 64 && g.pool) return launch_p3<64, 32, true, true>(a, g.tag, stream);
  if (g.Cin == 64 && g.Cout == 64 && !g.pool) return launch_p3<64, 64, false, true>(a, g.tag, stream);
  set_error("conv shape has no persistent 3x3 instantiation (backward)");
  return 1;
}

}  // namespace advb
