// LCNN "light" convolution blocks on tcgen05: the first block (1 -> 64, 5x5, as an im2col GEMM with K = 25 taps), its
// backward GEMM, and the 1x1 blocks, as PERSISTENT software-pipelined kernels (sm_100a).
//
// These layers have almost no arithmetic per byte (K <= 64): conv_tc.cu's one-tile-per-CTA version spent its time on
// per-CTA fixed costs and on the serial latency chain  global load -> convert -> MMA -> TMEM load -> store  (7 us per
// 128-pixel tile, profiles/r01_launches_tc.md).  Here a CTA loops over 128-pixel tiles and keeps, per SM:
//   * all weight slices resident in shared memory (one TMA bulk copy per CTA),
//   * the NEXT tile's global loads in flight in registers while the current tile is converted / multiplied / drained,
//   * two TMEM accumulators and a dedicated MMA-issue warp, so the MMAs of tile i overlap the epilogue of tile i-1
//     (TMEM -> registers -> bias / Max-Feature-Map -> shared staging -> pool / BatchNorm -> coalesced stores) run by
//     the 8 worker warps.  (Issuing tcgen05.mma blocks the issuing thread for about as long as the MMAs execute - the
//     queue is shallow - so an issuer that also takes part in the epilogue's barriers serialises the two; measured
//     with the kernel's own clock64 phase counters, ADVB_LIGHT_PROF=1.)
// Numerics are identical to conv_tc.cu (3xTF32 split, fp32 accumulation in TMEM, same epilogue arithmetic).
//
// Tile -> pixel maps (no halo: K spans channels or im2col taps only):
//   FLAT  1x1 blocks: 128 consecutive pixels of the clip's H*W grid;
//   BLK   first block forward: 8 rows x 16 columns, so the 2x2 max-pool stays inside the tile;
//   OVL4  first block backward: 128 consecutive pixels, consecutive tiles overlap by 4 so the horizontal half of
//         col2im (T[y][x][dy] = sum_dx Z[y][x - dx + 2][5 dy + dx]) finds its +-2 neighbours inside the tile.
#include <stdlib.h>

#include "conv.cuh"
#include "tc_common.cuh"

namespace advb {

namespace {

using namespace tc;

constexpr int LW = 256;       // worker threads (load / convert / epilogue)
constexpr int LT = LW + 32;   // + one dedicated MMA-issue warp

__host__ __device__ constexpr int pow2_ceil32(int v) {
  int p = 32;
  while (p < v) p <<= 1;
  return p;
}

struct LArgs {
  int B, H, W, Ho, Wo;
  int tiles_per_clip, n_tiles, tiles_x;
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
  FastDiv dTiles, dTx, dW;  // tiles per clip, column tiles (BLK map), conv grid width
  long long* prof;  // optional (ADVB_LIGHT_PROF=1): per-phase cycle counts of CTA 0, thread 0
};

template <int KTOT, int NOUT, bool POOL, bool BWD, bool IM2COL>
struct LCfg {
  static constexpr int NKC = (KTOT + 31) / 32;
  static constexpr int NSTRIDE = pow2_ceil32(NOUT);
  static constexpr int TMEM_COLS = 2 * NSTRIDE;          // two accumulators (power of two)
  static constexpr int W_BYTES = NKC * 2 * NOUT * 128;   // resident weights: per chunk hi image then lo image
  static constexpr int BAND_BYTES = NKC * 2 * 128 * 128; // per chunk: hi rows then lo rows (128 rows x 128 B)
  static constexpr int CS = BWD ? NOUT : NOUT / 2;       // channels per staged pixel
  static constexpr int SS = CS + 4;
  static constexpr int MAP = IM2COL ? (BWD ? 1 : 2) : 0;  // 0 FLAT, 1 OVL4, 2 BLK
  // FLAT (1x1) kernels, round 2: a tile's operand rows are ONE contiguous span of the NHWC tensor (forward: 128 pixels x KTOT
  // floats; backward: 128 x 32 gradient floats + 128 x 32 code bytes), so they arrive by 1-D TMA bulk copies into a raw ring
  // instead of through the workers' registers one tile ahead (ncu: long_scoreboard was the top stall; the backward's workers
  // spent 1.8 k of 5.7 k cycles per tile just ISSUING their prefetch loads).  Forward: 3-4 slots (block 1: 100 -> 90 us).
  // Backward: the ring must not cost the second CTA per SM (4 slots at one CTA per SM measured no gain), so ONE slot, refilled
  // as soon as the tile in it has been converted - the same look-ahead as the register prefetch without its issue cost.
  static constexpr bool RAW = MAP == 0;
  static constexpr int CH = KTOT / 2;                                     // backward: channels of the stage gradient (C_out / 2)
  static constexpr int RAW_MAIN = BWD ? 128 * CH * 4 : 128 * KTOT * 4;   // gradient / activation rows of one tile
  static constexpr int RAW_SLOT = RAW_MAIN + (BWD ? 128 * CH : 0);       // + the code bytes (backward)
  static constexpr int STAGE_BYTES = RAW ? 0 : 128 * SS * 4 + 128 * 8;
  static constexpr int FIXED_BYTES = W_BYTES + BAND_BYTES + STAGE_BYTES + 1024;
  // as many slots as keep the CTAs-per-SM count the kernel had without the ring (2 where it was 2), at most 4, at least 2
  static constexpr int RAW_BUDGET = (FIXED_BYTES + 2 * RAW_SLOT + 1024 <= 113 * 1024 ? 113 * 1024 : 226 * 1024) - 1024 - FIXED_BYTES;
  static constexpr bool BWD_ONE = BWD && FIXED_BYTES + RAW_SLOT + 1024 <= 113 * 1024;  // one slot keeps two CTAs per SM
  static constexpr int RAW_SLOTS = !RAW ? 0 : BWD_ONE ? 1 : (RAW_BUDGET / RAW_SLOT >= 4 ? 4 : (RAW_BUDGET / RAW_SLOT >= 3 ? 3 : 2));
  static constexpr size_t SMEM = (size_t)FIXED_BYTES + (size_t)RAW_SLOTS * RAW_SLOT;
  static constexpr int STEP = MAP == 1 ? 124 : 128;       // new pixels per tile (FLAT / OVL4)
#ifndef ADVB_LIGHT_CTAS3
#define ADVB_LIGHT_CTAS3 0
#endif
  // CTAs per SM the kernel is compiled for (register cap).  -DADVB_LIGHT_CTAS3=1 asks for 3 where shared memory and TMEM
  // allow it (first block forward, 32 -> 64 1x1): measured slower, 13.5 vs 12.1 ms and 4.6 vs 4.2 ms per PGD-40 call (72
  // registers per thread: spills, and the per-tile latency chain does not shorten), so the default stays 2
  static constexpr int CTAS = SMEM > 113 * 1024 ? 1 : ((ADVB_LIGHT_CTAS3 && 3 * (SMEM + 1024) <= 227 * 1024 && 3 * TMEM_COLS <= 512) ? 3 : 2);
};

// pixel of tile-local row m: (y, x) on the conv grid; false when the row is padding
template <int MAP>
__device__ __forceinline__ bool tile_pixel(const LArgs& a, int tl, int m, int Heff, int& y, int& x) {
  if (MAP == 2) {
    const int ty = fdiv(tl, a.dTx), tx = tl - ty * a.tiles_x;
    y = 8 * ty + (m >> 4);
    x = 16 * tx + (m & 15);
    return y < Heff && x < a.W;
  }
  const int px = (MAP == 1 ? 124 * tl - 2 : 128 * tl) + m;
  const bool ok = px >= 0 && px < Heff * a.W;
  const int p = ok ? px : 0;
  y = fdiv(p, a.dW);
  x = p - y * a.W;
  return ok;
}

template <int KTOT, int NOUT, bool POOL, bool BWD, bool IM2COL>
__global__ void __launch_bounds__(LT, LCfg<KTOT, NOUT, POOL, BWD, IM2COL>::CTAS)
conv_light_kernel(LArgs a) {
  using Cfg = LCfg<KTOT, NOUT, POOL, BWD, IM2COL>;
  constexpr int NKC = Cfg::NKC, CS = Cfg::CS, SS = Cfg::SS, MAP = Cfg::MAP;
  constexpr uint32_t IDESC = idesc_tf32(128, NOUT);
  static_assert(!BWD || KTOT == 64 || Cfg::RAW, "the register-prefetch backward assumes Cout = 64 (one 32-channel chunk per MFM half)");
  static_assert(!(Cfg::RAW && BWD && POOL), "the raw-ring backward assumes an un-pooled 1x1 block (gradient and codes on the conv grid)");
  static_assert(!(IM2COL && !BWD) || KTOT == 32, "first-block forward: 25 taps padded to one 32-wide chunk");

  extern __shared__ unsigned char smem_raw[];
  unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  unsigned char* w_s = base;                       // [NKC][hi NOUT x 128 B | lo NOUT x 128 B]
  unsigned char* band = w_s + Cfg::W_BYTES;        // [NKC][hi 128 x 128 B | lo 128 x 128 B]   (W_BYTES is a multiple of 1024)
  float* stage = reinterpret_cast<float*>(band + Cfg::BAND_BYTES);
  unsigned long long* flags = reinterpret_cast<unsigned long long*>(stage + 128 * SS);  // (unused when RAW: no staging tile)
  unsigned char* raw = band + Cfg::BAND_BYTES + Cfg::STAGE_BYTES;  // RAW: [RAW_SLOTS][RAW_SLOT] ring of unconverted tiles
  constexpr bool RAW = Cfg::RAW;
  constexpr int RSL = Cfg::RAW_SLOTS > 0 ? Cfg::RAW_SLOTS : 1;
  __shared__ uint64_t bar_w, bar_mma[2], bar_full, bar_raw[RSL];
  __shared__ uint32_t tmem_base_s;
  __shared__ __align__(16) float s_bias[BWD ? 4 : NOUT];
  __shared__ __align__(16) float s_bnm[BWD ? 4 : NOUT / 2], s_bni[BWD ? 4 : NOUT / 2];  // BatchNorm(eval) mean / inverse std (0 / 1 without BN)

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int Heff = (!BWD && POOL) ? 2 * a.Ho : a.H;

  if (tid == 0) {
    mbar_init(&bar_w, 1);
    mbar_init(&bar_mma[0], 1);
    mbar_init(&bar_mma[1], 1);
    mbar_init(&bar_full, LW / 32);
    for (int i = 0; i < RSL; ++i) mbar_init(&bar_raw[i], 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc<Cfg::TMEM_COLS>(&tmem_base_s);
  if (!BWD) {
    for (int i = tid; i < NOUT; i += LT) s_bias[i] = __ldg(a.bias + i);
    for (int i = tid; i < NOUT / 2; i += LT) {
      s_bnm[i] = a.bn_mean != nullptr ? __ldg(a.bn_mean + i) : 0.f;
      s_bni[i] = a.bn_mean != nullptr ? __ldg(a.bn_invstd + i) : 1.f;
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  if (tid == 0) {
    mbar_expect_tx(&bar_w, Cfg::W_BYTES);
    bulk_g2s(w_s, a.wpack, Cfg::W_BYTES, &bar_w);
  }

  if (warp == LW / 32) {
    // ================= MMA-issue warp =================
    const bool leader = elect_one();
    // RAW: this warp is also the producer of the raw ring.  Slot it % RSL holds tile `it`; it is refilled with tile it + RSL as soon
    // as every worker warp has converted tile `it` out of it (= bar_full of tile `it`, which this warp waits for anyway).
    auto raw_load = [&](int tile, int slot) {
      const int b = fdiv(tile, a.dTiles), tl = tile - b * a.tiles_per_clip;
      const int npx = min(128, a.H * a.W - 128 * tl);
      unsigned char* dst = raw + (size_t)slot * Cfg::RAW_SLOT;
      const size_t px0 = (size_t)b * a.H * a.W + (size_t)128 * tl;
      if (BWD) {
        mbar_expect_tx(&bar_raw[slot], (uint32_t)(npx * Cfg::CH * 5));
        bulk_g2s(dst, a.gout + px0 * Cfg::CH, (uint32_t)(npx * Cfg::CH * 4), &bar_raw[slot]);
        bulk_g2s(dst + Cfg::RAW_MAIN, a.codes_in + px0 * Cfg::CH, (uint32_t)(npx * Cfg::CH), &bar_raw[slot]);
      } else {
        mbar_expect_tx(&bar_raw[slot], (uint32_t)(npx * KTOT * 4));
        bulk_g2s(dst, a.in + px0 * KTOT, (uint32_t)(npx * KTOT * 4), &bar_raw[slot]);
      }
    };
    if (RAW && leader) {
      int t = blockIdx.x;
      for (int i = 0; i < RSL && t < a.n_tiles; ++i, t += gridDim.x) raw_load(t, i);
    }
    mbar_wait(&bar_w, 0u);
    int it = 0;
    for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x, ++it) {
      const int buf = it & 1;
      mbar_wait(&bar_full, (uint32_t)(it & 1));  // band of tile `it` is staged (and accumulator `buf` has been drained)
      tc_fence_after();
      if (RAW && leader) {
        const long long nt = (long long)tile + (long long)RSL * gridDim.x;
        if (nt < a.n_tiles) raw_load((int)nt, it % RSL);
      }
      const uint32_t dcol = tmem + buf * Cfg::NSTRIDE;
#pragma unroll
      for (int kc = 0; kc < NKC; ++kc) {
        const int kvalid = (KTOT - 32 * kc) >= 32 ? 32 : (KTOT - 32 * kc);
        const uint32_t a_hi = smem_u32(band + (size_t)kc * 2 * 128 * 128), a_lo = a_hi + 128 * 128;
        const uint32_t w_hi = smem_u32(w_s + (size_t)kc * 2 * NOUT * 128), w_lo = w_hi + NOUT * 128;
#pragma unroll 1
        for (int ks = 0; ks < kvalid / 8; ++ks) {
          const uint64_t ah = desc_sw128(a_hi + ks * 32), al = desc_sw128(a_lo + ks * 32);
          const uint64_t bh = desc_sw128(w_hi + ks * 32), bl = desc_sw128(w_lo + ks * 32);
          if (leader) {
            mma_tf32(dcol, ah, bh, IDESC, (kc > 0 || ks > 0) ? 1u : 0u);
            if (a.passes == 3) {
              mma_tf32(dcol, ah, bl, IDESC, 1u);
              mma_tf32(dcol, al, bh, IDESC, 1u);
            }
          }
        }
      }
      if (leader) mma_commit(&bar_mma[buf]);
      __syncwarp();
    }
  } else {
  // ================= worker warps =================

  // ---- per-thread item geometry: 4 band rows (m = tid/8 + 32 u), one 16-byte channel group c4 of every chunk ----
  const int c4 = tid & 7, m0 = tid >> 3;
  int toff[4];
  bool tok[4];
  if (IM2COL && !BWD) {
    const int Wi = a.W + 4;
#pragma unroll
    for (int u4 = 0; u4 < 4; ++u4) {
      const int tap = 4 * c4 + u4;
      tok[u4] = tap < 25;
      const int dy = tap / 5, dx = tap - dy * 5;
      toff[u4] = tok[u4] ? dy * Wi + dx : 0;
    }
  }
  float4 sc = make_float4(1.f, 1.f, 1.f, 1.f);
  if (BWD && a.bn_invstd != nullptr) sc = __ldg(reinterpret_cast<const float4*>(a.bn_invstd + 4 * c4));
  float4 scr[(BWD && Cfg::RAW) ? NKC : 1];  // RAW backward: BatchNorm(eval) scale of this thread's 4 channels in every chunk
  if (BWD && Cfg::RAW) {
#pragma unroll
    for (int kc = 0; kc < NKC; ++kc) {
      const int k0 = 32 * kc + 4 * c4, c = k0 >= Cfg::CH ? k0 - Cfg::CH : k0;
      scr[kc] = make_float4(1.f, 1.f, 1.f, 1.f);
      if (a.bn_invstd != nullptr && k0 < KTOT) scr[kc] = __ldg(reinterpret_cast<const float4*>(a.bn_invstd + c));
    }
  }

  // registers holding the next tile's raw loads
  float4 rv[4][BWD ? 1 : NKC];
  uchar4 rc[4];
  unsigned rwant = 0;  // BWD: per item the pool position code it corresponds to (8 bits each), 0xff = padding row
  unsigned rok = 0;    // forward: bit u = item u is a real pixel (loads are kept RAW until convert_store, so that they
                       // stay in flight across the epilogue: any use of a loaded value would stall the in-order warp)

  auto issue_loads = [&](int tile) {
    const int b = fdiv(tile, a.dTiles), tl = tile - b * a.tiles_per_clip;
    rwant = 0;
    rok = 0;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      int y, x;
      bool ok = tile_pixel<MAP>(a, tl, m0 + 32 * u, Heff, y, x);
      if (BWD) {
        const int py = POOL ? (y >> 1) : y, px = POOL ? (x >> 1) : x;
        ok = ok && py < a.Ho && px < a.Wo;
        const size_t o = ok ? (((size_t)b * a.Ho + py) * a.Wo + px) * 32 + 4 * c4 : 0;
        rv[u][0] = __ldg(reinterpret_cast<const float4*>(a.gout + o));
        rc[u] = __ldg(reinterpret_cast<const uchar4*>(a.codes_in + o));
        const unsigned w = ok ? (POOL ? (unsigned)(((y & 1) << 1) | (x & 1)) : 0u) : 0xffu;
        rwant |= w << (8 * u);
      } else if (IM2COL) {
        const float* img = a.in + ((size_t)b * (a.H + 4) + (ok ? y : 0)) * (a.W + 4) + (ok ? x : 0);
        rv[u][0].x = __ldg(img + toff[0]);
        rv[u][0].y = __ldg(img + toff[1]);
        rv[u][0].z = __ldg(img + toff[2]);
        rv[u][0].w = __ldg(img + toff[3]);
        rok |= (ok ? 1u : 0u) << u;
      } else {
        const size_t pix = ((size_t)b * a.H + (ok ? y : 0)) * a.W + (ok ? x : 0);
#pragma unroll
        for (int kc = 0; kc < NKC; ++kc) {
          const int ch = 32 * kc + 4 * c4;
          const bool okc = ok && ch < KTOT;
          rv[u][kc] = __ldg(reinterpret_cast<const float4*>(a.in + (okc ? pix * KTOT + ch : 0)));
        }
        rok |= (ok ? 1u : 0u) << u;
      }
    }
  };

  auto store_item = [&](int kc, int m, float4 v) {
    float4 hi, lo;
    split_tf32(v.x, hi.x, lo.x);
    split_tf32(v.y, hi.y, lo.y);
    split_tf32(v.z, hi.z, lo.z);
    split_tf32(v.w, hi.w, lo.w);
    unsigned char* hb = band + (size_t)kc * 2 * 128 * 128;
    const uint32_t off = sw128_chunk(m, c4);
    *reinterpret_cast<float4*>(hb + off) = hi;
    *reinterpret_cast<float4*>(hb + 128 * 128 + off) = lo;
  };

  // RAW: tile `it` out of its ring slot (rows >= npx of a clip's last tile are zero)
  auto convert_store_raw = [&](int tile, int it) {
    const int slot = it % RSL;
    mbar_wait(&bar_raw[slot], (uint32_t)((it / RSL) & 1));
    const int b = fdiv(tile, a.dTiles), tl = tile - b * a.tiles_per_clip;
    const int npx = min(128, a.H * a.W - 128 * tl);
    const unsigned char* src = raw + (size_t)slot * Cfg::RAW_SLOT;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int m = m0 + 32 * u;
      const bool ok = m < npx;
      if (BWD) {
        // K index k = 32 kc + 4 c4 of the expanded gradient = (MFM half k / CH, stage channel k % CH); a 1x1 block does not pool,
        // so the stored code of a channel is just its winning half << 2
        constexpr int CH = Cfg::CH;
        float4 g0 = make_float4(0.f, 0.f, 0.f, 0.f);
        uchar4 cd0 = make_uchar4(0xff, 0xff, 0xff, 0xff);
        if (CH == 32 && ok) {  // both chunks are the two halves of the SAME 32 channels: one load serves them
          g0 = *reinterpret_cast<const float4*>(src + ((size_t)m * CH + 4 * c4) * 4);
          cd0 = *reinterpret_cast<const uchar4*>(src + Cfg::RAW_MAIN + (size_t)m * CH + 4 * c4);
        }
#pragma unroll
        for (int kc = 0; kc < NKC; ++kc) {
          const int k0 = 32 * kc + 4 * c4;
          const int half = k0 >= CH ? 1 : 0, c = k0 - half * CH;
          float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
          if (ok && k0 < KTOT) {
            const float4 g = CH == 32 ? g0 : *reinterpret_cast<const float4*>(src + ((size_t)m * CH + c) * 4);
            const uchar4 cd = CH == 32 ? cd0 : *reinterpret_cast<const uchar4*>(src + Cfg::RAW_MAIN + (size_t)m * CH + c);
            const unsigned want = (unsigned)(half << 2);
            v.x = cd.x == want ? g.x * scr[kc].x : 0.f;
            v.y = cd.y == want ? g.y * scr[kc].y : 0.f;
            v.z = cd.z == want ? g.z * scr[kc].z : 0.f;
            v.w = cd.w == want ? g.w * scr[kc].w : 0.f;
          }
          store_item(kc, m, v);
        }
      } else {
#pragma unroll
        for (int kc = 0; kc < NKC; ++kc) {
          const int ch = 32 * kc + 4 * c4;
          float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
          if (ok && ch < KTOT) v = *reinterpret_cast<const float4*>(src + ((size_t)m * KTOT + ch) * 4);
          store_item(kc, m, v);
        }
      }
    }
  };

  auto convert_store = [&]() {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int m = m0 + 32 * u;
      if (BWD) {
        const unsigned w = (rwant >> (8 * u)) & 0xffu;
        const float4 g = make_float4(rv[u][0].x * sc.x, rv[u][0].y * sc.y, rv[u][0].z * sc.z, rv[u][0].w * sc.w);
#pragma unroll
        for (int kc = 0; kc < NKC; ++kc) {  // chunk kc = MFM half kc of the same 32 channels
          const unsigned want = w == 0xffu ? 0xffu : (w | (unsigned)(kc << 2));
          float4 v;
          v.x = rc[u].x == want ? g.x : 0.f;
          v.y = rc[u].y == want ? g.y : 0.f;
          v.z = rc[u].z == want ? g.z : 0.f;
          v.w = rc[u].w == want ? g.w : 0.f;
          store_item(kc, m, v);
        }
      } else if (IM2COL) {
        const bool ok = (rok >> u) & 1u;
        store_item(0, m, make_float4((ok && tok[0]) ? rv[u][0].x : 0.f, (ok && tok[1]) ? rv[u][0].y : 0.f,
                                     (ok && tok[2]) ? rv[u][0].z : 0.f, (ok && tok[3]) ? rv[u][0].w : 0.f));
      } else {
        const bool ok = (rok >> u) & 1u;
#pragma unroll
        for (int kc = 0; kc < NKC; ++kc) {
          const bool okc = ok && (32 * kc + 4 * c4) < KTOT;
          store_item(kc, m, okc ? rv[u][kc] : make_float4(0.f, 0.f, 0.f, 0.f));
        }
      }
    }
  };

  // ---- epilogue of one tile: TMEM accumulator `buf` -> staging -> global ----
  auto epilogue = [&](int tile, int buf) {
    const int b = fdiv(tile, a.dTiles), tl = tile - b * a.tiles_per_clip;
    if (MAP == 0) {
      // ---- 1x1 blocks (FLAT map), round 2: straight from the tcgen05.ld registers to global memory.  Thread = TMEM lane =
      // pixel holds every channel of its pixel, so the Max-Feature-Map pair (c, c + C/2), BatchNorm and the code bytes need
      // no staging tile, no second pass and no barrier; a pixel's channels are one contiguous row of the NHWC output, and a
      // warp's 32 pixels are consecutive rows.  (The staged version: TMEM -> shared -> barrier -> coalesced copy with its own
      // index math, ~2x the instructions; these layers moved 35 % of the HBM peak.)
      const int wq = warp & 3, hsel = warp >> 2;
      const int r = wq * 32 + lane;
      const int px = 128 * tl + r;
      const bool valid = px < a.H * a.W;
      const uint32_t taddr = tmem + ((uint32_t)(wq * 32) << 16) + buf * Cfg::NSTRIDE;
      if (BWD) {
        float* dst = a.gin + ((size_t)b * a.H * a.W + (valid ? px : 0)) * NOUT;
#pragma unroll 1
        for (int c0 = hsel * 16; c0 < NOUT; c0 += 32) {
          uint32_t v[16];
          tmem_ld16_issue(taddr + c0, v);
          tmem_ld_wait();
          if (valid) {
#pragma unroll
            for (int j = 0; j < 16; j += 4)
              *reinterpret_cast<float4*>(dst + c0 + j) = make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]),
                                                                     __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
          }
        }
      } else {
        const int p = valid ? px : 0;
        const int y = fdiv(p, a.dW), x = p - y * a.W;
        const int Hop = a.Ho + 2 * a.out_pad, Wop = a.Wo + 2 * a.out_pad;
        float* dst = a.out + (((size_t)b * Hop + y + a.out_pad) * Wop + x + a.out_pad) * CS;
        unsigned char* cdst = a.codes + ((size_t)b * a.H * a.W + p) * CS;
#pragma unroll 1
        for (int c0 = hsel * 16; c0 < CS; c0 += 32) {
          uint32_t lo[16], hi[16];
          tmem_ld16_issue(taddr + c0, lo);
          tmem_ld16_issue(taddr + CS + c0, hi);
          tmem_ld_wait();
          float o[16];
          unsigned cw[4] = {0u, 0u, 0u, 0u};
#pragma unroll
          for (int j4 = 0; j4 < 16; j4 += 4) {  // per-channel vectors as LDS.128 (they were 64 scalar LDS per 16 channels)
            const float4 bl4 = *reinterpret_cast<const float4*>(s_bias + c0 + j4);
            const float4 bh4 = *reinterpret_cast<const float4*>(s_bias + CS + c0 + j4);
            const float4 mn4 = *reinterpret_cast<const float4*>(s_bnm + c0 + j4);
            const float4 is4 = *reinterpret_cast<const float4*>(s_bni + c0 + j4);
            const float bl[4] = {bl4.x, bl4.y, bl4.z, bl4.w}, bh[4] = {bh4.x, bh4.y, bh4.z, bh4.w};
            const float mn[4] = {mn4.x, mn4.y, mn4.z, mn4.w}, is[4] = {is4.x, is4.y, is4.z, is4.w};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const int j = j4 + k;
              const float l = __uint_as_float(lo[j]) + bl[k];
              const float h = __uint_as_float(hi[j]) + bh[k];
              const bool sel = h > l;
              o[j] = ((sel ? h : l) - mn[k]) * is[k];
              cw[j >> 2] |= sel ? (4u << (8 * k)) : 0u;  // code byte: MFM half << 2 (no pool position)
            }
          }
          if (valid) {
#pragma unroll
            for (int j = 0; j < 16; j += 4) *reinterpret_cast<float4*>(dst + c0 + j) = make_float4(o[j], o[j + 1], o[j + 2], o[j + 3]);
            *reinterpret_cast<uint4*>(cdst + c0) = make_uint4(cw[0], cw[1], cw[2], cw[3]);
          }
        }
      }
      tc_fence_before();
      return;
    }
    asm volatile("bar.sync 1, %0;" ::"n"(LW) : "memory");  // every worker has finished reading the previous staging tile
    {
      // TMEM -> registers -> staging.  Warps w and w+4 share TMEM lane quadrant w%4 and split the columns.
      const int wq = warp & 3, hsel = warp >> 2;
      const int r = wq * 32 + lane;
      const uint32_t taddr = tmem + ((uint32_t)(wq * 32) << 16) + buf * Cfg::NSTRIDE;
      float* srow = stage + (size_t)r * SS;
      if (BWD) {
#pragma unroll 1
        for (int c0 = hsel * 16; c0 < NOUT; c0 += 32) {
          uint32_t v[16];
          tmem_ld16_issue(taddr + c0, v);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; j += 4)
            *reinterpret_cast<float4*>(srow + c0 + j) = make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]),
                                                                   __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
        }
      } else {
        unsigned fl = 0u;  // this warp's 16-channel groups: bit (c0 + j) of the pixel's MFM flags
#pragma unroll 1
        for (int c0 = hsel * 16; c0 < CS; c0 += 32) {
          uint32_t lo[16], hi[16];
          tmem_ld16_issue(taddr + c0, lo);
          tmem_ld16_issue(taddr + CS + c0, hi);
          tmem_ld_wait();
          float m4[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const float l = __uint_as_float(lo[j]) + s_bias[c0 + j];
            const float h = __uint_as_float(hi[j]) + s_bias[CS + c0 + j];
            const bool sel = h > l;
            m4[j] = sel ? h : l;
            fl |= (sel ? 1u : 0u) << j;
          }
#pragma unroll
          for (int j = 0; j < 16; j += 4)
            *reinterpret_cast<float4*>(srow + c0 + j) = make_float4(m4[j], m4[j + 1], m4[j + 2], m4[j + 3]);
          // flags as 16-bit fields: field index = c0 / 16
          reinterpret_cast<unsigned short*>(flags + r)[c0 >> 4] = (unsigned short)fl;
          fl = 0u;
        }
      }
    }
    tc_fence_before();
    asm volatile("bar.sync 1, %0;" ::"n"(LW) : "memory");
    constexpr int C4 = CS / 4;
    if (BWD && IM2COL) {
      // T[y][x][dy] for the 124 interior pixels of the tile
      for (int i = tid; i < 124 * 5; i += LW) {
        const int dy = i % 5, ml = 2 + i / 5;
        const int px = 124 * tl - 2 + ml;
        if (px >= a.H * a.W) continue;
        const int y = fdiv(px, a.dW), x = px - y * a.W;
        float acc = 0.f;
#pragma unroll
        for (int dx = 0; dx < 5; ++dx) {
          const int xs = x - dx + 2;
          if (xs >= 0 && xs < a.W) acc += stage[(size_t)(ml - dx + 2) * SS + 5 * dy + dx];
        }
        a.gin[((size_t)b * a.H * a.W + px) * 5 + dy] = acc;
      }
    } else if (BWD) {
      const int npx = min(128, a.H * a.W - 128 * tl);
      for (int i = tid; i < npx * C4; i += LW) {
        const int c4i = i % C4, ml = i / C4;
        const float4 v = *reinterpret_cast<const float4*>(stage + (size_t)ml * SS + 4 * c4i);
        *reinterpret_cast<float4*>(a.gin + ((size_t)b * a.H * a.W + 128 * tl + ml) * NOUT + 4 * c4i) = v;
      }
    } else if (POOL) {
      // 8 x 16 pixel tile -> 4 x 8 pooled cells
      const int ty = fdiv(tl, a.dTx), tx = tl - ty * a.tiles_x;
      const int Hop = a.Ho + 2 * a.out_pad, Wop = a.Wo + 2 * a.out_pad;
      for (int i = tid; i < 32 * C4; i += LW) {
        const int c4i = i % C4, cell = i / C4;
        const int cy = cell >> 3, cx = cell & 7;
        const int oy = 4 * ty + cy, ox = 8 * tx + cx;
        if (oy >= a.Ho || ox >= a.Wo) continue;
        const int c = 4 * c4i;
        const int r00 = (2 * cy) * 16 + 2 * cx;
        const int rr[4] = {r00, r00 + 1, r00 + 16, r00 + 17};
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
        float4 v = make_float4(best[0], best[1], best[2], best[3]);
        if (a.bn_mean != nullptr) {
          const float4 is = __ldg(reinterpret_cast<const float4*>(a.bn_invstd + c));
          v.x = (v.x - __ldg(a.bn_mean + c)) * is.x;
          v.y = (v.y - __ldg(a.bn_mean + c + 1)) * is.y;
          v.z = (v.z - __ldg(a.bn_mean + c + 2)) * is.z;
          v.w = (v.w - __ldg(a.bn_mean + c + 3)) * is.w;
        }
        *reinterpret_cast<float4*>(a.out + (((size_t)b * Hop + oy + a.out_pad) * Wop + ox + a.out_pad) * CS + c) = v;
        *reinterpret_cast<uchar4*>(a.codes + (((size_t)b * a.Ho + oy) * a.Wo + ox) * CS + c) =
            make_uchar4((unsigned char)code[0], (unsigned char)code[1], (unsigned char)code[2], (unsigned char)code[3]);
      }
    } else {
      const int Hop = a.Ho + 2 * a.out_pad, Wop = a.Wo + 2 * a.out_pad;
      const int npx = min(128, a.H * a.W - 128 * tl);
      for (int i = tid; i < npx * C4; i += LW) {
        const int c4i = i % C4, ml = i / C4;
        const int c = 4 * c4i;
        const int px = 128 * tl + ml;
        const int y = fdiv(px, a.dW), x = px - y * a.W;
        float4 v = *reinterpret_cast<const float4*>(stage + (size_t)ml * SS + c);
        const unsigned f = (unsigned)(flags[ml] >> c) & 15u;
        if (a.bn_mean != nullptr) {
          const float4 is = __ldg(reinterpret_cast<const float4*>(a.bn_invstd + c));
          v.x = (v.x - __ldg(a.bn_mean + c)) * is.x;
          v.y = (v.y - __ldg(a.bn_mean + c + 1)) * is.y;
          v.z = (v.z - __ldg(a.bn_mean + c + 2)) * is.z;
          v.w = (v.w - __ldg(a.bn_mean + c + 3)) * is.w;
        }
        *reinterpret_cast<float4*>(a.out + (((size_t)b * Hop + y + a.out_pad) * Wop + x + a.out_pad) * CS + c) = v;
        *reinterpret_cast<uchar4*>(a.codes + ((size_t)b * a.H * a.W + px) * CS + c) =
            make_uchar4((unsigned char)((f & 1u) << 2), (unsigned char)(((f >> 1) & 1u) << 2),
                        (unsigned char)(((f >> 2) & 1u) << 2), (unsigned char)(((f >> 3) & 1u) << 2));
      }
    }
  };

  // ---- persistent loop (workers) ----
  int tile = blockIdx.x;
  if (!RAW && tile < a.n_tiles) issue_loads(tile);
  int it = 0, prev_tile = -1;
  long long pc[4] = {0, 0, 0, 0};
  const bool prof = a.prof != nullptr && blockIdx.x == 0 && tid == 0;
#define LPROF(k)                      \
  if (prof) {                         \
    const long long now_ = clock64(); \
    pc[k] += now_ - t_last;           \
    t_last = now_;                    \
  }
  long long t_last = clock64();
  for (; tile < a.n_tiles; tile += gridDim.x, ++it) {
    const int buf = it & 1;
    if (it > 0) {  // MMAs of the previous tile have finished reading the band (and its accumulator is complete)
      mbar_wait(&bar_mma[buf ^ 1], (uint32_t)(((it - 1) >> 1) & 1));
      tc_fence_after();
    }
    LPROF(0)
    if (RAW) convert_store_raw(tile, it);
    else convert_store();
    fence_proxy_async();
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(&bar_full);  // 8 worker warps -> the MMA warp may issue tile `it`
    LPROF(1)
    const int next = tile + gridDim.x;
    if (!RAW && next < a.n_tiles) issue_loads(next);
    LPROF(2)
    if (it > 0) epilogue(prev_tile, buf ^ 1);  // overlaps the MMAs of tile `it`
    LPROF(3)
    prev_tile = tile;
  }
  if (prof) {
    for (int k = 0; k < 4; ++k) a.prof[k] = pc[k];
    a.prof[6] = it;
  }
  if (it > 0) {
    const int buf = (it - 1) & 1;
    mbar_wait(&bar_mma[buf], (uint32_t)(((it - 1) >> 1) & 1));
    tc_fence_after();
    epilogue(prev_tile, buf);
  }
  }  // worker warps
  // one barrier instruction for every warp of the CTA (compute-sanitizer synccheck reports role branches that each end in
  // their own __syncthreads as divergent barriers)
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<Cfg::TMEM_COLS>(tmem);
}

int tune_light_persistent() {
  static const int v = [] {
    const char* e = getenv("ADVB_LIGHT_PERSISTENT");
    return e != nullptr ? atoi(e) : 1;
  }();
  return v;
}

template <int KTOT, int NOUT, bool POOL, bool BWD, bool IM2COL>
int launch_light(LArgs a, const char* tag, cudaStream_t stream) {
  using Cfg = LCfg<KTOT, NOUT, POOL, BWD, IM2COL>;
  static_assert(Cfg::SMEM <= 227 * 1024, "light conv kernel does not fit shared memory");
  static_assert(Cfg::W_BYTES % 1024 == 0, "weight image must keep the band 1024-byte aligned");
  const int Heff = (!BWD && POOL) ? 2 * a.Ho : a.H;
  if (Cfg::MAP == 2) {
    a.tiles_x = cdiv(a.W, 16);
    a.tiles_per_clip = cdiv(Heff, 8) * a.tiles_x;
  } else {
    a.tiles_x = 1;
    a.tiles_per_clip = cdiv(Heff * a.W, Cfg::STEP);
  }
  a.n_tiles = a.B * a.tiles_per_clip;
  a.dTiles = make_fastdiv(a.tiles_per_clip);
  a.dTx = make_fastdiv(a.tiles_x);
  a.dW = make_fastdiv(a.W);
  auto kern = conv_light_kernel<KTOT, NOUT, POOL, BWD, IM2COL>;
  ADVB_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM));
  ADVB_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
  // CTAs per SM by shared memory (227 KB, + 1 KB the driver reserves per CTA), TMEM columns (512) and registers
  // (__launch_bounds__ asks for 2).  A larger grid than the SMs can hold at once is harmless: tiles are strided by
  // gridDim.x and the surplus CTAs simply start later.
  int per_sm = (int)((227 * 1024) / (Cfg::SMEM + 1024));
  if (per_sm > 512 / Cfg::TMEM_COLS) per_sm = 512 / Cfg::TMEM_COLS;
  if (per_sm > Cfg::CTAS) per_sm = Cfg::CTAS;
  if (per_sm < 1) per_sm = 1;
  int n_sm = 148;
  int dev = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
  int grid = n_sm * per_sm;
  if (grid > a.n_tiles) grid = a.n_tiles;
  static long long* prof_buf = nullptr;
  static const bool want_prof = getenv("ADVB_LIGHT_PROF") != nullptr;
  if (want_prof && prof_buf == nullptr) cudaMalloc(reinterpret_cast<void**>(&prof_buf), 8 * sizeof(long long));
  a.prof = want_prof ? prof_buf : nullptr;
  kern<<<grid, LT, Cfg::SMEM, stream>>>(a);
  ADVB_KERNEL_OK(tag, stream);
  if (want_prof) {
    long long hp[8];
    cudaStreamSynchronize(stream);
    cudaMemcpy(hp, prof_buf, sizeof(hp), cudaMemcpyDeviceToHost);
    fprintf(stderr, "[light %s] grid %d per_sm %d tiles/cta %lld | cycles/tile: wait_mma %lld convert %lld issue_loads %lld epilogue %lld\n",
            tag, grid, per_sm, hp[6], hp[0] / hp[6], hp[1] / hp[6], hp[2] / hp[6], hp[3] / hp[6]);
  }
  return 0;
}

}  // namespace

bool conv_light_supported(int Cin, int Cout, int KS, bool pool, bool bwd) {
  if (tune_light_persistent() == 0 || g_conv_sched != 0) return false;
  if (KS == 5) return Cin == 1 && Cout == 64 && pool;
  if (KS != 1 || pool) return false;
  if (!bwd) return (Cin == 32 && Cout == 64) || (Cin == 48 && Cout == 96) || (Cin == 64 && Cout == 128);
  return (Cout == 64 && Cin == 32) || (Cout == 96 && Cin == 48);
}

int conv_light_forward(const ConvFwdArgs& f, const unsigned char* wpack, int passes, cudaStream_t stream) {
  LArgs a{};
  a.B = f.B, a.H = f.H, a.W = f.W, a.Ho = f.Ho, a.Wo = f.Wo;
  a.wpack = wpack;
  a.in = f.in, a.out = f.out, a.out_pad = f.out_pad, a.codes = f.codes, a.bias = f.bias;
  a.bn_mean = f.bn_mean, a.bn_invstd = f.bn_invstd;
  a.passes = passes;
  if (f.KS == 5) return launch_light<32, 64, true, false, true>(a, f.tag, stream);
  ADVB_CHECK(f.in_pad == 0, "1x1 light conv expects a border-free input");
  if (f.Cin == 32 && f.Cout == 64) return launch_light<32, 64, false, false, false>(a, f.tag, stream);
  if (f.Cin == 48 && f.Cout == 96) return launch_light<48, 96, false, false, false>(a, f.tag, stream);
  if (f.Cin == 64 && f.Cout == 128) return launch_light<64, 128, false, false, false>(a, f.tag, stream);
  set_error("conv shape has no light instantiation");
  return 1;
}

int conv_light_backward(const ConvBwdArgs& g, const unsigned char* wpack, int passes, cudaStream_t stream) {
  LArgs a{};
  a.B = g.B, a.H = g.H, a.W = g.W, a.Ho = g.Ho, a.Wo = g.Wo;
  a.wpack = wpack;
  a.gout = g.gout, a.codes_in = g.codes, a.gin = g.gin, a.bn_invstd = g.bn_invstd;
  a.passes = passes;
  if (g.KS == 1 && g.Cout == 64 && g.Cin == 32 && !g.pool) return launch_light<64, 32, false, true, false>(a, g.tag, stream);
  if (g.KS == 1 && g.Cout == 96 && g.Cin == 48 && !g.pool) return launch_light<96, 48, false, true, false>(a, g.tag, stream);
  set_error("conv shape has no light instantiation (backward)");
  return 1;
}

int conv0_light_backward_gemm(const float* gout, const unsigned char* codes, const unsigned char* wpack, float* T, int B,
                              int H, int W, int Ho, int Wo, int passes, cudaStream_t stream) {
  LArgs a{};
  a.B = B, a.H = H, a.W = W, a.Ho = Ho, a.Wo = Wo;
  a.wpack = wpack;
  a.gout = gout, a.codes_in = codes, a.gin = T, a.bn_invstd = nullptr;
  a.passes = passes;
  return launch_light<64, 32, true, true, true>(a, "conv0_bwd_gemm", stream);
}

}  // namespace advb
