// LCNN convolution blocks on the 5th-generation tensor cores (tcgen05 / TMEM, sm_100a), forward and input-gradient
// backward, fp32-class accuracy through the 3xTF32 split (tc_common.cuh).
//
// Same contract as conv.cu (which stays as the fp32 SIMT cross-check): forward = conv + bias -> Max-Feature-Map ->
// [2x2 max-pool] -> [BatchNorm(eval)] with a 3-bit winner code per output; backward = un-BN, un-pool, un-MFM of the
// stage gradient followed by the transposed convolution.  Replaces src/models/lcnn.py:120-157 (Conv2d,
// MaxFeatureMap2D :89-95, MaxPool2d, BatchNorm2d(affine=False)) and its autograd input gradient.
//
// Implicit GEMM over FLATTENED PADDED PIXELS.  Activations are NHWC with a zero border equal to the conv padding,
// so pixel q = yp * Wp + xp of a clip is row q of a [Hp*Wp, C] matrix and the tap (dy, dx) of a KSxKS filter is the
// same matrix shifted by dy * Wp + dx rows:
//     out[q, n] = sum_tap sum_k A[q + dy*Wp + dx, k] * Wt[tap][n][k]        (border rows q are computed and dropped)
// A CTA owns a "super-tile" of R full image rows of one clip = up to 4 M-tiles of 128 pixels, each with its own
// fp32 accumulator in TMEM (<= 4 x 128 columns).  It stages ONE halo band (all rows any tap of any of its M-tiles
// touches) in shared memory as K-major SWIZZLE_128B rows of 32 channels, split into tf32 hi/lo parts, and every
// (M-tile, tap) operand is just a different start row of that band (descriptor start address; see tc_common.cuh
// for the hardware check).  Weights are pre-packed per call into the exact shared-memory image (hi/lo, swizzled)
// and streamed per (32-channel chunk, tap) slice through a ring of TMA bulk copies tracked by mbarriers.  One
// thread issues the MMAs; tcgen05.commit releases ring slots and signals chunk completion.  The epilogue reads the
// accumulators with tcgen05.ld (thread = pixel, so MFM's channel pair c / c + C/2 is register-local), stages the
// tile in shared memory and finishes pooling / BN / coalesced stores from there.
//
// Algorithmic HBM bytes per launch: stage input read once + stage output written once + 1 byte of
// winner code per output element; weights (<= 221 KB per layer) stay L2-resident.
#include "conv.cuh"
#include "tc_common.cuh"

#include <stdlib.h>

namespace advb {

namespace {

using namespace tc;

constexpr int TC_THREADS = 256;
// NM = M-tiles (128 pixels each) per CTA is a template parameter: 4 for the weight-heavy 3x3 layers (each weight
// slice streamed from L2 is reused by 512 pixels; one CTA per SM), 2 for the light layers (first block, 1x1
// convolutions), whose <= 113 KB footprint lets two CTAs share an SM and hide each other's fill / epilogue.

__host__ __device__ constexpr int pow2_ceil(int v) {
  int p = 32;
  while (p < v) p <<= 1;
  return p;
}

template <int NOUT, int NM>
struct TcCfg {
  static constexpr int NSTRIDE = pow2_ceil(NOUT);         // TMEM columns reserved per M-tile accumulator
  static constexpr int TMEM_COLS = pow2_ceil(NM * NSTRIDE);  // 32 .. 512 (power of two)
  static constexpr int SLICE_BYTES = 2 * NOUT * 128;      // hi + lo image of one (chunk, tap) weight slice
  // ring depth: deeper for small slices, but never so deep that an NM = 2 kernel loses its second CTA per SM
  static constexpr int NST = (SLICE_BYTES <= 8192 && NM > 2) ? 4 : 2;
};

struct TcArgs {
  int B, H, W;      // conv grid (pre-pool)
  int Ho, Wo;       // block output grid
  int R;            // image rows per super-tile
  int band_rows;    // rows of the shared-memory band (multiple of 8)
  const unsigned char* wpack;
  // forward
  const float* in;
  float* out;
  int out_pad;
  unsigned char* codes;
  const float* bias;
  const float* bn_mean;
  const float* bn_invstd;
  // backward
  const float* gout;
  const unsigned char* codes_in;
  float* gin;
  int passes;  // 3 = 3xTF32 (default), 1 = single-pass tf32 (fast, reduced precision)
  FastDiv dWp, dW, dWo;  // padded row length, conv grid width, output grid width
};

// KS: filter size; KTOT: contraction channels per tap (fwd: Cin, bwd: Cout); NOUT: GEMM N (fwd: Cout, bwd: Cin).
template <int KS, int KTOT, int NOUT, bool POOL, bool BWD, bool IM2COL = false, int NM = 4>
__global__ void __launch_bounds__(TC_THREADS, NM == 1 ? 3 : (NM == 2 ? 2 : 1)) conv_tc_kernel(TcArgs a) {
  using Cfg = TcCfg<NOUT, NM>;
  constexpr int NM_MAX = NM;
  constexpr int PC = KS / 2, NTAP = KS * KS, NKC = (KTOT + 31) / 32;
  constexpr int NSLICE = NKC * NTAP;
  constexpr int NST = Cfg::NST;
  constexpr uint32_t IDESC = idesc_tf32(128, NOUT);

  extern __shared__ unsigned char smem_raw[];
  unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  unsigned char* a_hi = base;
  unsigned char* a_lo = a_hi + (size_t)a.band_rows * 128;
  unsigned char* wring = a_lo + (size_t)a.band_rows * 128;
  __shared__ uint64_t bar_full[NST], bar_empty[NST], bar_chunk;
  __shared__ uint32_t tmem_base_s;
  __shared__ float s_bias[BWD ? 1 : NOUT];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int b = blockIdx.y, tile = blockIdx.x;
  const int Wp = a.W + 2 * PC, Hp = a.H + 2 * PC;
  const int Heff = (!BWD && POOL) ? 2 * a.Ho : a.H;
  const int y0 = tile * a.R;
  const int rows = min(a.R, Heff - y0);
  const int npx = rows * Wp;
  const int nM = (npx + 127) >> 7;
  const int q_lo = (y0 + PC) * Wp - PC * Wp - PC;            // flat padded pixel of band row 0
  const int band_used = nM * 128 + 2 * (PC * Wp + PC);

  if (tid == 0) {
    for (int s = 0; s < NST; ++s) {
      mbar_init(&bar_full[s], 1);
      mbar_init(&bar_empty[s], 1);
    }
    mbar_init(&bar_chunk, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc<Cfg::TMEM_COLS>(&tmem_base_s);
  if (!BWD)
    for (int i = tid; i < NOUT; i += TC_THREADS) s_bias[i] = __ldg(a.bias + i);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;

  // weight producer (thread 32): keeps the ring NST slices ahead of the MMA thread
  int produced = 0;
  auto produce_upto = [&](int limit) {  // called by all lanes of warp 1
    if (limit > NSLICE) limit = NSLICE;
    for (; produced < limit; ++produced) {
      const int s = produced % NST, use = produced / NST;
      if (use > 0) mbar_wait(&bar_empty[s], (uint32_t)((use - 1) & 1));
      if (lane == 0) {
        mbar_expect_tx(&bar_full[s], Cfg::SLICE_BYTES);
        bulk_g2s(wring + (size_t)s * Cfg::SLICE_BYTES, a.wpack + (size_t)produced * Cfg::SLICE_BYTES,
                 Cfg::SLICE_BYTES, &bar_full[s]);
      }
      __syncwarp();
    }
  };
  if (warp == 1) produce_upto(NST);

#pragma unroll 1
  for (int kc = 0; kc < NKC; ++kc) {
    if (kc > 0) {  // the MMAs of the previous chunk still read the band
      mbar_wait(&bar_chunk, (uint32_t)((kc - 1) & 1));
      tc_fence_after();
    }
    // ---- stage chunk kc of the band: 32 channels per row, tf32 hi / lo, SWIZZLE_128B ----
    // Two phases per batch of FILL_U items (16-byte channel groups): first every global load of the batch is issued
    // from branch-free, clamped addresses, then the batch is post-processed.  The band is read exactly once, so
    // latency - not bandwidth - is the enemy: an earlier version whose per-item bounds checks were branches had its
    // loads serialised by the compiler and spent half of its samples waiting on them (profiles/r01_conv_tc_fill.md).
    {
      constexpr int FILL_U = 8;
      const int total = band_used * 8;
      const int c4 = tid & 7;  // TC_THREADS is a multiple of 8: a thread always owns the same channel group
      const int ch = 32 * kc + 4 * c4;
      const bool ch_ok = ch < KTOT;
      const int npix = Hp * Wp;
      if (BWD) {
        constexpr int Ch = KTOT / 2;
        const int half = ch >= Ch ? 1 : 0;
        const int c = ch - half * Ch;
        float4 sc = make_float4(1.f, 1.f, 1.f, 1.f);
        if (a.bn_invstd != nullptr && ch_ok) sc = __ldg(reinterpret_cast<const float4*>(a.bn_invstd + c));
#pragma unroll 1
        for (int i0 = tid; i0 < total; i0 += FILL_U * TC_THREADS) {
          float4 g[FILL_U];
          uchar4 cd[FILL_U];
          unsigned want[FILL_U];
#pragma unroll
          for (int u = 0; u < FILL_U; ++u) {
            const int i = i0 + u * TC_THREADS;
            const int q = q_lo + (i >> 3);
            bool ok = ch_ok && i < total && q >= 0 && q < npix;
            const int qq = ok ? q : 0;
            const int yp = fdiv(qq, a.dWp), xp = qq - yp * Wp;
            const int y = yp - PC, x = xp - PC;
            ok = ok && y >= 0 && y < a.H && x >= 0 && x < a.W;
            const int py = POOL ? (y >> 1) : y, px = POOL ? (x >> 1) : x;
            ok = ok && py < a.Ho && px < a.Wo;
            const size_t o = ok ? (((size_t)b * a.Ho + py) * a.Wo + px) * Ch + c : 0;
            g[u] = __ldg(reinterpret_cast<const float4*>(a.gout + o));
            cd[u] = __ldg(reinterpret_cast<const uchar4*>(a.codes_in + o));
            want[u] = ok ? ((POOL ? (unsigned)(((y & 1) << 1) | (x & 1)) : 0u) | (unsigned)(half << 2)) : 0xffu;
          }
#pragma unroll
          for (int u = 0; u < FILL_U; ++u) {
            const int i = i0 + u * TC_THREADS;
            if (i < total) {
              float4 v;
              v.x = cd[u].x == want[u] ? g[u].x * sc.x : 0.f;
              v.y = cd[u].y == want[u] ? g[u].y * sc.y : 0.f;
              v.z = cd[u].z == want[u] ? g[u].z * sc.z : 0.f;
              v.w = cd[u].w == want[u] ? g[u].w * sc.w : 0.f;
              float4 hi, lo;
              split_tf32(v.x, hi.x, lo.x);
              split_tf32(v.y, hi.y, lo.y);
              split_tf32(v.z, hi.z, lo.z);
              split_tf32(v.w, hi.w, lo.w);
              const uint32_t off = sw128_chunk(i >> 3, c4);
              *reinterpret_cast<float4*>(a_hi + off) = hi;
              *reinterpret_cast<float4*>(a_lo + off) = lo;
            }
          }
        }
      } else if (IM2COL) {
        // first block: contraction index = tap (4 per channel group), gathered from the zero-bordered image
        const int Wi = a.W + 4;
        int toff[4];
        bool tok[4];
#pragma unroll
        for (int u4 = 0; u4 < 4; ++u4) {
          const int tap = 4 * c4 + u4;
          tok[u4] = tap < 25;
          const int dy = tap / 5, dx = tap - dy * 5;
          toff[u4] = tok[u4] ? dy * Wi + dx : 0;
        }
        const float* imgb = a.in + (size_t)b * (a.H + 4) * Wi;
#pragma unroll 1
        for (int i0 = tid; i0 < total; i0 += FILL_U * TC_THREADS) {
          float t[FILL_U][4];
#pragma unroll
          for (int u = 0; u < FILL_U; ++u) {
            const int i = i0 + u * TC_THREADS;
            const int q = q_lo + (i >> 3);
            const bool ok = i < total && q >= 0 && q < npix;
            const int qq = ok ? q : 0;
            const int y = fdiv(qq, a.dWp), x = qq - y * Wp;
            const float* img = imgb + (size_t)y * Wi + x;
#pragma unroll
            for (int u4 = 0; u4 < 4; ++u4) {
              const float val = __ldg(img + toff[u4]);
              t[u][u4] = (ok && tok[u4]) ? val : 0.f;
            }
          }
#pragma unroll
          for (int u = 0; u < FILL_U; ++u) {
            const int i = i0 + u * TC_THREADS;
            if (i < total) {
              float4 hi, lo;
              split_tf32(t[u][0], hi.x, lo.x);
              split_tf32(t[u][1], hi.y, lo.y);
              split_tf32(t[u][2], hi.z, lo.z);
              split_tf32(t[u][3], hi.w, lo.w);
              const uint32_t off = sw128_chunk(i >> 3, c4);
              *reinterpret_cast<float4*>(a_hi + off) = hi;
              *reinterpret_cast<float4*>(a_lo + off) = lo;
            }
          }
        }
      } else {
        const float* inb = a.in + (size_t)b * npix * KTOT + ch;
#pragma unroll 1
        for (int i0 = tid; i0 < total; i0 += FILL_U * TC_THREADS) {
          float4 v[FILL_U];
#pragma unroll
          for (int u = 0; u < FILL_U; ++u) {
            const int i = i0 + u * TC_THREADS;
            const int q = q_lo + (i >> 3);
            const bool ok = ch_ok && i < total && q >= 0 && q < npix;
            const float4 val = __ldg(reinterpret_cast<const float4*>(ok ? inb + (size_t)q * KTOT : a.in));
            v[u] = ok ? val : make_float4(0.f, 0.f, 0.f, 0.f);
          }
#pragma unroll
          for (int u = 0; u < FILL_U; ++u) {
            const int i = i0 + u * TC_THREADS;
            if (i < total) {
              float4 hi, lo;
              split_tf32(v[u].x, hi.x, lo.x);
              split_tf32(v[u].y, hi.y, lo.y);
              split_tf32(v[u].z, hi.z, lo.z);
              split_tf32(v[u].w, hi.w, lo.w);
              const uint32_t off = sw128_chunk(i >> 3, c4);
              *reinterpret_cast<float4*>(a_hi + off) = hi;
              *reinterpret_cast<float4*>(a_lo + off) = lo;
            }
          }
        }
      }
    }
    fence_proxy_async();
    __syncthreads();

    if (warp == 0) {
      // ---- MMA issue: the whole warp walks the (warp-uniform) loop nest so every operand stays in uniform
      // registers; one elected lane issues.  Order: tap slice > 8-wide k-step > pass > M-tile. ----
      const bool leader = elect_one();
      const int kvalid = (KTOT - 32 * kc) >= 32 ? 32 : (KTOT - 32 * kc);
      const uint32_t a_hi_addr = smem_u32(a_hi), a_lo_addr = smem_u32(a_lo);
#pragma unroll 1
      for (int tap = 0; tap < NTAP; ++tap) {
        const int sl = kc * NTAP + tap, s = sl % NST;
        mbar_wait(&bar_full[s], (uint32_t)((sl / NST) & 1));
        tc_fence_after();
        const uint32_t w_hi = smem_u32(wring + (size_t)s * Cfg::SLICE_BYTES), w_lo = w_hi + NOUT * 128;
        const int dy = tap / KS, dx = tap - dy * KS;
        const uint32_t row_off = (uint32_t)(dy * Wp + dx) * 128u;
#pragma unroll 1
        for (int ks = 0; ks < kvalid / 8; ++ks) {
          const uint64_t bh = desc_sw128(w_hi + ks * 32), bl = desc_sw128(w_lo + ks * 32);
          const uint64_t ah = desc_sw128(a_hi_addr + row_off + ks * 32), al = desc_sw128(a_lo_addr + row_off + ks * 32);
          const uint32_t acc0 = (sl > 0 || ks > 0) ? 1u : 0u;
          // fully unrolled over (pass, M-tile): the operand moves to uniform registers batch up ahead of a run of
          // back-to-back UTCHMMAs instead of serialising with each one
#pragma unroll
          for (int m = 0; m < NM; ++m) {
            if (m < nM && leader) {
              const uint32_t dcol = tmem + m * Cfg::NSTRIDE;
              const uint64_t moff = (uint64_t)(m * ((128 * 128) >> 4));
              mma_tf32(dcol, ah + moff, bh, IDESC, acc0);
              if (a.passes == 3) {
                mma_tf32(dcol, ah + moff, bl, IDESC, 1u);
                mma_tf32(dcol, al + moff, bh, IDESC, 1u);
              }
            }
          }
        }
        if (leader) mma_commit(&bar_empty[s]);  // slot free once these MMAs have read it
        __syncwarp();
      }
      if (leader) mma_commit(&bar_chunk);
      __syncwarp();
    } else if (warp == 1) {
      produce_upto((kc + 1) * NTAP + NST);
    }
  }
  mbar_wait(&bar_chunk, (uint32_t)((NKC - 1) & 1));
  tc_fence_after();

  // ---- epilogue 1: TMEM -> registers -> shared staging (overlays band + ring, both idle now) ----
  constexpr int CS = BWD ? NOUT : NOUT / 2;  // channels per staged pixel
  constexpr int SS = CS + 4;                 // padded row (floats): conflict-free 128-bit stores
  float* stage = reinterpret_cast<float*>(base);
  unsigned long long* flags = reinterpret_cast<unsigned long long*>(stage + (size_t)NM_MAX * 128 * SS);
  {
    const int wq = warp & 3, wg = warp >> 2;
    for (int m = wg; m < nM; m += 2) {
      const int r = m * 128 + wq * 32 + lane;
      const uint32_t taddr = tmem + ((uint32_t)(wq * 32) << 16) + m * Cfg::NSTRIDE;
      float* srow = stage + (size_t)r * SS;
      if (BWD) {
#pragma unroll 1
        for (int c0 = 0; c0 < NOUT; c0 += 16) {
          uint32_t v[16];
          tmem_ld16_issue(taddr + c0, v);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; j += 4)
            *reinterpret_cast<float4*>(srow + c0 + j) = make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]),
                                                                   __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
        }
      } else {
        unsigned long long fl = 0ull;
#pragma unroll 1
        for (int c0 = 0; c0 < CS; c0 += 16) {
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
            fl |= (unsigned long long)(sel ? 1u : 0u) << (c0 + j);
          }
#pragma unroll
          for (int j = 0; j < 16; j += 4)
            *reinterpret_cast<float4*>(srow + c0 + j) = make_float4(m4[j], m4[j + 1], m4[j + 2], m4[j + 3]);
        }
        flags[r] = fl;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<Cfg::TMEM_COLS>(tmem);

  // ---- epilogue 2: pooling / BatchNorm / coalesced stores ----
  constexpr int C4 = CS / 4;
  if (BWD && IM2COL) {
    // First block (1 input channel): the GEMM produced Z[pixel][tap] = sum_co Gexp[pixel][co] * W[co][tap].  The
    // horizontal half of col2im stays inside one image row, hence inside this tile:
    //   T[y][x][dy] = sum_dx Z[y][x - dx + 2][5 dy + dx];   conv0_col2im_rows() then sums the 5 rows.
    const int items = rows * a.W * 5;
    for (int i = tid; i < items; i += TC_THREADS) {
      const int i5 = i / 5, dy = i - 5 * i5, yl = fdiv(i5, a.dW), x = i5 - yl * a.W;
      float acc = 0.f;
#pragma unroll
      for (int dx = 0; dx < 5; ++dx) {
        const int xs = x - dx + 2;
        if (xs >= 0 && xs < a.W) acc += stage[(size_t)(yl * Wp + xs) * SS + 5 * dy + dx];
      }
      a.gin[((size_t)b * a.H + y0) * a.W * 5 + i] = acc;
    }
  } else if (BWD) {
    const int items = rows * a.W * C4;
    for (int i = tid; i < items; i += TC_THREADS) {
      const int ic = i / C4, c4 = i - C4 * ic, yl = fdiv(ic, a.dW), x = ic - yl * a.W;
      const float4 v = *reinterpret_cast<const float4*>(stage + (size_t)(yl * Wp + x + PC) * SS + 4 * c4);
      *reinterpret_cast<float4*>(a.gin + (((size_t)b * a.H + y0 + yl) * a.W + x) * NOUT + 4 * c4) = v;
    }
  } else {
    const int Hop = a.Ho + 2 * a.out_pad, Wop = a.Wo + 2 * a.out_pad;
    const int orows = POOL ? rows / 2 : rows, oy0 = POOL ? y0 / 2 : y0;
    const int items = orows * a.Wo * C4;
    for (int i = tid; i < items; i += TC_THREADS) {
      const int ic = i / C4, c4 = i - C4 * ic, yl = fdiv(ic, a.dWo), x = ic - yl * a.Wo;
      const int c = 4 * c4;
      float4 v;
      uchar4 cd;
      if (POOL) {
        const int r00 = (2 * yl) * Wp + 2 * x + PC;
        const int rr[4] = {r00, r00 + 1, r00 + Wp, r00 + Wp + 1};
        float best[4];
        unsigned code[4];
        {
          const float4 t = *reinterpret_cast<const float4*>(stage + (size_t)rr[0] * SS + c);
          const unsigned f = (unsigned)(flags[rr[0]] >> c) & 15u;
          best[0] = t.x, best[1] = t.y, best[2] = t.z, best[3] = t.w;
#pragma unroll
          for (int k = 0; k < 4; ++k) code[k] = ((f >> k) & 1u) << 2;
        }
#pragma unroll
        for (int p = 1; p < 4; ++p) {
          const float4 t = *reinterpret_cast<const float4*>(stage + (size_t)rr[p] * SS + c);
          const unsigned f = (unsigned)(flags[rr[p]] >> c) & 15u;
          const float tv[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
          for (int k = 0; k < 4; ++k)
            if (tv[k] > best[k]) {
              best[k] = tv[k];
              code[k] = (((f >> k) & 1u) << 2) | (unsigned)p;
            }
        }
        v = make_float4(best[0], best[1], best[2], best[3]);
        cd = make_uchar4((unsigned char)code[0], (unsigned char)code[1], (unsigned char)code[2], (unsigned char)code[3]);
      } else {
        const int r = yl * Wp + x + PC;
        v = *reinterpret_cast<const float4*>(stage + (size_t)r * SS + c);
        const unsigned f = (unsigned)(flags[r] >> c) & 15u;
        cd = make_uchar4((unsigned char)((f & 1u) << 2), (unsigned char)(((f >> 1) & 1u) << 2),
                         (unsigned char)(((f >> 2) & 1u) << 2), (unsigned char)(((f >> 3) & 1u) << 2));
      }
      if (a.bn_mean != nullptr) {  // running_mean is a borrowed PyTorch tensor: no alignment assumption
        const float4 mu = make_float4(__ldg(a.bn_mean + c), __ldg(a.bn_mean + c + 1), __ldg(a.bn_mean + c + 2),
                                      __ldg(a.bn_mean + c + 3));
        const float4 is = __ldg(reinterpret_cast<const float4*>(a.bn_invstd + c));
        v.x = (v.x - mu.x) * is.x;
        v.y = (v.y - mu.y) * is.y;
        v.z = (v.z - mu.z) * is.z;
        v.w = (v.w - mu.w) * is.w;
      }
      const int oy = oy0 + yl;
      *reinterpret_cast<float4*>(a.out + (((size_t)b * Hop + oy + a.out_pad) * Wop + x + a.out_pad) * CS + c) = v;
      *reinterpret_cast<uchar4*>(a.codes + (((size_t)b * a.Ho + oy) * a.Wo + x) * CS + c) = cd;
    }
  }
}

// Weight slices in consumption order [chunk kc][tap]: each slice = hi image then lo image of the K-major
// SWIZZLE_128B matrix [NOUT rows][32 k].  Forward: (n, k) = W[co = n][ci = 32 kc + k][tap]; backward:
// (n, k) = W[co = 32 kc + k][ci = n][KS*KS - 1 - tap] (flipped taps, transposed channels).
__global__ void pack_tc_kernel(const float* __restrict__ w, unsigned char* __restrict__ dst, int Cout, int Cin, int KS,
                               int bwd, int nout_pad) {
  const int NTAP = KS * KS;
  const int KTOT = bwd ? Cout : Cin, NOUT = nout_pad, n_valid = bwd ? Cin : Cout;
  const int NKC = (KTOT + 31) / 32;
  const int total = NKC * NTAP * NOUT * 32;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int k = i & 31, n = (i >> 5) % NOUT, sl = i / (32 * NOUT);
    const int tap = sl % NTAP, kc = sl / NTAP;
    const int kk = 32 * kc + k;
    float v = 0.f;
    if (kk < KTOT && n < n_valid) {
      const int co = bwd ? kk : n, ci = bwd ? n : kk;
      const int t = bwd ? NTAP - 1 - tap : tap;
      v = w[((size_t)co * Cin + ci) * NTAP + t];
    }
    float hi, lo;
    tc::split_tf32(v, hi, lo);
    const size_t slice = (size_t)sl * 2 * NOUT * 128;
    const uint32_t off = (uint32_t)(n * 128 + ((((k >> 2) ^ (n & 7)) << 4) | ((k & 3) << 2)));
    *reinterpret_cast<float*>(dst + slice + off) = hi;
    *reinterpret_cast<float*>(dst + slice + (size_t)NOUT * 128 + off) = lo;
  }
}

// Tuning knob (environment, read once): M-tiles per CTA of the 1x1 layers.  Measured on B200 (profiles/): see DESIGN.md.
int tune_light_nm() {
  static const int v = [] {
    const char* e = getenv("ADVB_LIGHT_NM");
    return e != nullptr ? atoi(e) : 2;
  }();
  return v;
}

struct TcPlan {
  int R, tiles, band_rows;
  size_t smem;
};

template <int NOUT, int NM_MAX>
TcPlan make_plan(int H, int W, int Ho, int KS, bool pool, bool bwd) {
  using Cfg = TcCfg<NOUT, NM_MAX>;
  const int pc = KS / 2, Wp = W + 2 * pc;
  const int Heff = (!bwd && pool) ? 2 * Ho : H;
  const bool even = !bwd && pool;
  int Rmax = (NM_MAX * 128) / Wp;
  if (even) Rmax &= ~1;
  if (Rmax < (even ? 2 : 1)) return TcPlan{0, 0, 0, 0};
  TcPlan p;
  p.tiles = (Heff + Rmax - 1) / Rmax;
  p.R = (Heff + p.tiles - 1) / p.tiles;
  if (even && (p.R & 1)) ++p.R;
  if (p.R > Rmax) p.R = Rmax;
  p.tiles = (Heff + p.R - 1) / p.R;
  const int nM = (p.R * Wp + 127) / 128;
  p.band_rows = (nM * 128 + 2 * (pc * Wp + pc) + 7) & ~7;
  const int CS = bwd ? NOUT : NOUT / 2;
  const size_t stage = (size_t)NM_MAX * 128 * (CS + 4) * 4 + (size_t)NM_MAX * 128 * 8;
  size_t s = (size_t)2 * p.band_rows * 128 + (size_t)Cfg::NST * Cfg::SLICE_BYTES;
  if (s < stage) s = stage;
  p.smem = s + 1024;
  return p;
}

template <int KS, int KTOT, int NOUT, bool POOL, bool BWD, bool IM2COL = false, int NM = 4>
int launch_tc(TcArgs a, const char* tag, cudaStream_t stream) {
  const TcPlan p = make_plan<NOUT, NM>(a.H, a.W, a.Ho, KS, POOL, BWD);
  ADVB_CHECK(p.tiles > 0 && p.smem <= 227 * 1024, "tensor-core conv tile does not fit (image too wide)");
  a.R = p.R;
  a.band_rows = p.band_rows;
  a.dWp = make_fastdiv(a.W + 2 * (KS / 2));
  a.dW = make_fastdiv(a.W);
  a.dWo = make_fastdiv(a.Wo);
  auto kern = conv_tc_kernel<KS, KTOT, NOUT, POOL, BWD, IM2COL, NM>;
  ADVB_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem));
  dim3 grid(p.tiles, a.B);
  kern<<<grid, TC_THREADS, p.smem, stream>>>(a);
  ADVB_KERNEL_OK(tag, stream);
  return 0;
}

}  // namespace

size_t conv_tc_pack_bytes(int Cout, int Cin, int KS, bool bwd) {
  if (KS == 5 && Cin == 1) {
    Cin = bwd ? 32 : 25;  // backward GEMM N = 25 taps padded to 32
    KS = 1;
  }
  const int KTOT = bwd ? Cout : Cin, NOUT = bwd ? Cin : Cout;
  return (size_t)((KTOT + 31) / 32) * KS * KS * 2 * NOUT * 128;
}

int conv_tc_pack(const float* w, unsigned char* wf, unsigned char* wd, int Cout, int Cin, int KS, cudaStream_t stream) {
  int nout_bwd = Cin;
  if (KS == 5 && Cin == 1) {  // (Cout,1,5,5) read as a 1x1 conv over 25 "channels" (the taps)
    Cin = 25;
    KS = 1;
    nout_bwd = 32;
  }
  const int n = (int)(conv_tc_pack_bytes(Cout, Cin, KS, false) / 8);
  pack_tc_kernel<<<cdiv(n, 256), 256, 0, stream>>>(w, wf, Cout, Cin, KS, 0, Cout);
  ADVB_KERNEL_OK("pack_tc_fwd", stream);
  if (wd != nullptr) {
    const int nb = ((Cout + 31) / 32) * KS * KS * nout_bwd * 32;
    pack_tc_kernel<<<cdiv(nb, 256), 256, 0, stream>>>(w, wd, Cout, Cin, KS, 1, nout_bwd);
    ADVB_KERNEL_OK("pack_tc_bwd", stream);
  }
  return 0;
}

size_t conv_tc_pack_bytes_padded(int Cout, int Cin, int KS, bool bwd, int npad) {
  const int KTOT = bwd ? Cout : Cin;
  return (size_t)((KTOT + 31) / 32) * KS * KS * 2 * npad * 128;
}

int conv_tc_pack_padded(const float* w, unsigned char* wf, unsigned char* wd, int Cout, int Cin, int KS, int npad_f, int npad_b,
                        cudaStream_t stream) {
  ADVB_CHECK(npad_f >= Cout && npad_b >= Cin && npad_f % 16 == 0 && npad_b % 16 == 0, "padded pack: the padded N must cover the channels");
  const int nf = ((Cin + 31) / 32) * KS * KS * npad_f * 32;
  pack_tc_kernel<<<cdiv(nf, 256), 256, 0, stream>>>(w, wf, Cout, Cin, KS, 0, npad_f);
  ADVB_KERNEL_OK("pack_tc_fwd", stream);
  const int nb = ((Cout + 31) / 32) * KS * KS * npad_b * 32;
  pack_tc_kernel<<<cdiv(nb, 256), 256, 0, stream>>>(w, wd, Cout, Cin, KS, 1, npad_b);
  ADVB_KERNEL_OK("pack_tc_bwd", stream);
  return 0;
}

bool conv_tc_supported(int Cin, int Cout, int KS, bool pool) {
  if (KS == 5 && pool) return Cin == 1 && Cout == 64;  // forward only (im2col); its backward is conv0_backward
  if (KS == 1 && !pool) return (Cin == 32 && Cout == 64) || (Cin == 48 && Cout == 96) || (Cin == 64 && Cout == 128);
  if (KS == 3 && pool) return (Cin == 32 && Cout == 96) || (Cin == 48 && Cout == 128) || (Cin == 32 && Cout == 64);
  if (KS == 3 && !pool) return Cin == 64 && Cout == 64;
  return false;
}

int conv_tc_forward(const ConvFwdArgs& f, const unsigned char* wpack, int passes, cudaStream_t stream) {
  ADVB_CHECK(f.in_pad == f.KS / 2, "tensor-core conv needs the input border to equal the conv padding");
  if (conv_light_supported(f.Cin, f.Cout, f.KS, f.pool, false)) return conv_light_forward(f, wpack, passes, stream);
  if (conv_p3_supported(f.Cin, f.Cout, f.KS, f.pool, f.W)) return conv_p3_forward(f, wpack, passes, stream);
  TcArgs a{};
  if (f.KS == 5) {  // first block: contraction over the 25 taps (padded to 32), pixel grid without border
    ADVB_CHECK(f.Cin == 1 && f.Cout == 64 && f.pool, "5x5 tensor-core conv is the LCNN first block only");
    a.B = f.B, a.H = f.H, a.W = f.W, a.Ho = f.Ho, a.Wo = f.Wo;
    a.wpack = wpack;
    a.in = f.in, a.out = f.out, a.out_pad = f.out_pad, a.codes = f.codes, a.bias = f.bias;
    a.bn_mean = f.bn_mean, a.bn_invstd = f.bn_invstd;
    a.passes = passes;
    return launch_tc<1, 32, 64, true, false, true, 2>(a, f.tag, stream);
  }
  a.B = f.B, a.H = f.H, a.W = f.W, a.Ho = f.Ho, a.Wo = f.Wo;
  a.wpack = wpack;
  a.in = f.in, a.out = f.out, a.out_pad = f.out_pad, a.codes = f.codes, a.bias = f.bias;
  a.bn_mean = f.bn_mean, a.bn_invstd = f.bn_invstd;
  a.passes = passes;
#define ADVB_TCF(KS_, CI_, CO_, POOL_)                                                              \
  if (f.KS == KS_ && f.Cin == CI_ && f.Cout == CO_ && f.pool == POOL_) {                            \
    if (KS_ == 1 && tune_light_nm() == 1) return launch_tc<KS_, CI_, CO_, POOL_, false, false, 1>(a, f.tag, stream); \
    return launch_tc<KS_, CI_, CO_, POOL_, false, false, (KS_ == 1 ? 2 : 4)>(a, f.tag, stream);     \
  }
  ADVB_TCF(1, 32, 64, false);
  ADVB_TCF(1, 48, 96, false);
  ADVB_TCF(1, 64, 128, false);
  ADVB_TCF(3, 32, 96, true);
  ADVB_TCF(3, 48, 128, true);
  ADVB_TCF(3, 32, 64, true);
  ADVB_TCF(3, 64, 64, false);
#undef ADVB_TCF
  set_error("conv shape has no tensor-core instantiation");
  return 1;
}

// d loss / d cepstral image of the first block: row sums of T (see the col2im epilogue)
__global__ void conv0_col2im_rows_kernel(const float* __restrict__ T, float* __restrict__ gin, int H, int W, int n) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int x = i % W, y = (i / W) % H, b = i / (W * H);
    float acc = 0.f;
#pragma unroll
    for (int dy = 0; dy < 5; ++dy) {
      const int ys = y - dy + 2;
      if (ys >= 0 && ys < H) acc += __ldg(T + (((size_t)b * H + ys) * W + x) * 5 + dy);
    }
    gin[i] = acc;
  }
}

int conv0_tc_backward(const float* gout, const unsigned char* codes, const unsigned char* wpack, float* T, float* gin,
                      int B, int H, int W, int Ho, int Wo, int passes, cudaStream_t stream) {
  TcArgs a{};
  a.B = B, a.H = H, a.W = W, a.Ho = Ho, a.Wo = Wo;
  a.wpack = wpack;
  a.gout = gout, a.codes_in = codes, a.gin = T, a.bn_invstd = nullptr;
  a.passes = passes;
  if (conv_light_supported(1, 64, 5, true, true))
    ADVB_TRY(conv0_light_backward_gemm(gout, codes, wpack, T, B, H, W, Ho, Wo, passes, stream));
  else
    ADVB_TRY((launch_tc<1, 64, 32, true, true, true, 2>(a, "conv0_bwd_gemm", stream)));
  const int n = B * H * W;
  conv0_col2im_rows_kernel<<<cdiv(n, 256), 256, 0, stream>>>(T, gin, H, W, n);
  ADVB_KERNEL_OK("conv0_bwd_rows", stream);
  return 0;
}

int conv_tc_backward(const ConvBwdArgs& g, const unsigned char* wpack, int passes, cudaStream_t stream) {
  if (conv_light_supported(g.Cin, g.Cout, g.KS, g.pool, true)) return conv_light_backward(g, wpack, passes, stream);
  if (conv_p3_supported(g.Cin, g.Cout, g.KS, g.pool, g.W)) return conv_p3_backward(g, wpack, passes, stream);
  TcArgs a{};
  a.B = g.B, a.H = g.H, a.W = g.W, a.Ho = g.Ho, a.Wo = g.Wo;
  a.wpack = wpack;
  a.gout = g.gout, a.codes_in = g.codes, a.gin = g.gin, a.bn_invstd = g.bn_invstd;
  a.passes = passes;
#define ADVB_TCB(KS_, CI_, CO_, POOL_)                                                              \
  if (g.KS == KS_ && g.Cin == CI_ && g.Cout == CO_ && g.pool == POOL_) {                            \
    if (KS_ == 1 && tune_light_nm() == 1) return launch_tc<KS_, CO_, CI_, POOL_, true, false, 1>(a, g.tag, stream); \
    return launch_tc<KS_, CO_, CI_, POOL_, true, false, 2>(a, g.tag, stream);                       \
  }
  ADVB_TCB(1, 32, 64, false);
  ADVB_TCB(1, 48, 96, false);
  ADVB_TCB(1, 64, 128, false);
  ADVB_TCB(3, 32, 96, true);
  ADVB_TCB(3, 48, 128, true);
  ADVB_TCB(3, 32, 64, true);
  ADVB_TCB(3, 64, 64, false);
#undef ADVB_TCB
  set_error("conv shape has no tensor-core instantiation (backward)");
  return 1;
}

}  // namespace advb
