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

constexpr int PW = 256;            // worker threads
constexpr int PT = PW + 64;        // + MMA warp + weight warp
constexpr int NM3 = 2;             // M-tiles (128 pixels) per tile
constexpr int NI_MAX = 11;         // prefetch items per worker thread: ceil((2*128 + 2*(42+1)) * 8 / 256)

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
  FastDiv dTiles, dWp, dW, dWo;
};

template <int KTOT, int NOUT, bool POOL, bool BWD>
struct P3Cfg {
  static constexpr int NKC = (KTOT + 31) / 32;
  static constexpr int NSLICE = NKC * 9;
  static constexpr int NSTRIDE = p3_pow2(NOUT);
  static constexpr int TMEM_COLS = p3_pow2(2 * NM3 * NSTRIDE);
  static constexpr int SLICE_BYTES = 2 * NOUT * 128;
  static constexpr int CS = BWD ? NOUT : NOUT / 2;
  static constexpr int SS = CS + 4;
  static constexpr int STAGE_BYTES = NM3 * 128 * SS * 4 + NM3 * 128 * 8;
  static constexpr int BAND_BYTES = 2 * 344 * 128;
  static constexpr int ROOM = 227 * 1024 - 1024 - BAND_BYTES - STAGE_BYTES;
  static constexpr int NST = ROOM / SLICE_BYTES >= 4 ? 4 : (ROOM / SLICE_BYTES >= 3 ? 3 : 2);
  static constexpr size_t SMEM = (size_t)NST * SLICE_BYTES + BAND_BYTES + STAGE_BYTES + 1024;
  static_assert(TMEM_COLS <= 512, "two accumulator sets must fit the 512 TMEM columns");
  static_assert(ROOM / SLICE_BYTES >= 2, "weight ring does not fit");
};

template <int KTOT, int NOUT, bool POOL, bool BWD>
__global__ void __launch_bounds__(PT, 1) conv_p3_kernel(P3Args a) {
  using Cfg = P3Cfg<KTOT, NOUT, POOL, BWD>;
  constexpr int NKC = Cfg::NKC, NSLICE = Cfg::NSLICE, NST = Cfg::NST, CS = Cfg::CS, SS = Cfg::SS;
  constexpr uint32_t IDESC = idesc_tf32(128, NOUT);

  extern __shared__ unsigned char smem_raw[];
  unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  unsigned char* a_hi = base;
  unsigned char* a_lo = a_hi + 344 * 128;
  unsigned char* wring = a_lo + 344 * 128;  // 88 064 = 86 * 1024: stays 1024-byte aligned
  float* stage = reinterpret_cast<float*>(wring + (size_t)NST * Cfg::SLICE_BYTES);
  unsigned long long* flags = reinterpret_cast<unsigned long long*>(stage + (size_t)NM3 * 128 * SS);
  __shared__ uint64_t bar_wfull[NST], bar_wempty[NST], bar_band_full, bar_unit_done;
  __shared__ uint32_t tmem_base_s;
  __shared__ float s_bias[BWD ? 1 : NOUT];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int Wp = a.W + 2, Hp = a.H + 2;
  const int Heff = (!BWD && POOL) ? 2 * a.Ho : a.H;
  const int npix = Hp * Wp;

  if (tid == 0) {
    for (int s = 0; s < NST; ++s) {
      mbar_init(&bar_wfull[s], 1);
      mbar_init(&bar_wempty[s], 1);
    }
    mbar_init(&bar_band_full, PW / 32);
    mbar_init(&bar_unit_done, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc<Cfg::TMEM_COLS>(&tmem_base_s);
  if (!BWD)
    for (int i = tid; i < NOUT; i += PT) s_bias[i] = __ldg(a.bias + i);
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

  if (warp == PW / 32 + 1) {
    // ================= weight warp: TMA ring =================
    int s_glob = 0;
    for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x) {
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
  } else if (warp == PW / 32) {
    // ================= MMA warp =================
    const bool leader = elect_one();
    const uint32_t a_hi_addr = smem_u32(a_hi), a_lo_addr = smem_u32(a_lo);
    int u_glob = 0, s_glob = 0, it = 0;
    for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x, ++it) {
      int b, y0, rows, nM;
      tile_geom(tile, b, y0, rows, nM);
      const uint32_t dbase = tmem + (uint32_t)(it & 1) * NM3 * Cfg::NSTRIDE;
#pragma unroll 1
      for (int kc = 0; kc < NKC; ++kc, ++u_glob) {
        mbar_wait(&bar_band_full, (uint32_t)(u_glob & 1));
        tc_fence_after();
        const int kvalid = (KTOT - 32 * kc) >= 32 ? 32 : (KTOT - 32 * kc);
#pragma unroll 1
        for (int tap = 0; tap < 9; ++tap, ++s_glob) {
          const int slot = s_glob % NST;
          mbar_wait(&bar_wfull[slot], (uint32_t)((s_glob / NST) & 1));
          tc_fence_after();
          const uint32_t w_hi = smem_u32(wring + (size_t)slot * Cfg::SLICE_BYTES), w_lo = w_hi + NOUT * 128;
          const int dy = tap / 3, dx = tap - dy * 3;
          const uint32_t row_off = (uint32_t)(dy * Wp + dx) * 128u;
#pragma unroll 1
          for (int ks = 0; ks < kvalid / 8; ++ks) {
            const uint64_t bh = desc_sw128(w_hi + ks * 32), bl = desc_sw128(w_lo + ks * 32);
            const uint64_t ah = desc_sw128(a_hi_addr + row_off + ks * 32), al = desc_sw128(a_lo_addr + row_off + ks * 32);
            const uint32_t acc0 = (kc > 0 || tap > 0 || ks > 0) ? 1u : 0u;
#pragma unroll
            for (int m = 0; m < NM3; ++m) {
              if (m < nM && leader) {
                const uint32_t dcol = dbase + m * Cfg::NSTRIDE;
                const uint64_t moff = (uint64_t)(m * ((128 * 128) >> 4));
                mma_tf32(dcol, ah + moff, bh, IDESC, acc0);
                if (a.passes == 3) {
                  mma_tf32(dcol, ah + moff, bl, IDESC, 1u);
                  mma_tf32(dcol, al + moff, bh, IDESC, 1u);
                }
              }
            }
          }
          if (leader) mma_commit(&bar_wempty[slot]);  // slot free once these MMAs have read it
          __syncwarp();
        }
        if (leader) mma_commit(&bar_unit_done);  // band free; after the last chunk the accumulator is complete
        __syncwarp();
      }
    }
  } else {
    // ================= worker warps =================
    const int c4 = tid & 7, r0 = tid >> 3;  // channel group; first band row (rows r0 + 32 u)
    float4 rv[NI_MAX];
    uchar4 rc[NI_MAX];
    unsigned rok = 0;  // bit u: item u is a real element (loads stay RAW in registers until convert)
    unsigned char rwant[NI_MAX];

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
#pragma unroll
        for (int u = 0; u < NI_MAX; ++u) {
          const int r = r0 + 32 * u;
          const int q = q_lo + r;
          bool ok = ch_ok && r < band_used && q >= 0 && q < npix;
          const int qq = ok ? q : 0;
          const int yp = fdiv(qq, a.dWp), xp = qq - yp * Wp;
          const int y = yp - 1, x = xp - 1;
          ok = ok && y >= 0 && y < a.H && x >= 0 && x < a.W;
          const int py = POOL ? (y >> 1) : y, px = POOL ? (x >> 1) : x;
          ok = ok && py < a.Ho && px < a.Wo;
          const size_t o = ok ? (((size_t)b * a.Ho + py) * a.Wo + px) * Ch + c : 0;
          rv[u] = __ldg(reinterpret_cast<const float4*>(a.gout + o));
          rc[u] = __ldg(reinterpret_cast<const uchar4*>(a.codes_in + o));
          rwant[u] = (unsigned char)(ok ? ((POOL ? (((y & 1) << 1) | (x & 1)) : 0) | (half << 2)) : 0xff);
        }
      } else {
        const float* inb = a.in + (size_t)b * npix * KTOT + ch;
#pragma unroll
        for (int u = 0; u < NI_MAX; ++u) {
          const int r = r0 + 32 * u;
          const int q = q_lo + r;
          const bool ok = ch_ok && r < band_used && q >= 0 && q < npix;
          rv[u] = __ldg(reinterpret_cast<const float4*>(ok ? inb + (size_t)q * KTOT : a.in));
          rok |= (ok ? 1u : 0u) << u;
        }
      }
    };

    auto convert_store = [&](int tile, int kc) {
      int b, y0, rows, nM;
      tile_geom(tile, b, y0, rows, nM);
      const int band_used = nM * 128 + 2 * (Wp + 1);
      float4 sc = make_float4(1.f, 1.f, 1.f, 1.f);
      if (BWD && a.bn_invstd != nullptr) {
        constexpr int Ch = KTOT / 2;
        const int ch = 32 * kc + 4 * c4;
        if (ch < KTOT) sc = __ldg(reinterpret_cast<const float4*>(a.bn_invstd + (ch >= Ch ? ch - Ch : ch)));
      }
#pragma unroll
      for (int u = 0; u < NI_MAX; ++u) {
        const int r = r0 + 32 * u;
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
          const uint32_t off = sw128_chunk(r, c4);
          *reinterpret_cast<float4*>(a_hi + off) = hi;
          *reinterpret_cast<float4*>(a_lo + off) = lo;
        }
      }
    };

    auto epilogue = [&](int tile, int buf) {
      int b, y0, rows, nM;
      tile_geom(tile, b, y0, rows, nM);
      asm volatile("bar.sync 1, %0;" ::"n"(PW) : "memory");  // previous staging tile fully consumed
      {
        const int wq = warp & 3, wg = warp >> 2;
        for (int m = wg; m < nM; m += 2) {
          const int r = m * 128 + wq * 32 + lane;
          const uint32_t taddr = tmem + ((uint32_t)(wq * 32) << 16) + (uint32_t)(buf * NM3 + m) * Cfg::NSTRIDE;
          float* srow = stage + (size_t)r * SS;
          if (BWD) {
#pragma unroll 1
            for (int c0 = 0; c0 < NOUT; c0 += 16) {
              uint32_t v[16];
              tmem_ld16_issue(taddr + c0, v);
              tmem_ld_wait();
#pragma unroll
              for (int j = 0; j < 16; j += 4)
                *reinterpret_cast<float4*>(srow + c0 + j) = make_float4(
                    __uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
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
      asm volatile("bar.sync 1, %0;" ::"n"(PW) : "memory");
      constexpr int C4 = CS / 4;
      if (BWD) {
        const int items = rows * a.W * C4;
        for (int i = tid; i < items; i += PW) {
          const int ic = i / C4, c4i = i - C4 * ic, yl = fdiv(ic, a.dW), x = ic - yl * a.W;
          const float4 v = *reinterpret_cast<const float4*>(stage + (size_t)(yl * Wp + x + 1) * SS + 4 * c4i);
          *reinterpret_cast<float4*>(a.gin + (((size_t)b * a.H + y0 + yl) * a.W + x) * NOUT + 4 * c4i) = v;
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
    while (tile < a.n_tiles) {
      if (u_glob > 0) {  // MMAs of the previous unit have finished reading the band (and, at kc = 0, tile it-1 is complete)
        mbar_wait(&bar_unit_done, (uint32_t)((u_glob - 1) & 1));
        tc_fence_after();
      }
      convert_store(tile, kc);
      fence_proxy_async();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bar_band_full);
      // next unit
      int ntile = tile, nkc = kc + 1;
      if (nkc == NKC) {
        nkc = 0;
        ntile = tile + gridDim.x;
      }
      if (ntile < a.n_tiles) issue_loads(ntile, nkc);
      if (kc == 0 && it > 0) epilogue(prev_tile, (it - 1) & 1);  // overlaps the MMAs of this tile
      if (nkc == 0) {
        prev_tile = tile;
        ++it;
      }
      tile = ntile;
      kc = nkc;
      ++u_glob;
    }
    if (it > 0) {
      mbar_wait(&bar_unit_done, (uint32_t)((u_glob - 1) & 1));
      tc_fence_after();
      epilogue(prev_tile, (it - 1) & 1);
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

template <int KTOT, int NOUT, bool POOL, bool BWD>
int launch_p3(P3Args a, const char* tag, cudaStream_t stream) {
  using Cfg = P3Cfg<KTOT, NOUT, POOL, BWD>;
  const int Wp = a.W + 2;
  const int Heff = (!BWD && POOL) ? 2 * a.Ho : a.H;
  const bool even = !BWD && POOL;
  int Rmax = (NM3 * 128) / Wp;
  if (even) Rmax &= ~1;
  ADVB_CHECK(Rmax >= (even ? 2 : 1) && Wp <= 42, "persistent 3x3 conv: image too wide for the tile");
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
  ADVB_CHECK((cdiv(R * Wp, 128) * 128 + 2 * (Wp + 1)) * 8 <= NI_MAX * PW, "persistent 3x3 conv: band exceeds the prefetch registers");
  auto kern = conv_p3_kernel<KTOT, NOUT, POOL, BWD>;
  ADVB_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM));
  int n_sm = 148, dev = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
  const int grid = a.n_tiles < n_sm ? a.n_tiles : n_sm;
  kern<<<grid, PT, Cfg::SMEM, stream>>>(a);
  ADVB_KERNEL_OK(tag, stream);
  return 0;
}

}  // namespace

bool conv_p3_supported(int Cin, int Cout, int KS, bool pool, int W) {
  if (tune_p3() == 0 || KS != 3 || W > 40) return false;
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

int conv_p3_backward(const ConvBwdArgs& g, const unsigned char* wpack, int passes, cudaStream_t stream) {
  P3Args a{};
  a.B = g.B, a.H = g.H, a.W = g.W, a.Ho = g.Ho, a.Wo = g.Wo;
  a.wpack = wpack;
  a.gout = g.gout, a.codes_in = g.codes, a.gin = g.gin, a.bn_invstd = g.bn_invstd;
  a.passes = passes;
  if (g.Cin == 32 && g.Cout == 96 && g.pool) return launch_p3<96, 32, true, true>(a, g.tag, stream);
  if (g.Cin == 48 && g.Cout == 128 && g.pool) return launch_p3<128, 48, true, true>(a, g.tag, stream);
  if (g.Cin == 32 && g.Cout == 64 && g.pool) return launch_p3<64, 32, true, true>(a, g.tag, stream);
  if (g.Cin == 64 && g.Cout == 64 && !g.pool) return launch_p3<64, 64, false, true>(a, g.tag, stream);
  set_error("conv shape has no persistent 3x3 instantiation (backward)");
  return 1;
}

}  // namespace advb
