// GEMM with fused epilogue for RawNet3's Conv1d contractions (contract: gemm.cuh).
//
// gemm_tc_kernel<NT, NSTAGE> (the product path): one CTA = one 128-row x NT-column output tile (NT = 256, or 128 when
// N = 128), fp32 accumulator in TMEM, 3xTF32 on tcgen05 (tc_common.cuh).  Roles:
//   * 8 worker warps: prefetch the next 128 x 32 fp32 chunk of A into registers (kept raw until converted, as in
//     conv_light.cu), split it into tf32 hi / lo and store it K-major SWIZZLE_128B into a ring stage; thread 0 also
//     launches the 1-D TMA bulk copy of the matching pre-packed weight slice into the same stage;
//   * warp 8: waits for the stage (8 warp arrivals + the copy's bytes on one mbarrier), issues 4 k-steps x 3 passes of
//     M=128, N=NT, K=8 MMAs and releases the stage with tcgen05.commit;
//   * epilogue (the worker warps again): tcgen05.ld 32x32b (thread = row) -> per-warp shared-memory transpose ->
//     thread = (row, 4 columns) so that every load / store of the fused epilogue (bias, ReLU + mask, BatchNorm, addend,
//     gate, second output) is a coalesced 16-byte access.
// blockIdx.x walks the N tiles fastest, so the CTAs that share a 128-row slab of A run together and the slab is read
// from HBM once and from L2 N/NT - 1 times.
//
// gemm_simt_kernel: the same contract in plain fp32 FMA (64 x 64 tiles), engine option conv_path = 1 (cross-check).
#include "gemm.cuh"
#include "tc_common.cuh"

#include <stdlib.h>

#include <algorithm>

namespace advb {

namespace {

using namespace tc;

constexpr int GW = 256;        // worker threads
constexpr int GT = GW + 32;    // + the MMA warp
constexpr int SINC_K = 251, SINC_STRIDE = 10;


__host__ __device__ constexpr int tile_n(int N) { return (N % 256 == 0) ? 256 : 128; }

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

// Per-chunk hand-offs (ring stage full / empty) are waited for with a bare try_wait loop: try_wait already suspends the
// thread in hardware for a bounded time, and the __nanosleep back-off of tc::mbar_wait (right for the once-per-tile waits of
// the convolution kernels) adds a sleep quantum to EVERY link of the fill -> MMA -> fill chain.
__device__ __forceinline__ void wait_hot(uint64_t* bar, uint32_t parity) {
  while (!mbar_try(bar, parity)) {
  }
}

// x = hi + lo with BOTH parts rounded to nearest tf32.  kind::tf32 truncates its inputs (tc_common.cuh); an fp32 `lo`
// (13 significant bits) would be truncated towards zero - a biased 2^-21 |x| error that adds up linearly over RawNet3's
// 1024..3072-long contractions (measured: 4e-6 systematic logit offset).  Rounded, the residual is unbiased and 2x smaller.
__device__ __forceinline__ void split_rn(float x, float& hi, float& lo) {
  uint32_t h, l;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(h) : "f"(x));
  hi = __uint_as_float(h);
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(l) : "f"(x - hi));
  lo = __uint_as_float(l);
}

// The fused epilogue on 4 consecutive columns n..n+3 of valid row r (clip = r / Tp).
// Per-column operands (EpiCol) are fetched once per 32-column block, not once per row: with 227 KB of the SM's 256 KB given to
// shared memory the L1 is ~28 KB and the output stores keep evicting those vectors, so every row paid 3 L2 round trips
// (measured: +12 us per 128 x 256 tile for the bias / ReLU / BatchNorm epilogue over the plain one).
// Per-row operands (EpiRow: addend, gate masks) of ROWS_AHEAD rows are fetched back to back before the first of them is
// finished: written one row at a time, the possible aliasing of `out` with `add` pins every load behind the previous row's
// stores and each row pays a full HBM latency.  (Fetching all 8 rows of a block up front needs ~100 registers; with 17 warps
// the kernel is capped at 96 and the spills landed on the fill warps' prefetch registers.)
struct EpiCol {
  float4 bias, bn_s, bn_t, gs, gs2;
};
struct EpiRow {  // `add` and `add2` are never used by the same GEMM: one register set serves both
  float4 addv;
  uchar4 gm, gm2;
};
__device__ __forceinline__ EpiCol epi_load_col(const GemmArgs& a, int n) {
  EpiCol c;
  c.bias = c.bn_s = c.bn_t = c.gs = c.gs2 = make_float4(0.f, 0.f, 0.f, 0.f);
  if (a.bias != nullptr && !a.bias_per_clip) c.bias = ldg4(a.bias + n);
  if (a.bn_scale != nullptr) c.bn_s = ldg4(a.bn_scale + n), c.bn_t = ldg4(a.bn_shift + n);
  if (a.gate_scale != nullptr) c.gs = ldg4(a.gate_scale + n);
  if (a.gate2_scale != nullptr) c.gs2 = ldg4(a.gate2_scale + n);
  return c;
}
__device__ __forceinline__ EpiRow epi_load_row(const GemmArgs& a, int r, int n) {
  EpiRow w;
  w.addv = make_float4(0.f, 0.f, 0.f, 0.f);
  w.gm = w.gm2 = make_uchar4(0, 0, 0, 0);
  if (a.add != nullptr) w.addv = *reinterpret_cast<const float4*>(a.add + (size_t)r * a.ld_add + n);  // may alias out: plain load
  else if (a.add2 != nullptr) w.addv = ldg4(a.add2 + (size_t)r * a.ld_add2 + n);
  if (a.gate_scale != nullptr) w.gm = __ldg(reinterpret_cast<const uchar4*>(a.gate_mask + (size_t)r * a.ld_gate + n));
  if (a.gate2_scale != nullptr) w.gm2 = __ldg(reinterpret_cast<const uchar4*>(a.gate2_mask + (size_t)r * a.ld_gate2 + n));
  return w;
}
__device__ __forceinline__ void epi_finish(const GemmArgs& a, int r, int clip, int n, float4 v, const EpiCol& c, const EpiRow& w) {
  if (a.bias != nullptr) {
    // (per-clip bias: the attention layer only, 128 floats per clip)
    const float4 b = a.bias_per_clip ? ldg4(a.bias + (size_t)clip * a.N + n) : c.bias;
    v.x += b.x, v.y += b.y, v.z += b.z, v.w += b.w;
  }
  if (a.relu) {
    if (a.mask_out != nullptr)
      *reinterpret_cast<uchar4*>(a.mask_out + (size_t)r * a.ld_mask + n) =
          make_uchar4(v.x > 0.f ? 1 : 0, v.y > 0.f ? 1 : 0, v.z > 0.f ? 1 : 0, v.w > 0.f ? 1 : 0);
    v.x = fmaxf(v.x, 0.f), v.y = fmaxf(v.y, 0.f), v.z = fmaxf(v.z, 0.f), v.w = fmaxf(v.w, 0.f);
  }
  if (a.bn_scale != nullptr) {
    v.x = fmaf(v.x, c.bn_s.x, c.bn_t.x), v.y = fmaf(v.y, c.bn_s.y, c.bn_t.y);
    v.z = fmaf(v.z, c.bn_s.z, c.bn_t.z), v.w = fmaf(v.w, c.bn_s.w, c.bn_t.w);
  }
  if (a.add != nullptr) v.x += w.addv.x, v.y += w.addv.y, v.z += w.addv.z, v.w += w.addv.w;
  if (a.out != nullptr) {
    float4 o = v;
    if (a.gate_scale != nullptr)
      o.x = w.gm.x ? v.x * c.gs.x : 0.f, o.y = w.gm.y ? v.y * c.gs.y : 0.f, o.z = w.gm.z ? v.z * c.gs.z : 0.f,
      o.w = w.gm.w ? v.w * c.gs.w : 0.f;
    *reinterpret_cast<float4*>(a.out + (size_t)r * a.ldc + n) = o;
  }
  if (a.out2 != nullptr) {
    float4 o = v;
    if (a.add2 != nullptr) o.x += w.addv.x, o.y += w.addv.y, o.z += w.addv.z, o.w += w.addv.w;
    if (a.gate2_scale != nullptr)
      o.x = w.gm2.x ? o.x * c.gs2.x : 0.f, o.y = w.gm2.y ? o.y * c.gs2.y : 0.f, o.z = w.gm2.z ? o.z * c.gs2.z : 0.f,
      o.w = w.gm2.w ? o.w * c.gs2.w : 0.f;
    *reinterpret_cast<float4*>(a.out2 + (size_t)r * a.ld2 + n) = o;
  }
}
__device__ __forceinline__ void epi_apply(const GemmArgs& a, int r, int clip, int n, float4 v) {  // SIMT kernel
  epi_finish(a, r, clip, n, v, epi_load_col(a, n), epi_load_row(a, r, n));
}

// Tensor-core kernels: 16 consecutive columns n0..n0+15 of row r, straight from the tcgen05.ld registers (thread = row).
// No shared-memory transpose: the phase counters (ADVB_GEMM_PROF) showed the transposing epilogue at 33-53 thousand cycles
// per 128 x 256 tile - ~130 cycles per LDS / STS, because the tensor core's operand reads own the shared-memory pipe - which
// made every GEMM with K <= 384 epilogue-bound (the MMA warp waited for a free accumulator 40 % of its time).  Here each
// thread reads / writes 64 contiguous bytes of its own row (4 x 16 B; the warp touches 32 rows per instruction, L2 merges
// the sectors), per-column vectors are warp-uniform broadcast loads, and the ReLU / gate masks move as one 16-byte word.
__device__ __forceinline__ void epi_direct16(const GemmArgs& a, int r, int clip, int n0, const uint32_t (&v)[16]) {
  if (a.n_store > 0 && n0 >= a.n_store) return;  // padding columns of a GEMM whose true N is not a multiple of 128
  float4 addv[4];
  uint4 gm = make_uint4(0, 0, 0, 0), gm2 = make_uint4(0, 0, 0, 0);
  const float* addp = a.add != nullptr ? a.add + (size_t)r * a.ld_add + n0
                                       : (a.add2 != nullptr ? a.add2 + (size_t)r * a.ld_add2 + n0 : nullptr);
  if (addp != nullptr) {
#pragma unroll
    for (int j = 0; j < 4; ++j) addv[j] = *reinterpret_cast<const float4*>(addp + 4 * j);  // `add` may alias out: plain loads
  }
  if (a.gate_scale != nullptr) gm = *reinterpret_cast<const uint4*>(a.gate_mask + (size_t)r * a.ld_gate + n0);
  if (a.gate2_scale != nullptr) gm2 = *reinterpret_cast<const uint4*>(a.gate2_mask + (size_t)r * a.ld_gate2 + n0);
  const uint32_t gmw[4] = {gm.x, gm.y, gm.z, gm.w}, gm2w[4] = {gm2.x, gm2.y, gm2.z, gm2.w};
  uint32_t mb[4] = {0u, 0u, 0u, 0u};
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int n = n0 + 4 * j;
    float4 x = make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]), __uint_as_float(v[4 * j + 2]),
                           __uint_as_float(v[4 * j + 3]));
    if (a.bias != nullptr) {
      const float4 b = ldg4(a.bias + (a.bias_per_clip ? (size_t)clip * a.N : (size_t)0) + n);
      x.x += b.x, x.y += b.y, x.z += b.z, x.w += b.w;
    }
    if (a.relu) {
      mb[j] = (x.x > 0.f ? 1u : 0u) | (x.y > 0.f ? 0x100u : 0u) | (x.z > 0.f ? 0x10000u : 0u) | (x.w > 0.f ? 0x1000000u : 0u);
      x.x = fmaxf(x.x, 0.f), x.y = fmaxf(x.y, 0.f), x.z = fmaxf(x.z, 0.f), x.w = fmaxf(x.w, 0.f);
    }
    if (a.bn_scale != nullptr) {
      const float4 sc = ldg4(a.bn_scale + n), sh = ldg4(a.bn_shift + n);
      x.x = fmaf(x.x, sc.x, sh.x), x.y = fmaf(x.y, sc.y, sh.y), x.z = fmaf(x.z, sc.z, sh.z), x.w = fmaf(x.w, sc.w, sh.w);
    }
    if (a.add != nullptr) x.x += addv[j].x, x.y += addv[j].y, x.z += addv[j].z, x.w += addv[j].w;
    if (a.out != nullptr) {
      float4 o = x;
      if (a.gate_scale != nullptr) {
        const float4 sc = ldg4(a.gate_scale + n);
        const uint32_t m = gmw[j];
        o.x = (m & 0xffu) ? x.x * sc.x : 0.f, o.y = (m & 0xff00u) ? x.y * sc.y : 0.f;
        o.z = (m & 0xff0000u) ? x.z * sc.z : 0.f, o.w = (m & 0xff000000u) ? x.w * sc.w : 0.f;
      }
      *reinterpret_cast<float4*>(a.out + (size_t)r * a.ldc + n) = o;
    }
    if (a.out2 != nullptr) {
      float4 o = x;
      if (a.add2 != nullptr) o.x += addv[j].x, o.y += addv[j].y, o.z += addv[j].z, o.w += addv[j].w;
      if (a.gate2_scale != nullptr) {
        const float4 sc = ldg4(a.gate2_scale + n);
        const uint32_t m = gm2w[j];
        o.x = (m & 0xffu) ? o.x * sc.x : 0.f, o.y = (m & 0xff00u) ? o.y * sc.y : 0.f;
        o.z = (m & 0xff0000u) ? o.z * sc.z : 0.f, o.w = (m & 0xff000000u) ? o.w * sc.w : 0.f;
      }
      *reinterpret_cast<float4*>(a.out2 + (size_t)r * a.ld2 + n) = o;
    }
  }
  if (a.relu && a.mask_out != nullptr)
    *reinterpret_cast<uint4*>(a.mask_out + (size_t)r * a.ld_mask + n0) = make_uint4(mb[0], mb[1], mb[2], mb[3]);
}

__device__ __forceinline__ bool row_valid(const GemmArgs& a, int r, int& clip) {
  clip = r / a.Tp;
  const int tp = r - clip * a.Tp;
  return r < a.M && tp >= a.pad && tp < a.pad + a.Tv;
}

// ---------------------------------------------------------------------------------------------------------------
template <int NT, int NSTAGE>
__global__ void __launch_bounds__(GT, 1) gemm_tc_kernel(const GemmArgs a, const int passes) {
  constexpr int A_BYTES = 128 * 128;                 // one tf32 part of a 128 x 32 chunk
  constexpr int W_BYTES = NT * 128;                  // one tf32 part of an NT x 32 weight slice
  constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * W_BYTES;
  constexpr uint32_t IDESC = idesc_tf32(128, NT);

  extern __shared__ unsigned char smem_raw[];
  unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar_full[NSTAGE], bar_empty[NSTAGE], bar_done;
  __shared__ uint32_t tmem_base_s;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int ntile = blockIdx.x, m0 = blockIdx.y * 128;
  const int kchunks = a.K >> 5, NKC = a.ntap * kchunks;

  if (tid == 0) {
    for (int s = 0; s < NSTAGE; ++s) {
      mbar_init(&bar_full[s], GW / 32 + 1);  // 8 worker warps + the expect_tx arrival of the weight copy
      mbar_init(&bar_empty[s], 1);
    }
    mbar_init(&bar_done, 1);
    fence_barrier_init();
  }
  if (warp == GW / 32) tmem_alloc<NT>(&tmem_base_s);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;

  if (warp == GW / 32) {
    // ---- MMA warp ----
    const bool leader = elect_one();
#pragma unroll 1
    for (int kc = 0; kc < NKC; ++kc) {
      const int s = kc % NSTAGE;
      wait_hot(&bar_full[s], (uint32_t)((kc / NSTAGE) & 1));
      tc_fence_after();
      const uint32_t a_hi = smem_u32(base + (size_t)s * STAGE_BYTES), a_lo = a_hi + A_BYTES;
      const uint32_t w_hi = a_lo + A_BYTES, w_lo = w_hi + W_BYTES;
      if (leader) {
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          const uint64_t ah = desc_sw128(a_hi + ks * 32), al = desc_sw128(a_lo + ks * 32);
          const uint64_t bh = desc_sw128(w_hi + ks * 32), bl = desc_sw128(w_lo + ks * 32);
          mma_tf32(tmem, ah, bh, IDESC, (kc > 0 || ks > 0) ? 1u : 0u);
          if (passes == 3) {
            mma_tf32(tmem, ah, bl, IDESC, 1u);
            mma_tf32(tmem, al, bh, IDESC, 1u);
          }
        }
        mma_commit(&bar_empty[s]);                    // the stage may be refilled once these MMAs have read it
        if (kc == NKC - 1) mma_commit(&bar_done);     // accumulator complete
      }
      __syncwarp();
    }
  } else {
    // ---- worker warps: A chunk -> registers -> tf32 hi / lo -> stage ----
    const int c4 = tid & 7, r0 = tid >> 3;  // this thread's 16-byte column group and first row (rows r0 + 32 u)
    // Register prefetch PF chunks ahead: one chunk's MMAs take 0.4-0.6 us, a global load ~0.8 us, so a distance of one
    // chunk left the loop latency-bound (first version: 19 us per 128 x 128, K = 384 tile against a 4.8 us MMA floor).
    constexpr int PF = 4;
    float4 rv[PF][4];
    unsigned rok[PF];
    auto issue_loads = [&](int kc, float4 (&dst)[4], unsigned& okm) {
      const int tap = kc / kchunks, kk = kc - tap * kchunks;
      okm = 0;
      if (a.im2col_T > 0) {
        const int k = kk * 32 + c4 * 4;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int r = m0 + r0 + 32 * u;
          const int rr = r < a.M ? r : 0;
          const int clip = rr / a.Tp, l = rr - clip * a.Tp;
          const long long idx = (long long)SINC_STRIDE * l + k;
          const float* src = a.A + (size_t)clip * a.im2col_T;
          float t[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const bool ok = r < a.M && (k + j) < SINC_K && (idx + j) < a.im2col_T;
            t[j] = ok ? __ldg(src + idx + j) : 0.f;  // (scalar gather: the sinc layer is 2 % of the model's MACs)
          }
          dst[u] = make_float4(t[0], t[1], t[2], t[3]);
          okm |= 1u << u;
        }
      } else {
        const int sh = a.shift[tap];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int r = m0 + r0 + 32 * u + sh;
          const bool ok = r >= 0 && r < a.M;
          dst[u] = ldg4(a.A + (size_t)(ok ? r : 0) * a.lda + kk * 32 + c4 * 4);
          okm |= (ok ? 1u : 0u) << u;
        }
      }
    };
#pragma unroll
    for (int d = 0; d < PF; ++d)
      if (d < NKC) issue_loads(d, rv[d], rok[d]);
#pragma unroll 1
    for (int kc0 = 0; kc0 < NKC; kc0 += PF) {
#pragma unroll
      for (int d = 0; d < PF; ++d) {
        const int kc = kc0 + d;
        if (kc < NKC) {
          const int s = kc % NSTAGE, use = kc / NSTAGE;
          if (use > 0) {
            wait_hot(&bar_empty[s], (uint32_t)((use - 1) & 1));
            tc_fence_after();
          }
          unsigned char* st = base + (size_t)s * STAGE_BYTES;
          if (tid == 0) {
            mbar_expect_tx(&bar_full[s], 2 * W_BYTES);
            bulk_g2s(st + 2 * A_BYTES, a.wpack + ((size_t)ntile * NKC + kc) * (2 * W_BYTES), 2 * W_BYTES, &bar_full[s]);
          }
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const bool ok = (rok[d] >> u) & 1u;
            const float4 v = ok ? rv[d][u] : make_float4(0.f, 0.f, 0.f, 0.f);
            float4 hi, lo;
            split_rn(v.x, hi.x, lo.x);
            split_rn(v.y, hi.y, lo.y);
            split_rn(v.z, hi.z, lo.z);
            split_rn(v.w, hi.w, lo.w);
            const uint32_t off = sw128_chunk(r0 + 32 * u, c4);
            *reinterpret_cast<float4*>(st + off) = hi;
            *reinterpret_cast<float4*>(st + A_BYTES + off) = lo;
          }
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) mbar_arrive(&bar_full[s]);
          if (kc + PF < NKC) issue_loads(kc + PF, rv[d], rok[d]);
        }
      }
    }

    // ---- epilogue ----
    mbar_wait(&bar_done, 0u);
    tc_fence_after();
    const int q = warp & 3, half = warp >> 2;
    const int r = m0 + 32 * q + lane;  // thread = row
    int clip;
    const bool valid = row_valid(a, r, clip);
    const uint32_t taddr = tmem + ((uint32_t)(32 * q) << 16);
#pragma unroll 1
    for (int cb = 0; cb < NT / 64; ++cb) {
      const int col0 = half * (NT / 2) + cb * 32;
      uint32_t v0[16], v1[16];
      tmem_ld16_issue(taddr + col0, v0);
      tmem_ld16_issue(taddr + col0 + 16, v1);
      tmem_ld_wait();
      if (valid) {
        epi_direct16(a, r, clip, ntile * NT + col0, v0);
        epi_direct16(a, r, clip, ntile * NT + col0 + 16, v1);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == GW / 32) tmem_dealloc<NT>(tmem);
}

// ---------------------------------------------------------------------------------------------------------------
// Persistent, warp-specialised version (the default schedule): one CTA per SM walks the output tiles (N tile fastest),
// the stage ring and the register prefetch run straight across tile boundaries, and two TMEM accumulators let the MMAs
// of tile i+1 overlap the epilogue of tile i.  Roles: warps 0-7 fill A (+ thread 0 the weight TMA), warp 8 issues the
// MMAs, warps 9-12 run the epilogue (one per TMEM lane quarter, all of the tile's columns).
// Hand-offs: bar_full / bar_empty per ring stage as above; acc_full[2] (tcgen05.commit after a tile's last MMA ->
// epilogue warps), acc_empty[2] (4 epilogue warps, as soon as their last tcgen05.ld of the tile has landed -> MMA warp).
// ---- thread-block cluster helpers (weight multicast) ----
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// 1-D TMA bulk copy global -> the SAME shared-memory offset of every CTA in `mask`; each destination CTA's mbarrier (same
// offset) receives complete_tx for the bytes it got.
__device__ __forceinline__ void bulk_g2s_multicast(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar, uint16_t mask) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;" ::"r"(
          smem_u32(dst_smem)),
      "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)), "h"(mask)
      : "memory");
}
// tcgen05.commit that arrives on the mbarrier at this offset in every CTA of `mask`
__device__ __forceinline__ void mma_commit_multicast(uint64_t* bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                   smem_u32(bar)),
               "h"(mask)
               : "memory");
}

// CL = cluster size along M (1, 2 or 4).  The CTAs of a cluster work on CL consecutive M tiles of the SAME N tile in
// lock-step, so every weight slice is fetched from L2 once per cluster: CTA r loads the r-th 1/CL of the slice and
// multicasts it into all CL shared memories.  Without it the kernel is bound by L2 -> SM traffic, not by the tensor pipe:
// a 128 x 256 tile streams 64 KB of weights + 16 KB of activations per 32-wide K chunk, 8.5 TB/s over 148 SMs at the
// measured 1.4 us per chunk (MMA time per chunk: 0.6 us).  A ring stage is released only when all CL CTAs have consumed
// it (multicast tcgen05.commit onto every CTA's bar_empty, count CL).
// EPW = epilogue warps: 4 (one per TMEM lane quarter; 13 warps => 128 registers per thread) or 8 (two per quarter, each half of
// the columns; 17 warps put 5 on one SM sub-partition and cap every thread at 96 registers).
template <int NT, int NSTAGE, int EPW, int CL>
__global__ void __launch_bounds__(GW + 32 + 32 * EPW, 1) gemm_tc_persistent_kernel(const GemmArgs a, const int passes, const int n_items) {
  // work item w = (M group of CL tiles, N tile), N tile fastest; cluster c takes items c, c + n_clusters, ...
  constexpr int A_BYTES = 128 * 128;
  constexpr int W_BYTES = NT * 128;
  constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * W_BYTES;
  constexpr int TM_COLS = 2 * NT;  // two accumulators
  constexpr uint32_t IDESC = idesc_tf32(128, NT);

  extern __shared__ unsigned char smem_raw[];
  unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar_full[NSTAGE], bar_empty[NSTAGE], acc_full[2], acc_empty[2];
  __shared__ uint32_t tmem_base_s;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int ntn = a.N / NT;
  const int kchunks = a.K >> 5, NKC = a.ntap * kchunks;
  const int crank = CL > 1 ? (int)cluster_ctarank() : 0;
  const int item0 = blockIdx.x / CL, item_step = gridDim.x / CL;
  constexpr uint16_t CMASK = (uint16_t)((1u << CL) - 1u);

  if (tid == 0) {
    for (int s = 0; s < NSTAGE; ++s) {
      mbar_init(&bar_full[s], GW / 32 + 1);
      mbar_init(&bar_empty[s], CL);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&acc_full[i], 1);
      mbar_init(&acc_empty[i], EPW);
    }
    fence_barrier_init();
  }
  if (warp == GW / 32) tmem_alloc<TM_COLS>(&tmem_base_s);
  tc_fence_before();
  __syncthreads();
  if (CL > 1) cluster_sync_all();  // every CTA's barriers are initialised before any peer signals them
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;

  if (warp == GW / 32) {
    // ---- MMA warp ----
    const bool leader = elect_one();
    int g = 0, it = 0;
    const bool prof = a.prof != nullptr && blockIdx.x == 0;
    long long pc_acc = 0, pc_full = 0, pc_issue = 0, t_last = clock64();
#ifdef ADVB_GEMM_PROFILE  // diagnostic build only (python build.py with NVCC_EXTRA=-DADVB_GEMM_PROFILE): the counters cost registers
#define GPROF(acc_)                    \
  if (prof) {                          \
    const long long now_ = clock64();  \
    acc_ += now_ - t_last;             \
    t_last = now_;                     \
  }
#else
#define GPROF(acc_)
#endif
#pragma unroll 1
    for (int t = item0; t < n_items; t += item_step, ++it) {
      const int buf = it & 1;
      if (it >= 2) {  // the epilogue of the tile that used this accumulator two tiles ago has drained it
        mbar_wait(&acc_empty[buf], (uint32_t)(((it >> 1) - 1) & 1));
        tc_fence_after();
      }
      GPROF(pc_acc)
      const uint32_t dcol = tmem + buf * NT;
#pragma unroll 1
      for (int kc = 0; kc < NKC; ++kc, ++g) {
        const int s = g % NSTAGE;
        wait_hot(&bar_full[s], (uint32_t)((g / NSTAGE) & 1));
        tc_fence_after();
        GPROF(pc_full)
        const uint32_t a_hi = smem_u32(base + (size_t)s * STAGE_BYTES), a_lo = a_hi + A_BYTES;
        const uint32_t w_hi = a_lo + A_BYTES, w_lo = w_hi + W_BYTES;
        if (leader) {
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            const uint64_t ah = desc_sw128(a_hi + ks * 32), al = desc_sw128(a_lo + ks * 32);
            const uint64_t bh = desc_sw128(w_hi + ks * 32), bl = desc_sw128(w_lo + ks * 32);
            mma_tf32(dcol, ah, bh, IDESC, (kc > 0 || ks > 0) ? 1u : 0u);
            if (passes == 3) {
              mma_tf32(dcol, ah, bl, IDESC, 1u);
              mma_tf32(dcol, al, bh, IDESC, 1u);
            }
          }
          if (CL > 1) mma_commit_multicast(&bar_empty[s], CMASK);
          else mma_commit(&bar_empty[s]);
          if (kc == NKC - 1) mma_commit(&acc_full[buf]);
        }
        __syncwarp();
        GPROF(pc_issue)
      }
    }
    if (prof && lane == 0) a.prof[0] = pc_acc, a.prof[1] = pc_full, a.prof[2] = pc_issue, a.prof[3] = it;
  } else if (warp < GW / 32) {
    // ---- fill warps ----
    const int c4 = tid & 7, r0 = tid >> 3;
    constexpr int PF = 4;
    float4 rv[PF][4];
    unsigned rok[PF];
    auto issue_loads = [&](int m0, int kc, float4 (&dst)[4], unsigned& okm) {
      const int tap = kc / kchunks, kk = kc - tap * kchunks;
      okm = 0;
      if (a.im2col_T > 0) {
        const int k = kk * 32 + c4 * 4;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int r = m0 + r0 + 32 * u;
          const int rr = r < a.M ? r : 0;
          const int clip = rr / a.Tp, l = rr - clip * a.Tp;
          const long long idx = (long long)SINC_STRIDE * l + k;
          const float* src = a.A + (size_t)clip * a.im2col_T;
          float tt[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const bool ok = r < a.M && (k + j) < SINC_K && (idx + j) < a.im2col_T;
            tt[j] = ok ? __ldg(src + idx + j) : 0.f;
          }
          dst[u] = make_float4(tt[0], tt[1], tt[2], tt[3]);
          okm |= 1u << u;
        }
      } else {
        const int sh = a.shift[tap];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int r = m0 + r0 + 32 * u + sh;
          const bool ok = r >= 0 && r < a.M;
          dst[u] = ldg4(a.A + (size_t)(ok ? r : 0) * a.lda + kk * 32 + c4 * 4);
          okm |= (ok ? 1u : 0u) << u;
        }
      }
    };
    // prefetch cursor (tile, chunk) runs PF chunks ahead of the store cursor, across tile boundaries
    int pt = item0, pk = 0;
    auto prefetch_next = [&](float4 (&dst)[4], unsigned& okm) {
      if (pt < n_items) {
        issue_loads(((pt / ntn) * CL + crank) * 128, pk, dst, okm);
        if (++pk == NKC) pk = 0, pt += item_step;
      }
    };
#pragma unroll
    for (int d = 0; d < PF; ++d) prefetch_next(rv[d], rok[d]);
    int t = item0, kc = 0, g = 0;
    const bool prof = a.prof != nullptr && blockIdx.x == 0 && warp == 0;
    long long pc_empty = 0, pc_work = 0, pc_fence = 0, pc_pref = 0, t_last = clock64();
#pragma unroll 1
    while (t < n_items) {
#pragma unroll
      for (int d = 0; d < PF; ++d) {
        if (t < n_items) {
          const int s = g % NSTAGE, use = g / NSTAGE;
          GPROF(pc_work)
          if (use > 0) {
            wait_hot(&bar_empty[s], (uint32_t)((use - 1) & 1));
            tc_fence_after();
          }
          GPROF(pc_empty)
          unsigned char* st = base + (size_t)s * STAGE_BYTES;
          if (tid == 0) {
            mbar_expect_tx(&bar_full[s], 2 * W_BYTES);
            const unsigned char* src = a.wpack + ((size_t)(t % ntn) * NKC + kc) * (2 * W_BYTES);
            if (CL > 1) {
              constexpr uint32_t PART = 2 * W_BYTES / CL;
              bulk_g2s_multicast(st + 2 * A_BYTES + crank * PART, src + crank * PART, PART, &bar_full[s], CMASK);
            } else {
              bulk_g2s(st + 2 * A_BYTES, src, 2 * W_BYTES, &bar_full[s]);
            }
          }
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const bool ok = (rok[d] >> u) & 1u;
            const float4 v = ok ? rv[d][u] : make_float4(0.f, 0.f, 0.f, 0.f);
            float4 hi, lo;
            split_rn(v.x, hi.x, lo.x);
            split_rn(v.y, hi.y, lo.y);
            split_rn(v.z, hi.z, lo.z);
            split_rn(v.w, hi.w, lo.w);
            const uint32_t off = sw128_chunk(r0 + 32 * u, c4);
            *reinterpret_cast<float4*>(st + off) = hi;
            *reinterpret_cast<float4*>(st + A_BYTES + off) = lo;
          }
          GPROF(pc_work)
          fence_proxy_async();
          GPROF(pc_fence)
          __syncwarp();
          if (lane == 0) mbar_arrive(&bar_full[s]);
          prefetch_next(rv[d], rok[d]);
          GPROF(pc_pref)
          ++g;
          if (++kc == NKC) kc = 0, t += item_step;
        }
      }
    }
    if (prof && lane == 0) a.prof[4] = pc_empty, a.prof[5] = pc_work, a.prof[9] = pc_fence, a.prof[10] = pc_pref;
  } else {
    // ---- epilogue warps ----
    const int ew = warp - (GW / 32 + 1);      // 0..EPW-1
    const int q = warp & 3, half = ew >> 2;   // TMEM lane quarter is fixed by the hardware warp id
    constexpr int NBLK = NT / 32 / (EPW / 4);  // 32-column blocks per warp
    int it = 0;
    const bool prof = a.prof != nullptr && blockIdx.x == 0 && ew == 0;
    long long pc_wait = 0, pc_tmem = 0, pc_store = 0, t_last = clock64();
#pragma unroll 1
    for (int t = item0; t < n_items; t += item_step, ++it) {
      const int buf = it & 1;
      const int m0 = ((t / ntn) * CL + crank) * 128, ntile = t % ntn;
      const int r = m0 + 32 * q + lane;  // thread = row
      int clip;
      const bool valid = row_valid(a, r, clip);
      GPROF(pc_store)
      mbar_wait(&acc_full[buf], (uint32_t)((it >> 1) & 1));
      tc_fence_after();
      GPROF(pc_wait)
      const uint32_t taddr = tmem + ((uint32_t)(32 * q) << 16) + buf * NT;
#pragma unroll 1
      for (int cb = 0; cb < NBLK; ++cb) {
        const int col0 = (half * NBLK + cb) * 32;
        uint32_t v0[16], v1[16];
        GPROF(pc_store)
        tmem_ld16_issue(taddr + col0, v0);
        tmem_ld16_issue(taddr + col0 + 16, v1);
        tmem_ld_wait();
        GPROF(pc_tmem)
        if (cb == NBLK - 1) {  // accumulator drained: hand it back before the global stores of this block
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&acc_empty[buf]);
        }
        if (valid) {
          epi_direct16(a, r, clip, ntile * NT + col0, v0);
          epi_direct16(a, r, clip, ntile * NT + col0 + 16, v1);
        }
      }
    }
    GPROF(pc_store)
    if (prof && lane == 0) a.prof[6] = pc_wait, a.prof[7] = pc_tmem, a.prof[8] = pc_store;
  }
#undef GPROF
  tc_fence_before();
  __syncthreads();
  if (CL > 1) cluster_sync_all();  // no CTA leaves while a peer may still signal its barriers
  if (warp == GW / 32) tmem_dealloc<TM_COLS>(tmem);
}

// ---------------------------------------------------------------------------------------------------------------
// Weight image: for N-tile j and chunk kc = tap * (K/32) + kk: [hi: NT rows x 128 B, SWIZZLE_128B][lo: same].
__global__ void gemm_pack_kernel(GemmW w, int N, int K, int ntap, unsigned char* __restrict__ dst) {
  const int NT = tile_n(N), kchunks = K >> 5, NKC = ntap * kchunks;
  const long long total = (long long)N * NKC * 32;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int k = (int)(i & 31);
    const int nl = (int)((i >> 5) % NT);
    const long long sl = i / (32LL * NT);  // slice index = ntile * NKC + kc
    const int kc = (int)(sl % NKC), ntile = (int)(sl / NKC);
    const int tap = kc / kchunks, kk = kc - tap * kchunks;
    const int n = ntile * NT + nl, kg = kk * 32 + k;
    float v = 0.f;
    if (n < w.n_valid && kg < w.k_valid) v = w.w[(long long)n * w.s_n + (long long)tap * w.s_tap + (long long)kg * w.s_k];
    float hi, lo;
    split_rn(v, hi, lo);
    unsigned char* slice = dst + (size_t)sl * (2 * (size_t)NT * 128);
    const uint32_t off = (uint32_t)(nl * 128 + ((((k >> 2) ^ (nl & 7)) << 4) | ((k & 3) << 2)));
    *reinterpret_cast<float*>(slice + off) = hi;
    *reinterpret_cast<float*>(slice + (size_t)NT * 128 + off) = lo;
  }
}

// ---------------------------------------------------------------------------------------------------------------
// fp32 SIMT cross-check: 64 x 64 tile, 256 threads, 4 x 4 outputs per thread, K in steps of 16.
__global__ void __launch_bounds__(256) gemm_simt_kernel(const GemmArgs a) {
  __shared__ float As[16][64 + 4];
  __shared__ float Ws[16][64 + 4];
  const int tid = threadIdx.x;
  const int n0 = blockIdx.x * 64, m0 = blockIdx.y * 64;
  const int tx = tid & 15, ty = tid >> 4;  // outputs: rows m0 + 4 ty + i, columns n0 + 4 tx + j
  float acc[4][4] = {};
  const int lr = tid >> 2, lk = (tid & 3) * 4;  // loader: row / column lr of the tile, k offset lk..lk+3
  for (int tap = 0; tap < a.ntap; ++tap) {
    for (int k0 = 0; k0 < a.K; k0 += 16) {
      {
        float t[4] = {0.f, 0.f, 0.f, 0.f};
        if (a.im2col_T > 0) {
          const int r = m0 + lr;
          if (r < a.M) {
            const int clip = r / a.Tp, l = r - clip * a.Tp;
            for (int j = 0; j < 4; ++j) {
              const int k = k0 + lk + j;
              const long long idx = (long long)SINC_STRIDE * l + k;
              if (k < SINC_K && idx < a.im2col_T) t[j] = a.A[(size_t)clip * a.im2col_T + idx];
            }
          }
        } else {
          const int r = m0 + lr + a.shift[tap];
          if (r >= 0 && r < a.M && (m0 + lr) < a.M) {
            const float4 v = ldg4(a.A + (size_t)r * a.lda + k0 + lk);
            t[0] = v.x, t[1] = v.y, t[2] = v.z, t[3] = v.w;
          }
        }
        for (int j = 0; j < 4; ++j) As[lk + j][lr] = t[j];
        const int n = n0 + lr;
        for (int j = 0; j < 4; ++j) {
          const int k = k0 + lk + j;
          float wv = 0.f;
          if (n < a.w.n_valid && k < a.w.k_valid)
            wv = a.w.w[(long long)n * a.w.s_n + (long long)tap * a.w.s_tap + (long long)k * a.w.s_k];
          Ws[lk + j][lr] = wv;
        }
      }
      __syncthreads();
#pragma unroll
      for (int k = 0; k < 16; ++k) {
        const float4 av = *reinterpret_cast<const float4*>(&As[k][4 * ty]);
        const float4 wv = *reinterpret_cast<const float4*>(&Ws[k][4 * tx]);
        const float ar[4] = {av.x, av.y, av.z, av.w}, wr[4] = {wv.x, wv.y, wv.z, wv.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(ar[i], wr[j], acc[i][j]);
      }
      __syncthreads();
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = m0 + 4 * ty + i;
    int clip;
    if (row_valid(a, r, clip)) epi_apply(a, r, clip, n0 + 4 * tx, make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]));
  }
}

template <int NT, int NSTAGE>
int launch_tc(const GemmArgs& a, int passes, cudaStream_t stream) {
  constexpr size_t smem = (size_t)NSTAGE * (2 * 128 * 128 + 2 * NT * 128) + 1024;
  ADVB_CUDA_OK(cudaFuncSetAttribute(gemm_tc_kernel<NT, NSTAGE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid(a.N / NT, cdiv(a.M, 128));
  gemm_tc_kernel<NT, NSTAGE><<<grid, GT, smem, stream>>>(a, passes);
  ADVB_KERNEL_OK(a.tag, stream);
  return 0;
}

int tune_env(const char* name, int dflt) {
  const char* e = getenv(name);
  return e != nullptr ? atoi(e) : dflt;
}
int tune_epw() {
  static const int v = tune_env("ADVB_GEMM_EPW", 8) == 4 ? 4 : 8;
  return v;
}
int tune_cluster() {  // cluster size along M for the weight multicast (1 = off)
  static const int v = [] {
    // measured on PGDL2-10, B = 16 with the direct epilogue: 145.8 / 155.9 / 147.4 clips/s for cluster sizes 1 / 2 / 4
    // (with the older transposing epilogue the kernel was epilogue-bound and the multicast was worth only +1.4 %)
    const int c = tune_env("ADVB_GEMM_CLUSTER", 2);
    return (c == 1 || c == 2 || c == 4) ? c : 2;
  }();
  return v;
}

template <int NT, int NSTAGE, int EPW, int CL>
int launch_tc_persistent(const GemmArgs& a, int passes, cudaStream_t stream) {
  constexpr size_t smem = (size_t)NSTAGE * (2 * 128 * 128 + 2 * NT * 128) + 1024;
  constexpr int PT = GW + 32 + 32 * EPW;
  auto kern = gemm_tc_persistent_kernel<NT, NSTAGE, EPW, CL>;
  ADVB_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaLaunchConfig_t cfg{};
  cfg.blockDim = dim3(PT);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CL, attr[0].val.clusterDim.y = 1, attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  static int max_clusters = 0;  // co-resident clusters of this shape (persistent kernel: never launch more)
  if (max_clusters == 0) {
    cfg.gridDim = dim3(CL);
    int n = 0;
    if (CL > 1) {
      ADVB_CUDA_OK(cudaOccupancyMaxActiveClusters(&n, kern, &cfg));
    } else {
      int dev = 0;
      ADVB_CUDA_OK(cudaGetDevice(&dev));
      ADVB_CUDA_OK(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev));
    }
    ADVB_CHECK(n > 0, "no co-resident cluster of the persistent GEMM fits on this device");
    max_clusters = n;
  }
  const int n_items = (a.N / NT) * cdiv(cdiv(a.M, 128), CL);
  cfg.gridDim = dim3(std::min(n_items, max_clusters) * CL);
  GemmArgs ac = a;
  static const bool prof_on = tune_env("ADVB_GEMM_PROF", 0) != 0;
  static long long* prof_dev = nullptr;
  if (prof_on) {
    if (prof_dev == nullptr) ADVB_CUDA_OK(cudaMalloc(&prof_dev, 16 * sizeof(long long)));
    ADVB_CUDA_OK(cudaMemsetAsync(prof_dev, 0, 16 * sizeof(long long), stream));
    ac.prof = prof_dev;
  }
  int p = passes, ni = n_items;
  void* params[3] = {&ac, &p, &ni};
  ADVB_CUDA_OK(cudaLaunchKernelExC(&cfg, reinterpret_cast<const void*>(kern), params));
  ADVB_KERNEL_OK(a.tag, stream);
  if (prof_on) {  // diagnostics only: synchronous read-back of CTA 0's phase counters (cycles)
    long long hcnt[16];
    ADVB_CUDA_OK(cudaMemcpyAsync(hcnt, prof_dev, sizeof(hcnt), cudaMemcpyDeviceToHost, stream));
    ADVB_CUDA_OK(cudaStreamSynchronize(stream));
    fprintf(stderr,
            "GEMMPROF %s M=%d N=%d K=%d taps=%d NT=%d tiles(cta0)=%lld | mma: acc_wait %lld full_wait %lld issue %lld | fill: "
            "empty_wait %lld convert+sts %lld fence %lld arrive+prefetch %lld | epi: full_wait %lld tmem %lld store %lld\n",
            a.tag, a.M, a.N, a.K, a.ntap, NT, hcnt[3], hcnt[0], hcnt[1], hcnt[2], hcnt[4], hcnt[5], hcnt[9], hcnt[10], hcnt[6], hcnt[7],
            hcnt[8]);
  }
  return 0;
}

template <int NT, int NSTAGE>
int launch_tc_persistent_tuned(const GemmArgs& a, int passes, cudaStream_t stream) {
  const int cl = tune_cluster();
  if (tune_epw() == 4) {
    if (cl == 4) return launch_tc_persistent<NT, NSTAGE, 4, 4>(a, passes, stream);
    if (cl == 2) return launch_tc_persistent<NT, NSTAGE, 4, 2>(a, passes, stream);
    return launch_tc_persistent<NT, NSTAGE, 4, 1>(a, passes, stream);
  }
  if (cl == 4) return launch_tc_persistent<NT, NSTAGE, 8, 4>(a, passes, stream);
  if (cl == 2) return launch_tc_persistent<NT, NSTAGE, 8, 2>(a, passes, stream);
  return launch_tc_persistent<NT, NSTAGE, 8, 1>(a, passes, stream);
}

}  // namespace

size_t gemm_pack_bytes(int N, int K, int ntap) { return (size_t)N * ntap * K * 8; }

int gemm_pack(const GemmW& w, int N, int K, int ntap, unsigned char* dst, cudaStream_t stream) {
  ADVB_CHECK(N % 128 == 0 && K % 32 == 0 && ntap >= 1 && ntap <= 3, "unsupported GEMM weight shape");
  const long long total = (long long)N * ntap * K;
  gemm_pack_kernel<<<(int)std::min<long long>(cdiv64(total, 256), 148 * 16), 256, 0, stream>>>(w, N, K, ntap, dst);
  ADVB_KERNEL_OK("gemm_pack", stream);
  return 0;
}

int gemm_run(const GemmArgs& a, int path, int passes, cudaStream_t stream) {
  ADVB_CHECK(a.N % 128 == 0 && a.K % 32 == 0 && a.ntap >= 1 && a.ntap <= 3 && a.M > 0, "unsupported GEMM shape");
  ADVB_CHECK(a.add == nullptr || a.add2 == nullptr, "add and add2 are mutually exclusive");
  ADVB_CHECK(a.im2col_T > 0 || (a.lda % 4 == 0 && (reinterpret_cast<uintptr_t>(a.A) & 15) == 0), "A operand must be 16-byte aligned");
  if (path == 1) {
    dim3 grid(a.N / 64, cdiv(a.M, 64));
    gemm_simt_kernel<<<grid, 256, 0, stream>>>(a);
    ADVB_KERNEL_OK(a.tag, stream);
    return 0;
  }
  ADVB_CHECK(a.wpack != nullptr, "tcgen05 GEMM needs the packed weight image");
  if (g_conv_sched == 1) {  // one tile per CTA: the first tcgen05 version, kept as an in-process cross-check
    if (tile_n(a.N) == 256) return launch_tc<256, 2>(a, passes, stream);
    return launch_tc<128, 3>(a, passes, stream);
  }
  if (tile_n(a.N) == 256) return launch_tc_persistent_tuned<256, 2>(a, passes, stream);
  return launch_tc_persistent_tuned<128, 3>(a, passes, stream);
}

}  // namespace advb
