// LCNN first block forward (Conv2d(1 -> 64, 5x5) + Max-Feature-Map + MaxPool 2x2, src/models/lcnn.py:121-126) as a
// TOEPLITZ GEMM on tcgen05 with NO im2col (sm_100a).
//
// The im2col version (conv_light.cu, K = 25 taps padded to 32) was issue-bound on its worker warps: every thread gathered
// 25 taps per pixel, split them into tf32 hi/lo and stored them swizzled, then pooled through a staging tile - 675
// instructions per thread per 128-pixel tile, 0.30 ms per launch against an MMA floor of 0.07 ms.
//
// Here the A operand is the image itself.  The cepstral image of a clip is transposed into shared memory as rows of
// consecutive FRAMES (one row per coefficient), and GEMM row r = (coefficient i, frame unit jj) is the 8 consecutive frames
// starting at frame 4 jj of coefficient row i + dc:
//     A_dc[r][k]      = img[c0 + i + dc - 2][f0 + 4 jj + k - 2]                         k = 0..7
//     B_dc[k][fl, co] = w[co][k - fl][dc]  if 0 <= k - fl <= 4 else 0                    fl = 0..3  (Toeplitz in the weights)
//     D[r][fl, co]    = sum_dc A_dc B_dc  = conv output at frame f0 + 4 jj + fl, coefficient c0 + i, channel co
// In the K-major no-swizzle ("interleave") shared-memory layout a core matrix is 8 rows x 16 bytes with rows 16 bytes apart,
// and the second K chunk sits LBO bytes further: with LBO = 16 bytes, row r's frames 4..7 ARE row r+1's frames 0..3, so the
// overlapping Toeplitz rows are one plain copy of the image row (SBO = the row pitch selects coefficient i).  One MMA of
// M = 128, N = 256, K = 8 per coefficient tap dc and 3xTF32 pass: 15 MMAs produce 512 pixels x 64 channels.
//
// Epilogue without a staging tile: thread = TMEM lane = (coefficient i, frame unit jj) holds 4 frames x 64 channels; the
// Max-Feature-Map pair (c, c + 32) and the frame pair of a pooled cell are in its own registers, the coefficient pair is in
// lane ^ 8 (one shuffle).  8 epilogue warps: warps w and w + 4 share TMEM lane quadrant w % 4 and take one frame cell each.
//
// Numerics: 3xTF32 split with fp32 accumulation in TMEM like every other LCNN convolution; the summation order over the 25
// taps differs from the im2col kernel (by coefficient tap, frame taps inside one MMA), i.e. ~1e-7 relative differences.
#include "conv.cuh"
#include "tc_common.cuh"

namespace advb {

namespace {

using namespace tc;

constexpr int TW = 256;                  // worker threads: strip loader + epilogue (16 warps were slower: 96-register cap, 0.132 vs 0.117 ms)
constexpr int TT = TW + 32;              // + one MMA-issue warp
constexpr int NT_MAX = 7;                // frame tiles (32 frames) per work item
constexpr int SW = 32 * NT_MAX + 4;      // strip width in frames: 2 + 2 halo (the K = 8 window of the last unit ends at +3)
constexpr int SROWS = 20;                // 16 coefficients + 2 + 2 halo
constexpr int PLANE = SROWS * SW * 4;    // bytes of one tf32 plane of a strip
constexpr int STRIP = 2 * PLANE;         // hi plane, lo plane
constexpr int WSLICE = 2 * 256 * 16;     // one (dc, part) weight image: [kc][n][16 B]
constexpr int WCONV = 5 * 2 * WSLICE;     // [dc][hi | lo]
constexpr int WBYTES = WCONV + WSLICE;   // + the bias image (see pack_c0t_kernel)
constexpr int ONES = 2304;               // >= 128 rows x 16 B + 16 B of 1.0f: the A operand that adds the bias inside the MMA
constexpr size_t C0T_SMEM = (size_t)WBYTES + ONES + 2 * STRIP + 1024;
static_assert(PLANE % 16 == 0 && (SW * 4) % 16 == 0, "descriptor start addresses and strides are in 16-byte units");
static_assert(C0T_SMEM <= 227 * 1024, "conv0 Toeplitz kernel does not fit shared memory");

struct C0TArgs {
  int B, H, W, Ho, Wo;            // conv grid (frames x coefficients) and pooled grid
  int chunks, tiles_lo, n_big;    // frame tiles per clip are split into `chunks` items: the first n_big have tiles_lo + 1 tiles
  int n_items;
  const float* in;                // (B, H + 4, W + 4) zero-bordered cepstral image
  const unsigned char* wpack;     // WBYTES, see pack_c0t_kernel
  const float* bias;              // (64)
  float* out;                     // (B, Ho + 2 pad, Wo + 2 pad, 32)
  int out_pad;
  unsigned char* codes;           // (B, Ho, Wo, 32)
  int passes;
  int swap_strides;               // diagnostic: exchange the roles of LBO / SBO in the descriptors
};

// K-major, no swizzle: [0,14) start >> 4, [16,30) LBO >> 4, [32,46) SBO >> 4, [46,48) version = 1, [61,64) layout = 0.
// Canonical layout (cute::UMMA, units of 16 bytes): ((8, m), 2) : ((1, SBO), LBO).
__device__ __forceinline__ uint64_t desc_interleave(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t)((smem_addr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) |
         ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) | ((uint64_t)1 << 46);
}

// Weight image of tap column dc, part (0 = tf32 hi, 1 = lo): B[n = fl * 64 + co][k = 4 kc + kk] = w[co][k - fl][dc].
// Bias image (one more slice): B[n][0..2] = the three tf32 pieces of bias[co] (hi + mid + lo = bias to 2^-33), multiplied by an
// all-ones A operand as the first MMA of every tile, so that the accumulator starts at the bias and the epilogue has no
// per-channel add (64 shared-memory loads and 128 adds per thread and tile in the first version).
__global__ void pack_c0t_kernel(const float* __restrict__ w, const float* __restrict__ bias, float* __restrict__ out) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;  // ((dc * 2 + kc) * 256 + n) * 4 + kk
  if (e >= 5 * 2 * 256 * 4 && e < 5 * 2 * 256 * 4 + 2 * 256 * 4) {
    const int i = e - 5 * 2 * 256 * 4, kk = i & 3, n = (i >> 2) & 255, kc = i >> 10;
    float v = 0.f;
    if (kc == 0 && kk < 3) {
      float r = bias[n & 63], h = 0.f, l = 0.f;
      for (int t = 0; t <= kk; ++t) {
        split_tf32(r, h, l);
        r = l;
      }
      v = h;
    }
    out[(size_t)(WCONV / 4) + i] = v;
    return;
  }
  if (e >= 5 * 2 * 256 * 4) return;
  const int kk = e & 3, n = (e >> 2) & 255, kc = (e >> 10) & 1, dc = e >> 11;
  const int fl = n >> 6, co = n & 63, df = 4 * kc + kk - fl;
  const float v = (df >= 0 && df <= 4) ? w[co * 25 + df * 5 + dc] : 0.f;
  float hi, lo;
  split_tf32(v, hi, lo);
  const int o = (kc * 256 + n) * 4 + kk;
  out[(size_t)(dc * 2 + 0) * (WSLICE / 4) + o] = hi;
  out[(size_t)(dc * 2 + 1) * (WSLICE / 4) + o] = lo;
}

__device__ __forceinline__ void item_geometry(const C0TArgs& a, int item, int& b, int& cb, int& t0, int& nt) {
  const int per_clip = 5 * a.chunks;
  b = item / per_clip;
  const int r = item - b * per_clip;
  cb = r / a.chunks;
  const int ch = r - cb * a.chunks;
  t0 = ch * a.tiles_lo + min(ch, a.n_big);
  nt = a.tiles_lo + (ch < a.n_big ? 1 : 0);
}

__global__ void __launch_bounds__(TT, 1) conv0_toeplitz_kernel(C0TArgs a) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  unsigned char* w_s = base;              // WBYTES
  float* ones = reinterpret_cast<float*>(w_s + WBYTES);
  unsigned char* strips = w_s + WBYTES + ONES;   // 2 x STRIP
  __shared__ uint64_t bar_w, bar_strip[2], bar_mma[2], bar_free[2];
  __shared__ uint32_t tmem_base_s;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) {
    mbar_init(&bar_w, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&bar_strip[i], TW / 32);
      mbar_init(&bar_mma[i], 1);
      mbar_init(&bar_free[i], TW / 32);
    }
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc<512>(&tmem_base_s);
  for (int i = tid; i < ONES / 4; i += TT) ones[i] = 1.0f;
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  if (tid == 0) {
    mbar_expect_tx(&bar_w, WBYTES);
    bulk_g2s(w_s, a.wpack, WBYTES, &bar_w);
  }
  constexpr uint32_t IDESC = idesc_tf32(128, 256);
  const uint32_t a_lbo = a.swap_strides ? SW * 4 : 16, a_sbo = a.swap_strides ? 16 : SW * 4;
  const uint32_t b_lbo = a.swap_strides ? 128 : 256 * 16, b_sbo = a.swap_strides ? 256 * 16 : 128;

  if (warp == TW / 32) {
    // ================= MMA-issue warp =================
    mbar_wait(&bar_w, 0u);
    const bool leader = elect_one();
    int k = 0, tcount = 0;
    for (int item = blockIdx.x; item < a.n_items; item += gridDim.x, ++k) {
      int b, cb, t0, nt;
      item_geometry(a, item, b, cb, t0, nt);
      const int sb = k & 1;
      mbar_wait(&bar_strip[sb], (uint32_t)((k >> 1) & 1));
      tc_fence_after();
      const uint32_t s_hi = smem_u32(strips + (size_t)sb * STRIP), s_lo = s_hi + PLANE;
      for (int t = 0; t < nt; ++t, ++tcount) {
        const int acc = tcount & 1;
        if (tcount >= 2) {  // the workers have drained this accumulator (tile tcount - 2)
          mbar_wait(&bar_free[acc], (uint32_t)(((tcount >> 1) - 1) & 1));
          tc_fence_after();
        }
        const uint32_t dcol = tmem + acc * 256;
        if (leader)  // accumulator := 1 * bias  (all-ones A rows: any 16-byte unit of the region)
          mma_tf32(dcol, desc_interleave(smem_u32(ones), a.swap_strides ? 128 : 16, a.swap_strides ? 16 : 128),
                   desc_interleave(smem_u32(w_s + WCONV), b_lbo, b_sbo), IDESC, 0u);
#pragma unroll
        for (int dc = 0; dc < 5; ++dc) {
          const uint32_t off = (uint32_t)(dc * SW + 32 * t) * 4u;
          const uint64_t ah = desc_interleave(s_hi + off, a_lbo, a_sbo), al = desc_interleave(s_lo + off, a_lbo, a_sbo);
          const uint32_t wb = smem_u32(w_s + (size_t)dc * 2 * WSLICE);
          const uint64_t bh = desc_interleave(wb, b_lbo, b_sbo), bl = desc_interleave(wb + WSLICE, b_lbo, b_sbo);
          if (leader) {
            mma_tf32(dcol, ah, bh, IDESC, 1u);
            if (a.passes == 3) {
              mma_tf32(dcol, ah, bl, IDESC, 1u);
              mma_tf32(dcol, al, bh, IDESC, 1u);
            }
          }
        }
        if (leader) mma_commit(&bar_mma[acc]);
        __syncwarp();
      }
    }
  } else {
  // ================= worker warps =================
  const int Hp = a.H + 4, Wp = a.W + 4;
  float4 rv[5];  // the next strip's 20 coefficients of this thread's frame, kept raw until convert_store
  bool rok = false;
  auto issue_loads = [&](int item) {
    int b, cb, t0, nt;
    item_geometry(a, item, b, cb, t0, nt);
    const int row = 32 * t0 + tid;  // padded frame row of strip column fs = tid (frame 32 t0 - 2 + fs, border 2)
    rok = tid < 32 * nt + 4 && row < Hp;
    const float4* src = reinterpret_cast<const float4*>(a.in + ((size_t)b * Hp + (rok ? row : 0)) * Wp + 16 * cb);
#pragma unroll
    for (int q = 0; q < 5; ++q) rv[q] = __ldg(src + q);
  };
  auto convert_store = [&](int sb) {
    if (tid >= SW) return;
    float* hi = reinterpret_cast<float*>(strips + (size_t)sb * STRIP);
    float* lo = hi + PLANE / 4;
#pragma unroll
    for (int q = 0; q < 5; ++q) {
      const float v[4] = {rv[q].x, rv[q].y, rv[q].z, rv[q].w};
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        float h, l;
        split_tf32(rok ? v[u] : 0.f, h, l);
        hi[(4 * q + u) * SW + tid] = h;
        lo[(4 * q + u) * SW + tid] = l;
      }
    }
  };
  auto publish = [&](int sb) {
    fence_proxy_async();
    __syncwarp();
    if (lane == 0) mbar_arrive(&bar_strip[sb]);
  };

  // ---- epilogue geometry of this thread ----
  const int wq = warp & 3, cf = warp >> 2;           // TMEM lane quadrant, frame cell of the unit (frames 2 cf, 2 cf + 1)
  const int r = wq * 32 + lane, ci = r >> 3, jj = r & 7;
  const bool odd = (ci & 1) != 0;                    // coefficient parity = pool column
  const int Hop = a.Ho + 2 * a.out_pad, Wop = a.Wo + 2 * a.out_pad;

  // spread the low 4 bits of m to the low bit of the 4 bytes of a word (bit q -> bit 8 q): the shifted copies do not overlap
  auto spread4 = [](unsigned m) { return ((m & 15u) * 0x00204081u) & 0x01010101u; };

  auto epilogue = [&](int b, int cb, int tile, int acc) {
    const int oy = 16 * tile + 2 * jj + cf, ox = 8 * cb + (ci >> 1);
    const uint32_t taddr = tmem + ((uint32_t)(wq * 32) << 16) + acc * 256 + (uint32_t)(2 * cf) * 64;
#pragma unroll 1
    for (int c0 = 0; c0 < 32; c0 += 16) {
      uint32_t e0[16], e1[16], o0[16], o1[16];
      tmem_ld16_issue(taddr + c0, e0);             // even frame, channels c0..      (first MFM half)
      tmem_ld16_issue(taddr + 32 + c0, e1);        // even frame, channels 32 + c0.. (second MFM half)
      tmem_ld16_issue(taddr + 64 + c0, o0);        // odd frame
      tmem_ld16_issue(taddr + 96 + c0, o1);
      tmem_ld_wait();
      // Winner bookkeeping as three 16-bit masks instead of 16 code registers: SE / SO = the second MFM half wins at the even /
      // odd frame (strictly greater), TK = the odd frame wins the frame pair (strictly greater: first wins ties).
      float v[16];
      unsigned SE = 0u, SO = 0u, TK = 0u;
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const float el = __uint_as_float(e0[j]), eh = __uint_as_float(e1[j]);
        const float ol = __uint_as_float(o0[j]), oh = __uint_as_float(o1[j]);
        const float ve = fmaxf(el, eh), vo = fmaxf(ol, oh);
        SE |= (eh > el ? 1u : 0u) << j;
        SO |= (oh > ol ? 1u : 0u) << j;
        TK |= (vo > ve ? 1u : 0u) << j;
        v[j] = fmaxf(ve, vo);
      }
      const unsigned Hm = (TK & SO) | (~TK & SE);   // MFM half of the winner, per channel
      // coefficient pair: lane ^ 8.  The even lane finalises channels c0..c0+7, the odd lane c0+8..c0+15 (shift `sh`).
      const unsigned sh = odd ? 8u : 0u;
      const unsigned Ho8 = (__shfl_xor_sync(0xffffffffu, Hm, 8) >> sh) & 0xffu;
      const unsigned Do8 = (__shfl_xor_sync(0xffffffffu, TK, 8) >> sh) & 0xffu;
      const unsigned Hm8 = (Hm >> sh) & 0xffu, Dm8 = (TK >> sh) & 0xffu;
      // ATen's max-pool keeps the FIRST maximum in the order (0,0), (0,1), (1,0), (1,1): on equal values the partner wins iff
      // its position is earlier.  Positions: mine = 2 dy_m + dx, partner = 2 dy_o + (dx ^ 1).
      const unsigned earlier = odd ? ~(Do8 & ~Dm8) : (~Do8 & Dm8);
      float res[8];
      unsigned GT = 0u, EQ = 0u;
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const float send = odd ? v[q] : v[8 + q];
        const float mine = odd ? v[8 + q] : v[q];
        const float other = __shfl_xor_sync(0xffffffffu, send, 8);
        GT |= (other > mine ? 1u : 0u) << q;
        EQ |= (other == mine ? 1u : 0u) << q;
        res[q] = fmaxf(mine, other);
      }
      const unsigned T = (GT | (EQ & earlier)) & 0xffu;       // the partner's element is the pooled winner
      const unsigned Hf = (T & Ho8) | (~T & Hm8), Df = (T & Do8) | (~T & Dm8), Xf = T ^ (odd ? 0xffu : 0u);
      if (oy < a.Ho) {
        const int cs = c0 + (odd ? 8 : 0);
        float* o = a.out + (((size_t)b * Hop + oy + a.out_pad) * Wop + ox + a.out_pad) * 32 + cs;
        *reinterpret_cast<float4*>(o) = make_float4(res[0], res[1], res[2], res[3]);
        *reinterpret_cast<float4*>(o + 4) = make_float4(res[4], res[5], res[6], res[7]);
        // code byte = MFM half << 2 | frame parity << 1 | coefficient parity
        const unsigned lo4 = (spread4(Hf) << 2) | (spread4(Df) << 1) | spread4(Xf);
        const unsigned hi4 = (spread4(Hf >> 4) << 2) | (spread4(Df >> 4) << 1) | spread4(Xf >> 4);
        *reinterpret_cast<uint2*>(a.codes + (((size_t)b * a.Ho + oy) * a.Wo + ox) * 32 + cs) = make_uint2(lo4, hi4);
      }
    }
  };

  // ---- persistent loop: strip k + 1 is converted and published before the epilogues of item k, strip k + 2 is in flight ----
  const int G = gridDim.x;
  int item = blockIdx.x;
  if (item < a.n_items) {
    issue_loads(item);
    convert_store(0);
    publish(0);
    if (item + G < a.n_items) issue_loads(item + G);
  }
  int k = 0, tcount = 0;
  for (; item < a.n_items; item += G, ++k) {
    if (item + G < a.n_items) {
      // strip buffer (k + 1) & 1 was last read by the MMAs of item k - 1, which completed before the workers finished that
      // item's epilogues (every tile's bar_mma was waited for)
      convert_store((k + 1) & 1);
      publish((k + 1) & 1);
      if (item + 2 * G < a.n_items) issue_loads(item + 2 * G);
    }
    int b, cb, t0, nt;
    item_geometry(a, item, b, cb, t0, nt);
    for (int t = 0; t < nt; ++t, ++tcount) {
      const int acc = tcount & 1;
      mbar_wait(&bar_mma[acc], (uint32_t)((tcount >> 1) & 1));
      tc_fence_after();
      epilogue(b, cb, t0 + t, acc);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bar_free[acc]);
    }
  }
  }  // worker warps
  // one barrier instruction for every warp of the CTA (compute-sanitizer synccheck flags role branches that end in their own
  // __syncthreads as divergent barriers)
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tmem);
}

}  // namespace

size_t conv0t_pack_bytes() { return WBYTES; }

bool conv0t_supported(int H, int W, int Ho, int Wo) {
  return g_conv_sched == 0 && W == 80 && Wo == 40 && Ho == H / 2 && Ho >= 1;
}

int conv0t_pack(const float* w, const float* bias, unsigned char* wpack, cudaStream_t stream) {
  pack_c0t_kernel<<<cdiv(5 * 2 * 256 * 4 + 2 * 256 * 4, 256), 256, 0, stream>>>(w, bias, reinterpret_cast<float*>(wpack));
  ADVB_KERNEL_OK("pack_c0t", stream);
  return 0;
}

int conv0t_forward(const ConvFwdArgs& f, const unsigned char* wpack, int passes, cudaStream_t stream) {
  ADVB_CHECK(f.Cin == 1 && f.Cout == 64 && f.KS == 5 && f.pool && f.in_pad == 2, "conv0 Toeplitz kernel is the LCNN first block");
  ADVB_CHECK(conv0t_supported(f.H, f.W, f.Ho, f.Wo), "conv0 Toeplitz kernel: unsupported geometry");
  ADVB_CHECK(f.bn_mean == nullptr, "the LCNN first block has no BatchNorm");
  C0TArgs a{};
  a.B = f.B, a.H = f.H, a.W = f.W, a.Ho = f.Ho, a.Wo = f.Wo;
  const int n_ft = cdiv(2 * f.Ho, 32);
  a.chunks = cdiv(n_ft, NT_MAX);
  a.tiles_lo = n_ft / a.chunks;
  a.n_big = n_ft - a.tiles_lo * a.chunks;
  a.n_items = f.B * 5 * a.chunks;
  a.in = f.in, a.wpack = wpack, a.bias = f.bias, a.out = f.out, a.out_pad = f.out_pad, a.codes = f.codes;
  a.passes = passes;
  static const int swap = [] {
    const char* e = getenv("ADVB_C0T_SWAP");
    return e != nullptr ? atoi(e) : 0;
  }();
  a.swap_strides = swap;
  static bool attr_set = false;
  if (!attr_set) {
    ADVB_CUDA_OK(cudaFuncSetAttribute(conv0_toeplitz_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C0T_SMEM));
    attr_set = true;
  }
  int n_sm = 148, dev = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
  const int grid = a.n_items < n_sm ? a.n_items : n_sm;
  conv0_toeplitz_kernel<<<grid, TT, C0T_SMEM, stream>>>(a);
  ADVB_KERNEL_OK(f.tag, stream);
  return 0;
}

}  // namespace advb
