// sm_100a building blocks shared by the tensor-core kernels: mbarrier, 1-D TMA bulk copy (cp.async.bulk ->
// SASS UBLKCP), tcgen05 MMA (kind::tf32, SS mode -> SASS UTCHMMA), TMEM allocation / loads (-> LDTM) and the
// shared-memory matrix descriptor for K-major SWIZZLE_128B operands.
//
// Conventions verified on a B200 with tools/tc_probe.cu (gpurun_out/tc_probe.log, summarised in DESIGN.md):
//   * kind::tf32 TRUNCATES fp32 inputs to tf32 (error vs truncated inputs 1e-6, vs rounded inputs 6e-3);
//   * the 128-byte swizzle is a function of the absolute shared-memory address, so an operand may start at ANY
//     128-byte row of a 1024-byte-aligned band with matrix-descriptor base_offset = 0 (base_offset = (addr>>7)&7
//     gives wrong results) - this is what lets the 9 taps of a 3x3 convolution share one halo band;
//   * SS-mode issue rate is shared-memory-read bound: (128 + N) / 4 cycles per M=128, K=8 MMA (N = 64: 48,
//     N = 96: 56, N = 128: 64 cycles measured).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace advb {
namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- mbarrier ----------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Spin with back-off: a bare try_wait loop of the 7 waiting warps of a CTA was 28 % of all executed instructions of the
// 3x3 backward kernel and stole issue slots from the co-resident CTA's fill (profiles/r01_ncu_full_persistent.md).
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_try(bar, parity)) __nanosleep(40);
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
// generic-proxy shared-memory writes -> visible to the async proxy (tcgen05.mma operand reads, bulk copies)
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- TMA 1-D bulk copy global -> shared, completion on an mbarrier ------------------------------------------
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// ---- TMEM ---------------------------------------------------------------------------------------------------
template <int COLS>
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem) {  // one full warp
  static_assert(COLS >= 32 && COLS <= 512 && (COLS & (COLS - 1)) == 0, "TMEM columns: power of two in [32, 512]");
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(COLS));
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
}
template <int COLS>
__device__ __forceinline__ void tmem_dealloc(uint32_t addr) {  // the allocating warp
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(COLS));
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// 32 lanes x 16 consecutive 32-bit columns: thread t of the warp receives lane (addr.lane + t).  A warp may only
// touch TMEM lanes [32 * (warp_id % 4), +32).
__device__ __forceinline__ void tmem_ld16_issue(uint32_t addr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(addr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---- descriptors --------------------------------------------------------------------------------------------
// K-major operand, 32 fp32 (128 bytes) per row, SWIZZLE_128B, 8-row groups 1024 bytes apart.
// Bit layout (cute::UMMA::SmemDescriptor): [0,14) start>>4, [16,30) LBO>>4, [32,46) SBO>>4, [46,48) version = 1,
// [49,52) base_offset = 0, [61,64) layout type (2 = SWIZZLE_128B).
__device__ __forceinline__ uint64_t desc_sw128(uint32_t smem_addr) {
  return (uint64_t)((smem_addr >> 4) & 0x3FFF) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) |
         ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
// Instruction descriptor (cute::UMMA::InstrDescriptor): D = F32 (bits [4,6) = 1), A = B = TF32 (bits [7,10),
// [10,13) = 2), both K-major, N>>3 at [17,23), M>>4 at [24,29).
__host__ __device__ constexpr uint32_t idesc_tf32(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// D[tmem] (+)= A[smem] * B[smem]^T, M = 128, K = 8 (tf32), issued by ONE thread.
__device__ __forceinline__ void mma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                         uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// bf16 operands (kind::f16), fp32 accumulator: same descriptor layout, a_format = b_format = 1 (BF16); K = 16 per instruction.
__host__ __device__ constexpr uint32_t idesc_bf16(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void mma_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                         uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Two floats -> one 32-bit word of two bf16 (round to nearest even): element 0 in the low half.
__device__ __forceinline__ uint32_t pack_bf16x2(float e0, float e1) {
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(e1), "f"(e0));
  return r;
}
// Arrive on an mbarrier when every previously issued MMA of this thread has completed (implies
// tcgen05.fence::before_thread_sync).
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

// One lane of a converged warp (warp-uniform predicate the compiler understands: operands stay uniform).
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

// ---- division by a launch-time constant ------------------------------------------------------------------------
// n / d as one multiply-high: m = ceil(2^32 / d) is exact whenever n * d < 2^32 (pixel and tile indices here are
// < 2^17 and divisors < 2^16).  A runtime integer division costs ~25 SASS instructions; the fill and epilogue loops
// of the convolution kernels were issue-bound on them (profiles/r01_ncu_full_persistent.md).
struct FastDiv {
  uint32_t d, m;
};
inline FastDiv make_fastdiv(int d) {
  FastDiv f;
  f.d = (uint32_t)d;
  f.m = d > 1 ? (uint32_t)((0x100000000ull + (uint64_t)d - 1) / (uint64_t)d) : 0u;
  return f;
}
__device__ __forceinline__ int fdiv(int n, const FastDiv& f) { return f.d > 1 ? (int)__umulhi((uint32_t)n, f.m) : n; }

// ---- 3xTF32 split -------------------------------------------------------------------------------------------
// x = hi + lo exactly, hi = round-to-nearest tf32(x).  The tensor core then computes
// hi_a*hi_b + hi_a*lo_b + lo_a*hi_b with fp32 accumulation: products carry ~2^-22 relative error, i.e. fp32-class
// accuracy (measured 3e-6 absolute on K=32 dot products of U(-1,1) values, tools/tc_probe.cu).
__device__ __forceinline__ void split_tf32(float x, float& hi, float& lo) {
  uint32_t h;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(h) : "f"(x));
  hi = __uint_as_float(h);
  lo = x - hi;
}
// byte offset of the 16-byte chunk c (0..7) of row r in a SWIZZLE_128B K-major tile whose base is 1024-aligned
__device__ __forceinline__ uint32_t sw128_chunk(int r, int c) { return (uint32_t)(r * 128 + ((c ^ (r & 7)) << 4)); }

}  // namespace tc
}  // namespace advb
