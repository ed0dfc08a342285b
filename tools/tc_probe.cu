// tcgen05 probe (diagnostic tool, not product code): validates on a real B200 the descriptor conventions that
// csrc/conv_tc.cu relies on, before they are baked into the convolution kernels:
//   1. kind::tf32 SS-mode MMA, M=128, N in {32,48,64,96,128}, K-major SWIZZLE_128B operands written by threads;
//   2. A-operand start rows that are NOT multiples of 8 (the implicit-GEMM "tap shift" is a row offset into one
//      shared halo band): which matrix-descriptor base_offset convention gives the right answer;
//   3. whether the tensor core truncates or rounds fp32 inputs to tf32;
//   4. 3xTF32 (hi/lo split) accuracy against fp64;
//   5. issue-to-completion cycles per MMA for the N values used.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/tc_probe tools/tc_probe.cu
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#define CK(x)                                                                      \
  do {                                                                             \
    cudaError_t e = (x);                                                           \
    if (e != cudaSuccess) {                                                        \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); \
      exit(2);                                                                     \
    }                                                                              \
  } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
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
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t sbo_bytes, uint32_t base_offset) {
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;                               // LBO (ignored for swizzled K-major)
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;     // SBO
  d |= (uint64_t)1 << 46;                               // descriptor version (Blackwell)
  d |= (uint64_t)(base_offset & 7) << 49;
  d |= (uint64_t)2 << 61;                               // SWIZZLE_128B
  return d;
}
__device__ __forceinline__ uint32_t make_idesc_tf32(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void mma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t addr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(addr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

constexpr int A_ROWS = 256;  // band rows available to the probe
constexpr int KC = 32;       // floats per 128-byte row

__device__ __forceinline__ uint32_t swz(int r, int k) {  // byte offset of element (row r, k) in a SW128 K-major tile
  return (uint32_t)(r * 128 + ((((k >> 2) ^ (r & 7)) << 4) | ((k & 3) << 2)));
}

// mode bit0: base_offset = (addr>>7)&7 instead of 0.  mode bit1: 3xTF32 (hi/lo split on device).
// Output D (128 x N).  timing: if reps > 0, re-issue the MMA group `reps` times and report cycles.
__global__ void __launch_bounds__(128) probe_kernel(const float* __restrict__ A, const float* __restrict__ Bm, float* D,
                                                     int N, int r0, int mode, int reps, long long* cycles, int* err) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* base = (unsigned char*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  unsigned char* a_hi = base;                         // 256 x 128 B
  unsigned char* a_lo = a_hi + A_ROWS * 128;          // 256 x 128 B
  unsigned char* b_hi = a_lo + A_ROWS * 128;          // 128 x 128 B
  unsigned char* b_lo = b_hi + 128 * 128;             // 128 x 128 B
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5;
  const bool split = (mode & 2) != 0;

  for (int i = tid; i < A_ROWS * KC; i += 128) {
    const int r = i / KC, k = i % KC;
    const float v = A[i];
    float hi = v, lo = 0.f;
    if (split) {
      uint32_t h;
      asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(h) : "f"(v));
      hi = __uint_as_float(h);
      lo = v - hi;
    }
    *(float*)(a_hi + swz(r, k)) = hi;
    *(float*)(a_lo + swz(r, k)) = lo;
  }
  for (int i = tid; i < 128 * KC; i += 128) {
    const int r = i / KC, k = i % KC;
    const float v = r < N ? Bm[i] : 0.f;
    float hi = v, lo = 0.f;
    if (split) {
      uint32_t h;
      asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(h) : "f"(v));
      hi = __uint_as_float(h);
      lo = v - hi;
    }
    *(float*)(b_hi + swz(r, k)) = hi;
    *(float*)(b_lo + swz(r, k)) = lo;
  }
  if (tid == 0) mbar_init(&bar, 1);
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "r"(128));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy smem writes -> visible to the MMA
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tm = tmem_base;

  long long t0 = 0, t1 = 0;
  if (tid == 0) {
    const uint32_t idesc = make_idesc_tf32(128, N);
    const int groups = reps > 0 ? reps : 1;
    t0 = clock64();
    for (int g = 0; g < groups; ++g) {
      uint32_t acc = 0;
      for (int ks = 0; ks < 4; ++ks) {
        const uint32_t aaddr_hi = smem_u32(a_hi) + r0 * 128 + ks * 32;
        const uint32_t aaddr_lo = smem_u32(a_lo) + r0 * 128 + ks * 32;
        const uint32_t bo = (mode & 1) ? ((aaddr_hi >> 7) & 7) : 0;
        const uint64_t ah = make_desc(aaddr_hi, 1024, bo), al = make_desc(aaddr_lo, 1024, bo);
        const uint64_t bh = make_desc(smem_u32(b_hi) + ks * 32, 1024, 0), bl = make_desc(smem_u32(b_lo) + ks * 32, 1024, 0);
        mma_tf32(tm, ah, bh, idesc, acc);
        acc = 1;
        if (split) {
          mma_tf32(tm, ah, bl, idesc, 1);
          mma_tf32(tm, al, bh, idesc, 1);
        }
      }
    }
    mma_commit(&bar);
  }
  // bounded wait
  bool done = false;
  for (int it = 0; it < (1 << 22) && !done; ++it) done = mbar_try(&bar, 0);
  if (tid == 0) {
    t1 = clock64();
    if (cycles) *cycles = t1 - t0;
  }
  if (!done) {
    if (tid == 0) *err = 1;
  } else {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int row = tid;  // warp w reads TMEM lanes 32w..32w+31
    for (int c0 = 0; c0 < N; c0 += 16) {
      uint32_t v[16];
      tmem_ld16(tm + ((uint32_t)(warp * 32) << 16) + c0, v);
      for (int j = 0; j < 16; ++j) D[row * N + c0 + j] = __uint_as_float(v[j]);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(128));
}

static float tf32_trunc(float v) {
  uint32_t u;
  memcpy(&u, &v, 4);
  u &= 0xFFFFE000u;
  memcpy(&v, &u, 4);
  return v;
}
static float tf32_rna(float v) {
  uint32_t u;
  memcpy(&u, &v, 4);
  u += 0x1000u;
  u &= 0xFFFFE000u;
  memcpy(&v, &u, 4);
  return v;
}

int main() {
  const size_t smem = (size_t)(2 * A_ROWS * 128 + 2 * 128 * 128) + 1024;
  CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  std::vector<float> hA(A_ROWS * KC), hB(128 * KC);
  srand(1);
  for (auto& v : hA) v = (float)rand() / RAND_MAX * 2.f - 1.f;
  for (auto& v : hB) v = (float)rand() / RAND_MAX * 2.f - 1.f;
  float *dA, *dB, *dD;
  long long* dcyc;
  int* derr;
  CK(cudaMalloc(&dA, hA.size() * 4));
  CK(cudaMalloc(&dB, hB.size() * 4));
  CK(cudaMalloc(&dD, 128 * 128 * 4));
  CK(cudaMalloc(&dcyc, 8));
  CK(cudaMalloc(&derr, 4));
  CK(cudaMemcpy(dA, hA.data(), hA.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, hB.data(), hB.size() * 4, cudaMemcpyHostToDevice));
  std::vector<float> hD(128 * 128);

  auto run = [&](int N, int r0, int mode, int reps, long long* cyc) -> int {
    CK(cudaMemset(derr, 0, 4));
    CK(cudaMemset(dD, 0, 128 * 128 * 4));
    probe_kernel<<<1, 128, smem>>>(dA, dB, dD, N, r0, mode, reps, dcyc, derr);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
      printf("  kernel failed: %s\n", cudaGetErrorString(e));
      exit(3);
    }
    int herr;
    CK(cudaMemcpy(&herr, derr, 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(hD.data(), dD, 128 * N * 4, cudaMemcpyDeviceToHost));
    if (cyc) CK(cudaMemcpy(cyc, dcyc, 8, cudaMemcpyDeviceToHost));
    return herr;
  };
  auto max_err = [&](int N, int r0, int kind) {  // kind 0: trunc inputs, 1: rna inputs, 2: exact fp64
    double worst = 0;
    for (int r = 0; r < 128; ++r)
      for (int n = 0; n < N; ++n) {
        double s = 0;
        for (int k = 0; k < KC; ++k) {
          const float a = hA[(r0 + r) * KC + k], b = hB[n * KC + k];
          if (kind == 0) s += (double)tf32_trunc(a) * tf32_trunc(b);
          else if (kind == 1) s += (double)tf32_rna(a) * tf32_rna(b);
          else s += (double)a * b;
        }
        worst = fmax(worst, fabs(s - (double)hD[r * N + n]));
      }
    return worst;
  };

  printf("== 1. shapes, r0=0, single-pass tf32 ==\n");
  const int Ns[5] = {32, 48, 64, 96, 128};
  for (int N : Ns) {
    int err = run(N, 0, 0, 0, nullptr);
    printf("N=%3d hang=%d  err_vs_trunc=%.3e err_vs_rna=%.3e err_vs_exact=%.3e\n", N, err, max_err(N, 0, 0),
           max_err(N, 0, 1), max_err(N, 0, 2));
  }
  printf("== 2. A start-row offsets (N=96): base_offset=0 vs base_offset=(addr>>7)&7 ==\n");
  const int r0s[10] = {0, 1, 2, 3, 5, 8, 9, 42, 43, 85};
  for (int r0 : r0s) {
    int e0 = run(96, r0, 0, 0, nullptr);
    double m0 = max_err(96, r0, 0);
    int e1 = run(96, r0, 1, 0, nullptr);
    double m1 = max_err(96, r0, 0);
    printf("r0=%3d  bo=0: hang=%d err=%.3e   bo=calc: hang=%d err=%.3e\n", r0, e0, m0, e1, m1);
  }
  printf("== 4. 3xTF32 (N=96) ==\n");
  for (int r0 : {0, 43}) {
    for (int mode : {2, 3}) {
      int e = run(96, r0, mode, 0, nullptr);
      printf("r0=%2d mode=%d hang=%d err_vs_exact=%.3e\n", r0, mode, e, max_err(96, r0, 2));
    }
  }
  printf("== 5. cycles per MMA (4 k-steps per group, 256 groups) ==\n");
  for (int N : Ns) {
    long long c1 = 0, c3 = 0;
    run(N, 0, 0, 256, &c1);
    run(N, 0, 2, 256, &c3);
    printf("N=%3d single-pass: %.1f cyc/MMA   3xTF32: %.1f cyc/MMA\n", N, (double)c1 / (256 * 4), (double)c3 / (256 * 12));
  }
  printf("probe done\n");
  return 0;
}
