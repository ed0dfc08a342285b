// Recurrent tail of LCNN: 2 x bidirectional LSTM(160 -> 2x80), residual mean pooling, Linear(160 -> 1), the 2-class
// cross-entropy head and the backward of all of it (sm_100a, fp32).
//
// Replaces BLSTMLayer (src/models/lcnn.py:24-46), the embedding tail (lcnn.py:196-206) and the patched attack head
// cat([-o, o]) + CrossEntropyLoss + autograd (fgsm.py:43-57, pgd.py:61-72).
//
// Structure: the input projection of all time steps is one small GEMM (gemm_kernel); the recurrence keeps W_hh^T
// (80x320 fp32 = 100 KB) resident in shared memory for the whole sequence, one CTA per (pair of clips, direction);
// gate activations overwrite the projection buffer in place and are reused by BPTT, which again keeps W_hh in
// shared memory and overwrites the same buffer with the pre-activation gate gradients; the input gradient is a
// second small GEMM.  PyTorch gate order i,f,g,o; zero initial state.
#include "rnn.cuh"

#include "gemm.cuh"

namespace advb {

namespace {

constexpr int HID = 80;
constexpr int G4 = 4 * HID;  // 320
#ifndef ADVB_LSTM_CL
#define ADVB_LSTM_CL 2
#endif
constexpr int CL = ADVB_LSTM_CL;  // clips per recurrence CTA

// C[M,N] = A[M,K] * Bm[K,N] (+ bias[N]) (+ Cadd[M,N]); row-major, K and N multiples of 4, 16-byte aligned rows.
// BM x BN tile, 16-deep k-steps, 256 threads with a (BM/16) x (BN/16) register tile; 128-bit global and shared loads.
// BN = 32 is used when 64-wide tiles would leave most SMs idle (the narrow N = 160 backward projections).
template <int BM, int BN>
__global__ void __launch_bounds__(256) gemm_kernel(const float* __restrict__ A, const float* __restrict__ Bm,
                                                    const float* __restrict__ bias, const float* __restrict__ Cadd,
                                                    float* __restrict__ C, int M, int N, int K) {
  constexpr int RM = BM / 16, RN = BN / 16;
  static_assert(BM == 64 && (BN == 64 || BN == 32), "tile shapes used by the BLSTM projections");
  __shared__ __align__(16) float As[16][BM + 4];  // k-major: As[k][m]
  __shared__ __align__(16) float Bs[16][BN];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  float acc[RM][RN] = {};
  const int a_row = tid >> 2, a_kq = (tid & 3) * 4;            // A tile: 64 rows x 4 float4 along k
  const int b_row = tid / (BN / 4), b_c4 = (tid % (BN / 4)) * 4;  // B tile: 16 rows x BN/4 float4
  // register prefetch: the global loads of k-step i+1 are in flight while k-step i is multiplied (one exposed load
  // latency per launch instead of one per k-step; the K = 640 backward projection has 40 of them)
  auto load_a = [&](int k0) {
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (m0 + a_row < M && k0 + a_kq < K) v = __ldg(reinterpret_cast<const float4*>(A + (size_t)(m0 + a_row) * K + k0 + a_kq));
    return v;
  };
  auto load_b = [&](int k0) {
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (b_row < 16 && k0 + b_row < K && n0 + b_c4 < N)
      v = __ldg(reinterpret_cast<const float4*>(Bm + (size_t)(k0 + b_row) * N + n0 + b_c4));
    return v;
  };
  float4 pa = load_a(0), pb = load_b(0);
  for (int k0 = 0; k0 < K; k0 += 16) {
    As[a_kq + 0][a_row] = pa.x;
    As[a_kq + 1][a_row] = pa.y;
    As[a_kq + 2][a_row] = pa.z;
    As[a_kq + 3][a_row] = pa.w;
    if (b_row < 16) *reinterpret_cast<float4*>(&Bs[b_row][b_c4]) = pb;
    if (k0 + 16 < K) pa = load_a(k0 + 16), pb = load_b(k0 + 16);
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < 16; ++kk) {
      float a[RM], bv[RN];
      const float4 av = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      a[0] = av.x, a[1] = av.y, a[2] = av.z, a[3] = av.w;
      if (RN == 4) {
        const float4 b4 = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
        bv[0] = b4.x, bv[1] = b4.y, bv[RN - 2] = b4.z, bv[RN - 1] = b4.w;
      } else {
        const float2 b2 = *reinterpret_cast<const float2*>(&Bs[kk][tx * 2]);
        bv[0] = b2.x, bv[1] = b2.y;
      }
#pragma unroll
      for (int i = 0; i < RM; ++i)
#pragma unroll
        for (int j = 0; j < RN; ++j) acc[i][j] = fmaf(a[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < RM; ++i) {
    const int m = m0 + ty * RM + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < RN; ++j) {
      const int n = n0 + tx * RN + j;
      if (n >= N) continue;
      float v = acc[i][j];
      if (bias != nullptr) v += bias[n];
      if (Cadd != nullptr) v += Cadd[(size_t)m * N + n];
      C[(size_t)m * N + n] = v;
    }
  }
}

// General register-tiled variant (round 2): BM x BN tile, (BM / TM) x (BN / TN) threads with a TM x TN register tile each.  Every
// output is still ONE sequential fmaf chain over k = 0 .. K-1, so the results are bit-identical to gemm_kernel's whatever the tile
// shape; what changes is the LDS : FMA ratio (gemm_kernel's 4 x 4 tile issues two LDS.128 per 16 FMAs, its 4 x 2 tile for the narrow
// N = 160 backward projection one LDS.128 + one LDS.64 per 8).
template <int BM, int BN, int TM, int TN>
__global__ void __launch_bounds__((BM / TM) * (BN / TN)) gemm_rt_kernel(const float* __restrict__ A, const float* __restrict__ Bm,
                                                                         const float* __restrict__ bias,
                                                                         const float* __restrict__ Cadd, float* __restrict__ C,
                                                                         int M, int N, int K) {
  constexpr int NX = BN / TN, NY = BM / TM, NT = NX * NY;
  constexpr int A4 = BM * 4, B4 = 4 * BN;  // float4 loads per k-step of the A / B tile
  constexpr int PA = (A4 + NT - 1) / NT, PB = (B4 + NT - 1) / NT;
  static_assert(TM % 4 == 0 && (TN == 4 || TN == 2) && BN % 4 == 0, "register tile shapes");
  __shared__ __align__(16) float As[16][BM + 4];  // k-major
  __shared__ __align__(16) float Bs[16][BN + 4];
  const int tid = threadIdx.x, tx = tid % NX, ty = tid / NX;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  float acc[TM][TN] = {};
  float4 pa[PA], pb[PB];
  auto load_tiles = [&](int k0) {
#pragma unroll
    for (int q = 0; q < PA; ++q) {
      const int i = tid + q * NT, row = i >> 2, kq = (i & 3) * 4;
      pa[q] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (i < A4 && m0 + row < M && k0 + kq < K) pa[q] = __ldg(reinterpret_cast<const float4*>(A + (size_t)(m0 + row) * K + k0 + kq));
    }
#pragma unroll
    for (int q = 0; q < PB; ++q) {
      const int i = tid + q * NT, row = i / (BN / 4), c4 = (i % (BN / 4)) * 4;
      pb[q] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (i < B4 && k0 + row < K && n0 + c4 < N) pb[q] = __ldg(reinterpret_cast<const float4*>(Bm + (size_t)(k0 + row) * N + n0 + c4));
    }
  };
  load_tiles(0);
  for (int k0 = 0; k0 < K; k0 += 16) {
#pragma unroll
    for (int q = 0; q < PA; ++q) {
      const int i = tid + q * NT, row = i >> 2, kq = (i & 3) * 4;
      if (i < A4) {
        As[kq + 0][row] = pa[q].x;
        As[kq + 1][row] = pa[q].y;
        As[kq + 2][row] = pa[q].z;
        As[kq + 3][row] = pa[q].w;
      }
    }
#pragma unroll
    for (int q = 0; q < PB; ++q) {
      const int i = tid + q * NT, row = i / (BN / 4), c4 = (i % (BN / 4)) * 4;
      if (i < B4) *reinterpret_cast<float4*>(&Bs[row][c4]) = pb[q];
    }
    if (k0 + 16 < K) load_tiles(k0 + 16);
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < 16; ++kk) {
      float a[TM], bv[TN];
#pragma unroll
      for (int i = 0; i < TM; i += 4) {
        const float4 av = *reinterpret_cast<const float4*>(&As[kk][ty * TM + i]);
        a[i] = av.x, a[i + 1] = av.y, a[i + 2] = av.z, a[i + 3] = av.w;
      }
      if (TN == 4) {
        const float4 b4 = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
        bv[0] = b4.x, bv[1] = b4.y, bv[TN - 2] = b4.z, bv[TN - 1] = b4.w;
      } else {
        const float2 b2 = *reinterpret_cast<const float2*>(&Bs[kk][tx * 2]);
        bv[0] = b2.x, bv[1] = b2.y;
      }
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int m = m0 + ty * TM + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int n = n0 + tx * TN + j;
      if (n >= N) continue;
      float v = acc[i][j];
      if (bias != nullptr) v += bias[n];
      if (Cadd != nullptr) v += Cadd[(size_t)m * N + n];
      C[(size_t)m * N + n] = v;
    }
  }
}

// Pack one BLSTM layer: wihT [160][640] (col = dir*320 + gate row), bias [640] = b_ih + b_hh,
// whhT [2][80][320], wih_cat [640][160], whh [2][320][80] (straight copies of the live tensors).
// perm_wf > 0: the layer's input arrives in the convolution's NHWC order (index w * C + c, C = I / perm_wf) instead of the
// reference's (c * Wf + w) of lcnn.py:196-199: the input-projection weights are permuted here, once per call, so that
// the block output feeds the GEMM as it is and the input gradient lands in the block's gradient buffer directly (the
// feats_gather / feats_scatter kernels of every iteration are gone).
__global__ void lstm_pack_kernel(LstmWeights w, LstmPacked p, int I, int perm_wf) {
  const int tid = blockIdx.x * blockDim.x + threadIdx.x, nt = gridDim.x * blockDim.x;
  for (int i = tid; i < 2 * G4 * I; i += nt) {
    const int k = i % I, j = (i / I) % G4, d = i / (I * G4);
    const float v = (d == 0 ? w.w_ih[0] : w.w_ih[1])[(size_t)j * I + k];
    const int kp = perm_wf > 0 ? (k % perm_wf) * (I / perm_wf) + k / perm_wf : k;
    p.wihT[(size_t)kp * (2 * G4) + d * G4 + j] = v;
    p.wih_cat[(size_t)(d * G4 + j) * I + kp] = v;
  }
  for (int i = tid; i < 2 * G4 * HID; i += nt) {
    const int k = i % HID, j = (i / HID) % G4, d = i / (HID * G4);
    const float v = (d == 0 ? w.w_hh[0] : w.w_hh[1])[(size_t)j * HID + k];
    p.whhT[((size_t)d * HID + k) * G4 + j] = v;
    p.whh[((size_t)d * G4 + j) * HID + k] = v;
  }
  for (int i = tid; i < 2 * G4; i += nt) {
    const int j = i % G4, d = i / G4;
    p.bias[i] = (d == 0 ? w.b_ih[0] : w.b_ih[1])[j] + (d == 0 ? w.b_hh[0] : w.b_hh[1])[j];
  }
}

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

// gates (B,L,640): in = input projection (+bias), out = activated gates i,f,g,o.  hout (B,L,160), cs (B,L,2,80).
__global__ void __launch_bounds__(G4) lstm_rec_fwd_kernel(float* __restrict__ gates, const float* __restrict__ whhT,
                                                           float* __restrict__ hout, float* __restrict__ cs, int B,
                                                           int L) {
  extern __shared__ __align__(16) float smem[];
  float* s_h = smem;                 // [CL][80]
  float* s_a = s_h + CL * HID;       // [CL][320]
  const int j = threadIdx.x, dir = blockIdx.y, b0 = blockIdx.x * CL;
  // persistent-RNN style: thread j keeps its 80 recurrent weights (column j of W_hh^T) in registers for the whole
  // sequence, so a step reads only h from shared memory (the per-step W reads were 2/3 of the shared-memory traffic)
  const float* wsrc = whhT + (size_t)dir * HID * G4;
  float wreg[HID];
#pragma unroll
  for (int k = 0; k < HID; ++k) wreg[k] = __ldg(wsrc + (size_t)k * G4 + j);
  if (j < CL * HID) s_h[j] = 0.f;
  float c_reg[CL];
#pragma unroll
  for (int cl = 0; cl < CL; ++cl) c_reg[cl] = 0.f;
  __syncthreads();
  const int gate = j / HID;
  // the input projection of step s+1 is fetched while step s runs: the L2 round trip (~0.5 us) was a dependent load
  // at the head of each of the 25 sequential steps
  float pre[CL];
  {
    const int t0 = dir == 0 ? 0 : L - 1;
#pragma unroll
    for (int cl = 0; cl < CL; ++cl)
      pre[cl] = (b0 + cl < B) ? gates[((size_t)(b0 + cl) * L + t0) * (2 * G4) + dir * G4 + j] : 0.f;
  }
  for (int step = 0; step < L; ++step) {
    const int t = dir == 0 ? step : L - 1 - step;
    float acc[CL];
#pragma unroll
    for (int cl = 0; cl < CL; ++cl) acc[cl] = pre[cl];
    if (step + 1 < L) {
      const int tn = dir == 0 ? step + 1 : L - 2 - step;
#pragma unroll
      for (int cl = 0; cl < CL; ++cl)
        pre[cl] = (b0 + cl < B) ? gates[((size_t)(b0 + cl) * L + tn) * (2 * G4) + dir * G4 + j] : 0.f;
    }
    float acc2[CL];  // second accumulator chain per clip (halves the dependent FMA chain)
#pragma unroll
    for (int cl = 0; cl < CL; ++cl) acc2[cl] = 0.f;
#pragma unroll
    for (int k = 0; k < HID; k += 8) {
#pragma unroll
      for (int cl = 0; cl < CL; ++cl) {
        const float4 h0 = *reinterpret_cast<const float4*>(s_h + cl * HID + k);
        const float4 h1 = *reinterpret_cast<const float4*>(s_h + cl * HID + k + 4);
        acc[cl] = fmaf(wreg[k + 3], h0.w, fmaf(wreg[k + 2], h0.z, fmaf(wreg[k + 1], h0.y, fmaf(wreg[k], h0.x, acc[cl]))));
        acc2[cl] = fmaf(wreg[k + 7], h1.w, fmaf(wreg[k + 6], h1.z, fmaf(wreg[k + 5], h1.y, fmaf(wreg[k + 4], h1.x, acc2[cl]))));
      }
    }
#pragma unroll
    for (int cl = 0; cl < CL; ++cl) acc[cl] += acc2[cl];
#pragma unroll
    for (int cl = 0; cl < CL; ++cl) {
      const float a = gate == 2 ? tanhf(acc[cl]) : sigmoidf_(acc[cl]);
      s_a[cl * G4 + j] = a;
      if (b0 + cl < B) gates[((size_t)(b0 + cl) * L + t) * (2 * G4) + dir * G4 + j] = a;
    }
    __syncthreads();
    if (j < HID) {
#pragma unroll
      for (int cl = 0; cl < CL; ++cl) {
        const float* a = s_a + cl * G4;
        const float c = a[HID + j] * c_reg[cl] + a[j] * a[2 * HID + j];
        const float h = a[3 * HID + j] * tanhf(c);
        c_reg[cl] = c;
        s_h[cl * HID + j] = h;
        if (b0 + cl < B) {
          hout[((size_t)(b0 + cl) * L + t) * (2 * HID) + dir * HID + j] = h;
          cs[(((size_t)(b0 + cl) * L + t) * 2 + dir) * HID + j] = c;
        }
      }
    }
    __syncthreads();
  }
}

// BPTT.  gates (B,L,640): in = activated gates, out = gradient w.r.t. gate pre-activations.  dout (B,L,160).
__global__ void __launch_bounds__(G4) lstm_rec_bwd_kernel(float* __restrict__ gates, const float* __restrict__ whh,
                                                           const float* __restrict__ dout,
                                                           const float* __restrict__ cs, int B, int L) {
  extern __shared__ __align__(16) float smem[];
  float* s_dg = smem;                  // [CL][320]
  float* s_dh = s_dg + CL * G4;        // [CL][80]
  float* s_part = s_dh + CL * HID;     // [CL][4][80]
  const int j = threadIdx.x, dir = blockIdx.y, b0 = blockIdx.x * CL;
  const float* wsrc = whh + (size_t)dir * G4 * HID;
  if (j < CL * HID) s_dh[j] = 0.f;
  float dc[CL];
#pragma unroll
  for (int cl = 0; cl < CL; ++cl) dc[cl] = 0.f;
  __syncthreads();
  const int part = j / HID, u = j % HID;
  // thread (part, u) keeps W_hh[part*80 + jj][u], jj = 0..79, in registers for the whole sequence
  float wreg[HID];
#pragma unroll
  for (int jj = 0; jj < HID; ++jj) wreg[jj] = __ldg(wsrc + (size_t)(part * HID + jj) * HID + u);
  // operands of a step (gates, cell, previous cell, dout) are fetched one step ahead (threads j < HID)
  float n_g[CL][4], n_c[CL], n_cp[CL], n_do[CL];
  auto fetch = [&](int step) {
    const int t = dir == 0 ? step : L - 1 - step;
    const int tp = dir == 0 ? t - 1 : t + 1;
#pragma unroll
    for (int cl = 0; cl < CL; ++cl) {
      n_g[cl][0] = n_g[cl][1] = n_g[cl][2] = n_g[cl][3] = 0.f;
      n_c[cl] = n_cp[cl] = n_do[cl] = 0.f;
      if (j < HID && b0 + cl < B) {
        const size_t bt = (size_t)(b0 + cl) * L + t;
        const float* g = gates + bt * (2 * G4) + dir * G4;
        n_g[cl][0] = g[j], n_g[cl][1] = g[HID + j], n_g[cl][2] = g[2 * HID + j], n_g[cl][3] = g[3 * HID + j];
        n_c[cl] = cs[(bt * 2 + dir) * HID + j];
        n_cp[cl] = step > 0 ? cs[((((size_t)(b0 + cl) * L + tp) * 2) + dir) * HID + j] : 0.f;
        n_do[cl] = dout[bt * (2 * HID) + dir * HID + j];
      }
    }
  };
  fetch(L - 1);
  for (int step = L - 1; step >= 0; --step) {
    const int t = dir == 0 ? step : L - 1 - step;
    float c_g[CL][4], c_c[CL], c_cp[CL], c_do[CL];
#pragma unroll
    for (int cl = 0; cl < CL; ++cl) {
      c_g[cl][0] = n_g[cl][0], c_g[cl][1] = n_g[cl][1], c_g[cl][2] = n_g[cl][2], c_g[cl][3] = n_g[cl][3];
      c_c[cl] = n_c[cl], c_cp[cl] = n_cp[cl], c_do[cl] = n_do[cl];
    }
    // NOTE: the gates row of step-1 is still the forward's activations: this step only overwrites row `t` below
    if (step > 0) fetch(step - 1);
    if (j < HID) {
#pragma unroll
      for (int cl = 0; cl < CL; ++cl) {
        float dgi = 0.f, dgf = 0.f, dgg = 0.f, dgo = 0.f;
        if (b0 + cl < B) {
          const float gi = c_g[cl][0], gf = c_g[cl][1], gg = c_g[cl][2], go = c_g[cl][3];
          const float c = c_c[cl];
          const float cp = c_cp[cl];
          const float dh = c_do[cl] + s_dh[cl * HID + j];
          const float tc = tanhf(c);
          const float d_o = dh * tc;
          const float dcc = dc[cl] + dh * go * (1.f - tc * tc);
          dc[cl] = dcc * gf;
          dgi = dcc * gg * gi * (1.f - gi);
          dgf = dcc * cp * gf * (1.f - gf);
          dgg = dcc * gi * (1.f - gg * gg);
          dgo = d_o * go * (1.f - go);
        }
        float* d = s_dg + cl * G4;
        d[j] = dgi;
        d[HID + j] = dgf;
        d[2 * HID + j] = dgg;
        d[3 * HID + j] = dgo;
      }
    }
    __syncthreads();
#pragma unroll
    for (int cl = 0; cl < CL; ++cl) {
      if (b0 + cl < B) gates[((size_t)(b0 + cl) * L + t) * (2 * G4) + dir * G4 + j] = s_dg[cl * G4 + j];
      float s = 0.f, s2 = 0.f;
      const float* d = s_dg + cl * G4 + part * HID;
#pragma unroll
      for (int jj = 0; jj < HID; jj += 8) {
        const float4 d0 = *reinterpret_cast<const float4*>(d + jj), d1 = *reinterpret_cast<const float4*>(d + jj + 4);
        s = fmaf(d0.w, wreg[jj + 3], fmaf(d0.z, wreg[jj + 2], fmaf(d0.y, wreg[jj + 1], fmaf(d0.x, wreg[jj], s))));
        s2 = fmaf(d1.w, wreg[jj + 7], fmaf(d1.z, wreg[jj + 6], fmaf(d1.y, wreg[jj + 5], fmaf(d1.x, wreg[jj + 4], s2))));
      }
      s_part[(cl * 4 + part) * HID + u] = s + s2;
    }
    __syncthreads();
    if (j < HID) {
#pragma unroll
      for (int cl = 0; cl < CL; ++cl) {
        const float* p = s_part + cl * 4 * HID + j;
        s_dh[cl * HID + j] = (p[0] + p[HID]) + (p[2 * HID] + p[3 * HID]);
      }
    }
    __syncthreads();
  }
}

// NHWC (B,L,Wf,C) block output -> (B,L,C*Wf) features with index c*Wf + w  (lcnn.py:196-199), and back.
__global__ void feats_gather_kernel(const float* __restrict__ act, float* __restrict__ feats, int n, int Wf, int C) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int c = i % C, w = (i / C) % Wf;
  const size_t bl = i / (C * Wf);
  feats[bl * (C * Wf) + c * Wf + w] = act[i];
}
__global__ void feats_scatter_kernel(const float* __restrict__ gfeats, float* __restrict__ gact, int n, int Wf, int C) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int c = i % C, w = (i / C) % Wf;
  const size_t bl = i / (C * Wf);
  gact[i] = gfeats[bl * (C * Wf) + c * Wf + w];
}

// logits[b] = w . mean_t(l2 + feats) + bias   (lcnn.py:205)
// feats is read at index `kp`: k itself, or its position in the NHWC block output (perm_wf > 0, see lstm_pack_kernel)
__global__ void __launch_bounds__(160) head_fwd_kernel(const float* __restrict__ l2, const float* __restrict__ feats,
                                                        const float* __restrict__ w, const float* __restrict__ bias,
                                                        float* __restrict__ logits, int L, int perm_wf) {
  __shared__ float s_red[5];
  const int b = blockIdx.x, k = threadIdx.x;
  const int kp = perm_wf > 0 ? (k % perm_wf) * (160 / perm_wf) + k / perm_wf : k;
  float s = 0.f;
  for (int t = 0; t < L; ++t) {
    const size_t o = ((size_t)b * L + t) * 160;
    s += l2[o + k] + feats[o + kp];
  }
  float v = (s / (float)L) * w[k];
  v = warp_sum(v);
  if ((k & 31) == 0) s_red[k >> 5] = v;
  __syncthreads();
  if (k == 0) logits[b] = (((s_red[0] + s_red[1]) + (s_red[2] + s_red[3])) + s_red[4]) + bias[0];
}

// d loss / d logit, then d/d(l2) = d/d(feats residual) = g_o * w / L for every time step.
__global__ void __launch_bounds__(160) head_bwd_kernel(const float* __restrict__ logits, const long long* __restrict__ y,
                                                        const float* __restrict__ w, float* __restrict__ dl2, int L,
                                                        int mode, float inv_n, const float* __restrict__ coef,
                                                        float* __restrict__ dfeat_add, int perm_wf) {
  const int b = blockIdx.x, k = threadIdx.x;
  const int kp = perm_wf > 0 ? (k % perm_wf) * (160 / perm_wf) + k / perm_wf : k;
  float go = 1.0f;
  if (mode == 2) go = coef[b];
  if (mode == 0) {
    const float o = logits[b];
    const float p1 = 1.0f / (1.0f + expf(-2.0f * o));
    go = 2.0f * (p1 - (float)y[b]) * inv_n;
  }
  const float v = go * w[k] / (float)L;
  for (int t = 0; t < L; ++t) dl2[((size_t)b * L + t) * 160 + k] = v;
  if (dfeat_add != nullptr)  // the same vector for the residual branch, in the feature buffer's own order
    for (int t = 0; t < L; ++t) dfeat_add[((size_t)b * L + t) * 160 + kp] = v;
}

}  // namespace

size_t lstm_rec_fwd_smem() { return (size_t)(CL * HID + CL * G4) * sizeof(float); }
size_t lstm_rec_bwd_smem() { return (size_t)(CL * G4 + CL * HID + CL * 4 * HID) * sizeof(float); }

int rnn_init() {
  ADVB_CUDA_OK(cudaFuncSetAttribute(lstm_rec_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)lstm_rec_fwd_smem()));
  ADVB_CUDA_OK(cudaFuncSetAttribute(lstm_rec_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)lstm_rec_bwd_smem()));
  return 0;
}

int gemm(const float* A, const float* Bm, const float* bias, const float* Cadd, float* C, int M, int N, int K,
         cudaStream_t stream, const char* tag) {
  ADVB_CHECK(K % 4 == 0 && N % 4 == 0, "gemm: K and N must be multiples of 4");
  static const int cfg = [] {  // ADVB_GEMM_RT=0: the first 4 x 4 / 4 x 2 register-tile kernel (bit-identical results, A/B only)
    const char* e = getenv("ADVB_GEMM_RT");
    return e != nullptr ? atoi(e) : 1;
  }();
  const bool narrow = cdiv(N, 64) * cdiv(M, 64) < 2 * 148;
  // measured on the BLSTM projections (B = 128: 3 200 x 640 x 160 forward, 3 200 x 160 x 640 backward): 30.5 -> 26.7 us and
  // 36.9 -> 33.3 us; (128 x 32, 8 x 2) for the narrow one: 35.9 us
  if (cfg == 1 && !narrow) {  // 128 x 64 tile, 8 x 4 per thread
    dim3 grid(cdiv(N, 64), cdiv(M, 128));
    gemm_rt_kernel<128, 64, 8, 4><<<grid, 256, 0, stream>>>(A, Bm, bias, Cadd, C, M, N, K);
    ADVB_KERNEL_OK(tag, stream);
    return 0;
  }
  if (cfg == 1 && narrow) {  // 64 x 32 tile, 4 x 4 per thread (128 threads)
    dim3 grid(cdiv(N, 32), cdiv(M, 64));
    gemm_rt_kernel<64, 32, 4, 4><<<grid, 128, 0, stream>>>(A, Bm, bias, Cadd, C, M, N, K);
    ADVB_KERNEL_OK(tag, stream);
    return 0;
  }
  if (narrow) {
    dim3 grid(cdiv(N, 32), cdiv(M, 64));
    gemm_kernel<64, 32><<<grid, 256, 0, stream>>>(A, Bm, bias, Cadd, C, M, N, K);
  } else {
    dim3 grid(cdiv(N, 64), cdiv(M, 64));
    gemm_kernel<64, 64><<<grid, 256, 0, stream>>>(A, Bm, bias, Cadd, C, M, N, K);
  }
  ADVB_KERNEL_OK(tag, stream);
  return 0;
}

size_t lstm_tc_fwd_bytes() { return gemm_pack_bytes(2 * G4, 160, 1); }
size_t lstm_tc_bwd_bytes() { return gemm_pack_bytes(256, 2 * G4, 1); }

int lstm_pack(const LstmWeights& w, const LstmPacked& p, cudaStream_t stream, int perm_wf) {
  lstm_pack_kernel<<<64, 256, 0, stream>>>(w, p, 160, perm_wf);
  ADVB_KERNEL_OK("lstm_pack", stream);
  if (p.tc_fwd != nullptr) {  // the two input projections as tcgen05 3xTF32 GEMMs: views of the (already permuted) wih_cat [640][160]
    GemmW f;
    f.w = p.wih_cat, f.s_n = 160, f.s_k = 1, f.n_valid = 2 * G4, f.k_valid = 160;
    ADVB_TRY(gemm_pack(f, 2 * G4, 160, 1, p.tc_fwd, stream));
    GemmW b;
    b.w = p.wih_cat, b.s_n = 1, b.s_k = 160, b.n_valid = 160, b.k_valid = 2 * G4;
    ADVB_TRY(gemm_pack(b, 256, 2 * G4, 1, p.tc_bwd, stream));
  }
  return 0;
}

// The input projections are 0.66 GFLOP GEMMs (3 200 x 640 x 160): 30 / 37 us each on the fp32 SIMT kernel, i.e. 135 us of every
// PGD iteration; on the persistent tcgen05 GEMM of gemm_tc.cu (3xTF32, fp32-class) they are one wave of 125 / 50 tiles.
static int lstm_proj_tc(const float* A, int K, const unsigned char* wpack, int N, int n_store, const float* bias, const float* add,
                        float* out, int M, int passes, cudaStream_t stream, const char* tag) {
  GemmArgs a;
  a.A = A, a.lda = K, a.M = M, a.K = K, a.ntap = 1;
  a.wpack = wpack, a.N = N, a.n_store = n_store;
  a.Tp = 1, a.pad = 0, a.Tv = 1;
  a.bias = bias;
  a.add = add, a.ld_add = n_store > 0 ? n_store : N;
  a.out = out, a.ldc = n_store > 0 ? n_store : N;
  a.tag = tag;
  return gemm_run(a, 0, passes, stream);
}

int blstm_forward(const LstmPacked& p, const float* x, float* gates, float* hout, float* cs, int B, int L,
                  cudaStream_t stream) {
  if (p.tc_fwd != nullptr)
    ADVB_TRY(lstm_proj_tc(x, 160, p.tc_fwd, 2 * G4, 0, p.bias, nullptr, gates, B * L, 3, stream, "lstm_inproj_fwd"));
  else
    ADVB_TRY(gemm(x, p.wihT, p.bias, nullptr, gates, B * L, 2 * G4, 160, stream, "lstm_inproj_fwd"));
  dim3 grid(cdiv(B, CL), 2);
  lstm_rec_fwd_kernel<<<grid, G4, lstm_rec_fwd_smem(), stream>>>(gates, p.whhT, hout, cs, B, L);
  ADVB_KERNEL_OK("lstm_rec_fwd", stream);
  return 0;
}

int blstm_backward(const LstmPacked& p, float* gates, const float* dout, const float* cs, const float* dx_add,
                   float* dx, int B, int L, cudaStream_t stream) {
  dim3 grid(cdiv(B, CL), 2);
  lstm_rec_bwd_kernel<<<grid, G4, lstm_rec_bwd_smem(), stream>>>(gates, p.whh, dout, cs, B, L);
  ADVB_KERNEL_OK("lstm_rec_bwd", stream);
  if (p.tc_bwd != nullptr)
    ADVB_TRY(lstm_proj_tc(gates, 2 * G4, p.tc_bwd, 256, 160, nullptr, dx_add, dx, B * L, 3, stream, "lstm_inproj_bwd"));
  else
    ADVB_TRY(gemm(gates, p.wih_cat, nullptr, dx_add, dx, B * L, 160, 2 * G4, stream, "lstm_inproj_bwd"));
  return 0;
}

int feats_gather(const float* act, float* feats, int B, int L, int Wf, int C, cudaStream_t stream) {
  const int n = B * L * Wf * C;
  feats_gather_kernel<<<cdiv(n, 256), 256, 0, stream>>>(act, feats, n, Wf, C);
  ADVB_KERNEL_OK("feats_gather", stream);
  return 0;
}
int feats_scatter(const float* gfeats, float* gact, int B, int L, int Wf, int C, cudaStream_t stream) {
  const int n = B * L * Wf * C;
  feats_scatter_kernel<<<cdiv(n, 256), 256, 0, stream>>>(gfeats, gact, n, Wf, C);
  ADVB_KERNEL_OK("feats_scatter", stream);
  return 0;
}

int head_forward(const float* l2, const float* feats, const float* w, const float* bias, float* logits, int B, int L,
                 cudaStream_t stream, int perm_wf) {
  head_fwd_kernel<<<B, 160, 0, stream>>>(l2, feats, w, bias, logits, L, perm_wf);
  ADVB_KERNEL_OK("head_fwd", stream);
  return 0;
}
int head_backward(const float* logits, const long long* y, const float* w, float* dl2, int B, int L, int mode,
                  int n_global, cudaStream_t stream, const float* coef, float* dfeat_add, int perm_wf) {
  head_bwd_kernel<<<B, 160, 0, stream>>>(logits, y, w, dl2, L, mode, 1.0f / (float)n_global, coef, dfeat_add, perm_wf);
  ADVB_KERNEL_OK("head_bwd", stream);
  return 0;
}

}  // namespace advb
