// FAB (Fast Adaptive Boundary, L-inf / L2, untargeted, 2 classes) and CW (Carlini-Wagner L2) update kernels (sm_100a).
//
// Replaces the tensor arithmetic of adversarial_attacks/torchattacks/attacks/fab.py:131-307 (attack_single_run),
// :562-614 (projection_linf) and cw.py:46-134.  The model forward / input-gradient backward between these kernels is
// the same engine path FGSM/PGD use; the per-step control flow of the reference that syncs with the host
// (`is_adv.sum() > 0`, `c_l.any()`, index compaction) is replaced by per-row predication, so one FAB step is a fixed
// sequence of launches with no host round trip.
//
// projection_linf without a sort.  The reference sorts p = |distance to the box face in the direction -sign(w)|,
// builds cumulative sums in that order and bisects over the sorted INDEX for the face-saturation level lambda
// (fab.py:575-612).  In closed form its cumulative sums are   b2(k) = sb[k] - s[k] p_(k) = -sum_j |w_j| min(p_j, p_(k)),
// so the bisection solves the monotone piecewise-linear equation  G(lambda) = -sum_j |w_j| min(p_j, lambda) = b  and
// the final formula (b - sb[lb]) / (-s[lb]) is the exact root on the linear piece that contains it.  Here each row is
// one CTA that searches lambda directly over the float bit patterns of [p_min, p_max] (8-ary: 7 thresholds per pass over
// the L2-resident row, <= 11 passes) and then evaluates the same closed-form root.  The result is independent of the
// order of equal keys, which torch.argsort leaves unspecified (SURVEY.md "hard parts"), and every reduction has a fixed
// order (deterministic).  Algorithmic bytes per row: read t, w (8 B/sample), write d (4 B/sample).
#include "fabcw.cuh"

#include <math.h>

namespace advb {

namespace {

constexpr int RT = 1024;  // threads of a one-CTA-per-row kernel
constexpr int CW_CHUNKS = 8;

struct OpSum {
  __device__ static float apply(float a, float b) { return a + b; }
  __device__ static float identity() { return 0.f; }
};
struct OpMin {
  __device__ static float apply(float a, float b) { return fminf(a, b); }
  __device__ static float identity() { return INFINITY; }
};
struct OpMax {
  __device__ static float apply(float a, float b) { return fmaxf(a, b); }
  __device__ static float identity() { return -INFINITY; }
};

// fixed-order block reduction; the result is returned to every thread.  s_red: 33 floats.
template <typename Op>
__device__ __forceinline__ float block_reduce(float v, float* s_red) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = Op::apply(v, __shfl_xor_sync(0xffffffffu, v, o));
  __syncthreads();  // protect s_red from the previous use
  if (lane == 0) s_red[warp] = v;
  __syncthreads();
  if (warp == 0) {
    float r = lane < nw ? s_red[lane] : Op::identity();
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) r = Op::apply(r, __shfl_xor_sync(0xffffffffu, r, o));
    if (lane == 0) s_red[32] = r;
  }
  __syncthreads();
  return s_red[32];
}

__device__ __forceinline__ float clamp01(float v) { return fminf(fmaxf(v, 0.f), 1.f); }

// ---------------------------------------------------------------------------------------------------------
// FAB
// ---------------------------------------------------------------------------------------------------------
__global__ void fab_init_kernel(const float* __restrict__ x, float* __restrict__ adv, float* __restrict__ x1,
                                float* __restrict__ res2, int B, int64_t n) {
  const int64_t i0 = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  for (int64_t i = i0; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float v = x[i];
    adv[i] = v;
    x1[i] = v;
  }
  if (i0 < B) res2[i0] = 1e10f;
}

__global__ void __launch_bounds__(RT) fab_hyperplane_kernel(const float* __restrict__ g, const float* __restrict__ logits,
                                                             const long long* __restrict__ y, const float* __restrict__ x1,
                                                             float* __restrict__ w, float* __restrict__ bh, int T) {
  __shared__ float s_red[33];
  const int r = blockIdx.x;
  const float c = y[r] == 1 ? -2.0f : 2.0f;  // dg = g_other - g_label = -+2 do/dx ; df = z_other - z_label = -+2 o
  float acc = 0.f;
  for (int i = threadIdx.x; i < T; i += RT) {
    const size_t o = (size_t)r * T + i;
    const float wi = c * g[o];
    w[o] = wi;
    acc = fmaf(wi, x1[o], acc);
  }
  acc = block_reduce<OpSum>(acc, s_red);
  if (threadIdx.x == 0) bh[r] = -(c * logits[r]) + acc;
}

constexpr int NTH = 7;  // interior thresholds per search pass

__global__ void __launch_bounds__(RT) fab_project_kernel(const float* __restrict__ x1, const float* __restrict__ x0,
                                                          const float* __restrict__ wmat, const float* __restrict__ bh,
                                                          float* __restrict__ d3, float* __restrict__ a0, int B, int T) {
  __shared__ float s_red[33];
  const int r = blockIdx.x, rb = r < B ? r : r - B;
  const float* t = (r < B ? x1 : x0) + (size_t)rb * T;
  const float* w = wmat + (size_t)rb * T;
  float* d = d3 + (size_t)r * T;
  const int tid = threadIdx.x;

  // pass 1: <w, t>, sum |w|
  float wt = 0.f, S = 0.f;
  for (int i = tid; i < T; i += RT) {
    const float wi = w[i];
    wt = fmaf(wi, t[i], wt);
    S += fabsf(wi);
  }
  wt = block_reduce<OpSum>(wt, s_red);
  S = block_reduce<OpSum>(S, s_red);
  const float b_in = bh[rb];
  const float sgn = (wt - b_in >= 0.f) ? 1.f : -1.f;   // fab.py:566
  const float b = sgn * b_in - sgn * wt;               // fab.py:567-568,576  (<= 0)

  // pass 2: WP = sum |w| p, p range.  a = (sgn w < 0); p = a ? 1 - t : t           (fab.py:570-573)
  float WP = 0.f, pmin = INFINITY, pmax = -INFINITY;
  for (int i = tid; i < T; i += RT) {
    const float wi = sgn * w[i], ti = t[i];
    const float p = wi < 0.f ? 1.f - ti : ti;
    WP = fmaf(fabsf(wi), p, WP);
    pmin = fminf(pmin, p);
    pmax = fmaxf(pmax, p);
  }
  WP = block_reduce<OpSum>(WP, s_red);
  pmin = block_reduce<OpMin>(pmin, s_red);
  pmax = block_reduce<OpMax>(pmax, s_red);
  const float b0 = -WP;                                  // <w, d> with d = (a - t)(w != 0)          (fab.py:577)
  // G(tau) = -sum |w| min(p, tau):  G(p_min) = -S p_min is the reference's b2 of :586, G(p_max) = b0
  const bool c_l = b - (-S * pmin) > 0.f;                // :587
  const bool c2 = (b - b0 > 0.f) && !c_l;                // :588

  float lambda = 0.f;
  int mode = 0;  // 0: keep the box corner d, 1: uniform level (c_l), 2: saturating level (c2)
  if (c_l) {
    mode = 1;
    lambda = fmaxf(b / (-S), 0.f);                       // :607-608 with sb[-1] = b0 + WP = 0
  } else if (c2) {
    mode = 2;
    // invariant: cond(lo) false, cond(hi) true, cond(tau) := b - G(tau) > 0
    unsigned lo = __float_as_uint(fmaxf(pmin, 0.f)), hi = __float_as_uint(fmaxf(pmax, 0.f));
    while (hi - lo > 1u) {
      const unsigned width = hi - lo;
      float th[NTH], acc[NTH];
#pragma unroll
      for (int k = 0; k < NTH; ++k) {
        th[k] = __uint_as_float(lo + (unsigned)(((unsigned long long)width * (k + 1)) >> 3));
        acc[k] = 0.f;
      }
      for (int i = tid; i < T; i += RT) {
        const float wi = sgn * w[i], ti = t[i];
        const float p = wi < 0.f ? 1.f - ti : ti, aw = fabsf(wi);
#pragma unroll
        for (int k = 0; k < NTH; ++k) acc[k] = fmaf(aw, fminf(p, th[k]), acc[k]);
      }
      unsigned new_lo = lo, new_hi = hi;
      bool found = false;
#pragma unroll
      for (int k = 0; k < NTH; ++k) {
        const float Gk = -block_reduce<OpSum>(acc[k], s_red);
        const unsigned bits = __float_as_uint(th[k]);
        if (!found) {
          if (b - Gk > 0.f && bits > lo) {
            new_hi = bits;
            found = true;
          } else {
            new_lo = bits;
          }
        }
      }
      lo = new_lo;
      hi = new_hi;
    }
    // root on the linear piece {p >= hi saturate at lambda}: -(sum_{p<hi} |w| p + lambda sum_{p>=hi} |w|) = b
    const float hf = __uint_as_float(hi);
    float below = 0.f, Sh = 0.f;
    for (int i = tid; i < T; i += RT) {
      const float wi = sgn * w[i], ti = t[i];
      const float p = wi < 0.f ? 1.f - ti : ti, aw = fabsf(wi);
      if (p >= hf) Sh += aw;
      else below = fmaf(aw, p, below);
    }
    below = block_reduce<OpSum>(below, s_red);
    Sh = block_reduce<OpSum>(Sh, s_red);
    lambda = fmaxf((-below - b) / Sh, 0.f);              // :611 == (b - sb[lb]) / (-s[lb])
  }

  float amax = 0.f;
  for (int i = tid; i < T; i += RT) {
    const float wi = sgn * w[i], ti = t[i];
    const bool a = wi < 0.f;
    float di = (a ? 1.f - ti : -ti);                     // (a - t)
    if (mode == 1) di = a ? lambda : -lambda;            // (2a - 1) lambda
    else if (mode == 2) di = a ? fminf(lambda, di) : fmaxf(-lambda, di);  // :612
    if (wi == 0.f) di = 0.f;                             // :571,614
    d[i] = di;
    amax = fmaxf(amax, fabsf(di));
  }
  amax = block_reduce<OpMax>(amax, s_red);
  if (tid == 0) a0[r] = amax;
}

// projection_l2 (fab.py:617-665) without a sort.  After the sign flip (c = <w,t> - b >= 0) the projection is
//   d_i = -min(alpha, r_i) w_i,   r_i = max(t_i / w_i, (t_i - 1) / w_i)   (the alpha at which coordinate i reaches its box face),
// with alpha the root of the monotone piecewise-linear  H(alpha) = sum_i w_i^2 min(alpha, r_i) = c.  The reference sorts r,
// builds s[k] = -H(rs[k]) by cumulative sums and bisects over the sorted INDEX; its three branches are
//   c4: c < w5 r_min            -> alpha = c / w5 (no coordinate saturates)                 (:644,659-661)
//   c3: c > H(inf)              -> every coordinate at its face (the hyperplane is out of the box's reach)  (:645)
//   c2: otherwise               -> alpha = (s[lb] + c) / ws[lb] + rs[lb] on the linear piece that contains the root (:662-666)
// Here, as in the L-inf kernel above: one CTA per row, an 8-ary search over the float bit patterns of [r_min, r_max] for the
// piece, then the same closed-form root.  Coordinates with |w| < 1e-8 never move (:627,667).
__device__ __forceinline__ float fab_l2_r(float wi, float ti) {
  if (fabsf(wi) < 1e-8f) return 1e12f;                                             // :627
  float r = fmaxf(__fdiv_rn(ti, wi), __fdiv_rn(ti - 1.f, wi));                      // :626
  r = fminf(fmaxf(r, -1e12f), 1e12f);
  return r == -1e12f ? 1e12f : r;                                                  // :628
}

__global__ void __launch_bounds__(RT) fab_project_l2_kernel(const float* __restrict__ x1, const float* __restrict__ x0,
                                                             const float* __restrict__ wmat, const float* __restrict__ bh,
                                                             float* __restrict__ d3, float* __restrict__ a0, int B, int T) {
  __shared__ float s_red[33];
  const int r = blockIdx.x, rb = r < B ? r : r - B;
  const float* t = (r < B ? x1 : x0) + (size_t)rb * T;
  const float* w = wmat + (size_t)rb * T;
  float* d = d3 + (size_t)r * T;
  const int tid = threadIdx.x;

  float wt = 0.f, w5 = 0.f;
  for (int i = tid; i < T; i += RT) {
    const float wi = w[i];
    wt = fmaf(wi, t[i], wt);
    w5 = fmaf(wi, wi, w5);
  }
  wt = block_reduce<OpSum>(wt, s_red);
  w5 = block_reduce<OpSum>(w5, s_red);
  const float c_in = wt - bh[rb];
  const float sgn = c_in >= 0.f ? 1.f : -1.f;   // :622
  const float c = sgn * c_in;                    // :624  (>= 0)

  // H(inf) over the movable coordinates, range of r over them
  float Hall = 0.f, rmin = INFINITY, rmax = -INFINITY;
  for (int i = tid; i < T; i += RT) {
    const float wi = sgn * w[i];
    const float ri = fab_l2_r(wi, t[i]);
    if (ri < 1e12f) {
      Hall = fmaf(wi * wi, ri, Hall);
      rmin = fminf(rmin, ri);
      rmax = fmaxf(rmax, ri);
    }
  }
  Hall = block_reduce<OpSum>(Hall, s_red);
  rmin = block_reduce<OpMin>(rmin, s_red);
  rmax = block_reduce<OpMax>(rmax, s_red);
  const bool any_movable = rmin <= rmax;
  const bool c4 = any_movable ? (c - w5 * rmin < 0.f) : true;   // s[:,0] + c < 0 with s[:,0] = -w5 rs[0]
  const bool c3 = !c4 && (c - Hall > 0.f);                       // <d,w> + c > 0 with every coordinate at its face
  float alpha = 0.f;
  int mode = 0;  // 0: all faces (c3), 1: uniform alpha (c4), 2: saturating alpha (c2)
  if (c4) {
    mode = 1;
    alpha = w5 > 0.f ? c / w5 : 0.f;
  } else if (!c3) {
    mode = 2;
    // invariant: H(lo) <= c < ... cond(tau) := H(tau) - c >= 0 is false at lo and true at hi
    unsigned lo = __float_as_uint(fmaxf(rmin, 0.f)), hi = __float_as_uint(fmaxf(rmax, 0.f));
    while (hi - lo > 1u) {
      const unsigned width = hi - lo;
      float th[NTH], acc[NTH];
#pragma unroll
      for (int k = 0; k < NTH; ++k) {
        th[k] = __uint_as_float(lo + (unsigned)(((unsigned long long)width * (k + 1)) >> 3));
        acc[k] = 0.f;
      }
      for (int i = tid; i < T; i += RT) {
        const float wi = sgn * w[i];
        const float ri = fab_l2_r(wi, t[i]), w2 = wi * wi;
        if (ri < 1e12f) {
#pragma unroll
          for (int k = 0; k < NTH; ++k) acc[k] = fmaf(w2, fminf(ri, th[k]), acc[k]);
        }
      }
      unsigned new_lo = lo, new_hi = hi;
      bool found = false;
#pragma unroll
      for (int k = 0; k < NTH; ++k) {
        const float Hk = block_reduce<OpSum>(acc[k], s_red);
        const unsigned bits = __float_as_uint(th[k]);
        if (!found) {
          if (Hk - c >= 0.f && bits > lo) {
            new_hi = bits;
            found = true;
          } else {
            new_lo = bits;
          }
        }
      }
      lo = new_lo;
      hi = new_hi;
    }
    // root on the linear piece {r >= hi still free}: sum_{r < hi} w^2 r + alpha sum_{r >= hi} w^2 = c
    const float hf = __uint_as_float(hi);
    float below = 0.f, Wh = 0.f;
    for (int i = tid; i < T; i += RT) {
      const float wi = sgn * w[i];
      const float ri = fab_l2_r(wi, t[i]), w2 = wi * wi;
      if (ri >= hf) Wh += w2;  // includes the immovable coordinates, as the reference's ws does (their w^2 < 1e-16)
      else below = fmaf(w2, ri, below);
    }
    below = block_reduce<OpSum>(below, s_red);
    Wh = block_reduce<OpSum>(Wh, s_red);
    alpha = Wh > 0.f ? (c - below) / Wh : 0.f;               // :663-664
  }

  float ss = 0.f;
  for (int i = tid; i < T; i += RT) {
    const float wi = sgn * w[i];
    const float ri = fab_l2_r(wi, t[i]);
    float di = 0.f;
    if (fabsf(wi) > 1e-8f) {
      const float face = -__fmul_rn(ri, wi);                   // :637
      if (mode == 0) di = face;
      else if (mode == 1) di = -__fmul_rn(alpha, wi);          // :661
      else di = alpha > ri ? face : -__fmul_rn(alpha, wi);     // :665-666
    }
    d[i] = di;
    ss = fmaf(di, di, ss);
  }
  ss = block_reduce<OpSum>(ss, s_red);
  if (tid == 0) a0[r] = sqrtf(ss);
}

__global__ void fab_combine_kernel(const float* __restrict__ x0, float* __restrict__ x1, const float* __restrict__ d3,
                                   const float* __restrict__ a0, float eta, float alpha_max, int B, int T) {
  const int r = blockIdx.y;
  const float a1 = fmaxf(a0[r], 1e-8f), a2 = fmaxf(a0[B + r], 1e-8f);   // fab.py:259-262
  const float alpha = fminf(fmaxf(__fdiv_rn(a1, __fadd_rn(a1, a2)), 0.f), alpha_max);
  const float one_m = __fsub_rn(1.f, alpha);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < T; i += gridDim.x * blockDim.x) {
    const size_t o = (size_t)r * T + i;
    const float u = __fmul_rn(__fadd_rn(x1[o], __fmul_rn(eta, d3[o])), one_m);
    const float v = __fmul_rn(__fadd_rn(x0[o], __fmul_rn(d3[(size_t)B * T + o], eta)), alpha);
    x1[o] = clamp01(__fadd_rn(u, v));                                   // :266-267
  }
}

__global__ void __launch_bounds__(RT) fab_bookkeep_kernel(const float* __restrict__ x0, const float* __restrict__ logits,
                                                           const long long* __restrict__ y, float* __restrict__ adv,
                                                           float* __restrict__ x1, float* __restrict__ res2, float beta,
                                                           int T, int norm_l2) {
  __shared__ float s_red[33];
  const int r = blockIdx.x;
  const long long pred = logits[r] > 0.f ? 1 : 0;  // argmax of [-o, o]; a tie goes to class 0 (torch.max)
  if (pred == y[r]) return;                        // not adversarial: nothing changes (fab.py:269-271)
  float tmax = 0.f;
  if (norm_l2) {                                   // :277-279
    for (int i = threadIdx.x; i < T; i += RT) {
      const size_t o = (size_t)r * T + i;
      const float df = __fsub_rn(x1[o], x0[o]);
      tmax = fmaf(df, df, tmax);
    }
    tmax = sqrtf(block_reduce<OpSum>(tmax, s_red));
  } else {                                         // :274-276
    for (int i = threadIdx.x; i < T; i += RT) {
      const size_t o = (size_t)r * T + i;
      tmax = fmaxf(tmax, fabsf(__fsub_rn(x1[o], x0[o])));
    }
    tmax = block_reduce<OpMax>(tmax, s_red);
  }
  const bool better = tmax < res2[r];
  for (int i = threadIdx.x; i < T; i += RT) {
    const size_t o = (size_t)r * T + i;
    const float v = x1[o], c = x0[o];
    if (better) adv[o] = v;                                             // :282-285
    x1[o] = __fadd_rn(c, __fmul_rn(__fsub_rn(v, c), beta));              // :288-289
  }
  __syncthreads();
  if (threadIdx.x == 0 && better) res2[r] = tmax;
}

// ---------------------------------------------------------------------------------------------------------
// CW
// ---------------------------------------------------------------------------------------------------------
__global__ void cw_init_kernel(const float* __restrict__ x, float* __restrict__ best_adv, float* __restrict__ w,
                               float* __restrict__ m, float* __restrict__ v, float* __restrict__ best_l2, int B,
                               int64_t n) {
  const int64_t i0 = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  for (int64_t i = i0; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float xi = x[i];
    const float u = __fsub_rn(__fmul_rn(xi, 2.f), 1.f);
    w[i] = 0.5f * logf(__fdiv_rn(__fadd_rn(1.f, u), __fsub_rn(1.f, u)));  // cw.py:117-123 (+-inf at x in {0,1})
    m[i] = 0.f;
    v[i] = 0.f;
    best_adv[i] = xi;
  }
  if (i0 < B) best_l2[i0] = 1e10f;
}

__global__ void __launch_bounds__(256) cw_image_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                        float* __restrict__ adv, float* __restrict__ partial, int T) {
  __shared__ float s_red[33];
  const int b = blockIdx.y, ch = blockIdx.x;
  const int len = (T + CW_CHUNKS - 1) / CW_CHUNKS;
  const int lo = ch * len, hi = min(T, lo + len);
  float s = 0.f;
  for (int i = lo + threadIdx.x; i < hi; i += 256) {
    const size_t o = (size_t)b * T + i;
    const float a = __fmul_rn(0.5f, __fadd_rn(tanhf(w[o]), 1.f));       // cw.py:114-115
    adv[o] = a;
    const float d = __fsub_rn(a, x[o]);
    s = fmaf(d, d, s);
  }
  s = block_reduce<OpSum>(s, s_red);
  if (threadIdx.x == 0) partial[b * CW_CHUNKS + ch] = s;
}

__global__ void __launch_bounds__(256) cw_head_kernel(const float* __restrict__ logits, const long long* __restrict__ y,
                                                       const long long* __restrict__ y_target,
                                                       const float* __restrict__ partial, float* __restrict__ cur_l2,
                                                       float* __restrict__ best_l2, float* __restrict__ coef,
                                                       float* __restrict__ mask, float* __restrict__ cost, float c,
                                                       float kappa, int B) {
  __shared__ float s_red[33];
  float l2_sum = 0.f, f_sum = 0.f;
  for (int b = threadIdx.x; b < B; b += 256) {
    float l2 = 0.f;
    for (int k = 0; k < CW_CHUNKS; ++k) l2 += partial[b * CW_CHUNKS + k];
    cur_l2[b] = l2;
    const float o = logits[b];
    // z = [-o, o]; j = z[y]; i = max((1 - onehot) * z) = max(z[other], 0)      (cw.py:125-134)
    // targeted (cw.py:82-83,131-132): the one-hot is built from the target labels and f = clamp(i - j, min=-kappa);
    // the best-adversarial bookkeeping below always compares with the given labels (cw.py:95)
    const bool tgt = y_target != nullptr;
    const bool y1 = (tgt ? y_target[b] : y[b]) == 1;
    const float zj = y1 ? o : -o, zo = y1 ? -o : o;
    const float fi = fmaxf(zo, 0.f);
    const float diff = tgt ? fi - zj : zj - fi;
    float f = diff, df = y1 ? 1.f : -1.f;                 // d zj / d o
    if (zo > 0.f) df += y1 ? 1.f : -1.f;                  // - d zo / d o when the other logit is the max
    if (tgt) df = -df;
    if (diff < -kappa) {                                  // clamp(min=-kappa): value -kappa, zero gradient
      f = -kappa;
      df = 0.f;
    }
    coef[b] = c * df;
    const bool correct = ((o > 0.f) ? 1 : 0) == ((y[b] == 1) ? 1 : 0);
    const bool take = !correct && best_l2[b] > l2;        // cw.py:94-101
    mask[b] = take ? 1.f : 0.f;
    if (take) best_l2[b] = l2;
    l2_sum += l2;
    f_sum += f;
  }
  l2_sum = block_reduce<OpSum>(l2_sum, s_red);
  f_sum = block_reduce<OpSum>(f_sum, s_red);
  if (threadIdx.x == 0) *cost = l2_sum + c * f_sum;       // cw.py:88
}

__global__ void cw_adam_kernel(const float* __restrict__ x, const float* __restrict__ g_model,
                               const float* __restrict__ adv, float* __restrict__ best_adv, float* __restrict__ w,
                               float* __restrict__ m, float* __restrict__ v, const float* __restrict__ mask,
                               float step_size, float bc2_sqrt, int T) {
  const int b = blockIdx.y;
  const bool take = mask[b] != 0.f;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < T; i += gridDim.x * blockDim.x) {
    const size_t o = (size_t)b * T + i;
    const float a = adv[o];
    if (take) best_adv[o] = a;
    // d cost / d adv = 2 (adv - x) + c f'(o) d o / d adv ;  adv = 1/2 (tanh w + 1)
    const float tw = tanhf(w[o]);
    const float ga = __fadd_rn(__fmul_rn(2.f, __fsub_rn(a, x[o])), g_model[o]);
    const float gw = __fmul_rn(__fmul_rn(ga, 0.5f), __fsub_rn(1.f, __fmul_rn(tw, tw)));
    // torch.optim.Adam (single-tensor): lerp, addcmul, sqrt / bias-correction + eps, addcdiv
    const float mi = __fadd_rn(m[o], __fmul_rn(0.1f, __fsub_rn(gw, m[o])));
    const float vi = __fadd_rn(__fmul_rn(v[o], 0.999f), __fmul_rn(__fmul_rn(0.001f, gw), gw));
    m[o] = mi;
    v[o] = vi;
    const float denom = __fadd_rn(__fdiv_rn(sqrtf(vi), bc2_sqrt), 1e-8f);
    w[o] = __fsub_rn(w[o], __fmul_rn(step_size, __fdiv_rn(mi, denom)));
  }
}

__global__ void __launch_bounds__(RT) row_diff_norms_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                                             float* __restrict__ linf, float* __restrict__ l2, int T) {
  __shared__ float s_red[33];
  const int r = blockIdx.x;
  float mx = 0.f, ss = 0.f;
  for (int i = threadIdx.x; i < T; i += RT) {
    const float d = __fsub_rn(a[(size_t)r * T + i], b[(size_t)r * T + i]);
    mx = fmaxf(mx, fabsf(d));
    ss = fmaf(d, d, ss);
  }
  mx = block_reduce<OpMax>(mx, s_red);
  ss = block_reduce<OpSum>(ss, s_red);
  if (threadIdx.x == 0) {
    if (linf != nullptr) linf[r] = mx;
    if (l2 != nullptr) l2[r] = sqrtf(ss);
  }
}

inline int ew_blocks(int64_t n) {
  int64_t b = (n + 255) / 256;
  return (int)(b > 148 * 16 ? 148 * 16 : b);
}

}  // namespace

int fab_init(const float* x, float* adv, const FabScratch& s, int B, int T, cudaStream_t stream) {
  fab_init_kernel<<<ew_blocks((int64_t)B * T), 256, 0, stream>>>(x, adv, s.x1, s.res2, B, (int64_t)B * T);
  ADVB_KERNEL_OK("fab_init", stream);
  return 0;
}
int fab_hyperplane(const float* g, const float* logits, const long long* y, const FabScratch& s, int B, int T,
                   cudaStream_t stream) {
  fab_hyperplane_kernel<<<B, RT, 0, stream>>>(g, logits, y, s.x1, s.w, s.bh, T);
  ADVB_KERNEL_OK("fab_hyperplane", stream);
  return 0;
}
int fab_project(const float* x0, const FabScratch& s, int B, int T, int norm_l2, cudaStream_t stream) {
  if (norm_l2) fab_project_l2_kernel<<<2 * B, RT, 0, stream>>>(s.x1, x0, s.w, s.bh, s.d3, s.a0, B, T);
  else fab_project_kernel<<<2 * B, RT, 0, stream>>>(s.x1, x0, s.w, s.bh, s.d3, s.a0, B, T);
  ADVB_KERNEL_OK(norm_l2 ? "fab_project_l2" : "fab_project", stream);
  return 0;
}
int fab_project_rows(const FabScratch& s, int R, int T, int norm_l2, cudaStream_t stream) {
  // r < B = R for every row
  if (norm_l2) fab_project_l2_kernel<<<R, RT, 0, stream>>>(s.x1, s.x1, s.w, s.bh, s.d3, s.a0, R, T);
  else fab_project_kernel<<<R, RT, 0, stream>>>(s.x1, s.x1, s.w, s.bh, s.d3, s.a0, R, T);
  ADVB_KERNEL_OK(norm_l2 ? "fab_project_l2" : "fab_project", stream);
  return 0;
}
int fab_combine(const float* x0, const FabScratch& s, float eta, float alpha_max, int B, int T, cudaStream_t stream) {
  dim3 grid(cdiv(T, 256 * 8), B);
  fab_combine_kernel<<<grid, 256, 0, stream>>>(x0, s.x1, s.d3, s.a0, eta, alpha_max, B, T);
  ADVB_KERNEL_OK("fab_combine", stream);
  return 0;
}
int fab_bookkeep(const float* x0, const float* logits, const long long* y, float* adv, const FabScratch& s, float beta,
                 int B, int T, int norm_l2, cudaStream_t stream) {
  fab_bookkeep_kernel<<<B, RT, 0, stream>>>(x0, logits, y, adv, s.x1, s.res2, beta, T, norm_l2);
  ADVB_KERNEL_OK("fab_bookkeep", stream);
  return 0;
}

int cw_init(const float* x, float* best_adv, const CwScratch& s, int B, int T, cudaStream_t stream) {
  cw_init_kernel<<<ew_blocks((int64_t)B * T), 256, 0, stream>>>(x, best_adv, s.w, s.m, s.v, s.best_l2, B, (int64_t)B * T);
  ADVB_KERNEL_OK("cw_init", stream);
  return 0;
}
int cw_forward_image(const float* x, const CwScratch& s, int B, int T, cudaStream_t stream) {
  dim3 grid(CW_CHUNKS, B);
  cw_image_kernel<<<grid, 256, 0, stream>>>(x, s.w, s.adv, s.l2_partial, T);
  ADVB_KERNEL_OK("cw_image", stream);
  return 0;
}
int cw_head(const float* logits, const long long* y, const long long* y_target, const CwScratch& s, float c, float kappa, int B,
            cudaStream_t stream) {
  cw_head_kernel<<<1, 256, 0, stream>>>(logits, y, y_target, s.l2_partial, s.cur_l2, s.best_l2, s.coef, s.mask, s.cost, c, kappa, B);
  ADVB_KERNEL_OK("cw_head", stream);
  return 0;
}
int cw_adam(const float* x, const float* g_model, float* best_adv, const CwScratch& s, float lr, int step, int B, int T,
            cudaStream_t stream) {
  // bias corrections in double like torch (python floats), applied as fp32 scalars
  const double bc1 = 1.0 - pow(0.9, (double)step), bc2 = 1.0 - pow(0.999, (double)step);
  dim3 grid(cdiv(T, 256 * 8), B);
  cw_adam_kernel<<<grid, 256, 0, stream>>>(x, g_model, s.adv, best_adv, s.w, s.m, s.v, s.mask, (float)((double)lr / bc1),
                                           (float)sqrt(bc2), T);
  ADVB_KERNEL_OK("cw_adam", stream);
  return 0;
}

int row_diff_norms(const float* a, const float* b, float* linf, float* l2, int B, int T, cudaStream_t stream) {
  row_diff_norms_kernel<<<B, RT, 0, stream>>>(a, b, linf, l2, T);
  ADVB_KERNEL_OK("row_diff_norms", stream);
  return 0;
}

}  // namespace advb
