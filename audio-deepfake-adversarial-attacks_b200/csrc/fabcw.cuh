// FAB (L-inf / L2, untargeted, 2 classes) and CW kernels — see fabcw.cu.
#pragma once
#include "common.cuh"

namespace advb {

struct FabScratch {
  float* x1;      // (B,T) current iterate
  float* d3;      // (2B,T) projections of x1 (rows 0..B-1) and of the original clip (rows B..2B-1)
  float* w;       // (B,T) hyperplane normal  dg = +-2 d o / d x1
  float* bh;      // (B)   hyperplane offset b
  float* a0;      // (2B)  L-inf (L2) norm of each projection row
  float* res2;    // (B)   best L-inf (L2) distance of an adversarial iterate so far (1e10 = none)
};

// x1 <- x, adv <- x, res2 <- 1e10                                                   (fab.py:165-170)
int fab_init(const float* x, float* adv, const FabScratch& s, int B, int T, cudaStream_t stream);
// w = c * g, b = -df + <w, x1>  with (c, df) = (-2, -2 o) for label 1 and (+2, +2 o) for label 0   (fab.py:108-110,226-229)
int fab_hyperplane(const float* g, const float* logits, const long long* y, const FabScratch& s, int B, int T,
                   cudaStream_t stream);
// d3 = projection_linf(cat(x1, x0), cat(w, w), cat(b, b)); a0 = row L-inf norms      (fab.py:232-235,562-614)
// norm_l2: projection_l2 and row L2 norms instead                                  (fab.py:236-240,251-253,617-665)
int fab_project(const float* x0, const FabScratch& s, int B, int T, int norm_l2, cudaStream_t stream);
// R independent rows: t = s.x1, w = s.w, b = s.bh -> d = s.d3, a0 = s.a0 (test entries advb_projection_linf / _l2)
int fab_project_rows(const FabScratch& s, int R, int T, int norm_l2, cudaStream_t stream);
// alpha = clamp(a1 / (a1 + a2), 0, alpha_max); x1 = clamp((x1 + eta d1)(1 - alpha) + (x0 + eta d2) alpha, 0, 1)   (fab.py:249-267)
int fab_combine(const float* x0, const FabScratch& s, float eta, float alpha_max, int B, int T, cudaStream_t stream);
// rows whose new prediction differs from the label: keep the closest one in adv/res2, step back by beta (fab.py:269-290)
int fab_bookkeep(const float* x0, const float* logits, const long long* y, float* adv, const FabScratch& s, float beta,
                 int B, int T, int norm_l2, cudaStream_t stream);

struct CwScratch {
  float *w, *m, *v;   // (B,T) tanh-space variable and Adam moments
  float* adv;         // (B,T) current adversarial batch  1/2 (tanh w + 1)
  float* l2_partial;  // (B, ROW_CHUNKS_CW)
  float* cur_l2;      // (B)
  float* best_l2;     // (B)
  float* coef;        // (B) d cost / d o  (= c * f'(o))
  float* mask;        // (B) 1 where this step's iterate becomes the new best
  float* cost;        // (1) batch cost of this step
};

int cw_init(const float* x, float* best_adv, const CwScratch& s, int B, int T, cudaStream_t stream);
// adv = 1/2 (tanh w + 1); cur_l2 = row sums of (adv - x)^2                          (cw.py:73-78,114-115)
int cw_forward_image(const float* x, const CwScratch& s, int B, int T, cudaStream_t stream);
// f, f', cost, best-L2 bookkeeping mask from the logits of adv                       (cw.py:80-101,125-134)
// y_target: nullable target labels of the targeted mode (cw.py:82-83,131-132)
int cw_head(const float* logits, const long long* y, const long long* y_target, const CwScratch& s, float c, float kappa, int B,
            cudaStream_t stream);
// g_w = (2 (adv - x) + g_model) * 1/2 (1 - tanh^2 w); Adam step on w; best_adv <- adv where mask   (cw.py:88-101)
int cw_adam(const float* x, const float* g_model, float* best_adv, const CwScratch& s, float lr, int step, int B, int T,
            cudaStream_t stream);

// per-row L-inf and L2 norms of a - b  (perturbation norms: fab.py:515-521, tests, bench)
int row_diff_norms(const float* a, const float* b, float* linf, float* l2, int B, int T, cudaStream_t stream);

}  // namespace advb
