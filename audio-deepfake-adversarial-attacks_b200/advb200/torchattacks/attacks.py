"""FGSM / PGD / PGDL2 on the native engine (fgsm.py:7-62, pgd.py:7-78, pgdl2.py:7-90)."""
import torch

from .. import _lib
from .attack import Attack


def _desc(kind, **kw):
    d = _lib.AttackDesc()
    d.kind = kind
    for k, v in kw.items():
        setattr(d, k, v)
    return d


class FGSM(Attack):
    def __init__(self, model, eps=0.007):
        super().__init__("FGSM", model)
        self.eps = eps
        self._supported_mode = ["default", "targeted"]

    def forward(self, images, labels):
        images, labels = self._prepare(images, labels)
        target = self._get_target_label(images, labels) if self._targeted else None
        return self._engine(images).attack(_desc(_lib.ATTACK_FGSM, eps=self.eps), images, labels, minmax=self._fused_minmax,
                                           target=target)


class PGD(Attack):
    def __init__(self, model, eps=0.3, alpha=2 / 255, steps=40, random_start=True):
        super().__init__("PGD", model)
        self.eps = eps
        self.alpha = alpha
        self.steps = steps
        self.random_start = random_start
        self._supported_mode = ["default", "targeted"]

    def forward(self, images, labels, noise=None):
        images, labels = self._prepare(images, labels)
        target = self._get_target_label(images, labels) if self._targeted else None
        if self.random_start and noise is None:
            # same draw, on the same device RNG stream, as pgd.py:56
            noise = torch.empty_like(images).uniform_(-self.eps, self.eps)
        if noise is not None:
            noise = noise.to(images.device)
        d = _desc(_lib.ATTACK_PGD, eps=self.eps, alpha=self.alpha, steps=self.steps)
        return self._engine(images).attack(d, images, labels, noise, minmax=self._fused_minmax, target=target)


class PGDL2(Attack):
    def __init__(self, model, eps=1.0, alpha=0.2, steps=40, random_start=True, eps_for_division=1e-10):
        super().__init__("PGDL2", model)
        self.eps = eps
        self.alpha = alpha
        self.steps = steps
        self.random_start = random_start
        self.eps_for_division = eps_for_division
        self._supported_mode = ["default", "targeted"]

    def forward(self, images, labels, delta=None):
        images, labels = self._prepare(images, labels)
        target = self._get_target_label(images, labels) if self._targeted else None
        if self.random_start and delta is None:
            # pgdl2.py:57-61: same RNG draws in the same order (normal_ then uniform_)
            delta = torch.empty_like(images).normal_()
            n = delta.view(images.size(0), -1).norm(p=2, dim=1).view(images.size(0), 1)
            r = torch.zeros_like(n).uniform_(0, 1)
            delta *= r / n * self.eps
        if delta is not None:
            delta = delta.to(images.device)
        d = _desc(_lib.ATTACK_PGDL2, eps=self.eps, alpha=self.alpha, steps=self.steps, eps_div=self.eps_for_division)
        return self._engine(images).attack(d, images, labels, delta, minmax=self._fused_minmax, target=target)


class FAB(Attack):
    """fab.py:11-78, :495-526 (perturb), :131-307 (attack_single_run): ``norm='Linf'`` (what the repo's ``AttackEnum``
    reaches) and ``norm='L2'``, untargeted, any ``n_restarts``.  ``norm='L1'`` cannot run in the reference either
    (``FAB.perturb`` never assigns ``res`` for it, fab.py:515-521: NameError) and raises here.  The index selection of the
    reference (attack only the clips that are still correctly classified) stays here as tensor plumbing; every forward,
    gradient, projection and norm runs in libadvb200."""

    def __init__(self, model, norm="Linf", eps=None, steps=100, n_restarts=1, alpha_max=0.1, eta=1.05, beta=0.9,
                 verbose=False, seed=0, targeted=False, n_classes=10):
        super().__init__("FAB", model)
        if norm == "L1":
            raise NotImplementedError("FAB norm='L1' does not run in the reference (fab.py:515-521 leaves `res` undefined); "
                                      "advb200 provides 'Linf' and 'L2'")
        if norm not in ("Linf", "L2"):
            raise ValueError("norm not supported")  # fab.py:224
        # (the reference's patched copy ignores `targeted` too: fab.py:63 sets self.targeted = False unconditionally)
        self.norm = norm
        self.n_restarts = n_restarts
        self.eps = eps if eps is not None else {"Linf": 0.3, "L2": 1.0, "L1": 5.0}[norm]
        self.alpha_max = alpha_max
        self.eta = eta
        self.beta = beta
        self.steps = steps
        self.targeted = False
        self.verbose = verbose
        self.seed = seed
        self.target_class = None
        self.n_target_classes = n_classes - 1
        self._supported_mode = ["default"]

    def forward(self, images, labels):
        images, labels = self._prepare(images, labels)
        return self.perturb(images, labels)

    def _get_predicted_label(self, x):
        # argmax of cat([-o, o]) (fab.py:80-85): class 1 iff o > 0 (a tie goes to class 0)
        return (self._engine(x).forward(x)[:, 0] > 0).long()

    def attack_single_run(self, x, y=None, use_rand_start=False):
        y_pred = self._get_predicted_label(x)
        y = y_pred.clone() if y is None else y.clone().long().to(x.device)
        pred = y_pred == y
        if pred.sum() == 0:
            return x
        idx = pred.nonzero().flatten()
        im2, la2 = x[idx].contiguous(), y[idx].contiguous()
        start = None
        if use_rand_start:
            # fab.py:176-194: res2 is still 1e10 at this point, so the radius is eps; same RNG draw (torch.rand / torch.randn on
            # the CPU generator, then moved - exactly what fab.py:178,185 do), same op order
            radius = torch.full((im2.shape[0], 1), float(self.eps), device=im2.device)
            if self.norm == "Linf":
                t = 2 * torch.rand(im2.shape).to(im2.device) - 1
                start = im2 + radius * t / t.abs().max(dim=1, keepdim=True)[0] * .5
            else:
                t = torch.randn(im2.shape).to(im2.device)
                start = im2 + radius * t / (t ** 2).sum(dim=-1, keepdim=True).sqrt() * .5
            start = start.clamp(0.0, 1.0)
        d = _desc(_lib.ATTACK_FAB, eps=self.eps, steps=self.steps, alpha_max=self.alpha_max, eta=self.eta, beta=self.beta,
                  norm=_lib.NORM_L2 if self.norm == "L2" else _lib.NORM_LINF)
        adv = self._engine(x).attack(d, im2, la2, start)  # rows never found adversarial come back equal to im2
        adv_c = x.clone()
        adv_c[idx] = adv
        return adv_c

    def perturb(self, x, y):
        adv = x.clone()
        acc = self._get_predicted_label(x) == y
        # fab.py:504-505: the reference reseeds the global RNGs on every call (kept: it affects later shuffles)
        torch.random.manual_seed(self.seed)
        torch.cuda.random.manual_seed(self.seed)
        for counter in range(self.n_restarts):  # fab.py:507-526
            ind_to_fool = acc.nonzero().flatten()
            if ind_to_fool.numel() != 0:
                x_to_fool, y_to_fool = x[ind_to_fool].contiguous(), y[ind_to_fool].contiguous()
                adv_curr = self.attack_single_run(x_to_fool, y_to_fool, use_rand_start=(counter > 0))
                eng = self._engine(x)
                acc_curr = self._get_predicted_label(adv_curr) == y_to_fool
                linf, l2 = eng.row_diff_norms(x_to_fool, adv_curr)
                res = l2 if self.norm == "L2" else linf  # fab.py:515-519
                acc_curr = torch.max(acc_curr, res > self.eps)
                ind_curr = (acc_curr == 0).nonzero().flatten()
                acc[ind_to_fool[ind_curr]] = 0
                adv[ind_to_fool[ind_curr]] = adv_curr[ind_curr].clone()
        return adv


class CW(Attack):
    """cw.py:10-134: tanh-space Adam on  sum ||x' - x||^2 + c sum f(x')  with the batch-wide early stop."""

    def __init__(self, model, c=1e-4, kappa=0, steps=1000, lr=0.01):
        super().__init__("CW", model)
        self.c = c
        self.kappa = kappa
        self.steps = steps
        self.lr = lr
        self._supported_mode = ["default", "targeted"]

    def forward(self, images, labels):
        images, labels = self._prepare(images, labels)
        target = self._get_target_label(images, labels) if self._targeted else None
        d = _desc(_lib.ATTACK_CW, c=self.c, kappa=self.kappa, steps=self.steps, lr=self.lr)
        return self._engine(images).attack(d, images, labels, target=target)
