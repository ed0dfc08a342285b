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
        self._supported_mode = ["default"]

    def forward(self, images, labels):
        images, labels = self._prepare(images, labels)
        return self._engine(images).attack(_desc(_lib.ATTACK_FGSM, eps=self.eps), images, labels)


class PGD(Attack):
    def __init__(self, model, eps=0.3, alpha=2 / 255, steps=40, random_start=True):
        super().__init__("PGD", model)
        self.eps = eps
        self.alpha = alpha
        self.steps = steps
        self.random_start = random_start
        self._supported_mode = ["default"]

    def forward(self, images, labels, noise=None):
        images, labels = self._prepare(images, labels)
        if self.random_start and noise is None:
            # same draw, on the same device RNG stream, as pgd.py:56
            noise = torch.empty_like(images).uniform_(-self.eps, self.eps)
        if noise is not None:
            noise = noise.to(images.device)
        d = _desc(_lib.ATTACK_PGD, eps=self.eps, alpha=self.alpha, steps=self.steps)
        return self._engine(images).attack(d, images, labels, noise)


class PGDL2(Attack):
    def __init__(self, model, eps=1.0, alpha=0.2, steps=40, random_start=True, eps_for_division=1e-10):
        super().__init__("PGDL2", model)
        self.eps = eps
        self.alpha = alpha
        self.steps = steps
        self.random_start = random_start
        self.eps_for_division = eps_for_division
        self._supported_mode = ["default"]

    def forward(self, images, labels, delta=None):
        images, labels = self._prepare(images, labels)
        if self.random_start and delta is None:
            # pgdl2.py:57-61: same RNG draws in the same order (normal_ then uniform_)
            delta = torch.empty_like(images).normal_()
            n = delta.view(images.size(0), -1).norm(p=2, dim=1).view(images.size(0), 1)
            r = torch.zeros_like(n).uniform_(0, 1)
            delta *= r / n * self.eps
        if delta is not None:
            delta = delta.to(images.device)
        d = _desc(_lib.ATTACK_PGDL2, eps=self.eps, alpha=self.alpha, steps=self.steps, eps_div=self.eps_for_division)
        return self._engine(images).attack(d, images, labels, delta)
