"""Oracle: the patched torchattacks update rules (test infrastructure — see oracle/__init__.py).

``model_fn(x) -> logit (B,1)``.  Every attack builds 2-class logits ``cat([-o, o])`` and uses mean
CrossEntropy with int64 labels (``fgsm.py:43-53``, ``pgd.py:61-68``, ``pgdl2.py:66-73``).
Random starts are passed in (drawn by the caller with torch) so both sides see identical noise (F9).
"""
import torch
import torch.nn.functional as F


def loss_and_grad(model_fn, x, y):
    """returns (logit (B,1) detached, d CE / d x)."""
    x = x.clone().detach().requires_grad_(True)
    o = model_fn(x)
    z = torch.cat([-o, o], dim=1)
    cost = F.cross_entropy(z, y)
    (g,) = torch.autograd.grad(cost, x)
    return o.detach(), g


def fgsm(model_fn, x, y, eps):
    """fgsm.py:33-62."""
    _, g = loss_and_grad(model_fn, x, y)
    return torch.clamp(x + eps * g.sign(), 0, 1).detach()


def pgd(model_fn, x, y, eps, alpha=2 / 255, steps=40, noise=None, trace=None):
    """pgd.py:40-78.  ``noise`` = the U(-eps,eps) tensor the reference draws at :56 (None = no random start)."""
    adv = x.clone().detach()
    if noise is not None:
        adv = torch.clamp(adv + noise, 0, 1).detach()
    for _ in range(steps):
        o, g = loss_and_grad(model_fn, adv, y)
        if trace is not None:
            trace.append((o, g))
        adv = adv.detach() + alpha * g.sign()
        delta = torch.clamp(adv - x, min=-eps, max=eps)
        adv = torch.clamp(x + delta, 0, 1).detach()
    return adv


def pgdl2_start(x, eps, normal, r):
    """pgdl2.py:55-62 given the N(0,1) tensor ``normal`` (B,T) and the U(0,1) column ``r`` (B,1)."""
    n = normal.view(x.shape[0], -1).norm(p=2, dim=1).view(-1, 1)
    return torch.clamp(x + normal * (r / n * eps), 0, 1).detach()


def pgdl2(model_fn, x, y, eps, alpha=0.2, steps=40, start=None, eps_div=1e-10):
    """pgdl2.py:40-90.  ``start`` = the already-built random start (see pgdl2_start) or None."""
    B = x.shape[0]
    adv = x.clone().detach() if start is None else start.clone().detach()
    for _ in range(steps):
        _, g = loss_and_grad(model_fn, adv, y)
        gn = torch.norm(g.view(B, -1), p=2, dim=1) + eps_div
        g = g / gn.view(B, 1)
        adv = adv.detach() + alpha * g
        delta = adv - x
        dn = torch.norm(delta.view(B, -1), p=2, dim=1)
        factor = torch.min(eps / dn, torch.ones_like(dn))
        delta = delta * factor.view(-1, 1)
        adv = torch.clamp(x + delta, 0, 1).detach()
    return adv


def to_minmax(x):
    """src/aa/utils.py:4-9."""
    mn = x.min(dim=1, keepdim=True)[0]
    mx = x.max(dim=1, keepdim=True)[0]
    return (x - mn) / (mx - mn), mn, mx


def revert_minmax(x, mn, mx):
    """src/aa/utils.py:12-14."""
    return x * (mx - mn) + mn
