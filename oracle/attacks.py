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


# ------------------------------------------------------------------------------------------------------------
# FAB (untargeted, L-inf, 2 classes) — fab.py:70-78 (forward), :495-526 (perturb), :131-307 (attack_single_run),
# :90-112 (get_diff_logits_grads_batch), :562-614 (projection_linf).  SURVEY.md App. A.6.
# ------------------------------------------------------------------------------------------------------------
def logit_and_grad(model_fn, x):
    """(o (B,1) detached, d o_i / d x_i).  The reference runs two backwards (classes 0 and 1 of z = [-o, o],
    fab.py:96-103); the class-0 gradient is the exact negation of the class-1 gradient."""
    x = x.clone().detach().requires_grad_(True)
    o = model_fn(x)
    (g,) = torch.autograd.grad(o.sum(), x)
    return o.detach(), g


def predicted_label(model_fn, x):
    """fab.py:80-85: argmax of cat([-o, o]) (ties -> class 0, like torch.max)."""
    with torch.no_grad():
        o = model_fn(x)
    return torch.cat([-o, o], dim=1).max(dim=1)[1]


def projection_linf(t, w, b):
    """fab.py:562-614, statement by statement (rows independent)."""
    import math

    w, b = w.clone(), b.clone()
    sign = 2 * ((w * t).sum(1) - b >= 0) - 1
    w.mul_(sign.unsqueeze(1))
    b.mul_(sign)
    a = (w < 0).float()
    d = (a - t) * (w != 0).float()
    p = a - t * (2 * a - 1)
    indp = torch.argsort(p, dim=1)
    b = b - (w * t).sum(1)
    b0 = (w * d).sum(1)
    indp2 = indp.flip((1,))
    ws = w.gather(1, indp2)
    bs2 = -ws * d.gather(1, indp2)
    s = torch.cumsum(ws.abs(), dim=1)
    sb = torch.cumsum(bs2, dim=1) + b0.unsqueeze(1)
    b2 = sb[:, -1] - s[:, -1] * p.gather(1, indp[:, 0:1]).squeeze(1)
    c_l = b - b2 > 0
    c2 = (b - b0 > 0) & (~c_l)
    lb = torch.zeros(int(c2.sum()))
    ub = torch.full_like(lb, w.shape[1] - 1)
    nitermax = math.ceil(math.log2(w.shape[1]))
    indp_, sb_, s_, p_, b_ = indp[c2], sb[c2], s[c2], p[c2], b[c2]
    for _ in range(nitermax):
        c4 = torch.floor((lb + ub) / 2)
        c2i = c4.long().unsqueeze(1)
        indcurr = indp_.gather(1, indp_.size(1) - 1 - c2i)
        b2 = (sb_.gather(1, c2i) - s_.gather(1, c2i) * p_.gather(1, indcurr)).squeeze(1)
        c = b_ - b2 > 0
        lb = torch.where(c, c4, lb)
        ub = torch.where(c, ub, c4)
    lb = lb.long()
    if c_l.any():
        lm = torch.clamp_min((b[c_l] - sb[c_l, -1]) / (-s[c_l, -1]), min=0).unsqueeze(-1)
        d[c_l] = (2 * a[c_l] - 1) * lm
    lm = torch.clamp_min((b[c2] - sb[c2, lb]) / (-s[c2, lb]), min=0).unsqueeze(-1)
    d[c2] = torch.min(lm, d[c2]) * a[c2] + torch.max(-lm, d[c2]) * (1 - a[c2])
    return d * (w != 0).float()


def projection_l2(t, w, b):
    """fab.py:617-665, statement by statement (rows independent)."""
    import math

    import torch.nn.functional as F

    w = w.clone()
    c = (w * t).sum(1) - b
    ind2 = 2 * (c >= 0) - 1
    w.mul_(ind2.unsqueeze(1))
    c = c * ind2
    r = torch.max(t / w, (t - 1) / w).clamp(min=-1e12, max=1e12)
    r.masked_fill_(w.abs() < 1e-8, 1e12)
    r[r == -1e12] *= -1
    rs, indr = torch.sort(r, dim=1)
    rs2 = F.pad(rs[:, 1:], (0, 1))
    rs.masked_fill_(rs == 1e12, 0)
    rs2.masked_fill_(rs2 == 1e12, 0)
    w3s = (w ** 2).gather(1, indr)
    w5 = w3s.sum(dim=1, keepdim=True)
    ws = w5 - torch.cumsum(w3s, dim=1)
    d = -(r * w)
    d.mul_((w.abs() > 1e-8).float())
    s = torch.cat((-w5 * rs[:, 0:1], torch.cumsum((-rs2 + rs) * ws, dim=1) - w5 * rs[:, 0:1]), 1)
    c4 = s[:, 0] + c < 0
    c3 = (d * w).sum(dim=1) + c > 0
    c2 = ~(c4 | c3)
    lb = torch.zeros(int(c2.sum()))
    ub = torch.full_like(lb, w.shape[1] - 1)
    nitermax = math.ceil(math.log2(w.shape[1]))
    s_, c_ = s[c2], c[c2]
    for _ in range(nitermax):
        counter4 = torch.floor((lb + ub) / 2)
        counter2 = counter4.long().unsqueeze(1)
        c3b = s_.gather(1, counter2).squeeze(1) + c_ > 0
        lb = torch.where(c3b, counter4, lb)
        ub = torch.where(c3b, ub, counter4)
    lb = lb.long()
    if c4.any():
        alpha = c[c4] / w5[c4].squeeze(-1)
        d[c4] = -alpha.unsqueeze(-1) * w[c4]
    if c2.any():
        alpha = (s[c2, lb] + c[c2]) / ws[c2, lb] + rs[c2, lb]
        alpha[ws[c2, lb] == 0] = 0
        c5 = (alpha.unsqueeze(-1) > r[c2]).float()
        d[c2] = d[c2] * c5 - alpha.unsqueeze(-1) * w[c2] * (1 - c5)
    return d * (w.abs() > 1e-8).float()


def _row_norm(v, norm):
    """The per-row norm FAB uses for its distances (fab.py:248-256,273-281)."""
    v = v.reshape(v.shape[0], -1)
    return v.abs().max(dim=1)[0] if norm == "Linf" else (v ** 2).sum(dim=-1).sqrt()


def fab_single_run(model_fn, x, y, steps=100, alpha_max=0.1, eta=1.05, beta=0.9, trace=None, norm="Linf", start=None):
    """fab.py:131-307 with norm='Linf' / 'L2'; ``start`` = the random restart point of :176-194 (None: no random start)."""
    x = x.detach().clone()
    pred = predicted_label(model_fn, x) == y
    if pred.sum() == 0:
        return x
    pred = pred.nonzero().flatten()
    im2, la2 = x[pred].clone(), y[pred].clone()
    bs = im2.shape[0]
    u1 = torch.arange(bs)
    adv, adv_c = im2.clone(), x.clone()
    res2 = 1e10 * torch.ones(bs)
    x1, x0 = im2.clone(), im2.clone().reshape(bs, -1)
    if start is not None:
        x1 = start[pred].clone()
    proj = projection_linf if norm == "Linf" else projection_l2
    for _ in range(steps):
        o, g = logit_and_grad(model_fn, x1)
        z = torch.cat([-o, o], dim=1)
        g2 = torch.stack([-g, g], dim=1)  # (bs, 2, T)
        df = z - z[u1, la2].unsqueeze(1)
        dg = g2 - g2[u1, la2].unsqueeze(1)
        df[u1, la2] = 1e10
        if norm == "Linf":
            dist1 = df.abs() / (1e-12 + dg.abs().view(bs, 2, -1).sum(dim=-1))
        else:
            dist1 = df.abs() / (1e-12 + (dg ** 2).view(bs, 2, -1).sum(dim=-1).sqrt())
        ind = dist1.min(dim=1)[1]
        dg2 = dg[u1, ind]
        b = -df[u1, ind] + (dg2 * x1).view(bs, -1).sum(dim=-1)
        w = dg2.reshape(bs, -1)
        d3 = proj(torch.cat((x1.reshape(bs, -1), x0), 0), torch.cat((w, w), 0), torch.cat((b, b), 0))
        d1, d2 = d3[:bs].reshape(x1.shape), d3[-bs:].reshape(x1.shape)
        a0 = _row_norm(d3, norm).unsqueeze(1)
        a0 = torch.max(a0, 1e-8 * torch.ones_like(a0))
        a1, a2 = a0[:bs], a0[-bs:]
        alpha = torch.min(torch.max(a1 / (a1 + a2), torch.zeros_like(a1)), alpha_max * torch.ones_like(a1))
        x1 = ((x1 + eta * d1) * (1 - alpha) + (im2 + d2 * eta) * alpha).clamp(0.0, 1.0)
        is_adv = predicted_label(model_fn, x1) != la2
        if trace is not None:
            trace.append(dict(o=o.clone(), linf_d3=a0.clone(), x1=x1.clone(), is_adv=is_adv.clone()))
        if is_adv.sum() > 0:
            ia = is_adv.nonzero().flatten()
            t = _row_norm(x1[ia] - im2[ia], norm)
            better = (t < res2[ia]).float().unsqueeze(1)
            adv[ia] = x1[ia] * better + adv[ia] * (1 - better)
            res2[ia] = t * (t < res2[ia]).float() + res2[ia] * (t >= res2[ia]).float()
            x1[ia] = im2[ia] + (x1[ia] - im2[ia]) * beta
    succ = (res2 < 1e10).nonzero().flatten()
    adv_c[pred[succ]] = adv[succ].clone()
    return adv_c


def fab(model_fn, x, y, eps=0.3, steps=100, alpha_max=0.1, eta=1.05, beta=0.9, norm="Linf"):
    """fab.py:495-526 (n_restarts=1, untargeted, Linf / L2).  The reference reseeds torch's global RNG here (:504-505);
    with one restart no random number is consumed."""
    adv = x.clone()
    acc = predicted_label(model_fn, x) == y
    idx = acc.nonzero().flatten()
    if idx.numel() == 0:
        return adv
    xf, yf = x[idx].clone(), y[idx].clone()
    cur = fab_single_run(model_fn, xf, yf, steps, alpha_max, eta, beta, norm=norm)
    still = predicted_label(model_fn, cur) == yf
    res = _row_norm(xf - cur, norm)
    still = torch.max(still, res > eps)
    ok = (still == 0).nonzero().flatten()
    adv[idx[ok]] = cur[ok].clone()
    return adv


# ------------------------------------------------------------------------------------------------------------
# CW — cw.py:46-134.  SURVEY.md App. A.7.
# ------------------------------------------------------------------------------------------------------------
def cw(model_fn, x, y, c=1e-4, kappa=0.0, steps=1000, lr=0.01, target=None):
    """``target``: target labels of the targeted mode (cw.py:56-57,82-83,131-132); the best-adversarial bookkeeping keeps
    comparing with ``y`` (cw.py:95)."""
    x = x.clone().detach()
    u = x * 2 - 1
    w = (0.5 * torch.log((1 + u) / (1 - u))).detach()  # cw.py:117-123 (+-inf where x is exactly 0 or 1)
    w.requires_grad = True
    best_adv = x.clone()
    best_l2 = 1e10 * torch.ones(len(x))
    prev_cost = 1e10
    opt = torch.optim.Adam([w], lr=lr)
    onehot = torch.eye(2)[y if target is None else target]
    for step in range(steps):
        adv = 0.5 * (torch.tanh(w) + 1)
        cur_l2 = ((adv.flatten(1) - x.flatten(1)) ** 2).sum(dim=1)
        o = model_fn(adv)
        z = torch.cat([-o, o], dim=1)
        i, _ = torch.max((1 - onehot) * z, dim=1)
        j = torch.masked_select(z, onehot.bool())
        f_loss = torch.clamp((j - i) if target is None else (i - j), min=-kappa).sum()
        cost = cur_l2.sum() + c * f_loss
        opt.zero_grad()
        cost.backward()
        opt.step()
        pre = torch.max(z.detach(), 1)[1]
        correct = (pre == y).float()
        mask = (1 - correct) * (best_l2 > cur_l2.detach())
        best_l2 = mask * cur_l2.detach() + (1 - mask) * best_l2
        best_adv = mask.view(-1, 1) * adv.detach() + (1 - mask.view(-1, 1)) * best_adv
        if step % max(steps // 10, 1) == 0:
            if cost.item() > prev_cost:
                return best_adv
            prev_cost = cost.item()
    return best_adv
