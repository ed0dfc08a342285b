"""Oracle: RawNet3 forward from a state_dict (test infrastructure — see oracle/__init__.py).

Follows ``src/models/rawnet3.py``: forward ``:73-137``, PreEmphasis ``:151-158``, AFMS ``:176-182``, Bottle2neck
``:242-274`` (conv -> ReLU -> BN order), ``prepare_model`` ``:277-291``; the sinc filters restate asteroid-filterbanks
0.4.0 ``ParamSincFB.filters`` (third party, source absent: PARITY UNPINNED for the filter formula).
BatchNorm in eval mode; InstanceNorm always uses the clip's own statistics.  Differentiable through
torch autograd.
"""
import torch
import torch.nn.functional as F

LAYERS = (("layer1", 2, 5), ("layer2", 3, 3), ("layer3", 4, 0))  # name, dilation, pool


def sinc_filters(state, prefix="conv1.filterbank."):
    """(256, 1, 251) band-pass filters from low_hz_/band_hz_ (min_low_hz = min_band_hz = 50, sample rate 16 kHz)."""
    low_p, band_p = state[prefix + "low_hz_"], state[prefix + "band_hz_"]
    window, n_ = state[prefix + "window_"], state[prefix + "n_"]
    low = 50 + torch.abs(low_p)
    high = torch.clamp(low + 50 + torch.abs(band_p), 50, 8000.0)
    band = (high - low)[:, 0]
    ft_low, ft_high = torch.matmul(low, n_), torch.matmul(high, n_)
    out = []
    for kind in ("cos", "sin"):
        if kind == "cos":
            left = ((torch.sin(ft_high) - torch.sin(ft_low)) / (n_ / 2)) * window
            centre, right = 2 * band.view(-1, 1), torch.flip(left, dims=[1])
        else:
            left = ((torch.cos(ft_low) - torch.cos(ft_high)) / (n_ / 2)) * window
            centre, right = torch.zeros_like(band.view(-1, 1)), -torch.flip(left, dims=[1])
        bp = torch.cat([left, centre, right], dim=1) / (2 * band[:, None])
        out.append(bp.view(-1, 1, bp.shape[1]))
    return torch.cat(out, dim=0)


def bn(x, state, prefix, eps=1e-5):
    """nn.BatchNorm1d (affine) in eval mode on (B,C,T) or (B,C)."""
    shape = (1, -1, 1) if x.dim() == 3 else (1, -1)
    rm, rv = state[prefix + ".running_mean"].view(shape), state[prefix + ".running_var"].view(shape)
    return (x - rm) / torch.sqrt(rv + eps) * state[prefix + ".weight"].view(shape) + state[prefix + ".bias"].view(shape)


def bottle2neck(x, state, p, dilation, pool, taps=None):
    """rawnet3.py:242-274."""
    rkey = p + ".residual.0.weight"
    residual = F.conv1d(x, state[rkey]) if rkey in state else x
    out = bn(F.relu(F.conv1d(x, state[p + ".conv1.weight"], state[p + ".conv1.bias"])), state, p + ".bn1")
    spx = torch.split(out, 128, 1)
    outs = []
    sp = None
    for i in range(7):
        sp = spx[i] if i == 0 else sp + spx[i]
        sp = F.conv1d(sp, state[f"{p}.convs.{i}.weight"], state[f"{p}.convs.{i}.bias"], dilation=dilation,
                      padding=dilation)
        sp = bn(F.relu(sp), state, f"{p}.bns.{i}")
        outs.append(sp)
    outs.append(spx[7])
    out = torch.cat(outs, 1)
    out = bn(F.relu(F.conv1d(out, state[p + ".conv3.weight"], state[p + ".conv3.bias"])), state, p + ".bn3")
    out = out + residual
    if taps is not None:
        taps[p + ".pre_pool"] = out
    if pool:
        out = F.max_pool1d(out, pool)
    y = torch.sigmoid(out.mean(dim=2) @ state[p + ".afms.fc.weight"].t() + state[p + ".afms.fc.bias"])
    return (out + state[p + ".afms.alpha"]) * y.unsqueeze(2)


def preprocess(x, state):
    """PreEmphasis (rawnet3.py:151-158) + InstanceNorm1d(1, eps=1e-4, affine) (:24-26): (B,T) -> (B,1,T)."""
    v = x.unsqueeze(1)
    v = F.conv1d(F.pad(v, (1, 0), "reflect"), state["preprocess.0.flipped_filter"])
    return F.instance_norm(v, weight=state["preprocess.1.weight"], bias=state["preprocess.1.bias"], eps=1e-4)


def tail(s, state, taps=None):
    """Everything after the sinc convolution: raw filter outputs (B,256,L) -> logit (B,1).  rawnet3.py:80-137."""
    v = torch.log(torch.abs(s) + 1e-6)
    v = v - v.mean(dim=-1, keepdim=True)
    if taps is not None:
        taps["sinc"] = v
    x1 = bottle2neck(v, state, "layer1", 2, 5, taps)
    x2 = bottle2neck(x1, state, "layer2", 3, 3, taps)
    m1 = F.max_pool1d(x1, 3)
    x3 = bottle2neck(m1 + x2, state, "layer3", 4, 0, taps)
    if taps is not None:
        taps["x1"], taps["x2"], taps["x3"] = x1, x2, x3
    h = F.relu(F.conv1d(torch.cat((m1, x2, x3), dim=1), state["layer4.weight"], state["layer4.bias"]))
    if taps is not None:
        taps["layer4"] = h
    t = h.shape[-1]
    g = torch.cat((h, h.mean(dim=2, keepdim=True).repeat(1, 1, t),
                   torch.sqrt(h.var(dim=2, keepdim=True).clamp(min=1e-4, max=1e4)).repeat(1, 1, t)), dim=1)
    a = bn(F.relu(F.conv1d(g, state["attention.0.weight"], state["attention.0.bias"])), state, "attention.2")
    w = torch.softmax(F.conv1d(a, state["attention.3.weight"], state["attention.3.bias"]), dim=2)
    mu = torch.sum(h * w, dim=2)
    sg = torch.sqrt((torch.sum((h ** 2) * w, dim=2) - mu ** 2).clamp(min=1e-4, max=1e4))
    pooled = torch.cat((mu, sg), 1)
    if taps is not None:
        taps["pooled"] = pooled
    return bn(pooled, state, "bn5") @ state["fc6.weight"].t() + state["fc6.bias"]


def forward(x, state, taps=None):
    """waveform (B,T) -> logit (B,1).  rawnet3.py:73-137."""
    v = preprocess(x, state)
    if taps is not None:
        taps["pre"] = v
    s = F.conv1d(v, sinc_filters(state), stride=10)
    if taps is not None:
        taps["sinc_raw"] = s
    return tail(s, state, taps)
