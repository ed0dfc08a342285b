"""Oracle: LCNN forward from a state_dict (test infrastructure — see oracle/__init__.py).

Follows ``src/models/lcnn.py``: layer stack ``:120-157``, Max-Feature-Map ``:89-95``, BLSTM ``:37-46``,
embedding ``:186-206``, frontend call ``:233-243``.  BatchNorm/Dropout are in eval mode during an attack
(``attack.py:308-321``).  Differentiable through torch autograd.
"""
import torch
import torch.nn.functional as F

from . import frontend as fe

# (conv idx, pool after MFM?, bn idx or None) — src/models/lcnn.py:121-153
BLOCKS = (
    (0, True, None),
    (3, False, 5),
    (6, True, 9),
    (10, False, 12),
    (13, True, None),
    (16, False, 18),
    (19, False, 21),
    (22, False, 24),
    (25, True, None),
)


def mfm(x):
    """lcnn.py:89-95: view (B,2,C/2,H,W), max over dim 1."""
    B, C, H, W = x.shape
    return x.view(B, 2, C // 2, H, W).max(1)[0]


def lstm_dir(x, w_ih, w_hh, b_ih, b_hh, reverse):
    """One direction of nn.LSTM, gate order i,f,g,o, zero initial state.  x (L,B,I) -> (L,B,H)."""
    L, B, _ = x.shape
    H = w_hh.shape[1]
    h = x.new_zeros(B, H)
    c = x.new_zeros(B, H)
    out = [None] * L
    steps = range(L - 1, -1, -1) if reverse else range(L)
    xp = x @ w_ih.t() + b_ih + b_hh
    for t in steps:
        g = xp[t] + h @ w_hh.t()
        i, f, gg, o = g.chunk(4, dim=1)
        c = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(gg)
        h = torch.sigmoid(o) * torch.tanh(c)
        out[t] = h
    return torch.stack(out, 0)


def blstm(x, state, prefix):
    """lcnn.py:37-46: x (B,L,D) -> time-major bidirectional LSTM -> (B,L,D)."""
    xt = x.permute(1, 0, 2)
    p = prefix + ".l_blstm."
    fwd = lstm_dir(xt, state[p + "weight_ih_l0"], state[p + "weight_hh_l0"], state[p + "bias_ih_l0"],
                   state[p + "bias_hh_l0"], False)
    bwd = lstm_dir(xt, state[p + "weight_ih_l0_reverse"], state[p + "weight_hh_l0_reverse"],
                   state[p + "bias_ih_l0_reverse"], state[p + "bias_hh_l0_reverse"], True)
    return torch.cat([fwd, bwd], dim=2).permute(1, 0, 2)


def embedding(feat, state, taps=None):
    """feat (B,1,80,F) cepstral image -> logit (B,1).  lcnn.py:186-206."""
    x = feat.permute(0, 1, 3, 2)  # (B,1,F,80)
    for idx, pool, bn in BLOCKS:
        w = state[f"m_transform.{idx}.weight"]
        x = F.conv2d(x, w, state[f"m_transform.{idx}.bias"], padding=w.shape[-1] // 2)
        if taps is not None:  # winner bookkeeping in the engine's encoding: (MFM half of the winning pixel) << 2 | (dy << 1 | dx)
            Bc, Cc, Hc, Wc = x.shape
            half = x.view(Bc, 2, Cc // 2, Hc, Wc).max(1)[1]
        x = mfm(x)
        if pool:
            if taps is not None:
                _, flat = F.max_pool2d(x, 2, 2, return_indices=True)
                wy, wx = flat // x.shape[3], flat % x.shape[3]
                taps[f"codes{idx}"] = (half.flatten(2).gather(2, flat.flatten(2)).view_as(flat) << 2) | ((wy & 1) << 1) | (wx & 1)
            x = F.max_pool2d(x, 2, 2)
        elif taps is not None:
            taps[f"codes{idx}"] = half << 2
        if bn is not None:
            rm = state[f"m_transform.{bn}.running_mean"].view(1, -1, 1, 1)
            rv = state[f"m_transform.{bn}.running_var"].view(1, -1, 1, 1)
            x = (x - rm) / torch.sqrt(rv + 1e-5)
        if taps is not None:
            taps[f"block{idx}"] = x
    B = x.shape[0]
    hf = x.permute(0, 2, 1, 3).contiguous()
    hf = hf.view(B, hf.shape[1], -1)  # (B, 25, 160), feature = channel*5 + w
    l1 = blstm(hf, state, "m_before_pooling.0")
    l2 = blstm(l1, state, "m_before_pooling.1")
    if taps is not None:
        taps["feats"], taps["lstm1"], taps["lstm2"] = hf, l1, l2
    pooled = (l2 + hf).mean(1)
    return pooled @ state["m_output_act.weight"].t() + state["m_output_act.bias"]


def forward(x, state, taps=None):
    """waveform (B,T) -> logit (B,1).  lcnn.py:233-243."""
    fb, dct, window, _ = fe.tables_from_state(state)
    feat = fe.cepstral_frontend(x, fb, dct, window).unsqueeze(1)
    if taps is not None:
        taps["frontend"] = feat
    return embedding(feat, state, taps)
