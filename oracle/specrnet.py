"""Oracle: SpecRNet forward from a state_dict (test infrastructure — see oracle/__init__.py).

Follows ``src/models/specrnet.py``: residual block ``:73-91`` (its ``bn1``/``lrelu`` output is DISCARDED — ``conv1`` is
applied to the raw block input, ``:76-81``; reproduced here), embedding ``:139-181``, frontend call ``:203-214``.
BatchNorm in eval mode.  Differentiable through torch autograd.
"""
import torch
import torch.nn.functional as F

from . import frontend as fe


def bn(x, state, prefix, eps=1e-5):
    """nn.BatchNorm2d (affine) in eval mode."""
    shape = (1, -1, 1, 1)
    rm, rv = state[prefix + ".running_mean"].view(shape), state[prefix + ".running_var"].view(shape)
    w, b = state[prefix + ".weight"].view(shape), state[prefix + ".bias"].view(shape)
    return (x - rm) / torch.sqrt(rv + eps) * w + b


def residual_block(x, state, prefix, downsample):
    """specrnet.py:73-91."""
    out = F.conv2d(x, state[prefix + ".conv1.weight"], state[prefix + ".conv1.bias"], padding=1)
    out = F.leaky_relu(bn(out, state, prefix + ".bn2"), 0.3)
    out = F.conv2d(out, state[prefix + ".conv2.weight"], state[prefix + ".conv2.bias"], padding=1)
    identity = x
    if downsample:
        identity = F.conv2d(x, state[prefix + ".conv_downsample.weight"], state[prefix + ".conv_downsample.bias"])
    return F.max_pool2d(out + identity, 2)


def gru_dir(x, w_ih, w_hh, b_ih, b_hh, reverse):
    """One direction of nn.GRU (gate order r, z, n), zero initial state.  x (B,L,I) -> (B,L,H)."""
    B, L, _ = x.shape
    H = w_hh.shape[1]
    h = x.new_zeros(B, H)
    out = [None] * L
    gi_all = x @ w_ih.t() + b_ih
    for t in (range(L - 1, -1, -1) if reverse else range(L)):
        gi, gh = gi_all[:, t], h @ w_hh.t() + b_hh
        i_r, i_z, i_n = gi.chunk(3, dim=1)
        h_r, h_z, h_n = gh.chunk(3, dim=1)
        r = torch.sigmoid(i_r + h_r)
        z = torch.sigmoid(i_z + h_z)
        n = torch.tanh(i_n + r * h_n)
        h = (1 - z) * n + z * h
        out[t] = h
    return torch.stack(out, 1)


def gru(x, state, layers=2):
    for l in range(layers):
        fwd = gru_dir(x, state[f"gru.weight_ih_l{l}"], state[f"gru.weight_hh_l{l}"], state[f"gru.bias_ih_l{l}"],
                      state[f"gru.bias_hh_l{l}"], False)
        bwd = gru_dir(x, state[f"gru.weight_ih_l{l}_reverse"], state[f"gru.weight_hh_l{l}_reverse"],
                      state[f"gru.bias_ih_l{l}_reverse"], state[f"gru.bias_hh_l{l}_reverse"], True)
        x = torch.cat([fwd, bwd], dim=2)
    return x


def embedding(feat, state, taps=None):
    """feat (B,1,80,F) -> logit (B,1).  specrnet.py:139-181."""
    x = F.selu(bn(feat, state, "first_bn"))
    for name, down in (("0", True), ("2", True), ("4", False)):
        xb = residual_block(x, state, f"block{name}.0", down)
        y = xb.mean(dim=(2, 3))
        y = torch.sigmoid(y @ state[f"fc_attention{name}.0.weight"].t() + state[f"fc_attention{name}.0.bias"])
        y = y.view(y.shape[0], -1, 1, 1)
        x = F.max_pool2d(xb * y + y, 2)
        if taps is not None:
            taps[f"block{name}"] = xb
            taps[f"stage{name}"] = x
    x = F.selu(bn(x, state, "bn_before_gru"))
    x = x.squeeze(-2).permute(0, 2, 1)  # (B, L, 64)
    if taps is not None:
        taps["gru_in"] = x
    x = gru(x, state)[:, -1, :]
    if taps is not None:
        taps["gru_last"] = x
    x = x @ state["fc1_gru.weight"].t() + state["fc1_gru.bias"]
    return x @ state["fc2_gru.weight"].t() + state["fc2_gru.bias"]


def forward(x, state, taps=None):
    """waveform (B,T) -> logit (B,1).  specrnet.py:203-214."""
    fb, dct, window, _ = fe.tables_from_state(state)
    feat = fe.cepstral_frontend(x, fb, dct, window).unsqueeze(1)
    if taps is not None:
        taps["frontend"] = feat
    return embedding(feat, state, taps)
