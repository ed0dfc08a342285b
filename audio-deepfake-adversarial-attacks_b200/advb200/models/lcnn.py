"""LCNN parameter holder.

Same constructor, parameter names and shapes as the reference's ``src/models/lcnn.py:102-243``
(``LCNN(device=..., input_channels=1, frontend_algorithm=[...])``), so reference checkpoints load with
``load_state_dict`` and ``nn.DataParallel``/``.to(device)`` behave the same.  The layer stack is kept
as an index-compatible ``nn.Sequential`` (``m_transform.0`` … ``m_transform.25``) whose parametric
slots are real ``nn.Conv2d`` / ``nn.BatchNorm2d(affine=False)`` modules used only as storage: no
PyTorch op of theirs ever runs.  ``forward`` hands the waveform to the CUDA engine
(``csrc/engine.cu``), which reads the live parameter storage on every call.
"""
import torch
from torch import nn

from .. import frontends


class _Slot(nn.Module):
    """Non-parametric position in ``m_transform`` (Max-Feature-Map, MaxPool, Dropout in the reference)."""

    def __init__(self, what: str):
        super().__init__()
        self.what = what

    def extra_repr(self):
        return self.what


class BLSTMLayer(nn.Module):
    def __init__(self, input_dim, output_dim):
        super().__init__()
        self.l_blstm = nn.LSTM(input_dim, output_dim // 2, bidirectional=True)


# (index in m_transform, C_in, C_out, kernel, pad)  — src/models/lcnn.py:121-153
CONV_LAYERS = (
    (0, None, 64, 5, 2),
    (3, 32, 64, 1, 0),
    (6, 32, 96, 3, 1),
    (10, 48, 96, 1, 0),
    (13, 48, 128, 3, 1),
    (16, 64, 128, 1, 0),
    (19, 64, 64, 3, 1),
    (22, 32, 64, 1, 0),
    (25, 32, 64, 3, 1),
)
BN_LAYERS = ((5, 32), (9, 48), (12, 48), (18, 64), (21, 32), (24, 32))
POOL_SLOTS = (2, 8, 15, 27)


class LCNN(nn.Module):
    def __init__(self, device: str = "cuda", **kwargs):
        super().__init__()
        input_channels = kwargs.get("input_channels", 1)
        num_coefficients = kwargs.get("num_coefficients", 80)
        if input_channels != 1:
            raise ValueError("advb200 LCNN supports input_channels=1 (lfcc / mfcc frontends)")
        self.num_coefficients = num_coefficients
        self.v_emd_dim = 1
        self.device = device

        slots = [_Slot("mfm") for _ in range(29)]
        for idx, cin, cout, k, p in CONV_LAYERS:
            slots[idx] = nn.Conv2d(input_channels if cin is None else cin, cout, (k, k), 1, padding=(p, p))
        for idx, c in BN_LAYERS:
            slots[idx] = nn.BatchNorm2d(c, affine=False)
        for idx in POOL_SLOTS:
            slots[idx] = _Slot("maxpool2x2")
        slots[28] = _Slot("dropout(identity under attack/eval)")
        self.m_transform = nn.Sequential(*slots)

        feat = (num_coefficients // 16) * 32
        self.m_before_pooling = nn.Sequential(BLSTMLayer(feat, feat), BLSTMLayer(feat, feat))
        self.m_output_act = nn.Linear(feat, self.v_emd_dim)

        frontend_name = kwargs.get("frontend_algorithm", [])
        self.frontend = frontends.get_frontend(frontend_name)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        from .. import engine

        return engine.model_forward(self, x)
