"""RawNet3 parameter holder.

Same constructor surface, parameter names and shapes as the reference's ``src/models/rawnet3.py:11-71,161-291``
(``prepare_model()`` = Bottle2neck, scale 8, context, summed, ECA, nOut 1, sinc stride 10, log + mean-normalised sinc
features), so reference checkpoints load with ``load_state_dict``.  The sinc front layer holds the four state-dict
entries of asteroid-filterbanks' ``ParamSincFB`` (``conv1.filterbank.{low_hz_,band_hz_,window_,n_}``); the filters are
evaluated from them on the GPU on every call.  ``bn1`` / ``bn6`` exist in the checkpoint although the reference's
forward never applies them (rawnet3.py:35,69,80-137).  ``forward`` hands the waveform to the CUDA engine.
"""
import math

import numpy as np
import torch
from torch import nn

C_MAIN, N_SINC, SINC_K, SINC_STRIDE, SCALE = 1024, 256, 251, 10, 8


class SincParams(nn.Module):
    """Storage of ParamSincFB(256, 251, stride=10): mel-spaced initial cut-offs, Hamming half window, time axis."""

    def __init__(self, n_filters=N_SINC, kernel_size=SINC_K, sample_rate=16000.0, min_low_hz=50, min_band_hz=50):
        super().__init__()
        half = kernel_size // 2
        to_mel = lambda hz: 2595 * np.log10(1 + hz / 700)  # noqa: E731
        mel = np.linspace(to_mel(30), to_mel(sample_rate / 2 - (min_low_hz + min_band_hz)), n_filters // 2 + 1,
                          dtype="float32")
        hz = 700 * (10 ** (mel / 2595) - 1)
        self.low_hz_ = nn.Parameter(torch.from_numpy(hz[:-1]).view(-1, 1))
        self.band_hz_ = nn.Parameter(torch.from_numpy(np.diff(hz)).view(-1, 1))
        self.register_buffer("window_", torch.from_numpy(np.hamming(kernel_size)[:half]).float())
        self.register_buffer("n_", 2 * math.pi * (torch.arange(-half, 0.0).view(1, -1) / sample_rate))


class SincEncoder(nn.Module):
    def __init__(self):
        super().__init__()
        self.filterbank = SincParams()


class PreEmphasis(nn.Module):
    def __init__(self, coef: float = 0.97):
        super().__init__()
        self.coef = coef
        self.register_buffer("flipped_filter", torch.FloatTensor([-coef, 1.0]).unsqueeze(0).unsqueeze(0))


class AFMS(nn.Module):
    def __init__(self, nb_dim: int):
        super().__init__()
        self.alpha = nn.Parameter(torch.ones((nb_dim, 1)))
        self.fc = nn.Linear(nb_dim, nb_dim)


class Bottle2neck(nn.Module):
    def __init__(self, inplanes, planes, kernel_size=None, dilation=None, scale=4, pool=False):
        super().__init__()
        width = int(math.floor(planes / scale))
        self.conv1 = nn.Conv1d(inplanes, width * scale, kernel_size=1)
        self.bn1 = nn.BatchNorm1d(width * scale)
        self.nums = scale - 1
        pad = math.floor(kernel_size / 2) * dilation
        self.convs = nn.ModuleList(
            [nn.Conv1d(width, width, kernel_size=kernel_size, dilation=dilation, padding=pad) for _ in range(self.nums)])
        self.bns = nn.ModuleList([nn.BatchNorm1d(width) for _ in range(self.nums)])
        self.conv3 = nn.Conv1d(width * scale, planes, kernel_size=1)
        self.bn3 = nn.BatchNorm1d(planes)
        self.width = width
        self.afms = AFMS(planes)
        if inplanes != planes:
            self.residual = nn.Sequential(nn.Conv1d(inplanes, planes, kernel_size=1, stride=1, bias=False))
        else:
            self.residual = nn.Identity()


class RawNet3(nn.Module):
    def __init__(self, block=Bottle2neck, model_scale=SCALE, context=True, summed=True, C=C_MAIN, **kwargs):
        super().__init__()
        supported = dict(encoder_type="ECA", nOut=1, out_bn=False, sinc_stride=SINC_STRIDE, log_sinc=True,
                         norm_sinc="mean")
        for k, v in supported.items():
            if kwargs.get(k, v) != v:
                raise ValueError(f"advb200 RawNet3 supports the reference's prepare_model() configuration only ({k}={v})")
        if not (context and summed and model_scale == SCALE and C == C_MAIN):
            raise ValueError("advb200 RawNet3 supports the reference's prepare_model() configuration only")
        self.preprocess = nn.Sequential(PreEmphasis(), nn.InstanceNorm1d(1, eps=1e-4, affine=True))
        self.conv1 = SincEncoder()
        self.bn1 = nn.BatchNorm1d(C // 4)
        self.layer1 = block(C // 4, C, kernel_size=3, dilation=2, scale=model_scale, pool=5)
        self.layer2 = block(C, C, kernel_size=3, dilation=3, scale=model_scale, pool=3)
        self.layer3 = block(C, C, kernel_size=3, dilation=4, scale=model_scale)
        self.layer4 = nn.Conv1d(3 * C, 1536, kernel_size=1)
        self.attention = nn.Sequential(
            nn.Conv1d(1536 * 3, 128, kernel_size=1), nn.ReLU(), nn.BatchNorm1d(128), nn.Conv1d(128, 1536, kernel_size=1),
            nn.Softmax(dim=2))
        self.bn5 = nn.BatchNorm1d(3072)
        self.fc6 = nn.Linear(3072, 1)
        self.bn6 = nn.BatchNorm1d(1)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        from .. import engine

        return engine.model_forward(self, x)


def prepare_model():
    """rawnet3.py:277-291."""
    return RawNet3(Bottle2neck, model_scale=8, context=True, summed=True, encoder_type="ECA", nOut=1, out_bn=False,
                   sinc_stride=10, log_sinc=True, norm_sinc="mean", grad_mult=1)
