"""Frontend buffer holders (LFCC / MFCC) for the B200 engine.

Mirrors the plugin surface of the reference's ``src/frontends.py:13-79``: module-level singletons
``LFCC_FN`` / ``MFCC_FN`` / ``MEL_SCALE_FN``, ``get_frontend(names)`` and the `mel_spec` functions
``prepare_mel_scale_vector`` / ``prepare_stft_features`` (forward only).  The holders carry exactly the buffers the
reference's torchaudio transforms register, under the same state_dict keys:

    LFCC : ``filter_mat`` (257,128), ``dct_mat`` (128,80), ``Spectrogram.window`` (400,)
    MFCC : ``dct_mat`` (128,80), ``MelSpectrogram.spectrogram.window`` (400,),
           ``MelSpectrogram.mel_scale.fb`` (257,128)

so a reference checkpoint loads into them unchanged and the engine reads the filterbank / DCT /
window from live buffer storage instead of regenerating them.  The arithmetic itself (framing, rFFT,
filterbank, dB, DCT and the backward of all of it) lives in ``csrc/frontend.cu``; these classes own no
math beyond building the constant tables, and calling one runs the CUDA kernels (CUDA input only).

Table formulas follow torchaudio 2.11 ``functional.linear_fbanks`` / ``melscale_fbanks`` /
``create_dct`` (functional.py:507-513, 518-587, 624-667), which the reference instantiates with
n_fft=512, win_length=400, hop=160, 128 filters, 80 coefficients, sr=16 kHz.
"""
import math
from typing import List

import torch
from torch import nn

SAMPLING_RATE = 16_000
N_FFT = 512
WIN_LENGTH = 400
HOP_LENGTH = 160
N_FILTER = 128
N_COEFF = 80
N_FREQS = N_FFT // 2 + 1

FRONTEND_LFCC = 1
FRONTEND_MFCC = 2


def _triangular_filterbank(all_freqs: torch.Tensor, f_pts: torch.Tensor) -> torch.Tensor:
    f_diff = f_pts[1:] - f_pts[:-1]
    slopes = f_pts.unsqueeze(0) - all_freqs.unsqueeze(1)
    down = (-1.0 * slopes[:, :-2]) / f_diff[:-1]
    up = slopes[:, 2:] / f_diff[1:]
    return torch.max(torch.zeros(1), torch.min(down, up))


def linear_fbanks() -> torch.Tensor:
    all_freqs = torch.linspace(0, SAMPLING_RATE // 2, N_FREQS)
    f_pts = torch.linspace(0.0, float(SAMPLING_RATE // 2), N_FILTER + 2)
    return _triangular_filterbank(all_freqs, f_pts)


def mel_fbanks(n_mels: int = N_FILTER) -> torch.Tensor:
    all_freqs = torch.linspace(0, SAMPLING_RATE // 2, N_FREQS)
    m_min = 2595.0 * math.log10(1.0 + 0.0 / 700.0)
    m_max = 2595.0 * math.log10(1.0 + (float(SAMPLING_RATE // 2) / 700.0))
    m_pts = torch.linspace(m_min, m_max, n_mels + 2)
    f_pts = 700.0 * (10.0 ** (m_pts / 2595.0) - 1.0)
    return _triangular_filterbank(all_freqs, f_pts)


def dct_matrix() -> torch.Tensor:
    n = torch.arange(float(N_FILTER))
    k = torch.arange(float(N_COEFF)).unsqueeze(1)
    dct = torch.cos(math.pi / float(N_FILTER) * (n + 0.5) * k)
    dct[0] *= 1.0 / math.sqrt(2.0)
    dct *= math.sqrt(2.0 / float(N_FILTER))
    return dct.t().contiguous()


class _Window(nn.Module):
    def __init__(self):
        super().__init__()
        self.register_buffer("window", torch.hann_window(WIN_LENGTH))


class _MelScale(nn.Module):
    def __init__(self):
        super().__init__()
        self.register_buffer("fb", mel_fbanks())


class _MelSpectrogram(nn.Module):
    def __init__(self):
        super().__init__()
        self.spectrogram = _Window()
        self.mel_scale = _MelScale()


class _Frontend(nn.Module):
    kind = 0
    top_db = 80.0

    def tables(self):
        """(filterbank (257,128), dct (128,80), window (400,)) — live buffers."""
        raise NotImplementedError

    def forward(self, waveform: torch.Tensor) -> torch.Tensor:
        from . import engine

        return engine.frontend_forward(self, waveform)


class LFCC(_Frontend):
    kind = FRONTEND_LFCC

    def __init__(self):
        super().__init__()
        self.Spectrogram = _Window()
        self.register_buffer("filter_mat", linear_fbanks())
        self.register_buffer("dct_mat", dct_matrix())

    def tables(self):
        return self.filter_mat, self.dct_mat, self.Spectrogram.window


class MFCC(_Frontend):
    kind = FRONTEND_MFCC

    def __init__(self):
        super().__init__()
        self.MelSpectrogram = _MelSpectrogram()
        self.register_buffer("dct_mat", dct_matrix())

    def tables(self):
        return self.MelSpectrogram.mel_scale.fb, self.dct_mat, self.MelSpectrogram.spectrogram.window


class MelScale(nn.Module):
    """Buffer holder of ``torchaudio.transforms.MelScale(n_mels=80, n_stft=257, sample_rate=16000)`` (src/frontends.py:34-38):
    state_dict key ``fb`` (257,80).  Calling it on a (..., 257, F) tensor is the linear map ``fb^T @ x`` of the reference;
    the `mel_spec` path itself runs in ``csrc/frontend.cu`` (fe_melspec_kernel)."""

    def __init__(self, n_mels: int = N_COEFF):
        super().__init__()
        self.register_buffer("fb", mel_fbanks(n_mels))


# module-level singletons shared by every model in the process, like the reference
MFCC_FN = MFCC()
LFCC_FN = LFCC()
MEL_SCALE_FN = MelScale()


def prepare_stft_features(audio: torch.Tensor, win_length: int = WIN_LENGTH, hop_length: int = HOP_LENGTH):
    """src/frontends.py:61-79: (|mel(STFT)|, angle(mel(STFT))), each (B, 80, F)."""
    if win_length != WIN_LENGTH or hop_length != HOP_LENGTH:
        raise NotImplementedError("advb200 implements the reference's own framing (win_length=400, hop_length=160)")
    from . import engine

    out = engine.mel_spec_forward(audio, MEL_SCALE_FN.fb)
    return out[..., 0, :, :], out[..., 1, :, :]


def prepare_mel_scale_vector(audio: torch.Tensor, win_length: int = WIN_LENGTH, hop_length: int = HOP_LENGTH) -> torch.Tensor:
    """src/frontends.py:53-58: torch.stack([abs, angle], dim=1) -> (B, 2, 80, F)."""
    if win_length != WIN_LENGTH or hop_length != HOP_LENGTH:
        raise NotImplementedError("advb200 implements the reference's own framing (win_length=400, hop_length=160)")
    from . import engine

    return engine.mel_spec_forward(audio, MEL_SCALE_FN.fb)


def get_frontend(frontends: List[str]):
    if "mfcc" in frontends:
        return MFCC_FN
    elif "lfcc" in frontends:
        return LFCC_FN
    elif "mel_spec" in frontends:
        return prepare_mel_scale_vector
    raise ValueError(f"{frontends} frontend is not supported!")


def frontend_kind_of(module) -> int:
    """Identify the frontend of a (reference or advb200) model by its registered buffers."""
    keys = set(dict(module.named_buffers()).keys())
    if "filter_mat" in keys:
        return FRONTEND_LFCC
    if "MelSpectrogram.mel_scale.fb" in keys:
        return FRONTEND_MFCC
    raise ValueError(f"unsupported frontend module {type(module).__name__} (buffers: {sorted(keys)})")


def frontend_tables(module):
    """(fb, dct, window) tensors from either a torchaudio transform or an advb200 holder."""
    bufs = dict(module.named_buffers())
    if "filter_mat" in bufs:
        return bufs["filter_mat"], bufs["dct_mat"], bufs["Spectrogram.window"]
    return bufs["MelSpectrogram.mel_scale.fb"], bufs["dct_mat"], bufs["MelSpectrogram.spectrogram.window"]
