"""Oracle: LFCC / MFCC frontend (test infrastructure — see oracle/__init__.py).

Restates what ``src/frontends.py:13-32`` instantiates: ``torchaudio.transforms.LFCC`` /``MFCC`` with
n_fft=512, win_length=400, hop=160 (torchaudio ``transforms/_transforms.py:701-718, 807-828``;
``functional/functional.py:116-145`` spectrogram, ``:390-399`` amplitude_to_DB).  The filterbank, DCT matrix
and window come from the model's state_dict buffers.  SURVEY.md App. A.1 gives the formulas.
"""
import torch

N_FFT = 512
HOP = 160
WIN = 400


def frames_of(n_samples: int) -> int:
    return 1 + n_samples // HOP


def power_spectrogram(x: torch.Tensor, window: torch.Tensor) -> torch.Tensor:
    """x (B,T) -> |STFT|^2 (B, F, 257).  functional.py:116-145 with center=True, reflect pad, power=2."""
    B, T = x.shape
    xp = torch.nn.functional.pad(x.unsqueeze(1), (N_FFT // 2, N_FFT // 2), mode="reflect").squeeze(1)
    w512 = torch.zeros(N_FFT, dtype=x.dtype)
    off = (N_FFT - WIN) // 2
    w512[off:off + WIN] = window.to(x.dtype)
    fr = xp.unfold(1, N_FFT, HOP)  # (B, F, 512)
    spec = torch.fft.rfft(fr * w512, dim=-1)  # (B, F, 257)
    return spec.real * spec.real + spec.imag * spec.imag


def db_scale(E: torch.Tensor, top_db: float = 80.0) -> torch.Tensor:
    """functional.py:390-399 for a 3-D input: ONE cutoff for the whole batch."""
    D = 10.0 * torch.log10(torch.clamp(E, min=1e-10))
    D = D - 10.0 * 0.0  # db_multiplier = log10(max(amin, ref=1.0)) = 0   (_transforms.py:330-333)
    return torch.max(D, D.amax() - top_db)


def cepstral_frontend(x: torch.Tensor, fb: torch.Tensor, dct: torch.Tensor, window: torch.Tensor) -> torch.Tensor:
    """x (B,T) -> coefficients (B, 80, F), the layout torchaudio returns (_transforms.py:818-828 / 714-718)."""
    P = power_spectrogram(x, window)  # (B,F,257)
    E = P @ fb.to(x.dtype)  # (B,F,128)
    D = db_scale(E)
    C = D @ dct.to(x.dtype)  # (B,F,80)
    return C.transpose(1, 2)


def tables_from_state(state: dict):
    """(fb, dct, window, kind) from a model state_dict (keys per SURVEY.md F6)."""
    if "frontend.filter_mat" in state:
        return state["frontend.filter_mat"], state["frontend.dct_mat"], state["frontend.Spectrogram.window"], "lfcc"
    return (state["frontend.MelSpectrogram.mel_scale.fb"], state["frontend.dct_mat"],
            state["frontend.MelSpectrogram.spectrogram.window"], "mfcc")


def mel_spec(x: torch.Tensor, fb: torch.Tensor) -> torch.Tensor:
    """src/frontends.py:53-79 (prepare_mel_scale_vector / prepare_stft_features): torch.stft without a window (:65-71), the mel
    filterbank applied to the real and to the imaginary part separately (:74-75; MelScale.forward is ``fb^T @ x``), abs and angle of
    the complex result (:77-79), stacked on dim 1 (:58).  x (B,T), fb (257,n_mels) -> (B,2,n_mels,F)."""
    st = torch.stft(x, n_fft=N_FFT, return_complex=True, hop_length=HOP, win_length=WIN)
    mr = torch.matmul(st.real.transpose(-1, -2), fb).transpose(-1, -2)
    mi = torch.matmul(st.imag.transpose(-1, -2), fb).transpose(-1, -2)
    c = torch.complex(mr, mi)
    return torch.stack([c.abs(), c.angle()], dim=1)
