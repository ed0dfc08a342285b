"""Seeded synthetic inputs shared by tests, bench and golden generation.

Test infrastructure — see oracle/__init__.py.
"""
import hashlib

import torch


def clips(cfg_id: int, batch: int, n_samples: int = 64000, silence: bool = False):
    """x in [0,1] (min-max scaled 0.1*randn), y in {0,1}; generator seed 1000+cfg_id."""
    g = torch.Generator("cpu").manual_seed(1000 + cfg_id)
    raw = 0.1 * torch.randn(batch, n_samples, generator=g)
    y = torch.randint(0, 2, (batch,), generator=g)
    if silence:
        a, b = (n_samples * 5) // 16, (n_samples * 15) // 32
        raw[:, a:b] = 0.0
    mn = raw.min(dim=1, keepdim=True)[0]
    mx = raw.max(dim=1, keepdim=True)[0]
    return (raw - mn) / (mx - mn), y


def pgd_noise(cfg_id: int, shape, eps: float):
    g = torch.Generator("cpu").manual_seed(2000 + cfg_id)
    return torch.empty(shape).uniform_(-eps, eps, generator=g)


def pgdl2_noise(cfg_id: int, shape):
    g = torch.Generator("cpu").manual_seed(2000 + cfg_id)
    normal = torch.empty(shape).normal_(generator=g)
    r = torch.empty(shape[0], 1).uniform_(0, 1, generator=g)
    return normal, r


def randomize_norm_stats(state: dict, seed: int = 7) -> dict:
    """Give every BatchNorm non-trivial running stats (and affine where present) so placement errors show."""
    g = torch.Generator("cpu").manual_seed(seed)
    out = {}
    for k, v in state.items():
        if k.endswith("running_mean"):
            out[k] = 0.2 * torch.randn(v.shape, generator=g)
        elif k.endswith("running_var"):
            out[k] = 0.5 + torch.rand(v.shape, generator=g)
        elif (k.endswith(".weight") or k.endswith(".bias")) and k.rsplit(".", 1)[0] + ".running_mean" in state:
            # affine BatchNorm (SpecRNet): non-trivial scale / shift
            out[k] = (0.75 + 0.5 * torch.rand(v.shape, generator=g)) if k.endswith(".weight") else \
                0.1 * torch.randn(v.shape, generator=g)
        else:
            out[k] = v.clone()
    return out


def state_digest(state: dict) -> str:
    h = hashlib.sha256()
    for k in sorted(state):
        v = state[k]
        if v.dtype.is_floating_point:
            h.update(k.encode())
            h.update(v.detach().cpu().contiguous().float().numpy().tobytes())
    return h.hexdigest()[:16]
