"""Golden-case definitions shared by oracle/make_golden.py and the tests (test infrastructure)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "audio-deepfake-adversarial-attacks_b200")
if PKG not in sys.path:
    sys.path.insert(0, PKG)

from . import synth  # noqa: E402

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")

# name -> dict(model, frontend, T, B, cfg_id, silence)
CASES = {
    "lcnn_lfcc_t16000": dict(model="lcnn", frontend="lfcc", T=16000, B=2, cfg_id=11, silence=False),
    "lcnn_lfcc_t16000_silence": dict(model="lcnn", frontend="lfcc", T=16000, B=2, cfg_id=12, silence=True),
    "lcnn_lfcc_t64000": dict(model="lcnn", frontend="lfcc", T=64000, B=1, cfg_id=13, silence=False),
}

ATTACKS = {
    "fgsm": dict(eps=0.005),
    "pgd": dict(eps=0.001, alpha=2 / 255, steps=3),
    "pgdl2": dict(eps=0.1, alpha=0.2, steps=3),
}


def build_holder(model: str, frontend: str, seed: int = 42):
    """advb200 parameter holder with seeded default init (the YAML ``data.seed`` of the reference is 42)."""
    from advb200.models import get_model

    torch.manual_seed(seed)
    cfg = {"input_channels": 1, "frontend_algorithm": [frontend]}
    return get_model(model, cfg, "cpu")


def build_state(model: str, frontend: str, seed: int = 42, calibrate_on=None, forward_fn=None):
    """Seeded state_dict with randomised BatchNorm statistics; optionally shift the output bias so that the clean
    logits of ``calibrate_on`` straddle zero (otherwise no label ever flips)."""
    holder = build_holder(model, frontend, seed)
    # (.cpu(): the frontend singleton may have been moved to a GPU by an earlier model in this process)
    state = {k: v.detach().cpu().clone() for k, v in holder.state_dict().items()}
    state = synth.randomize_norm_stats(state)
    if calibrate_on is not None:
        with torch.no_grad():
            o = forward_fn(calibrate_on, state)
        key = {"lcnn": "m_output_act.bias", "specrnet": "fc2_gru.bias", "rawnet3": "fc6.bias"}[model]
        state[key] = state[key] - o.median()
    return holder, state


def case_inputs(case: dict):
    x, y = synth.clips(case["cfg_id"], case["B"], case["T"], silence=case["silence"])
    return x, y
