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
    "specrnet_mfcc_t16000": dict(model="specrnet", frontend="mfcc", T=16000, B=3, cfg_id=15, silence=False),
    "specrnet_lfcc_t64000": dict(model="specrnet", frontend="lfcc", T=64000, B=1, cfg_id=16, silence=False),
    # RawNet3 is a raw-waveform model (no spectral frontend, rawnet3.py:73-137)
    "rawnet3_t16000": dict(model="rawnet3", frontend="none", T=16000, B=2, cfg_id=17, silence=False),
    "rawnet3_t64000": dict(model="rawnet3", frontend="none", T=64000, B=1, cfg_id=18, silence=False,
                           attacks=("fgsm", "pgdl2")),
    "rawnet3_t16000_margin": dict(model="rawnet3", frontend="none", T=16000, B=3, cfg_id=19, silence=False,
                                  margin=0.01, labels=(1, 0, 1), attacks=("fab",)),
    # FAB / CW: clean logits sit at +margin (all predicted bonafide) and labels are fixed, so the clips labelled 1 are
    # correctly classified and must be pushed across a real margin; the clip labelled 0 is already misclassified and
    # exercises FAB's "attack only the correctly classified clips" path (fab.py:506-513)
    "lcnn_lfcc_t16000_margin": dict(model="lcnn", frontend="lfcc", T=16000, B=4, cfg_id=14, silence=False,
                                    margin=0.01, labels=(1, 1, 0, 1), attacks=("fab", "cw", "cw_strong")),
    # same clips, weights and labels (same cfg_id): FAB with norm='L2' (fab.py:184-194,216-219,236-240,617-665)
    "lcnn_lfcc_t16000_margin_l2": dict(model="lcnn", frontend="lfcc", T=16000, B=4, cfg_id=14, silence=False,
                                       margin=0.01, labels=(1, 1, 0, 1), attacks=("fab_l2",)),
}
DEFAULT_ATTACKS = ("fgsm", "pgd", "pgdl2")

ATTACKS = {
    "fgsm": dict(eps=0.005),
    "pgd": dict(eps=0.001, alpha=2 / 255, steps=3),
    "pgdl2": dict(eps=0.1, alpha=0.2, steps=3),
    "fab": dict(eps=0.3, steps=8, eta=10.0, alpha_max=0.1, beta=0.9),   # AttackEnum.FAB preset, fewer steps
    "fab_l2": dict(eps=2.0, steps=8, eta=1.05, alpha_max=0.1, beta=0.9, norm="L2"),
    "cw": dict(c=1e-4, kappa=0.0, steps=20, lr=0.01),
    # CW with the classification term dominating the fp32 rounding noise of tanh(atanh(.)): elements are comparable, clips flip
    # at steps 4-5 (best-adversarial mask path) and the batch-wide early stop fires at step 8 (cw.py:107-110)
    "cw_strong": dict(c=1e4, kappa=0.0, steps=20, lr=5e-4),
}


def build_holder(model: str, frontend: str, seed: int = 42):
    """advb200 parameter holder with seeded default init (the YAML ``data.seed`` of the reference is 42)."""
    from advb200.models import get_model

    torch.manual_seed(seed)
    cfg = {"input_channels": 1, "frontend_algorithm": [frontend]} if model != "rawnet3" else {}
    return get_model(model, cfg, "cpu")


def build_state(model: str, frontend: str, seed: int = 42, calibrate_on=None, forward_fn=None, margin: float = 0.0):
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
        state[key] = state[key] - o.median() + margin
    return holder, state


def case_inputs(case: dict):
    x, y = synth.clips(case["cfg_id"], case["B"], case["T"], silence=case["silence"])
    if case.get("labels") is not None:
        y = torch.tensor(case["labels"], dtype=torch.int64)
    return x, y
