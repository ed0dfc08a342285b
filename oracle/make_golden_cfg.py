"""Goldens at the BENCHMARKED configurations (BASELINE.json configs[0..3]), produced by the UNMODIFIED reference on CPU.

    python -m oracle.make_golden_cfg [case ...]          (build container; minutes of CPU per case)

For every case the reference's own nn.Module (src/models/*) and attack class (adversarial_attacks/torchattacks) run
exactly as evaluate_models_on_adversarial_attacks.py:211-238 drives them: model.eval(), atk(x, y) with
set_training_mode(True, False), clean re-inference of the attacked batch, sigmoid, (p + .5).int().  Stored per case
(small: no waveforms): the calibrated output bias, clean / adversarial logits, predicted labels, per-clip L-inf / L2 of
the perturbation and 1 bit per sample (x_adv > x).  Test infrastructure.
"""
import os
import sys
import time
import warnings

import numpy as np
import torch

warnings.filterwarnings("ignore")

from . import cases, ref, synth  # noqa: E402

BIAS_KEY = {"lcnn": "m_output_act.bias", "specrnet": "fc2_gru.bias", "rawnet3": "fc6.bias"}

# name -> model / frontend / batch / cfg_id (inputs = synth.clips(cfg_id, B, 64000) = bench.py's synthetic_batch(B, 1000+cfg_id))
CFG_CASES = {
    # configs[0]: FGSM eps=0.005 on LCNN+LFCC, batch 8 (also run through generate_attacks() in tests/test_gpu_dropin.py)
    "cfg1_lcnn_fgsm_b8": dict(model="lcnn", frontend="lfcc", B=8, cfg_id=1, attack="fgsm", params=dict(eps=0.005)),
    # configs[1], the headline: torchattacks.PGD(model, eps=0.001) -> alpha 2/255, 40 steps, random start
    "cfg2_lcnn_pgd40_b128": dict(model="lcnn", frontend="lfcc", B=128, cfg_id=2, attack="pgd",
                                 params=dict(eps=0.001, alpha=2 / 255, steps=40)),
    # configs[2] at a CPU-affordable batch: same PGD-40 on SpecRNet+MFCC (the batch-wide dB floor is always active)
    "cfg3_specrnet_mfcc_pgd40_b32": dict(model="specrnet", frontend="mfcc", B=32, cfg_id=3, attack="pgd",
                                         params=dict(eps=0.001, alpha=2 / 255, steps=40)),
    # configs[3]: AttackEnum.PGDL2 (eps 0.1, alpha 0.2, 10 steps) on RawNet3, the 16 clips one GPU of the 8 holds
    "cfg4_rawnet3_pgdl2_b16": dict(model="rawnet3", frontend="none", B=16, cfg_id=4, attack="pgdl2",
                                   params=dict(eps=0.1, alpha=0.2, steps=10)),
}
T = 64000


def cfg_state(case, x=None):
    """Seeded weights (torch.manual_seed(42) + randomised BN statistics) WITHOUT the output-bias calibration."""
    _, state = cases.build_state(case["model"], case["frontend"])
    return state


def make_attack(ta, model, case):
    p = case["params"]
    if case["attack"] == "fgsm":
        return ta.FGSM(model, eps=p["eps"])
    if case["attack"] == "pgd":
        return ta.PGD(model, eps=p["eps"], alpha=p["alpha"], steps=p["steps"], random_start=True)
    if case["attack"] == "pgdl2":
        return ta.PGDL2(model, eps=p["eps"], alpha=p["alpha"], steps=p["steps"], random_start=True)
    raise ValueError(case["attack"])


def main():
    ta = ref.torchattacks()
    torch.set_num_threads(os.cpu_count() or 1)
    only = set(sys.argv[1:])
    for name, case in CFG_CASES.items():
        if only and name not in only:
            continue
        t0 = time.time()
        x, y = synth.clips(case["cfg_id"], case["B"], T)
        state = cfg_state(case)
        model = ref.model(case["model"], case["frontend"], state)
        model.eval()
        with torch.no_grad():
            clean0 = model(x)
        # calibrated synthetic checkpoint: shift the output bias so the clean logits straddle 0
        key = BIAS_KEY[case["model"]]
        state[key] = state[key] - clean0.median()
        model.load_state_dict(state)
        with torch.no_grad():
            clean = model(x)
        atk = make_attack(ta, model, case)
        atk.set_training_mode(model_training=True, batchnorm_training=False)
        torch.manual_seed(2000 + case["cfg_id"])  # the reference draws its random start from the global RNG (pgd.py:56)
        model.eval()
        xa = atk(x, y)
        model.eval()
        with torch.no_grad():
            la = model(xa)
        first = None
        if case["attack"] == "pgd":
            # the same attack stopped after ONE step: the only point where two fp32 implementations are comparable sample by
            # sample (PGD with alpha > 2 eps is chaotic over many steps: tools/pgd_divergence.py)
            one = dict(case, params=dict(case["params"], steps=1))
            atk1 = make_attack(ta, model, one)
            atk1.set_training_mode(model_training=True, batchnorm_training=False)
            torch.manual_seed(2000 + case["cfg_id"])
            model.eval()
            first = atk1(x, y)
            model.eval()
        pred_clean = (torch.sigmoid(clean.squeeze(1)) + .5).int()  # evaluate_...py:236-238
        pred_adv = (torch.sigmoid(la.squeeze(1)) + .5).int()
        out = {
            "digest": np.array(synth.state_digest(state)),
            "bias": state[key].numpy().astype(np.float32),
            "x_sum": np.array(x.double().sum().item()),
            "y": y.numpy(),
            "logits_clean": clean.numpy(),
            "logits_adv": la.numpy(),
            "pred_clean": pred_clean.numpy(),
            "pred_adv": pred_adv.numpy(),
            "delta_linf": (xa - x).abs().amax(dim=1).numpy(),
            "delta_l2": (xa - x).norm(p=2, dim=1).numpy(),
            "sign_bits": np.packbits((xa > x).numpy()),
            "moved_bits": np.packbits((xa != x).numpy()),
            **({"step1_sign_bits": np.packbits((first > x).numpy())} if first is not None else {}),
            "seconds": np.array(time.time() - t0),
            "threads": np.array(torch.get_num_threads()),
        }
        path = os.path.join(cases.GOLDEN_DIR, name + ".npz")
        np.savez_compressed(path, **out)
        acc_c = float((pred_clean == y.int()).float().mean())
        acc_a = float((pred_adv == y.int()).float().mean())
        print(f"{name}: clean acc {acc_c:.4f} adv acc {acc_a:.4f} flips {int((pred_clean != pred_adv).sum())} "
              f"min|logit_adv| {la.abs().min().item():.2e}  {time.time() - t0:.0f} s  {os.path.getsize(path) // 1024} KiB",
              flush=True)


if __name__ == "__main__":
    main()
