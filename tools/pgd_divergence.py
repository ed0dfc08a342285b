#!/usr/bin/env python
"""How far do two CORRECT fp32 runs of the reference's own PGD drift apart?  (CPU, reference classes from oracle/_ref.)

PGD as the repo configures it is degenerate (alpha = 2/255 > 2 eps): after every step the iterate is
x + eps * sign(g), so a sample whose gradient is an fp32 tie flips by 2 eps, the next gradient is evaluated at a slightly
different point, more near-zero samples flip, and so on.  This script runs torchattacks.PGD on the reference LCNN from the same
random start with 8 threads and with 1 thread (different summation order inside MKL / ATen, nothing else) and reports the
fraction of samples whose final sign differs after 1, 5, 10, 20 and 40 steps, next to the attack's actual success criterion
(predicted labels).  Output: tests/golden/pgd_divergence_reference.json (committed; re-run with `python tools/pgd_divergence.py`).
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["CUDA_VISIBLE_DEVICES"] = ""
import torch  # noqa: E402

from oracle import cases, ref, synth  # noqa: E402
from oracle.make_golden_cfg import BIAS_KEY  # noqa: E402
import numpy as np  # noqa: E402


def main():
    ta = ref.torchattacks()
    B, T, eps = 8, 64000, 0.001
    which = sys.argv[1] if len(sys.argv) > 1 else "lcnn"   # "lcnn" (cfg 2) or "specrnet" (cfg 3: SpecRNet + MFCC)
    if which == "specrnet":
        x, y = synth.clips(3, B, T)
        _, state = cases.build_state("specrnet", "mfcc")
        state[BIAS_KEY["specrnet"]] = torch.from_numpy(
            np.load(os.path.join(cases.GOLDEN_DIR, "cfg3_specrnet_mfcc_pgd40_b32.npz"))["bias"])
        model = ref.model("specrnet", "mfcc", state)
    else:
        x, y = synth.clips(2, B, T)
        _, state = cases.build_state("lcnn", "lfcc")
        state[BIAS_KEY["lcnn"]] = torch.from_numpy(np.load(os.path.join(cases.GOLDEN_DIR, "cfg2_lcnn_pgd40_b128.npz"))["bias"])
        model = ref.model("lcnn", "lfcc", state)
    rows = []
    for steps in ((1, 10, 40) if which == "specrnet" else (1, 5, 10, 20, 40)):
        outs = {}
        for threads in (8, 1):
            torch.set_num_threads(threads)
            atk = ta.PGD(model, eps=eps, alpha=2 / 255, steps=steps, random_start=True)
            atk.set_training_mode(model_training=True, batchnorm_training=False)
            torch.manual_seed(2002)
            model.eval()
            adv = atk(x, y)
            model.eval()
            with torch.no_grad():
                outs[threads] = (adv, model(adv).flatten())
        a, b = outs[8], outs[1]
        row = {"steps": steps, "clips": B, "sign_mismatch_frac": float(((a[0] > x) != (b[0] > x)).float().mean()),
               "label_mismatch": int(((a[1] > 0) != (b[1] > 0)).sum()), "max_abs_dlogit": float((a[1] - b[1]).abs().max())}
        print(row, flush=True)
        rows.append(row)
    out = {"what": f"reference torchattacks.PGD on the reference {which} (CPU), 8 threads vs 1 thread, same random start",
           "eps": eps, "alpha": 2 / 255, "T": T, "rows": rows}
    name = "pgd_divergence_reference.json" if which == "lcnn" else f"pgd_divergence_reference_{which}.json"
    json.dump(out, open(os.path.join(cases.GOLDEN_DIR, name), "w"), indent=1)


if __name__ == "__main__":
    main()
