#!/usr/bin/env python
"""Experiment (VERDICT r01 item 3 i): forward 3x3 blocks with the tf32 main term + ONE bf16 MMA for both cross terms
(tf32_passes = 2) against 3xTF32 (tf32_passes = 3): accuracy on the reference fixtures and per-kernel time.  Run under gpurun."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "audio-deepfake-adversarial-attacks_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402
import helpers  # noqa: E402
from advb200 import engine  # noqa: E402

dev = torch.device("cuda:0")
out = {}
# ---- accuracy on the small goldens: logits vs the reference, gradient vs 3xTF32
for name in ("lcnn_lfcc_t16000", "lcnn_lfcc_t64000"):
    case, x, y, holder, state, fwd = helpers.case_setup(name)
    holder = helpers.load_holder_state(holder, state, dev)
    eng = engine.engine_for(holder, x.shape[0], x.shape[1])
    g = helpers.load_golden(name)
    res = {}
    for p in (3, 2, 1):
        eng.set_option("tf32_passes", p)
        grad, logits = eng.grad(x.to(dev), y.to(dev))
        res[p] = (grad.cpu(), logits.cpu())
    eng.set_option("tf32_passes", 3)
    ref_g = torch.from_numpy(g["grad"])
    for p in (3, 2, 1):
        gr, lg = res[p]
        out[f"{name}_p{p}"] = {"max_abs_dlogit_vs_reference": float(np.abs(lg.numpy() - g["logits"]).max()),
                               "grad_cosine_vs_reference": helpers.cosine(gr, ref_g),
                               "grad_sign_mismatch_vs_reference": float((gr.sign() != ref_g.sign()).float().mean()),
                               "grad_trimmed_rel_err_vs_reference": helpers.trimmed_rel_err(gr, ref_g)}
# ---- the benchmarked configuration: first PGD step and the 40-step result against the B = 128 fixture
job = bench.Job("lcnn", 0, dev, 0)
fx = job.fixture
B, T = job.B, bench.T_SAMPLES
sign1 = np.unpackbits(fx["step1_sign_bits"])[: B * T].reshape(B, T).astype(bool)
sign40 = np.unpackbits(fx["sign_bits"])[: B * T].reshape(B, T).astype(bool)
from advb200 import torchattacks as ta  # noqa: E402

torch.manual_seed(2002)
noise = torch.empty(B, T).uniform_(-bench.EPS, bench.EPS).to(dev)
for p in (3, 2, 1):
    job.eng.set_option("tf32_passes", p)
    a1 = ta.PGD(job.holder, eps=bench.EPS, alpha=bench.ALPHA, steps=1).forward(job.x, job.y, noise=noise)
    a40 = job.atk.forward(job.x, job.y, noise=noise)
    la = job.eng.forward(a40).flatten().cpu()
    pred = (torch.sigmoid(la) + .5).int().numpy()
    for _ in range(2):
        job.atk(job.x, job.y)
    job.eng.profile_begin()
    job.atk(job.x, job.y)
    prof = {r["name"]: r for r in job.eng.profile_end()}
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(3):
        job.atk(job.x, job.y)
    ev1.record()
    torch.cuda.synchronize()
    out[f"cfg2_p{p}"] = {"step1_sign_mismatch": int(((a1 > job.x).cpu().numpy() != sign1).sum()),
                         "step40_sign_mismatch_frac": float(((a40 > job.x).cpu().numpy() != sign40).mean()),
                         "label_mismatch": int((pred != fx["pred_adv"]).sum()),
                         "max_abs_dlogit_adv": float(np.abs(la.numpy() - fx["logits_adv"].ravel()).max()),
                         "clips_per_s": 3 * B / (ev0.elapsed_time(ev1) * 1e-3),
                         "kernel_us": {k: round(1e3 * prof[k]["total_ms"] / prof[k]["count"], 1)
                                       for k in ("conv_fwd_b2", "conv_fwd_b4", "conv_fwd_b6", "conv_fwd_b8",
                                                 "conv_bwd_b2", "conv_bwd_b4", "conv_bwd_b6", "conv_bwd_b8")}}
job.eng.set_option("tf32_passes", 3)
print("MIXCHECK " + json.dumps(out))
for k, v in out.items():
    print(k, v)
