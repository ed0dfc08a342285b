#!/usr/bin/env python
"""Diagnostic: the Toeplitz first-block kernel (conv0_fwd=0) against the im2col kernel (conv0_fwd=1) on the same input:
block-0 output, logits, and the per-kernel time of both at the benchmark size.  Run under gpurun."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "audio-deepfake-adversarial-attacks_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402

import helpers  # noqa: E402
from advb200 import engine  # noqa: E402

dev = torch.device("cuda:0")
for name in ("lcnn_lfcc_t16000", "lcnn_lfcc_t64000"):
    case, x, y, holder, state, fwd = helpers.case_setup(name)
    holder = helpers.load_holder_state(holder, state, dev)
    eng = engine.engine_for(holder, x.shape[0], x.shape[1])
    xd = x.to(dev)
    outs = {}
    for mode in (1, 0):
        eng.set_option("conv0_fwd", mode)
        logits = eng.forward(xd).cpu()
        blk, p = eng.debug_stage("block0")
        outs[mode] = (logits, blk.cpu())
    g = helpers.load_golden(name)
    d = (outs[0][1] - outs[1][1]).abs()
    print(name, "swap", os.environ.get("ADVB_C0T_SWAP", "0"), "block0 max|toeplitz - im2col|", d.max().item(), "mean", d.mean().item(),
          "| logits toeplitz", outs[0][0].flatten().tolist(), "im2col", outs[1][0].flatten().tolist(), "golden", g["logits"].ravel().tolist())
    gr = {}
    for mode in (1, 0):
        eng.set_option("conv0_fwd", mode)
        gr[mode] = eng.grad(xd, y.to(dev))[0].cpu()
    print("   grad cosine toeplitz vs im2col", helpers.cosine(gr[0], gr[1]), "sign agree", (gr[0].sign() == gr[1].sign()).float().mean().item())

if "--time" in sys.argv:
    from oracle import cases

    holder, state = cases.build_state("lcnn", "lfcc")
    holder = helpers.load_holder_state(holder, state, dev)
    B = 128
    x = torch.rand(B, 64000, device=dev)
    eng = engine.engine_for(holder, B, 64000)
    for mode in (1, 0):
        eng.set_option("conv0_fwd", mode)
        for _ in range(2):
            eng.forward(x)
        eng.profile_begin()
        for _ in range(5):
            eng.forward(x)
        prof = {r["name"]: r for r in eng.profile_end()}
        r = prof["conv_fwd_b0"]
        print("conv0_fwd", mode, "conv_fwd_b0 avg ms", r["total_ms"] / r["count"])
