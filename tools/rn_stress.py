"""RawNet3 determinism stress: run the gradient evaluation N times and compare every debug stage bit-for-bit with the
first run (a race in the GEMM pipeline shows up as a run-to-run difference).  Diagnostic tool.

    python tools/rn_stress.py [--runs 40] [--batch 4] [--T 16000]
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "audio-deepfake-adversarial-attacks_b200"))

import torch  # noqa: E402

import bench  # noqa: E402

STAGES = ["rn_pre", "rn_sinc_raw", "rn_sinc", "rn_o1", "rn_cat1", "rn_y1", "rn_x1", "rn_y2", "rn_y3", "rn_cat4", "rn_layer4",
          "rn_pooled", "rn_gcat4", "rn_gx1", "rn_gsinc", "rn_gs", "rn_gn"]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--runs", type=int, default=40)
    ap.add_argument("--batch", type=int, default=4)
    ap.add_argument("--T", type=int, default=16000)
    args = ap.parse_args()
    from advb200 import engine

    dev = torch.device("cuda:0")
    holder, state = bench.build_lcnn_state("rawnet3", "none")
    holder.load_state_dict(state)
    holder = holder.to(dev)
    g = torch.Generator("cpu").manual_seed(3)
    x = torch.rand(args.batch, args.T, generator=g).to(dev)
    y = torch.randint(0, 2, (args.batch,), generator=g).to(dev)
    eng = engine.engine_for(holder, args.batch, args.T)
    ref = None
    bad = 0
    for i in range(args.runs):
        grad, logits = eng.grad(x, y)
        cur = {"grad": grad.clone(), "logits": logits.clone()}
        for s in STAGES:
            cur[s] = eng.debug_stage(s)[0].clone()
        if ref is None:
            ref = cur
            continue
        diffs = [k for k in cur if not torch.equal(cur[k], ref[k])]
        if diffs:
            bad += 1
            first = diffs[0]
            d = (cur[first].float() - ref[first].float()).abs()
            print(f"run {i}: {len(diffs)} stages differ, first = {first}, max abs diff {d.max().item():.3e}, "
                  f"elements {int((d > 0).sum())} of {d.numel()}", flush=True)
    print(f"{bad} of {args.runs - 1} runs differ from the first one")


if __name__ == "__main__":
    main()
