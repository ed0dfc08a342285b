"""One forward+backward (advb_grad) of LCNN+LFCC (or --model rawnet3 / specrnet) at the bench shape, for ncu captures (diagnostic tool).

    ncu --set full --clock-control none --import-source on -k regex:conv_tc -s 16 -c 16 -o gpurun_out/prof \
        python tools/profile_grad.py [--batch 128] [--calls 2]
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "audio-deepfake-adversarial-attacks_b200"))

import torch  # noqa: E402

import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=128)
    ap.add_argument("--calls", type=int, default=2)
    ap.add_argument("--conv-path", type=int, default=0)
    ap.add_argument("--model", default="lcnn", choices=["lcnn", "specrnet", "rawnet3"])
    ap.add_argument("--samples", type=int, default=bench.T_SAMPLES)
    ap.add_argument("--opt", action="append", default=[], help="engine option name=value (repeatable)")
    ap.add_argument("--attack", action="store_true", help="run a 2-step PGD call (graph replay + fused update) instead of advb_grad")
    args = ap.parse_args()
    from advb200 import engine

    dev = torch.device("cuda:0")
    fe = {"lcnn": "lfcc", "specrnet": "mfcc", "rawnet3": "none"}[args.model]
    holder, state = bench.build_lcnn_state(args.model, fe)
    holder.load_state_dict(state)
    holder = holder.to(dev)
    x, y = bench.synthetic_batch(args.batch, 1002)
    x, y = x[:, :args.samples].contiguous().to(dev), y.to(dev)
    eng = engine.engine_for(holder, args.batch, args.samples)
    eng.set_option("conv_path", args.conv_path)
    for kv in args.opt:
        k, v = kv.split("=")
        eng.set_option(k, int(v))
    if args.attack:
        from advb200 import torchattacks as ta

        atk = ta.PGD(holder, eps=0.001, alpha=2 / 255, steps=4) if args.model != "rawnet3" else ta.PGDL2(holder, eps=0.1, alpha=0.2, steps=2)
        for _ in range(args.calls):
            adv = atk(x, y)
        torch.cuda.synchronize()
        print("attack linf", (adv - x).abs().max().item())
        return
    for _ in range(args.calls):
        g, logits = eng.grad(x, y)
    torch.cuda.synchronize()
    print("grad norm", g.norm().item(), "logit mean", logits.mean().item())


if __name__ == "__main__":
    main()
