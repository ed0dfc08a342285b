"""Locate disagreements between the tcgen05 and fp32-SIMT conv paths, stage by stage (diagnostic tool)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "audio-deepfake-adversarial-attacks_b200"))
import torch  # noqa: E402

import helpers  # noqa: E402
from advb200 import engine  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "lcnn_lfcc_t64000"
dev = torch.device("cuda:0")
case, x, y, holder, state, fwd = helpers.case_setup(name)
holder = helpers.load_holder_state(holder, state, dev)
eng = engine.engine_for(holder, x.shape[0], x.shape[1])
xd, yd = x.to(dev), y.to(dev)
stages = [f"block{i}" for i in range(9)] + [f"gblock{i}" for i in range(9)] + ["gcoef"]
snap = {}
for path in (1, 0):
    eng.set_option("conv_path", path)
    eng.grad(xd, yd)
    snap[path] = {s: eng.debug_stage(s)[0].clone() for s in stages}
B = x.shape[0]
for s in stages:
    a, b = snap[0][s][:B], snap[1][s][:B]
    d = (a - b).abs()
    rel = helpers.rel_err(a, b)
    line = f"{s:8s} shape {tuple(a.shape)} rel {rel:.3e} max {d.max().item():.3e}"
    if rel > 1e-4:
        bad = (d > 1e-3 * b.abs().max()).nonzero()
        ys = bad[:, 1].unique().tolist()
        xs = bad[:, 2].unique().tolist()
        cs = bad[:, 3].unique().tolist()
        line += f"\n   bad n={bad.shape[0]} rows(y)={ys[:40]} cols(x)={xs[:40]} ch={cs[:64]}"
    print(line)
