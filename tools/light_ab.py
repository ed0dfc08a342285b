"""A/B of the persistent light conv kernels against the one-tile-per-CTA kernels: dump stage tensors of one gradient
evaluation (run once per setting of ADVB_LIGHT_PERSISTENT), then `compare` the two dumps (diagnostic tool)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "audio-deepfake-adversarial-attacks_b200")):
    sys.path.insert(0, p)
import torch  # noqa: E402


def dump(tag, name):
    import helpers
    from advb200 import engine

    dev = torch.device("cuda:0")
    case, x, y, holder, state, fwd = helpers.case_setup(name)
    holder = helpers.load_holder_state(holder, state, dev)
    eng = engine.engine_for(holder, x.shape[0], x.shape[1])
    g, logits = eng.grad(x.to(dev), y.to(dev))
    out = {"grad": g.cpu(), "logits": logits.cpu()}
    for i in range(9):
        out[f"block{i}"] = eng.debug_stage(f"block{i}")[0].cpu()
        out[f"gblock{i}"] = eng.debug_stage(f"gblock{i}")[0].cpu()
    out["gcoef"] = eng.debug_stage("gcoef")[0].cpu()
    torch.save(out, os.path.join(ROOT, "gpurun_out", f"ab_{tag}.pt"))


def compare():
    a = torch.load(os.path.join(ROOT, "gpurun_out", "ab_0.pt"))
    b = torch.load(os.path.join(ROOT, "gpurun_out", "ab_1.pt"))
    for k in a:
        d = (a[k] - b[k]).abs()
        nz = int((d > 0).sum())
        print(f"{k:10s} max|d| {d.max().item():.3e} differing {nz}/{d.numel()} max|a| {a[k].abs().max().item():.3e}")
        if nz and a[k].dim() == 4:
            idx = (d > 0).nonzero()
            print("   first", idx[0].tolist(), "last", idx[-1].tolist(), "rows", sorted(set(idx[:, 1].tolist()))[:12])


if __name__ == "__main__":
    if sys.argv[1] == "compare":
        compare()
    else:
        dump(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else "lcnn_lfcc_t64000")
