"""SpecRNet parity diagnostics on the GPU box (prints error metrics per stage)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "audio-deepfake-adversarial-attacks_b200")):
    sys.path.insert(0, p)
import torch  # noqa: E402

import helpers  # noqa: E402


def main():
    from advb200 import engine

    name = sys.argv[1] if len(sys.argv) > 1 else "specrnet_mfcc_t16000"
    dev = torch.device("cuda:0")
    case, x, y, holder, state, fwd = helpers.case_setup(name)
    holder = helpers.load_holder_state(holder, state, dev)
    eng = engine.engine_for(holder, x.shape[0], x.shape[1])
    grad, logits = eng.grad(x.to(dev), y.to(dev))
    B = x.shape[0]
    taps = {}
    xc = x.clone().requires_grad_(True)
    o = fwd(xc, state, taps)
    for v in taps.values():
        v.retain_grad()
    torch.nn.functional.cross_entropy(torch.cat([-o, o], dim=1), y).backward()

    def report(tag, a, b):
        sa = (torch.sign(a) == torch.sign(b)).float().mean().item()
        z = ((a == 0) != (b == 0)).float().mean().item()
        print(f"{tag:10s} rel {helpers.rel_err(a, b):.3e} trimmed {helpers.trimmed_rel_err(a, b):.3e} cos {helpers.cosine(a, b):.7f} "
              f"sign {sa:.5f} zero-mismatch {z:.5f} |b|max {b.abs().max().item():.3e}")

    for nm in ("0", "2", "4"):
        t, _ = eng.debug_stage(f"sr_gn{nm}")
        want = taps[f"stage{nm}"].grad
        report("g_xn" + nm, t[:B, :, :, :want.shape[1]].permute(0, 3, 2, 1).cpu(), want)
    t, _ = eng.debug_stage("gcoef")
    gc, want = t[:B].permute(0, 3, 2, 1).cpu(), taps["frontend"].grad
    report("gcoef", gc, want)
    for c in range(0, 80, 13):
        report(f" coef{c}", gc[:, :, c], want[:, :, c])
    for f in (0, 1, 50, gc.shape[-1] - 2, gc.shape[-1] - 1):
        report(f" frame{f}", gc[..., f], want[..., f])
    report("grad", grad.cpu(), xc.grad)


if __name__ == "__main__":
    main()
