"""Stage-by-stage parity report of the CUDA engine against the oracle (run on the GPU box).

    python tools/gpu_diag.py [--case NAME] [--out gpurun_out/diag.json]

Prints one line per check and writes them all to a JSON file.  Diagnostic tooling (uses oracle/ as the checker).
"""
import argparse
import json
import os
import sys
import traceback

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "audio-deepfake-adversarial-attacks_b200"))

import torch  # noqa: E402

import helpers  # noqa: E402
from oracle import attacks as oatk  # noqa: E402
from oracle import cases  # noqa: E402
from oracle import frontend as ofe  # noqa: E402

RESULTS = []


def report(check, **kw):
    rec = {"check": check}
    rec.update({k: (float(v) if isinstance(v, (int, float)) or hasattr(v, "item") else v) for k, v in kw.items()})
    RESULTS.append(rec)
    print(json.dumps(rec), flush=True)


def section(fn):
    def wrapped(*a, **k):
        try:
            return fn(*a, **k)
        except Exception as e:  # keep going: one broken stage must not hide the others
            report(fn.__name__ + ":EXCEPTION", error=repr(e), tb=traceback.format_exc()[-1500:])
    return wrapped


def interior(t, p):
    return t[:, p:t.shape[1] - p, p:t.shape[2] - p, :] if p > 0 else t


def cmp(name, got, want):
    got, want = got.detach().double().cpu(), want.detach().double().cpu()
    diff = (got - want).abs()
    report(name, max_abs=diff.max(), ref_max=want.abs().max(), rel=helpers.rel_err(got, want),
           cos=helpers.cosine(got, want), shape=list(got.shape))


@section
def check_frontend(eng, x, state):
    fb, dct, win, _ = ofe.tables_from_state(state)
    xc = x.cpu().clone().requires_grad_(True)
    want = ofe.cepstral_frontend(xc, fb, dct, win)
    got = eng.frontend_fwd(x)
    cmp("frontend_fwd", got, want)
    g = torch.Generator("cpu").manual_seed(5)
    gc = torch.randn(want.shape, generator=g)
    (gx_want,) = torch.autograd.grad((want * gc).sum(), xc)
    gx = eng.frontend_bwd(x, gc.to(x.device))
    cmp("frontend_bwd", gx, gx_want)


@section
def check_stages(eng, holder, x, y, state, fwd):
    taps = {}
    xc = x.cpu().clone().requires_grad_(True)
    o = fwd(xc, state, taps)
    for v in taps.values():
        v.retain_grad()
    z = torch.cat([-o, o], dim=1)
    cost = torch.nn.functional.cross_entropy(z, y.cpu())
    cost.backward()
    g, logits = eng.grad(x, y)
    cmp("logits", logits, o)
    t, p = eng.debug_stage("frontend")
    B = x.shape[0]
    cmp("stage_frontend", interior(t, p)[:B, :, :, 0], taps["frontend"][:, 0].transpose(1, 2))
    from oracle.lcnn import BLOCKS
    for i, (idx, _, _) in enumerate(BLOCKS):
        t, p = eng.debug_stage(f"block{i}")
        cmp(f"stage_block{i}", interior(t, p)[:B].permute(0, 3, 1, 2), taps[f"block{idx}"])
    for nm in ("feats", "lstm1", "lstm2"):
        t, _ = eng.debug_stage(nm)
        cmp("stage_" + nm, t[:B, :, 0, :], taps[nm])
    t, _ = eng.debug_stage("dfeats")
    cmp("grad_feats", t[:B, :, 0, :], taps["feats"].grad)
    for i, (idx, _, _) in reversed(list(enumerate(BLOCKS))):
        t, _ = eng.debug_stage(f"gblock{i}")
        cmp(f"grad_block{i}", t[:B].permute(0, 3, 1, 2), taps[f"block{idx}"].grad)
    t, _ = eng.debug_stage("gcoef")
    cmp("grad_frontend", t[:B, :, :, 0], taps["frontend"].grad[:, 0].transpose(1, 2))
    cmp("grad_x", g, xc.grad)
    agree = (torch.sign(g.cpu()) == torch.sign(xc.grad)).float().mean()
    report("grad_sign_agreement", frac=agree)


@section
def check_attacks(eng, holder, name, case, x, y, state, fwd):
    from advb200 import torchattacks as ta
    g = helpers.load_golden(name)
    for an, ap in cases.ATTACKS.items():
        want = helpers.oracle_attack(name, an, x.cpu(), y.cpu(), state, fwd, case)
        if an == "fgsm":
            got = ta.FGSM(holder, eps=ap["eps"])(x, y)
        elif an == "pgd":
            atk = ta.PGD(holder, eps=ap["eps"], alpha=ap["alpha"], steps=ap["steps"])
            got = atk.forward(x, y, noise=helpers.reference_start(case, "pgd", x.cpu(), ap["eps"]).to(x.device))
        else:
            atk = ta.PGDL2(holder, eps=ap["eps"], alpha=ap["alpha"], steps=ap["steps"])
            got = atk.forward(x, y, delta=helpers.reference_start(case, "pgdl2", x.cpu(), ap["eps"]).to(x.device))
        got = got.cpu()
        d_got, d_want = got - x.cpu(), want - x.cpu()
        with torch.no_grad():
            l_want = fwd(want, state)
        l_got = eng.forward(got.to(x.device)).cpu()
        report(f"attack_{an}", mismatch_frac=(got != want).float().mean(), max_abs=(got - want).abs().max(),
               linf_got=d_got.abs().max(), linf_want=d_want.abs().max(),
               l2_err=(d_got.norm(dim=1) - d_want.norm(dim=1)).abs().max(),
               golden_linf_err=abs(d_got.abs().amax(dim=1).numpy() - g[f"{an}_delta_linf"]).max(),
               golden_l2_err=abs(d_got.norm(dim=1).numpy() - g[f"{an}_delta_l2"]).max(),
               logits_got=l_got.flatten().tolist(), logits_want=l_want.flatten().tolist(),
               golden_logits=g[f"{an}_logits_adv"].ravel().tolist())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--case", default="all")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "diag.json"))
    args = ap.parse_args()
    from advb200 import engine
    dev = torch.device("cuda:0")
    names = list(cases.CASES) if args.case == "all" else [args.case]
    for name in names:
        report("CASE", name=name)
        case, x, y, holder, state, fwd = helpers.case_setup(name)
        holder = helpers.load_holder_state(holder, state, dev)
        xd, yd = x.to(dev), y.to(dev)
        eng = engine.engine_for(holder, x.shape[0], x.shape[1])
        report("engine", workspace_mb=eng.workspace_bytes / 2**20)
        check_frontend(eng, xd, state)
        check_stages(eng, holder, xd, yd, state, fwd)
        check_attacks(eng, holder, name, case, xd, yd, state, fwd)
        torch.cuda.synchronize()
        report("launches", n=eng.launches)
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    with open(args.out, "w") as f:
        json.dump(RESULTS, f, indent=1)


if __name__ == "__main__":
    main()
