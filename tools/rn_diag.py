"""RawNet3: stage-by-stage parity report of the CUDA engine against the oracle, on both GEMM paths (run on the GPU box).

    python tools/rn_diag.py [--case rawnet3_t16000] [--out gpurun_out/rn_diag.json] [--bench B]

Diagnostic tooling (uses oracle/ as the checker).
"""
import argparse
import json
import os
import sys
import time
import traceback

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "audio-deepfake-adversarial-attacks_b200"))

import torch  # noqa: E402

import helpers  # noqa: E402
from oracle import attacks as oatk  # noqa: E402
from oracle import cases  # noqa: E402

RESULTS = []


def report(check, **kw):
    rec = {"check": check}
    rec.update({k: (float(v) if isinstance(v, (int, float)) or hasattr(v, "item") else v) for k, v in kw.items()})
    RESULTS.append(rec)
    print(json.dumps(rec), flush=True)


def cmp(name, got, want):
    got, want = got.detach().double().cpu(), want.detach().double().cpu()
    diff = (got - want).abs()
    report(name, max_abs=diff.max(), med_abs=diff.median(), ref_max=want.abs().max(), rel=helpers.rel_err(got, want),
           cos=helpers.cosine(got, want), nan=int(torch.isnan(got).sum()), shape=list(got.shape))


def stage(eng, name, B):
    t, _ = eng.debug_stage(name)  # (Bmax, rows, 1, C)
    return t[:B, :, 0, :]


def oracle_taps(x, y, state, fwd):
    taps = {}
    xc = x.clone().requires_grad_(True)
    o = fwd(xc, state, taps)
    keep = {k: v for k, v in taps.items() if k in ("sinc", "x1", "x3", "layer4", "pooled", "sinc_raw", "pre")}
    for v in keep.values():
        v.retain_grad()
    cost = torch.nn.CrossEntropyLoss()(torch.cat([-o, o], dim=1), y)
    cost.backward()
    return o.detach(), taps, xc.grad, {k: v.grad for k, v in keep.items()}


def run_path(tag, eng, conv_path, x, y, xd, yd, o_want, taps, g_want, tgrads, B):
    eng.set_option("conv_path", conv_path)
    torch.cuda.synchronize()
    try:
        g, logits = eng.grad(xd, yd)
        torch.cuda.synchronize()
    except Exception as e:
        report(tag + ":EXCEPTION", error=repr(e), tb=traceback.format_exc()[-1500:])
        return None
    cmp(tag + "/pre", stage(eng, "rn_pre", B)[:, :, 0], taps["pre"][:, 0, :])
    cmp(tag + "/sinc_raw", stage(eng, "rn_sinc_raw", B), taps["sinc_raw"].transpose(1, 2))
    pad = 2
    s = stage(eng, "rn_sinc", B)
    cmp(tag + "/sinc", s[:, pad:s.shape[1] - pad], taps["sinc"].transpose(1, 2))
    for nm, key, pd in (("rn_y1", "layer1.pre_pool", 2), ("rn_y2", "layer2.pre_pool", 3), ("rn_y3", "layer3.pre_pool", 4)):
        t = stage(eng, nm, B)
        cmp(tag + "/" + key, t[:, pd:t.shape[1] - pd], taps[key].transpose(1, 2))
    t = stage(eng, "rn_x1", B)
    cmp(tag + "/x1", t[:, 3:t.shape[1] - 3], taps["x1"].transpose(1, 2))
    c4 = stage(eng, "rn_cat4", B)
    cmp(tag + "/x2", c4[:, :, 1024:2048], taps["x2"].transpose(1, 2))
    cmp(tag + "/x3", c4[:, :, 2048:], taps["x3"].transpose(1, 2))
    cmp(tag + "/layer4", stage(eng, "rn_layer4", B), taps["layer4"].transpose(1, 2))
    cmp(tag + "/pooled", stage(eng, "rn_pooled", B)[:, 0], taps["pooled"])
    cmp(tag + "/logits", logits.cpu(), o_want)
    # backward stages
    gc4 = stage(eng, "rn_gcat4", B)
    cmp(tag + "/g_x3", gc4[:, :, 2048:], tgrads["x3"].transpose(1, 2))
    cmp(tag + "/g_x1", stage(eng, "rn_gx1", B), tgrads["x1"].transpose(1, 2))
    gs = stage(eng, "rn_gsinc", B)
    cmp(tag + "/g_sinc", gs[:, pad:gs.shape[1] - pad], tgrads["sinc"].transpose(1, 2))
    cmp(tag + "/g_pre", stage(eng, "rn_gn", B)[:, :, 0], tgrads["pre"][:, 0, :])
    cmp(tag + "/grad", g.cpu(), g_want)
    report(tag + "/grad_metrics", trimmed=helpers.trimmed_rel_err(g.cpu(), g_want),
           sign_agree=(torch.sign(g.cpu()) == torch.sign(g_want)).float().mean())
    return g.cpu(), logits.cpu()


def at_engine_features(tag, eng, conv_path, x, y, xd, yd, state, B):
    """Backward segments pinned separately: the oracle tail evaluated at the ENGINE's sinc outputs, then the two linear
    segments (transposed sinc convolution, preprocess VJP) fed with the engine's own intermediate gradients."""
    import torch.nn.functional as F
    from oracle import rawnet3 as orn

    eng.set_option("conv_path", conv_path)
    g, logits = eng.grad(xd, yd)
    S = stage(eng, "rn_sinc_raw", B).cpu().transpose(1, 2).contiguous().requires_grad_(True)
    o = orn.tail(S, state)
    cmp(tag + "/at_S/logits", logits.cpu(), o.detach())
    cost = torch.nn.CrossEntropyLoss()(torch.cat([-o, o], dim=1), y)
    (gS,) = torch.autograd.grad(cost, S)
    GS = stage(eng, "rn_gs", B).cpu().transpose(1, 2).contiguous()
    cmp(tag + "/at_S/g_s", GS, gS)
    report(tag + "/at_S/g_s_metrics", trimmed=helpers.trimmed_rel_err(GS, gS), sign_agree=(torch.sign(GS) == torch.sign(gS)).float().mean())
    filt = orn.sinc_filters(state)
    gpre = F.conv_transpose1d(GS, filt, stride=10)
    gpre = F.pad(gpre, (0, x.shape[1] - gpre.shape[-1]))[:, 0]
    gn = stage(eng, "rn_gn", B).cpu()[:, :, 0]
    cmp(tag + "/at_GS/g_pre", gn, gpre)
    xc = x.clone().requires_grad_(True)
    (gx,) = torch.autograd.grad(orn.preprocess(xc, state), xc, grad_outputs=gn.unsqueeze(1))
    cmp(tag + "/at_gn/grad", g.cpu(), gx)
    cmp(tag + "/filters", stage(eng, "rn_filt", 1).cpu().view(256, 251), filt[:, 0])


def attacks(tag, eng, conv_path, name, case, x, y, xd, yd, holder, state, fwd, g):
    from advb200 import torchattacks as ta

    eng.set_option("conv_path", conv_path)
    for attack in case.get("attacks", cases.DEFAULT_ATTACKS):
        p = cases.ATTACKS[attack]
        if attack == "fgsm":
            atk = ta.FGSM(holder, eps=p["eps"])
            got = atk(xd, yd)
        elif attack == "pgd":
            atk = ta.PGD(holder, eps=p["eps"], alpha=p["alpha"], steps=p["steps"])
            got = atk.forward(xd, yd, noise=helpers.reference_start(case, "pgd", x, p["eps"]).to(xd.device))
        elif attack == "pgdl2":
            atk = ta.PGDL2(holder, eps=p["eps"], alpha=p["alpha"], steps=p["steps"])
            got = atk.forward(xd, yd, delta=helpers.reference_start(case, "pgdl2", x, p["eps"]).to(xd.device))
        else:
            atk = ta.FAB(holder, norm="Linf", eps=p["eps"], steps=p["steps"], eta=p["eta"], alpha_max=p["alpha_max"],
                         beta=p["beta"], n_classes=2)
            got = atk(xd, yd)
        got = got.cpu()
        d = got - x
        la = eng.forward(got.to(xd.device)).cpu().numpy().ravel().tolist()
        rec = dict(linf=d.abs().amax(1).tolist(), linf_ref=g[f"{attack}_delta_linf"].tolist(), l2=d.norm(dim=1).tolist(),
                   l2_ref=g[f"{attack}_delta_l2"].tolist(), logits_adv=la, logits_adv_ref=g[f"{attack}_logits_adv"].ravel().tolist())
        if g[f"{attack}_adv"].size:
            ref = torch.from_numpy(g[f"{attack}_adv"])
            rec.update(mismatch=(got != ref).float().mean().item(), maxdiff=(got - ref).abs().max().item(),
                       cos=helpers.cosine(d, ref - x))
            lr = eng.forward(ref.to(xd.device)).cpu().numpy().ravel().tolist()
            rec.update(logits_at_ref_adv=lr)
        report(f"{tag}/attack/{attack}", **rec)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--case", default="rawnet3_t16000")
    ap.add_argument("--out", default="gpurun_out/rn_diag.json")
    ap.add_argument("--bench", type=int, default=0)
    ap.add_argument("--bench-T", type=int, default=64000)
    args = ap.parse_args()
    from advb200 import engine

    dev = torch.device("cuda:0")
    case, x, y, holder, state, fwd = helpers.case_setup(args.case)
    holder = helpers.load_holder_state(holder, state, dev)
    B = x.shape[0]
    eng = engine.engine_for(holder, B, x.shape[1])
    report("workspace", MiB=eng.workspace_bytes / 2 ** 20)
    xd, yd = x.to(dev), y.to(dev)
    o_want, taps, g_want, tgrads = oracle_taps(x, y, state, fwd)
    g = helpers.load_golden(args.case)
    res = {}
    for tag, path in (("simt", 1), ("tc", 0)):
        res[tag] = run_path(tag, eng, path, x, y, xd, yd, o_want, taps, g_want, tgrads, B)
        if res[tag] is not None:
            cmp(tag + "/logits_vs_golden", res[tag][1], torch.from_numpy(g["logits"]))
            cmp(tag + "/grad_vs_golden", res[tag][0], torch.from_numpy(g["grad"]))
    for tag, path in (("simt", 1), ("tc", 0)):
        try:
            at_engine_features(tag, eng, path, x, y, xd, yd, state, B)
            attacks(tag, eng, path, args.case, case, x, y, xd, yd, holder, state, fwd, g)
        except Exception as e:
            report(tag + ":EXCEPTION2", error=repr(e), tb=traceback.format_exc()[-1500:])
    if res.get("simt") and res.get("tc"):
        cmp("tc_vs_simt/grad", res["tc"][0], res["simt"][0])
        cmp("tc_vs_simt/logits", res["tc"][1], res["simt"][1])
    if args.bench:
        from oracle import synth

        xb, yb = synth.clips(4, args.bench, args.bench_T)
        eb = engine.engine_for(holder, args.bench, args.bench_T)
        report("bench_workspace", GiB=eb.workspace_bytes / 2 ** 30)
        xb, yb = xb.to(dev), yb.to(dev)
        for tag, path in (("tc", 0), ("simt", 1)):
            eb.set_option("conv_path", path)
            eb.grad(xb, yb)
            torch.cuda.synchronize()
            t0 = time.time()
            n = 3 if path == 0 else 1
            for _ in range(n):
                eb.grad(xb, yb)
            torch.cuda.synchronize()
            dt = (time.time() - t0) / n
            report("bench_grad_eval/" + tag, B=args.bench, T=args.bench_T, ms=dt * 1e3, clips_per_s=args.bench / dt,
                   tflops=76.5e9 * args.bench * (args.bench_T / 64000) / dt / 1e12)
        eb.set_option("conv_path", 0)
        eb.profile_begin()
        eb.grad(xb, yb)
        prof = eb.profile_end()
        prof.sort(key=lambda r: -r["total_ms"])
        report("bench_kernels", top=prof[:25])
    os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
    with open(args.out, "w") as f:
        json.dump(RESULTS, f, indent=1)


if __name__ == "__main__":
    main()
