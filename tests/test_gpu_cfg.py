"""GPU: parity AT THE BENCHMARKED CONFIGURATIONS (BASELINE.json configs[0..3]) against fixtures produced by the unmodified
reference on CPU (oracle/make_golden_cfg.py), plus the schedule variants of the native loop (CUDA graph, fused update,
fused min-max) against each other.

Success criterion of the reference: evaluate_models_on_adversarial_attacks.py:236-265 -- labels (sigmoid(o) + .5).int() of the
attacked batch, accuracy.  Gates (north star): labels bit-exact, attack-success-rate within +-0.1 %, perturbation norms
within 1e-5; the gradient-sign mismatch count is REPORTED for every case and bounded for the spectral models.
"""
import json
import os

import numpy as np
import pytest
import torch

from oracle import cases, synth
from oracle.make_golden_cfg import BIAS_KEY, CFG_CASES, T

import helpers

pytestmark = pytest.mark.gpu

REPORT = os.path.join(cases.ROOT, "gpurun_out", "cfg_parity.json")


def cfg_setup(name, dev):
    from advb200 import engine

    case, gold = CFG_CASES[name], helpers.load_golden(name)
    x, y = synth.clips(case["cfg_id"], case["B"], T)
    holder, state = cases.build_state(case["model"], case["frontend"])
    state[BIAS_KEY[case["model"]]] = torch.from_numpy(gold["bias"])
    assert synth.state_digest(state) == str(gold["digest"]), "seeded weights differ from the ones the reference ran"
    assert np.array_equal(y.numpy(), gold["y"])
    holder = helpers.load_holder_state(holder, state, dev)
    return case, gold, x, y, holder, engine.engine_for(holder, case["B"], T)


def native_attack(case, holder, x, y, dev):
    """The native attack on the reference's own random start (drawn from torch's CPU generator under the seed the golden
    generator used: pgd.py:56 / pgdl2.py:57-61)."""
    from advb200 import torchattacks as ta

    p, kind = case["params"], case["attack"]
    xd, yd = x.to(dev), y.to(dev)
    if kind == "fgsm":
        atk = ta.FGSM(holder, eps=p["eps"])
        atk.set_training_mode(model_training=True, batchnorm_training=False)
        return atk(xd, yd)
    start = helpers.reference_start(case, kind, x, p["eps"]).to(dev)
    if kind == "pgd":
        atk = ta.PGD(holder, eps=p["eps"], alpha=p["alpha"], steps=p["steps"], random_start=True)
        atk.set_training_mode(model_training=True, batchnorm_training=False)
        return atk.forward(xd, yd, noise=start)
    atk = ta.PGDL2(holder, eps=p["eps"], alpha=p["alpha"], steps=p["steps"], random_start=True)
    atk.set_training_mode(model_training=True, batchnorm_training=False)
    return atk.forward(xd, yd, delta=start)


def _labels(logits):
    return (torch.sigmoid(logits.flatten()) + .5).int()  # evaluate_models_on_adversarial_attacks.py:236-238


@pytest.mark.parametrize("name", list(CFG_CASES))
def test_attack_matches_reference_at_benchmarked_config(name, cuda_device, record_property):
    case, gold, x, y, holder, eng = cfg_setup(name, cuda_device)
    clean = eng.forward(x.to(cuda_device)).cpu()
    np.testing.assert_allclose(clean.numpy(), gold["logits_clean"], atol=2e-5)
    assert np.array_equal(_labels(clean).numpy(), gold["pred_clean"])

    adv = native_attack(case, holder, x, y, cuda_device)
    la = eng.forward(adv).cpu()
    adv = adv.cpu()
    d = adv - x
    eps = case["params"]["eps"]
    rawnet = case["model"] == "rawnet3"
    # perturbation norms: the attack's own norm within 1e-5 (north star)
    if case["attack"] == "pgdl2":
        np.testing.assert_allclose(d.norm(p=2, dim=1).numpy(), gold["delta_l2"], rtol=1e-5, atol=1e-5)
    else:
        np.testing.assert_allclose(d.abs().amax(dim=1).numpy(), gold["delta_linf"], atol=1e-5)
        np.testing.assert_allclose(d.norm(p=2, dim=1).numpy(), gold["delta_l2"], rtol=1e-4)

    # predicted labels of the attacked batch and the attack success rate
    pred = _labels(la).numpy()
    y_np = y.numpy()
    asr_ref = float((gold["pred_adv"] != y_np).mean())
    asr = float((pred != y_np).mean())
    label_mismatch = int((pred != gold["pred_adv"]).sum())
    n = x.numel()
    sign_ref = np.unpackbits(gold["sign_bits"])[:n].reshape(x.shape).astype(bool)
    sign_mismatch = int(((adv > x).numpy() != sign_ref).sum())
    dlogit = float(np.abs(la.numpy() - gold["logits_adv"]).max())
    row = {"case": name, "clips": int(x.shape[0]), "label_mismatch": label_mismatch, "asr": asr, "asr_reference": asr_ref,
           "sign_mismatch": sign_mismatch, "sign_mismatch_frac": sign_mismatch / n, "max_abs_dlogit_adv": dlogit,
           "min_abs_logit_adv_reference": float(np.abs(gold["logits_adv"]).min())}
    for k, v in row.items():
        record_property(k, v)
    print("cfg parity:", json.dumps(row))
    try:  # collected by the round's profile notes (gpurun_out is scratch; never read back by any test)
        os.makedirs(os.path.dirname(REPORT), exist_ok=True)
        rows = json.load(open(REPORT)) if os.path.exists(REPORT) else {}
        rows[name] = row
        json.dump(rows, open(REPORT, "w"), indent=1)
    except OSError:
        pass
    # Sample-by-sample agreement is only defined for ONE gradient evaluation: PGD as configured is degenerate (alpha > 2 eps:
    # every step lands on x +- eps), so one fp32 tie moves a sample by 2 eps, the next gradient is evaluated at a different
    # point, more near-zero samples flip ...  The reference's OWN 8-thread and 1-thread runs differ in 1.1 % of the samples after
    # 5 steps, 3.5 % after 10, 7.2 % after 20, 9.9 % after 40, with adversarial logits 6e-5 apart (tools/pgd_divergence.py ->
    # tests/golden/pgd_divergence_reference.json).  So: the FIRST step is gated tightly, sample by sample; the full attack
    # is gated by what it is for - the predicted labels and the attack success rate - on every clip whose reference logit
    # is farther from the decision boundary than the run-to-run logit noise of that model (LCNN: all 128 clips qualify).
    if case["attack"] == "pgd":
        one = dict(case, params=dict(case["params"], steps=1))
        adv1 = native_attack(one, holder, x, y, cuda_device).cpu()
        s1_ref = np.unpackbits(gold["step1_sign_bits"])[:n].reshape(x.shape).astype(bool)
        s1 = int(((adv1 > x).numpy() != s1_ref).sum())
        record_property("step1_sign_mismatch", s1)
        print("cfg parity: first PGD step sign mismatch", s1, "of", n, "=", s1 / n)
        assert s1 / n < 1e-3, (s1, n)
    elif case["attack"] == "fgsm":
        assert sign_mismatch / n < 1e-3, row
    # (band, bound on the adversarial-logit difference): measured native-vs-reference differences are 1.7e-4 (LCNN), 1.1e-3 - 3.1e-3
    # (SpecRNet+MFCC, batch-coupled dB floor; fp32 SIMT convolutions / 3xTF32 tensor-core convolutions) and 1.9e-3 (RawNet3,
    # ill-conditioned log|sinc| gradient).  SpecRNet's band is set from the reference's OWN noise: its 8-thread and
    # 1-thread PGD-40 runs end 2.9e-3 apart in the adversarial logit on 8 clips (15.7 % of the samples differ; 9e-8 after one step:
    # tools/pgd_divergence.py specrnet -> tests/golden/pgd_divergence_reference_specrnet.json): band = that, bound = twice that.
    band, dmax = {"lcnn": (0.0, 5e-4), "specrnet": (3e-3, 6e-3), "rawnet3": (2.2e-3, 3e-3)}[case["model"]]
    decided = np.abs(gold["logits_adv"].ravel()) > band
    print("cfg parity: clips outside the +-%g logit band: %d of %d" % (band, int(decided.sum()), decided.size))
    assert decided.sum() >= 0.7 * decided.size
    assert np.array_equal(pred[decided], gold["pred_adv"][decided]), row
    assert abs(float((pred[decided] != y_np[decided]).mean()) - float((gold["pred_adv"][decided] != y_np[decided]).mean())) <= 1e-3
    if case["model"] == "lcnn":
        assert label_mismatch == 0 and abs(asr - asr_ref) <= 1e-3, row
    assert dlogit < dmax, row


def test_clips_are_independent_at_the_headline_size(cuda_device):
    """Size-independent property at BASELINE.json configs[1] (B = 128, T = 64 000): clips shard with no data-path coupling
    (SURVEY.md §8e), so the attack of the whole batch is, bit for bit, the attacks of its two halves - whatever the persistent
    kernels' tile schedules (148 CTAs over 128 or 64 clips) and the CUDA-graph replay do.  (LFCC on these clips never reaches the
    batch-wide dB floor, the one coupling the reference itself has: F5.)  PGD-6 and the gradient of one evaluation."""
    from advb200 import torchattacks as ta

    name = "cfg2_lcnn_pgd40_b128"
    case, gold, x, y, holder, eng = cfg_setup(name, cuda_device)
    p = case["params"]
    xd, yd = x.to(cuda_device), y.to(cuda_device)
    noise = helpers.reference_start(case, "pgd", x, p["eps"]).to(cuda_device)
    atk = ta.PGD(holder, eps=p["eps"], alpha=p["alpha"], steps=6, random_start=True)
    atk.set_training_mode(model_training=True, batchnorm_training=False)
    whole = atk.forward(xd, yd, noise=noise)
    halves = torch.cat([atk.forward(xd[:64], yd[:64], noise=noise[:64]), atk.forward(xd[64:], yd[64:], noise=noise[64:])])
    assert torch.equal(whole, halves)
    # the CE mean is over the batch the call sees: compare per-clip gradients at the same normalisation
    g_all, l_all = eng.grad(xd, yd)
    g_a, l_a = eng.grad(xd[:64], yd[:64])
    assert torch.equal(l_all[:64], l_a)
    assert (g_all[:64] * 2.0 - g_a).abs().max().item() <= 1e-12  # 1 / 128 vs 1 / 64: a power of two, exact


def test_pgd_schedule_variants_are_bit_identical(cuda_device):
    """CUDA-graph replay vs host loop, update fused into the frontend backward vs its own kernel, even and odd step counts
    (the fused loop ping-pongs two iterate buffers): every variant must return the same bits."""
    from advb200 import torchattacks as ta

    name = "lcnn_lfcc_t16000"
    case, x, y, holder, state, fwd = helpers.case_setup(name)
    holder = helpers.load_holder_state(holder, state, cuda_device)
    from advb200 import engine

    eng = engine.engine_for(holder, x.shape[0], x.shape[1])
    xd, yd = x.to(cuda_device), y.to(cuda_device)
    noise = helpers.reference_start(case, "pgd", x, 0.001).to(cuda_device)
    try:
        for steps in (4, 5, 1):
            outs = {}
            for graph in (0, 1):
                for fuse in (0, 1):
                    eng.set_option("graph", graph)
                    eng.set_option("fuse_update", fuse)
                    atk = ta.PGD(holder, eps=0.001, alpha=2 / 255, steps=steps)
                    a = atk.forward(xd, yd, noise=noise)
                    b = atk.forward(xd, yd, noise=noise)  # second call: cached graph exec
                    assert torch.equal(a, b)
                    outs[(graph, fuse)] = a
            ref = outs[(0, 0)]
            for k, v in outs.items():
                assert torch.equal(v, ref), (steps, k)
        # FGSM fused vs unfused, PGDL2 graph vs host loop
        for fuse in (0, 1):
            eng.set_option("fuse_update", fuse)
            outs[("fgsm", fuse)] = ta.FGSM(holder, eps=0.005)(xd, yd)
        assert torch.equal(outs[("fgsm", 0)], outs[("fgsm", 1)])
        delta = helpers.reference_start(case, "pgdl2", x, 0.1).to(cuda_device)
        for graph in (0, 1):
            eng.set_option("graph", graph)
            outs[("l2", graph)] = ta.PGDL2(holder, eps=0.1, alpha=0.2, steps=4).forward(xd, yd, delta=delta)
        assert torch.equal(outs[("l2", 0)], outs[("l2", 1)])
    finally:
        eng.set_option("graph", 1)
        eng.set_option("fuse_update", 1)


def test_graph_replay_sees_weight_updates(cuda_device):
    """Adversarial training mutates the weights between attack calls (src/trainer.py:328-329): the cached graph reads the
    repacked weights, and the version-stamp cache must notice an in-place optimiser-style update."""
    from advb200 import engine
    from advb200 import torchattacks as ta

    case, x, y, holder, state, fwd = helpers.case_setup("lcnn_lfcc_t16000")
    holder = helpers.load_holder_state(holder, state, cuda_device)
    eng = engine.engine_for(holder, x.shape[0], x.shape[1])
    xd, yd = x.to(cuda_device), y.to(cuda_device)
    noise = helpers.reference_start(case, "pgd", x, 0.001).to(cuda_device)
    atk = ta.PGD(holder, eps=0.001, alpha=2 / 255, steps=4)
    a = atk.forward(xd, yd, noise=noise)
    l0 = eng.launches
    a2 = atk.forward(xd, yd, noise=noise)
    per_call_cached = eng.launches - l0
    with torch.no_grad():
        for p in holder.parameters():
            p.mul_(-1.0 if p.dim() == 4 and p.shape[1] == 1 else 1.0)  # flip the first convolution: the gradient changes
    l0 = eng.launches
    b = atk.forward(xd, yd, noise=noise)
    per_call_repacked = eng.launches - l0
    assert torch.equal(a, a2)
    assert not torch.equal(a, b)
    assert per_call_repacked > per_call_cached, "a weight change must trigger the repack kernels"
    eng.set_option("graph", 0)
    try:
        c = atk.forward(xd, yd, noise=noise)
    finally:
        eng.set_option("graph", 1)
    assert torch.equal(b, c)


def test_fused_minmax_matches_the_reference_three_line_sequence(cuda_device):
    """evaluate_models_on_adversarial_attacks.py:219-221 with the reference's own to_minmax / revert_minmax arithmetic executed
    by torch (src/aa/utils.py:4-14) around the native attack, against the single native call."""
    from advb200 import aa
    from advb200 import torchattacks as ta

    case, x, y, holder, state, fwd = helpers.case_setup("lcnn_lfcc_t16000")
    holder = helpers.load_holder_state(holder, state, cuda_device)
    g = torch.Generator("cpu").manual_seed(77)
    raw = (0.1 * torch.randn(x.shape, generator=g) - 0.03).to(cuda_device)
    yd = y.to(cuda_device)
    mn, mx = raw.min(dim=1, keepdim=True)[0], raw.max(dim=1, keepdim=True)[0]
    x01 = (raw - mn) / (mx - mn)
    noise = helpers.reference_start(case, "pgd", x, 0.001).to(cuda_device)
    for make, kw in ((lambda: ta.FGSM(holder, eps=0.005), {}),
                     (lambda: ta.PGD(holder, eps=0.001, alpha=2 / 255, steps=3), {"noise": noise})):
        atk = make()
        want = atk.forward(x01, yd, **kw) * (mx - mn) + mn
        atk._fused_minmax = True
        got = atk.forward(raw, yd, **kw)
        assert torch.equal(got, want)
    # constant clip: max - min = 0 -> NaN row, exactly like the reference (src/aa/utils.py:8-9); other rows unaffected
    raw2 = raw.clone()
    raw2[0] = 0.25
    x01b, mnb, mxb = aa.to_minmax(raw2)
    assert torch.isnan(x01b[0]).all() and torch.equal(x01b[1], x01[1])
    back = aa.revert_minmax(x01b, mnb, mxb)
    assert torch.isnan(back[0]).all() and torch.allclose(back[1], raw2[1], atol=1e-6)


@pytest.mark.parametrize("kind", ["fgsm", "pgd", "pgdl2", "cw"])
def test_targeted_modes_against_oracle(kind, cuda_device):
    """attack.py:60-108: set_mode_targeted_by_function on the native classes against the oracle's targeted variants (which
    tests/test_oracle_golden.py pins to the reference's own classes)."""
    from advb200 import torchattacks as ta

    name = "lcnn_lfcc_t16000_margin"
    case, x, y, holder, state, fwd = helpers.case_setup(name)
    holder = helpers.load_holder_state(holder, state, cuda_device)
    xd, yd = x.to(cuda_device), y.to(cuda_device)
    fn = lambda v: fwd(v, state)  # noqa: E731
    want = helpers.oracle_targeted(kind, fn, x, y, 1 - y, case)
    if kind == "fgsm":
        atk, kw = ta.FGSM(holder, eps=0.005), {}
    elif kind == "pgd":
        atk, kw = ta.PGD(holder, eps=0.001, alpha=2 / 255, steps=3), {"noise": helpers.reference_start(case, "pgd", x, 0.001)}
    elif kind == "pgdl2":
        atk, kw = ta.PGDL2(holder, eps=0.1, alpha=0.2, steps=3), {"delta": helpers.reference_start(case, "pgdl2", x, 0.1)}
    else:
        atk, kw = ta.CW(holder, c=1e4, kappa=0.0, steps=5, lr=5e-4), {}
    atk.set_mode_targeted_by_function(lambda images, labels: 1 - labels)
    got = atk.forward(xd, yd, **{k: v.to(cuda_device) for k, v in kw.items()}).cpu()
    if kind in ("fgsm", "pgd"):
        assert (got != want).float().mean().item() < 2e-3
    elif kind == "pgdl2":
        np.testing.assert_allclose((got - x).norm(p=2, dim=1).numpy(), (want - x).norm(p=2, dim=1).numpy(), rtol=1e-5)
        assert (got - want).abs().max().item() < 2e-4  # (the reference's own run-to-run element noise on PGDL2 is 2.4e-4)
    else:  # sign-like Adam steps amplify gradient ties (see tests/test_oracle_golden.py); norms and flips are stable
        np.testing.assert_allclose((got - x).norm(p=2, dim=1).numpy(), (want - x).norm(p=2, dim=1).numpy(), rtol=2e-2, atol=1e-6)
        assert ((got - want).abs() > 1e-5).float().mean().item() < 0.05
    # untargeted result differs (same ascent direction for t = 1 - y only up to the factor of the f term / sigma)
    with pytest.raises(ValueError):
        ta.FAB(holder, n_classes=2).set_mode_targeted_by_function()


def test_cw_strong_elementwise_and_mask_path(cuda_device):
    """CW with the classification term dominating rounding noise (c = 1e4, lr = 5e-4): three clips flip at steps 4-5 (the
    best-adversarial mask path, cw.py:94-101) and the batch-wide early stop fires at step 8 (cw.py:107-110).  Element-wise
    against the reference golden."""
    from advb200 import torchattacks as ta

    name = "lcnn_lfcc_t16000_margin"
    case, x, y, holder, state, fwd = helpers.case_setup(name)
    holder = helpers.load_holder_state(holder, state, cuda_device)
    from advb200 import engine

    eng = engine.engine_for(holder, x.shape[0], x.shape[1])
    g = helpers.load_golden(name)
    p = cases.ATTACKS["cw_strong"]
    atk = ta.CW(holder, c=p["c"], kappa=p["kappa"], steps=p["steps"], lr=p["lr"])
    l0 = eng.launches
    got = atk(x.to(cuda_device), y.to(cuda_device)).cpu()
    launches = eng.launches - l0
    ref = torch.from_numpy(g["cw_strong_adv"])
    d = (got - ref).abs()
    print("cw_strong: max |d|", d.max().item(), "frac > 1e-5:", (d > 1e-5).float().mean().item(), "launches", launches)
    np.testing.assert_allclose((got - x).norm(p=2, dim=1).numpy(), g["cw_strong_delta_l2"], rtol=1e-4, atol=1e-6)
    assert (d > 1e-5).float().mean().item() < 2e-3 and d.max().item() < 1e-3
    assert torch.equal(got[2], x[2]) or (got[2] - x[2]).abs().max().item() < 1e-7  # misclassified from the start: step-0 snapshot
    la = eng.forward(got.to(cuda_device)).cpu().numpy()
    assert np.array_equal(la > 0, g["cw_strong_logits_adv"] > 0)
    np.testing.assert_allclose(la, g["cw_strong_logits_adv"], atol=2e-5)
    # batch-wide early stop (cw.py:107-110): with 20 steps the cost is checked every 2 steps and first increases at step 8,
    # so 9 of the 20 iterations run; with 8 steps it is checked every step and stops at step 6 (7 iterations)
    short = ta.CW(holder, c=p["c"], kappa=p["kappa"], steps=8, lr=p["lr"])
    l0 = eng.launches
    short(x.to(cuda_device), y.to(cuda_device))
    per_iter = (eng.launches - l0) / 7.0
    assert 8.4 * per_iter < launches < 10.3 * per_iter, (launches, per_iter)


def test_frontend_singletons_run_standalone(cuda_device):
    """LFCC_FN(x) / MFCC_FN(x) (src/frontends.py:13-32) are part of the plugin surface: the holders, and torchaudio's own
    transforms, run through a frontend-only engine handle."""
    from advb200 import engine, frontends
    from oracle import frontend as ofe

    g = torch.Generator("cpu").manual_seed(3)
    x = torch.rand(3, 16000, generator=g)
    for fe in (frontends.LFCC().to(cuda_device), frontends.MFCC().to(cuda_device)):
        got = fe(x.to(cuda_device)).cpu()
        fb, dct, win = [t.cpu() for t in fe.tables()]
        want = ofe.cepstral_frontend(x, fb, dct, win)
        assert got.shape == want.shape == (3, 80, 101)
        assert (got - want).abs().max().item() < 1e-3
        assert fe(x[0].to(cuda_device)).shape == (80, 101)
    with pytest.raises(RuntimeError, match="frontend only"):
        engine.engine_for(fe, 3, 16000).forward(x.to(cuda_device))


def test_fab_random_restarts(cuda_device):
    """fab.py:507-526 with n_restarts = 2: the second run starts from the reference's random point (same CPU RNG draw) and only
    attacks the clips the first run left robust; results stay inside the eps ball and never lose an adversarial clip."""
    from advb200 import torchattacks as ta

    name = "lcnn_lfcc_t16000_margin"
    case, x, y, holder, state, fwd = helpers.case_setup(name)
    holder = helpers.load_holder_state(holder, state, cuda_device)
    p = cases.ATTACKS["fab"]
    xd, yd = x.to(cuda_device), y.to(cuda_device)
    one = ta.FAB(holder, norm="Linf", eps=p["eps"], steps=p["steps"], eta=p["eta"], n_restarts=1, n_classes=2)(xd, yd)
    two = ta.FAB(holder, norm="Linf", eps=p["eps"], steps=p["steps"], eta=p["eta"], n_restarts=2, n_classes=2)(xd, yd)
    assert (two - xd).abs().max().item() <= p["eps"] + 1e-6
    moved1 = (one != xd).any(dim=1)
    moved2 = (two != xd).any(dim=1)
    assert (moved2 | ~moved1).all(), "a restart may add adversarial clips, never drop one"
    assert torch.equal(two[moved1], one[moved1]), "clips already fooled keep their first adversarial example (fab.py:523-524)"


def test_calls_on_different_streams_are_ordered(cuda_device):
    """A handle's activations and scratch belong to the handle: a call enqueued on another CUDA stream must wait for the
    previous call (completion event in the C ABI), so interleaving streams gives the serial results."""
    from advb200 import engine
    from advb200 import torchattacks as ta

    case, x, y, holder, state, fwd = helpers.case_setup("lcnn_lfcc_t16000")
    holder = helpers.load_holder_state(holder, state, cuda_device)
    eng = engine.engine_for(holder, x.shape[0], x.shape[1])
    xd, yd = x.to(cuda_device), y.to(cuda_device)
    noise = helpers.reference_start(case, "pgd", x, 0.001).to(cuda_device)
    atk = ta.PGD(holder, eps=0.001, alpha=2 / 255, steps=6)
    want_adv = atk.forward(xd, yd, noise=noise)
    want_logits = eng.forward(want_adv)
    torch.cuda.synchronize()
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    for _ in range(3):
        with torch.cuda.stream(s1):
            adv = atk.forward(xd, yd, noise=noise)
        ready = torch.cuda.Event()
        ready.record(s1)
        with torch.cuda.stream(s2):
            s2.wait_event(ready)           # the caller's own dependency on `adv` ...
            logits = eng.forward(adv)      # ... and the engine's: its buffers are still in use by s1's call until it completes
            g, _ = eng.grad(xd, yd)
        with torch.cuda.stream(s1):
            adv2 = atk.forward(xd, yd, noise=noise)  # back on s1 while s2's calls may still run
        torch.cuda.synchronize()
        assert torch.equal(adv, want_adv) and torch.equal(adv2, want_adv) and torch.equal(logits, want_logits)


def test_mel_spec_frontend_matches_reference(cuda_device):
    """get_frontend(["mel_spec"]) = prepare_mel_scale_vector (src/frontends.py:53-79) on the GPU (fe_melspec_kernel)
    against the fixture the unmodified reference produced and against the oracle; odd frame count and a 1-D clip included."""
    from advb200 import frontends
    from oracle import frontend as ofe
    from oracle import synth

    g = helpers.load_golden("mel_spec")
    fn = frontends.get_frontend(["mel_spec"])
    for tag, cfg_id, B, T in (("t16000", 31, 2, 16000), ("t16150", 32, 1, 16150)):
        x, _ = synth.clips(cfg_id, B, T)
        got = fn(x.to(cuda_device)).cpu()
        want = torch.from_numpy(g[f"{tag}_out"])
        assert got.shape == want.shape
        zg, zw = torch.polar(got[:, 0], got[:, 1]), torch.polar(want[:, 0], want[:, 1])
        scale = zw.abs().max().item()
        assert (zg - zw).abs().max().item() < 2e-5 * scale, tag            # the complex mel spectrum, element-wise
        np.testing.assert_allclose(got[:, 0].numpy(), want[:, 0].numpy(), atol=2e-5 * scale)   # channel 0: abs
        big = want[:, 0] > 1e-2 * scale                                      # channel 1: angle, where it is well conditioned
        dphi = torch.remainder(got[:, 1] - want[:, 1] + np.pi, 2 * np.pi) - np.pi
        assert dphi[big].abs().max().item() < 2e-3
        zo = ofe.mel_spec(x, torch.from_numpy(g["fb"]))
        assert (zg - torch.polar(zo[:, 0], zo[:, 1])).abs().max().item() < 2e-5 * scale
    mag, ang = frontends.prepare_stft_features(x.to(cuda_device))
    assert torch.equal(mag.cpu(), got[:, 0]) and torch.equal(ang.cpu(), got[:, 1])
    one = fn(x[0].to(cuda_device))
    assert one.shape == (2, 80, 101) and torch.equal(one.cpu(), got[0])
    with pytest.raises(NotImplementedError):
        fn(x.to(cuda_device).requires_grad_(True))
