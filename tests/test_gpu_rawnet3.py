"""GPU: RawNet3 on the CUDA path (through the C ABI) against the oracle and the reference-generated golden fixtures.

Conditioning note: RawNet3 takes log(|s| + 1e-6) of its sinc-filter outputs, so d logit / d waveform is
dominated by the handful of filter outputs that sit within ~1e-5 of a zero crossing, where the fp32 rounding of the 251-tap
dot product itself (1e-5 absolute) changes 1 / (|s| + 1e-6) by tens of percent.  Two correct fp32 implementations
therefore agree on the logits to ~1e-6 and on the waveform gradient only in direction (cosine ~0.99, signs ~99.5 %); the
reference's own multi-step attacks drift apart between 1 and 8 CPU threads (measured: cosine 0.83 after 3 PGDL2 steps).
The tests pin every segment tightly where that is meaningful - the forward stage by stage, the backward with the oracle
evaluated AT THE ENGINE'S OWN sinc outputs / fed with the engine's own intermediate gradients - and gate the end-to-end
quantities with the tolerances the conditioning allows.
"""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import attacks as oatk
from oracle import cases
from oracle import rawnet3 as orn

import helpers

pytestmark = pytest.mark.gpu

NAME = "rawnet3_t16000"
PATHS = pytest.mark.parametrize("conv_path", [0, 1], ids=["tcgen05", "simt"])


def _setup(name, dev, conv_path=0):
    from advb200 import engine

    case, x, y, holder, state, fwd = helpers.case_setup(name)
    holder = helpers.load_holder_state(holder, state, dev)
    eng = engine.engine_for(holder, x.shape[0], x.shape[1])
    eng.set_option("conv_path", conv_path)
    return case, x, y, holder, state, fwd, eng


def _stage(eng, name, B):
    t, _ = eng.debug_stage(name)  # (Bmax, rows per clip incl. zero border rows, 1, C)
    return t[:B, :, 0, :].cpu()


def _valid(t, pad):
    return t[:, pad:t.shape[1] - pad]


@PATHS
def test_forward_stage_by_stage(conv_path, cuda_device):
    case, x, y, holder, state, fwd, eng = _setup(NAME, cuda_device, conv_path)
    B = x.shape[0]
    taps = {}
    with torch.no_grad():
        want = fwd(x, state, taps)
    got = eng.forward(x.to(cuda_device)).cpu()
    assert (_stage(eng, "rn_pre", B)[:, :, 0] - taps["pre"][:, 0]).abs().max().item() < 2e-6
    filt = _stage(eng, "rn_filt", 1).view(256, 251)
    assert (filt - orn.sinc_filters(state)[:, 0]).abs().max().item() < 5e-6
    assert helpers.rel_err(_stage(eng, "rn_sinc_raw", B), taps["sinc_raw"].transpose(1, 2)) < 5e-6
    # log|s|: exact to fp32 everywhere except within ~1e-5 of a zero crossing of s (median error ~1e-6, L2 error 5e-4)
    s = _valid(_stage(eng, "rn_sinc", B), 2)
    assert (s - taps["sinc"].transpose(1, 2)).abs().median().item() < 1e-5
    assert helpers.rel_err(s, taps["sinc"].transpose(1, 2)) < 3e-3
    for stage, key, pad in (("rn_y1", "layer1.pre_pool", 2), ("rn_y2", "layer2.pre_pool", 3), ("rn_y3", "layer3.pre_pool", 4),
                            ("rn_x1", "x1", 3)):
        assert helpers.rel_err(_valid(_stage(eng, stage, B), pad), taps[key].transpose(1, 2)) < 2e-3, stage
        t = _stage(eng, stage, B)
        assert t[:, :pad].abs().max().item() == 0.0 and t[:, t.shape[1] - pad:].abs().max().item() == 0.0, \
            "zero border rows must stay zero"
    c4 = _stage(eng, "rn_cat4", B)
    assert helpers.rel_err(c4[:, :, 1024:2048], taps["x2"].transpose(1, 2)) < 1e-3
    assert helpers.rel_err(c4[:, :, 2048:], taps["x3"].transpose(1, 2)) < 1e-3
    assert helpers.rel_err(_stage(eng, "rn_layer4", B), taps["layer4"].transpose(1, 2)) < 1e-3
    assert helpers.rel_err(_stage(eng, "rn_pooled", B)[:, 0], taps["pooled"]) < 2e-4
    np.testing.assert_allclose(got.numpy(), want.numpy(), atol=1e-5)
    np.testing.assert_allclose(got.numpy(), helpers.load_golden(NAME)["logits"], atol=1e-5)


@PATHS
@pytest.mark.parametrize("name", [NAME, "rawnet3_t16000_margin"])
def test_backward_segments_pinned_at_engine_state(name, conv_path, cuda_device):
    """Tail backward with the oracle evaluated at the engine's own sinc outputs; transposed sinc convolution and
    preprocess VJP fed with the engine's own intermediate gradients: every segment tight, no conditioning excuse."""
    case, x, y, holder, state, fwd, eng = _setup(name, cuda_device, conv_path)
    B = x.shape[0]
    g, logits = eng.grad(x.to(cuda_device), y.to(cuda_device))
    S = _stage(eng, "rn_sinc_raw", B).transpose(1, 2).contiguous().requires_grad_(True)
    o = orn.tail(S, state)
    # tcgen05 accumulates in fp32 with truncation: over RawNet3's 1024..3072-long contractions of non-negative (post-ReLU)
    # activations that is a systematic ~4e-6 logit offset (measured); the fp32 FMA path agrees to 3e-7
    np.testing.assert_allclose(logits.cpu().numpy(), o.detach().numpy(), atol=1e-5 if conv_path == 0 else 3e-6)
    cost = torch.nn.CrossEntropyLoss()(torch.cat([-o, o], dim=1), y)
    (gS,) = torch.autograd.grad(cost, S)
    GS = _stage(eng, "rn_gs", B).transpose(1, 2).contiguous()
    # a ReLU / max-pool winner decided by a ~1e-7 margin may flip between two correct fp32 implementations
    assert helpers.trimmed_rel_err(GS, gS) < 1e-4
    assert helpers.cosine(GS, gS) > 0.99999
    assert (torch.sign(GS) == torch.sign(gS)).float().mean().item() > 0.9995
    gpre = F.conv_transpose1d(GS, orn.sinc_filters(state), stride=10)
    gpre = F.pad(gpre, (0, x.shape[1] - gpre.shape[-1]))[:, 0]
    gn = _stage(eng, "rn_gn", B)[:, :, 0]
    assert helpers.rel_err(gn, gpre) < 1e-5
    xc = x.clone().requires_grad_(True)
    (gx,) = torch.autograd.grad(orn.preprocess(xc, state), xc, grad_outputs=gn.unsqueeze(1))
    assert helpers.rel_err(g.cpu(), gx) < 1e-5


@PATHS
@pytest.mark.parametrize("name", [NAME, "rawnet3_t64000"])
def test_logits_and_gradient_against_reference_golden(name, conv_path, cuda_device):
    case, x, y, holder, state, fwd, eng = _setup(name, cuda_device, conv_path)
    B = x.shape[0]
    gold = helpers.load_golden(name)
    g, logits = eng.grad(x.to(cuda_device), y.to(cuda_device))
    np.testing.assert_allclose(logits.cpu().numpy(), gold["logits"], atol=1e-5)
    # End to end from the clean waveform, the last WELL-CONDITIONED gradient is the one w.r.t. the log-sinc features
    # (upstream of 1 / (|s| + 1e-6)): direction against the oracle (which test_oracle_golden pins to the reference).
    taps = {}
    xc = x.clone().requires_grad_(True)
    o = fwd(xc, state, taps)
    taps["sinc"].retain_grad()
    torch.nn.CrossEntropyLoss()(torch.cat([-o, o], dim=1), y).backward()
    gs = _valid(_stage(eng, "rn_gsinc", B), 2)
    gs_ref = taps["sinc"].grad.transpose(1, 2)
    assert helpers.cosine(gs, gs_ref) > 0.999
    assert helpers.trimmed_rel_err(gs, gs_ref) < 2e-2
    # ... and tightly where the filter output is well above the rounding noise of a 251-tap fp32 dot product (~1e-5 absolute:
    # at |s| >= 0.05, 96 % of the outputs, the 1 / (|s| + 1e-6) factor is known to 2e-4 relative).  What is left there are
    # the isolated patches behind a ReLU / max-pool winner that two correct fp32 forwards decide differently (measured:
    # untrimmed 1.8e-2 on the tcgen05 AND on the fp32 FMA path alike); the trimmed error drops them.
    s_abs = taps["sinc"].detach().transpose(1, 2).abs()
    well = s_abs >= 0.05
    assert well.float().mean().item() > 0.95
    t01, t1, t10 = (helpers.trimmed_rel_err(gs[well], gs_ref[well], drop=d) for d in (0.001, 0.01, 0.10))
    print(f"rn_gsinc on |s| >= 0.05 [{name}, conv_path {conv_path}]: trimmed relative error {t01:.2e} (0.1 %), {t1:.2e} (1 %), "
          f"{t10:.2e} (10 %)")
    assert t10 < 2e-3  # (the unmasked comparison above: 2e-2)
    assert (torch.sign(gs[well]) == torch.sign(gs_ref[well])).float().mean().item() > 0.995
    # The waveform gradient itself is dominated by the few filter outputs nearest to a zero crossing (|s| ~ 1e-6..1e-5,
    # below the 1e-5 rounding noise of the 251-tap dot product), and InstanceNorm's backward spreads them over every
    # sample: against the reference only its scale is comparable (measured cosine: 0.98 at T = 16 000, 0.1-0.25 at
    # T = 64 000, for the fp32 FMA path just as for the tensor-core path).  Its exactness is pinned segment by segment in
    # test_backward_segments_pinned_at_engine_state.
    ref = torch.from_numpy(gold["grad"])
    assert torch.isfinite(g).all()
    ratio = (g.cpu().abs().median() / ref.abs().median()).item()
    assert 0.5 < ratio < 2.0
    # The yardstick for the waveform gradient is the FLOAT64 run of the reference model (tools/rn_conditioning.py ->
    # tests/golden/rawnet3_fp64_grad_*.npz).  The reference's own float32 gradient agrees with it in 99.0 % / 98.6 % of the
    # signs and with cosine 0.906 / -0.085 (T = 16 000 / 64 000; 8 threads and 1 thread alike) - a handful of 1 / (|s| + 1e-6)
    # terms carry most of the norm.  The engine is held to the same standard: at least as close to float64 as the reference's
    # float32 runs are (signs decide FGSM / PGD; the norm-dominating terms decide nothing an L-inf attack uses).
    f64 = np.load(os.path.join(cases.GOLDEN_DIR, name.replace("rawnet3_", "rawnet3_fp64_grad_") + ".npz"))
    g64 = torch.from_numpy(f64["g64"])
    sign_ref = min(float(f64["ref32_t8_sign"]), float(f64["ref32_t1_sign"]))
    sign_eng = (torch.sign(g.cpu().double()) == torch.sign(g64)).double().mean().item()
    cos_eng, cos_ref = helpers.cosine(g.cpu(), g64), min(float(f64["ref32_t8_cos"]), float(f64["ref32_t1_cos"]))
    print(f"rawnet3 waveform gradient vs float64 [{name}, conv_path {conv_path}]: signs engine {sign_eng:.4f} / reference fp32 "
          f"{sign_ref:.4f}; cosine engine {cos_eng:.4f} / reference fp32 {cos_ref:.4f}")
    assert sign_eng > sign_ref - 0.01
    if cos_ref > 0.5:  # (T = 64 000: the reference's own cosine with float64 is -0.09 - nothing to hold the engine to)
        assert cos_eng > cos_ref - 0.1
    # logit-gradient mode (FAB, fab.py:90-105): the CE gradient is its per-clip multiple
    from advb200 import _lib

    gl, _ = eng.grad(x.to(cuda_device), None, what=_lib.GRAD_LOGIT)
    coef = 2 * (torch.sigmoid(2 * logits.cpu()) - y.view(-1, 1).float()) / x.shape[0]
    assert helpers.rel_err(gl.cpu() * coef, g.cpu()) < 1e-4


def test_tensor_core_path_matches_simt_path_and_schedules_agree(cuda_device):
    case, x, y, holder, state, fwd, eng = _setup(NAME, cuda_device, 0)
    xd, yd = x.to(cuda_device), y.to(cuda_device)
    g0, l0 = eng.grad(xd, yd)
    gs0 = _stage(eng, "rn_gsinc", x.shape[0])
    g0b, l0b = eng.grad(xd, yd)
    assert torch.equal(g0, g0b) and torch.equal(l0, l0b), "the engine must be deterministic run to run"
    # schedules: the persistent warp-specialised GEMM and the one-tile-per-CTA GEMM issue the same MMAs in the same order
    eng.set_option("conv_sched", 1)
    g2, l2 = eng.grad(xd, yd)
    eng.set_option("conv_sched", 0)
    assert torch.equal(l0, l2) and torch.equal(g0, g2)
    eng.set_option("conv_path", 1)
    g1, l1 = eng.grad(xd, yd)
    gs1 = _stage(eng, "rn_gsinc", x.shape[0])
    eng.set_option("conv_path", 0)
    np.testing.assert_allclose(l0.cpu().numpy(), l1.cpu().numpy(), atol=1e-5)
    assert helpers.cosine(gs0, gs1) > 0.999  # (the waveform gradient itself is ill-conditioned: see the golden test)


@pytest.mark.parametrize("attack", ["fgsm", "pgd", "pgdl2"])
def test_attacks_against_oracle_and_golden(attack, cuda_device):
    from advb200 import torchattacks as ta

    case, x, y, holder, state, fwd, eng = _setup(NAME, cuda_device)
    g = helpers.load_golden(NAME)
    p = cases.ATTACKS[attack]
    xd, yd = x.to(cuda_device), y.to(cuda_device)
    model_fn = lambda v: fwd(v, state)  # noqa: E731

    def run(steps):
        if attack == "fgsm":
            atk = ta.FGSM(holder, eps=p["eps"])
            atk.set_training_mode(True, False)
            return atk(xd, yd).cpu()
        if attack == "pgd":
            atk = ta.PGD(holder, eps=p["eps"], alpha=p["alpha"], steps=steps)
            atk.set_training_mode(True, False)
            return atk.forward(xd, yd, noise=helpers.reference_start(case, "pgd", x, p["eps"]).to(cuda_device)).cpu()
        atk = ta.PGDL2(holder, eps=p["eps"], alpha=p["alpha"], steps=steps)
        atk.set_training_mode(True, False)
        return atk.forward(xd, yd, delta=helpers.reference_start(case, "pgdl2", x, p["eps"]).to(cuda_device)).cpu()

    got = run(p.get("steps", 1))
    assert torch.equal(xd.cpu(), x), "inputs must not be mutated"
    assert got.min().item() >= 0.0 and got.max().item() <= 1.0
    d = got - x
    # the attack's own perturbation norm within 1e-5 of the reference (north-star tolerance)
    if attack == "pgdl2":
        np.testing.assert_allclose(d.norm(p=2, dim=1).numpy(), g["pgdl2_delta_l2"], rtol=1e-5, atol=1e-5)
    else:
        np.testing.assert_allclose(d.abs().amax(dim=1).numpy(), g[f"{attack}_delta_linf"], atol=1e-5)
        np.testing.assert_allclose(d.norm(p=2, dim=1).numpy(), g[f"{attack}_delta_l2"], rtol=1e-4, atol=1e-5)
    # element-wise: only the FIRST step is comparable (the reference's own iterates are chaotic beyond it)
    if attack == "fgsm":
        ref = torch.from_numpy(g["fgsm_adv"])
        assert (got != ref).float().mean().item() < 2e-2
    elif attack == "pgd":
        one = run(1)
        want = oatk.pgd(model_fn, x, y, p["eps"], p["alpha"], 1, noise=helpers.reference_start(case, "pgd", x, p["eps"]))
        assert (one != want).float().mean().item() < 2e-2
    else:
        # the L2-normalised direction is the ill-conditioned waveform gradient itself (see the golden test): check the
        # update rule (pgdl2.py:78-88) on the engine's own gradient at the reference's start point instead
        one = run(1)
        start = torch.clamp(x + helpers.reference_start(case, "pgdl2", x, p["eps"]), 0, 1)
        ge, _ = eng.grad(start.to(cuda_device), yd)
        ge = ge.cpu()
        gn = ge / (ge.view(ge.shape[0], -1).norm(p=2, dim=1).view(-1, 1) + 1e-10)
        adv = start + p["alpha"] * gn
        delta = adv - x
        dn = delta.view(delta.shape[0], -1).norm(p=2, dim=1).view(-1, 1)
        want = torch.clamp(x + delta * torch.min(p["eps"] / dn, torch.ones_like(dn)), 0, 1)
        assert (one - want).abs().max().item() < 2e-6
    # predicted labels of the attacked batch as in the reference, and at the reference's own adversarial batch the
    # engine must reproduce the reference's logits
    la = eng.forward(got.to(cuda_device)).cpu().numpy()
    assert np.array_equal(la > 0, g[f"{attack}_logits_adv"] > 0)
    ref = torch.from_numpy(g[f"{attack}_adv"])
    lr = eng.forward(ref.to(cuda_device)).cpu().numpy()
    np.testing.assert_allclose(lr, g[f"{attack}_logits_adv"], atol=2e-5)


def test_fab_against_golden(cuda_device):
    """FAB on RawNet3 is chaotic in the reference itself (its L-inf moves by 25 % between 1 and 8 CPU threads and a clip
    may need 4 or 8 steps): gate what is stable - the misclassified clip comes back untouched, every perturbation FAB
    returns flips the label and has the reference's order of magnitude, and the reference's adversarial batch gets the
    reference's logits."""
    from advb200 import torchattacks as ta

    name = "rawnet3_t16000_margin"
    case, x, y, holder, state, fwd, eng = _setup(name, cuda_device)
    g = helpers.load_golden(name)
    p = cases.ATTACKS["fab"]
    xd, yd = x.to(cuda_device), y.to(cuda_device)
    atk = ta.FAB(holder, norm="Linf", eps=p["eps"], steps=p["steps"], eta=p["eta"], alpha_max=p["alpha_max"],
                 beta=p["beta"], n_classes=2)
    atk.set_training_mode(True, False)
    got = atk(xd, yd).cpu()
    assert torch.equal(xd.cpu(), x)
    assert got.min().item() >= 0.0 and got.max().item() <= 1.0
    assert torch.equal(got[1], x[1]), "a clip that is misclassified from the start must come back untouched"
    linf = (got - x).abs().amax(dim=1).numpy()
    la = eng.forward(got.to(cuda_device)).cpu().numpy().ravel()
    found = linf > 0
    assert found[[0, 2]].any(), "FAB found no adversarial example at all"
    for i in np.nonzero(found)[0]:
        assert (la[i] > 0) != bool(y[i].item() == 1), "a returned perturbation must flip the label"
        assert linf[i] < 3 * g["fab_delta_linf"][i] and g["fab_delta_linf"][i] < 3 * linf[i]
    ref = torch.from_numpy(g["fab_adv"])
    lr = eng.forward(ref.to(cuda_device)).cpu().numpy()
    np.testing.assert_allclose(lr, g["fab_logits_adv"], atol=2e-5)
    assert np.array_equal(lr > 0, g["fab_logits_adv"] > 0)


def test_native_clip_length_and_ragged_batch(cuda_device):
    """64 600 samples (the reference's native clip length, 6 435 sinc frames -> 1 287 -> 429) and a batch smaller than the
    handle's max_batch: logits against the oracle."""
    from advb200 import engine
    from oracle import synth

    case, x0, y0, holder, state, fwd = helpers.case_setup(NAME)
    holder = helpers.load_holder_state(holder, state, cuda_device)
    x, y = synth.clips(31, 3, 64600)
    eng = engine.engine_for(holder, 3, 64600)
    with torch.no_grad():
        want = fwd(x[:2], state)
    got = eng.forward(x[:2].to(cuda_device)).cpu()
    np.testing.assert_allclose(got.numpy(), want.numpy(), atol=1e-5)
    g, _ = eng.grad(x.to(cuda_device), y.to(cuda_device))
    assert torch.isfinite(g).all()


def test_errors_are_loud(cuda_device):
    from advb200 import engine

    case, x, y, holder, state, fwd, eng = _setup(NAME, cuda_device)
    with pytest.raises(RuntimeError):
        eng.frontend_fwd(x.to(cuda_device))  # RawNet3 has no spectral frontend
    with pytest.raises((RuntimeError, ValueError)):
        eng.forward(x[:, :8000].contiguous().to(cuda_device))
