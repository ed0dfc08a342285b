"""CPU: pin the oracle restatement against fixtures produced by the reference itself (oracle/make_golden.py)."""
import numpy as np
import pytest
import torch

from oracle import attacks as oatk
from oracle import cases, synth

import helpers

CASES = list(cases.CASES)


@pytest.mark.parametrize("name", CASES)
def test_oracle_forward_and_gradient_match_reference(name):
    case, x, y, holder, state, fwd = helpers.case_setup(name)
    g = helpers.load_golden(name)
    assert synth.state_digest(state) == str(g["digest"]), "seeded weights differ from the ones the golden used"
    assert abs(x.double().sum().item() - float(g["x_sum"])) < 1e-6
    taps = {}
    with torch.no_grad():
        o = fwd(x, state, taps)
    np.testing.assert_allclose(o.numpy(), g["logits"], atol=2e-6)
    if "frontend" in g.files:
        np.testing.assert_allclose(taps["frontend"].squeeze(1).numpy(), g["frontend"], atol=2e-3)
    _, grad = oatk.loss_and_grad(lambda v: fwd(v, state), x, y)
    assert helpers.rel_err(grad, torch.from_numpy(g["grad"])) < 1e-4


@pytest.mark.parametrize("name", CASES[:2] + ["rawnet3_t16000"])
@pytest.mark.parametrize("attack", ["fgsm", "pgd", "pgdl2"])
def test_oracle_attacks_match_reference(name, attack):
    case, x, y, holder, state, fwd = helpers.case_setup(name)
    g = helpers.load_golden(name)
    xa = helpers.oracle_attack(name, attack, x, y, state, fwd, case)
    ref = torch.from_numpy(g[f"{attack}_adv"])
    # perturbation norms within 1e-5 (north-star tolerance); elements: identical up to rare gradient-sign ties
    # (the 1e-5 gate applies to the attack's own norm: L-inf for FGSM/PGD, L2 for PGDL2 — SURVEY.md §4)
    if attack != "pgdl2":
        np.testing.assert_allclose((xa - x).abs().amax(dim=1).numpy(), g[f"{attack}_delta_linf"], atol=1e-5)
    np.testing.assert_allclose((xa - x).norm(p=2, dim=1).numpy(), g[f"{attack}_delta_l2"], rtol=1e-5, atol=1e-5)
    if attack == "pgdl2":
        # With the dB floor active (silence case) the loss is discontinuous in x (clamp membership + arg-max
        # routing): ulp-level differences flip borderline elements and iterates drift apart, so
        # element-wise comparison is only meaningful on the smooth case; norms and labels are checked on both.
        # RawNet3: log(|sinc| + 1e-6) makes the gradient ill-conditioned near the zero crossings of the filter outputs;
        # the reference run with 1 vs 8 threads drifts to cosine 0.83 between its own 3-step PGDL2 perturbations
        # (measured), so only the first step is comparable element-wise (tested on the GPU per step).
        if not case["silence"] and case["model"] != "rawnet3":
            assert (xa - ref).abs().max().item() < 5e-6
    else:
        assert (xa != ref).float().mean().item() < 2e-3
    with torch.no_grad():
        la = fwd(xa, state)
    assert np.array_equal((la.numpy() > 0), (g[f"{attack}_logits_adv"] > 0)), "label flips differ"


@pytest.mark.parametrize("name,attack", [("lcnn_lfcc_t16000_margin", "fab"), ("lcnn_lfcc_t16000_margin", "cw"),
                                         ("lcnn_lfcc_t16000_margin", "cw_strong"), ("rawnet3_t16000_margin", "fab"),
                                         ("lcnn_lfcc_t16000_margin_l2", "fab_l2")])
def test_oracle_fab_cw_match_reference(name, attack):
    case, x, y, holder, state, fwd = helpers.case_setup(name)
    g = helpers.load_golden(name)
    xa = helpers.oracle_attack(name, attack, x, y, state, fwd, case)
    ref = torch.from_numpy(g[f"{attack}_adv"])
    # FAB's own norm is L-inf, CW's is L2; the clip labelled 0 is misclassified from the start and must
    # come back untouched by FAB (fab.py:506-513)
    if case["model"] == "rawnet3":
        # ill-conditioned gradient (see the PGDL2 note above): the reference's own FAB L-inf moves by 25 % between 1 and
        # 8 threads (measured); gate the order of magnitude, the untouched clips and the label flips only
        got, want = (xa - x).abs().amax(dim=1).numpy(), g[f"{attack}_delta_linf"]
        assert np.array_equal(got == 0, want == 0)
        assert np.all(got <= 2 * want + 1e-12) and np.all(want <= 2 * got + 1e-12)
    else:
        np.testing.assert_allclose((xa - x).abs().amax(dim=1).numpy(), g[f"{attack}_delta_linf"], rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose((xa - x).norm(p=2, dim=1).numpy(), g[f"{attack}_delta_l2"], rtol=1e-4, atol=1e-5)
        if attack == "fab_l2":
            # the L2 step is proportional to the gradient itself (not to its sign): over 8 boundary-hugging steps the direction
            # of two fp32 implementations drifts by ~1e-3 relative while the norm FAB minimises agrees to 5e-6 (measured)
            assert helpers.rel_err(xa - x, ref - x) < 5e-3
        else:
            assert (xa - ref).abs().max().item() < 1e-5
    if attack.startswith("fab"):
        wrong = [i for i, lab in enumerate(case["labels"]) if lab == 0]  # predicted bonafide, labelled spoof
        assert all(torch.equal(xa[i], x[i]) for i in wrong)
    with torch.no_grad():
        la = fwd(xa, state)
    if attack == "fab_l2":  # the returned iterates sit ON the boundary (reference logits -2e-7 ... -4e-6): compare the values
        np.testing.assert_allclose(la.numpy(), g[f"{attack}_logits_adv"], atol=2e-5)
    else:
        assert np.array_equal((la.numpy() > 0), (g[f"{attack}_logits_adv"] > 0)), "label flips differ"


def test_oracle_projection_linf_solves_the_box_hyperplane_problem():
    """fab.py:562-614: the returned d satisfies <w, t + d> = b with t + d in [0,1]^n whenever the hyperplane meets the
    box (property check of the restatement, independent of the golden vectors)."""
    g = torch.Generator().manual_seed(5)
    t = torch.rand(6, 500, generator=g)
    w = torch.randn(6, 500, generator=g)
    z = torch.rand(6, 500, generator=g)  # a point of the box: choose b so the hyperplane passes through it
    b = (w * z).sum(1)
    d = oatk.projection_linf(t, w, b)
    assert ((t + d) >= -1e-6).all() and ((t + d) <= 1 + 1e-6).all()
    np.testing.assert_allclose((w * (t + d)).sum(1).numpy(), b.numpy(), rtol=1e-4, atol=1e-4)


def test_oracle_projection_l2_solves_the_box_hyperplane_problem():
    """fab.py:617-665: <w, t + d> = b inside the box, and no other feasible move is shorter in L2 (checked against the
    L-inf projection's move, which is feasible for the same constraint)."""
    g = torch.Generator().manual_seed(6)
    t = torch.rand(6, 500, generator=g)
    w = torch.randn(6, 500, generator=g)
    z = torch.rand(6, 500, generator=g)
    b = (w * z).sum(1)
    d = oatk.projection_l2(t, w, b)
    assert ((t + d) >= -1e-6).all() and ((t + d) <= 1 + 1e-6).all()
    np.testing.assert_allclose((w * (t + d)).sum(1).numpy(), b.numpy(), rtol=1e-4, atol=1e-4)
    d_inf = oatk.projection_linf(t, w, b)
    assert (d.norm(dim=1) <= d_inf.norm(dim=1) * (1 + 1e-5)).all()


def test_oracle_mel_spec_matches_reference():
    """src/frontends.py:53-79: the oracle's `mel_spec` against what the unmodified prepare_mel_scale_vector produced
    (oracle/make_golden_melspec.py), and the table advb200's MEL_SCALE_FN holds against the reference's MelScale buffer."""
    from advb200 import frontends
    from oracle import frontend as ofe

    g = helpers.load_golden("mel_spec")
    fb = torch.from_numpy(g["fb"])
    assert torch.equal(frontends.MEL_SCALE_FN.fb, fb) and frontends.get_frontend(["mel_spec"]) is frontends.prepare_mel_scale_vector
    for tag, cfg_id, B, T in (("t16000", 31, 2, 16000), ("t16150", 32, 1, 16150)):
        x, _ = synth.clips(cfg_id, B, T)
        assert abs(x.double().sum().item() - float(g[f"{tag}_x_sum"])) < 1e-9
        got, want = ofe.mel_spec(x, fb), torch.from_numpy(g[f"{tag}_out"])
        assert got.shape == want.shape == (B, 2, 80, 101)
        zg, zw = torch.polar(got[:, 0], got[:, 1]), torch.polar(want[:, 0], want[:, 1])
        assert (zg - zw).abs().max().item() < 1e-5 * zw.abs().max().item()


def test_oracle_matches_reference_at_config1():
    """BASELINE.json configs[0] (FGSM eps=0.005, LCNN+LFCC, batch 8, 64 000 samples): the oracle port against the fixture the
    unmodified reference produced (oracle/make_golden_cfg.py)."""
    from oracle import lcnn as olcnn
    from oracle.make_golden_cfg import BIAS_KEY, CFG_CASES, T

    name = "cfg1_lcnn_fgsm_b8"
    case, g = CFG_CASES[name], helpers.load_golden(name)
    x, y = synth.clips(case["cfg_id"], case["B"], T)
    _, state = cases.build_state(case["model"], case["frontend"])
    state[BIAS_KEY[case["model"]]] = torch.from_numpy(g["bias"])
    assert synth.state_digest(state) == str(g["digest"])
    assert np.array_equal(y.numpy(), g["y"])
    fn = lambda v: olcnn.forward(v, state)  # noqa: E731
    with torch.no_grad():
        np.testing.assert_allclose(fn(x).numpy(), g["logits_clean"], atol=3e-6)
    xa = oatk.fgsm(fn, x, y, case["params"]["eps"])
    np.testing.assert_allclose((xa - x).abs().amax(dim=1).numpy(), g["delta_linf"], atol=1e-5)
    np.testing.assert_allclose((xa - x).norm(p=2, dim=1).numpy(), g["delta_l2"], rtol=1e-5)
    sign = np.unpackbits(g["sign_bits"])[: x.numel()].reshape(x.shape).astype(bool)
    assert ((xa > x).numpy() != sign).mean() < 2e-3
    with torch.no_grad():
        la = fn(xa)
    np.testing.assert_allclose(la.numpy(), g["logits_adv"], atol=1e-4)
    assert np.array_equal((torch.sigmoid(la.squeeze(1)) + .5).int().numpy(), g["pred_adv"])


def test_oracle_targeted_modes_match_reference():
    """attack.py:60-108 + fgsm.py:49-50 / pgd.py:64-65 / pgdl2.py:69-70 / cw.py:82-83,131-132: the oracle's targeted variants
    (negated ascent on the target labels' loss) against the reference's own classes, run live on CPU."""
    from oracle import ref

    if not ref.available():
        pytest.skip("reference not staged (oracle/_ref)")
    ta = ref.torchattacks()
    name = "lcnn_lfcc_t16000_margin"
    case, x, y, holder, state, fwd = helpers.case_setup(name)
    model = ref.model("lcnn", "lfcc", state)
    fn = lambda v: fwd(v, state)  # noqa: E731
    tmap = lambda images, labels: 1 - labels  # noqa: E731
    for kind in ("fgsm", "pgd", "pgdl2", "cw"):
        if kind == "fgsm":
            atk = ta.FGSM(model, eps=0.005)
        elif kind == "pgd":
            atk = ta.PGD(model, eps=0.001, alpha=2 / 255, steps=3, random_start=True)
        elif kind == "pgdl2":
            atk = ta.PGDL2(model, eps=0.1, alpha=0.2, steps=3, random_start=True)
        else:
            atk = ta.CW(model, c=1e4, kappa=0.0, steps=5, lr=5e-4)
        # model_training=False (torchattacks' default): with model_training=True the reference's _get_target_label
        # (attack.py:262-268) calls model.eval(); ...; model.train(), which switches BatchNorm and Dropout(0.7) back to
        # TRAINING mode for the attack forward - a stock-torchattacks quirk that makes targeted attacks on LCNN random
        # (measured: 50 % of the FGSM signs differ run to run).  The native engine always runs the eval forward.
        atk.set_training_mode(model_training=False)
        atk.set_mode_targeted_by_function(tmap)
        torch.manual_seed(2000 + case["cfg_id"])
        model.eval()
        want = atk(x, y)
        # CW: tanh-space Adam takes sign-like +-lr steps, which amplify the 1e-6 gradient differences between two model
        # implementations wherever the classification and distance terms nearly cancel (10 % of one clip's samples move by up
        # to 3e-4 here) - so the targeted CW *logic* is pinned with the reference's own nn.Module as the model function
        # (bit-exact), and the oracle's model is pinned by every other test in this file
        got = helpers.oracle_targeted(kind, (lambda v: model(v)) if kind == "cw" else fn, x, y, 1 - y, case)
        if kind in ("fgsm", "pgd"):
            assert (got != want).float().mean().item() < 2e-3, kind
        else:
            assert (got - want).abs().max().item() < 1e-5, kind
