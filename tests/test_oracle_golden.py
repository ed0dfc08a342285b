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
                                         ("rawnet3_t16000_margin", "fab")])
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
        assert (xa - ref).abs().max().item() < 1e-5
    if attack == "fab":
        wrong = [i for i, lab in enumerate(case["labels"]) if lab == 0]  # predicted bonafide, labelled spoof
        assert all(torch.equal(xa[i], x[i]) for i in wrong)
    with torch.no_grad():
        la = fwd(xa, state)
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
