"""GPU: the CUDA path (through the C ABI) against the oracle and the reference-generated golden fixtures."""
import numpy as np
import pytest
import torch

from oracle import attacks as oatk
from oracle import cases
from oracle import frontend as ofe
from oracle.lcnn import BLOCKS

import helpers

pytestmark = pytest.mark.gpu

LCNN_CASES = [n for n, c in cases.CASES.items() if c["model"] == "lcnn"]
SPEC_CASES = [n for n, c in cases.CASES.items() if c["model"] == "specrnet"]


def _setup(name, dev):
    from advb200 import engine

    case, x, y, holder, state, fwd = helpers.case_setup(name)
    holder = helpers.load_holder_state(holder, state, dev)
    eng = engine.engine_for(holder, x.shape[0], x.shape[1])
    return case, x, y, holder, state, fwd, eng


def _interior(t, p):
    return t[:, p:t.shape[1] - p, p:t.shape[2] - p, :] if p > 0 else t


@pytest.mark.parametrize("name", [n for n, c in cases.CASES.items() if c["frontend"] != "none"])
def test_frontend_forward_backward(name, cuda_device):
    case, x, y, holder, state, fwd, eng = _setup(name, cuda_device)
    fb, dct, win, _ = ofe.tables_from_state(state)
    xc = x.clone().requires_grad_(True)
    want = ofe.cepstral_frontend(xc, fb, dct, win)
    got = eng.frontend_fwd(x.to(cuda_device)).cpu()
    # fp32 features of magnitude <= ~500: 1e-3 absolute (SURVEY.md App. A.1 measured 5e-4 between torch paths)
    assert (got - want).abs().max().item() < 1e-3
    np.testing.assert_allclose(got.numpy(), helpers.load_golden(name)["frontend"], atol=2e-3)
    gc = torch.randn(want.shape, generator=torch.Generator("cpu").manual_seed(5))
    (gx_want,) = torch.autograd.grad((want * gc).sum(), xc)
    gx = eng.frontend_bwd(x.to(cuda_device), gc.to(cuda_device)).cpu()
    if not case["silence"]:
        assert helpers.rel_err(gx, gx_want) < 1e-5
    else:  # floor active: clamp membership of borderline elements may differ by ulps (see test_oracle_golden)
        assert helpers.cosine(gx, gx_want) > 0.999


@pytest.mark.parametrize("T", [6399, 6400, 6401, 12801, 19999])
def test_frontend_clip_lengths_around_the_backward_tile(T, cuda_device):
    """fe_bwd works on tiles of 40 hops = 6 400 samples, one warp per frame pair: lengths just below / at / above a tile
    boundary (a last tile of ONE sample included) and odd lengths, forward and backward against the oracle."""
    from advb200 import engine

    case, x0, y0, holder, state, fwd = helpers.case_setup("lcnn_lfcc_t16000")
    holder = helpers.load_holder_state(holder, state, cuda_device)
    g = torch.Generator("cpu").manual_seed(T)
    x = torch.rand(3, T, generator=g)
    eng = engine.engine_for(holder, 3, T)
    fb, dct, win, _ = ofe.tables_from_state(state)
    xc = x.clone().requires_grad_(True)
    want = ofe.cepstral_frontend(xc, fb, dct, win)
    got = eng.frontend_fwd(x.to(cuda_device)).cpu()
    assert got.shape == want.shape
    assert (got - want).abs().max().item() < 1e-3
    gc = torch.randn(want.shape, generator=g)
    (gx_want,) = torch.autograd.grad((want * gc).sum(), xc)
    gx = eng.frontend_bwd(x.to(cuda_device), gc.to(cuda_device)).cpu()
    assert helpers.rel_err(gx, gx_want) < 1e-5


def _oracle_taps(x, y, state, fwd, feat=None):
    """Oracle forward/backward with every stage tapped.  With ``feat`` (B,1,80,F) the embedding is evaluated at those
    features (a leaf) and the waveform gradient is the oracle frontend's VJP of the resulting feature gradient."""
    from oracle import lcnn as olcnn

    taps = {}
    xc = x.clone().requires_grad_(True)
    if feat is None:
        o = fwd(xc, state, taps)
    else:
        leaf = feat.clone().requires_grad_(True)
        taps["frontend"] = leaf
        o = olcnn.embedding(leaf, state, taps)
    for v in taps.values():
        if v.requires_grad:  # the integer winner codes are not differentiable
            v.retain_grad()
    torch.nn.functional.cross_entropy(torch.cat([-o, o], dim=1), y).backward()
    if feat is None:
        return o.detach(), taps, xc.grad
    fb, dct, win, _ = ofe.tables_from_state(state)
    (gx,) = torch.autograd.grad((ofe.cepstral_frontend(xc, fb, dct, win).unsqueeze(1) * taps["frontend"].grad).sum(), xc)
    return o.detach(), taps, gx


@pytest.mark.parametrize("conv_path", [0, 1], ids=["tcgen05", "simt"])
@pytest.mark.parametrize("name", LCNN_CASES[:3])
def test_every_stage_forward_and_backward(name, conv_path, cuda_device):
    case, x, y, holder, state, fwd, eng = _setup(name, cuda_device)
    eng.set_option("conv_path", conv_path)
    try:
        g, logits = eng.grad(x.to(cuda_device), y.to(cuda_device))
    finally:
        eng.set_option("conv0_fwd", 0)
        eng.set_option("conv_path", 0)
    B = x.shape[0]
    feat = None
    if case["silence"]:
        # Silent frames clamp to the dB floor, their cepstrum is DC-only and every other coefficient is fp32 noise
        # (~1e-5): the 2x2 max-pool of block 0 then picks its winner among noise-level ties, so d loss/d features
        # moves by 28 % under 1e-5 feature noise IN THE ORACLE ITSELF (measured; see DESIGN.md "silence case").
        # The meaningful check is therefore the oracle evaluated at the engine's own features.
        t, p = eng.debug_stage("frontend")
        feat = _interior(t, p)[:B].permute(0, 3, 2, 1).contiguous().cpu()  # (B,F,80,1) -> (B,1,80,F)
    o, taps, gx_want = _oracle_taps(x, y, state, fwd, feat)
    tol = 2e-5 if not case["silence"] else 5e-3
    assert (logits.cpu() - o).abs().max().item() < 2e-6
    for i, (idx, _, _) in enumerate(BLOCKS):
        t, p = eng.debug_stage(f"block{i}")
        assert helpers.rel_err(_interior(t, p)[:B].permute(0, 3, 1, 2).cpu(), taps[f"block{idx}"].detach()) < 2e-5, i
        t, _ = eng.debug_stage(f"gblock{i}")
        assert helpers.trimmed_rel_err(t[:B].permute(0, 3, 1, 2).cpu(), taps[f"block{idx}"].grad) < tol, i
    for nm in ("feats", "lstm1", "lstm2"):
        t, _ = eng.debug_stage(nm)
        assert helpers.rel_err(t[:B, :, 0, :].cpu(), taps[nm].detach()) < 2e-5, nm
    if conv_path == 0 and not case["silence"]:
        # COUNT of Max-Feature-Map / max-pool winners the engine decides differently from the oracle: the
        # gradient gates above are flip-robust, this bounds how many flips they absorb.  A winner decided by a margin below the
        # fp32 noise of the conv sums (~1e-6 relative) may legitimately differ; measured 0-3 per block of up to 128 000 elements.
        flips = {}
        for i, (idx, _, _) in enumerate(BLOCKS):
            t, _ = eng.debug_stage(f"codes{i}")
            got = t[:B].permute(0, 3, 1, 2).cpu().long()
            want = taps[f"codes{idx}"]
            assert got.shape == want.shape, (i, got.shape, want.shape)
            flips[f"block{i}"] = (int((got != want).sum()), want.numel())
        print("winner flips vs oracle", name, flips)
        assert all(n <= max(4, 1e-4 * tot) for n, tot in flips.values()), flips
    t, _ = eng.debug_stage("gcoef")  # (B,F,80,1): d loss / d cepstral image
    gc = t[:B].permute(0, 3, 2, 1).cpu()
    if not case["silence"]:
        assert helpers.grads_agree(gc, taps["frontend"].grad, tol)
        assert helpers.grads_agree(g.cpu(), gx_want, tol)
        assert helpers.grads_agree(g.cpu(), torch.from_numpy(helpers.load_golden(name)["grad"]), tol)
    else:
        # block 0's pool winners in the silent frames are decided by ~1e-7 differences of the conv sums themselves
        # (any two correct fp32 convolutions disagree there), so d/d features is compared tightly only on the
        # frames / samples of the second half of the clip (the silence spans [5/16 T, 15/32 T)) and by direction
        # overall.
        fh, th = gc.shape[-1] // 2 + 4, x.shape[1] // 2 + 1024
        assert helpers.rel_err(gc[..., fh:], taps["frontend"].grad[..., fh:]) < 1e-4
        assert helpers.cosine(gc, taps["frontend"].grad) > 0.99
        assert helpers.rel_err(g.cpu()[:, th:], gx_want[:, th:]) < 1e-4
        assert helpers.cosine(g.cpu(), gx_want) > 0.98


@pytest.mark.parametrize("conv_path", [0, 1], ids=["tcgen05", "simt"])
@pytest.mark.parametrize("name", LCNN_CASES)
def test_logits_and_gradient_against_reference_golden(name, conv_path, cuda_device):
    case, x, y, holder, state, fwd, eng = _setup(name, cuda_device)
    g = helpers.load_golden(name)
    eng.set_option("conv_path", conv_path)
    try:
        grad, logits = eng.grad(x.to(cuda_device), y.to(cuda_device))
    finally:
        eng.set_option("conv_path", 0)
    np.testing.assert_allclose(logits.cpu().numpy(), g["logits"], atol=3e-6)
    ref = torch.from_numpy(g["grad"])
    if not case["silence"]:
        # tight on >= 90 % of the samples; an isolated pool / MFM winner decided by a ~1e-7 margin may flip between
        # two correct fp32 implementations (helpers.trimmed_rel_err) - the fp32 SIMT path happens to have none here
        assert helpers.grads_agree(grad.cpu(), ref)
        if conv_path == 1:
            assert helpers.rel_err(grad.cpu(), ref) < 2e-5
    else:
        # the reference's own gradient is chaotic in silent frames (pool arg-max among fp32-noise ties, see
        # test_every_stage_forward_and_backward); away from them (second half of each clip: the silence spans
        # [5/16 T, 15/32 T) and the oracle's own noise sensitivity there is 5e-7) it must agree tightly
        h = x.shape[1] // 2
        assert helpers.rel_err(grad.cpu()[:, h:], ref[:, h:]) < 1e-4
        assert (torch.sign(grad.cpu()) == torch.sign(ref)).float().mean().item() > 0.95
    # holder(x) is the drop-in model call: same logits
    np.testing.assert_allclose(holder(x.to(cuda_device)).cpu().numpy(), g["logits"], atol=3e-6)


@pytest.mark.parametrize("name", LCNN_CASES[:2] + SPEC_CASES[:1])
@pytest.mark.parametrize("attack", ["fgsm", "pgd", "pgdl2"])
def test_attacks_against_oracle_and_golden(name, attack, cuda_device):
    from advb200 import torchattacks as ta

    case, x, y, holder, state, fwd, eng = _setup(name, cuda_device)
    g = helpers.load_golden(name)
    p = cases.ATTACKS[attack]
    xd, yd = x.to(cuda_device), y.to(cuda_device)
    if attack == "fgsm":
        atk = ta.FGSM(holder, eps=p["eps"])
        atk.set_training_mode(True, False)
        got = atk(xd, yd)
    elif attack == "pgd":
        atk = ta.PGD(holder, eps=p["eps"], alpha=p["alpha"], steps=p["steps"])
        atk.set_training_mode(True, False)
        got = atk.forward(xd, yd, noise=helpers.reference_start(case, "pgd", x, p["eps"]).to(cuda_device))
    else:
        atk = ta.PGDL2(holder, eps=p["eps"], alpha=p["alpha"], steps=p["steps"])
        atk.set_training_mode(True, False)
        got = atk.forward(xd, yd, delta=helpers.reference_start(case, "pgdl2", x, p["eps"]).to(cuda_device))
    assert got.data_ptr() != xd.data_ptr() and torch.equal(xd.cpu(), x), "inputs must not be mutated"
    got = got.cpu()
    assert got.min().item() >= 0.0 and got.max().item() <= 1.0
    d = got - x
    # the attack's own perturbation norm within 1e-5 of the reference (north-star tolerance)
    if attack == "pgdl2":
        np.testing.assert_allclose(d.norm(p=2, dim=1).numpy(), g["pgdl2_delta_l2"], rtol=1e-5, atol=1e-5)
    else:
        np.testing.assert_allclose(d.abs().amax(dim=1).numpy(), g[f"{attack}_delta_linf"], atol=1e-5)
        np.testing.assert_allclose(d.norm(p=2, dim=1).numpy(), g[f"{attack}_delta_l2"], rtol=1e-5, atol=1e-5)
    want = helpers.oracle_attack(name, attack, x, y, state, fwd, case)
    ref = torch.from_numpy(g[f"{attack}_adv"])
    if not case["silence"]:
        if attack == "pgdl2":
            assert (got - want).abs().max().item() < 5e-4 and (got - ref).abs().max().item() < 5e-4
        else:  # sign steps: identical up to gradient-sign ties
            assert (got != want).float().mean().item() < 2e-3 and (got != ref).float().mean().item() < 2e-3
    # predicted labels of the attacked batch: bit-exact with the reference
    la = eng.forward(got.to(cuda_device)).cpu().numpy()
    if not case["silence"]:
        assert np.array_equal(la > 0, g[f"{attack}_logits_adv"] > 0)
    # ... and at the reference's own adversarial batch the engine must reproduce the reference's logits (this also
    # covers the silence case, whose gradient signs - hence adversarial samples - are chaotic in the oracle itself)
    lr = eng.forward(ref.to(cuda_device)).cpu().numpy()
    np.testing.assert_allclose(lr, g[f"{attack}_logits_adv"], atol=3e-6)
    assert np.array_equal(lr > 0, g[f"{attack}_logits_adv"] > 0)


def test_minmax_roundtrip_and_edge_cases(cuda_device):
    from advb200 import engine

    raw = 0.1 * torch.randn(5, 16000, generator=torch.Generator("cpu").manual_seed(3))
    x01, mn, mx = engine.to_minmax(raw.to(cuda_device))
    w01, wmn, wmx = oatk.to_minmax(raw)
    assert torch.equal(x01.cpu(), w01) and torch.equal(mn.cpu(), wmn) and torch.equal(mx.cpu(), wmx)
    back = engine.revert_minmax(x01, mn, mx).cpu()
    assert torch.equal(back, oatk.revert_minmax(w01, wmn, wmx))


def test_fused_minmax_attack_matches_the_three_call_sequence(cuda_device):
    """to_minmax -> atk -> revert_minmax in one native call (advb_attack_minmax) is bit-identical to the
    three calls the reference makes (evaluate_models_on_adversarial_attacks.py:219-221)."""
    from advb200 import aa
    from advb200 import torchattacks as ta

    name = "lcnn_lfcc_t16000"
    case, x, y, holder, state, fwd, eng = _setup(name, cuda_device)
    raw = (0.3 * (x - 0.4)).to(cuda_device)  # an unscaled waveform batch
    yd = y.to(cuda_device)
    for atk in (ta.FGSM(holder, eps=0.005), ta.PGD(holder, eps=0.001, steps=3), ta.PGDL2(holder, eps=0.1, steps=2)):
        atk.set_training_mode(True, False)
        torch.manual_seed(7)
        torch.cuda.manual_seed(7)
        x01, mn, mx = aa.to_minmax(raw)
        want = aa.revert_minmax(atk(x01, yd), mn, mx)
        torch.manual_seed(7)
        torch.cuda.manual_seed(7)
        got = aa.attack_minmax(atk, raw, yd)
        assert torch.equal(got, want), type(atk).__name__
        assert not atk._fused_minmax
    ref = raw.cpu()
    mn, mx = ref.min(dim=1, keepdim=True)[0], ref.max(dim=1, keepdim=True)[0]
    assert (got.cpu() - ref).abs().max().item() <= 0.11 * (mx - mn).max().item()  # an L2 ball of 0.1 in the scaled domain


def test_handle_errors_are_loud(cuda_device):
    from advb200 import engine

    case, x, y, holder, state, fwd, eng = _setup("lcnn_lfcc_t16000", cuda_device)
    with pytest.raises(ValueError):
        eng.forward(torch.zeros(2, 8000, device=cuda_device))
    with pytest.raises(RuntimeError, match="no CPU path"):
        eng.forward(torch.zeros(2, 16000))
    # batch grows -> handle is rebuilt, results unchanged for the first clips
    xb = torch.cat([x, x], 0).to(cuda_device)
    l4 = engine.engine_for(holder, 4, 16000).forward(xb).cpu()
    assert torch.allclose(l4[:2], l4[2:], atol=0, rtol=0)


def test_live_weights_are_reread(cuda_device):
    """Adversarial training mutates the attacked model between calls (src/trainer.py:309-331)."""
    case, x, y, holder, state, fwd, eng = _setup("lcnn_lfcc_t16000", cuda_device)
    xd = x.to(cuda_device)
    a = eng.forward(xd).cpu()
    with torch.no_grad():
        holder.m_output_act.bias += 0.25
    b = eng.forward(xd).cpu()
    assert torch.allclose(b - a, torch.full_like(a, 0.25), atol=1e-6)


@pytest.mark.parametrize("name", ["lcnn_lfcc_t16000", "lcnn_lfcc_t64000"])
def test_tensor_core_path_matches_simt_path(name, cuda_device):
    """conv_path 0 (tcgen05, 3xTF32) and 1 (fp32 SIMT) are two independent implementations of the same blocks."""
    case, x, y, holder, state, fwd, eng = _setup(name, cuda_device)
    xd, yd = x.to(cuda_device), y.to(cuda_device)
    try:
        eng.set_option("conv_path", 1)
        g_simt, l_simt = eng.grad(xd, yd)
        eng.set_option("conv_path", 0)
        g_tc, l_tc = eng.grad(xd, yd)
        assert (l_tc - l_simt).abs().max().item() < 2e-6
        assert helpers.grads_agree(g_tc.cpu(), g_simt.cpu())
        # the one-tile-per-CTA kernels (conv_tc.cu only) and the persistent warp-specialised kernels (conv_light.cu,
        # conv_p3.cu) share weight images, MMA order and epilogue arithmetic: forward bit-identical; the backward of the
        # pooled 3x3 blocks differs only by the summation order of the horizontal-scatter formulation
        # (first block forward through the im2col kernel for this comparison: its Toeplitz replacement sums the 25 taps in a
        # different order)
        eng.set_option("conv0_fwd", 1)
        g_i2c, l_i2c = eng.grad(xd, yd)
        eng.set_option("conv_sched", 1)
        g_classic, l_classic = eng.grad(xd, yd)
        eng.set_option("conv_sched", 0)
        eng.set_option("conv0_fwd", 0)
        assert torch.equal(l_classic, l_i2c)
        assert helpers.rel_err(g_classic.cpu(), g_i2c.cpu()) < 2e-5
        # Toeplitz first block (default) against the im2col first block: same 3xTF32 products, different summation order
        assert (l_i2c - l_tc).abs().max().item() < 1e-6
        assert helpers.grads_agree(g_tc.cpu(), g_i2c.cpu())
        # first block backward: the fp32 cell kernel (default, conv0_bwd.cu) against the tcgen05 GEMM + col2im version
        eng.set_option("conv0_bwd", 1)
        g_c0tc, _ = eng.grad(xd, yd)
        eng.set_option("conv0_bwd", 0)
        assert helpers.rel_err(g_c0tc.cpu(), g_tc.cpu()) < 2e-5
        # single-pass tf32: reduced precision, documented as an opt-in; sanity only
        eng.set_option("tf32_passes", 1)
        g_fast, l_fast = eng.grad(xd, yd)
        assert (l_fast - l_simt).abs().max().item() < 5e-3
        assert helpers.cosine(g_fast, g_simt) > 0.98
    finally:
        eng.set_option("tf32_passes", 3)
        eng.set_option("conv0_fwd", 0)
        eng.set_option("conv_sched", 0)
        eng.set_option("conv_path", 0)
        eng.set_option("conv_sched", 0)
        eng.set_option("conv0_bwd", 0)


def test_specrnet_tensor_core_conv2_matches_simt(cuda_device):
    """Option sr_tc: conv2 of SpecRNet's 64-channel blocks on the persistent tcgen05 kernel (3xTF32; conv_p3's PLAIN variant, and the
    same kernel fed with the flipped / channel-transposed image as the transposed convolution) against the fp32 SIMT kernels."""
    name = SPEC_CASES[0]
    case, x, y, holder, state, fwd, eng = _setup(name, cuda_device)
    xd, yd = x.to(cuda_device), y.to(cuda_device)
    try:
        g1, l1 = eng.grad(xd, yd)
        eng.set_option("sr_tc", 0)
        g0, l0 = eng.grad(xd, yd)
    finally:
        eng.set_option("sr_tc", 1)
    assert (l1 - l0).abs().max().item() < 2e-6
    assert helpers.grads_agree(g1.cpu(), g0.cpu())
    assert not torch.equal(g1, g0), "the two paths are different kernels: identical bits would mean the option is not wired"


def test_projection_linf_against_oracle(cuda_device):
    """fab.py:562-614: the sort-free CUDA projection against the sort-based restatement, all three branches
    (unreachable hyperplane -> box corner, uniform level, saturating level), ragged row length."""
    from advb200 import engine

    g = torch.Generator("cpu").manual_seed(11)
    R, T = 12, 16001
    t = torch.rand(R, T, generator=g)
    w = torch.randn(R, T, generator=g) * 1e-3
    w[3, ::7] = 0.0  # exact zeros: (w != 0) mask
    z = torch.rand(R, T, generator=g)
    b = (w * z).sum(1)                       # hyperplane through a box point: saturating level
    b[0] = (w[0] * t[0]).sum() + 1e-4        # almost on the plane: uniform level (c_l)
    b[1] = (w[1] * t[1]).sum() - 2e-4
    b[2] = (w[2].abs()).sum() * 5            # does not meet the box: corner
    t[4, :100] = 0.0                         # ties at the box faces
    t[4, 100:200] = 1.0
    want = oatk.projection_linf(t, w, b)
    got = engine.projection_linf(t.to(cuda_device), w.to(cuda_device), b.to(cuda_device)).cpu()
    # the attack only uses d through x + eta d and ||d||_inf: compare those scales
    scale = want.abs().amax(dim=1).clamp_min(1e-12)
    assert ((got - want).abs().amax(dim=1) / scale).max().item() < 2e-4
    np.testing.assert_allclose(got.abs().amax(dim=1).numpy(), want.abs().amax(dim=1).numpy(), rtol=2e-4)
    hit = [i for i in range(R) if i != 2]
    np.testing.assert_allclose((w * (t + got)).sum(1)[hit].numpy(), b[hit].numpy(), rtol=1e-3, atol=2e-5)


def test_projection_l2_against_oracle(cuda_device):
    """fab.py:617-665: the sort-free CUDA projection against the sort-based restatement, all three branches (no coordinate
    saturates / some saturate / the hyperplane is out of the box's reach), immovable coordinates, ragged row length."""
    from advb200 import engine

    g = torch.Generator("cpu").manual_seed(12)
    R, T = 12, 16001
    t = torch.rand(R, T, generator=g)
    w = torch.randn(R, T, generator=g) * 1e-3
    w[3, ::7] = 0.0                          # |w| < 1e-8: never move
    w[5, ::5] = 5e-9
    z = torch.rand(R, T, generator=g)
    b = (w * z).sum(1)                       # hyperplane through a box point: saturating alpha (c2)
    t[:2] = 0.3 + 0.4 * t[:2]                # away from the faces, so that a well-conditioned offset stays in branch c4
    b[0] = (w[0] * t[0]).sum() + 1e-2        # close to the plane: uniform alpha, no coordinate saturates (c4)
    b[1] = (w[1] * t[1]).sum() - 2e-2
    b[2] = (w[2].abs()).sum() * 5            # does not meet the box: every coordinate at its face (c3)
    t[4, :100] = 0.0                         # already at a face
    t[4, 100:200] = 1.0
    want = oatk.projection_l2(t, w, b)
    got = engine.projection_l2(t.to(cuda_device), w.to(cuda_device), b.to(cuda_device)).cpu()
    # the attack only uses d through x + eta d and ||d||_2: compare at those scales
    scale = want.abs().amax(dim=1).clamp_min(1e-12)
    assert ((got - want).abs().amax(dim=1) / scale).max().item() < 2e-4
    np.testing.assert_allclose(got.norm(dim=1).numpy(), want.norm(dim=1).numpy(), rtol=2e-4)
    hit = [i for i in range(R) if i != 2]
    np.testing.assert_allclose((w * (t + got)).sum(1)[hit].numpy(), b[hit].numpy(), rtol=1e-3, atol=2e-5)
    assert (got[3, ::7] == 0).all() and (got[5, ::5] == 0).all()
    assert ((t + got) >= -1e-6).all() and ((t + got) <= 1 + 1e-6).all()


def test_fab_l2_against_oracle_and_golden(cuda_device):
    """torchattacks.FAB(norm='L2') (fab.py:184-194,216-219,236-240,251-253,277-279,617-665) against the
    fixture the unmodified reference produced; norm='L1' cannot run upstream either and must say so."""
    from advb200 import torchattacks as ta

    name, attack = "lcnn_lfcc_t16000_margin_l2", "fab_l2"
    case, x, y, holder, state, fwd, eng = _setup(name, cuda_device)
    g = helpers.load_golden(name)
    p = cases.ATTACKS[attack]
    xd, yd = x.to(cuda_device), y.to(cuda_device)
    atk = ta.FAB(holder, norm="L2", eps=p["eps"], steps=p["steps"], eta=p["eta"], alpha_max=p["alpha_max"], beta=p["beta"],
                 n_classes=2)
    atk.set_training_mode(True, False)
    got = atk(xd, yd).cpu()
    assert torch.equal(xd.cpu(), x) and got.min().item() >= 0.0 and got.max().item() <= 1.0
    d = got - x
    ref = torch.from_numpy(g["fab_l2_adv"])
    # FAB-L2's own norm is L2
    np.testing.assert_allclose(d.norm(p=2, dim=1).numpy(), g["fab_l2_delta_l2"], rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(d.abs().amax(dim=1).numpy(), g["fab_l2_delta_linf"], rtol=1e-3, atol=1e-6)
    assert torch.equal(got[2], x[2]), "a clip that is misclassified from the start must come back untouched"
    # the L2 step is proportional to the gradient itself: elements agree to the gradient's own parity
    assert helpers.rel_err(d, ref - x) < 5e-3  # oracle vs reference: 1e-3 (tests/test_oracle_golden.py)
    la = eng.forward(got.to(cuda_device)).cpu().numpy()
    # the returned iterates sit ON the decision boundary (reference logits -2e-7 ... -4e-6): their sign is below the fp32 noise
    # of any two forwards; what is comparable is that they are as close to it
    np.testing.assert_allclose(la[[0, 1, 3]], g["fab_l2_logits_adv"][[0, 1, 3]], atol=2e-5)
    assert la[2] > 0
    lr = eng.forward(ref.to(cuda_device)).cpu().numpy()
    np.testing.assert_allclose(lr, g["fab_l2_logits_adv"], atol=3e-6)
    # restarts (fab.py:184-194): the L2 random start stays inside the eps / 2 ball and the result never loses an adversarial clip
    two = ta.FAB(holder, norm="L2", eps=p["eps"], steps=p["steps"], eta=p["eta"], alpha_max=p["alpha_max"], beta=p["beta"],
                 n_classes=2, n_restarts=2)
    two.set_training_mode(True, False)
    got2 = two(xd, yd).cpu()
    assert ((got2 != x).any(dim=1) | ~(got != x).any(dim=1)).all()
    assert (got2 - x).norm(p=2, dim=1).max().item() <= p["eps"] + 1e-5
    with pytest.raises(NotImplementedError):
        ta.FAB(holder, norm="L1")


@pytest.mark.parametrize("attack", ["fab", "cw"])
def test_fab_cw_against_oracle_and_golden(attack, cuda_device):
    from advb200 import torchattacks as ta

    name = "lcnn_lfcc_t16000_margin"
    case, x, y, holder, state, fwd, eng = _setup(name, cuda_device)
    g = helpers.load_golden(name)
    p = cases.ATTACKS[attack]
    xd, yd = x.to(cuda_device), y.to(cuda_device)
    if attack == "fab":
        atk = ta.FAB(holder, norm="Linf", eps=p["eps"], steps=p["steps"], eta=p["eta"], alpha_max=p["alpha_max"],
                     beta=p["beta"], n_classes=2)
    else:
        atk = ta.CW(holder, c=p["c"], kappa=p["kappa"], steps=p["steps"], lr=p["lr"])
    atk.set_training_mode(True, False)
    got = atk(xd, yd)
    assert torch.equal(xd.cpu(), x), "inputs must not be mutated"
    got = got.cpu()
    assert got.min().item() >= 0.0 and got.max().item() <= 1.0
    d = got - x
    ref = torch.from_numpy(g[f"{attack}_adv"])
    if attack == "fab":
        # FAB's own norm is L-inf; the reference's run-to-run noise on it is 5e-7: 1e-5 relative + 1e-6
        np.testing.assert_allclose(d.abs().amax(dim=1).numpy(), g["fab_delta_linf"], rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(d.norm(p=2, dim=1).numpy(), g["fab_delta_l2"], rtol=1e-4, atol=1e-5)
        assert torch.equal(got[2], x[2]), "a clip that is misclassified from the start must come back untouched"
        # like FGSM/PGD, the L-inf projection moves every sample by +-lambda along -sign(w): elements agree except where
        # the gradient sign itself is a tie between two fp32 implementations
        assert ((got - ref).abs() > 1e-5).float().mean().item() < 2e-3
    else:
        # CW's own norm is L2.  Its first Adam steps are sign-like (m / sqrt(v) = +-1) on a gradient dominated by the
        # fp32 rounding of tanh(atanh(.)), so elements differ between any two libm's; the norm is what is stable.
        np.testing.assert_allclose(d.norm(p=2, dim=1).numpy(), g["cw_delta_l2"], rtol=2e-2, atol=1e-5)
    la = eng.forward(got.to(cuda_device)).cpu().numpy()
    assert np.array_equal(la > 0, g[f"{attack}_logits_adv"] > 0), "label flips differ"
    lr = eng.forward(ref.to(cuda_device)).cpu().numpy()
    np.testing.assert_allclose(lr, g[f"{attack}_logits_adv"], atol=3e-6)


@pytest.mark.parametrize("name", SPEC_CASES)
def test_specrnet_stages_logits_and_gradient(name, cuda_device):
    """SpecRNet (+ MFCC, whose dB floor is always active: mel filter 0 is identically zero) against the
    oracle at every block boundary and against the reference-generated golden logits / gradient."""
    case, x, y, holder, state, fwd, eng = _setup(name, cuda_device)
    g = helpers.load_golden(name)
    grad, logits = eng.grad(x.to(cuda_device), y.to(cuda_device))
    B = x.shape[0]
    taps = {}
    xc = x.clone().requires_grad_(True)
    o = fwd(xc, state, taps)
    for v in taps.values():
        v.retain_grad()
    torch.nn.functional.cross_entropy(torch.cat([-o, o], dim=1), y).backward()
    for nm in ("0", "2", "4"):
        # engine layout (B, frames, coeffs, C padded) -> oracle (B, C, coeffs, frames)
        t, p = eng.debug_stage(f"sr_xb{nm}")
        want = taps[f"block{nm}"].detach()
        assert helpers.rel_err(t[:B, :, :, :want.shape[1]].permute(0, 3, 2, 1).cpu(), want) < 2e-5, nm
        t, p = eng.debug_stage(f"sr_xn{nm}")
        want = taps[f"stage{nm}"].detach()
        assert helpers.rel_err(_interior(t, p)[:B, :, :, :want.shape[1]].permute(0, 3, 2, 1).cpu(), want) < 2e-5, nm
        t, _ = eng.debug_stage(f"sr_gn{nm}")
        assert helpers.trimmed_rel_err(t[:B, :, :, :want.shape[1]].permute(0, 3, 2, 1).cpu(), taps[f"stage{nm}"].grad) < 1e-4, nm
    assert (logits.cpu() - o.detach()).abs().max().item() < 2e-6
    np.testing.assert_allclose(logits.cpu().numpy(), g["logits"], atol=3e-6)
    t, _ = eng.debug_stage("gcoef")
    # (no sign criterion here: where SELU saturates in fp32 the engine's derivative is exactly 0, torch's is ~e^-20)
    # trimmed: an isolated pool winner decided by a ~1e-7 margin may flip between two correct fp32 implementations
    gcf = t[:B].permute(0, 3, 2, 1).cpu()
    assert helpers.trimmed_rel_err(gcf, taps["frontend"].grad) < 2e-5 and helpers.cosine(gcf, taps["frontend"].grad) > 0.9995
    assert helpers.grads_agree(grad.cpu(), xc.grad, 1e-4)
    assert helpers.grads_agree(grad.cpu(), torch.from_numpy(g["grad"]), 1e-4)
    np.testing.assert_allclose(holder(x.to(cuda_device)).cpu().numpy(), g["logits"], atol=3e-6)


@pytest.mark.parametrize("model,frontend", [("lcnn", "lfcc"), ("specrnet", "mfcc")])
def test_native_clip_length_64600_ragged_batch(model, frontend, cuda_device):
    """The reference's own clip length is 64 600 samples = 404 frames (src/datasets/base_dataset.py:27),
    BASELINE.json uses 64 000 = 401; nothing may hard-code either.  B = 3 (odd), checked against the pinned oracle, and the
    PGD step at that length against the oracle's update rule."""
    from advb200 import engine
    from advb200 import torchattacks as ta
    from oracle import synth

    T, B = 64600, 3
    x, y = synth.clips(21, B, T)
    fwd = helpers.ORACLE_FWD[model]
    holder, state = cases.build_state(model, frontend, calibrate_on=x, forward_fn=fwd)
    holder = helpers.load_holder_state(holder, state, cuda_device)
    eng = engine.engine_for(holder, B, T)
    xd, yd = x.to(cuda_device), y.to(cuda_device)
    grad, logits = eng.grad(xd, yd)
    o, g_want = oatk.loss_and_grad(lambda v: fwd(v, state), x, y)
    assert (logits.cpu() - o).abs().max().item() < 3e-6
    assert helpers.grads_agree(grad.cpu(), g_want, 1e-4)
    # one FGSM step from the engine's own gradient sign reproduces the engine's FGSM output bit for bit
    eps = 0.002
    got = ta.FGSM(holder, eps=eps)(xd, yd).cpu()
    want = torch.clamp(x + eps * grad.cpu().sign(), 0, 1)
    assert torch.equal(got, want)
    assert (got != oatk.fgsm(lambda v: fwd(v, state), x, y, eps)).float().mean().item() < 2e-3
