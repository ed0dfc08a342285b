"""GPU: the drop-in proof.  The reference's own scripts / trainer classes run UNEDITED with `advb200.install()` on the GPU
and are compared with the same code running the reference's vendored torchattacks on the box's CPU (oracle/_ref).

Both sides run in subprocesses: `install()` rebinds `adversarial_attacks.torchattacks` process-wide, and the CPU side needs
CUDA_VISIBLE_DEVICES="" before torch is imported.
"""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

from oracle import cases, ref

pytestmark = pytest.mark.gpu

TOOLS = os.path.join(cases.ROOT, "tools")


def _run(script, *argv, timeout=900):
    r = subprocess.run([sys.executable, os.path.join(TOOLS, script), *argv], capture_output=True, text=True, timeout=timeout,
                       cwd=cases.ROOT)
    assert r.returncode == 0, (r.stdout[-3000:], r.stderr[-3000:])
    return r.stdout


def _generate_attacks(impl, attack="FGSM", eps=0.005):
    out = _run("dropin_generate_attacks.py", "--impl", impl, "--attack", attack, "--eps", str(eps), "--clips", "16", "--batch", "8")
    line = [ln for ln in out.splitlines() if ln.startswith("DROPIN ")][-1]
    return json.loads(line[len("DROPIN "):])


@pytest.mark.skipif(not ref.available(), reason="reference not staged (oracle/_ref)")
def test_reference_generate_attacks_runs_on_the_native_engine(cuda_device, record_property):
    """BASELINE.json configs[0] / SURVEY.md §8(a) a15: evaluate_models_on_adversarial_attacks.generate_attacks() with
    AttackEnum bound to the native classes and the reference's own LCNN (load_model + nn.DataParallel) on cuda:0, against the
    same function with the vendored torchattacks on CPU."""
    want = _generate_attacks("reference")
    got = _generate_attacks("native")
    assert got["device"] == "cuda" and got["attack_class"] == "advb200.torchattacks.attacks.FGSM"
    assert want["device"] == "cpu" and want["attack_class"].startswith("adversarial_attacks.torchattacks")
    assert got["clips"] == want["clips"] == 16 and set(got["records"]) == set(want["records"])
    for name, w in want["records"].items():
        g = got["records"][name]
        assert g["y"] == w["y"] and g["pred_clean"] == w["pred_clean"], name
        assert g["pred"] == w["pred"], (name, g, w)                       # predicted label of the attacked clip: bit-exact
        # FGSM's own norm is L-inf; L2 = eps * sqrt(#samples not clamped at 0 / 1) moves by 8e-6 relative per sign tie there
        assert abs(g["linf"] - w["linf"]) < 1e-5 and abs(g["l2"] - w["l2"]) < 1e-4 * w["l2"], (name, g["l2"], w["l2"])
        assert abs(g["score"] - w["score"]) < 5e-5 and abs(g["score_clean"] - w["score_clean"]) < 5e-6, (name, g, w)
    assert got["metrics"]["accuracy"] == want["metrics"]["accuracy"]      # attack success rate: identical
    for k in ("eer", "auc", "f1_score"):
        assert abs(got["metrics"][k] - want["metrics"][k]) < 1e-6, k
    record_property("cfg1_native_seconds", got["seconds_generate_attacks"])
    record_property("cfg1_reference_cpu_seconds", want["seconds_generate_attacks"])
    print("generate_attacks(): native", got["seconds_generate_attacks"], "s, reference CPU", want["seconds_generate_attacks"], "s",
          got["metrics"])


@pytest.mark.skipif(not ref.available(), reason="reference not staged (oracle/_ref)")
def test_reference_generate_attacks_pgd_preset(cuda_device):
    """Same path with AttackEnum.PGD_eps001 (random start drawn on the device by the native class, on the CPU by the reference:
    different noise, so only what does not depend on it is compared: the L-inf norm, and the labels / accuracy, which PGD-10
    at eps 1e-3 decides by a wide margin on this checkpoint)."""
    want = _generate_attacks("reference", "PGD_eps001", 0.0)
    got = _generate_attacks("native", "PGD_eps001", 0.0)
    assert got["attack_class"] == "advb200.torchattacks.attacks.PGD"
    for name, w in want["records"].items():
        g = got["records"][name]
        assert abs(g["linf"] - w["linf"]) < 1e-5, (name, g, w)
        if abs(w["score"] - 0.5) > 2e-4:
            assert g["pred"] == w["pred"], (name, g, w)
    assert abs(got["metrics"]["accuracy"] - want["metrics"]["accuracy"]) <= 100.0 / 16


@pytest.mark.skipif(not ref.available(), reason="reference not staged (oracle/_ref)")
def test_reference_trainer_strategies_on_the_native_engine(cuda_device, tmp_path):
    """src/trainer.py's RANDOM / EQUAL (in-place half batch) / ONLY_ADV / ADAPTIVE attack call sites, as
    the reference's own trainer classes make them (CPU labels, raw waveforms, min-max inside), native vs reference-on-CPU."""
    a, b = str(tmp_path / "native.npz"), str(tmp_path / "reference.npz")
    _run("dropin_trainer.py", "--impl", "reference", "--out", b)
    _run("dropin_trainer.py", "--impl", "native", "--out", a)
    got, want = np.load(a), np.load(b)
    assert set(got.files) == set(want.files) and len(got.files) == 14
    raw = want["raw"]
    assert np.array_equal(got["raw"], raw) and np.array_equal(got["y"], want["y"])
    n_attacked = 0
    for k in want.files:
        if k in ("y", "raw"):
            continue
        g, w = got[k], want[k]
        # which clips the strategy attacked (RANDOM / ADAPTIVE may attack none; EQUAL attacks a random half in place)
        hit_g, hit_w = (g != raw).any(axis=1), (w != raw).any(axis=1)
        assert np.array_equal(hit_g, hit_w), k
        if k.startswith("EQUAL"):
            assert hit_w.sum() == raw.shape[0] // 2 and np.array_equal(g[~hit_g], raw[~hit_g]), k
        if k.startswith("ONLY_ADV"):
            assert hit_w.all(), k
        n_attacked += int(hit_w.sum())
        # FGSM in min-max space, reverted: every sample moves by +-eps (max - min); elements agree except at gradient-sign ties
        assert float((np.abs(g - w) > 1e-6).mean()) < 2e-3, k
        np.testing.assert_allclose(np.abs(g - raw).max(axis=1), np.abs(w - raw).max(axis=1), atol=1e-6)
    assert n_attacked >= 3 * raw.shape[0] + 3 * (raw.shape[0] // 2)
