"""CPU: host-side logic, the C-ABI library surface, and the 'no CPU fallback' contract."""
import ctypes
import os
import re

import pytest
import torch

from oracle import cases

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from advb200 import _lib

    header = open(os.path.join(ROOT, "include", "advb200.h")).read()
    declared = set(re.findall(r"ADVB_API [\w\s\*]+?(advb_\w+)\(", header))
    assert declared, "no prototypes parsed"
    lib = _lib.load()
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    for name in declared:
        assert isinstance(getattr(lib, name), ctypes._CFuncPtr)
    assert lib.advb_version() == 202


def test_struct_layout_matches_header():
    from advb200 import _lib

    assert ctypes.sizeof(_lib.TensorRef) == 24
    assert ctypes.sizeof(_lib.AttackDesc) == 12 * 4 + 4 + 4 + 8  # + targeted, norm (the former padding), target_labels pointer
    assert _lib.AttackDesc.target_labels.offset == 56 and _lib.AttackDesc.targeted.offset == 48 and _lib.AttackDesc.norm.offset == 52
    assert ctypes.sizeof(_lib.ModelDesc) == 6 * 4 + 8


def test_frontend_tables_match_torchaudio():
    torchaudio = pytest.importorskip("torchaudio")
    from advb200 import frontends

    lf = torchaudio.transforms.LFCC(sample_rate=16000, n_lfcc=80,
                                    speckwargs={"n_fft": 512, "win_length": 400, "hop_length": 160})
    mf = torchaudio.transforms.MFCC(sample_rate=16000, n_mfcc=80,
                                    melkwargs={"n_fft": 512, "win_length": 400, "hop_length": 160})
    for ref, mine in ((lf, frontends.LFCC()), (mf, frontends.MFCC())):
        a, b = ref.state_dict(), mine.state_dict()
        assert list(a) == list(b)
        for k in a:
            assert torch.equal(a[k], b[k]), k


def test_lcnn_holder_has_reference_state_dict_keys():
    holder = cases.build_holder("lcnn", "lfcc")
    keys = list(holder.state_dict())
    assert "m_transform.6.weight" in keys and "m_transform.9.running_var" in keys
    assert "m_before_pooling.1.l_blstm.weight_hh_l0_reverse" in keys and "frontend.filter_mat" in keys
    assert sum(p.numel() for p in holder.parameters()) == 467425
    assert type(holder).__name__ == "LCNN"


def test_specrnet_holder_has_reference_state_dict_keys():
    holder = cases.build_holder("specrnet", "mfcc")
    keys = list(holder.state_dict())
    for k in ("first_bn.running_var", "block0.0.conv_downsample.weight", "block2.0.bn1.weight", "block4.0.conv2.bias",
              "fc_attention4.0.weight", "bn_before_gru.bias", "gru.weight_ih_l1_reverse", "fc2_gru.bias",
              "frontend.MelSpectrogram.mel_scale.fb"):
        assert k in keys, k
    assert "block4.0.conv_downsample.weight" not in keys
    assert sum(p.numel() for p in holder.parameters()) == 277963
    assert type(holder).__name__ == "SpecRNet"


def test_rawnet3_holder_has_reference_state_dict_keys():
    from advb200.models import get_model
    from advb200.models.rawnet3 import RawNet3

    holder = get_model("rawnet3", {}, "cpu")
    keys = list(holder.state_dict())
    for k in ("preprocess.0.flipped_filter", "preprocess.1.weight", "conv1.filterbank.low_hz_", "conv1.filterbank.band_hz_",
              "conv1.filterbank.window_", "conv1.filterbank.n_", "bn1.running_var", "layer1.residual.0.weight",
              "layer1.convs.6.weight", "layer2.bns.3.running_mean", "layer3.afms.alpha", "layer3.afms.fc.bias", "layer4.weight",
              "attention.0.weight", "attention.2.running_var", "attention.3.bias", "bn5.weight", "fc6.bias", "bn6.weight"):
        assert k in keys, k
    assert "layer2.residual.0.weight" not in keys  # identity residual when in == out (rawnet3.py:234-240)
    sd = holder.state_dict()
    assert tuple(sd["layer1.convs.0.weight"].shape) == (128, 128, 3) and tuple(sd["attention.0.weight"].shape) == (128, 4608, 1)
    assert tuple(sd["conv1.filterbank.low_hz_"].shape) == (128, 1) and tuple(sd["conv1.filterbank.n_"].shape) == (1, 125)
    assert sum(p.numel() for p in holder.parameters()) == 15496197  # the reference's prepare_model() (measured in the build container)
    assert type(holder).__name__ == "RawNet3"
    with pytest.raises(ValueError):
        RawNet3(encoder_type="ASP")  # only the reference's prepare_model() configuration is built natively
    with pytest.raises(RuntimeError, match="no CPU path"):
        holder(torch.rand(1, 16000))


def test_rawnet3_sinc_filter_restatement_properties():
    """The restated asteroid ParamSincFB (oracle side): 128 cosine (even) + 128 sine (odd) filters of 251 taps, unit centre
    tap for the cosine half, zero for the sine half."""
    from oracle import rawnet3 as orn

    holder = cases.build_holder("rawnet3", "none")
    f = orn.sinc_filters(holder.state_dict())[:, 0]
    assert tuple(f.shape) == (256, 251)
    assert torch.allclose(f[:128], f[:128].flip(1)) and torch.allclose(f[128:], -f[128:].flip(1))
    assert torch.allclose(f[:128, 125], torch.ones(128)) and torch.equal(f[128:, 125], torch.zeros(128))


def test_attack_api_surface_and_no_cpu_fallback():
    from advb200 import aa
    from advb200 import torchattacks as ta

    holder = cases.build_holder("lcnn", "lfcc")
    cls, params = aa.AttackEnum["PGD_eps001"].value
    atk = cls(holder, **params)
    assert (atk.eps, atk.steps, atk.alpha, atk.random_start, atk.attack) == (0.001, 10, 2 / 255, True, "PGD")
    atk.set_training_mode(model_training=True, batchnorm_training=False)
    assert "PGD(" in str(atk) and "eps=0.001" in str(atk)
    x = torch.rand(2, 16000)
    y = torch.tensor([0, 1])
    for a in (atk, ta.FGSM(holder, eps=0.005), ta.PGDL2(holder, eps=0.1, steps=2)):
        with pytest.raises(RuntimeError, match="no CPU path"):
            a(x, y)
    with pytest.raises(RuntimeError, match="no CPU path"):
        holder(x)
    with pytest.raises(NotImplementedError):
        atk.set_training_mode(True, True)
    # targeted modes (attack.py:60-108): by-function works; the least-likely / random pickers fail on a 1-logit model, as in
    # the reference (attack.py:280: list(range(1)).remove(label))
    atk.set_mode_targeted_by_function(lambda images, labels: 1 - labels)
    assert atk.get_mode() == "targeted" and torch.equal(atk._get_target_label(x, y), 1 - y)
    atk.set_mode_targeted_least_likely()
    with pytest.raises(ValueError):
        atk._get_target_label(x, y)
    atk.set_mode_default()
    with pytest.raises(ValueError, match="not supported"):
        ta.FAB(holder, n_classes=2).set_mode_targeted_by_function()


def test_reference_staging_manifest():
    """oracle/_ref (when staged) holds byte-identical copies of the reference files: sha256 per file in MANIFEST.json."""
    import hashlib
    import json

    from oracle import ref

    if not os.path.isdir(ref.STAGED):
        pytest.skip("oracle/_ref not staged (run python -m oracle.make_ref in the build container)")
    man = json.load(open(os.path.join(ref.STAGED, "MANIFEST.json")))["files"]
    assert "adversarial_attacks/torchattacks/attacks/pgd.py" in man and "src/models/lcnn.py" in man
    for rel, sha in man.items():
        path = os.path.join(ref.STAGED, rel)
        assert hashlib.sha256(open(path, "rb").read()).hexdigest() == sha, rel
        if os.path.exists(os.path.join("/root/reference", rel)):
            assert open(path, "rb").read() == open(os.path.join("/root/reference", rel), "rb").read(), rel


def test_install_swaps_reference_namespace():
    import sys

    import advb200
    from advb200 import torchattacks as native

    saved = {k: sys.modules.get(k) for k in ("adversarial_attacks", "adversarial_attacks.torchattacks")}
    try:
        advb200.install()
        from adversarial_attacks import torchattacks  # the import line of src/aa/aa_types.py:2

        assert torchattacks.PGD is native.PGD and torchattacks.FGSM is native.FGSM
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
