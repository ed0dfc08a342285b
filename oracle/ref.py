"""Import the reference's own modules (test infrastructure: tests/, bench.py reference / cpu_baseline legs only).

Root = oracle/_ref (staged by oracle/make_ref.py; travels to the GPU box) or, in the build container, /root/reference.
The reference needs three things that are absent from this image:
  * `asteroid_filterbanks` (RawNet3's sinc filterbank)        -> oracle/third_party restatement on sys.path
  * `src.datasets.*` (soundfile / sox / removed torchaudio)    -> a stub module exposing `DetectionDataset`
  * a writable CWD (the evaluate script creates `logs/` at import)
"""
import contextlib
import importlib
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
STAGED = os.path.join(HERE, "_ref")
THIRD_PARTY = os.path.join(HERE, "third_party")


def root():
    """Directory holding the reference tree, or None."""
    if os.path.exists(os.path.join(STAGED, "src", "frontends.py")):
        return STAGED
    if os.path.exists("/root/reference/src/frontends.py"):
        return "/root/reference"
    return None


def available() -> bool:
    return root() is not None


def activate():
    """Put the reference (and the asteroid restatement) on sys.path and stub its dataset package.  Idempotent."""
    r = root()
    if r is None:
        raise RuntimeError("reference not available: run `python -m oracle.make_ref` in the build container")
    for p in (THIRD_PARTY, r):
        if p not in sys.path:
            sys.path.insert(0, p)
    if "src.datasets.detection_dataset" not in sys.modules:
        import src  # noqa: F401  (the reference's package)

        pkg = types.ModuleType("src.datasets")
        pkg.__path__ = []
        mod = types.ModuleType("src.datasets.detection_dataset")

        class DetectionDataset:  # only the attribute generate_attacks() may touch (evaluate_...py:230-233)
            @staticmethod
            def wavefake_preprocessing_on_batch(batch_x, batch_sr):
                return batch_x, batch_sr

        mod.DetectionDataset = DetectionDataset
        pkg.detection_dataset = mod
        sys.modules["src.datasets"] = pkg
        sys.modules["src.datasets.detection_dataset"] = mod
    return r


@contextlib.contextmanager
def _cwd(path):
    prev = os.getcwd()
    os.makedirs(path, exist_ok=True)
    os.chdir(path)
    try:
        yield
    finally:
        os.chdir(prev)


def import_script(name: str, workdir: str):
    """Import `evaluate_models_on_adversarial_attacks` / `train_models_on_adversarial_attacks` from a writable CWD."""
    activate()
    with _cwd(workdir):
        return importlib.import_module(name)


def torchattacks():
    """The reference's vendored torchattacks package (adversarial_attacks/torchattacks), unmodified."""
    activate()
    native = sys.modules.get("adversarial_attacks.torchattacks")
    if native is not None and getattr(native, "__advb200__", False):
        kept = sys.modules.get("advb200.reference_torchattacks")
        if kept is not None:
            return kept
        raise RuntimeError("advb200.install() replaced adversarial_attacks.torchattacks before the reference was imported")
    from adversarial_attacks import torchattacks as ta

    return ta


def model(kind: str, frontend: str = "lfcc", state=None):
    """Reference nn.Module on CPU: LCNN / SpecRNet / RawNet3 built the way src/models/models.py:6-18 does."""
    activate()
    if kind == "lcnn":
        from src.models.lcnn import LCNN

        m = LCNN(device="cpu", input_channels=1, frontend_algorithm=[frontend])
    elif kind == "specrnet":
        from src.models.specrnet import SpecRNet, get_config

        m = SpecRNet(get_config(1), device="cpu", input_channels=1, frontend_algorithm=[frontend])
    elif kind == "rawnet3":
        from src.models.rawnet3 import prepare_model

        m = prepare_model()
    else:
        raise ValueError(kind)
    if state is not None:
        res = m.load_state_dict(state, strict=True)
        assert not res.missing_keys and not res.unexpected_keys
    return m
