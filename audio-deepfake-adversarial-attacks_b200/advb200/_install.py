"""Drop-in installation: make the reference's own ``from adversarial_attacks import torchattacks``
(src/aa/aa_types.py:2) resolve to the native classes, without editing the reference."""
import sys
import types


def install():
    from . import torchattacks as native

    pkg = sys.modules.get("adversarial_attacks")
    if pkg is None:
        pkg = types.ModuleType("adversarial_attacks")
        pkg.__path__ = []  # namespace-like package
        sys.modules["adversarial_attacks"] = pkg
    previous = sys.modules.get("adversarial_attacks.torchattacks")
    if previous is not None and previous is not native:
        sys.modules["advb200.reference_torchattacks"] = previous  # keep the original reachable for A/B runs
    sys.modules["adversarial_attacks.torchattacks"] = native
    pkg.torchattacks = native
    return native
