"""Native replacements of the reference's patched ``adversarial_attacks.torchattacks`` classes.

Same constructor signatures, public attributes, ``set_training_mode`` and ``__call__(images, labels)`` contract
(attack.py:14-35,132-147,308-331; fgsm.py:28; pgd.py:31-32; pgdl2.py:31); the per-iteration pipeline (frontend,
model forward, CE, input gradient, update rule) runs in libadvb200's CUDA kernels.
"""
from .attack import Attack
from .attacks import CW, FAB, FGSM, PGD, PGDL2

__version__ = "3.2.7+advb200"
__advb200__ = True  # marker: this module is the native replacement, not the vendored package
__all__ = ["Attack", "FGSM", "PGD", "PGDL2", "FAB", "CW"]
