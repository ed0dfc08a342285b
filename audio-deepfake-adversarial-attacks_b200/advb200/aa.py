"""AA glue — same names as the reference's src/aa/aa_types.py:5-24 and src/aa/utils.py:4-14."""
from enum import Enum

from . import torchattacks
from .engine import revert_minmax, to_minmax  # noqa: F401  (GPU kernels; same return convention)


class AttackEnum(Enum):
    PGD = (torchattacks.PGD, {"eps": 0.0005, "steps": 10})
    PGD_eps00075 = (torchattacks.PGD, {"eps": 0.00075, "steps": 10})
    PGD_eps001 = (torchattacks.PGD, {"eps": 0.001, "steps": 10})

    PGDL2 = (torchattacks.PGDL2, {"eps": 0.1, "steps": 10})
    PGDL2_eps15 = (torchattacks.PGDL2, {"eps": 0.15, "steps": 10})
    PGDL2_eps20 = (torchattacks.PGDL2, {"eps": 0.20, "steps": 10})

    FGSM = (torchattacks.FGSM, {"eps": 0.0005})
    FGSM_eps00075 = (torchattacks.FGSM, {"eps": 0.00075})
    FGSM_eps001 = (torchattacks.FGSM, {"eps": 0.001})

    FAB = (torchattacks.FAB, {"n_classes": 2, "eta": 10})
    FAB_eta20 = (torchattacks.FAB, {"n_classes": 2, "eta": 20})
    FAB_eta30 = (torchattacks.FAB, {"n_classes": 2, "eta": 30})

    NO_ATTACK = (None, {})


def attack_minmax(atk, x, y):
    """``revert_minmax(atk(to_minmax(x)), ...)`` - the three lines every call site of the reference wraps around an attack
    (evaluate_models_on_adversarial_attacks.py:219-221, src/trainer.py:425-427) - as ONE native call for FGSM / PGD / PGDL2
    (SURVEY.md §8 f2); FAB and CW, whose Python shims issue several engine calls, go through the three GPU kernels."""
    if isinstance(atk, (torchattacks.FGSM, torchattacks.PGD, torchattacks.PGDL2)):
        atk._fused_minmax = True
        try:
            return atk(x, y)
        finally:
            atk._fused_minmax = False
    x01, mn, mx = to_minmax(x.to(atk.device))
    return revert_minmax(atk(x01, y), mn, mx)
