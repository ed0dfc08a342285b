"""advb200 — B200-native adversarial-perturbation engine (host side).

Public surface mirrors the reference (piotrkawa/audio-deepfake-adversarial-attacks):

* ``advb200.torchattacks.{FGSM,PGD,PGDL2,FAB,CW}`` — same constructors / ``set_training_mode`` / ``__call__`` as
  ``adversarial_attacks/torchattacks`` (attack.py:14-35,132-147,308-331);
* ``advb200.models.get_model(name, config, device)`` — ``src/models/models.py:6-18``;
* ``advb200.frontends.{LFCC_FN,MFCC_FN,get_frontend}`` — ``src/frontends.py:13-50``;
* ``advb200.aa.{AttackEnum,to_minmax,revert_minmax}`` — ``src/aa/aa_types.py``, ``src/aa/utils.py``;
* ``advb200.install()`` — registers the native attacks as ``adversarial_attacks.torchattacks`` so the
  reference's unedited scripts bind to them.

All arithmetic runs in ``libadvb200.so`` (hand-written sm_100a CUDA behind the C ABI of ``include/advb200.h``).
There is no CPU fallback: calling an attack or a model without the library or without a CUDA device raises.
"""
__version__ = "0.1.0"


def install():
    from ._install import install as _install

    return _install()
