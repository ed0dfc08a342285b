"""CPU oracle for the adversarial-perturbation hot path — TEST INFRASTRUCTURE ONLY.

A restatement, in plain torch CPU tensor arithmetic, of the reference's algorithm for the path
BASELINE.json names (waveform -> LFCC/MFCC -> LCNN/SpecRNet/RawNet3 -> 2-class CE -> d loss/d waveform ->
FGSM/PGD/PGDL2/FAB/CW update).  Every function cites the reference file:line it follows.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` leg may
import this package, and only as the checker / the reported CPU baseline — never on the product path
(``audio-deepfake-adversarial-attacks_b200/advb200`` fails loudly when its CUDA library is missing).

Pinning: the reference ships no tests or golden vectors for this path, so the oracle is
pinned against outputs of the reference itself, executed in the build container by ``oracle/make_golden.py``
(imports /root/reference) and committed as fixtures under ``tests/golden/``; ``tests/test_oracle_golden.py``
re-checks the restatement against them on every run.  The floating-point arithmetic is fp32 like the
reference (fp64 available through ``dtype=`` for tolerance calibration).
"""
