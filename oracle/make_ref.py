"""Stage the UNMODIFIED reference files this path needs from /root/reference into oracle/_ref/ (test infrastructure).

    python -m oracle.make_ref            (build container only; __graft_entry__.build() calls it when /root/reference exists)

The reference is pure Python (no setup.py / pyproject, nothing to compile), so "building" it means making its own
modules importable where /root/reference does not exist: oracle/_ref/ is git-ignored (reference sources never enter
the history) but NOT gpurun-ignored, so it travels to the GPU box like a built .so.  Files are byte-for-byte copies;
MANIFEST.json records the sha256 of every staged file so that `--impl reference` / `cpu_baseline.kind == "reference"`
can state that the timed code is the reference's own.

Only tests/, bench.py's reference / cpu_baseline legs and __graft_entry__.smoke() may import anything from here.
"""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = "/root/reference"
REF_DST = os.path.join(HERE, "_ref")

# the hot path plus the two scripts whose inner loops call it
FILES = [
    "evaluate_models_on_adversarial_attacks.py",
    "train_models_on_adversarial_attacks.py",
    "src/__init__.py",
    "src/frontends.py",
    "src/metrics.py",
    "src/utils.py",
    "src/trainer.py",
    "src/aa/__init__.py",
    "src/aa/aa_types.py",
    "src/aa/aa_trainer_types.py",
    "src/aa/utils.py",
    "src/aa/qualitative/__init__.py",
    "src/aa/qualitative/attacks_analysis.py",
    "src/models/__init__.py",
    "src/models/models.py",
    "src/models/lcnn.py",
    "src/models/specrnet.py",
    "src/models/rawnet3.py",
    "configs/training/lcnn.yaml",
    "configs/training/specrnet.yaml",
    "configs/training/rawnet3.yaml",
    "configs/aa_evaluation/lcnn.yaml",
    "configs/aa_evaluation/specrnet.yaml",
    "configs/aa_evaluation/rawnet3.yaml",
]
DIRS = ["adversarial_attacks/torchattacks"]  # the vendored torchattacks 3.2.7 (26 attack files; __init__ imports them all)


def _sha(path):
    h = hashlib.sha256()
    with open(path, "rb") as f:
        h.update(f.read())
    return h.hexdigest()


def stage(verbose: bool = False) -> str:
    if not os.path.isdir(REF_SRC):
        raise RuntimeError(f"{REF_SRC} is not present: oracle/_ref can only be staged in the build container")
    files = list(FILES)
    for d in DIRS:
        for root, _, names in os.walk(os.path.join(REF_SRC, d)):
            for n in sorted(names):
                if n.endswith(".py"):
                    files.append(os.path.relpath(os.path.join(root, n), REF_SRC))
    manifest = {}
    for rel in files:
        src, dst = os.path.join(REF_SRC, rel), os.path.join(REF_DST, rel)
        if not os.path.exists(src):
            continue
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        if not os.path.exists(dst) or _sha(dst) != _sha(src):
            shutil.copyfile(src, dst)
        manifest[rel] = _sha(dst)
    with open(os.path.join(REF_DST, "MANIFEST.json"), "w") as f:
        json.dump({"source": "piotrkawa/audio-deepfake-adversarial-attacks (unmodified copies)", "files": manifest}, f, indent=1)
    if verbose:
        print(f"staged {len(manifest)} reference files into {REF_DST}")
    return REF_DST


if __name__ == "__main__":
    stage(verbose=True)
    sys.exit(0)
