"""Golden for the `mel_spec` frontend, produced by the UNMODIFIED reference function on CPU.

    python -m oracle.make_golden_melspec            (build container; seconds)

Calls src.frontends.prepare_mel_scale_vector (src/frontends.py:53-58) from the staged reference on seeded clips: B = 2 at
T = 16 000 and one clip of the reference-native odd length 16 150 (a frame count that is not a multiple of the two frames a
warp packs).  Stores the outputs and the filterbank the reference's MEL_SCALE_FN holds.  Test infrastructure.
"""
import os

import numpy as np
import torch

from . import cases, ref, synth


def main():
    ref.activate()
    import src.frontends as rf  # the reference's own module

    torch.set_num_threads(os.cpu_count() or 1)
    out = {"fb": rf.MEL_SCALE_FN.fb.numpy()}
    for tag, cfg_id, B, T in (("t16000", 31, 2, 16000), ("t16150", 32, 1, 16150)):
        x, _ = synth.clips(cfg_id, B, T)
        with torch.no_grad():
            y = rf.prepare_mel_scale_vector(x)
        out[f"{tag}_out"] = y.numpy()
        out[f"{tag}_x_sum"] = np.array(x.double().sum().item())
        print(tag, tuple(y.shape), float(y[:, 0].mean()))
    path = os.path.join(cases.GOLDEN_DIR, "mel_spec.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
