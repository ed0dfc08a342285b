#!/usr/bin/env python
"""Drive the attack call sites of the reference's adversarial-training strategies (src/trainer.py:455-542: RANDOM, EQUAL's
in-place half batch, ONLY_ADV, ADAPTIVE) UNEDITED -- with the reference's torchattacks on CPU (`--impl reference`) or under
`advb200.install()` on the GPU (`--impl native`) -- and save what each strategy returned.  SURVEY.md §8(a) row a16.

The trainers are the reference's own classes (AdversarialGDTrainerEnum); `init_adv_attacks` builds the attacks from
AttackEnum exactly as training does, `apply_adv_attack(batch_x, batch_y)` is called with CPU labels like the training loop
(src/trainer.py:305), and Python's `random` is seeded so both sides pick the same clips / attacks.

Test infrastructure (imports oracle/): used by tests/test_gpu_dropin.py.
"""
import argparse
import os
import random
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "audio-deepfake-adversarial-attacks_b200"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--impl", choices=["native", "reference"], required=True)
    ap.add_argument("--out", required=True)
    ap.add_argument("--clips", type=int, default=8)
    ap.add_argument("--samples", type=int, default=16000)
    args = ap.parse_args()
    if args.impl == "reference":
        os.environ["CUDA_VISIBLE_DEVICES"] = ""
    import numpy as np
    import torch

    if args.impl == "native":
        import advb200

        advb200.install()
    from oracle import cases, ref, synth

    ref.activate()
    from src.aa.aa_trainer_types import AdversarialGDTrainerEnum

    device = "cuda" if torch.cuda.is_available() else "cpu"
    g = torch.Generator("cpu").manual_seed(4242)
    raw = 0.1 * torch.randn(args.clips, args.samples, generator=g) + 0.02  # RAW waveforms: the strategies min-max them
    y = torch.randint(0, 2, (args.clips,), generator=g)
    _, state = cases.build_state("lcnn", "lfcc", calibrate_on=synth.clips(11, 2, 16000)[0],
                                 forward_fn=__import__("oracle.lcnn", fromlist=["forward"]).forward)
    model = ref.model("lcnn", "lfcc", state).to(device)
    attack_model = torch.nn.DataParallel(model) if device == "cuda" else model
    out = {"y": y.numpy(), "raw": raw.numpy()}
    for name, attacks in (("ONLY_ADV", ["FGSM_eps001"]), ("EQUAL", ["FGSM_eps001"]), ("RANDOM", ["FGSM", "FGSM_eps001"]),
                          ("ADAPTIVE", ["FGSM", "FGSM_eps00075"])):
        trainer = AdversarialGDTrainerEnum[name].value(epochs=1, batch_size=args.clips, device=device)
        trainer.init_adv_attacks(attack_model, attacks)
        if args.impl == "native":
            assert all(type(a).__module__.startswith("advb200") for _, a in trainer.attacks)
        random.seed(7)
        was_training = model.training
        for rep in range(3):  # three draws: RANDOM / ADAPTIVE pick differently each time
            batch_x = raw.clone().to(device)
            got = trainer.apply_adv_attack(batch_x, y.clone())  # CPU labels, like the training loop
            trainer.update_adv_attack(np.float32(0.4 + 0.1 * rep), None, iter=rep, epoch=0)
            out[f"{name}_{rep}"] = got.detach().cpu().numpy()
        assert model.training == was_training, "the attack must restore the model's training flag"
    np.savez_compressed(args.out, **out)
    print("TRAINER_DROPIN ok", args.impl, device, sorted(k for k in out if k not in ("y", "raw")))


if __name__ == "__main__":
    main()
