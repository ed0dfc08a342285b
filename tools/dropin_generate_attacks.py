#!/usr/bin/env python
"""Run the reference's own `generate_attacks()` (evaluate_models_on_adversarial_attacks.py:133-298), UNEDITED, on a synthetic
dataset -- with the reference's vendored torchattacks on CPU (`--impl reference`) or with `advb200.install()` on the GPU
(`--impl native`) -- and print one JSON line with the per-clip results.

This is BASELINE.json configs[0] (FGSM eps=0.005 on LCNN+LFCC, batch 8, 64 000-sample clips) and SURVEY.md §8(a) row a15: the
model that reaches the attack is the reference's `src.models.lcnn.LCNN`, loaded by the reference's `load_model` from a
checkpoint file and wrapped in `nn.DataParallel` (evaluate_...py:162-169).  Recipe: SURVEY.md §8(c).

Test infrastructure (imports oracle/): used by tests/test_gpu_dropin.py and as bench.py's config-1 line.
"""
import argparse
import json
import logging
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "audio-deepfake-adversarial-attacks_b200"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--impl", choices=["native", "reference"], required=True)
    ap.add_argument("--attack", default="FGSM", help="AttackEnum member name (class is taken from it)")
    ap.add_argument("--eps", type=float, default=0.005, help="override of the preset's eps (configs[0] uses 0.005)")
    ap.add_argument("--model", default="lcnn", choices=["lcnn", "specrnet"])
    ap.add_argument("--clips", type=int, default=16)
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--samples", type=int, default=64000)
    ap.add_argument("--cfg-id", type=int, default=1)
    args = ap.parse_args()

    if args.impl == "reference":
        os.environ["CUDA_VISIBLE_DEVICES"] = ""  # the reference's CPU path (there is no --cpu flag on that script)
    import torch

    if args.impl == "native":
        import advb200

        advb200.install()  # BEFORE the reference imports AttackEnum (src/aa/aa_types.py:2 binds the classes at import)
    import yaml

    from oracle import cases, ref, synth
    from oracle.make_golden_cfg import BIAS_KEY

    work = tempfile.mkdtemp(prefix="advb_dropin_")
    ev = ref.import_script("evaluate_models_on_adversarial_attacks", work)
    device = "cuda" if torch.cuda.is_available() else "cpu"
    if args.impl == "native":
        assert device == "cuda", "the native engine has no CPU path"
        from adversarial_attacks import torchattacks as bound

        assert getattr(bound, "__advb200__", False) and ev.AttackEnum.FGSM.value[0].__module__.startswith("advb200")

    # seeded calibrated checkpoint FILE (white-box: target and attack model load the same weights)
    x, y = synth.clips(args.cfg_id, args.clips, args.samples)
    _, state = cases.build_state(args.model, "lfcc")
    gold_path = os.path.join(cases.GOLDEN_DIR, "cfg1_lcnn_fgsm_b8.npz")
    if args.model == "lcnn" and os.path.exists(gold_path):
        import numpy as np

        state[BIAS_KEY["lcnn"]] = torch.from_numpy(np.load(gold_path)["bias"])
    ckpt = os.path.join(work, "ckpt.pth")
    torch.save(state, ckpt)
    cfg = yaml.safe_load(open(os.path.join(ref.root(), "configs", "training", f"{args.model}.yaml")))
    cfg["checkpoint"]["path"] = ckpt

    class Synthetic(torch.utils.data.Dataset):  # item layout of DetectionDataset(return_label, return_meta): evaluate_...py:211
        def __len__(self):
            return args.clips

        def __getitem__(self, i):
            return x[i], 16000, int(y[i]), ("synthetic", f"clip_{i:04d}", "val", 4.0)

    ev.get_dataset = lambda **kw: Synthetic()
    records = {}

    def on_attack_end(batch_x, batch_x_attacked, batch_y, batch_preds_label, batch_preds, batch_preds_noattack_label,
                      batch_preds_noattack, batch_metadata):
        names = batch_metadata[1]
        d = (batch_x_attacked - batch_x).float()
        for i, n in enumerate(names):
            records[n] = dict(y=int(batch_y[i]), pred=int(batch_preds_label[i]), score=float(batch_preds[i]),
                              pred_clean=int(batch_preds_noattack_label[i]), score_clean=float(batch_preds_noattack[i]),
                              linf=float(d[i].abs().max()), l2=float(d[i].norm()))

    captured = []

    class Grab(logging.Handler):
        def emit(self, record):
            msg = record.getMessage()
            if "adv_eval/eer" in msg:
                captured.append(msg)

    ev.LOGGER.addHandler(Grab())
    cls, params = ev.AttackEnum[args.attack].value
    params = dict(params)
    if args.eps > 0:
        params["eps"] = args.eps
    torch.manual_seed(42)
    t0 = time.perf_counter()
    ev.generate_attacks(datasets_paths=[None] * 3, model_config=cfg, attack_model_config=cfg, attack_method=cls,
                        attack_params=params, device=device, batch_size=args.batch, on_attack_end_callback=on_attack_end)
    if device == "cuda":
        torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    metrics = {}
    if captured:
        for part in captured[-1].split(","):
            k, v = part.rsplit(":", 1)
            metrics[k.strip().split("/")[-1]] = float(v)
    print("DROPIN " + json.dumps({"impl": args.impl, "device": device, "attack_class": f"{cls.__module__}.{cls.__name__}",
                                  "model_class": "src.models." + args.model, "params": params, "clips": len(records),
                                  "seconds_generate_attacks": dt, "metrics": metrics, "records": records}))


if __name__ == "__main__":
    main()
