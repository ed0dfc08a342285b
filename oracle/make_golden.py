"""Generate tests/golden/*.npz by running the UNMODIFIED reference (/root/reference) on CPU.

Run in the build container only:   python -m oracle.make_golden
The GPU box has no /root/reference; it consumes the committed fixtures.  Test infrastructure.
"""
import os
import sys
import warnings

import numpy as np
import torch

warnings.filterwarnings("ignore")
sys.path.insert(0, "/root/reference")
# third-party dependency of the reference that is not installed here (restated, see its docstring)
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "third_party"))

from . import cases, synth  # noqa: E402
from . import lcnn as olcnn  # noqa: E402
from . import rawnet3 as orn  # noqa: E402
from . import specrnet as ospec  # noqa: E402


def reference_model(case, state):
    if case["model"] == "lcnn":
        from src.models.lcnn import LCNN

        m = LCNN(device="cpu", input_channels=1, frontend_algorithm=[case["frontend"]])
    elif case["model"] == "specrnet":
        from src.models.specrnet import SpecRNet, get_config

        m = SpecRNet(get_config(1), device="cpu", input_channels=1, frontend_algorithm=[case["frontend"]])
    elif case["model"] == "rawnet3":
        from src.models.rawnet3 import prepare_model

        m = prepare_model()
    else:
        raise NotImplementedError(case["model"])
    missing = m.load_state_dict(state, strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    return m


def main():
    from adversarial_attacks import torchattacks

    torch.set_num_threads(8)
    os.makedirs(cases.GOLDEN_DIR, exist_ok=True)
    only = set(sys.argv[1:])  # optional: regenerate just the named cases
    for name, case in cases.CASES.items():
        if only and name not in only:
            continue
        x, y = cases.case_inputs(case)
        fwd = {"lcnn": olcnn.forward, "specrnet": ospec.forward, "rawnet3": orn.forward}[case["model"]]
        _, state = cases.build_state(case["model"], case["frontend"], calibrate_on=x, forward_fn=fwd,
                                     margin=case.get("margin", 0.0))
        ref = reference_model(case, state)
        ref.eval()
        out = {"digest": np.array(synth.state_digest(state)), "x_sum": np.array(x.double().sum().item())}
        with torch.no_grad():
            out["logits"] = ref(x).numpy()
            if hasattr(ref, "frontend"):
                out["frontend"] = ref.frontend(x).numpy()
        # CE gradient exactly as the attacks build it (fgsm.py:43-57)
        xr = x.clone().requires_grad_(True)
        ref.train()
        for m in ref.modules():
            if "BatchNorm" in type(m).__name__ or "Dropout" in type(m).__name__:
                m.eval()
        o = ref(xr)
        cost = torch.nn.CrossEntropyLoss()(torch.cat([-o, o], dim=1), y)
        out["grad"] = torch.autograd.grad(cost, xr)[0].numpy()
        ref.eval()

        for an in case.get("attacks", cases.DEFAULT_ATTACKS):
            ap = cases.ATTACKS[an]
            if an.startswith("fab"):
                atk = torchattacks.FAB(ref, norm=ap.get("norm", "Linf"), eps=ap["eps"], steps=ap["steps"], eta=ap["eta"],
                                       alpha_max=ap["alpha_max"], beta=ap["beta"], n_classes=2)
            elif an.startswith("cw"):
                atk = torchattacks.CW(ref, c=ap["c"], kappa=ap["kappa"], steps=ap["steps"], lr=ap["lr"])
            elif an == "fgsm":
                atk = torchattacks.FGSM(ref, eps=ap["eps"])
            elif an == "pgd":
                atk = torchattacks.PGD(ref, eps=ap["eps"], alpha=ap["alpha"], steps=ap["steps"], random_start=True)
            else:
                atk = torchattacks.PGDL2(ref, eps=ap["eps"], alpha=ap["alpha"], steps=ap["steps"], random_start=True)
            atk.set_training_mode(model_training=True, batchnorm_training=False)
            # reproduce the reference's own random start: seed torch's global RNG, record what it draws
            torch.manual_seed(2000 + case["cfg_id"])
            ref.eval()  # generate_attacks() calls model.eval() before every batch (evaluate_...py:212)
            xa = atk(x, y)
            ref.eval()  # Attack.__call__ leaves train() on; the clean re-inference runs in eval mode (:236)
            with torch.no_grad():
                la = ref(xa)
            out[f"{an}_adv"] = xa.numpy() if case["T"] <= 16000 else np.zeros(0, np.float32)
            out[f"{an}_delta_linf"] = (xa - x).abs().amax(dim=1).numpy()
            out[f"{an}_delta_l2"] = (xa - x).norm(p=2, dim=1).numpy()
            out[f"{an}_sign_bits"] = np.packbits((xa > x).numpy())
            out[f"{an}_logits_adv"] = la.numpy()
        path = os.path.join(cases.GOLDEN_DIR, name + ".npz")
        np.savez_compressed(path, **out)
        first = case.get("attacks", cases.DEFAULT_ATTACKS)[-1]
        print(name, "logits", out["logits"].ravel(), first, "logits", out[f"{first}_logits_adv"].ravel(),
              os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
