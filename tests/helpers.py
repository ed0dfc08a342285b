"""Shared test helpers (oracle side = checker; product side = advb200 through the C ABI)."""
import os

import numpy as np
import torch

from oracle import attacks as oatk
from oracle import cases, synth
from oracle import lcnn as olcnn

ORACLE_FWD = {"lcnn": olcnn.forward}


def load_golden(name):
    return np.load(os.path.join(cases.GOLDEN_DIR, name + ".npz"))


def case_setup(name):
    case = cases.CASES[name]
    x, y = cases.case_inputs(case)
    fwd = ORACLE_FWD[case["model"]]
    holder, state = cases.build_state(case["model"], case["frontend"], calibrate_on=x, forward_fn=fwd)
    return case, x, y, holder, state, fwd


def reference_start(case, attack, x, eps):
    """The random start the reference drew under torch.manual_seed(2000+cfg_id) (oracle/make_golden.py)."""
    torch.manual_seed(2000 + case["cfg_id"])
    if attack == "pgd":
        return torch.empty_like(x).uniform_(-eps, eps)
    delta = torch.empty_like(x).normal_()
    n = delta.view(x.size(0), -1).norm(p=2, dim=1).view(x.size(0), 1)
    r = torch.zeros_like(n).uniform_(0, 1)
    delta *= r / n * eps
    return delta


def oracle_attack(name, attack, x, y, state, fwd, case):
    p = cases.ATTACKS[attack]
    model_fn = lambda v: fwd(v, state)  # noqa: E731
    if attack == "fgsm":
        return oatk.fgsm(model_fn, x, y, p["eps"])
    if attack == "pgd":
        return oatk.pgd(model_fn, x, y, p["eps"], p["alpha"], p["steps"], noise=reference_start(case, "pgd", x, p["eps"]))
    delta = reference_start(case, "pgdl2", x, p["eps"])
    start = torch.clamp(x + delta, 0, 1)
    return oatk.pgdl2(model_fn, x, y, p["eps"], p["alpha"], p["steps"], start=start)


def rel_err(a, b):
    a, b = a.double(), b.double()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def cosine(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return (a @ b / (a.norm() * b.norm()).clamp_min(1e-30)).item()


def load_holder_state(holder, state, device):
    holder.load_state_dict(state)
    return holder.to(device)
