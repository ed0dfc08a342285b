#!/usr/bin/env python
"""RawNet3: how well-conditioned is d loss / d waveform in the REFERENCE itself?  (CPU; reference classes from oracle/_ref.)

RawNet3 takes log(|s| + 1e-6) of 256 x 6375 sinc-filter outputs (rawnet3.py:80-82).  The smallest |s| of a clip sit below the
rounding noise of a 251-tap fp32 dot product, 1 / (|s| + 1e-6) of those few outputs dominates the waveform gradient, and the
InstanceNorm backward spreads them over every sample.  This script measures it instead of asserting it:

  * the reference model in float64 gives the gradient every fp32 implementation approximates;
  * the reference's own float32 gradient is computed with 8 threads and with 1 thread (same code, different summation order);
  * cosine / sign agreement of each float32 run with float64, over all samples and over the samples that remain when the
    contribution of the ill-conditioned filter outputs (|s| < 1e-4) is removed from BOTH gradients (the gradient is linear in
    d loss / d log-feature, so masking those feature gradients before the transposed sinc convolution is exact).

Output: tests/golden/rawnet3_fp64_grad_{t16000,t64000}.npz (float64 gradient, masked float64 gradient, the reference-fp32
statistics).  tests/test_gpu_rawnet3.py holds the engine to the same standard: at least as close to float64 as the reference's
own float32 runs are.  Re-run with `python tools/rn_conditioning.py`.
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
os.environ["CUDA_VISIBLE_DEVICES"] = ""
import numpy as np  # noqa: E402
import torch  # noqa: E402

import helpers  # noqa: E402
from oracle import cases, ref  # noqa: E402

S_FLOOR = 1e-4  # filter outputs below this are "ill-conditioned": their 1/(|s|+1e-6) amplifies fp32 rounding noise >= 1e-5/1e-4


def grads(model, x, y, dtype, mask_floor=None):
    """(d CE / d x, d CE / d x with the feature gradients of outputs |s| < mask_floor zeroed, |s| tensor)."""
    m = model.double() if dtype == torch.float64 else model.float()
    xx = x.to(dtype).clone().requires_grad_(True)
    keep = {}

    def hook(mod, inp, out):  # the sinc encoder's output s (B, 256, L): rawnet3.py:80
        keep["s"] = out
        if mask_floor is not None:
            out.register_hook(lambda g: g * (keep["s"].detach().abs() >= mask_floor).to(g.dtype))

    h = m.conv1.register_forward_hook(hook)
    try:
        m.train()
        for mod in m.modules():
            if "BatchNorm" in type(mod).__name__ or "Dropout" in type(mod).__name__:
                mod.eval()
        o = m(xx)
        cost = torch.nn.CrossEntropyLoss()(torch.cat([-o, o], dim=1), y)
        (g,) = torch.autograd.grad(cost, xx)
    finally:
        h.remove()
        m.eval()
    return g.detach(), keep["s"].detach().abs()


def cos(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return float(a @ b / (a.norm() * b.norm()).clamp_min(1e-300))


def main():
    for name in ("rawnet3_t16000", "rawnet3_t64000"):
        case, x, y, holder, state, fwd = helpers.case_setup(name)
        model = ref.model("rawnet3", "none", state)
        torch.set_num_threads(8)
        g64, s_abs = grads(model, x, y, torch.float64)
        g64m, _ = grads(model, x, y, torch.float64, S_FLOOR)
        model = ref.model("rawnet3", "none", state)  # fresh float32 copy (the double() above converted in place)
        out = {"g64": g64.numpy(), "g64_masked": g64m.numpy(), "s_floor": np.array(S_FLOOR),
               "n_small": np.array(int((s_abs < S_FLOOR).sum())), "n_outputs": np.array(s_abs.numel()),
               "min_abs_s": np.array(float(s_abs.min()))}
        for threads in (8, 1):
            torch.set_num_threads(threads)
            g32, _ = grads(model, x, y, torch.float32)
            g32m, _ = grads(model, x, y, torch.float32, S_FLOOR)
            out[f"ref32_t{threads}_cos"] = np.array(cos(g32, g64))
            out[f"ref32_t{threads}_sign"] = np.array(float((g32.sign() == g64.sign()).double().mean()))
            out[f"ref32_t{threads}_cos_masked"] = np.array(cos(g32m, g64m))
            out[f"ref32_t{threads}_sign_masked"] = np.array(float((g32m.sign() == g64m.sign()).double().mean()))
            if threads == 8:
                g8 = g32
            else:
                out["ref32_t8_vs_t1_cos"] = np.array(cos(g8, g32))
        path = os.path.join(cases.GOLDEN_DIR, name.replace("rawnet3_", "rawnet3_fp64_grad_") + ".npz")
        np.savez_compressed(path, **out)
        print(name, {k: (float(v) if v.ndim == 0 else v.shape) for k, v in out.items()}, os.path.getsize(path) // 1024, "KiB", flush=True)


if __name__ == "__main__":
    main()
