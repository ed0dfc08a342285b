#!/usr/bin/env python
"""One small call of every entry point the per-model gradient runs of tools/sanitize.sh do not reach: FAB (L-inf and L2, with a
restart), CW, PGDL2, a targeted FGSM, the fused min-max call, projection_linf / projection_l2 rows and the mel_spec frontend.
Run under compute-sanitizer (tools/sanitize.sh)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "audio-deepfake-adversarial-attacks_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402

import helpers  # noqa: E402
from advb200 import aa, engine, frontends  # noqa: E402
from advb200 import torchattacks as ta  # noqa: E402

dev = torch.device("cuda:0")
case, x, y, holder, state, fwd = helpers.case_setup("lcnn_lfcc_t16000_margin")
holder = helpers.load_holder_state(holder, state, dev)
xd, yd = x.to(dev), y.to(dev)
out = {}
for name, atk in (("fab_linf", ta.FAB(holder, norm="Linf", eps=0.3, steps=2, eta=10.0, n_classes=2, n_restarts=2)),
                  ("fab_l2", ta.FAB(holder, norm="L2", eps=2.0, steps=2, n_classes=2, n_restarts=2)),
                  ("cw", ta.CW(holder, c=1e4, steps=3, lr=5e-4)),
                  ("pgdl2", ta.PGDL2(holder, eps=0.1, alpha=0.2, steps=2)),
                  ("pgd", ta.PGD(holder, eps=0.001, steps=3))):
    atk.set_training_mode(True, False)
    out[name] = float((atk(xd, yd) - xd).abs().max())
tg = ta.FGSM(holder, eps=0.005)
tg.set_mode_targeted_by_function(lambda images, labels: 1 - labels)
out["fgsm_targeted"] = float((tg(xd, yd) - xd).abs().max())
raw = 0.1 * torch.randn(4, 16000, device=dev) + 0.02
mm = ta.FGSM(holder, eps=0.005)
mm.set_training_mode(True, False)
out["fgsm_minmax"] = float((aa.attack_minmax(mm, raw, yd) - raw).abs().max())
g = torch.Generator("cpu").manual_seed(1)
t, w = torch.rand(3, 4001, generator=g).to(dev), (torch.randn(3, 4001, generator=g) * 1e-3).to(dev)
b = (w * torch.rand(3, 4001, generator=g).to(dev)).sum(1)
out["proj_linf"] = float(engine.projection_linf(t, w, b).abs().max())
out["proj_l2"] = float(engine.projection_l2(t, w, b).abs().max())
out["mel_spec"] = float(frontends.prepare_mel_scale_vector(xd)[:, 0].mean())
torch.cuda.synchronize()
print("extras ok", out)
