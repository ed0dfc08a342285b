"""Strict multi-GPU mode, one PROCESS per rank (SURVEY.md §8(e) ii): a clip-sharded run with the cross-rank dB-floor exchange
(include/advb200.h, advb_xrank_*) reproduces the single-device run of the whole batch, per clip.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 \
        tools/strict_equiv.py [--out gpurun_out/strict_equiv.json]

Each rank uses cuda:LOCAL_RANK when the box has that many GPUs (mailbox stores travel over NVLink), else every rank shares
cuda:0 (the CUDA IPC mapping and the protocol are the same; the two contexts time-slice).  The control plane (IPC-handle
exchange, gathering the shards' results) is gloo, so the script runs on either.  Model: SpecRNet + MFCC, whose dB floor is
always active (mel filter 0 is identically zero); the second half of the batch is 6 dB quieter, so per-shard floors differ.

Rank 0 prints one JSON line (and writes --out): the per-clip differences between the whole batch on one device and
 (a) the shards under strict mode, (b) the shards with per-shard floors (the default; what nn.DataParallel does).
"""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "audio-deepfake-adversarial-attacks_b200"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

CASE = "specrnet_mfcc_t16000"
STEPS = 4


def batch(groups=2):
    """The golden case's clips, then the same clips 6 dB quieter, 12 dB quieter, ... (one group per rank: every shard has its own
    batch maximum)."""
    import helpers

    case, x, y, holder, state, _ = helpers.case_setup(CASE)
    return case, torch.cat([x * 0.5 ** g for g in range(groups)]), torch.cat([y] * groups), holder, state


def run_all(eng, holder, xs, ys, noise, n_global):
    """Logits, CE gradient and a PGD-`STEPS` result of the clips this handle is given."""
    from advb200 import torchattacks as ta

    g, logits = eng.grad(xs, ys, n_global=n_global)
    atk = ta.PGD(holder, eps=0.001, alpha=2 / 255, steps=STEPS, random_start=True)
    atk.set_training_mode(model_training=True, batchnorm_training=False)
    adv = atk.forward(xs, ys, noise=noise)
    return {"logits": logits.flatten().clone(), "grad": g.clone(), "adv": adv.clone(), "adv_logits": eng.forward(adv).flatten().clone()}


def compare(full, parts):
    d = {}
    d["logit_maxdiff"] = (full["logits"] - parts["logits"]).abs().max().item()
    d["grad_rel_err"] = ((full["grad"] - parts["grad"]).double().norm() / full["grad"].double().norm()).item()
    d["grad_sign_mismatch"] = (torch.sign(full["grad"]) != torch.sign(parts["grad"])).float().mean().item()
    d["pgd_element_mismatch"] = (full["adv"] != parts["adv"]).float().mean().item()
    d["pgd_adv_logit_maxdiff"] = (full["adv_logits"] - parts["adv_logits"]).abs().max().item()
    d["pgd_labels_equal"] = bool(torch.equal(full["adv_logits"] > 0, parts["adv_logits"] > 0))
    return d


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist.init_process_group("gloo")
    shared = torch.cuda.device_count() < world
    dev = torch.device("cuda", 0 if shared else local)
    torch.cuda.set_device(dev)

    import helpers
    from advb200 import engine, shard

    case, x, y, holder, state = batch(max(2, world))
    n = x.shape[0]
    holder = helpers.load_holder_state(holder, state, dev)
    eng = engine.engine_for(holder, n, x.shape[1])
    torch.manual_seed(77)
    noise = torch.empty_like(x).uniform_(-0.001, 0.001)
    xd, yd, nd = x.to(dev), y.to(dev), noise.to(dev)
    lo, hi = shard.shard_bounds(n, rank, world)

    full = run_all(eng, holder, xd, yd, nd, n)                         # the whole batch on this rank's device
    plain = run_all(eng, holder, xd[lo:hi], yd[lo:hi], nd[lo:hi], n)   # this rank's shard, per-shard floor (default)
    eng.enable_strict()
    strict = run_all(eng, holder, xd[lo:hi], yd[lo:hi], nd[lo:hi], n)  # this rank's shard, floor shared across the ranks
    timed_out = eng.strict_timed_out()
    eng.xrank_connect(0, 1)
    again = run_all(eng, holder, xd[lo:hi], yd[lo:hi], nd[lo:hi], n)   # disconnected: back to the per-shard result, bit for bit
    back = all(torch.equal(again[k], plain[k]) for k in plain)

    def gathered(res):
        return {k: shard.gather_rows(v.cpu(), n) for k, v in res.items()}

    g_plain, g_strict = gathered(plain), gathered(strict)
    flags = [None] * world
    dist.all_gather_object(flags, (bool(timed_out), bool(back)))
    if rank == 0:
        full = {k: v.cpu() for k, v in full.items()}
        line = {"case": CASE, "world": world, "clips": n, "devices": "one shared GPU" if shared else f"{world} GPUs",
                "pgd_steps": STEPS, "timed_out": [f[0] for f in flags], "disconnect_restores_plain": [f[1] for f in flags],
                "strict_vs_whole_batch": compare(full, g_strict), "per_shard_floor_vs_whole_batch": compare(full, g_plain)}
        print(json.dumps(line))
        if args.out:
            os.makedirs(os.path.dirname(os.path.abspath(args.out)), exist_ok=True)
            with open(args.out, "w") as f:
                json.dump(line, f, indent=1)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
