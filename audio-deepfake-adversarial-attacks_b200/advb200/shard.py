"""Clip sharding over ranks (one process per GPU) — SURVEY.md §8(e).

Clips are independent, so the attack path shards on the batch dimension with NO collective on the data path: rank r
attacks clips ``[lo_r, hi_r)`` with its own engine handle and replicated weights.  ``torch.distributed`` (NCCL on GPUs,
gloo in the CPU tests) is used only for what the reference's single-process loop does after the attack: gathering the
per-clip scores / labels for accuracy and EER (evaluate_models_on_adversarial_attacks.py:261-298) and, in the benchmark,
the max-over-ranks timing.  The reference's own multi-GPU mode is ``nn.DataParallel`` (batch scatter inside every forward,
evaluate_models_on_adversarial_attacks.py:163-167); like it, each shard computes its own dB floor.
"""
from typing import Tuple

import torch
import torch.distributed as dist


def shard_bounds(n_clips: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous split of ``n_clips`` over ``world`` ranks; the first ``n_clips % world`` ranks get one extra clip."""
    if not 0 <= rank < world:
        raise ValueError(f"rank {rank} outside world of {world}")
    base, rem = divmod(n_clips, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard(t: torch.Tensor, rank: int, world: int) -> torch.Tensor:
    lo, hi = shard_bounds(t.shape[0], rank, world)
    return t[lo:hi]


def gather_rows(t: torch.Tensor, n_total: int, group=None) -> torch.Tensor:
    """All ranks receive the row-wise concatenation (in rank order) of every rank's shard of an ``n_total``-row tensor.
    Ragged shards are padded to the largest shard for the collective and trimmed afterwards."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return t
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    sizes = [hi - lo for lo, hi in (shard_bounds(n_total, r, world) for r in range(world))]
    if t.shape[0] != sizes[rank]:
        raise ValueError(f"rank {rank} holds {t.shape[0]} rows, expected {sizes[rank]}")
    pad = max(sizes)
    buf = t.new_zeros((pad,) + tuple(t.shape[1:]))
    buf[: t.shape[0]] = t
    parts = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(parts, buf, group=group)
    return torch.cat([p[:n] for p, n in zip(parts, sizes)], dim=0)


def gather_evaluation(logits: torch.Tensor, labels: torch.Tensor, n_total: int, group=None) -> dict:
    """End of a sharded evaluation (evaluate_models_on_adversarial_attacks.py:236-298): every rank contributes the logits of its
    (attacked) clips and their labels; ONE all_gather of (score, label) rows (NCCL on GPUs), then the reference's host arithmetic on
    the gathered rows - labels ``(sigmoid(o) + .5).int()`` (:238), accuracy in percent (:262-265), EER of ``calculate_eer(1 - y,
    sigmoid(o))`` (:282-288, src/metrics.py:9-14).  Every rank returns the same dict, equal to the unsharded evaluation."""
    import numpy as np
    from scipy.interpolate import interp1d
    from scipy.optimize import brentq
    from sklearn.metrics import roc_curve

    rows = torch.stack([torch.sigmoid(logits.flatten().float()), labels.flatten().to(logits.device).float()], dim=1)
    rows = gather_rows(rows, n_total, group).cpu()
    score, y = rows[:, 0].numpy(), rows[:, 1].numpy().astype(np.int64)
    pred = (rows[:, 0] + 0.5).int().numpy()
    out = {"clips": int(y.shape[0]), "accuracy": float((pred == y).mean() * 100.0), "eer": None}
    if 0 < y.sum() < y.shape[0]:  # roc_curve needs both classes
        fpr, tpr, _ = roc_curve(1 - y, -score)
        out["eer"] = float(brentq(lambda v: 1.0 - v - interp1d(fpr, tpr)(v), 0.0, 1.0))
    return out


def max_over_ranks(value: float, device=None, group=None) -> float:
    """Slowest rank's value (multi-GPU timings are reported as the max over ranks)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())
