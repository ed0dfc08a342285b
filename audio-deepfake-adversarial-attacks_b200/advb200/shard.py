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


def max_over_ranks(value: float, device=None, group=None) -> float:
    """Slowest rank's value (multi-GPU timings are reported as the max over ranks)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())
