"""CPU, world_size 2 over gloo: the clip-sharding host logic of the multi-GPU path (advb200/shard.py)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from advb200 import shard


def test_shard_bounds_cover_every_clip_once():
    for n in (0, 1, 5, 8, 127, 128):
        for world in (1, 2, 3, 8):
            spans = [shard.shard_bounds(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard.shard_bounds(4, 2, 2)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_clips, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(3)
        logits = torch.randn(n_clips, 1, generator=g)          # what every rank would compute for the full batch
        labels = torch.randint(0, 2, (n_clips,), generator=g)
        mine = shard.shard(logits, rank, world) * 1.0            # this rank's clips only
        got = shard.gather_rows(mine, n_clips)
        got_y = shard.gather_rows(shard.shard(labels, rank, world), n_clips)
        assert torch.equal(got, logits) and torch.equal(got_y, labels)
        # accuracy from the gathered scores is identical on every rank and equals the unsharded value
        acc = ((got[:, 0] > 0).long() == got_y).float().mean().item()
        assert acc == ((logits[:, 0] > 0).long() == labels).float().mean().item()
        # (score, label) gather + the reference's accuracy / EER arithmetic: same result on every rank as the unsharded evaluation
        ev = shard.gather_evaluation(mine, shard.shard(labels, rank, world), n_clips)
        ref_path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "src", "metrics.py")
        if os.path.exists(ref_path):  # the reference's own calculate_eer (src/metrics.py:9-14), staged by oracle/make_ref.py
            import importlib.util

            spec = importlib.util.spec_from_file_location("ref_metrics", ref_path)
            ref_metrics = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(ref_metrics)
            y_np, score_np = labels.numpy(), torch.sigmoid(logits[:, 0]).numpy()
            _, eer_ref, _, _ = ref_metrics.calculate_eer(y=1 - y_np, y_score=score_np)
            assert abs(ev["eer"] - eer_ref) < 1e-12
        assert ev["clips"] == n_clips
        assert abs(ev["accuracy"] - ((torch.sigmoid(logits[:, 0]) + 0.5).int() == labels).float().mean().item() * 100) < 1e-4
        slowest = shard.max_over_ranks(10.0 + rank)
        assert slowest == 10.0 + world - 1
        with open(os.path.join(out_dir, f"ok{rank}"), "w") as f:
            f.write("ok")
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_clips", [8, 7])
def test_gather_over_two_ranks(tmp_path, n_clips):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), n_clips, str(tmp_path)), nprocs=world, join=True)
    assert all((tmp_path / f"ok{r}").exists() for r in range(world))
