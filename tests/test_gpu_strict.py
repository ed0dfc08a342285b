"""GPU: strict multi-GPU mode (SURVEY.md §8(e) ii; include/advb200.h advb_xrank_*).  amplitude_to_DB's top_db floor is relative to
the maximum of the WHOLE batch (F5), so a clip-sharded run equals the single-device run only when the ranks share that maximum
and the summed gradient of the clamped elements.  The exchange is a 32-thread kernel storing into the peers' mailboxes.

* in process: two / four engine handles ("ranks") on one device, one stream each, connected by device pointers;
* two processes under torchrun (CUDA IPC mapping; one GPU each when the box has two, else both on cuda:0): tools/strict_equiv.py.

Model: SpecRNet + MFCC (floor always active: mel filter 0 is identically zero); the second half of the batch is 6 dB quieter.
"""
import copy
import json
import os
import subprocess
import sys

import pytest
import torch

import helpers

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _check(line):
    s, p = line["strict_vs_whole_batch"], line["per_shard_floor_vs_whole_batch"]
    assert not any(line["timed_out"]), line
    assert all(line["disconnect_restores_plain"]), line
    # forward: same floor => same features => same logits; backward: only the rounding of the clamped-mass sum differs
    assert s["logit_maxdiff"] <= 1e-6, line
    assert s["grad_rel_err"] < 1e-5 and s["grad_sign_mismatch"] < 1e-3, line
    assert s["pgd_labels_equal"] and s["pgd_adv_logit_maxdiff"] < 1e-3 and s["pgd_element_mismatch"] < 0.02, line
    # the test has teeth: with per-shard floors (the default, = nn.DataParallel) the quieter shard's clips differ visibly
    assert p["logit_maxdiff"] > 100 * max(s["logit_maxdiff"], 1e-7), line
    assert p["grad_rel_err"] > 100 * s["grad_rel_err"], line


@pytest.mark.parametrize("world", [2, 4])
def test_strict_floor_handles_in_process(world, cuda_device, record_property):
    """`world` engine handles ("ranks") on one device, one stream each, coupled through their mailboxes' device pointers."""
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import strict_equiv as se
    from advb200 import engine

    case, x, y, holder, state = se.batch(world)
    n, half = x.shape[0], x.shape[0] // world
    holders = [helpers.load_holder_state(copy.deepcopy(holder), state, cuda_device) for _ in range(world + 1)]
    torch.manual_seed(77)
    noise = torch.empty_like(x).uniform_(-0.001, 0.001).to(cuda_device)
    xd, yd = x.to(cuda_device), y.to(cuda_device)
    full = se.run_all(engine.engine_for(holders[0], n, x.shape[1]), holders[0], xd, yd, noise, n)
    engs = [engine.engine_for(h, half, x.shape[1]) for h in holders[1:]]
    spans = [(r * half, (r + 1) * half) for r in range(world)]
    # per-shard floors first: this also makes every lazy allocation and graph capture happen before the handles are coupled
    plain = [se.run_all(e, h, xd[lo:hi], yd[lo:hi], noise[lo:hi], n) for e, h, (lo, hi) in zip(engs, holders[1:], spans)]
    ptrs = [e.xrank_export()[1] for e in engs]
    for r, e in enumerate(engs):
        e.xrank_connect(r, world, local_ptrs=ptrs)
    # the two "ranks" run on two streams: every call only enqueues, the exchange kernels of one handle wait for the other's
    streams = [torch.cuda.Stream(cuda_device) for _ in engs]
    for s in streams:
        s.wait_stream(torch.cuda.current_stream(cuda_device))
    from advb200 import torchattacks as ta

    atks = [ta.PGD(h, eps=0.001, alpha=2 / 255, steps=se.STEPS, random_start=True) for h in holders[1:]]
    for a in atks:
        a.set_training_mode(model_training=True, batchnorm_training=False)
    res = [dict() for _ in engs]
    for stage in ("grad", "pgd", "fwd"):  # same call sequence on both handles, interleaved like two ranks in lock step
        for r, (e, (lo, hi)) in enumerate(zip(engs, spans)):
            with torch.cuda.stream(streams[r]):
                if stage == "grad":
                    g, l = e.grad(xd[lo:hi], yd[lo:hi], n_global=n)
                    res[r].update(grad=g, logits=l.flatten())
                elif stage == "pgd":
                    res[r]["adv"] = atks[r].forward(xd[lo:hi], yd[lo:hi], noise=noise[lo:hi])
                else:
                    res[r]["adv_logits"] = e.forward(res[r]["adv"]).flatten()
    torch.cuda.synchronize(cuda_device)
    timed_out = [e.strict_timed_out() for e in engs]
    for e in engs:
        e.xrank_connect(0, 1)
    again = [se.run_all(e, h, xd[lo:hi], yd[lo:hi], noise[lo:hi], n) for e, h, (lo, hi) in zip(engs, holders[1:], spans)]
    cat = lambda parts: {k: torch.cat([p[k] for p in parts]) for k in full}  # noqa: E731
    line = {"timed_out": timed_out,
            "disconnect_restores_plain": [all(torch.equal(a[k], p[k]) for k in p) for a, p in zip(again, plain)],
            "strict_vs_whole_batch": se.compare(full, cat(res)), "per_shard_floor_vs_whole_batch": se.compare(full, cat(plain))}
    record_property("strict", json.dumps(line))
    print("strict in-process:", json.dumps(line))
    _check(line)
    # FAB / CW cannot keep the ranks' call sequences equal: refused under strict mode
    engs[0].xrank_connect(0, world, local_ptrs=ptrs)
    with pytest.raises(RuntimeError, match="strict"):
        ta.CW(holders[1], c=1e-4, kappa=0, steps=2, lr=0.01)(xd[:half], yd[:half])
    engs[0].xrank_connect(0, 1)


def _free_port():
    import socket

    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _two_contexts_possible():
    """Two ranks need two GPUs, or one GPU in the default compute mode (two processes may then share it)."""
    if torch.cuda.device_count() >= 2:
        return True
    try:
        import pynvml

        pynvml.nvmlInit()
        return pynvml.nvmlDeviceGetComputeMode(pynvml.nvmlDeviceGetHandleByIndex(0)) == pynvml.NVML_COMPUTEMODE_DEFAULT
    except Exception:
        return True


def test_strict_floor_two_processes(cuda_device, tmp_path, record_property):
    if not _two_contexts_possible():
        pytest.skip("one GPU in an exclusive compute mode: two rank processes cannot share it")
    out = tmp_path / "strict_equiv.json"
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "tools", "strict_equiv.py"), "--out", str(out)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    line = json.loads(out.read_text())
    record_property("strict", json.dumps(line))
    print("strict two processes:", json.dumps(line))
    _check(line)
