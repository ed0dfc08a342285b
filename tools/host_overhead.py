#!/usr/bin/env python
"""Diagnostic: host-side time of one attack call (enqueue only) vs its GPU time, device-resident and with host buffers."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "audio-deepfake-adversarial-attacks_b200"))
import torch
import bench
dev = torch.device("cuda:0")
job = bench.Job("lcnn", 0, dev, 0)
for _ in range(3): job.atk(job.x, job.y)
torch.cuda.synchronize()
for mode in ("resident", "host"):
    for i in range(6):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        if mode == "host":
            xd = job.x_host.to(dev, non_blocking=True); yd = job.y_host.to(dev, non_blocking=True)
            t_c = time.perf_counter()
            adv = job.atk(xd, yd)
            t1 = time.perf_counter()
            job.adv_host.copy_(adv, non_blocking=True)
        else:
            t_c = t0
            adv = job.atk(job.x, job.y)
            t1 = time.perf_counter()
        t1b = time.perf_counter()
        torch.cuda.synchronize(); t2 = time.perf_counter()
        print(mode, i, "copy-enqueue %.2f ms  atk enqueue %.2f ms  d2h enqueue %.2f  total %.2f ms" % (1e3*(t_c-t0), 1e3*(t1-t_c), 1e3*(t1b-t1), 1e3*(t2-t0)))
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(5):
    xd = job.x_host.to(dev, non_blocking=True); yd = job.y_host.to(dev, non_blocking=True)
    job.adv_host.copy_(job.atk(xd, yd), non_blocking=True)
torch.cuda.synchronize(); print("e2e back-to-back per step %.2f ms" % (1e3*(time.perf_counter()-t0)/5))
import cProfile, pstats
pr = cProfile.Profile(); pr.enable()
for _ in range(3): job.atk(job.x, job.y)
pr.disable(); torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
