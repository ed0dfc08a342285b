#!/bin/bash
# first GPU pass of round 2: Toeplitz kernel check (both descriptor conventions, time-boxed), then the whole GPU suite
out=gpurun_out/r02_first
mkdir -p $out
timeout 300 python tools/c0t_check.py --time > $out/c0t_swap0.log 2>&1; echo "c0t swap0 rc $?" >> $out/c0t_swap0.log
ADVB_C0T_SWAP=1 timeout 300 python tools/c0t_check.py > $out/c0t_swap1.log 2>&1; echo "c0t swap1 rc $?" >> $out/c0t_swap1.log
grep -v Warn $out/c0t_swap0.log | tail -12; grep -v Warn $out/c0t_swap1.log | tail -6
timeout 1500 python -m pytest tests -q -m gpu -x --deselect tests/test_gpu_dropin.py 2>&1 | tail -40 > $out/pytest_gpu.log; tail -25 $out/pytest_gpu.log
timeout 900 python -m pytest tests/test_gpu_dropin.py -q -m gpu -s 2>&1 | tail -30 > $out/pytest_dropin.log; tail -15 $out/pytest_dropin.log
timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu-baseline > $out/bench_lcnn.log 2>&1; tail -1 $out/bench_lcnn.log | cut -c1-3000
