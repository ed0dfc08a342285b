#!/bin/bash
# Round-end check on a GPU box: GPU tests, smoke, the bench lines of every workload, the reference arm, the ncu launch
# list and (NCU_FULL=1) one `--set full` capture of the kernels changed this round.
# Usage: gpurun -- '[NCU_FULL=1] bash tools/round_check.sh [tag]'
tag=${1:-final}
out=gpurun_out/$tag
mkdir -p $out
python -m pytest tests -q -m gpu 2>&1 | tail -4 > $out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > $out/smoke.log 2>&1
python bench.py --kernel-times $out/kernel_times_lcnn.json > $out/bench_lcnn.log 2>&1
python bench.py --impl reference --steps 1 --warmup 0 > $out/bench_reference.log 2>&1
python bench.py --workload specrnet --steps 3 --no-cpu-baseline > $out/bench_specrnet.log 2>&1
python bench.py --workload rawnet3 --steps 3 > $out/bench_rawnet3.log 2>&1
python bench.py --workload lcnn_advtrain --steps 3 --no-cpu-baseline > $out/bench_lcnn_advtrain.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -s 150 -c 400 --csv --log-file $out/launches.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > $out/ncu_launches.log 2>&1
if [ "${NCU_FULL:-0}" = 1 ]; then  # one --set full capture of the second gradient evaluation's frontend / first-block / 3x3 kernels
  ncu --set full --clock-control none --import-source on \
      -k regex:"fe_bwd_kernel|fe_dct_t_kernel|fe_floor_dct_kernel|fe_power_db_kernel|conv0_bwd_cells_kernel|conv_p3_kernel" -s 13 -c 13 \
      -o $out/full python tools/profile_grad.py --calls 2 > $out/ncu_full.log 2>&1
fi
tail -2 $out/pytest_gpu.log; cat $out/smoke.log | tail -1
for f in lcnn reference specrnet rawnet3 lcnn_advtrain; do tail -1 $out/bench_$f.log | cut -c1-160; done
