#!/bin/bash
# Round-end check on a GPU box: GPU tests, smoke, the headline bench (with other_workloads), the reference arm, every other
# workload's own bench line, the ncu launch list and (NCU_FULL=1) one `--set full` capture of the main kernels.
# Usage: gpurun -- '[NCU_FULL=1] bash tools/round_check.sh [tag]'
tag=${1:-final}
out=gpurun_out/$tag
mkdir -p $out
python -m pytest tests -q -m gpu -s 2>&1 | grep -v Warning > $out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > $out/smoke.log 2>&1
python bench.py --kernel-times $out/kernel_times_lcnn.json > $out/bench_lcnn.log 2>&1
python bench.py --impl reference --steps 2 --warmup 1 > $out/bench_reference.log 2>&1
python bench.py --workload specrnet --steps 3 --kernel-times $out/kernel_times_specrnet.json > $out/bench_specrnet.log 2>&1
python bench.py --workload rawnet3 --steps 3 --kernel-times $out/kernel_times_rawnet3.json > $out/bench_rawnet3.log 2>&1
python bench.py --workload rawnet3_fab --steps 2 --warmup 1 --no-cpu-baseline > $out/bench_rawnet3_fab.log 2>&1
python bench.py --workload lcnn_advtrain --steps 5 --no-cpu-baseline > $out/bench_lcnn_advtrain.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -s 150 -c 400 --csv --log-file $out/launches.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-other-workloads > $out/ncu_launches.log 2>&1
if [ "${NCU_FULL:-0}" = 1 ]; then  # one --set full capture of the second gradient evaluation's main kernels
  ncu --set full --clock-control none --import-source on \
      -k regex:"fe_bwd_kernel|fe_dct_t_kernel|fe_floor_dct_kernel|fe_power_db_kernel|conv0_bwd_cells_kernel|conv_p3_kernel|conv0_toeplitz_kernel|conv_light_kernel" -s 21 -c 21 \
      -f -o $out/full python tools/profile_grad.py --calls 2 > $out/ncu_full.log 2>&1
  # gpurun brings back at most 64 MiB: export the raw page here and drop the report (two reports do not fit)
  ncu -i $out/full.ncu-rep --page raw --csv > $out/full_raw.csv 2>/dev/null; rm -f $out/full.ncu-rep
  # SpecRNet's tensor-core convolutions (first forward + backward of one gradient evaluation)
  ncu --set full --clock-control none --import-source on -k regex:"conv_p3_kernel|sr_first_conv1|sr_expand_go|sr_first_bwd" -c 14 \
      -f -o $out/full_specrnet python tools/profile_grad.py --model specrnet --batch 256 --calls 1 > $out/ncu_full_specrnet.log 2>&1
  ncu -i $out/full_specrnet.ncu-rep --page raw --csv > $out/full_specrnet_raw.csv 2>/dev/null; rm -f $out/full_specrnet.ncu-rep
fi
grep -E "passed|failed" $out/pytest_gpu.log | tail -2; tail -1 $out/smoke.log
for f in lcnn reference specrnet rawnet3 rawnet3_fab lcnn_advtrain; do tail -1 $out/bench_$f.log | cut -c1-200; done
