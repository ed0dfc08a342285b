#!/bin/bash
# Reduced round-end check (bench lines + ncu evidence; the GPU tests and the unchanged workloads' own lines are in round_check.sh).
# The ncu reports are exported to CSV on the box and deleted: gpurun brings back at most 64 MiB.
tag=${1:-final_lite}
out=gpurun_out/$tag
mkdir -p $out
python bench.py --kernel-times $out/kernel_times_lcnn.json > $out/bench_lcnn.log 2>&1
python bench.py --workload specrnet --steps 3 --no-cpu-baseline --kernel-times $out/kernel_times_specrnet.json > $out/bench_specrnet.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -s 150 -c 400 --csv --log-file $out/launches.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-other-workloads > $out/ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on \
    -k regex:"fe_bwd_kernel|fe_dct_t_kernel|fe_floor_dct_kernel|fe_power_db_kernel|conv0_bwd_cells_kernel|conv_p3_kernel|conv0_toeplitz_kernel|conv_light_kernel" -s 21 -c 21 \
    -f -o $out/full python tools/profile_grad.py --calls 2 > $out/ncu_full.log 2>&1
ncu -i $out/full.ncu-rep --page raw --csv > $out/full_raw.csv 2>/dev/null; rm -f $out/full.ncu-rep
ncu --set full --clock-control none --import-source on -k regex:"conv_p3_kernel|sr_first_conv1|sr_expand_go|sr_first_bwd" -c 14 \
    -f -o $out/full_specrnet python tools/profile_grad.py --model specrnet --batch 256 --calls 1 > $out/ncu_full_specrnet.log 2>&1
ncu -i $out/full_specrnet.ncu-rep --page raw --csv > $out/full_specrnet_raw.csv 2>/dev/null; rm -f $out/full_specrnet.ncu-rep
for f in lcnn specrnet; do tail -1 $out/bench_$f.log | cut -c1-200; done
du -sh $out
