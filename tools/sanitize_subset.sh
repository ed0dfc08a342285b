#!/bin/bash
# compute-sanitizer on the kernels whose synchronisation changed late in round 2 (conv_p3: one MMA-issuing warp per M-tile, resident
# weight slices; fe_xrank exchange): racecheck / synccheck / memcheck of one gradient evaluation, LCNN and SpecRNet, small shapes.
out=gpurun_out/${1:-r02_sanitizer_subset}; mkdir -p $out
CS=/usr/local/cuda/bin/compute-sanitizer
run() {
  local tool=$1 tag=$2; shift 2
  timeout 600 $CS --tool $tool --print-limit 5 python tools/profile_grad.py "$@" > $out/${tool}_${tag}.log 2>&1
  echo "$tool $tag rc=$? : $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|hazard|grad norm|attack linf' $out/${tool}_${tag}.log | tr '\n' ' ' | cut -c1-300)"
}
run racecheck lcnn --model lcnn --batch 2 --samples 16000 --calls 1
run synccheck lcnn --model lcnn --batch 2 --samples 16000 --calls 1
run memcheck specrnet --model specrnet --batch 2 --samples 16000 --calls 1
run racecheck specrnet --model specrnet --batch 2 --samples 16000 --calls 1
run synccheck specrnet --model specrnet --batch 2 --samples 16000 --calls 1
