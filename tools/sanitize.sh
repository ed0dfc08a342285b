#!/bin/bash
# compute-sanitizer on one gradient evaluation / one short attack call per model: memcheck, racecheck
# (shared-memory hazards of the hand-rolled mbarrier / TMEM pipelines), synccheck.  Small shapes: the tools slow kernels 10-100x.
out=gpurun_out/${1:-r02_sanitizer}; mkdir -p $out
CS=/usr/local/cuda/bin/compute-sanitizer
run() {  # tool, tag, args...
  local tool=$1 tag=$2; shift 2
  timeout 900 $CS --tool $tool --print-limit 5 python tools/profile_grad.py "$@" > $out/${tool}_${tag}.log 2>&1
  echo "$tool $tag rc=$? : $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|hazard|grad norm|attack linf' $out/${tool}_${tag}.log | tr '\n' ' ' | cut -c1-300)"
}
run memcheck lcnn --model lcnn --batch 2 --samples 16000 --calls 1
run memcheck lcnn_attack --model lcnn --batch 2 --samples 16000 --calls 2 --attack
run racecheck lcnn --model lcnn --batch 2 --samples 16000 --calls 1
run synccheck lcnn --model lcnn --batch 2 --samples 16000 --calls 1
run memcheck specrnet --model specrnet --batch 2 --samples 16000 --calls 1
run racecheck specrnet --model specrnet --batch 2 --samples 16000 --calls 1
run memcheck rawnet3 --model rawnet3 --batch 1 --samples 16000 --calls 1
run racecheck rawnet3 --model rawnet3 --batch 1 --samples 16000 --calls 1
run synccheck rawnet3 --model rawnet3 --batch 1 --samples 16000 --calls 1
# the attack entry points the gradient runs above do not reach (FAB L-inf / L2 + restart, CW, PGDL2, targeted FGSM, projections, mel_spec)
runx() {
  local tool=$1
  timeout 1200 $CS --tool $tool --print-limit 5 python tools/sanitize_extras.py > $out/${tool}_extras.log 2>&1
  echo "$tool extras rc=$? : $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|hazard|extras ok' $out/${tool}_extras.log | tr '\n' ' ' | cut -c1-400)"
}
runx memcheck
runx racecheck
