#!/bin/bash
# one development cycle on a GPU box: GPU tests, bench with the full kernel table, optional ncu --set full of $NCU_K
tag=${1:-cycle}
out=gpurun_out/$tag
mkdir -p $out
if [ "${SKIP_TESTS:-0}" != 1 ]; then
  timeout 1800 python -m pytest tests -q -m gpu -s 2>&1 | grep -v Warning > $out/pytest_gpu.log
  grep -E "cfg parity|cw_strong|generate_attacks\(\)|passed|failed|FAILED|^E  " $out/pytest_gpu.log | cut -c1-400 | tail -${TAILN:-30}
fi
timeout 900 python bench.py --steps ${STEPS:-5} --warmup 3 --no-cpu-baseline ${BENCH_ARGS} --kernel-times $out/kernel_times.json > $out/bench.log 2>&1
tail -1 $out/bench.log | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('value', round(d['value'], 1), 'e2e', round(d['e2e']['value'], 1), 'ms', round(d['ms_per_step'], 2), 'launches', d['gpu_launches'], 'path', round(d['path_roofline']['frac'], 4), 'clk', d['clocks'])
print('parity', d['attack'].get('parity_vs_reference'))
print({k: round(v.get('value', 0), 1) for k, v in d.get('other_workloads', {}).items()})
" 2>&1 | tail -4
python - <<PY
import json
rows = json.load(open("$out/kernel_times.json"))
tot = sum(r["total_ms"] for r in rows)
print("profiled call total ms", round(tot, 2))
for r in rows[:40]:
    print(f'{r["name"]:22s} n={r["count"]:4d} total {r["total_ms"]:7.3f} ms  avg {1e3*r["total_ms"]/r["count"]:7.1f} us')
PY
if [ -n "$NCU_K" ]; then
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$NCU_K" -s ${NCU_S:-2} -c ${NCU_C:-2} -f -o $out/full \
      python tools/profile_grad.py --calls 2 > $out/ncu_full.log 2>&1
  tail -3 $out/ncu_full.log
fi
