#!/bin/bash
out=gpurun_out/r02_p3cl; mkdir -p $out
for cl in 1 2 4; do
  ADVB_P3_CL=$cl timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-other-workloads --kernel-times $out/kt_$cl.json > $out/bench_$cl.log 2>&1
  python - <<PY
import json
try:
    d = json.loads([x for x in open("$out/bench_$cl.log") if x.startswith("{")][-1])
    rows = {r["name"]: r for r in json.load(open("$out/kt_$cl.json"))}
    ks = ["conv_fwd_b2", "conv_bwd_b2", "conv_fwd_b4", "conv_bwd_b4", "conv_fwd_b6", "conv_bwd_b6", "conv_fwd_b8", "conv_bwd_b8"]
    print("CL=$cl value", round(d["value"], 1), "ms", round(d["ms_per_step"], 2), "parity", d["attack"]["parity_vs_reference"]["sign_mismatch_vs_reference"], {k: round(1e3 * rows[k]["total_ms"] / rows[k]["count"], 1) for k in ks})
except Exception as e:
    print("CL=$cl failed", e); print(open("$out/bench_$cl.log").read()[-1500:])
PY
done
