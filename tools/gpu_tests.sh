#!/bin/bash
# GPU suite + smoke on a gpurun box; logs under gpurun_out/<tag>/
tag=${1:-tests}
out=gpurun_out/$tag
mkdir -p $out
timeout 1800 python -m pytest tests -q -m gpu -s 2>&1 | grep -v Warning > $out/pytest_gpu.log; grep -E "cfg parity|cw_strong|generate_attacks\(\)|passed|failed|FAILED|Error" $out/pytest_gpu.log | cut -c1-600 | tail -40
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $out/smoke.log 2>&1; tail -2 $out/smoke.log
