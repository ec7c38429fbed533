#!/bin/bash
# last gpurun call of the round: the full GPU suite, smoke() and the default bench line on the frozen build
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out
mkdir -p $O
( time timeout 1200 python -m pytest tests -m gpu -q -x ) > $O/final_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/final_pytest_gpu.log; tail -n 7 $O/final_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/final_smoke.log 2>&1; tail -n 3 $O/final_smoke.log
timeout 400 python bench.py > $O/final_bench.json 2> $O/final_bench.err
python - <<PY
import json
try:
    d = json.load(open("$O/final_bench.json")); print("final", round(d["value"],1), round(d["ms_per_step"],4), round(d["e2e"]["value"],1), d["gpu_launches"], round(d["roofline"]["frac"],4), d["cpu_baseline"]["value"], d["same_box_reference"].get("ref_gpu_poses_per_s"), d["same_box_reference"].get("ball_query_r0.1x32_r0.2x64_us"))
except Exception as e: print("final failed", e)
PY
