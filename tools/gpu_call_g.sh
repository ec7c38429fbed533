#!/bin/bash
# sixth gpurun call: tail kernel with its parameters prefetched into L1 -- parity tests, bench, per-kernel times; N = 1 lines of the
# strong-scaling workloads on the final build
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out
mkdir -p $O
( timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_pem.py tests/test_gpu_graph.py -q -x ) > $O/g_pytest.log 2>&1; echo "rc=$?" >> $O/g_pytest.log; tail -n 4 $O/g_pytest.log
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > $O/g_bench.json 2> $O/g_bench.err
python - <<PY
import json
try:
    d = json.load(open("$O/g_bench.json")); g = d.get("roofline_geo") or {}
    print("g", round(d["value"],1), round(d["ms_per_step"],4), round(d["e2e"]["value"],1), d["gpu_launches"], round(d["roofline"]["frac"],4), g.get("avg_launch_ms"), d.get("same_box_reference",{}).get("ball_query_r0.1x32_r0.2x64_us"))
except Exception as e: print("g failed", e)
PY
timeout 200 python tools/kernel_times.py 2>&1 | tail -n 5 > $O/g_ktimes.txt; cat $O/g_ktimes.txt
for wl in ycbv lmo; do timeout 300 python bench.py --workload $wl --steps 5 --warmup 3 2> $O/g_${wl}_n1.err | tail -n 1 > $O/g_${wl}_n1.json; head -c 200 $O/g_${wl}_n1.json; echo; done
