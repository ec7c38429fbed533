#!/bin/bash
# one gpurun call: GPU parity suite, the bench lines (captured step / launch by launch / 8-warp tail comparator), launch list and
# ncu --set full capture of the step's main kernels.  Everything lands in gpurun_out/.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $O/a_gpu.txt 2>&1
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > $O/a_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/a_pytest_gpu.log; tail -5 $O/a_pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 3 > $O/a_bench_graph.json 2> $O/a_bench_graph.err; tail -c 600 $O/a_bench_graph.json
timeout 300 python bench.py --steps 20 --warmup 3 --no-graph --no-cpu-baseline --no-ref-gpu > $O/a_bench_nograph.json 2> $O/a_bench_nograph.err
SAM6D_TAIL_EPI_WARPS=8 timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-ref-gpu > $O/a_bench_tail8.json 2> $O/a_bench_tail8.err
for f in nograph tail8; do python - <<PY
import json
try:
    d = json.load(open("$O/a_bench_$f.json")); print("$f", d["value"], d["ms_per_step"], d["e2e"]["value"], d["gpu_launches"])
except Exception as e: print("$f failed", e)
PY
done
timeout 200 python tools/kernel_times.py > $O/a_ktimes_tail16.txt 2>&1
SAM6D_TAIL_EPI_WARPS=8 timeout 200 python tools/kernel_times.py > $O/a_ktimes_tail8.txt 2>&1
tail -5 $O/a_ktimes_tail16.txt $O/a_ktimes_tail8.txt
SAM6D_PROFILE_ONE_STEP=1 timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/a_launches_step.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-ref-gpu > $O/a_launches.log 2>&1
SAM6D_PROFILE_ONE_STEP=1 timeout 900 ncu --profile-from-start off --set full --clock-control none -k "regex:rpe_scores_tc|geo_embed_tc|pe_tc_kernel|tail_tc|coarse_select|ball_query_pair|fine_pass|linattn_tc" -c 44 -o $O/a_step_full -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-ref-gpu > $O/a_ncu_full.log 2>&1
ls -la $O | tail -20
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/a_smoke.log 2>&1; tail -4 $O/a_smoke.log
