#!/bin/bash
# second gpurun call: bench lines (captured step with the 16- / 8-warp tail), per-kernel times, launch list, ncu --set full of the
# step's main kernels exported as CSV (the .ncu-rep stays on the box: gpurun_out/ is capped at 64 MiB), new tests, ISM / RGB lines
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out
mkdir -p $O
( time timeout 600 python -m pytest tests/test_gpu_ism_geo.py tests/test_gpu_graph.py tests/test_gpu_cli.py -x -q ) > $O/b_pytest_new.log 2>&1; echo "pytest rc=$?" >> $O/b_pytest_new.log; tail -n 6 $O/b_pytest_new.log
timeout 600 python bench.py --steps 20 --warmup 3 > $O/b_bench_graph.json 2> $O/b_bench_graph.err
timeout 200 python tools/kernel_times.py > $O/b_ktimes_tail16.txt 2>&1
SAM6D_TAIL_EPI_WARPS=8 timeout 200 python tools/kernel_times.py > $O/b_ktimes_tail8.txt 2>&1
SAM6D_TAIL_EPI_WARPS=8 timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-ref-gpu > $O/b_bench_tail8.json 2> $O/b_bench_tail8.err
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-ref-gpu > $O/b_bench_graph2.json 2> $O/b_bench_graph2.err
for f in graph tail8 graph2; do python - <<PY
import json
try:
    d = json.load(open("$O/b_bench_$f.json")); print("$f", d["value"], d["ms_per_step"], d["e2e"]["value"], d["gpu_launches"], d["roofline"]["frac"], d["roofline"]["attention_frac"], d["roofline_tensor"]["frac"])
except Exception as e: print("$f failed", e)
PY
done
tail -n 5 $O/b_ktimes_tail16.txt $O/b_ktimes_tail8.txt
SAM6D_PROFILE_ONE_STEP=1 timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/b_launches_step.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-ref-gpu > $O/b_launches.log 2>&1
SAM6D_PROFILE_ONE_STEP=1 timeout 900 ncu --profile-from-start off --set full --clock-control none -k "regex:rpe_scores_tc|geo_embed_tc|pe_tc_kernel|tail_tc|coarse_select|ball_query_pair|fine_pass|linattn_tc|gemm_tma" -c 30 -o /tmp/b_step_full -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-ref-gpu > $O/b_ncu_full.log 2>&1
ncu -i /tmp/b_step_full.ncu-rep --page raw --csv > $O/b_step_full_raw.csv 2>> $O/b_ncu_full.log
timeout 600 python bench.py --workload ism --steps 5 --warmup 3 > $O/b_bench_ism.json 2> $O/b_bench_ism.err; tail -c 300 $O/b_bench_ism.json
timeout 600 python bench.py --rgb --steps 10 --warmup 3 --no-cpu-baseline --no-ref-gpu > $O/b_bench_rgb.json 2> $O/b_bench_rgb.err; tail -c 300 $O/b_bench_rgb.json
du -sh $O; ls -la $O | tail -n 25
