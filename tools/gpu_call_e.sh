#!/bin/bash
# fifth gpurun call: full GPU suite on the final build, the final bench records (PEM with CPU / same-box baselines, launch by launch
# comparator, ISM, RGB), launch list + ncu --set full CSV of the step, launch-shape variants of the table kernel
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out
mkdir -p $O
( time timeout 1500 python -m pytest tests -m gpu -q -x ) > $O/e_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/e_pytest_gpu.log; tail -n 7 $O/e_pytest_gpu.log
for c in 1 2; do SAM6D_GEO_LUT_CFG=$c timeout 300 python -m pytest tests/test_gpu_kernels.py -q -k geo_embed_lut 2>&1 | tail -n 1; done
timeout 600 python bench.py --steps 20 --warmup 3 > $O/e_bench_final.json 2> $O/e_bench_final.err
for c in 1 2; do SAM6D_GEO_LUT_CFG=$c timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-ref-gpu > $O/e_bench_cfg$c.json 2> $O/e_bench_cfg$c.err; done
timeout 300 python bench.py --steps 20 --warmup 3 --no-graph --no-cpu-baseline --no-ref-gpu > $O/e_bench_nograph.json 2> $O/e_bench_nograph.err
for f in final cfg1 cfg2 nograph; do python - <<PY
import json
try:
    d = json.load(open("$O/e_bench_$f.json")); g = d.get("roofline_geo") or {}
    print("$f", round(d["value"],1), round(d["ms_per_step"],4), round(d["e2e"]["value"],1), d["gpu_launches"], round(d["roofline"]["frac"],4), round(d["roofline"]["attention_frac"],4), g.get("avg_launch_ms"), g.get("frac"))
except Exception as e: print("$f failed", e)
PY
done
timeout 200 python tools/kernel_times.py 2>&1 | tail -n 5 > $O/e_ktimes.txt; cat $O/e_ktimes.txt
SAM6D_PROFILE_ONE_STEP=1 timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/e_launches_step.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-ref-gpu > $O/e_launches.log 2>&1
SAM6D_PROFILE_ONE_STEP=1 timeout 900 ncu --profile-from-start off --set full --clock-control none -k "regex:rpe_scores_tc|geo_embed|pe_tc_kernel|tail_tc|coarse_select|ball_query_pair|fine_pass|linattn_tc|gemm_tma|attn_tc" -c 36 -o /tmp/e_step_full -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-ref-gpu > $O/e_ncu_full.log 2>&1
ncu -i /tmp/e_step_full.ncu-rep --page raw --csv > $O/e_step_full_raw.csv 2>> $O/e_ncu_full.log
timeout 600 python bench.py --workload ism --steps 5 --warmup 3 --no-cpu-baseline > $O/e_bench_ism.json 2> $O/e_bench_ism.err; tail -c 200 $O/e_bench_ism.json
timeout 600 python bench.py --rgb --steps 10 --warmup 3 --no-cpu-baseline --no-ref-gpu > $O/e_bench_rgb.json 2> $O/e_bench_rgb.err; tail -c 200 $O/e_bench_rgb.json
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/e_smoke.log 2>&1; tail -n 4 $O/e_smoke.log
du -sh $O
