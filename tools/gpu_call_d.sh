#!/bin/bash
# fourth gpurun call: the table kernel with the [0, 32) distance table and hoisted index math, both arithmetic variants through the
# parity tests and the bench; per-warp barrier arrivals in the tail kernel
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out
mkdir -p $O
( timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_pem.py tests/test_gpu_graph.py -q -x ) > $O/d_pytest_precise.log 2>&1; echo "rc=$?" >> $O/d_pytest_precise.log; tail -n 4 $O/d_pytest_precise.log
( SAM6D_GEO_LUT_PRECISE=0 timeout 900 python -m pytest tests/test_gpu_pem.py -q ) > $O/d_pytest_packed.log 2>&1; echo "rc=$?" >> $O/d_pytest_packed.log; tail -n 12 $O/d_pytest_packed.log
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-ref-gpu > $O/d_bench_lut_precise.json 2> $O/d_bench_lut_precise.err
SAM6D_GEO_LUT_PRECISE=0 timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-ref-gpu > $O/d_bench_lut_packed.json 2> $O/d_bench_lut_packed.err
for f in lut_precise lut_packed; do python - <<PY
import json
try:
    d = json.load(open("$O/d_bench_$f.json")); g = d.get("roofline_geo") or {}
    print("$f", round(d["value"],1), round(d["ms_per_step"],4), round(d["e2e"]["value"],1), d["gpu_launches"], round(d["roofline"]["frac"],4), g.get("avg_launch_ms"), g.get("frac"))
except Exception as e: print("$f failed", e)
PY
done
timeout 200 python tools/kernel_times.py 2>&1 | tail -n 5 > $O/d_ktimes.txt; cat $O/d_ktimes.txt
SAM6D_PROFILE_ONE_STEP=1 timeout 600 ncu --profile-from-start off --set full --clock-control none -k "regex:geo_embed_lut" -c 1 -o /tmp/d_geo_precise -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-ref-gpu > $O/d_ncu_precise.log 2>&1
ncu -i /tmp/d_geo_precise.ncu-rep --page raw --csv > $O/d_geo_precise_raw.csv 2>> $O/d_ncu_precise.log
SAM6D_GEO_LUT_PRECISE=0 SAM6D_PROFILE_ONE_STEP=1 timeout 600 ncu --profile-from-start off --set full --clock-control none -k "regex:geo_embed_lut" -c 1 -o /tmp/d_geo_packed -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-ref-gpu > $O/d_ncu_packed.log 2>&1
ncu -i /tmp/d_geo_packed.ncu-rep --page raw --csv > $O/d_geo_packed_raw.csv 2>> $O/d_ncu_packed.log
du -sh $O
