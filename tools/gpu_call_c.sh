#!/bin/bash
# third gpurun call: full GPU suite with the table-interpolated geometric embedding as the default bf16 path, bench lines for its
# two arithmetic variants and the tensor-core comparator, launch list + ncu --set full CSV of the new kernel
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out
mkdir -p $O
( time timeout 1500 python -m pytest tests -m gpu -q -x ) > $O/c_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/c_pytest_gpu.log; tail -n 8 $O/c_pytest_gpu.log
timeout 300 python -m pytest tests/test_gpu_kernels.py -q -k "geo_embed_lut" -s 2>&1 | grep "rms error" > $O/c_geo_lut_rms.txt; cat $O/c_geo_lut_rms.txt | head -20
timeout 600 python bench.py --steps 20 --warmup 3 > $O/c_bench_lut_precise.json 2> $O/c_bench_lut_precise.err
SAM6D_GEO_LUT_PRECISE=0 timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-ref-gpu > $O/c_bench_lut_packed.json 2> $O/c_bench_lut_packed.err
SAM6D_GEO_LUT=0 timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-ref-gpu > $O/c_bench_geo_tc.json 2> $O/c_bench_geo_tc.err
for f in lut_precise lut_packed geo_tc; do python - <<PY
import json
try:
    d = json.load(open("$O/c_bench_$f.json")); g = d.get("roofline_geo") or d.get("roofline_tensor") or {}
    print("$f", round(d["value"],1), round(d["ms_per_step"],4), round(d["e2e"]["value"],1), d["gpu_launches"], round(d["roofline"]["frac"],4), g.get("avg_launch_ms", g.get("avg_call_ms")), g.get("frac"))
except Exception as e: print("$f failed", e)
PY
done
SAM6D_PROFILE_ONE_STEP=1 timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/c_launches_step.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-ref-gpu > $O/c_launches.log 2>&1
SAM6D_PROFILE_ONE_STEP=1 timeout 600 ncu --profile-from-start off --set full --clock-control none -k "regex:geo_embed_lut|geo_embed_tc|geo_indices" -c 4 -o /tmp/c_geo_full -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-ref-gpu > $O/c_ncu_full.log 2>&1
ncu -i /tmp/c_geo_full.ncu-rep --page raw --csv > $O/c_geo_full_raw.csv 2>> $O/c_ncu_full.log
SAM6D_GEO_LUT_PRECISE=0 SAM6D_PROFILE_ONE_STEP=1 timeout 600 ncu --profile-from-start off --set full --clock-control none -k "regex:geo_embed_lut" -c 1 -o /tmp/c_geo_packed -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-ref-gpu > $O/c_ncu_packed.log 2>&1
ncu -i /tmp/c_geo_packed.ncu-rep --page raw --csv > $O/c_geo_packed_raw.csv 2>> $O/c_ncu_packed.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/c_smoke.log 2>&1; tail -n 4 $O/c_smoke.log
du -sh $O
