#!/bin/bash
# seventh gpurun call: bf16 feature stack + TMA GEMMs for the bf16 token matrices -- parity suite of the matching path, bench
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out
mkdir -p $O
( timeout 900 python -m pytest tests/test_gpu_pem.py tests/test_gpu_graph.py tests/test_gpu_vit.py tests/test_gpu_cli.py -q -x ) > $O/h_pytest.log 2>&1; echo "rc=$?" >> $O/h_pytest.log; tail -n 6 $O/h_pytest.log
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-ref-gpu > $O/h_bench.json 2> $O/h_bench.err
timeout 300 python bench.py --rgb --steps 10 --warmup 3 --no-cpu-baseline --no-ref-gpu > $O/h_bench_rgb.json 2> $O/h_bench_rgb.err
for f in h_bench h_bench_rgb; do python - <<PY
import json
try:
    d = json.load(open("$O/$f.json")); print("$f", round(d["value"],1), round(d["ms_per_step"],4), round(d["e2e"]["value"],1), d["gpu_launches"], round(d["roofline"]["frac"],4))
except Exception as e: print("$f failed", e)
PY
done
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/h_smoke.log 2>&1; tail -n 3 $O/h_smoke.log
