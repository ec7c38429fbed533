#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out
mkdir -p $O
export PYTORCH_NO_CUDA_MEMORY_CACHING=1
timeout 60 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_ism_geo.py -q -x -k "frame12 or edge" > $O/san_ism_geo2.log 2>&1; echo "ism_geo rc=$?" | tee -a $O/san_ism_geo2.log; grep -c "Invalid" $O/san_ism_geo2.log; tail -n 4 $O/san_ism_geo2.log
