#!/bin/bash
# compute-sanitizer memcheck over the kernels written this round without a debugger at hand (geo_lut, ism_geo, geo distance pass on
# compact rows); PYTORCH_NO_CUDA_MEMORY_CACHING=1 so that every tensor is its own allocation and an out-of-bounds access is seen
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out
mkdir -p $O
export PYTORCH_NO_CUDA_MEMORY_CACHING=1
timeout 85 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_kernels.py -q -x -k "geo_embed_lut and 33" > $O/san_geo_lut.log 2>&1; echo "geo_lut rc=$?" | tee -a $O/san_geo_lut.log; tail -n 4 $O/san_geo_lut.log
timeout 70 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_ism_geo.py -q -x -k "frame12 or edge" > $O/san_ism_geo.log 2>&1; echo "ism_geo rc=$?" | tee -a $O/san_ism_geo.log; tail -n 4 $O/san_ism_geo.log
