#!/bin/bash
# usage (on the GPU box, N GPUs visible): bash tools/run_multigpu.sh N   -- NCCL parity test + the three workloads at N ranks
N=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --nproc-per-node $N"
timeout 600 $TR --master-port 29511 -m pytest tests/test_gpu_dist_nccl.py -q -x > gpurun_out/r2_nccl_test_n$N.log 2>&1; tail -3 gpurun_out/r2_nccl_test_n$N.log
for wl in ycbv lmo; do
  timeout 600 $TR --master-port 29512 bench.py --workload $wl --gpus $N --steps 5 --warmup 3 2> gpurun_out/r2_${wl}_n$N.err | tail -1 > gpurun_out/r2_${wl}_n$N.json
done
timeout 600 $TR --master-port 29513 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline --no-ref-gpu 2> gpurun_out/r2_pem_n$N.err | tail -1 > gpurun_out/r2_pem_n$N.json
head -c 400 gpurun_out/r2_*_n$N.json
