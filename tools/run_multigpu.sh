set -x
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 600 $TR --nproc-per-node 2 --master-port 29511 -m pytest tests/test_gpu_dist_nccl.py -q -x > gpurun_out/r2_nccl_test.log 2>&1; tail -3 gpurun_out/r2_nccl_test.log
for wl in ycbv lmo; do
  timeout 600 python bench.py --workload $wl --steps 5 --warmup 3 2> gpurun_out/r2_${wl}_n1.err | tail -1 > gpurun_out/r2_${wl}_n1.json
  timeout 600 $TR --nproc-per-node 2 --master-port 29512 bench.py --workload $wl --gpus 2 --steps 5 --warmup 3 2> gpurun_out/r2_${wl}_n2.err | tail -1 > gpurun_out/r2_${wl}_n2.json
done
timeout 600 $TR --nproc-per-node 2 --master-port 29513 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline --no-ref-gpu 2> gpurun_out/r2_pem_n2.err | tail -1 > gpurun_out/r2_pem_n2.json
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-ref-gpu 2> gpurun_out/r2_pem_n1.err | tail -1 > gpurun_out/r2_pem_n1.json
head -c 600 gpurun_out/r2_*_n?.json
