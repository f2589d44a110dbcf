#!/bin/bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
for w in svd multi; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --workload $w --steps 5 --warmup 3 --skip-cpu-baseline > gpurun_out/r2_bench_${w}_8gpu.json 2> gpurun_out/r2_bench_${w}_8gpu.err
echo "bench $w x8 rc=$?"; cut -c1-400 gpurun_out/r2_bench_${w}_8gpu.json; tail -2 gpurun_out/r2_bench_${w}_8gpu.err | cut -c1-300
done
