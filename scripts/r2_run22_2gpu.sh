#!/bin/bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2_bench_sdxl_2gpu.json 2> gpurun_out/r2_bench_sdxl_2gpu.err
echo "bench x2 rc=$?"; cut -c1-300 gpurun_out/r2_bench_sdxl_2gpu.json; tail -2 gpurun_out/r2_bench_sdxl_2gpu.err | cut -c1-200
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29522 bench.py --impl reference --gpus 2 --steps 1 --warmup 0 > gpurun_out/r2_bench_ref_2gpu.json 2> gpurun_out/r2_bench_ref_2gpu.err
echo "ref x2 rc=$?"; cut -c1-300 gpurun_out/r2_bench_ref_2gpu.json
