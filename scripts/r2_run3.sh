#!/bin/bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
K='regex:^(gemm_conv|attention)'
timeout 600 ncu --set full --clock-control none --import-source on -k "$K" -s 3 -c 1 -f -o gpurun_out/r2_attn4k_base python scripts/prof_kernels.py attn4k > gpurun_out/ncu_attn4k.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/ncu_attn4k.log
export CA_B200_LIB=$PWD/ctrl_adapter_b200/libctrl_adapter_b200_all.so CA_GEMM_BN320=1 CA_FOLD_SMALL_CONV=1
timeout 1500 python -m pytest tests/ -x -q -m gpu > gpurun_out/r2_pytest_gpu_all.log 2>&1
echo "pytest(all variants) rc=$?"; tail -3 gpurun_out/r2_pytest_gpu_all.log | cut -c1-200
timeout 600 python bench.py --steps 10 --warmup 3 --skip-cpu-baseline --skip-e2e > gpurun_out/r2_bench_all.json 2> gpurun_out/r2_bench_all.err
echo "bench all rc=$?"; cut -c1-300 gpurun_out/r2_bench_all.json
