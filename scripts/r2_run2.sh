#!/bin/bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
timeout 1500 python -m pytest tests/ -x -q -m gpu -rA > gpurun_out/r2_pytest_gpu.log 2>&1
echo "pytest rc=$?"; tail -25 gpurun_out/r2_pytest_gpu.log | cut -c1-200
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attention -s 5 -c 1 -o gpurun_out/r2_attn4k_base python scripts/prof_kernels.py attn4k > gpurun_out/ncu_attn4k.log 2>&1
echo "ncu rc=$?"
for v in pdl epi; do
CA_B200_LIB=$PWD/ctrl_adapter_b200/libctrl_adapter_b200_$v.so timeout 600 python bench.py --steps 10 --warmup 3 --skip-cpu-baseline --skip-e2e > gpurun_out/r2_bench_$v.json 2> gpurun_out/r2_bench_$v.err
echo "bench $v rc=$?"; cut -c1-300 gpurun_out/r2_bench_$v.json
done
for p in 0 2 3; do for s in attn attn4k attn1k; do CA_ATTN_POLY=$p python scripts/prof_kernels.py $s --time 2>&1 | tail -1 | sed "s/^/poly$p /"; done; done | tee gpurun_out/r2_attn_poly.txt
