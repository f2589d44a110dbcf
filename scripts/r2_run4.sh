#!/bin/bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
for sp in 1 2; do
CA_ATTN_SPLIT=$sp timeout 600 python -m tests.kernel_checks --group attn --json gpurun_out/r2_attn_split$sp.json > gpurun_out/r2_attn_split$sp.log 2>&1
echo "attn checks split=$sp rc=$?"; grep -c "\[ok" gpurun_out/r2_attn_split$sp.log; grep "FAIL\|EXC" gpurun_out/r2_attn_split$sp.log | cut -c1-300
done
for sp in 1 2; do for s in attn attn4k attn1k attn77; do CA_ATTN_SPLIT=$sp timeout 120 python scripts/prof_kernels.py $s --time 2>&1 | tail -1 | sed "s/^/split$sp /"; done; done | tee gpurun_out/r2_attn_split.txt
for p in 0 3 4; do for s in attn attn4k; do CA_ATTN_POLY=$p timeout 120 python scripts/prof_kernels.py $s --time 2>&1 | tail -1 | sed "s/^/split2 poly$p /"; done; done | tee -a gpurun_out/r2_attn_split.txt
K='regex:^(gemm_conv|attention)'
timeout 600 ncu --set full --clock-control none --import-source on -k "$K" -s 3 -c 1 -f -o gpurun_out/r2_attn4k_split2 python scripts/prof_kernels.py attn4k > gpurun_out/ncu_attn4k_split2.log 2>&1
echo "ncu rc=$?"
timeout 1500 python -m pytest tests/ -x -q -m gpu > gpurun_out/r2_pytest_gpu_run4.log 2>&1
echo "pytest rc=$?"; tail -5 gpurun_out/r2_pytest_gpu_run4.log | cut -c1-400
timeout 600 python bench.py --steps 10 --warmup 3 --skip-cpu-baseline > gpurun_out/r2_bench_run4.json 2> gpurun_out/r2_bench_run4.err
echo "bench rc=$?"; cut -c1-400 gpurun_out/r2_bench_run4.json; tail -3 gpurun_out/r2_bench_run4.err
