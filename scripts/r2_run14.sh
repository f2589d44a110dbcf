#!/bin/bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
CA_ATTN_PRE=1 timeout 300 python -m tests.kernel_checks --group attn > gpurun_out/r2_attn_pre.log 2>&1; echo "attn pre rc=$? ok=$(grep -c '\[ok' gpurun_out/r2_attn_pre.log)"; grep "FAIL\|EXC" gpurun_out/r2_attn_pre.log | cut -c1-300
CA_ATTN_PRE=1 timeout 300 python -m tests.module_checks --groups shapes > gpurun_out/r2_shapes_pre.log 2>&1; echo "shapes pre rc=$?"
for rep in 1 2; do for pre in 0 1; do for s in attn attn4k attn1k attn77; do CA_ATTN_PRE=$pre timeout 120 python scripts/prof_kernels.py $s --time 2>&1 | tail -1 | sed "s/^/pre$pre /"; done; done; done | tee gpurun_out/r2_attn_pre.txt
for s in attn attn4k; do CA_ATTN_PRE=1 CA_ATTN_THROTTLE=8 timeout 120 python scripts/prof_kernels.py $s --time 2>&1 | tail -1 | sed "s/^/pre1 th8 /"; done | tee -a gpurun_out/r2_attn_pre.txt
K='regex:^(gemm_conv|attention)'
CA_ATTN_PRE=1 timeout 600 ncu --set full --clock-control none --import-source on -k "$K" -s 3 -c 1 -f -o gpurun_out/r2_attn4k_pre python scripts/prof_kernels.py attn4k > gpurun_out/ncu_attn4k_pre.log 2>&1
echo "ncu rc=$?"
