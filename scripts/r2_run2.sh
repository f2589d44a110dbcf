#!/bin/bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
timeout 1200 python -m tests.module_checks --groups shapes,svd,sparse,svd_loop,fold --json gpurun_out/r2_pending_r1order.json > gpurun_out/r2_pending_r1order.log 2>&1
echo "r1-order rc=$?"; grep -c "\[ok" gpurun_out/r2_pending_r1order.log; grep "FAIL" gpurun_out/r2_pending_r1order.log | cut -c1-400
timeout 1500 python -m pytest tests/ -x -q -m gpu -rA > gpurun_out/r2_pytest_gpu.log 2>&1
echo "pytest rc=$?"; tail -30 gpurun_out/r2_pytest_gpu.log | cut -c1-300
for p in 0 2 3; do for s in attn attn4k attn1k attn77; do CA_ATTN_POLY=$p python scripts/prof_kernels.py $s --time 2>&1 | tail -1 | sed "s/^/poly$p /"; done; done | tee gpurun_out/r2_attn_poly.txt
