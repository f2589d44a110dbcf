#!/bin/bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
for cfg in "4 2" "104 2" "108 2" "304 2" "4 0" "4 1"; do set -- $cfg; for s in attn attn4k attn1k; do CA_ATTN_THROTTLE=$1 CA_ATTN_POLY=$2 timeout 120 python scripts/prof_kernels.py $s --time 2>&1 | tail -1 | sed "s/^/th$1 poly$2 /"; done; done | tee gpurun_out/r2_attn_lag.txt
CA_ATTN_THROTTLE=104 timeout 300 python -m tests.kernel_checks --group attn > gpurun_out/r2_attn_lag1.log 2>&1; echo "attn lag1 ok=$(grep -c '\[ok' gpurun_out/r2_attn_lag1.log)"
CA_ATTN_THROTTLE=4 CA_ATTN_POLY=1 timeout 300 python -m tests.kernel_checks --group attn > gpurun_out/r2_attn_p1.log 2>&1; echo "attn poly1 ok=$(grep -c '\[ok' gpurun_out/r2_attn_p1.log)"
