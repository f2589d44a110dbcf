#!/bin/bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
for st in 0 150 300 600; do for s in attn attn4k attn1k; do CA_ATTN_STAGGER=$st timeout 120 python scripts/prof_kernels.py $s --time 2>&1 | tail -1 | sed "s/^/stagger$st /"; done; done | tee gpurun_out/r2_attn_knobs.txt
for bo in 0 32 128 256; do for s in attn attn4k attn1k attn77; do CA_ATTN_BACKOFF=$bo timeout 120 python scripts/prof_kernels.py $s --time 2>&1 | tail -1 | sed "s/^/backoff$bo /"; done; done | tee -a gpurun_out/r2_attn_knobs.txt
CA_ATTN_GRID=i timeout 120 python scripts/prof_kernels.py attn4k --time 2>&1 | tail -1 | sed "s/^/grid-items /" | tee -a gpurun_out/r2_attn_knobs.txt
