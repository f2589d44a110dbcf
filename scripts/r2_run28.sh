#!/bin/bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out; export PYTHONPATH=$PWD
{
for rep in 1 2; do
for w in attn77 attn1k; do
echo "default      $(timeout 100 python scripts/prof_kernels.py $w --time 2>&1 | tail -1)"
for bo in 0 16; do echo "backoff=$bo    $(CA_ATTN_BACKOFF=$bo timeout 100 python scripts/prof_kernels.py $w --time 2>&1 | tail -1)"; done
echo "grid=items   $(CA_ATTN_GRID=i timeout 100 python scripts/prof_kernels.py $w --time 2>&1 | tail -1)"
done; done
} | tee gpurun_out/r2_attn77_knobs.txt
