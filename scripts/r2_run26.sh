#!/bin/bash
# where do the roles of the attention kernel wait at ONE KV tile per item (Lk = 77 cross attention)?
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out; export PYTHONPATH=$PWD
timeout 300 python scripts/gemm_trace.py --attn --events --only 1024x77 2>&1 | tee gpurun_out/r2_attn77_trace.txt | cut -c1-3000
timeout 120 python scripts/prof_kernels.py attn77 --time 2>&1 | tail -1
