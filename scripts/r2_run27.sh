#!/bin/bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out; export PYTHONPATH=$PWD
timeout 300 ncu --set full --clock-control none --import-source on -k 'regex:^attention' -s 3 -c 1 -f -o gpurun_out/r2h_attn77 python scripts/prof_kernels.py attn77 > gpurun_out/ncu_r2h_attn77.log 2>&1; echo "ncu rc=$?"
