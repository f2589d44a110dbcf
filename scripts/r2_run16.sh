#!/bin/bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
echo "--- BN160"; CA_GEMM_BN320=0 timeout 300 python scripts/gemm_trace.py --only mid 2>&1 | grep "==" | tee gpurun_out/r2_mid_bn.txt
echo "--- BN320"; timeout 300 python scripts/gemm_trace.py --only mid 2>&1 | grep "==" | tee -a gpurun_out/r2_mid_bn.txt
