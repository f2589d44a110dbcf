#!/bin/bash
# short-KV attention kernel: parity (attn kernel group) and same-box A/B against the general kernel
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out; export PYTHONPATH=$PWD
timeout 240 python tests/kernel_checks.py --group attn > gpurun_out/r2_attn_checks.log 2>&1; echo "attn checks rc=$?"; grep -E "FAIL|EXC|lk77|lk80|lk33|lk1 |checks ok" gpurun_out/r2_attn_checks.log | cut -c1-220
{
for rep in 1 2; do
echo "short   $(timeout 100 python scripts/prof_kernels.py attn77 --time 2>&1 | tail -1)"
echo "general $(CA_ATTN_SHORT=0 timeout 100 python scripts/prof_kernels.py attn77 --time 2>&1 | tail -1)"
done
} | tee gpurun_out/r2_attn_short_ab.txt
