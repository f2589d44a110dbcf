#!/bin/bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
K='regex:^(gemm_conv|attention|gn_|layernorm|softmax|temporal)'
for w in attn4k conv gn ln lin_small; do
timeout 300 ncu --set full --clock-control none --import-source on -k "$K" -s 3 -c 1 -f -o gpurun_out/r2f_$w python scripts/prof_kernels.py $w > gpurun_out/ncu_r2f_$w.log 2>&1; echo "ncu $w rc=$?"
done
timeout 300 ncu --set full --clock-control none --import-source on -k "$K" -s 6 -c 1 -f -o gpurun_out/r2f_gn_apply python scripts/prof_kernels.py gn > gpurun_out/ncu_r2f_gn_apply.log 2>&1; echo "ncu gn_apply rc=$?"
