#!/bin/bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
for rep in 1 2; do for mk in 0 1536; do for w in i2vgen sdxl; do
CA_GEMM_BN320_MINK=$mk timeout 900 python bench.py --workload $w --steps 8 --warmup 3 --skip-cpu-baseline --skip-eager-baseline --skip-e2e --skip-profile > gpurun_out/r2_ab_${w}_mk$mk.json 2> /dev/null
echo "mink=$mk $w $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/r2_ab_${w}_mk$mk.json)"
done; done; done | tee gpurun_out/r2_bn320_mink_ab.txt
