#!/bin/bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
for th in 0 4; do CA_ATTN_THROTTLE=$th timeout 300 python -m tests.kernel_checks --group attn > gpurun_out/r2_attn_sp_th$th.log 2>&1; echo "attn th=$th rc=$? ok=$(grep -c '\[ok' gpurun_out/r2_attn_sp_th$th.log)"; done
for rep in 1 2; do for th in 0 4 8; do for s in attn attn4k attn1k attn77; do CA_ATTN_THROTTLE=$th timeout 120 python scripts/prof_kernels.py $s --time 2>&1 | tail -1 | sed "s/^/th$th /"; done; done; done | tee gpurun_out/r2_attn_throttle3.txt
for th in 0 4 0 4; do CA_ATTN_THROTTLE=$th timeout 900 python bench.py --steps 10 --warmup 3 --skip-cpu-baseline --skip-eager-baseline --skip-e2e --skip-profile > gpurun_out/r2_bench_sdxl_sp_th$th.json 2> gpurun_out/r2_bench_sdxl_sp_th$th.err
echo "bench th$th rc=$?"; cut -c1-230 gpurun_out/r2_bench_sdxl_sp_th$th.json | grep -o '"ms_per_step": [0-9.]*'; done
