#!/bin/bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
CA_ATTN_SPLIT=2 timeout 600 python -m tests.kernel_checks --group attn > gpurun_out/r2_attn_split2.log 2>&1
echo "attn checks split=2 rc=$?"; grep -c "\[ok" gpurun_out/r2_attn_split2.log; grep "FAIL\|EXC" gpurun_out/r2_attn_split2.log | cut -c1-300
for sp in 1 2; do for s in attn attn4k attn1k attn77; do CA_ATTN_SPLIT=$sp timeout 120 python scripts/prof_kernels.py $s --time 2>&1 | tail -1 | sed "s/^/split$sp /"; done; done | tee gpurun_out/r2_attn_split_b.txt
for p in 0 3 4; do for s in attn attn4k; do CA_ATTN_POLY=$p timeout 120 python scripts/prof_kernels.py $s --time 2>&1 | tail -1 | sed "s/^/split2 poly$p /"; done; done | tee -a gpurun_out/r2_attn_split_b.txt
K='regex:^(gemm_conv|attention)'
timeout 600 ncu --set full --clock-control none --import-source on -k "$K" -s 3 -c 1 -f -o gpurun_out/r2_attn4k_split2b python scripts/prof_kernels.py attn4k > gpurun_out/ncu_attn4k_split2b.log 2>&1
echo "ncu rc=$?"
timeout 900 python -m tests.module_checks --group loops > gpurun_out/r2_loops.log 2>&1
echo "loops rc=$?"; grep "^\[" gpurun_out/r2_loops.log | cut -c1-260
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench_run5.json 2> gpurun_out/r2_bench_run5.err
echo "bench rc=$?"; python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_bench_run5.json'))
for k in ('value','ms_per_step','vs_eager','launches_per_step'): print(k, d.get(k))
print('e2e', d['e2e'] and d['e2e']['value']); print('eager', d['eager_gpu_baseline']); print('cpu', {k:v for k,v in (d['cpu_baseline'] or {}).items() if k!='sample'})
print('roofline', d['roofline']); print({k:(v['ms'],v['tflops'],v['gbs']) for k,v in d['kernel_families'].items()})
PY
tail -3 gpurun_out/r2_bench_run5.err
timeout 900 python bench.py --workload i2vgen --steps 5 --warmup 3 --skip-cpu-baseline > gpurun_out/r2_bench_i2vgen_run5.json 2> gpurun_out/r2_bench_i2vgen_run5.err
echo "bench i2vgen rc=$?"; cut -c1-300 gpurun_out/r2_bench_i2vgen_run5.json; tail -3 gpurun_out/r2_bench_i2vgen_run5.err
