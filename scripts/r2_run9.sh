#!/bin/bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
timeout 600 python -m tests.kernel_checks --group misc > gpurun_out/r2_misc_norm.log 2>&1; echo "misc rc=$?"; grep -c "\[ok" gpurun_out/r2_misc_norm.log; grep "FAIL\|EXC" gpurun_out/r2_misc_norm.log | cut -c1-300
timeout 600 python -m tests.kernel_checks --group attn > gpurun_out/r2_attn_q2.log 2>&1; echo "attn rc=$?"; grep -c "\[ok" gpurun_out/r2_attn_q2.log; grep "FAIL\|EXC" gpurun_out/r2_attn_q2.log | cut -c1-300
timeout 300 python -m tests.module_checks --groups shapes > gpurun_out/r2_shapes_q2.log 2>&1; echo "shapes rc=$?"
for s in attn attn4k attn1k attn77; do timeout 120 python scripts/prof_kernels.py $s --time 2>&1 | tail -1; done
for s in ln gn; do timeout 120 python scripts/prof_kernels.py $s --time 2>&1 | tail -1; done
timeout 900 python bench.py --steps 10 --warmup 3 --skip-cpu-baseline --skip-eager-baseline > gpurun_out/r2_bench_sdxl_run9.json 2> gpurun_out/r2_bench_sdxl_run9.err
echo "bench rc=$?"; python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_bench_sdxl_run9.json'))
print('value',round(d['value'],3),'ms',round(d['ms_per_step'],1)); print({k:(v['ms'],v['gbs']) for k,v in d['kernel_families'].items() if k in ('layernorm','groupnorm')})
PY
timeout 900 python bench.py --workload i2vgen --steps 5 --warmup 3 --skip-cpu-baseline --skip-eager-baseline > gpurun_out/r2_bench_i2vgen_run9.json 2> gpurun_out/r2_bench_i2vgen_run9.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_bench_i2vgen_run9.json'))
print('i2vgen value',round(d['value'],3),'ms',round(d['ms_per_step'],1)); print({k:(v['ms'],v['gbs']) for k,v in d['kernel_families'].items() if k in ('layernorm','groupnorm')})
PY
