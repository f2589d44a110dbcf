#!/bin/bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
timeout 600 python -m tests.kernel_checks --group gemm > gpurun_out/r2_gemm_thr.log 2>&1; echo "gemm ok=$(grep -c '\[ok' gpurun_out/r2_gemm_thr.log) rc=$?"
timeout 600 python -m tests.kernel_checks --group conv > gpurun_out/r2_conv_thr.log 2>&1; echo "conv ok=$(grep -c '\[ok' gpurun_out/r2_conv_thr.log)"
for w in sdxl i2vgen svd multi; do
timeout 900 python bench.py --workload $w --steps 8 --warmup 3 --skip-cpu-baseline --skip-eager-baseline > gpurun_out/r2_bench_${w}_run18.json 2> gpurun_out/r2_bench_${w}_run18.err
python - <<PY
import json
d=json.load(open('gpurun_out/r2_bench_${w}_run18.json'))
print('$w', 'value',round(d['value'],3),'ms',round(d['ms_per_step'],1),'e2e',d['e2e'] and round(d['e2e']['value'],3), {k:(v['ms'],v['tflops'] or v['gbs']) for k,v in list(d['kernel_families'].items())[:5]})
PY
done
