#!/bin/bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
./scripts/ubench/pipes > gpurun_out/r2_ubench_pipes.txt 2>&1; echo "ubench rc=$?"; cat gpurun_out/r2_ubench_pipes.txt
timeout 1800 python -m pytest tests/ -x -q -m gpu -rA > gpurun_out/r2_pytest_gpu_run8.log 2>&1
echo "pytest rc=$?"; grep -E "^(PASSED|FAILED|ERROR)|passed|failed" gpurun_out/r2_pytest_gpu_run8.log | cut -c1-200
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench_sdxl_run8.json 2> gpurun_out/r2_bench_sdxl_run8.err
echo "bench rc=$?"; python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_bench_sdxl_run8.json'))
print('value',round(d['value'],3),'ms',round(d['ms_per_step'],1),'e2e',d['e2e'] and round(d['e2e']['value'],3),'vs_eager',d.get('vs_eager'))
print('cpu', {k:v for k,v in (d['cpu_baseline'] or {}).items() if k!='sample'}); print('clocks', d['clocks'])
PY
K='regex:^(gemm_conv|attention|temporal_attention|gn_|layernorm|add_k|avgpool|cfg_|i2vgen|nchw|nhwc|router|silu|softmax|timestep|upsample2x)'
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k "$K" --csv --log-file gpurun_out/launches_r2.csv python bench.py --steps 1 --warmup 1 --no-graph --skip-cpu-baseline --skip-e2e --skip-profile --skip-eager-baseline > gpurun_out/ncu_bench_r2.log 2>&1
echo "ncu list rc=$?"; wc -l gpurun_out/launches_r2.csv; gzip -f gpurun_out/launches_r2.csv
timeout 300 ncu --set full --clock-control none --import-source on -k regex:^temporal_attention -s 3 -c 1 -f -o gpurun_out/r2_tattn python scripts/prof_kernels.py tattn > gpurun_out/ncu_tattn.log 2>&1; echo "ncu tattn rc=$?"
timeout 600 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/r2_bench_reference_arm.json 2> gpurun_out/r2_bench_reference_arm.err
echo "ref arm rc=$?"; cut -c1-600 gpurun_out/r2_bench_reference_arm.json
