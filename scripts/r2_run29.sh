#!/bin/bash
# final check of the committed tree: GPU suite, smoke, default bench line
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
timeout 1500 python -m pytest tests/ -x -q -m gpu -rA > gpurun_out/r2_pytest_gpu_final.log 2>&1
echo "pytest rc=$?"; grep -E "^(PASSED|FAILED|ERROR)|passed|failed" gpurun_out/r2_pytest_gpu_final.log | cut -c1-160
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/r2_bench_sdxl_final.json 2> gpurun_out/r2_bench_sdxl_final.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_bench_sdxl_final.json'))
print('value',round(d['value'],3),'ms',round(d['ms_per_step'],1),'e2e',d['e2e'],'vs_eager',d.get('vs_eager'),'roof',d['roofline'] and d['roofline']['frac'],'cpu',d['cpu_baseline'] and d['cpu_baseline']['value'], d['clocks'], 'launches', d['gpu_launches'])
PY
