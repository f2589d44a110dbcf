#!/bin/bash
# after the LayerNorm rewrite: full GPU suite, smoke, same-box whole-step A/B against the previous LayerNorm, final lines
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
timeout 1800 python -m pytest tests/ -x -q -m gpu -rA > gpurun_out/r2_pytest_gpu_final.log 2>&1
echo "pytest rc=$?"; grep -E "^(PASSED|FAILED|ERROR)|passed|failed" gpurun_out/r2_pytest_gpu_final.log | cut -c1-160
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
OLD=$PWD/ctrl_adapter_b200/libctrl_adapter_b200_oldln.so
for rep in 1 2; do
for w in sdxl i2vgen; do
 for v in new old; do
  lib=""; [ "$v" = "old" ] && lib="$OLD"
  CA_B200_LIB=${lib:-$PWD/ctrl_adapter_b200/libctrl_adapter_b200.so} timeout 600 python bench.py --workload $w --steps 10 --warmup 3 --skip-cpu-baseline --skip-eager-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('AB $w $v ms', round(d['ms_per_step'],2))"
 done
done
done | tee gpurun_out/r2_layernorm_step_ab.txt
for w in sdxl i2vgen svd multi; do
extra=""; [ "$w" != "sdxl" ] && extra="--skip-cpu-baseline"
timeout 900 python bench.py --workload $w --steps 10 --warmup 3 $extra > gpurun_out/r2_bench_${w}_final.json 2> gpurun_out/r2_bench_${w}_final.err
python - <<PY
import json
d=json.load(open('gpurun_out/r2_bench_${w}_final.json'))
print('$w', 'value',round(d['value'],3),'ms',round(d['ms_per_step'],1),'e2e',d['e2e'] and round(d['e2e']['value'],3),'vs_eager',d.get('vs_eager') and round(d['vs_eager'],3),'frac',d['config']['step_frac_of_sustained_peak'],'roof',d['roofline'] and (d['roofline']['kernel'][:12], d['roofline']['frac']), 'launches', d['launches_per_step'])
print('   ', {k:(v['ms'],v['tflops'] or v['gbs']) for k,v in list(d['kernel_families'].items())[:6]}, d['clocks'])
PY
done
