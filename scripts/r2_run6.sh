#!/bin/bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
timeout 600 python -m tests.kernel_checks --group attn > gpurun_out/r2_attn_tc.log 2>&1
echo "attn group rc=$?"; grep -c "\[ok" gpurun_out/r2_attn_tc.log; grep "temporal\|FAIL\|EXC" gpurun_out/r2_attn_tc.log | cut -c1-300
for s in tattn tattn14 tattn20; do CA_TATTN_FMA=1 timeout 120 python scripts/prof_kernels.py $s --time 2>&1 | tail -1 | sed "s/^/fma /"; timeout 120 python scripts/prof_kernels.py $s --time 2>&1 | tail -1 | sed "s/^/tc  /"; done | tee gpurun_out/r2_tattn.txt
timeout 600 python -m tests.module_checks --groups video,shapes > gpurun_out/r2_video_tc.log 2>&1
echo "video,shapes rc=$?"; grep "^\[" gpurun_out/r2_video_tc.log | cut -c1-200
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench_sdxl_run6.json 2> gpurun_out/r2_bench_sdxl_run6.err
echo "bench sdxl rc=$?"; tail -2 gpurun_out/r2_bench_sdxl_run6.err
for w in i2vgen svd multi; do
timeout 900 python bench.py --workload $w --steps 5 --warmup 3 --skip-cpu-baseline > gpurun_out/r2_bench_${w}_run6.json 2> gpurun_out/r2_bench_${w}_run6.err
echo "bench $w rc=$?"; tail -2 gpurun_out/r2_bench_${w}_run6.err
done
python - <<'PY'
import json
for w in ("sdxl","i2vgen","svd","multi"):
    try:
        d=json.load(open(f'gpurun_out/r2_bench_{w}_run6.json'))
        print(w, 'value',round(d['value'],3),'ms',round(d['ms_per_step'],1),'e2e',d['e2e'] and round(d['e2e']['value'],3),'eager',d['eager_gpu_baseline'] and d['eager_gpu_baseline'].get('value'),'vs_eager',d.get('vs_eager'),'frac',d['config']['step_frac_of_sustained_peak'], 'roof', d['roofline'] and d['roofline']['frac'])
        if d.get('cpu_baseline'): print('   cpu', {k:v for k,v in d['cpu_baseline'].items() if k!='sample'})
        if d['eager_gpu_baseline'] and d['eager_gpu_baseline'].get('error'): print('   eager err', d['eager_gpu_baseline']['error'])
    except Exception as e: print(w, 'ERR', e)
PY
for w in sdxl i2vgen svd; do
timeout 900 python scripts/drift_report.py $w --out gpurun_out/r2_drift_$w.json 2> gpurun_out/r2_drift_$w.err | tail -1
echo "drift $w rc=$?"; tail -2 gpurun_out/r2_drift_$w.err
done
timeout 900 python scripts/drift_report.py sdxl --guidance-end 0.6 --out gpurun_out/r2_drift_sdxl_end06.json 2> gpurun_out/r2_drift_sdxl_end06.err | tail -1
