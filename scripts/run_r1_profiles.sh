# Round-1 profile captures (run on the B200 box through gpurun; outputs land in gpurun_out/).
K='regex:^(gemm_conv|attention|gn_|layernorm|temporal_attention|add_k|avgpool|cfg_|i2vgen|nchw|nhwc|router|silu|timestep|upsample2x)'
# launch list of the bench command: duration + DRAM bytes per launch, our kernels only (one eager step after one warm-up)
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k "$K" \
  --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 1 --warmup 1 --no-graph --skip-cpu-baseline \
  --skip-e2e --skip-profile > gpurun_out/ncu_bench.log 2>&1
wc -l gpurun_out/launches_r1.csv
# full-set captures of the dominant kernels at config-2 shapes
for w in geglu lin_res conv attn4k attn1k ln; do
  timeout 200 ncu --set full --clock-control none --import-source on -k "$K" -s 3 -c 1 -o gpurun_out/r1_$w -f \
    python scripts/prof_kernels.py $w > gpurun_out/ncu_r1_$w.log 2>&1
  tail -1 gpurun_out/ncu_r1_$w.log
done
timeout 200 ncu --set full --clock-control none --import-source on -k "$K" -s 3 -c 2 -o gpurun_out/r1_gn -f \
  python scripts/prof_kernels.py gn > gpurun_out/ncu_r1_gn.log 2>&1
tail -1 gpurun_out/ncu_r1_gn.log
