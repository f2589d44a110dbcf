K='regex:^(gemm_conv|attention|gn_|layernorm|temporal_attention|add_k|avgpool|cfg_|i2vgen|nchw|nhwc|router|silu|timestep|upsample2x)'
timeout 200 python -m tests.kernel_checks --group attn 2>&1 | grep -v "^\[ok" | tail -4
for P in 0 2 3 4; do echo "POLY=$P"; for w in attn attn1k; do CA_ATTN_POLY=$P timeout 120 python scripts/prof_kernels.py $w --time 2>&1 | tail -1; done; done
timeout 600 python bench.py > gpurun_out/bench_r1_default.json 2> gpurun_out/bench_r1_default.err; tail -c 600 gpurun_out/bench_r1_default.json
timeout 400 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k "$K" --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 1 --warmup 1 --no-graph --skip-cpu-baseline --skip-e2e --skip-profile > gpurun_out/ncu_bench.log 2>&1; wc -l gpurun_out/launches_r1.csv
for w in geglu lin_res conv attn4k; do timeout 200 ncu --set full --clock-control none --import-source on -k "$K" -s 3 -c 1 -o gpurun_out/r1_$w -f python scripts/prof_kernels.py $w > gpurun_out/ncu_r1_$w.log 2>&1; tail -1 gpurun_out/ncu_r1_$w.log; done
