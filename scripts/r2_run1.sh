#!/bin/bash
# round-2 first GPU run: the groups round 1 hid behind xfail + the prepared experiments' kernel checks + base numbers
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 1200 python -m tests.module_checks --groups shapes,svd,sparse,svd_loop,fold --json gpurun_out/r2_pending.json > gpurun_out/r2_pending.log 2>&1
echo "pending rc=$?"
grep -E "^\[(ok|FAIL)" gpurun_out/r2_pending.log | cut -c1-300
for s in attn attn4k attn1k attn77 conv lin_small lin_res geglu ln gn; do timeout 120 python scripts/prof_kernels.py $s --time 2>&1 | tail -1; done | tee gpurun_out/r2_kernels_base.txt
CA_GEMM_BN320=1 timeout 600 python -m tests.kernel_checks --group conv --json gpurun_out/r2_bn320_conv.json > gpurun_out/r2_bn320_conv.log 2>&1
echo "bn320 conv rc=$?"; tail -3 gpurun_out/r2_bn320_conv.log
CA_GEMM_BN320=1 timeout 600 python -m tests.kernel_checks --group gemm --json gpurun_out/r2_bn320_gemm.json > gpurun_out/r2_bn320_gemm.log 2>&1
echo "bn320 gemm rc=$?"; tail -3 gpurun_out/r2_bn320_gemm.log
timeout 600 python bench.py --steps 10 --warmup 3 --skip-cpu-baseline > gpurun_out/r2_bench_base.json 2> gpurun_out/r2_bench_base.err
echo "bench base rc=$?"; cut -c1-700 gpurun_out/r2_bench_base.json
CA_GEMM_BN320=1 timeout 600 python bench.py --steps 10 --warmup 3 --skip-cpu-baseline --skip-e2e > gpurun_out/r2_bench_bn320.json 2> gpurun_out/r2_bench_bn320.err
echo "bench bn320 rc=$?"; cut -c1-300 gpurun_out/r2_bench_bn320.json
CA_FOLD_SMALL_CONV=1 timeout 600 python bench.py --steps 10 --warmup 3 --skip-cpu-baseline --skip-e2e > gpurun_out/r2_bench_fold.json 2> gpurun_out/r2_bench_fold.err
echo "bench fold rc=$?"; cut -c1-300 gpurun_out/r2_bench_fold.json
