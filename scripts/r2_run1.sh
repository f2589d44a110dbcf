#!/bin/bash
# round-2 first GPU run: the groups round 1 hid behind xfail + the prepared experiments' kernel checks
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 1200 python -m tests.module_checks --groups svd,sparse,svd_loop,fold,shapes --json gpurun_out/r2_pending.json > gpurun_out/r2_pending.log 2>&1
echo "pending rc=$?"
tail -40 gpurun_out/r2_pending.log
CA_GEMM_BN320=1 timeout 600 python -m tests.kernel_checks --group conv --json gpurun_out/r2_bn320_conv.json > gpurun_out/r2_bn320_conv.log 2>&1
echo "bn320 conv rc=$?"; tail -5 gpurun_out/r2_bn320_conv.log
CA_GEMM_BN320=1 timeout 600 python -m tests.kernel_checks --group gemm --json gpurun_out/r2_bn320_gemm.json > gpurun_out/r2_bn320_gemm.log 2>&1
echo "bn320 gemm rc=$?"; tail -5 gpurun_out/r2_bn320_gemm.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench_base.json 2> gpurun_out/r2_bench_base.err
echo "bench base rc=$?"; cat gpurun_out/r2_bench_base.json | cut -c1-600
CA_GEMM_BN320=1 timeout 600 python bench.py --steps 10 --warmup 3 --skip-cpu-baseline --skip-e2e > gpurun_out/r2_bench_bn320.json 2> gpurun_out/r2_bench_bn320.err
echo "bench bn320 rc=$?"; cat gpurun_out/r2_bench_bn320.json | cut -c1-300
CA_FOLD_SMALL_CONV=1 timeout 600 python bench.py --steps 10 --warmup 3 --skip-cpu-baseline --skip-e2e > gpurun_out/r2_bench_fold.json 2> gpurun_out/r2_bench_fold.err
echo "bench fold rc=$?"; cat gpurun_out/r2_bench_fold.json | cut -c1-300
