#!/bin/bash
# LayerNorm with bulk-copy row staging + register-resident gamma/beta: parity (misc kernel checks), timing at the three
# SDXL widths against the previous kernel (same box, HBM-rotating inputs), one ncu capture
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out; export PYTHONPATH=$PWD
timeout 600 python tests/kernel_checks.py --group misc > gpurun_out/r2_ln_checks.log 2>&1; echo "checks rc=$?"; tail -24 gpurun_out/r2_ln_checks.log
OLD=$PWD/ctrl_adapter_b200/libctrl_adapter_b200_oldln.so
{
for rep in 1 2; do
for w in ln ln640 ln320; do
  echo "new $(timeout 120 python scripts/prof_kernels.py $w --time 2>&1 | tail -1)"
  echo "old $(CA_B200_LIB=$OLD timeout 120 python scripts/prof_kernels.py $w --time 2>&1 | tail -1)"
done; done
} | tee gpurun_out/r2_ln_time.txt
K='regex:^(layernorm)'
timeout 300 ncu --set full --clock-control none --import-source on -k "$K" -s 3 -c 1 -f -o gpurun_out/r2g_ln python scripts/prof_kernels.py ln > gpurun_out/ncu_r2g_ln.log 2>&1; echo "ncu ln rc=$?"
