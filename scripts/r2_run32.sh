#!/bin/bash
# same-box whole-step A/B of the short-KV attention kernel (kernel-only legs of the bench)
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
for v in short general; do
  s=1; [ "$v" = "general" ] && s=0
  CA_ATTN_SHORT=$s timeout 40 python bench.py --steps 10 --warmup 3 --skip-cpu-baseline --skip-eager-baseline --skip-e2e --skip-profile 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('AB sdxl $v ms', round(d['ms_per_step'],2))"
done | tee gpurun_out/r2_attn_short_step_ab.txt
