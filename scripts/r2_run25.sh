#!/bin/bash
# SVD temporal VAE decoder: kernel check of frame_conv_small, module parity (vae group), decode time at 14 x 576 x 1024
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out; export PYTHONPATH=$PWD
timeout 600 python tests/kernel_checks.py --group misc > gpurun_out/r2_misc_checks.log 2>&1; echo "misc rc=$?"; grep -E "FAIL|frame_conv|checks ok" gpurun_out/r2_misc_checks.log
timeout 900 python -m pytest tests/test_modules_gpu.py -x -q -m gpu -k vae -rA > gpurun_out/r2_vae_group.log 2>&1; echo "vae group rc=$?"; tail -5 gpurun_out/r2_vae_group.log
python - <<'PY' 2>&1 | tee gpurun_out/r2_vae_temporal_time.txt
import json, torch
print(json.dumps([r for r in json.load(open('gpurun_out/gpu_group_vae.json'))], indent=0)[:1500]) if __import__('os').path.exists('gpurun_out/gpu_group_vae.json') else None
from ctrl_adapter_b200.vae import AutoencoderKLTemporalDecoder, AutoencoderKL, svd_decode_latents
torch.manual_seed(0)
with torch.device("cuda"):
    m = AutoencoderKLTemporalDecoder().to(torch.bfloat16).eval()
lat = torch.randn(1, 14, 4, 72, 128, device="cuda") * 0.18215
for chunk in (14, 7):
    for _ in range(2):
        v = svd_decode_latents(m, lat, 14, chunk)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); v = svd_decode_latents(m, lat, 14, chunk); e1.record(); torch.cuda.synchronize()
    print(f"svd temporal VAE decode 14 x 576 x 1024, chunk {chunk}: {e0.elapsed_time(e1):.1f} ms, out {tuple(v.shape)} finite={bool(torch.isfinite(v).all())} peak_mem={torch.cuda.max_memory_allocated()/2**30:.1f} GiB")
PY
