#!/usr/bin/env python
"""Benchmark of the Ctrl-Adapter denoising hot path on B200 (contract: see the task brief / DESIGN.md section 6).

  python bench.py --gpus N --steps K --warmup W            ours: CUDA path (this repo's kernels)
  python bench.py --impl reference --gpus N ...            reference arm: the oracle restatement of the reference's
                                                           PyTorch path on the host CPU cores (diffusers is not
                                                           installable here, so the reference itself cannot run)

Workload = BASELINE.json configs[1]: SDXL + depth ControlNet + Ctrl-Adapter, 1024x1024, batch 8 (CFG -> 16
frame-samples), synthetic latents / embeddings of the named shapes, random weights of the real architectures
(SD1.5 ControlNet 361 M, SDXL adapter 184 M, SDXL UNet 2.57 B parameters, bf16).
One "step" = one iteration of the pipeline loop for the whole batch:
[pool] -> ControlNet -> Ctrl-Adapter -> UNet(+injection) -> CFG -> Euler update.

Multi-GPU (--gpus N under torchrun): batch-axis sharding, every rank runs its own batch of 8 with no per-step
communication (weak scaling) and one NCCL all-gather of the final latents after the last step.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

BF16 = torch.bfloat16
# algorithmic FLOPs per frame-sample (SURVEY.md section 8d / BASELINE.md section 2): ControlNet@64^2, SDXL adapter, SDXL UNet@128^2
TFLOP_PER_SAMPLE = 0.2833 + 2.2565 + 6.761


def parse():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=10)
    p.add_argument("--warmup", type=int, default=3)
    p.add_argument("--impl", default="ours", choices=["ours", "reference"])
    p.add_argument("--batch", type=int, default=8, help="images per GPU (BASELINE config: 8)")
    p.add_argument("--res", type=int, default=1024)
    p.add_argument("--workload", default="sdxl", choices=["sdxl", "i2vgen", "svd"],
                   help="sdxl = BASELINE.json configs[1] (headline); i2vgen = configs[2] (I2VGen-XL 16f 512x512, batch 4); "
                        "svd = configs[3] (SVD 14f 576x1024, batch 2 per GPU)")
    p.add_argument("--no-graph", action="store_true", help="eager launches instead of CUDA-graph replay")
    p.add_argument("--skip-cpu-baseline", action="store_true")
    p.add_argument("--skip-e2e", action="store_true")
    p.add_argument("--eager-baseline", action="store_true",
                   help="also time the oracle (restated reference) as eager bf16-autocast PyTorch on this GPU")
    p.add_argument("--skip-profile", action="store_true", help="skip the per-kernel CUDA-event profile (roofline block)")
    p.add_argument("--cpu-sample-res", type=int, default=512, help="resolution of the one-frame-sample pass timed by the CPU legs")
    return p.parse_args()


# ----------------------------------------------------------------------------------------------------
def synthetic_inputs(batch, res, device, seed):
    g = torch.Generator(device="cpu").manual_seed(seed)
    lat = res // 8
    r = lambda *s: torch.randn(*s, generator=g)  # noqa: E731
    return dict(
        latents=r(batch, 4, lat, lat).to(device),
        prompt_embeds=r(2 * batch, 77, 2048).to(device),
        add_text_embeds=r(2 * batch, 1280).to(device),
        add_time_ids=torch.tensor([[res, res, 0, 0, res, res]] * (2 * batch), dtype=torch.float32).to(device),
        controlnet_prompt_embeds=r(2 * batch, 77, 768).to(device),
        control_images=torch.rand(2 * batch, 3, 512, 512, generator=g).to(device),
    )


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.idx)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm = sorted(float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit())
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            if len(r) >= 9:
                for nme, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(nme)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return d, "measured (MEASURED_PEAKS.json)"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback (B200_PROFILING.md)"


# ----------------------------------------------------------------------------------------------------
def _oracle_frame_sample(cn, ad, un, inp, lat, t, res):
    """One frame-sample through ControlNet (at res/2) -> adapter -> UNet, restated reference modules."""
    F = torch.nn.functional
    down, mid = cn(F.adaptive_avg_pool2d(lat, (res // 16, res // 16)), t,
                   encoder_hidden_states=inp["controlnet_prompt_embeds"][:1],
                   controlnet_cond=F.adaptive_avg_pool2d(inp["control_images"][:1], (res // 2, res // 2)),
                   conditioning_scale=1.0, return_dict=False)
    da, _ = ad(down, num_frames=1, timestep=t, encoder_hidden_states=inp["prompt_embeds"][:1])
    un(lat, t, encoder_hidden_states=inp["prompt_embeds"][:1],
       added_cond_kwargs={"text_embeds": inp["add_text_embeds"][:1], "time_ids": inp["add_time_ids"][:1]},
       down_block_additional_residuals=da, mid_block_additional_residual=0)


def _oracle_flops_per_frame_sample(res):
    """Exact matmul / conv / attention FLOPs of one frame-sample at `res`, counted on meta tensors (no compute)."""
    from torch.utils.flop_counter import FlopCounterMode
    from oracle.adapter import ControlNetAdapter
    from oracle.cases import ADAPTER_SDXL_KW, CONTROLNET_KW
    from oracle.controlnet import ControlNetModel
    from oracle.unet_sdxl import UNet2DConditionModel
    with torch.device("meta"):
        cn, ad, un = ControlNetModel(**CONTROLNET_KW).eval(), ControlNetAdapter(**ADAPTER_SDXL_KW).eval(), UNet2DConditionModel().eval()
    inp = {k: (v.to("meta") if torch.is_tensor(v) else v) for k, v in synthetic_inputs(1, res, "cpu", 1234).items()}
    with FlopCounterMode(display=False) as fc, torch.no_grad():
        _oracle_frame_sample(cn, ad, un, inp, inp["latents"][:1], torch.tensor(500.0), res)
    return float(fc.get_total_flops())


def cpu_reference_step_rate(sample_res, res, steps, warmup, budget_s=60.0):
    """Oracle (restated reference PyTorch path) on the host CPU, fp32 eager, all threads, on a BOUNDED sample.
    One full step (16 frame-samples at 1024x1024) takes ~45 minutes on a 128-thread Xeon, so a timed pass is ONE
    frame-sample at `sample_res` (default 512: ~20-40 s) and the rate is scaled to the full step by the exact
    FLOP ratio (torch FlopCounterMode on meta tensors): steps/s = 1 / (t_pass * F_step / F_pass).
    Returns (steps_per_s of the 16-frame-sample workload at `res`, info dict)."""
    from oracle.adapter import ControlNetAdapter
    from oracle.cases import ADAPTER_SDXL_KW, CONTROLNET_KW
    from oracle.controlnet import ControlNetModel
    from oracle.pipeline_sdxl import EulerDiscreteScheduler
    from oracle.unet_sdxl import UNet2DConditionModel
    cores = torch.get_num_threads()  # torch's default: one thread per physical core it may use
    torch.manual_seed(0)
    t0 = time.time()
    cn = ControlNetModel(**CONTROLNET_KW).eval()
    ad = ControlNetAdapter(**ADAPTER_SDXL_KW).eval()
    un = UNet2DConditionModel().eval()
    build_s = time.time() - t0
    f_pass = _oracle_flops_per_frame_sample(sample_res)
    f_step = 16.0 * (f_pass if sample_res == res else _oracle_flops_per_frame_sample(res))
    sch = EulerDiscreteScheduler()
    sch.set_timesteps(50)
    inp = synthetic_inputs(1, sample_res, "cpu", 1234)
    lat = (inp["latents"] * sch.init_noise_sigma)[:1]

    times, done = [], 0
    t_start = time.time()
    for i in range(warmup + steps):
        t1 = time.time()
        with torch.no_grad():
            _oracle_frame_sample(cn, ad, un, inp, sch.scale_model_input(lat, i % 50), sch.timesteps[i % 50], sample_res)
        dt = time.time() - t1
        done += 1
        if i >= warmup or (time.time() - t_start > budget_s):
            times.append(dt)  # a warm-up pass is promoted to a timed one when the budget is already spent
        if time.time() - t_start > budget_s and len(times) >= 1:
            break
    ms = 1000.0 * sum(times) / len(times)
    rate = 1.0 / ((ms / 1000.0) * (f_step / f_pass))
    cpu_model = ""
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                cpu_model = line.split(":", 1)[1].strip()
                break
    except Exception:
        pass
    info = {"value": rate, "unit": "steps/s", "cores": cores, "kind": "port",
            "sample": f"one frame-sample (ControlNet -> adapter -> UNet) at {sample_res}x{sample_res}, fp32 eager, "
                      f"{len(times)} timed pass(es) of {done} executed, {ms:.0f} ms per pass = {f_pass / 1e12:.3f} TFLOP; scaled "
                      f"to the 16-frame-sample {res}x{res} step ({f_step / 1e12:.1f} TFLOP) by the FLOP ratio; oracle "
                      f"restatement of the reference path (the reference itself needs diffusers, not installable here); "
                      f"cpu: {cpu_model}",
            "ms_per_sample_pass": ms, "sample_tflop": f_pass / 1e12, "step_tflop": f_step / 1e12,
            "cpu_tflops": f_pass / ms / 1e9, "model_build_s": build_s}
    return rate, info


def run_reference_arm(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    rate, info = cpu_reference_step_rate(a.cpu_sample_res, a.res, a.steps, a.warmup, budget_s=150.0)
    line = {"impl": "reference", "metric": "denoising steps/sec (SDXL 1024x1024 + depth ControlNet + Ctrl-Adapter, batch 8)",
            "value": rate, "unit": "steps/s", "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": 1000.0 / rate if rate > 0 else None, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"SDXL+depth ControlNet+Ctrl-Adapter {a.res}x{a.res} batch 8 (CFG: 16 frame-samples), "
                                   "one pipeline-loop iteration per step", "sample": info["sample"]},
            "cpu_baseline": info,
            "e2e": {"value": rate, "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


# ----------------------------------------------------------------------------------------------------
_RESULT_FD = None


def _quiet_stdout():
    """The contract is ONE JSON line on stdout: route everything libraries print there (e.g. NCCL's version banner)
    to stderr and keep the real stdout for emit()."""
    global _RESULT_FD
    if _RESULT_FD is None:
        sys.stdout.flush()
        _RESULT_FD = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    sys.stdout.flush()
    os.write(_RESULT_FD if _RESULT_FD is not None else 1, (json.dumps(line) + "\n").encode())


def main():
    a = parse()
    _quiet_stdout()
    if a.impl == "reference":
        run_reference_arm(a)
        return
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py (ours) needs a CUDA device: the hot path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)

    from ctrl_adapter_b200 import _lib, ops
    from ctrl_adapter_b200.adapter import ControlNetAdapter
    from ctrl_adapter_b200.controlnet import ControlNetModel
    from ctrl_adapter_b200.pipeline_sdxl import SDXLControlNetAdapterLoop
    from ctrl_adapter_b200.unet_sdxl import UNet2DConditionModel
    _lib.check(_lib.load().ca_device_ok(), "ca_device_ok")

    torch.manual_seed(1234 + rank)
    with torch.device(dev):
        cn = ControlNetModel(cross_attention_dim=768)
    # zero-initialised ControlNet heads would make every residual exactly 0: give them random values
    for m in list(cn.controlnet_down_blocks) + [cn.controlnet_mid_block, cn.controlnet_cond_embedding.conv_out]:
        torch.nn.init.normal_(m.weight, std=0.02)
    if a.workload == "sdxl":
        with torch.device(dev):
            ad = ControlNetAdapter("sdxl", num_blocks=1, num_frames=1, cross_attention_dim=2048, add_spatial_resnet=True,
                                   add_spatial_transformer=True, add_adapter_location_A=True,
                                   add_adapter_location_B=True, add_adapter_location_C=True)
            un = UNet2DConditionModel()
        cn, ad, un = (m.to(BF16).eval() for m in (cn, ad, un))
        loop = SDXLControlNetAdapterLoop(cn, ad, un, num_inference_steps=50, guidance_scale=5.0,
                                         controlnet_conditioning_scale=1.0)
        inp = synthetic_inputs(a.batch, a.res, dev, 1234 + rank)
        loop.prepare(**inp)
        n_samples = 2 * a.batch
        tflop_per_sample = TFLOP_PER_SAMPLE
        wl_name = (f"SDXL+depth ControlNet+Ctrl-Adapter {a.res}x{a.res}, batch {a.batch} per GPU "
                   f"(CFG: {n_samples} frame-samples), one pipeline-loop iteration per step")
        base_cfg = "BASELINE.json configs[1]"
    elif a.workload == "svd":
        from ctrl_adapter_b200.pipeline_svd import SVDControlNetAdapterLoop
        from ctrl_adapter_b200.unet_svd import UNetSpatioTemporalConditionModel
        frames, vb = 14, (a.batch if a.batch != 8 else 2)
        lh, lw = 72, 128  # 576 x 1024 video -> 72 x 128 latents; needs use_size_512=False (cfg 4)
        with torch.device(dev):
            ad = ControlNetAdapter("svd", num_blocks=1, num_frames=frames, cross_attention_dim=1024,
                                   add_spatial_resnet=True, add_temporal_resnet=True, add_spatial_transformer=True,
                                   add_temporal_transformer=True, add_adapter_location_A=True,
                                   add_adapter_location_B=True, add_adapter_location_C=True,
                                   add_adapter_location_D=True, add_adapter_location_M=True)
            un = UNetSpatioTemporalConditionModel(num_attention_heads=(5, 10, 20, 20), num_frames=frames)
        cn, ad, un = (m.to(BF16).eval() for m in (cn, ad, un))
        loop = SVDControlNetAdapterLoop(cn, ad, un, num_inference_steps=25, min_guidance_scale=1.0, max_guidance_scale=3.0,
                                        use_size_512=False, skip_conv_in=True)
        g = torch.Generator(device="cpu").manual_seed(1234 + rank)
        r = lambda *s_: torch.randn(*s_, generator=g).to(dev)  # noqa: E731
        n_samples = 2 * vb * frames
        il = r(vb, frames, 4, lh, lw)
        loop.prepare(latents=r(vb, frames, 4, lh, lw), image_latents=torch.cat([torch.zeros_like(il), il]),
                     image_embeddings=torch.cat([torch.zeros(vb, 1, 1024, device=dev), r(vb, 1, 1024)]),
                     added_time_ids=torch.tensor([[13.0, 127.0, 0.02]] * (2 * vb), device=dev),
                     controlnet_prompt_embeds=r(n_samples, 77, 768),
                     control_images=torch.rand(n_samples, 3, 8 * lh, 8 * lw, generator=g).to(dev))
        tflop_per_sample = 0.7735 + 1.679 + 3.192  # ControlNet, video adapter, SVD UNet @72x128 (rough)
        wl_name = (f"SVD+depth ControlNet+Ctrl-Adapter 14 frames 576x1024, batch {vb} per GPU "
                   f"(CFG: {n_samples} frame-samples), one pipeline-loop iteration per step")
        base_cfg = "BASELINE.json configs[3]"
    else:
        from ctrl_adapter_b200.pipeline_i2vgen import I2VGenXLControlNetAdapterLoop
        from ctrl_adapter_b200.unet_i2vgen import I2VGenXLUNet
        frames, vb = 16, (a.batch if a.batch != 8 else 4)
        with torch.device(dev):
            ad = ControlNetAdapter("i2vgenxl", num_blocks=1, num_frames=frames, cross_attention_dim=1024,
                                   add_spatial_resnet=True, add_temporal_resnet=True, add_spatial_transformer=True,
                                   add_temporal_transformer=True, add_adapter_location_A=True,
                                   add_adapter_location_B=True, add_adapter_location_C=True,
                                   add_adapter_location_D=True, add_adapter_location_M=True)
            un = I2VGenXLUNet()
        cn, ad, un = (m.to(BF16).eval() for m in (cn, ad, un))
        loop = I2VGenXLControlNetAdapterLoop(cn, ad, un, None, num_inference_steps=50, guidance_scale=9.0)
        g = torch.Generator(device="cpu").manual_seed(1234 + rank)
        r = lambda *s_: torch.randn(*s_, generator=g).to(dev)  # noqa: E731
        n_samples = 2 * vb * frames
        loop.prepare(latents=r(vb, 4, frames, 64, 64), prompt_embeds=r(2 * vb, 77, 1024),
                     image_latents=r(2 * vb, 4, frames, 64, 64), image_embeddings=r(2 * vb, 1, 1024),
                     fps=torch.full((2 * vb,), 16.0, device=dev), controlnet_prompt_embeds=r(n_samples, 77, 768),
                     control_images=torch.rand(n_samples, 3, 512, 512, generator=g).to(dev))
        tflop_per_sample = 0.2833 + 0.6564 + 1.308  # ControlNet, video adapter, I2VGen-XL UNet (rough)
        wl_name = (f"I2VGen-XL+depth ControlNet+Ctrl-Adapter 16 frames 512x512, batch {vb} per GPU "
                   f"(CFG: {n_samples} frame-samples), one pipeline-loop iteration per step")
        base_cfg = "BASELINE.json configs[2]"

    use_graph = not a.no_graph
    nsteps = loop.num_inference_steps  # the schedule wraps around when more steps are timed than it has
    l0 = ops.PROFILER.launches
    loop.step(0)  # packs weights, sets kernel attributes
    launches_per_step = ops.PROFILER.launches - l0
    torch.cuda.synchronize()
    if use_graph:
        loop.capture(warmup=1)
    stepfn = loop.step_graph if use_graph else loop.step
    for i in range(a.warmup):
        stepfn(i % nsteps)
    torch.cuda.synchronize()
    if dist:
        dist.barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for i in range(a.steps):
        stepfn((a.warmup + i) % nsteps)
    if dist:  # the single collective of the job: gather every rank's final latents (C1 in SURVEY.md)
        from ctrl_adapter_b200.distributed import gather_latents
        gathered = gather_latents(loop.latents, loop.latents.shape[0] * world)
    ev1.record()
    torch.cuda.synchronize()
    if dist:
        dist.barrier()
    ms_total = ev0.elapsed_time(ev1)
    clocks = sampler.stop()
    if dist:
        tt = torch.tensor([ms_total], device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms_total = float(tt)
    ms_step = ms_total / a.steps
    value = world * a.steps / (ms_total / 1000.0)
    finite = bool(torch.isfinite(loop.latents).all())

    # ---- e2e: public module API, host buffers, H2D/D2H of the step's inputs/outputs inside the timed region ----
    e2e = None
    if not a.skip_e2e:
        host_lat = torch.empty(loop.latents.shape, dtype=torch.float32).pin_memory()
        host_lat.copy_(loop.latents.cpu())
        host_in = torch.empty(loop.model_in.shape, dtype=BF16).pin_memory()
        host_in.copy_(loop.model_in.cpu())
        ke = max(3, min(a.steps, 10))
        for _ in range(2):
            loop.step(0)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(ke):
            loop.latents.copy_(host_lat, non_blocking=True)       # H2D: this step's latents
            loop.model_in.copy_(host_in, non_blocking=True)       # H2D: scaled model input
            loop.step(i % nsteps)                                 # ControlNet / adapter / UNet module forward()s
            host_lat.copy_(loop.latents, non_blocking=True)       # D2H: the step's result
            host_in.copy_(loop.model_in, non_blocking=True)
            torch.cuda.current_stream().synchronize()             # the host consumes the result every step
        e1.record()
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        ms_e = e0.elapsed_time(e1)
        if dist:
            tt = torch.tensor([ms_e], device=dev)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            ms_e = float(tt)
        nbytes = host_lat.numel() * 4 + host_in.numel() * 2
        e2e = {"value": world * ke / (ms_e / 1000.0), "unit": "steps/s", "h2d_bytes_per_step": nbytes,
               "d2h_bytes_per_step": nbytes, "steps": ke, "ms_per_step": ms_e / ke, "wall_ms_per_step": 1000 * wall / ke,
               "path": "eager module forward() calls through the C ABI, pinned host latents copied in/out every step"}

    # ---- per-kernel-family CUDA-event profile of one eager step -> roofline of the dominant kernel ----
    roofline, families = None, None
    peaks, peak_src = measured_peaks()
    if not a.skip_profile and rank == 0:
        ops.PROFILER.start()
        loop.step(0)
        ops.PROFILER.stop()
        fam = ops.PROFILER.summary()
        try:
            os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
            json.dump(ops.PROFILER.launches_table(80), open(os.path.join(ROOT, "gpurun_out", "launch_table.json"), "w"), indent=0)
        except Exception:
            pass
        tot = sum(v["ms"] for v in fam.values())
        families = {k: {"launches": v["launches"], "ms": round(v["ms"], 3), "share": round(v["ms"] / tot, 4),
                        "tflops": round(v["flops"] / (v["ms"] * 1e9), 2) if v["ms"] > 0 and v["flops"] else None,
                        "gbs": round(v["bytes"] / (v["ms"] * 1e6), 1) if v["ms"] > 0 and v["bytes"] else None}
                    for k, v in sorted(fam.items(), key=lambda kv: -kv[1]["ms"])}
        top = max(fam.items(), key=lambda kv: kv[1]["ms"])
        # DRAM bytes per launch of the dominant kernel: ncu dram__bytes_{read,write}.sum over every launch of one step
        # (profiles/r1_traffic_<workload>.json, made by scripts/launch_share.py from the committed launch list)
        traffic = None
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", f"r1_traffic_{a.workload}.json")))
            kn = {"gemm": "gemm_conv_kernel", "attention": "attention_kernel"}.get(top[0])
            if kn in tj:
                traffic = round(tj[kn]["dram_bytes_per_launch"], 0)
        except Exception:
            traffic = None
        if top[0] in ("gemm", "attention"):
            ach = top[1]["flops"] / (top[1]["ms"] * 1e9)
            peak = peaks["bf16_tflops_sustained"]
            roofline = {"kernel": {"gemm": "gemm_conv_kernel (tcgen05 multi-tap GEMM / implicit conv)",
                                   "attention": "attention_kernel (tcgen05 flash attention)"}[top[0]],
                        "bound": "tensor", "achieved": round(ach, 2), "peak": peak, "unit": "TFLOP/s",
                        "frac": round(ach / peak, 4), "traffic": traffic,
                        "how": f"sum of algorithmic FLOPs of the {top[1]['launches']} launches of one step / sum of their "
                               f"CUDA-event durations ({top[1]['ms']:.1f} ms = {100 * top[1]['ms'] / tot:.0f}% of the step); "
                               f"peak = bf16_tflops_sustained, {peak_src}"}
        else:
            ach = top[1]["bytes"] / (top[1]["ms"] * 1e6)
            roofline = {"kernel": top[0], "bound": "hbm", "achieved": round(ach, 1), "peak": peaks["hbm_gbs"],
                        "unit": "GB/s", "frac": round(ach / peaks["hbm_gbs"], 4), "traffic": None,
                        "how": f"algorithmic bytes / CUDA-event time of one step; peak {peak_src}"}

    eager_gpu = None
    if a.eager_baseline and rank == 0 and a.workload == "sdxl":
        # "reference single-GPU eager PyTorch" (BASELINE.md section 3): oracle modules, bf16 params, torch.autocast,
        # default SDPA backend, no compile, same synthetic tensors and batch
        try:
            del loop
            torch.cuda.empty_cache()
            from oracle.adapter import ControlNetAdapter as OA
            from oracle.cases import ADAPTER_SDXL_KW, CONTROLNET_KW
            from oracle.controlnet import ControlNetModel as OC
            from oracle.pipeline_sdxl import EulerDiscreteScheduler, sdxl_step
            from oracle.unet_sdxl import UNet2DConditionModel as OU
            with torch.device(dev):
                ocn, oad, oun = OC(**CONTROLNET_KW), OA(**ADAPTER_SDXL_KW), OU()
            ocn, oad, oun = (m.to(BF16).eval() for m in (ocn, oad, oun))
            sch = EulerDiscreteScheduler()
            sch.set_timesteps(50, device=dev)
            ei = {k: (v.to(BF16) if v.is_floating_point() else v) for k, v in inp.items()}
            lat = ei["latents"] * sch.init_noise_sigma

            def estep(i, lat):
                with torch.no_grad(), torch.autocast("cuda", dtype=BF16):
                    return sdxl_step(ocn, oad, oun, sch, i, lat, ei["prompt_embeds"], ei["add_text_embeds"],
                                     ei["add_time_ids"], ei["controlnet_prompt_embeds"], ei["control_images"])
            for i in range(2):
                lat = estep(i, lat)
            torch.cuda.synchronize()
            q0, q1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ne = 5
            q0.record()
            for i in range(ne):
                lat = estep(2 + i, lat)
            q1.record()
            torch.cuda.synchronize()
            ems = q0.elapsed_time(q1) / ne
            eager_gpu = {"value": 1000.0 / ems, "unit": "steps/s", "ms_per_step": ems, "steps": ne,
                         "what": "oracle restatement of the reference loop, eager PyTorch bf16 autocast on this GPU "
                                 "(cuDNN / cuBLAS / SDPA), same batch and shapes"}
            del ocn, oad, oun
            torch.cuda.empty_cache()
        except Exception as e:
            eager_gpu = {"value": None, "error": repr(e)[:300]}
        loop = None

    cpu_baseline = None
    if rank == 0 and world == 1 and not a.skip_cpu_baseline and a.workload == "sdxl":
        try:
            del loop
            torch.cuda.empty_cache()
            _, cpu_baseline = cpu_reference_step_rate(a.cpu_sample_res, a.res, 1, 0, budget_s=60.0)
        except Exception as e:  # the CPU leg must never hide the GPU numbers
            cpu_baseline = {"value": None, "error": repr(e)[:300]}

    if rank == 0:
        step_tflop = n_samples * tflop_per_sample
        line = {
            "metric": {"sdxl": "denoising steps/sec (SDXL 1024x1024 + depth ControlNet + Ctrl-Adapter, batch 8)",
                       "i2vgen": "denoising steps/sec (I2VGen-XL 16f 512x512 + depth ControlNet + Ctrl-Adapter, batch 4)",
                       "svd": "denoising steps/sec (SVD 14f 576x1024 + depth ControlNet + Ctrl-Adapter, batch 2 per GPU)"}[
                           a.workload],
            "value": value, "unit": "steps/s", "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
            "data": "synthetic",
            "config": {"workload": wl_name, "baseline_config": base_cfg, "parallelism": f"batch-sharded dp{world}",
                       "cuda_graph": use_graph, "l2": "per-step working set (6.3 GB bf16 weights + multi-GB activations) "
                                                      "exceeds the 126 MB L2; no explicit flush",
                       "algorithmic_tflop_per_step": round(step_tflop, 1),
                       "step_tflops_achieved": round(step_tflop / (ms_step / 1000.0), 1),
                       "step_frac_of_sustained_peak": round(step_tflop / (ms_step / 1000.0) / peaks["bf16_tflops_sustained"], 4)},
            "finite_outputs": finite, "clocks": clocks, "e2e": e2e, "gpu_launches": launches_per_step * a.steps,
            "launches_per_step": launches_per_step, "roofline": roofline, "kernel_families": families,
            "cpu_baseline": cpu_baseline, "eager_gpu_baseline": eager_gpu,
        }
        emit(line)
    if dist:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
