"""Developer tool: where do the roles of gemm_conv_kernel wait?  Builds a -DCA_TRACE copy of the library (on the CPU
box: `python scripts/gemm_trace.py --build`), then on a B200 runs BASELINE config-2 GEMM shapes once each and prints,
per case, the mean over CTAs of the cycles each role spent blocked (producer: empty smem stage, MMA issuer: full stage /
free accumulator, epilogue warps 0 and 7: accumulator ready, cp.async residual, tcgen05.wait::ld, arrive)."""
import math
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "ctrl_adapter_b200")
TRACE_LIB = os.path.join(PKG, "libctrl_adapter_b200_trace.so")


def build():
    bdir = os.path.join(PKG, "_build_trace")
    os.makedirs(bdir, exist_ok=True)
    objs = []
    procs = []
    sys.path.insert(0, ROOT)
    from ctrl_adapter_b200.build import SOURCES
    for src in SOURCES:
        o = os.path.join(bdir, src.replace(".cu", ".o"))
        objs.append(o)
        procs.append(subprocess.Popen(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
                                       "-DCA_TRACE", "-Xcompiler", "-fPIC", "-c", os.path.join(PKG, "csrc", src), "-o", o]))
    assert all(p.wait() == 0 for p in procs)
    subprocess.check_call(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-cudart", "static", "-o",
                           TRACE_LIB] + objs)
    print(TRACE_LIB)


def main():
    import torch
    n_slots = 16 * 4096 + 2 * 288
    trace = torch.zeros(n_slots, dtype=torch.int64, device="cuda")
    os.environ["CA_B200_LIB"] = TRACE_LIB
    os.environ["CA_GEMM_TRACE_PTR"] = hex(trace.data_ptr())
    sys.path.insert(0, ROOT)
    from ctrl_adapter_b200 import ops

    bf = torch.bfloat16

    def rnd(*s, scale=1.0):
        return (torch.randn(*s, device="cuda") * scale).to(bf)

    def lin(m, k, n, res=False, act=0, bn=0):
        x, w, b = rnd(m, k), rnd(n, k, scale=1 / math.sqrt(k)), torch.zeros(n, device="cuda")
        r = rnd(m, n) if res else None
        if act == ops.ACT_GEGLU:
            w, b = ops.pack_geglu_weight(w, b, 256)
        return lambda: ops.linear(x, w, b, residual=r, act=act, bn=bn)

    def conv(n, hw, c, co):
        x = rnd(n, hw, hw, c)
        w = ops.pack_conv_weight(rnd(co, c, 3, 3, scale=1 / math.sqrt(9 * c)))
        b = torch.zeros(co, device="cuda")
        return lambda: ops.conv2d(x, w, b)

    cases = {
        "lin 16384x1280->1280": (lin(16384, 1280, 1280), 2.0 * 16384 * 1280 * 1280),
        "lin 16384x1280->1280 +res": (lin(16384, 1280, 1280, res=True), 2.0 * 16384 * 1280 * 1280),
        "lin 65536x640->640 +res": (lin(65536, 640, 640, res=True), 2.0 * 65536 * 640 * 640),
        "lin 65536x640->640": (lin(65536, 640, 640), 2.0 * 65536 * 640 * 640),
        "lin 16384x1280->3840": (lin(16384, 1280, 3840), 2.0 * 16384 * 1280 * 3840),
        "lin 16384x5120->1280 +res": (lin(16384, 5120, 1280, res=True), 2.0 * 16384 * 5120 * 1280),
        "geglu 16384x1280->10240": (lin(16384, 1280, 10240, act=ops.ACT_GEGLU), 2.0 * 16384 * 1280 * 10240),
        "geglu 65536x640->5120": (lin(65536, 640, 5120, act=ops.ACT_GEGLU), 2.0 * 65536 * 640 * 5120),
        "conv3 16x128x128 320->320": (conv(16, 128, 320, 320), 2.0 * 16 * 128 * 128 * 2880 * 320),
        # thin (small K, small N) GEMMs of the video backbones: HBM-bound by their operands, far from it in round 2
        "thin 524288x320->320 +res": (lin(524288, 320, 320, res=True), 2.0 * 524288 * 320 * 320),
        "thin 524288x320->320": (lin(524288, 320, 320), 2.0 * 524288 * 320 * 320),
        "thin 524288x320->960": (lin(524288, 320, 960), 2.0 * 524288 * 320 * 960),
        "thin 131072x640->640 +res": (lin(131072, 640, 640, res=True), 2.0 * 131072 * 640 * 640),
        "thin 524288x512->320 +res": (lin(524288, 512, 320, res=True), 2.0 * 524288 * 512 * 320),
        "mid 524288x1280->320 +res": (lin(524288, 1280, 320, res=True), 2.0 * 524288 * 1280 * 320),
        "mid 131072x2560->640 +res": (lin(131072, 2560, 640, res=True), 2.0 * 131072 * 2560 * 640),
        "mid 65536x640->1920": (lin(65536, 640, 1920), 2.0 * 65536 * 640 * 1920),
        "mid 131072x1024->640": (lin(131072, 1024, 640), 2.0 * 131072 * 1024 * 640),
        "mid 65536x1920->640": (lin(65536, 1920, 640), 2.0 * 65536 * 1920 * 640),
        "mid conv3 16x64x64 320->320": (conv(16, 64, 320, 320), 2.0 * 16 * 64 * 64 * 2880 * 320),
        "mid conv3 128x64x64 320->320": (conv(128, 64, 320, 320), 2.0 * 128 * 64 * 64 * 2880 * 320),
    }
    def attn(b, h, l, lk=None):
        qkv = rnd(b, l, 3 * h * 64)
        q, k, v = qkv[:, :, : h * 64], qkv[:, :, h * 64: 2 * h * 64], qkv[:, :, 2 * h * 64:]
        if lk is not None:
            kv = rnd(b, lk, 2 * h * 64)
            k, v = kv[:, :, : h * 64], kv[:, :, h * 64:]
        return lambda: ops.attention(q, k, v, h, 64, 0.125)

    attn_cases = {
        "attn b16 h5 16384": (attn(16, 5, 16384), 4.0 * 16 * 5 * 16384 * 16384 * 64),
        "attn b16 h10 4096": (attn(16, 10, 4096), 4.0 * 16 * 10 * 4096 * 4096 * 64),
        "attn b16 h20 1024": (attn(16, 20, 1024), 4.0 * 16 * 20 * 1024 * 1024 * 64),
        "attn b16 h20 1024x77": (attn(16, 20, 1024, 77), 4.0 * 16 * 20 * 1024 * 77 * 64),
    }
    attn_names = ["sm_total", "sm_wait_s_full", "sm_wait_o_full", "sm_rescales", "sm_tmem_ld_wait", "mma_wait_kv_full",
                  "mma_wait_s_empty", "mma_wait_p_full", "mma_total", "prod_wait_kv_empty", "mma_wait_q_full",
                  "sm_wait_o_full_epilogue"]
    gemm_names = ["prod_total", "prod_wait_empty", "mma_wait_full", "mma_wait_acc", "e0_wait_accfull", "e0_total", "e0_wait_cp",
             "e0_arrive", "mma_total", "e0_tiles", "e7_wait_accfull", "e7_total", "e7_wait_cp", "e7_arrive", "e7_tiles",
             "e0_tmem_wait"]
    only_attn = "--attn" in sys.argv
    for name, (fn, flops) in ([] if only_attn else list(cases.items())) + list(attn_cases.items()):
        names = attn_names if name in attn_cases else gemm_names
        if "--only" in sys.argv and sys.argv[sys.argv.index("--only") + 1] not in name:
            continue
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        trace.zero_()
        fn()
        torch.cuda.synchronize()
        if "--events" in sys.argv and name in attn_cases:
            ev = trace[16 * 4096:].view(-1, 2).cpu().tolist()
            rows = sorted([(c, int(i), src) for src, lo in (("prod", 0), ("mma", 96), ("smx", 192))
                           for i, c in ev[lo:lo + 96] if c > 0])
            t0 = rows[0][0] if rows else 0
            print("   events (clk since first, src, id): " + " ".join(f"{c - t0}:{src}{i}" for c, i, src in rows[:120]))
        t = trace[:16 * 4096].view(-1, 16).cpu().double()
        used = t[t[:, 0] > 0][:2048]
        lead = used[used[:, 8] > 0]
        print(f"== {name}: {ms * 1e3:.1f} us  {flops / ms / 1e9:.0f} TFLOP/s  ctas={used.shape[0]}")
        if used.shape[0] == 0:
            continue
        mean = used.mean(0)
        mean_lead = lead.mean(0) if lead.shape[0] else mean
        parts = []
        for i, nm in enumerate(names):
            v = mean_lead[i] if nm.startswith("mma") else mean[i]
            parts.append(f"{nm}={v:.0f}")
        print("   " + "  ".join(parts))


if __name__ == "__main__":
    if "--build" in sys.argv:
        build()
    else:
        main()
