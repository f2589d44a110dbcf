"""Times ops.linear for several (M, K, N, bn, residual) combinations next to torch.matmul (cuBLAS) at the same shape."""
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ctrl_adapter_b200 import ops  # noqa: E402

BF16 = torch.bfloat16
dev = "cuda"


def timeit(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


cases = [(16384, 1280, 1280), (16384, 5120, 1280), (16384, 1280, 3840), (65536, 640, 640), (65536, 640, 1920),
         (262144, 320, 512), (262144, 512, 320), (16384, 1280, 10240), (8192, 8192, 8192)]
for (m, k, n) in cases:
    x = (torch.randn(m, k, device=dev)).to(BF16)
    w = (torch.randn(n, k, device=dev) / math.sqrt(k)).to(BF16)
    res = torch.randn(m, n, device=dev).to(BF16)
    fl = 2.0 * m * k * n
    t_cublas = timeit(lambda: torch.matmul(x, w.t()))
    line = f"m{m} k{k} n{n}: cublas {t_cublas*1e3:7.1f}us {fl/t_cublas/1e9:7.1f}TF |"
    for bn in (128, 160, 256):
        if bn == 160 and n % 160:
            continue
        t = timeit(lambda: ops.linear(x, w, None, bn=bn))
        line += f" bn{bn} {t*1e3:7.1f}us {fl/t/1e9:6.1f}TF |"
    t = timeit(lambda: ops.linear(x, w, None, residual=res))
    line += f" auto+res {t*1e3:7.1f}us {fl/t/1e9:6.1f}TF"
    print(line, flush=True)
    del x, w, res
