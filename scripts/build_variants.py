"""Developer tool: builds experimental variants of the library next to the default one (selected at run time with
CA_B200_LIB=<path>, see ctrl-adapter_b200/_lib.py).  Runs on the CPU box (nvcc cross-compiles).

  python scripts/build_variants.py nopdl    -> ctrl_adapter_b200/libctrl_adapter_b200_nopdl.so  (-DCA_NO_PDL: no programmatic dependent launch)
  python scripts/build_variants.py trace    -> ctrl_adapter_b200/libctrl_adapter_b200_trace.so  (-DCA_TRACE: role-wait counters)

A/B on the B200:   python bench.py --steps 10 ...   vs   CA_B200_LIB=$PWD/ctrl_adapter_b200/libctrl_adapter_b200_nopdl.so python bench.py ...
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "ctrl_adapter_b200")
sys.path.insert(0, ROOT)
VARIANTS = {"nopdl": ["-DCA_NO_PDL"], "trace": ["-DCA_TRACE"], "nopdl_trace": ["-DCA_NO_PDL", "-DCA_TRACE"]}


def build(name):
    from ctrl_adapter_b200.build import SOURCES
    bdir = os.path.join(PKG, f"_build_{name}")
    os.makedirs(bdir, exist_ok=True)
    objs, procs = [], []
    for src in SOURCES:
        o = os.path.join(bdir, src.replace(".cu", ".o"))
        objs.append(o)
        procs.append(subprocess.Popen(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
                                       "-Xcompiler", "-fPIC"] + VARIANTS[name] +
                                      ["-c", os.path.join(PKG, "csrc", src), "-o", o]))
    if not all(p.wait() == 0 for p in procs):
        raise SystemExit(f"nvcc failed for variant {name}")
    lib = os.path.join(PKG, f"libctrl_adapter_b200_{name}.so")
    subprocess.check_call(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-cudart", "static", "-o", lib] + objs)
    print(lib)


if __name__ == "__main__":
    for v in sys.argv[1:] or ["trace"]:
        build(v)
