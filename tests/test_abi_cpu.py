"""CPU tests of the drop-in boundary: the C-ABI library builds, loads without a CUDA driver, exports every symbol
that include/ctrl_adapter_b200.h declares, and fails loudly (no silent fallback) when no GPU is present."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from ctrl_adapter_b200 import _lib, build
    build.build()
    return _lib.load()


def test_header_symbols_exported(lib):
    hdr = open(os.path.join(ROOT, "include", "ctrl_adapter_b200.h")).read()
    declared = set(re.findall(r"^(?:int|const char\*)\s+(ca_\w+)\s*\(", hdr, flags=re.M))
    assert len(declared) >= 20
    from ctrl_adapter_b200 import _lib
    raw = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(raw, name), f"{name} declared in the header but not exported"
    assert declared == set(_lib.SIGNATURES), "ctypes SIGNATURES and header declarations differ"


def test_abi_version_and_struct_sizes(lib):
    from ctrl_adapter_b200 import _lib
    assert lib.ca_abi_version() == 1
    # the ctypes mirror must have the C layout: compile-time sizes are checked against gcc's view of the header
    import subprocess, tempfile
    src = '#include <stdio.h>\n#include "ctrl_adapter_b200.h"\nint main(){printf("%zu %zu", sizeof(ca_gemm_desc), sizeof(ca_attention_desc));return 0;}'
    with tempfile.TemporaryDirectory() as d:
        c = os.path.join(d, "s.c")
        open(c, "w").write(src)
        exe = os.path.join(d, "s")
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), c, "-o", exe])
        out = subprocess.check_output([exe]).decode().split()
    assert int(out[0]) == ctypes.sizeof(_lib.GemmDesc)
    assert int(out[1]) == ctypes.sizeof(_lib.AttentionDesc)


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_fails_loudly_without_gpu(lib):
    from ctrl_adapter_b200 import _lib, ops
    assert lib.ca_device_ok() != 0
    assert b"CUDA" in lib.ca_last_error() or b"device" in lib.ca_last_error()
    with pytest.raises(ValueError, match="no CPU path"):
        ops.linear(torch.zeros(4, 8, dtype=torch.bfloat16), torch.zeros(8, 8, dtype=torch.bfloat16))
    d = _lib.GemmDesc()
    assert lib.ca_gemm(ctypes.byref(d), None) == 1  # CA_ERR_INVALID, with a message
    assert len(lib.ca_last_error()) > 0


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "ctrl_adapter_b200")
    for f in os.listdir(pkg):
        if f.endswith(".py"):
            src = open(os.path.join(pkg, f)).read()
            assert "oracle" not in src.replace("# oracle", ""), f"{f} references the oracle"
