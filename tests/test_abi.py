"""CPU-only checks of the drop-in boundary: the C-ABI library loads, exports every symbol that
include/protograd_b200.h declares, the ctypes binding covers all of them, and the product path
fails loudly (no CPU fallback) when there is no GPU."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "protograd_b200.h")
LIB = os.path.join(ROOT, "protograd.jl_b200", "libprotograd_b200.so")


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(pg_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_symbols():
    syms = declared_symbols()
    assert len(syms) >= 50 and "pg_init" in syms and "pg_tc_linear_fwd" in syms


def test_library_exports_every_declared_symbol():
    assert os.path.exists(LIB), "run __graft_entry__.build() first"
    lib = ctypes.CDLL(LIB)
    missing = [s for s in declared_symbols() if not hasattr(lib, s)]
    assert not missing, missing


def test_binding_covers_header():
    import protograd_b200.capi as capi
    assert set(declared_symbols()) == set(capi.EXPORTS)


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import protograd_b200 as pg
    with pytest.raises(pg.capi.PGError, match="cuda|CUDA|device"):
        pg.Context.get(0)


def test_product_package_does_not_import_oracle():
    pkg = os.path.join(ROOT, "protograd.jl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in txt.replace("the oracle", "").replace("CPU oracle", "") or f in ("__init__.py",) and "import oracle" not in txt, f


def test_tensor_core_conv_envelope_is_a_host_query():
    """pg_tc_conv2d_supported is pure host logic (shape envelope + shared-memory group choice of
    csrc/tc_conv.cu): callable without a GPU; the plan uses it to pick the conv kernel per layer."""
    lib = ctypes.CDLL(LIB)
    sup = lib.pg_tc_conv2d_supported
    sup.restype = ctypes.c_int
    sup.argtypes = [ctypes.c_int] * 6
    # LeNet convs of BASELINE configs[2] (Cin, H, W, Cout, kH, kW)
    assert sup(1, 28, 28, 6, 5, 5) == 1 and sup(6, 12, 12, 16, 5, 5) == 1
    assert sup(16, 6, 6, 16, 3, 3) == 1 and sup(3, 10, 10, 4, 1, 1) == 1
    assert sup(17, 8, 8, 4, 3, 3) == 0          # more than 16 input channels
    assert sup(4, 8, 8, 17, 3, 3) == 0          # more than 16 output channels
    assert sup(16, 12, 12, 16, 5, 5) == 0       # im2col width 16*25 + 1 > 256 (weight-gradient UMMA N)
    assert sup(3, 224, 224, 8, 3, 3) == 0       # one sample does not fit a shared-memory slot
    assert sup(1, 4, 4, 1, 5, 5) == 0           # filter larger than the image
