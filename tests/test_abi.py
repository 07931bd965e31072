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
