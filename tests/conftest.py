import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run under gpurun)")
    # The suite exercises the built library (symbol / layout / planner checks on CPU, everything on GPU).  A fresh clone has
    # no libprotograd_b200.so (built artefacts are git-ignored): compile it the way __graft_entry__.build() does rather than
    # failing every test at import.  nvcc cross-compiles sm_100a without a GPU; without nvcc the import error stands.
    lib = os.path.join(ROOT, "protograd.jl_b200", "libprotograd_b200.so")
    if not os.path.exists(lib):
        import importlib.util
        spec = importlib.util.spec_from_file_location("pg_build", os.path.join(ROOT, "protograd.jl_b200", "build.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        if os.path.exists(mod.NVCC):
            mod.build_lib()


def pytest_collection_modifyitems(config, items):
    # GPU tests must never silently pass without a device: skip only when torch sees no GPU
    # AND the user did not ask for them explicitly with -m gpu.
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    mexpr = config.getoption("-m") or ""
    if "gpu" in mexpr and "not gpu" not in mexpr:
        return  # explicit request: let them fail loudly
    skip = pytest.mark.skip(reason="no GPU in this container")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)
