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


def test_struct_layouts_match_the_header(tmp_path):
    """pg_op and pg_mlp_desc as gcc lays them out from include/protograd_b200.h == the ctypes mirrors the host side fills."""
    import subprocess
    import ctypes
    from protograd_b200.capi import MlpDesc
    from protograd_b200.trace import PgOp
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cname = {"inp": "in"}                                 # `in` is a Python keyword: the ctypes mirror calls it `inp`
    fields = {"pg_op": [f[0] for f in PgOp._fields_], "pg_mlp_desc": [f[0] for f in MlpDesc._fields_]}
    src = ['#include <stdio.h>', '#include <stddef.h>', '#include "protograd_b200.h"', 'int main(void) {']
    for st, names in fields.items():
        src.append(f'  printf("{st} %zu\\n", sizeof({st}));')
        for n in names:
            src.append(f'  printf("{st}.{n} %zu\\n", offsetof({st}, {cname.get(n, n)}));')
    src += ['  return 0;', '}']
    c = tmp_path / "layout.c"
    c.write_text("\n".join(src))
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-I", os.path.join(root, "include"), str(c), "-o", str(exe)], check=True)
    out = dict(line.split() for line in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.splitlines())
    for st, cls in (("pg_op", PgOp), ("pg_mlp_desc", MlpDesc)):
        assert int(out[st]) == ctypes.sizeof(cls), st
        for n in fields[st]:
            assert int(out[f"{st}.{n}"]) == getattr(cls, n).offset, (st, n)


def _build_c_host(tmp_path):
    """tools/abi_smoke.c: a plain-C caller of the ABI, compiled with -Wall -Werror against the public header"""
    import subprocess
    exe = tmp_path / "abi_smoke"
    subprocess.run(["gcc", "-O2", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tools", "abi_smoke.c"),
                    "-L", os.path.dirname(LIB), "-lprotograd_b200", "-lm", "-o", str(exe)], check=True)
    env = dict(os.environ, LD_LIBRARY_PATH=os.path.dirname(LIB) + os.pathsep + os.environ.get("LD_LIBRARY_PATH", ""))
    return subprocess.run([str(exe)], capture_output=True, text=True, env=env)


def test_c_host_links_and_fails_loudly_without_gpu(tmp_path):
    """the header is valid C (not only C++), every symbol the C host uses links, and without a GPU the program stops at its
    first call with the library's error message — no CPU fallback below the ABI either"""
    import torch
    r = _build_c_host(tmp_path)
    if torch.cuda.is_available():
        assert r.returncode == 0 and "abi_smoke ok" in r.stdout, r.stdout + r.stderr
    else:
        assert r.returncode != 0 and "abi_smoke ok" not in r.stdout
        assert "pg_device_count" in r.stderr or "pg_init" in r.stderr, r.stderr


@pytest.mark.gpu
def test_c_host_trains_c1_through_the_abi(tmp_path):
    """one C1 step (784-100-10, minibatch 128) driven from C: loss and the logits-layer bias gradient vs the program's own
    double-precision evaluation of src/layers.jl:28-34,81,87 + src/losses.jl:7-10 (rel <= 1e-5), a stale pullback rejected,
    dot over the arena, then 30 Adam steps that must reduce the loss"""
    r = _build_c_host(tmp_path)
    assert r.returncode == 0 and "abi_smoke ok" in r.stdout, r.stdout + r.stderr


def test_no_floating_point_atomics_and_blackwell_instructions_in_the_sass():
    """Determinism and "B200-native" as properties of the shipped binary (cuobjdump works without a GPU): the library holds no
    floating-point ATOM / ATOMG / ATOMS / RED instruction — every gradient sum is a fixed-order reduction of partials, the only
    atomics are integer tickets, tile counters and timing — and its GEMM / conv kernels are tcgen05 + TMA code (UTCHMMA incl. the
    cta_group::2 form, LDTM, UTMALDG / UTMASTG, UBLKCP), not mma.sync recompiles (no HMMA)."""
    import shutil
    import subprocess
    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump not on PATH")
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    ops = re.findall(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\w+\s+)?([A-Za-z0-9_.]+)", sass, flags=re.M)
    atomics = {o for o in ops if o.split(".")[0] in ("ATOM", "ATOMG", "ATOMS", "RED")}
    assert atomics, "no atomics at all: the mnemonic filter is broken"
    fp = {o for o in atomics if re.search(r"\.(F16|F32|F64|BF16)", o)}
    assert not fp, fp
    have = {o.split(".")[0] for o in ops}
    for must in ("UTCHMMA", "LDTM", "UTMALDG", "UTMASTG", "UBLKCP", "UTCBAR"):
        assert must in have, must
    assert any(o.startswith("UTCHMMA.2CTA") for o in ops) and any(o.startswith("UTMALDG.2D.2CTA") for o in ops)
    assert "HMMA" not in have
