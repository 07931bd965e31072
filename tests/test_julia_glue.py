"""CPU-only: the Julia glue (julia/*.jl) cannot be executed here (no Julia toolchain in the image), so what can be
checked statically is: every `ccall((:pg_xxx, LIB), Cint, (argtypes...), ...)` names a symbol that
include/protograd_b200.h declares and the library exports, with the SAME NUMBER of arguments, and the Julia `PgOp`
struct has the header's `pg_op` layout."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "protograd_b200.h")
LIB = os.path.join(ROOT, "protograd.jl_b200", "libprotograd_b200.so")
JULIA = [os.path.join(ROOT, "julia", f) for f in ("ProtoGradB200.jl", "parity.jl", "bench.jl")]


def header_arity():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    out = {}
    for m in re.finditer(r"\b(?:int|void|const char\*)\s+(pg_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", src, flags=re.S):
        args = m.group(2).strip()
        out[m.group(1)] = 0 if args in ("", "void") else len(_split_top(args))
    return out


def _split_top(s):
    parts, depth, cur = [], 0, ""
    for ch in s:
        if ch in "([{":
            depth += 1
        elif ch in ")]}":
            depth -= 1
        if ch == "," and depth == 0:
            parts.append(cur); cur = ""
        else:
            cur += ch
    if cur.strip():
        parts.append(cur)
    return parts


def julia_ccalls(path):
    src = open(path).read()
    calls = []
    for m in re.finditer(r"ccall\(\(:(pg_[a-z0-9_]+),\s*LIB\),\s*(\w+),\s*\(", src):
        i = m.end()            # just after the '(' of the argument-type tuple
        depth, j = 1, i
        while depth:
            depth += {"(": 1, ")": -1}.get(src[j], 0)
            j += 1
        types = src[i:j - 1].strip()
        n = 0 if types in ("", ",") else len([t for t in _split_top(types) if t.strip()])
        calls.append((m.group(1), n, src.count("\n", 0, m.start()) + 1))
    return calls


def test_every_julia_ccall_matches_the_header():
    arity = header_arity()
    assert len(arity) >= 90 and arity["pg_plan_forward"] == 8 and arity["pg_last_error"] == 0
    lib = ctypes.CDLL(LIB)
    seen = set()
    for path in JULIA:
        assert os.path.exists(path), path
        for name, n, line in julia_ccalls(path):
            assert name in arity, f"{os.path.basename(path)}:{line}: {name} is not declared in the header"
            assert hasattr(lib, name), f"{name} is not exported by the library"
            assert n == arity[name], f"{os.path.basename(path)}:{line}: {name} called with {n} argument types, the header declares {arity[name]}"
            seen.add(name)
    # the glue really goes through the plan / arena / dp groups of SURVEY.md section 8(b)
    for must in ("pg_init", "pg_arena_create", "pg_arena_like", "pg_arena_upload", "pg_arena_download", "pg_vec_expr", "pg_vec_dot", "pg_plan_create",
                 "pg_plan_forward", "pg_plan_backward", "pg_plan_value", "pg_loss_read", "pg_opt_gd", "pg_opt_adam", "pg_opt_nesterov", "pg_opt_heavyball",
                 "pg_dp_init", "pg_dp_allreduce", "pg_dp_bucket_begin"):
        assert must in seen, must


def test_julia_pg_op_struct_matches_the_header_layout():
    from protograd_b200.trace import PgOp
    assert ctypes.sizeof(PgOp) == 80
    src = open(JULIA[0]).read()
    body = src[src.index("struct PgOp"):src.index("end", src.index("struct PgOp"))]
    fields = re.findall(r"(\w+)::(Int32|Int64|Float64)", body)
    size = {"Int32": 4, "Int64": 8, "Float64": 8}
    off = 0
    for _, t in fields:
        off = (off + size[t] - 1) // size[t] * size[t]
        off += size[t]
    assert off == 80 and [t for _, t in fields] == ["Int32"] * 6 + ["Int64"] * 4 + ["Int32"] * 2 + ["Float64"] * 2


def test_julia_mlp_desc_struct_matches_the_header_layout():
    """PgMlpDesc (julia/ProtoGradB200.jl) has the field order, types and size of pg_mlp_desc (= capi.MlpDesc)."""
    from protograd_b200.capi import MlpDesc
    src = open(JULIA[0]).read()
    body = src[src.index("struct PgMlpDesc"):src.index("\nend", src.index("struct PgMlpDesc"))]
    fields = re.findall(r"^\s+(\w+)::(NTuple\{(\d),(Int64|Ptr\{Cvoid\})\}|Int32|Int64|Float64|Ptr\{Cvoid\})", body, re.M)
    size = {"Int32": 4, "Int64": 8, "Float64": 8, "Ptr{Cvoid}": 8}
    off, names = 0, []
    for name, t, cnt, elt in fields:
        n, e = (int(cnt), elt) if cnt else (1, t)
        off = (off + size[e] - 1) // size[e] * size[e]
        assert getattr(MlpDesc, name).offset == off, (name, off, getattr(MlpDesc, name).offset)
        off += n * size[e]
        names.append(name)
    assert names == [f[0] for f in MlpDesc._fields_]
    assert (off + 7) // 8 * 8 == ctypes.sizeof(MlpDesc)
