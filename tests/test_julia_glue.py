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


def _julia_tokens(src):
    """(token, bracket_depth) for identifiers / keywords of a Julia source, with comments, strings, chars and symbols removed;
    bracket_depth counts enclosing (), [] and {} (an `end` inside [] is an index, a `for` / `if` inside brackets a generator)"""
    out, i, n, depth = [], 0, len(src), 0
    while i < n:
        c = src[i]
        if c == "#":
            if src.startswith("#=", i):
                i = src.index("=#", i) + 2
            else:
                while i < n and src[i] != "\n":
                    i += 1
        elif src.startswith('"""', i):
            i = src.index('"""', i + 3) + 3
        elif c == '"':
            i += 1
            while src[i] != '"':
                if src[i] == "\\":
                    i += 1
                elif src.startswith("$(", i):          # interpolation: skip to its closing parenthesis (may hold strings)
                    d, i = 1, i + 2
                    while d:
                        if src[i] == '"':              # a nested string literal inside the interpolation
                            i += 1
                            while src[i] != '"':
                                i += 2 if src[i] == "\\" else 1
                        d += {"(": 1, ")": -1}.get(src[i], 0)
                        i += 1
                    continue
                i += 1
            i += 1
        elif c == "'" and i + 2 < n and (src[i + 2] == "'" or (src[i + 1] == "\\" and src[i + 3] == "'")):
            i += 3 if src[i + 2] == "'" else 4          # character literal (a lone ' is the adjoint operator)
        elif c in "([{":
            depth += 1; i += 1
        elif c in ")]}":
            depth -= 1; i += 1
            assert depth >= 0, "unbalanced closing bracket near line %d" % (src.count("\n", 0, i) + 1)
        elif c == ":" and i + 1 < n and (src[i + 1].isalpha() or src[i + 1] == "_") and (i == 0 or not (src[i - 1].isalnum() or src[i - 1] in "_)]}")):
            i += 1                                       # a Symbol literal such as :none, :end — not a keyword
            while i < n and (src[i].isalnum() or src[i] in "_!"):
                i += 1
        elif c.isalpha() or c == "_" or c == "@":
            j = i + 1
            while j < n and (src[j].isalnum() or src[j] in "_!"):
                j += 1
            prev_dot = i > 0 and src[i - 1] == "."       # field access such as x.end never happens, but x.begin could
            if not prev_dot:
                out.append((src[i:j], depth, src.count("\n", 0, i) + 1))
            i = j
        else:
            i += 1
    assert depth == 0, "unbalanced brackets at end of file"
    return out


def _check_julia_blocks(src, name):
    openers = {"module", "baremodule", "function", "struct", "if", "for", "while", "try", "let", "do", "begin", "quote", "macro"}
    block_exprs = {"begin", "function", "let", "quote", "try", "do"}      # may also open inside a call's parentheses
    stack = []
    for t, depth, line in _julia_tokens(src):
        if t in openers and (depth == 0 or t in block_exprs):
            stack.append((t, line, depth))
        elif t == "end":
            if stack and stack[-1][2] == depth:
                stack.pop()
            else:
                assert depth > 0, f"{name}:{line}: `end` without an opener"      # inside brackets: the `end` of an index, a[2:end]
    assert not stack, f"{name}: unclosed `{stack[-1][0]}` opened at line {stack[-1][1]}"


def test_julia_sources_have_balanced_blocks_and_brackets():
    """No Julia here to parse the glue: a tokenizer (comments, strings, interpolations, symbols stripped) checks that every (), [], {}
    closes and that the block openers (module, function, struct, if, for, while, try, let, do, begin, quote, macro; generators and
    ternaries inside brackets excluded) are matched one to one by `end` at the same bracket depth — the class of slip a hand edit
    of a never-executed file makes.  The checker is itself checked on broken copies."""
    import pytest
    for path in JULIA:
        src = open(path).read()
        _check_julia_blocks(src, os.path.basename(path))
    src = open(JULIA[0]).read()
    k = src.index("function check(rc::Cint)")
    cut = src.index("\nend\n", k)
    with pytest.raises(AssertionError, match="unclosed"):
        _check_julia_blocks(src[:cut] + src[cut + 4:], "one `end` removed")
    with pytest.raises(AssertionError, match="without an opener|unclosed"):
        _check_julia_blocks(src[:cut] + "\nend" + src[cut:], "one `end` too many")
    with pytest.raises(AssertionError, match="unbalanced"):
        _check_julia_blocks(src.replace("ccall((:pg_sync, LIB), Cint, (Ptr{Cvoid},), ctx().h))", "ccall((:pg_sync, LIB), Cint, (Ptr{Cvoid},), ctx().h)", 1), "a parenthesis removed")
