"""Build libprotograd_b200.so (sm_100a only) with nvcc, in-tree, next to this file.

    python protograd.jl_b200/build.py [--force]

One object per .cu (compiled in parallel), linked into a single shared library with a plain C ABI
(include/protograd_b200.h).  No libcuda link dependency: the TMA descriptor encoder is resolved at
run time through cudaGetDriverEntryPoint, so the library also loads (for symbol checks) on a box
without a GPU driver.
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libprotograd_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
         "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr", "-Xptxas", "-v"]


def _newer(src, dst, extra=()):
    if not os.path.exists(dst):
        return True
    t = os.path.getmtime(dst)
    return any(os.path.getmtime(p) > t for p in (src,) + tuple(extra))


def build_lib(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    srcs = sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))
    hdrs = tuple(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h")))
    hdrs += (os.path.join(os.path.dirname(HERE), "include", "protograd_b200.h"),)
    jobs = []
    for s in srcs:
        src, obj = os.path.join(CSRC, s), os.path.join(OBJ, s[:-3] + ".o")
        if force or _newer(src, obj, hdrs):
            jobs.append((src, obj))

    def cc(job):
        src, obj = job
        r = subprocess.run([NVCC] + FLAGS + ["-c", src, "-o", obj], capture_output=True, text=True)
        log = r.stdout + r.stderr
        with open(obj[:-2] + ".ptxas.log", "w") as f:
            f.write(log)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed for %s:\n%s" % (src, log))
        return log

    if jobs:
        with ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            for log in ex.map(cc, jobs):
                if verbose:
                    print(log)
    objs = [os.path.join(OBJ, s[:-3] + ".o") for s in srcs]
    if jobs or force or not os.path.exists(LIB):
        r = subprocess.run([NVCC, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB] + objs,
                           capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n" + r.stdout + r.stderr)
    return LIB


def build_trace_lib():
    """Debug variant (tools/tc_trace.py): tc_gemm.cu and mlp_step.cu compiled with -DPG_TC_TRACE (per-tile clock stamps of the pair
    kernel's roles, phase stamps of the one-launch MLP step), everything else from the regular objects -> libprotograd_b200_trace.so; loaded with PG_LIB_PATH=<that file>."""
    build_lib()
    traced = []
    for name in ("tc_gemm", "mlp_step"):
        obj = os.path.join(OBJ, name + "_trace.o")
        r = subprocess.run([NVCC] + FLAGS + ["-DPG_TC_TRACE", "-c", os.path.join(CSRC, name + ".cu"), "-o", obj], capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(r.stdout + r.stderr)
        traced.append(obj)
    skip = ("tc_gemm.o", "mlp_step.o", "tc_gemm_trace.o", "mlp_step_trace.o")
    objs = [os.path.join(OBJ, f) for f in sorted(os.listdir(OBJ)) if f.endswith(".o") and f not in skip] + traced
    out = os.path.join(HERE, "libprotograd_b200_trace.so")
    r = subprocess.run([NVCC, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", out] + objs, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(r.stdout + r.stderr)
    return out


if __name__ == "__main__":
    if "--trace" in sys.argv:
        print(build_trace_lib())
        sys.exit(0)
    print(build_lib(force="--force" in sys.argv, verbose="-v" in sys.argv))
