"""Builds libsarpro_gpu.so (sm_100a) in-tree with nvcc. No JIT cache: the .so travels with the tree."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libsarpro_gpu.so")
SOURCES = ["api.cu", "api_f32.cu", "api_batch.cu", "api_jpeg.cu", "api_read.cu", "kernels_read.cu", "plan_read.cpp", "comm.cu", "kernels_dn.cu", "kernels_resize.cu", "kernels_hmma.cu", "kernels_plan.cu", "kernels_small.cu", "kernels_f32.cu", "plan.cpp", "plan_f32.cpp", "narrow.cpp"]
HEADERS = ["ctx.h", "kernels.h", "plan.h", "host_pool.h", "common.cuh", "clahe_exact.cuh", os.path.join("..", "..", "include", "sarpro_gpu.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "--fmad=false",            # Rust never contracts a*b+c; keep f32/f64 arithmetic IEEE op-by-op
    "-Xcompiler", "-fPIC,-O2,-ffp-contract=off,-fno-fast-math,-Wall",
    "-Xptxas", "-v" if os.environ.get("SARPRO_PTXAS_V") else "-O3",
] + os.environ.get("SARPRO_NVCC_EXTRA", "").split()


def nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def _stale(target: str, deps: list[str]) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False, variant: str = "", extra: list[str] | None = None) -> str:
    """variant/extra: an experimental build (extra nvcc flags) into libsarpro_gpu_<variant>.so, selected at run time with
    SARPRO_GPU_LIB=<path>; the default library is never touched by it."""
    hdrs = [os.path.join(CSRC, h) for h in HEADERS] + [os.path.abspath(__file__)]
    objdir = os.path.join(HERE, "build" + ("_" + variant if variant else ""))
    LIB = os.path.join(HERE, f"libsarpro_gpu_{variant}.so") if variant else globals()["LIB"]
    flags = NVCC_FLAGS + (extra or [])
    os.makedirs(objdir, exist_ok=True)
    objs = []
    procs = []
    for src in SOURCES:
        sp = os.path.join(CSRC, src)
        obj = os.path.join(objdir, src + ".o")
        objs.append(obj)
        if force or _stale(obj, [sp] + hdrs):
            cmd = [nvcc(), *flags, "-x", "cu", "-c", sp, "-o", obj]
            if verbose:
                print(" ".join(cmd), file=sys.stderr)
            procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0 or (verbose and out.strip()):
            print(f"--- {src}\n{out}", file=sys.stderr)
        failed |= p.returncode != 0
    if failed:
        raise RuntimeError("nvcc failed")
    if force or procs or _stale(LIB, objs):
        cmd = [nvcc(), "-shared", "-o", LIB, *objs, "-cudart", "static", "-ldl", "-lpthread", "-lrt",
               "-Xlinker", "--no-undefined",   # a missing definition must fail the build, not the first call
               "-gencode", "arch=compute_100a,code=sm_100a"]
        subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    var = [a.split("=", 1)[1] for a in sys.argv if a.startswith("--variant=")]
    ext = [a.split("=", 1)[1] for a in sys.argv if a.startswith("--extra=")]
    print(build(force="--force" in sys.argv, verbose=True, variant=var[0] if var else "", extra=ext))
