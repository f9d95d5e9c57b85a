"""In-tree build of the CUDA engine (sm_100a only) and of the oracle-side checkers.

``build_engine()`` compiles numpyro_b200/csrc/b200nuts.cu into libb200nuts.so next to the sources
so the library travels with the repository snapshot to the GPU box.
"""
from __future__ import annotations

import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(CSRC, "libb200nuts.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-fmad=false",            # det-f32 bookkeeping: no implicit FMA contraction; hot loops use explicit FMA
    "--extended-lambda", "-Xcompiler", "-fPIC", "-shared",
]


def _nvcc() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: the engine cannot be built (there is no CPU fallback)")


def _stale(target: str, sources) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def engine_sources():
    srcs = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith((".cu", ".cuh"))]
    srcs.append(os.path.join(os.path.dirname(HERE), "include", "b200nuts.h"))
    return srcs


STREAM_KS = (1, 2, 4, 7, 8)          # tile widths of the streaming kernel (stream_ks_for in stream_engine.cuh)


def _run(cmd):
    subprocess.check_call(cmd, cwd=CSRC)


def build_engine(force: bool = False, verbose: bool = False) -> str:
    """Compile the engine for sm_100a.  The streaming kernel's instances are split over one object per tile width and
    compiled in parallel; ``B200NUTS_FAST_KS=<ks>`` (development) builds a single translation unit with one instance."""
    srcs = engine_sources()
    if not (force or _stale(LIB, srcs)):
        return LIB
    nvcc = _nvcc()
    common = NVCC_FLAGS[:-1] + (["-Xptxas", "-v"] if verbose else [])          # (without -shared)
    if os.environ.get("B200NUTS_TICK_LAPS"):
        common = common + ["-DB2_TICK_LAPS"]
    fast = os.environ.get("B200NUTS_FAST_KS")
    if fast:
        _run([nvcc] + common + ["-shared", f"-DB2_STREAM_FAST_KS={int(fast)}", "-o", LIB, os.path.join(CSRC, "b200nuts.cu"),
                                os.path.join(CSRC, "gemm_engine.cu")])
        return LIB
    objdir = os.path.join(CSRC, "build")
    os.makedirs(objdir, exist_ok=True)
    jobs = [[nvcc] + common + ["-DB2_SPLIT_BUILD", "-c", "-o", os.path.join(objdir, "b200nuts.o"), os.path.join(CSRC, "b200nuts.cu")]]
    jobs.append([nvcc] + common + ["-c", "-o", os.path.join(objdir, "gemm_engine.o"), os.path.join(CSRC, "gemm_engine.cu")])
    for ks in STREAM_KS:
        jobs.append([nvcc] + common + [f"-DB2_INST_KS={ks}", "-c", "-o", os.path.join(objdir, f"stream_ks{ks}.o"),
                                       os.path.join(CSRC, "stream_instances.cu")])
    from concurrent.futures import ThreadPoolExecutor
    with ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 1)) as pool:
        list(pool.map(_run, jobs))
    objs = [os.path.join(objdir, "b200nuts.o"), os.path.join(objdir, "gemm_engine.o")] + \
           [os.path.join(objdir, f"stream_ks{ks}.o") for ks in STREAM_KS]
    _run([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", LIB] + objs)
    return LIB


if __name__ == "__main__":
    print(build_engine(force=True, verbose=True))
