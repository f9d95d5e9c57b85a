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


def build_engine(force: bool = False, verbose: bool = False) -> str:
    srcs = engine_sources()
    if force or _stale(LIB, srcs):
        fast = os.environ.get("B200NUTS_FAST_KS")          # development only: a single streaming-kernel instance
        cmd = ([_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ([f"-DB2_STREAM_FAST_KS={int(fast)}"] if fast else []) + (["-DB2_TICK_LAPS"] if os.environ.get("B200NUTS_TICK_LAPS") else [])
               + ["-o", LIB, os.path.join(CSRC, "b200nuts.cu")])
        subprocess.check_call(cmd, cwd=CSRC)
    return LIB


if __name__ == "__main__":
    print(build_engine(force=True, verbose=True))
