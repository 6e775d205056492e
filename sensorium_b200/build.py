"""Build libdwn_b200.so (hand-written sm_100a kernels + C ABI) in-tree with nvcc.

`python -m sensorium_b200.build` or `sensorium_b200.build.build()`; nvcc cross-compiles without a GPU.
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

CSRC = Path(__file__).resolve().parent / "csrc"
LIB = Path(__file__).resolve().parent / "libdwn_b200.so"
SOURCES = ["dwn_api.cu", "dwn_gemm.cu", "dwn_core_fwd.cu", "dwn_core_bwd.cu", "dwn_head.cu", "dwn_optim.cu",
           "dwn_pw_algebra.cu", "dwn_io.cu", "dwn_comm.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "--use_fast_math=false"]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def _stale(out: Path, deps) -> bool:
    if not out.exists():
        return True
    t = out.stat().st_mtime
    return any(Path(d).stat().st_mtime > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> Path:
    headers = list(CSRC.glob("*.cuh")) + [CSRC.parent.parent / "include" / "dwn_b200.h"]
    srcs = [CSRC / s for s in SOURCES if (CSRC / s).exists()]
    objdir = CSRC / "build"
    objdir.mkdir(exist_ok=True)
    flags = [f for f in NVCC_FLAGS if not f.startswith("--use_fast_math")]

    def compile_one(src: Path):
        obj = objdir / (src.stem + ".o")
        if force or _stale(obj, [src] + headers):
            cmd = [_nvcc(), *flags, "-c", str(src), "-o", str(obj)]
            if verbose:
                print(" ".join(cmd), flush=True)
            subprocess.run(cmd, check=True)
        return obj

    with ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        objs = list(ex.map(compile_one, srcs))
    if force or _stale(LIB, objs):
        cmd = [_nvcc(), "-shared", "-o", str(LIB), *map(str, objs), "-cudart", "static", "-Wno-deprecated-gpu-targets", "-ldl"]
        if verbose:
            print(" ".join(cmd), flush=True)
        subprocess.run(cmd, check=True)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
