"""Builds libmt_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m musicgeneration_b200.build [--force] [--verbose]
"""
from __future__ import annotations

import concurrent.futures as cf
import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
BUILD = os.path.join(HERE, "_build")
LIB = os.path.join(HERE, "libmt_b200.so")
SOURCES = ["elementwise.cu", "gemm_simt.cu", "rga_simt.cu", "decode.cu", "decode_step.cu", "gemm_tc.cu", "gemm_skinny.cu", "rga_tc.cu", "rga_tc_bwd.cu", "rga_tc_bwd2.cu", "rga_tc_bwd3.cu", "rga_tc_bwd4.cu", "rga_tc_bwd4q.cu",
           "feed.cu", "api.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-Xptxas", "-v", "--expt-relaxed-constexpr"] + os.environ.get("MT_NVCC_EXTRA", "").split()


def _nvcc() -> str:
    for c in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if c and os.path.isfile(c):
            return c
    raise RuntimeError("nvcc not found (needed to build libmt_b200.so)")


def _digest(path: str) -> str:
    h = hashlib.sha256()
    for dep in sorted(os.listdir(CSRC)) + ["../../include/mt_b200.h"]:
        if dep.endswith((".cuh", ".h")):
            with open(os.path.join(CSRC, dep), "rb") as f:
                h.update(f.read())
    with open(path, "rb") as f:
        h.update(f.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def _compile(src: str, force: bool, verbose: bool) -> str:
    os.makedirs(BUILD, exist_ok=True)
    path = os.path.join(CSRC, src)
    obj = os.path.join(BUILD, src.replace(".cu", ".o"))
    stamp = obj + ".sha"
    dig = _digest(path)
    if not force and os.path.isfile(obj) and os.path.isfile(stamp) and open(stamp).read() == dig:
        return obj
    cmd = [_nvcc()] + NVCC_FLAGS + ["-c", path, "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    log = r.stdout + r.stderr
    with open(obj + ".log", "w") as f:
        f.write(log)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed on {src}:\n{log[-4000:]}")
    if verbose:
        print(log)
    with open(stamp, "w") as f:
        f.write(dig)
    return obj


def build(force: bool = False, verbose: bool = False) -> str:
    with cf.ThreadPoolExecutor(max_workers=min(8, len(SOURCES))) as ex:
        objs = list(ex.map(lambda s: _compile(s, force, verbose), SOURCES))
    newest = max(os.path.getmtime(o) for o in objs)
    if force or not os.path.isfile(LIB) or os.path.getmtime(LIB) < newest:
        cmd = [_nvcc(), "-shared", "-o", LIB] + objs + ["-lcudart", "-lcuda"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n" + r.stdout + r.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
