"""In-tree build of libtempme_b200.so for sm_100a (explicit nvcc; no JIT cache)."""
from __future__ import annotations

import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
SOURCES = ["graph.cu", "sample.cu", "encoder.cu", "encoder_tc.cu", "tc_selftest.cu", "edge_imp.cu", "enhance.cu", "kl.cu"]
HEADERS = ["common.cuh", "tc.cuh", "timeenc.cuh", os.path.join("..", "..", "include", "tempme_b200.h")]
LIB = os.path.join(CSRC, "libtempme_b200.so")


def nvcc_path():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    return None


def stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(os.path.join(CSRC, f)) > t for f in SOURCES + HEADERS)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not stale():
        return LIB
    nvcc = nvcc_path()
    if nvcc is None:
        raise RuntimeError("nvcc not found: cannot build tempme_b200/csrc/libtempme_b200.so")
    cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
           "-Xcompiler", "-fPIC,-fopenmp", "-shared", "-o", LIB] + SOURCES
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    if os.environ.get("TEMPME_BUILD_TIMING"):          # diagnostic build: per-round clock stamps in the scorer (TEMPME_TC_TIMING=1 at run time)
        cmd.insert(1, "-DTM_TC_TIMING")
    subprocess.check_call(cmd, cwd=CSRC)
    return LIB


if __name__ == "__main__":
    print(build(force=True, verbose=True))
