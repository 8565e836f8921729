"""In-tree build of libtempme_b200.so for sm_100a (explicit nvcc; no JIT cache).

Staleness is decided by a content hash of the sources, headers and compiler flags written next to the library
(``libtempme_b200.so.stamp``), not by mtimes: a snapshot copy or a fresh checkout changes mtimes but not contents.
The build runs under an exclusive file lock, compiles to a temporary file in the same directory and renames it into
place, so concurrent ranks (torchrun, mp.spawn) never dlopen a half-written library: one of them builds, the others
wait on the lock and find the stamp current.
"""
from __future__ import annotations

import fcntl
import hashlib
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
SOURCES = ["graph.cu", "graph_build.cu", "sample.cu", "encoder.cu", "encoder_tc.cu", "tc_selftest.cu", "edge_imp.cu", "enhance.cu", "kl.cu", "train.cu", "gemm_tc.cu"]
HEADERS = ["common.cuh", "tc.cuh", "timeenc.cuh", "beta.cuh", os.path.join("..", "..", "include", "tempme_b200.h")]
LIB = os.path.join(CSRC, "libtempme_b200.so")
STAMP = LIB + ".stamp"
LOCK = os.path.join(CSRC, ".build.lock")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC,-fopenmp", "-shared"]


def nvcc_path():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    return None


def _extra_flags():
    # diagnostic build: per-round clock stamps in the scorer (TEMPME_TC_TIMING=1 at run time)
    # TEMPME_BUILD_DEFS: extra -D switches of A/B experiments (space separated)
    return (["-DTM_TC_TIMING"] if os.environ.get("TEMPME_BUILD_TIMING") else []) + os.environ.get("TEMPME_BUILD_DEFS", "").split()


def source_hash() -> str:
    h = hashlib.sha256()
    h.update(" ".join(FLAGS + _extra_flags()).encode())
    for f in SOURCES + HEADERS:
        h.update(f.encode())
        with open(os.path.join(CSRC, f), "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()


def _compile_objects(nvcc, verbose, force):
    """One object per translation unit, compiled concurrently; an object is reused when the hash of its source, the shared
    headers and the flags is unchanged (obj/<name>.o.stamp)."""
    from concurrent.futures import ThreadPoolExecutor
    objdir = os.path.join(CSRC, "obj")
    os.makedirs(objdir, exist_ok=True)
    cflags = [f for f in FLAGS if f != "-shared"] + _extra_flags()
    hh = hashlib.sha256(" ".join(cflags).encode())
    for f in HEADERS:
        with open(os.path.join(CSRC, f), "rb") as fh:
            hh.update(fh.read())

    def one(src):
        obj = os.path.join(objdir, src[:-3] + ".o")
        h = hh.copy()
        with open(os.path.join(CSRC, src), "rb") as fh:
            h.update(fh.read())
        digest = h.hexdigest()
        try:
            with open(obj + ".stamp") as fh:
                fresh = fh.read().strip() == digest and os.path.exists(obj)
        except OSError:
            fresh = False
        if force or not fresh:
            subprocess.check_call([nvcc] + (["-Xptxas=-v"] if verbose else []) + cflags + ["-c", "-o", obj, src], cwd=CSRC)
            with open(obj + ".stamp", "w") as fh:
                fh.write(digest + "\n")
        return obj

    with ThreadPoolExecutor(max_workers=min(len(SOURCES), os.cpu_count() or 1)) as ex:
        return list(ex.map(one, SOURCES))


def stale() -> bool:
    if not os.path.exists(LIB) or not os.path.exists(STAMP):
        return True
    try:
        with open(STAMP) as fh:
            return fh.read().strip() != source_hash()
    except OSError:
        return True


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not stale():
        return LIB
    nvcc = nvcc_path()
    if nvcc is None:
        raise RuntimeError("nvcc not found: cannot build tempme_b200/csrc/libtempme_b200.so")
    with open(LOCK, "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and not stale():          # another process built it while this one waited for the lock
                return LIB
            digest = source_hash()
            tmp = f"{LIB}.tmp.{os.getpid()}"
            try:
                objs = _compile_objects(nvcc, verbose, force)
                subprocess.check_call([nvcc] + FLAGS + ["-o", tmp] + objs, cwd=CSRC)
                os.replace(tmp, LIB)
            finally:
                if os.path.exists(tmp):
                    os.unlink(tmp)
            with open(STAMP + ".tmp", "w") as fh:
                fh.write(digest + "\n")
            os.replace(STAMP + ".tmp", STAMP)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)
    return LIB


if __name__ == "__main__":
    print(build(force=True, verbose=True))
