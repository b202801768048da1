"""Compile the CUDA engine in-tree:  freesasa_b200/csrc/*.cu  ->  freesasa_b200/csrc/libfsb200.so

sm_100a only (`-gencode arch=compute_100a,code=sm_100a`), `-lineinfo` so ncu's source page maps to
the .cu files.  nvcc cross-compiles without a GPU; the static CUDA runtime is linked in (nvcc's
default), so the library has no load-time dependency besides libstdc++/libc and can be dlopen'ed
from C, from Python (ctypes) or next to PyTorch's own runtime.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
ROOT = os.path.dirname(HERE)
LIB = os.path.join(CSRC, "libfsb200.so")
SOURCES = ["api.cu", "cells.cu", "integrate.cu"]
HEADERS = [os.path.join(CSRC, "engine.cuh"), os.path.join(CSRC, "cert_dirs.inc"), os.path.join(ROOT, "include", "fsb200.h")]
HOST_LIB = os.path.join(CSRC, "libfreesasa_b200_host.so")
HOST_SOURCES = ["host_shim.c", "radii.c", "ingest.c", "areas.c", "select.c", "workers.c"]
HOST_HEADERS = [os.path.join(ROOT, "include", "freesasa_b200_host.h"), os.path.join(ROOT, "include", "fsb200.h"),
                os.path.join(CSRC, "host_internal.h"), os.path.join(CSRC, "radius_tables.inc")]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def _stale(target: str, deps) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build_library(force: bool = False, verbose: bool = False) -> str:
    srcs = [os.path.join(CSRC, s) for s in SOURCES]
    if force or _stale(LIB, srcs + HEADERS + [os.path.abspath(__file__)]):
        cmd = [
            _nvcc(), "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
            "-Xcompiler", "-fPIC,-O2,-Wall", "-shared", "-o", LIB, *srcs,
        ]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
            print(" ".join(cmd), flush=True)
        # the image exports CC/CXX pointing at a wrapper toolchain; nvcc should use the distro g++
        env = dict(os.environ)
        if os.path.exists("/usr/bin/g++"):
            cmd[1:1] = ["-ccbin", "/usr/bin/g++"]
        subprocess.run(cmd, check=True, env=env)
    return LIB


def build_variant(name: str, defines, force: bool = False, replace=None) -> str:
    """Experiment build: the same sources with extra -D knobs (engine.cuh) -> csrc/libfsb200_<name>.so, used by
    tests/tools/ab_variants.py to measure one change at a time on the GPU box (FSB200_ENGINE_LIB selects it).
    `replace` maps a source name to another file (e.g. an older revision of integrate.cu taken from git)."""
    srcs = [(replace or {}).get(s, os.path.join(CSRC, s)) for s in SOURCES]
    out = os.path.join(CSRC, f"libfsb200_{name}.so")
    if force or _stale(out, srcs + HEADERS + [os.path.abspath(__file__)]):
        cmd = [_nvcc(), "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
               "-Xcompiler", "-fPIC,-O2,-Wall", "-shared", "-I", CSRC, "-o", out, *[f"-D{d}" for d in defines], *srcs]
        if os.path.exists("/usr/bin/g++"):
            cmd[1:1] = ["-ccbin", "/usr/bin/g++"]
        subprocess.run(cmd, check=True)
    return out


def build_host_shim(force: bool = False) -> str:
    """The C host layer that mirrors the reference's own entry points (freesasa_calc_coord ...)."""
    srcs = [os.path.join(CSRC, s) for s in HOST_SOURCES]
    if force or _stale(HOST_LIB, srcs + HOST_HEADERS + [LIB]):
        cc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc"
        cmd = [cc, "-std=gnu99", "-O2", "-fPIC", "-Wall", "-Wextra", "-shared", "-I", os.path.join(ROOT, "include"),
               "-o", HOST_LIB, *srcs, "-L", CSRC, "-lfsb200", "-Wl,-rpath,$ORIGIN", "-Wl,-Bsymbolic", "-lm", "-lpthread"]
        subprocess.run(cmd, check=True)
    return HOST_LIB


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv))
    print(build_host_shim(force="--force" in sys.argv))
