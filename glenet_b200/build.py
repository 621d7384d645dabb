"""Compile ``libglenet_geom.so`` in-tree with plain nvcc for sm_100a (no torch headers).

    python -m glenet_b200.build [--force] [--verbose]

The shared object lands in ``glenet_b200/lib/`` (git-ignored; it travels to the GPU box
with the working tree).  There is exactly one target architecture and no fallback.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
LIBPATH = os.path.join(LIBDIR, "libglenet_geom.so")
INCLUDE = os.path.join(os.path.dirname(HERE), "include")
SOURCES = ["iou.cu", "iou3d_v1.cu", "nms.cu", "pib.cu", "vnms.cu", "rotate_iou.cu", "crop.cu", "host.cpp"]
HEADERS = ["common.cuh", "geom.cuh", "exchange.cuh", "clip.cuh", os.path.join(INCLUDE, "glenet_geom.h")]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "--fmad=true",            # same default as the reference build; rounding-critical code uses *_rn intrinsics
    "-Xcompiler", "-fPIC", "-Xcompiler", "-O2",
]


def nvcc() -> str:
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.isfile(exe):
        raise RuntimeError("nvcc not found: libglenet_geom.so cannot be built")
    return exe


def _stale() -> bool:
    if not os.path.isfile(LIBPATH):
        return True
    t = os.path.getmtime(LIBPATH)
    deps = [os.path.join(CSRC, s) for s in SOURCES] + [h if os.path.isabs(h) else os.path.join(CSRC, h) for h in HEADERS]
    return any(os.path.getmtime(d) > t for d in deps)


def build_library(force: bool = False, verbose: bool = False) -> str:
    if not force and not _stale():
        return LIBPATH
    os.makedirs(LIBDIR, exist_ok=True)
    objs = []
    procs = []
    for src in SOURCES:
        obj = os.path.join(LIBDIR, os.path.splitext(src)[0] + ".o")
        cmd = [nvcc(), *NVCC_FLAGS, "-I", INCLUDE, "-c", os.path.join(CSRC, src), "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas")
            cmd.insert(2, "-v")
            print(" ".join(cmd), flush=True)
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    for src, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode:
            print(out, flush=True)
        if p.returncode:
            raise RuntimeError(f"nvcc failed on {src}")
    cmd = [nvcc(), "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIBPATH, *objs]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode:
        print(r.stdout)
        raise RuntimeError("nvcc link failed")
    for o in objs:
        os.remove(o)
    return LIBPATH


TOOLS_CUDA = os.path.join(os.path.dirname(HERE), "tools", "cuda")
TOOL_BINARIES = ["ffma_peak"]       # measurement microkernels bench.py runs beside the library (FP32 peak of this GPU)


def build_tools(force: bool = False) -> list:
    """Compile the measurement microkernels of tools/cuda into tools/cuda/bin (git-ignored, travels to the GPU box)."""
    out = []
    bindir = os.path.join(TOOLS_CUDA, "bin")
    os.makedirs(bindir, exist_ok=True)
    for name in TOOL_BINARIES:
        src, exe = os.path.join(TOOLS_CUDA, name + ".cu"), os.path.join(bindir, name)
        if force or not os.path.isfile(exe) or os.path.getmtime(src) > os.path.getmtime(exe):
            r = subprocess.run([nvcc(), "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-o", exe, src],
                               stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
            if r.returncode:
                print(r.stdout)
                raise RuntimeError(f"nvcc failed on {src}")
        out.append(exe)
    return out


if __name__ == "__main__":
    build_tools(force="--force" in sys.argv)
    path = build_library(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print(path)
