"""Build libstorm_b200.so in-tree with nvcc for sm_100a.

    python -m stormbitmaps_b200.build [--force] [--verbose]

The shared object lands next to this file (stormbitmaps_b200/libstorm_b200.so);
it is git-ignored and travels to the GPU box with the gpurun snapshot.  nvcc
cross-compiles without a GPU, so this also runs in the CPU-only container.
"""
from __future__ import annotations

import glob
import os
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
LIB = os.path.join(PKG, "libstorm_b200.so")
OBJ = os.path.join(PKG, "_obj")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC",
    "-I", os.path.join(ROOT, "include"),
    "-I", CSRC,
]


def _newer(src_files, target):
    if not os.path.isfile(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in src_files)


def build(force: bool = False, verbose: bool = False) -> str:
    nvcc = os.environ.get("NVCC", "nvcc")
    sources = sorted(glob.glob(os.path.join(CSRC, "*.cu")))
    headers = sorted(glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(CSRC, "*.h")) +
                     glob.glob(os.path.join(ROOT, "include", "*.h")))
    os.makedirs(OBJ, exist_ok=True)
    objs, procs = [], []
    for src in sources:
        obj = os.path.join(OBJ, os.path.basename(src)[:-3] + ".o")
        objs.append(obj)
        if force or _newer([src] + headers, obj):
            cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj]
            procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0 or verbose:
            sys.stderr.write(f"--- {os.path.basename(src)}\n{out}")
        failed |= p.returncode != 0
    if failed:
        raise RuntimeError("nvcc failed")
    if force or procs or _newer(objs, LIB):
        subprocess.check_call([nvcc, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a"])
    build_tools(force)
    return LIB


TOOLS = {"contig_api_bench": os.path.join(ROOT, "tools", "contig_api_bench.c")}


def tool_path(name: str) -> str:
    return os.path.join(OBJ, name)


def build_tools(force: bool = False) -> None:
    """C programs written against include/*.h and linked with the shared object (rpath relative to the binary, so
    that they run from the gpurun snapshot): the struct-API end-to-end driver bench.py times."""
    for name, src in TOOLS.items():
        exe = tool_path(name)
        if force or _newer([src, LIB], exe):
            subprocess.check_call(["gcc", "-std=c99", "-O2", "-Wall", "-I", os.path.join(ROOT, "include"), src,
                                   "-L", PKG, "-lstorm_b200", "-Wl,-rpath,$ORIGIN/..", "-o", exe])


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
