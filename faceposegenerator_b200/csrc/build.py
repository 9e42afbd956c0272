"""Builds libidb_b200.so (hand-written sm_100a kernels + C ABI) in-tree with nvcc.

    python -m faceposegenerator_b200.csrc.build [--force] [--verbose]

nvcc cross-compiles for sm_100a without a GPU; the resulting .so is git-ignored but
travels to the GPU box with the repo snapshot.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.dirname(HERE)
LIB = os.path.join(PKG, "libidb_b200.so")
SOURCES = ["api.cu", "gemm_tc.cu", "attention.cu", "attention_bwd.cu", "simt_kernels.cu", "backward_kernels.cu"]
HEADERS = ["idb_common.cuh", "idb_host.h", os.path.join("..", "..", "include", "idb.h")]
ARCH_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a"]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def is_stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(HERE, s) for s in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not is_stale():
        return LIB
    objs = []
    procs = []
    bdir = os.path.join(HERE, "build")
    os.makedirs(bdir, exist_ok=True)
    for src in SOURCES:
        obj = os.path.join(bdir, src.replace(".cu", ".o"))
        cmd = [_nvcc(), *ARCH_FLAGS, "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC",
               "--expt-relaxed-constexpr", "-c", os.path.join(HERE, src), "-o", obj]
        cmd += os.environ.get("IDB_NVCC_EXTRA", "").split()   # e.g. -DIDB_EPI_PROF=1 for the epilogue clock() breakdown
        if verbose:
            cmd += ["-Xptxas", "-v"]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    for src, pr in procs:
        out, _ = pr.communicate()
        if verbose or pr.returncode:
            sys.stderr.write(out)
        if pr.returncode:
            raise RuntimeError(f"nvcc failed on {src}")
    link = [_nvcc(), *ARCH_FLAGS, "-shared", "-o", LIB, *objs]
    r = subprocess.run(link, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode:
        sys.stderr.write(r.stdout)
        raise RuntimeError("link failed")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
