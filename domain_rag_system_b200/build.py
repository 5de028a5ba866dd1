"""Build the C-ABI CUDA library in-tree: ``python -m domain_rag_system_b200.build``.

nvcc cross-compiles for sm_100a without a GPU.  The resulting
``csrc/libdrs_b200.so`` is git-ignored but travels with the repository snapshot.
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(CSRC, "libdrs_b200.so")
SOURCES = ["api.cu", "select.cu", "dense_simt.cu", "dense_umma.cu", "dense_pair.cu", "bm25.cu", "fuse.cu", "rerank.cu", "shard.cu", "tokenize.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "--expt-extended-lambda", "--expt-relaxed-constexpr", "-Xcompiler", "-fPIC"]


def _newer(a, b):
    return (not os.path.exists(b)) or os.path.getmtime(a) > os.path.getmtime(b)


def build(force: bool = False, verbose: bool = False) -> str:
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh"))]
    deps.append(os.path.join(os.path.dirname(HERE), "include", "drs_b200.h"))
    objs = []
    for src in SOURCES:
        s = os.path.join(CSRC, src)
        o = os.path.join(CSRC, src[:-3] + ".o")
        objs.append(o)
        hdr_new = any(_newer(d, o) for d in deps if d.endswith((".cuh", ".h")))
        if force or _newer(s, o) or hdr_new:
            cmd = [nvcc, *NVCC_FLAGS, "-c", s, "-o", o] + (["-Xptxas", "-v"] if verbose else [])
            print("[drs build]", " ".join(cmd), flush=True)
            subprocess.check_call(cmd)
    if force or any(_newer(o, LIB) for o in objs):
        cmd = [nvcc, "-shared", "-o", LIB, *objs, "-gencode", "arch=compute_100a,code=sm_100a"]
        print("[drs build]", " ".join(cmd), flush=True)
        subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(LIB)
