"""In-tree nvcc build of libgmz.so (sm_100a only).  Used by __graft_entry__.build()."""
from __future__ import annotations

import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
SO = os.path.join(HERE, "libgmz.so")
SOURCES = ["gmz_engine.cu", "gmz_per.cu", "gmz_slices.cu", "gmz_tactics.cu", "gmz_hidden.cu"]
HEADERS = ["gmz_common.cuh", "gmz_tree.cuh", "gmz_play.cuh", os.path.join("..", "..", "include", "gmz.h")]
# -fmad=false: the search's float64 arithmetic must round once per operation, like the
# reference's Python floats; an FMA would change visit counts.
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-fmad=false",
              "-std=c++17", "-shared", "-Xcompiler", "-fPIC"]


def _stale() -> bool:
    if not os.path.exists(SO):
        return True
    t = os.path.getmtime(SO)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS]
    return any(os.path.exists(d) and os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not force and not _stale():
        return SO
    srcs = [os.path.join(CSRC, s) for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    cmd = [nvcc, *NVCC_FLAGS, *(["-Xptxas", "-v"] if verbose else []), "-o", SO, *srcs]
    subprocess.check_call(cmd)
    return SO
