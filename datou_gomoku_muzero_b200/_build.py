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


# The same library with every certified select decision re-derived by the exact float64 path and the
# disagreements counted (gmz_select_counters): loaded only by tests/test_certified_select_gpu.py and
# tools/soak_parity.py through GMZ_LIB, never by the package itself.
SO_VERIFY = os.path.join(HERE, "libgmz_verify.so")


def _stale(so=SO) -> bool:
    if not os.path.exists(so):
        return True
    t = os.path.getmtime(so)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS]
    return any(os.path.exists(d) and os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False, verify: bool = False) -> str:
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    so = SO_VERIFY if verify else SO
    if not force and not _stale(so):
        return so
    srcs = [os.path.join(CSRC, s) for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    cmd = [nvcc, *NVCC_FLAGS, *(["-DGMZ_VERIFY_FAST"] if verify else []), *(["-Xptxas", "-v"] if verbose else []), "-o", so, *srcs]
    subprocess.check_call(cmd)
    return so
