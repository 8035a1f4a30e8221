"""In-tree build of liblgcnhs.so (sm_100a only) with plain nvcc.

The shared library is written next to the sources (csrc/liblgcnhs.so) so that it travels
with the repository snapshot to the GPU box; objects go to csrc/build/.  nvcc
cross-compiles without a GPU, so this also is the CPU-side "does it build" check.
"""
from __future__ import annotations

import os
import shutil
import subprocess
from concurrent.futures import ThreadPoolExecutor

CSRC = os.path.normpath(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "csrc"))
LIB_PATH = os.path.join(CSRC, "liblgcnhs.so")
SOURCES = ["capi.cu", "graph.cu", "spmm.cu", "train_ops.cu", "score_topk.cu", "score_topk_tc.cu", "spread_ops.cu", "umma_gemm.cu", "metrics.cu", "sampler.cu", "probe.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found (set NVCC or put /usr/local/cuda/bin on PATH)")


def _deps_mtime() -> float:
    hdrs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    hdrs.append(os.path.normpath(os.path.join(CSRC, "..", "..", "include", "lgcnhs.h")))
    return max(os.path.getmtime(h) for h in hdrs if os.path.exists(h))


def build_library(force: bool = False, verbose: bool = False) -> str:
    """Compile every .cu under csrc/ and link csrc/liblgcnhs.so.  Returns the library path."""
    nvcc = _nvcc()
    bdir = os.path.join(CSRC, "build")
    os.makedirs(bdir, exist_ok=True)
    hdr_m = _deps_mtime()
    objs, jobs = [], []
    for src in SOURCES:
        s = os.path.join(CSRC, src)
        o = os.path.join(bdir, src.replace(".cu", ".o"))
        objs.append(o)
        stale = force or not os.path.exists(o) or os.path.getmtime(o) < max(os.path.getmtime(s), hdr_m)
        if stale:
            jobs.append([nvcc, *NVCC_FLAGS, "-c", s, "-o", o])

    def run(cmd):
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed:\n%s\n%s" % (" ".join(cmd), r.stderr[-4000:]))
        if verbose and r.stderr.strip():
            print(r.stderr)

    if jobs:
        with ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            list(ex.map(run, jobs))
    if jobs or not os.path.exists(LIB_PATH):
        run([nvcc, "-shared", "-o", LIB_PATH, *objs, "-gencode", "arch=compute_100a,code=sm_100a",
             "-Xcompiler", "-fPIC"])
    return LIB_PATH


if __name__ == "__main__":
    print(build_library(force="--force" in os.sys.argv, verbose=True))
