"""Build recipe for libb200ag.so (sm_100a only).

Invoked by ``__graft_entry__.build()`` and, on demand, by ``python -m autograd_b200.build``.
The library is built IN-TREE next to this file so that it travels with the repo
snapshot to the GPU box; nvcc cross-compiles for sm_100a without a GPU present.
"""

from __future__ import annotations

import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_PATH = os.path.join(HERE, "libb200ag.so")
STAMP_PATH = os.path.join(HERE, "csrc", ".build_stamp")

SOURCES = ["core.cu", "layout.cu", "reduce.cu", "gemm.cu", "gemm_tc.cu", "scan_sort.cu", "conv.cu", "conv_direct.cu", "comm.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC",
]


# extra compiler flags for experiments (e.g. B200AG_NVCC_EXTRA="-DB200AG_GEMM_STAGES=2"); part of the build stamp
NVCC_FLAGS += os.environ.get("B200AG_NVCC_EXTRA", "").split()


def _nvcc() -> str:
    cand = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "bin", "nvcc")
    return cand if os.path.exists(cand) else "nvcc"


def _source_hash() -> str:
    h = hashlib.sha256()
    names = sorted(os.listdir(CSRC))
    for name in names:
        if name.endswith((".cu", ".cuh", ".h")):
            with open(os.path.join(CSRC, name), "rb") as f:
                h.update(name.encode())
                h.update(f.read())
    with open(os.path.join(HERE, "..", "include", "b200ag.h"), "rb") as f:
        h.update(f.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def is_fresh() -> bool:
    if not (os.path.exists(LIB_PATH) and os.path.exists(STAMP_PATH)):
        return False
    with open(STAMP_PATH) as f:
        return f.read().strip() == _source_hash()


def build(force: bool = False, verbose: bool = True) -> str:
    """Compile every .cu under csrc/ for sm_100a and link libb200ag.so."""
    if not force and is_fresh():
        return LIB_PATH
    nvcc = _nvcc()
    objs = []

    def compile_one(src):
        obj = os.path.join(CSRC, src.replace(".cu", ".o"))
        cmd = [nvcc, *NVCC_FLAGS, "-c", os.path.join(CSRC, src), "-o", obj]
        if verbose:
            print(" ".join(cmd), flush=True)
        subprocess.run(cmd, check=True)
        return obj

    with ThreadPoolExecutor(max_workers=min(len(SOURCES), os.cpu_count() or 1)) as pool:
        objs = list(pool.map(compile_one, SOURCES))
    link = [nvcc, "-shared", "-o", LIB_PATH, *objs, "-gencode", "arch=compute_100a,code=sm_100a",
            "-lnvrtc", "-cudart", "static"]
    if verbose:
        print(" ".join(link), flush=True)
    subprocess.run(link, check=True)
    with open(STAMP_PATH, "w") as f:
        f.write(_source_hash())
    return LIB_PATH


if __name__ == "__main__":
    build(force="--force" in sys.argv)
    print(LIB_PATH)
