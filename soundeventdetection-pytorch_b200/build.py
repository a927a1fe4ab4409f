"""Builds libsedb.so (hand-written sm_100a kernels + C ABI) in-tree with nvcc.

    python soundeventdetection-pytorch_b200/build.py [--bf16] [--force] [-v]

nvcc cross-compiles without a GPU.  The .so is git-ignored but travels to the GPU box with the repo snapshot.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libsedb.so")
SOURCES = ["sedb.cu"]
HEADERS = ["umma.cuh", "logmel.cuh", "resample.cuh", "probe.cuh", "cnn.cuh", "conv_issue.cuh", "cnn_host.inl", "cnn_train.cuh", "cnn_train_host.inl", "host_tables.h",
           os.path.join("..", "..", "include", "sedb.h")]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; libsedb.so cannot be built")


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.exists(d) and os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, fp16: bool | None = None, verbose: bool = False) -> str:
    if fp16 is None:
        fp16 = os.environ.get("SEDB_SPLIT_FP16", "1") == "1"
    if not force and not needs_build():
        return LIB
    cmd = [_nvcc(), "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
           "-Xptxas", "-v" if verbose else "-warn-spills", "--shared", "-Xcompiler", "-fPIC",
           f"-DSEDB_SPLIT_FP16={1 if fp16 else 0}", "-o", LIB] + [os.path.join(CSRC, s) for s in SOURCES]
    cmd += os.environ.get("SEDB_EXTRA_NVCC_FLAGS", "").split()          # development experiments only
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stdout + res.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, fp16=False if "--bf16" in sys.argv else None, verbose="-v" in sys.argv))
