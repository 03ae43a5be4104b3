"""Compile the sm_100a CUDA library (and the CPU oracle) in-tree.

``nvcc`` cross-compiles without a GPU, so this runs in the build container; the resulting
``ralf_b200/libralf_b200.so`` travels to the GPU box with the repository snapshot.
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libralf_b200.so")
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared",
]


def _sources() -> list[str]:
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _stale(target: str, deps: list[str]) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build_cuda(force: bool = False, verbose: bool = False) -> str:
    srcs = _sources()
    deps = srcs + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh"))]
    deps.append(os.path.join(ROOT, "include", "ralf_b200.h"))
    if force or _stale(LIB, deps):
        nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
        hdrs = [d for d in deps if not d.endswith(".cu")]
        objdir = os.path.join(HERE, "build")
        os.makedirs(objdir, exist_ok=True)
        flags = [f for f in NVCC_FLAGS if f != "-shared"]
        procs, objs = [], []
        for src in srcs:  # one nvcc per translation unit, in parallel; only stale objects are recompiled
            obj = os.path.join(objdir, os.path.basename(src)[:-3] + ".o")
            objs.append(obj)
            if force or _stale(obj, [src] + hdrs):
                cmd = [nvcc, *flags, "-c", "-o", obj, src]
                if verbose:
                    cmd.insert(1, "-Xptxas=-v")
                procs.append((cmd, subprocess.Popen(cmd)))
        for cmd, pr in procs:
            if pr.wait() != 0:
                raise subprocess.CalledProcessError(pr.returncode, cmd)
        subprocess.run([nvcc, "-arch=sm_100a", "-shared", "-o", LIB, *objs], check=True)
    return LIB


def build_oracle(force: bool = False) -> str:
    odir = os.path.join(ROOT, "oracle")
    target = os.path.join(odir, "libknn_oracle.so")
    if force or _stale(target, [os.path.join(odir, "knn_oracle.c")]):
        subprocess.run(["make", "-C", odir, "-B", "libknn_oracle.so"], check=True)
    return target


if __name__ == "__main__":
    build_cuda(force="--force" in sys.argv, verbose="-v" in sys.argv)
    build_oracle(force="--force" in sys.argv)
    print("built", LIB)
