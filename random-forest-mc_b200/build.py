"""Build librfmc_b200.so in-tree with nvcc for sm_100a (the only target)."""

from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "librfmc_b200.so")
SOURCES = ["abi.cu", "grow.cu", "predict.cu", "rank_encode.cu"]
HOST_SOURCES = ["host_rng.cpp"]  # plain g++ (-O3, AVX2 clones dispatched at load time)
FLAGS = [
    "-O3",
    "-std=c++17",
    "-gencode",
    "arch=compute_100a,code=sm_100a",
    "-lineinfo",
    "-fmad=false",  # the reference's sums are multiply-then-add with two roundings (forest.py:211)
    "-Xcompiler",
    "-fPIC",
    "-shared",
]


def _nvcc() -> str:
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found")
    return exe


def needs_build() -> bool:
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "rfmc_b200.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build_library(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return OUT
    objs = []
    for src in HOST_SOURCES:
        obj = os.path.join(CSRC, src.rsplit(".", 1)[0] + ".o")
        r = subprocess.run(["g++", "-O3", "-std=c++17", "-fPIC", "-c", src, "-o", obj], cwd=CSRC, capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError(f"g++ failed on {src}")
        objs.append(obj)
    extra = os.environ.get("RFMC_NVCC_DEFS", "").split()
    cmd = [_nvcc()] + FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + ["-o", OUT] + SOURCES + objs
    res = subprocess.run(cmd, cwd=CSRC, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("nvcc failed building librfmc_b200.so")
    if verbose:
        sys.stderr.write(res.stderr)
    return OUT


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv))
