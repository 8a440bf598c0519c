"""Build librfmc_b200.so in-tree with nvcc for sm_100a (the only target).

Every source is compiled to its own object (in parallel, only when it or a header changed) and the
objects are linked into one shared library."""

from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(CSRC, "obj")
OUT = os.path.join(HERE, "librfmc_b200.so")
SOURCES = ["abi.cu", "abi_rng.cu", "abi_dist.cu", "grow.cu", "predict.cu", "rank_encode.cu", "rng_walk.cu"]
HOST_SOURCES = ["host_rng.cpp", "mt_jump.cpp"]  # plain g++ (-O3, SIMD clones dispatched at load time)
HEADERS = [os.path.join(CSRC, "common.cuh"), os.path.join(HERE, "..", "include", "rfmc_b200.h")]
FLAGS = [
    "-O3",
    "-std=c++17",
    "-gencode",
    "arch=compute_100a,code=sm_100a",
    "-lineinfo",
    "-fmad=false",  # the reference's sums are multiply-then-add with two roundings (forest.py:211)
    "-Xcompiler",
    "-fPIC",
]


def _nvcc() -> str:
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found")
    return exe


def _sources():
    return [s for s in SOURCES + HOST_SOURCES if os.path.exists(os.path.join(CSRC, s))]


def _obj_of(src: str) -> str:
    return os.path.join(OBJ, src.rsplit(".", 1)[0] + ".o")


def _stale(target: str, deps) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def needs_build() -> bool:
    return _stale(OUT, [os.path.join(CSRC, s) for s in _sources()] + HEADERS)


def build_library(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return OUT
    os.makedirs(OBJ, exist_ok=True)
    extra = os.environ.get("RFMC_NVCC_DEFS", "").split()
    stamp = os.path.join(OBJ, ".defs")
    if (open(stamp).read() if os.path.exists(stamp) else "") != " ".join(extra):
        force = True
        with open(stamp, "w") as f:
            f.write(" ".join(extra))

    def compile_one(src: str):
        obj = _obj_of(src)
        path = os.path.join(CSRC, src)
        if not force and not _stale(obj, [path] + HEADERS):
            return obj, ""
        if src.endswith(".cu"):
            cmd = [_nvcc()] + FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + ["-c", path, "-o", obj]
        else:
            cmd = ["g++", "-O3", "-std=c++17", "-fPIC", "-c", path, "-o", obj]
        r = subprocess.run(cmd, cwd=CSRC, capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError(f"compiling {src} failed")
        return obj, r.stderr

    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as pool:
        done = list(pool.map(compile_one, _sources()))
    if verbose:
        for _, log in done:
            sys.stderr.write(log)
    cmd = [_nvcc(), "-shared", "-o", OUT] + [o for o, _ in done] + ["-lcudart", "-ldl"]
    res = subprocess.run(cmd, cwd=CSRC, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("linking librfmc_b200.so failed")
    return OUT


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv))
