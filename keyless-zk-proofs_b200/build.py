"""Builds libkzp_b200.so (sm_100a only) in-tree with nvcc. No JIT cache, no torch extension machinery:
the shared object lands next to this file so it travels with the gpurun snapshot and is the library the
tests and bench load through ctypes.

    python keyless-zk-proofs_b200/build.py [--force] [--verbose]
"""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
BUILD = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libkzp_b200.so")
CLI = os.path.join(HERE, "kzp_prove")  # command-line prover over the C ABI (csrc/cli_main.cpp)
INCLUDE = os.path.join(os.path.dirname(HERE), "include")

SOURCES = ["kernels.cu", "msm_sort.cu", "msm_g1.cu", "msm_g2.cu", "prover.cu", "capi.cu", "pool.cpp", "verify.cpp", "fullprover_abi.cu"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17", "--expt-relaxed-constexpr",
    "-Xcompiler", "-fPIC,-fvisibility=default",
    "-I", INCLUDE,
]


def _digest() -> str:
    h = hashlib.sha256()
    for root in (CSRC, INCLUDE):
        for name in sorted(os.listdir(root)):
            p = os.path.join(root, name)
            if os.path.isfile(p):
                h.update(name.encode())
                h.update(open(p, "rb").read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(BUILD, exist_ok=True)
    stamp = os.path.join(BUILD, "stamp.sha256")
    digest = _digest()
    if not force and os.path.exists(LIB) and os.path.exists(CLI) and os.path.exists(stamp) and open(stamp).read() == digest:
        return LIB
    if not os.path.exists(NVCC):
        if os.path.exists(LIB):
            # GPU box without a toolchain mismatch: use the prebuilt library that travelled with the snapshot
            return LIB
        raise RuntimeError("nvcc not found and no prebuilt libkzp_b200.so present")

    def compile_one(src: str) -> str:
        obj = os.path.join(BUILD, os.path.splitext(src)[0] + ".o")
        cmd = [NVCC] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", os.path.join(CSRC, src), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (src, r.stdout, r.stderr))
        if verbose:
            sys.stderr.write(r.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=min(len(SOURCES), os.cpu_count() or 4)) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    cmd = [NVCC, "-shared", "-o", LIB] + objs + ["-cudart", "static", "-lpthread", "-ldl", "-lrt"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n%s\n%s" % (r.stdout, r.stderr))
    cmd = [os.environ.get("CXX", "g++"), "-O2", "-std=c++17", os.path.join(CSRC, "cli_main.cpp"), "-o", CLI,
           "-L", HERE, "-lkzp_b200", "-Wl,-rpath,$ORIGIN"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("CLI build failed:\n%s\n%s" % (r.stdout, r.stderr))
    with open(stamp, "w") as f:
        f.write(digest)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
