"""Build libhegpu.so (sm_100a) in-tree with nvcc.  Used by __graft_entry__.build()."""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libhegpu.so")
SOURCES = ["hegpu.cu"]
HEADERS = ["modarith.cuh", "ntt.cuh", "kernels.cuh", os.path.join("..", "..", "include", "hegpu.h")]
NVCC_FLAGS = [
    "-shared", "-Xcompiler", "-fPIC", "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
]


def _stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS]
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


HOST = os.path.join(HERE, "host")
HOST_SRC = ["ckks_encoder.cpp", "he_operators.cpp", "he_linalg.cpp", "he_fft.cpp"]
HOST_LIB = os.path.join(HERE, "libhe_host.so")
HOST_TEST = os.path.join(HERE, "he_host_test")


def build_host(force: bool = False) -> str:
    """C++ host mirror of the reference's interface (he_operators / he_linalg / he_fft / he_util)
    over the C ABI, plus the test driver binary."""
    deps = [os.path.join(HOST, f) for f in os.listdir(HOST)] + [LIB]
    if not force and os.path.exists(HOST_LIB) and os.path.exists(HOST_TEST):
        t = min(os.path.getmtime(HOST_LIB), os.path.getmtime(HOST_TEST))
        if all(os.path.getmtime(d) <= t for d in deps):
            return HOST_LIB
    cxx = os.environ.get("CXX", "/usr/bin/g++")
    inc = ["-I", os.path.join(HERE, "..", "include"), "-I", HOST]
    common = [cxx, "-std=c++20", "-O2", "-fPIC", "-Wall"] + inc
    link = ["-L", HERE, "-lhegpu", "-Wl,-rpath,$ORIGIN"]
    for cmd in (common + ["-shared", "-o", HOST_LIB] + [os.path.join(HOST, f) for f in HOST_SRC] + link,
                common + ["-o", HOST_TEST, os.path.join(HOST, "he_host_test.cpp"), "-L", HERE, "-lhe_host", "-lhegpu", "-Wl,-rpath,$ORIGIN"]):
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout)
            raise RuntimeError("g++ failed building the host mirror")
    return HOST_LIB


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not _stale():
        build_host()
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB] + [os.path.join(CSRC, f) for f in SOURCES]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if verbose:
        print(r.stdout)
    if r.returncode != 0:
        sys.stderr.write(r.stdout)
        raise RuntimeError("nvcc failed building libhegpu.so")
    build_host(force=True)
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(LIB)
