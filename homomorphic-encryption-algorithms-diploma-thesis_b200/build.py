"""Build libhegpu.so (sm_100a) in-tree with nvcc.  Used by __graft_entry__.build()."""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libhegpu.so")
SOURCES = ["hegpu.cu", "ntt_inst_plain_fwd.cu", "ntt_inst_plain_inv.cu", "ntt_inst_ks_intt.cu", "ntt_inst_ks_lift.cu",
           "ntt_inst_half_intt.cu", "ntt_inst_ks_moddown.cu", "ntt_inst_rescale.cu", "ntt_inst_final.cu"]
HEADERS = ["modarith.cuh", "ntt.cuh", "kernels.cuh", "mac_kernels.cuh", "imma_kernels.cuh", "internal.cuh", "ntt_launch.cuh", os.path.join("..", "..", "include", "hegpu.h")]
OBJDIR = os.environ.get("HEGPU_OBJDIR", "/tmp/hegpu_obj")  # objects are scratch: only the linked .so lives in-tree
NVCC_FLAGS = [
    "-Xcompiler", "-fPIC", "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
]


def _obj(src: str) -> str:
    return os.path.join(OBJDIR, os.path.splitext(src)[0] + ".o")


def _obj_stale(src: str) -> bool:
    o = _obj(src)
    if not os.path.exists(o):
        return True
    t = os.path.getmtime(o)
    deps = [os.path.join(CSRC, src)] + [os.path.join(CSRC, h) for h in HEADERS]
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def _stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(_obj_stale(f) or os.path.getmtime(_obj(f)) > t for f in SOURCES)


HOST = os.path.join(HERE, "host")
HOST_SRC = ["ckks_encoder.cpp", "he_operators.cpp", "he_linalg.cpp", "he_fft.cpp", "he_math.cpp", "seal_wire.cpp", "he_server.cpp"]
WIRE_LIB = os.path.join(HERE, "libseal_wire.so")  # the wire layer alone: pure host code, loadable without a GPU
HOST_LIB = os.path.join(HERE, "libhe_host.so")
HOST_TEST = os.path.join(HERE, "he_host_test")


def build_host(force: bool = False) -> str:
    """C++ host mirror of the reference's interface (he_operators / he_linalg / he_fft / he_util)
    over the C ABI, plus the test driver binary."""
    deps = [os.path.join(HOST, f) for f in os.listdir(HOST)] + [LIB]
    if not force and os.path.exists(HOST_LIB) and os.path.exists(HOST_TEST) and os.path.exists(WIRE_LIB):
        t = min(os.path.getmtime(HOST_LIB), os.path.getmtime(HOST_TEST), os.path.getmtime(WIRE_LIB))
        if all(os.path.getmtime(d) <= t for d in deps):
            return HOST_LIB
    cxx = os.environ.get("CXX", "/usr/bin/g++")
    inc = ["-I", os.path.join(HERE, "..", "include"), "-I", HOST]
    common = [cxx, "-std=c++20", "-O2", "-fPIC", "-Wall"] + inc
    link = ["-L", HERE, "-lhegpu", "-lz", "-ldl", "-Wl,-rpath,$ORIGIN"]
    for cmd in (common + ["-shared", "-o", WIRE_LIB, os.path.join(HOST, "seal_wire.cpp"), os.path.join(HOST, "seal_wire_c.cpp"), "-lz", "-ldl"],
                common + ["-shared", "-o", HOST_LIB] + [os.path.join(HOST, f) for f in HOST_SRC] + link,
                common + ["-o", HOST_TEST, os.path.join(HOST, "he_host_test.cpp"), "-L", HERE, "-lhe_host", "-lhegpu", "-lz", "-ldl", "-Wl,-rpath,$ORIGIN"]):
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout)
            raise RuntimeError("g++ failed building the host mirror")
    return HOST_LIB


def build(force: bool = False, verbose: bool = False) -> str:
    """One object per translation unit (the NTT kernels of each job resolver are their own unit),
    compiled in parallel, linked into libhegpu.so."""
    if not force and not _stale():
        build_host()
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    os.makedirs(OBJDIR, exist_ok=True)
    todo = [f for f in SOURCES if force or _obj_stale(f)]
    procs = []
    for f in todo:
        cmd = [nvcc, "-c"] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", _obj(f), os.path.join(CSRC, f)]
        procs.append((f, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for f, pr in procs:
        out, _ = pr.communicate()
        if verbose:
            print(out)
        if pr.returncode != 0:
            sys.stderr.write(out)
            failed = True
    if failed:
        raise RuntimeError("nvcc failed building libhegpu.so")
    cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB] + [_obj(f) for f in SOURCES]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout)
        raise RuntimeError("nvcc failed linking libhegpu.so")
    build_host(force=True)
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(LIB)
