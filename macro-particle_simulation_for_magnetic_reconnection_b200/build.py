"""Builds the in-tree native libraries (explicit nvcc / g++ commands).

libmrg_fulmov.so   CUDA kernels + C ABI (include/mrg_fulmov.h), sm_100a only
libmrg_host.so     C++ host mirror of the reference's fulmov interface
Both live next to this file so they travel with the repository snapshot.
"""
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libmrg_fulmov.so")
HOSTLIB = os.path.join(HERE, "libmrg_host.so")

NVCC_FLAGS = [
    "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
    "-Xcompiler", "-fPIC", "-shared",
]


def _nvcc():
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; the CUDA library cannot be built")


def _stale(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def cuda_sources():
    srcs = [os.path.join(CSRC, f) for f in ("mrg_api.cu", "mrg_kernels.cuh", "mrg_tile.cuh", "mrg_device.cuh")]
    srcs.append(os.path.join(ROOT, "include", "mrg_fulmov.h"))
    return srcs


def build_cuda(force=False, verbose=False, out=None, defines=()):
    """out/defines build an experimental variant next to the product library
    (tools/ only; select it at run time with MRG_LIB=<path>)."""
    srcs = cuda_sources()
    target = out or LIB
    if not force and not _stale(target, srcs):
        return target
    cmd = [_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-D" + d for d in defines] + [
        "-o", target, os.path.join(CSRC, "mrg_api.cu"), "-ldl"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + r.stdout + r.stderr)
    if verbose:
        print(r.stderr)
    return target


def build_host(force=False):
    srcs = [os.path.join(CSRC, "mrg_host.cpp"), os.path.join(CSRC, "mrg_host.h"),
            os.path.join(ROOT, "include", "mrg_fulmov.h"), os.path.join(CSRC, "mrg_restart.cpp"), os.path.join(CSRC, "mrg_restart.h")]
    if not force and not _stale(HOSTLIB, srcs + [LIB]):
        return HOSTLIB
    cmd = ["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-o", HOSTLIB, srcs[0], srcs[3],
           "-L" + HERE, "-lmrg_fulmov", "-Wl,-rpath,$ORIGIN"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("g++ failed:\n" + r.stdout + r.stderr)
    return HOSTLIB


def build_all(force=False):
    build_cuda(force)
    build_host(force)
    return LIB, HOSTLIB


if __name__ == "__main__":
    import sys
    build_cuda(force=True, verbose="-v" in sys.argv)
    build_host(force=True)
    print(LIB)
    print(HOSTLIB)
