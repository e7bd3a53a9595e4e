"""Builds gym_futbol_b200/csrc/libfutbol_b200.so in-tree with nvcc for sm_100a.

The library links only the CUDA runtime (static); torch is not involved in the build.
``python -m gym_futbol_b200.build`` or ``build_extension()``.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

CSRC = os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc")
LIB = os.path.join(CSRC, "libfutbol_b200.so")
SOURCES = ("capi.cu", "v0_kernels.cu", "v1_kernels.cu", "gae_kernel.cu", "selftest.cu")
HEADERS = ("philox.cuh", "v0_step.cuh", "v0_kernels.h", "v1_step.cuh", "v1_kernels.h", "ieee_fast.cuh", "sampler.cuh",
           os.path.join("..", "..", "include", "futbol_b200.h"))

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17", "--extended-lambda",
    "-fmad=false",            # parity: no FMA contraction in the fp64 state arithmetic
    "-Xcompiler", "-fPIC", "-Xcompiler", "-ffp-contract=off", "-shared", "-cudart", "static",
]


def _nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: cannot build libfutbol_b200.so")


def is_stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(os.path.join(CSRC, f)) > t for f in SOURCES + HEADERS)


def build_extension(force=False, verbose=False, extra_flags=(), out=None):
    """`extra_flags` / `out`: tuning variants (e.g. -DFUTBOL_MIN_BLOCKS=4 into libfutbol_b200_mb4.so).

    Safe when several processes call it at once (every rank of a torchrun job after a fresh clone): the build runs
    under an exclusive file lock, nvcc writes to a temporary name in the same directory and the finished library is
    moved into place atomically, so no process can dlopen a half-written file; a process that waited on the lock
    finds the library fresh and does not rebuild.
    """
    if out is None and not force and not is_stale():
        return LIB
    default_target = out is None
    out = LIB if out is None else os.path.join(CSRC, out)
    import fcntl
    with open(out + ".lock", "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if default_target and not force and not is_stale():      # built by another process while we waited
                return out
            tmp = "%s.tmp.%d" % (out, os.getpid())
            cmd = [_nvcc()] + NVCC_FLAGS + list(extra_flags) + (["-Xptxas", "-v"] if verbose else []) + ["-o", tmp] + list(SOURCES)
            try:
                proc = subprocess.run(cmd, cwd=CSRC, capture_output=True, text=True)
                if proc.returncode != 0:
                    raise RuntimeError("nvcc failed:\n%s\n%s" % (" ".join(cmd), proc.stderr))
                os.replace(tmp, out)
            finally:
                if os.path.exists(tmp):
                    os.remove(tmp)
            if verbose:
                sys.stderr.write(proc.stderr)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)
    return out


if __name__ == "__main__":
    print(build_extension(force="--force" in sys.argv, verbose=True))
