"""Builds libace_b200.so (hand-written sm_100a CUDA + C-ABI) in-tree with nvcc."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libace_b200.so")
SOURCES = ["kernels.cu", "kernels_ext.cu", "context.cu", "client.cu", "evaluator.cu",
           "chebyshev.cu", "bootstrap.cu", "sched.cu", "sched_selftest.cu", "batch.cu", "prof.cu", "capi.cu", "rt_shim.cu"]


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    deps.append(os.path.join(HERE, "..", "include", "ace_b200.h"))
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    cmd = ["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "--fmad=false",
           "-std=c++17", "-shared", "-Xcompiler", "-fPIC", 
           "-cudart", "shared", "-o", LIB] + [os.path.join(CSRC, s) for s in SOURCES]
    if verbose:
        cmd.insert(1, "-Xptxas")
        cmd.insert(2, "-v")
    subprocess.run(cmd, check=True)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
