"""Builds libace_b200.so (hand-written sm_100a CUDA + C-ABI) in-tree with nvcc.

One object per source file (compiled in parallel, rebuilt only when the file or a header
changed), linked into the product library.  The scheduler's host-simulated self test
(csrc/sched_selftest.cu, used by tests/test_cpu_sched.py) is test scaffolding and goes into
its own library, libace_b200_selftest.so, which links against the product one."""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libace_b200.so")
SELFTEST_LIB = os.path.join(HERE, "libace_b200_selftest.so")
SOURCES = ["kernels.cu", "ntt16.cu", "kernels_ext.cu", "context.cu", "client.cu", "keyfile.cu",
           "evaluator.cu", "chebyshev.cu", "bootstrap.cu", "sched.cu", "batch.cu", "prof.cu", "peaks.cu",
           "capi.cu", "rt_shim.cu"]
SELFTEST_SOURCES = ["sched_selftest.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "--fmad=false",
              "-std=c++17", "-Xcompiler", "-fPIC", "-cudart", "shared"]


def _headers_mtime():
    hs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh"))]
    hs.append(os.path.join(HERE, "..", "include", "ace_b200.h"))
    return max(os.path.getmtime(h) for h in hs)


def _compile(src, force, verbose, hdr_t):
    obj = os.path.join(OBJ, src.replace(".cu", ".o"))
    path = os.path.join(CSRC, src)
    if (not force and os.path.exists(obj) and os.path.getmtime(obj) > os.path.getmtime(path)
            and os.path.getmtime(obj) > hdr_t):
        return obj, False
    cmd = ["nvcc"] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", path, "-o", obj]
    subprocess.run(cmd, check=True)
    return obj, True


def needs_build():
    if not os.path.exists(LIB) or not os.path.exists(SELFTEST_LIB):
        return True
    t = min(os.path.getmtime(LIB), os.path.getmtime(SELFTEST_LIB))
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    deps.append(os.path.join(HERE, "..", "include", "ace_b200.h"))
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    os.makedirs(OBJ, exist_ok=True)
    hdr_t = _headers_mtime()
    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 4)) as ex:
        res = list(ex.map(lambda s: _compile(s, force, verbose, hdr_t), SOURCES + SELFTEST_SOURCES))
    objs = [o for o, _ in res[:len(SOURCES)]]
    st_objs = [o for o, _ in res[len(SOURCES):]]
    subprocess.run(["nvcc", "-shared", "-cudart", "shared", "-o", LIB] + objs, check=True)
    subprocess.run(["nvcc", "-shared", "-cudart", "shared", "-o", SELFTEST_LIB] + st_objs +
                   ["-L" + HERE, "-lace_b200", "-Xlinker", "-rpath,$ORIGIN"], check=True)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
