"""Host-only test of the deferred-execution scheduler (ace_compiler_b200/csrc/sched.h): random
programs over the recorded polynomial-level API must compute, deferred (waves, batching, mul+add
fusion, dead-store elimination, deferred frees), exactly what call-by-call execution computes.
A test-only library (libace_b200_selftest.so = csrc/sched_selftest.cu, linked against the product
library but not part of it) runs both on a host-simulated backend; no GPU involved.
Reference semantics = the reference executes every Hw_* / Decomp_modup / Mod_down / Rescale call
immediately (fhe-cmplr/rtlib/ant/src/poly/poly_arith.c:14-56, poly_eval.c:28-49)."""
import ctypes as C

import pytest

import ace_compiler_b200 as ace


class _Lib:
    _h = None


def _selftest_lib():
    if _Lib._h is None:
        ace.load_library()
        from ace_compiler_b200 import build as b
        h = C.CDLL(b.SELFTEST_LIB)
        h.ace_sched_selftest.restype = C.c_int
        h.ace_sched_selftest.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32, C.POINTER(C.c_size_t)]
        _Lib._h = h
    return _Lib._h


@pytest.mark.parametrize("seed", range(1, 41))
def test_deferred_equals_eager_random_program(seed):
    lib = _selftest_lib()
    stats = (C.c_size_t * 8)()
    rc = lib.ace_sched_selftest(seed, 6000, 30, stats)
    assert rc == 0, "mismatch at synchronisation point %d (seed %d)" % (rc - 1, seed)
    ops, flushes, waves, fused, dead, chains, modups, shared = list(stats)
    assert ops >= 6000 and waves >= flushes > 0


def test_scheduler_actually_defers():
    """the deferred run must batch (few waves per op), fuse and drop stores -- otherwise the test
    above compares eager with eager"""
    lib = _selftest_lib()
    stats = (C.c_size_t * 8)()
    assert lib.ace_sched_selftest(12345, 20000, 30, stats) == 0
    ops, flushes, waves, fused, dead, chains, modups, shared = list(stats)
    assert waves < ops / 3 and fused > 100 and dead > 100, list(stats)
    # repeated Decomp_modup calls of an unmodified polynomial are served from the first result
    assert modups > 500 and shared > 50, list(stats)


@pytest.mark.parametrize("seed,sync_permille", [(s, p) for s in range(200, 212) for p in (1, 3)])
def test_long_deferred_windows(seed, sync_permille):
    """windows of hundreds to thousands of recorded ops between synchronisation points (what a
    convolution layer between two bootstraps looks like)"""
    lib = _selftest_lib()
    rc = lib.ace_sched_selftest(seed, 30000, sync_permille, None)
    assert rc == 0, "mismatch at synchronisation point %d" % (rc - 1)


def test_one_window_hits_the_flush_threshold():
    """no synchronisation at all: the scheduler flushes by itself at 2^18 recorded ops, chains are
    longer than one launch, the limb table is rehashed"""
    lib = _selftest_lib()
    stats = (C.c_size_t * 8)()
    assert lib.ace_sched_selftest(99, 400000, 0, stats) == 0
    assert stats[1] >= 2  # flushed on its own at least once before the final synchronisation
