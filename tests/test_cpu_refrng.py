"""The reference's random sources restated in the product library (csrc/refrng.h: BLAKE2Xb PRNG,
Sample_uniform / Sample_ternary on it, Sample_triangle on glibc's rand()) against the compiled
reference (oracle/_ref/libace_ref.so) when it is present, and against known-answer values taken
from it otherwise.  No GPU involved.  Reference: fhe-cmplr/rtlib/ant/include/util/prng.h:42-90,
src/util/random_sample.c:38-152."""
import ctypes as C
import os

import numpy as np
import pytest

import ace_compiler_b200 as ace

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libace_ref.so")
SEED = (C.c_uint32 * 16)(*[(0x9e3779b9 * (i + 1)) & 0xFFFFFFFF for i in range(16)])


def _vp(a):
    return a.ctypes.data_as(C.c_void_p)


def test_blake2xb_known_answer():
    """first words of buffer 7 under the harness seed (value produced by the reference's BLAKE2)"""
    lib = ace.load_library()
    a = np.zeros(3, np.uint32)
    lib.ace_refrng_words(SEED, 7, _vp(a), 3)
    assert list(a) == [209987833, 135636564, 3183613642]


def test_glibc_random_matches_libc():
    """Sample_triangle's source is libc rand() == random(): compare with this process's libc"""
    lib = ace.load_library()
    libc = C.CDLL(None)
    libc.random.restype = C.c_long
    for seed in (1, 12345, 4000000000):
        libc.srandom(C.c_uint(seed))
        want = np.array([libc.random() % 4 for _ in range(5000)])
        want = np.where(want == 0, -1, np.where(want == 1, 1, 0))
        got = np.zeros(5000, np.int64)
        lib.ace_refrng_triangle(seed, _vp(got), 5000)
        assert (got == want).all(), seed


@pytest.mark.skipif(not os.path.exists(REF_SO), reason="compiled reference not present")
def test_streams_match_reference():
    lib = ace.load_library()
    ref = C.CDLL(REF_SO)
    ref.ref_prng_words.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_size_t]
    ref.ref_sample_uniform.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_size_t, C.c_uint64]
    ref.ref_sample_ternary.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_size_t, C.c_int64]
    ref.ref_sample_triangle.argtypes = [C.c_uint32, C.c_void_p, C.c_size_t]
    a, b = np.zeros(5000, np.uint32), np.zeros(5000, np.uint32)
    for ctr in (0, 7, 1 << 40):
        lib.ace_refrng_words(SEED, ctr, _vp(a), len(a))
        ref.ref_prng_words(SEED, ctr, _vp(b), len(b))
        assert (a == b).all(), ctr
    # the ResNet primes (51, 50 and 60 bits), a 14-bit and a 33-bit modulus
    for bound in (1125899947868161, 2251799813554177, 1152921504606584833, 12289, 2 ** 32 + 15):
        x, y = np.zeros(20000, np.int64), np.zeros(20000, np.int64)
        lib.ace_refrng_uniform(SEED, 3, _vp(x), len(x), bound)
        ref.ref_sample_uniform(SEED, 3, _vp(y), len(y), bound)
        assert (x == y).all() and x.max() < bound, bound
    for hw in (192, 0, 64):
        x, y = np.zeros(4096, np.int64), np.zeros(4096, np.int64)
        lib.ace_refrng_ternary(SEED, 11, _vp(x), len(x), hw)
        ref.ref_sample_ternary(SEED, 11, _vp(y), len(y), hw)
        assert (x == y).all() and (hw == 0 or (x != 0).sum() == hw), hw
    for s in (1, 12345, 777000):
        x, y = np.zeros(70000, np.int64), np.zeros(70000, np.int64)
        lib.ace_refrng_triangle(s, _vp(x), len(x))
        ref.ref_sample_triangle(s, _vp(y), len(y))
        assert (x == y).all(), s
