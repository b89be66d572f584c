"""CPU tests (-m "not gpu") of the host-side bootstrap set-up against the compiled reference
(oracle/_ref/libace_ref.so, when present) and committed golden values:

* the collapsed FFT diagonals of CoeffsToSlots / SlotsToCoeffs (Coeff_collapse,
  ckks_bootstrap_context.c:612-776) must be the same doubles bit for bit -- every bootstrap
  plaintext is rounded from them;
* the fixed 2N-th roots (Get_rou, fhe_std_parms.c:200-271, 336-344) that replace the generator
  search for a few primes: N = 16384 with 60-bit P primes hits them (golden psi values below
  were printed by the reference, tests/golden/make_golden.py --psi)."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest

from oracle_bindings import PortLib, REF_SO, build_oracles

HERE = os.path.dirname(os.path.abspath(__file__))





def test_fft_diagonals_match_reference():
    build_oracles()
    if not os.path.exists(REF_SO):
        pytest.skip("reference library not built")
    import ace_compiler_b200 as ace
    lib = ace.load_library()
    ref = C.CDLL(REF_SO)
    ref.ref_coeff_collapse.restype = C.c_size_t
    ref.ref_coeff_collapse.argtypes = [C.c_uint32, C.c_uint32, C.c_int, C.c_int, C.c_void_p]
    for slots in (8, 64, 512, 2048, 8192):  # 2048: remainder level; 8192: 5 collapsed layers
        for enc in (1, 0):
            for flag in (0, 1):
                cap = 3 * 64 * slots * 2
                a, b = np.zeros(cap), np.zeros(cap)
                na = lib.ace_bootstrap_fft_diagonals(slots, 3, flag, enc, a.ctypes.data)
                nb = ref.ref_coeff_collapse(slots, 3, flag, enc, b.ctypes.data)
                assert na == nb and na > 0
                assert (a[: 2 * na].view(np.int64) == b[: 2 * nb].view(np.int64)).all(), (slots, enc, flag)


def test_fft_diagonals_golden():
    """same check against committed values (runs where the reference library is absent)"""
    import ace_compiler_b200 as ace
    lib = ace.load_library()
    g = np.load(os.path.join(HERE, "golden", "bts_diagonals.npz"))
    for key in g.files:
        slots, enc, flag = (int(x) for x in key.split("_")[1:])
        a = np.zeros(3 * 64 * slots * 2)
        na = lib.ace_bootstrap_fft_diagonals(slots, 3, flag, enc, a.ctypes.data)
        assert 2 * na == g[key].size
        assert (a[: 2 * na].view(np.int64) == g[key].view(np.int64)).all(), key


def test_fixed_roots_n16384():
    build_oracles()
    g = np.load(os.path.join(HERE, "golden", "psi_n16384.npz"))
    P = PortLib(16384, 17, 51, 50, 3)
    assert (P.q == g["q"]).all() and (P.p == g["p"]).all()
    G = P.L + P.K
    assert [P.psi(i >= P.L, i - P.L if i >= P.L else i) for i in range(G)] == list(g["psi"])
