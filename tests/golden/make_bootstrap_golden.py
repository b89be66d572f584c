import sys, ctypes as C, numpy as np
sys.path.insert(0,'/root/repo/tests')
from oracle_bindings import RefLib, REF_SO
ref = C.CDLL(REF_SO)
ref.ref_coeff_collapse.restype = C.c_size_t
ref.ref_coeff_collapse.argtypes = [C.c_uint32, C.c_uint32, C.c_int, C.c_int, C.c_void_p]
out = {}
for slots in (8, 64, 512):
    for enc in (1, 0):
        for flag in (0, 1):
            b = np.zeros(3*64*slots*2)
            n = ref.ref_coeff_collapse(slots, 3, flag, enc, b.ctypes.data)
            out["d_%d_%d_%d" % (slots, enc, flag)] = b[:2*n].copy()
np.savez_compressed('/root/repo/tests/golden/bts_diagonals.npz', **out)
R = RefLib(16384, 17, 51, 50, 3, 192, [1], with_bootstrap=False)
psi = [R.psi(i >= R.L, i - R.L if i >= R.L else i) for i in range(R.L + R.K)]
np.savez('/root/repo/tests/golden/psi_n16384.npz', q=R.q, p=R.p, psi=np.array(psi, np.int64))
print("ok", len(out), psi[-3:])
