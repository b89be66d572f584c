"""Generates tests/golden/*.npz from the compiled reference (oracle/_ref/libace_ref.so).
Run here (needs /root/reference to have been compiled by oracle/Makefile):
    python tests/golden/make_golden.py
Each fixture is one small parameter set; every array is the reference's own output for the
committed, seeded inputs.  One process per parameter set (the reference keeps a global
context)."""
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))

SETS = {
    "n64": (64, 3, 33, 30, 2, 0),
    "n256": (256, 5, 51, 50, 3, 0),
    "n1024": (1024, 4, 60, 56, 2, 192),
}
ROTS = [1, -3, 5]


def make(name):
    from oracle_bindings import RefLib
    N, depth, q0, sf, parts, hw = SETS[name]
    R = RefLib(N, depth, q0, sf, parts, hw, ROTS)
    rng = np.random.default_rng(2025)
    mods = np.concatenate([R.q, R.p])
    G = R.L + R.K

    def poly(gs):
        return np.stack([rng.integers(0, mods[g], N, dtype=np.int64) for g in gs])

    out = dict(params=np.array([N, depth, q0, sf, parts, hw]), q=R.q, p=R.p,
               psi=np.array([R.psi(g >= R.L, g - R.L if g >= R.L else g) for g in range(G)]),
               rots=np.array(ROTS))
    x = poly(range(G))
    out["ntt_in"] = x
    out["ntt_out"] = np.stack([R.ntt(g, x[g]) for g in range(G)])
    out["intt_out"] = np.stack([R.intt(g, x[g]) for g in range(G)])
    a, b = poly(range(G)), poly(range(G))
    out["ew_a"], out["ew_b"] = a, b
    out["modadd"] = np.stack([R.hw("modadd", g, a[g], b[g]) for g in range(G)])
    out["modmul"] = np.stack([R.hw("modmul", g, a[g], b[g]) for g in range(G)])
    for r in ROTS:
        k, order = R.auto_order(r)
        out["auto_idx_%d" % r] = np.array([k])
        out["auto_order_%d" % r] = order
        out["rotate_%d" % r] = R.hw("rotate", 0, a[0], order)
    for nq in sorted({R.L, R.L - 1, R.part_size, 1} - {0}):
        c = poly(range(nq))
        out["modup_in_%d" % nq] = c
        beta = min(R.parts, -(-nq // R.part_size))
        for part in range(beta):
            out["modup_out_%d_%d" % (nq, part)] = R.decomp_modup(c, part)
        e = poly(list(range(nq)) + [R.L + i for i in range(R.K)])
        out["moddown_in_%d" % nq] = e
        out["moddown_out_%d" % nq] = R.mod_down(e)
        if nq > 1:
            out["rescale_out_%d" % nq] = R.rescale(c)
    # ciphertext level with the reference's own (seed-pinned) keys
    slots = N // 2
    v = rng.uniform(-1, 1, slots)
    ct = R.encrypt(v, R.L, slots)
    out["msg"], out["ct_c0"], out["ct_c1"] = v, ct.c0, ct.c1
    out["sk"] = R.sk()
    k0, k1 = R.swk(False, 0)
    out["relin_k0"], out["relin_k1"] = k0, k1
    m = R.ct_mul(ct, ct)
    out["mul_c0"], out["mul_c1"] = m.c0, m.c1
    rs = R.ct_rescale(m)
    out["rs_c0"], out["rs_c1"] = rs.c0, rs.c1
    out["rs_dec"] = R.decrypt(rs)
    r = ROTS[0]
    k0, k1 = R.swk(True, r)
    out["rot_k0"], out["rot_k1"] = k0, k1
    ro = R.ct_rotate(ct, r)
    out["rot_c0"], out["rot_c1"] = ro.c0, ro.c1
    pt, sc = R.encode(v[: slots // 2], R.L - 1, slots, 2)
    out["enc_pt"] = pt
    w = rng.uniform(-0.05, 0.05, slots).astype(np.float32)
    out["encf_in"] = w
    out["encf_out"], _, _ = R.encode_float(w, 1, 2)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, "ok", {k: v.shape for k, v in list(out.items())[:4]})


if __name__ == "__main__":
    if len(sys.argv) > 1:
        make(sys.argv[1])
    else:
        for n in SETS:
            subprocess.run([sys.executable, __file__, n], check=True)
