"""Exact key generation, run in its own process (the reference keeps one global context).

    python tests/keygen_case.py N depth hamming_weight [bootstrap]
    python tests/keygen_case.py model <name> [n]  (the whole key set of an emitted ResNet: its
                                                   rotation indices -- or the first n of them --
                                                   + the bootstrap keys)

The compiled reference (oracle/_ref/libace_ref.so) generates its keys with pinned randomness in
pin mode 1 (oracle/ref_harness.c: BLAKE2 PRNG seed words + counter pinned; the k-th Sample_triangle
draws from srandom(TRI_BASE + k)).  The B200 runtime generates ITS keys from the same seeds with
ace_keygen_reference (csrc/refrng.h restates the generators) and must reproduce, limb for limb: the
secret key, the public key, the relinearisation key, every rotation key (and, with `bootstrap`, the
bootstrap rotation keys and the conjugation key), and a public-key encryption made afterwards.
TEST INFRASTRUCTURE: the product never loads anything under oracle/."""
import ctypes as C
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)
TRI_BASE = 777000
SEED16 = [(0x9e3779b9 * (i + 1)) & 0xFFFFFFFF for i in range(16)]  # ref_harness.c Pin_random


def main():
    import ace_compiler_b200 as ace
    from oracle_bindings import REF_SO, RefLib, build_oracles
    build_oracles()
    if sys.argv[1] == "model":
        import json
        cfg = json.load(open(os.path.join(HERE, "emitted", sys.argv[2] + ".rots.json")))
        N, depth, hw, rots, with_bts = cfg["N"], cfg["mul_depth"], cfg["hamming_weight"], cfg["rot_idxs"], True
        assert (cfg["first_mod_size"], cfg["num_q_parts"]) == (51, 3)
        if len(sys.argv) > 3:  # only the first keys of the list (the streams are sequential)
            rots = rots[:int(sys.argv[3])]
        SF = cfg["scaling_mod_size"]
    else:
        N, depth, hw = (int(x) for x in sys.argv[1:4])
        with_bts = len(sys.argv) > 4 and sys.argv[4] == "bootstrap"
        rots = [1, -2, 5, 1, 16, -7]
        SF = 50
    h = C.CDLL(REF_SO)
    h.ref_pin_mode.argtypes = [C.c_int, C.c_uint32]
    h.ref_pin_mode(1, TRI_BASE)
    t = time.time()
    ref = RefLib(N, depth, 51, SF, 3, hw, rots, with_bootstrap=with_bts)
    t_ref = time.time() - t
    ctx = ace.Context(N, depth, 51, SF, 3, hamming_weight=hw)
    t = time.time()
    ctx.keygen_reference(SEED16, 0, TRI_BASE, rots)
    autos = []
    if with_bts:
        # Bootstrap_keygen (ckks_bootstrap_context.c:1194-1226): Generate_rot_maps over
        # Find_rot_indices, then the conjugation key
        brots = ctx.bootstrap_rot_indices(N // 2)
        ctx.keygen_rotations(0, brots)  # rotation VALUES: a value with a known automorphism regenerates it
        ctx.keygen_autos([2 * N - 1])
        autos = [ctx.auto_index(r) for r in brots] + [2 * N - 1]
    ctx.sync()
    t_gpu = time.time() - t
    print("keys: reference %.1f s, B200 runtime %.1f s" % (t_ref, t_gpu), flush=True)
    assert (ctx.export_secret_key() == ref.sk()).all(), "secret key differs"
    p0, p1 = np.zeros((ref.L, N), np.int64), np.zeros((ref.L, N), np.int64)
    ref.lib.ref_pk_export.argtypes = [C.c_void_p, C.c_void_p]
    ref.lib.ref_pk_export(p0.ctypes.data_as(C.c_void_p), p1.ctypes.data_as(C.c_void_p))
    g0, g1 = ctx.export_public_key()
    assert (g1 == p1).all() and (g0 == p0).all(), "public key differs"
    k0, k1 = ctx.export_switch_key(False)
    r0, r1 = ref.swk(False, 0)
    assert (k1 == r1).all() and (k0 == r0).all(), "relinearisation key differs"
    n_keys, bad = 0, []
    done = set()
    for rot in rots:
        a = ctx.auto_index(rot)
        if a in done:
            continue
        done.add(a)
        k0, k1 = ctx.export_switch_key(True, a)
        r0, r1 = ref.swk(True, rot)
        if not ((k1 == r1).all() and (k0 == r0).all()):
            bad.append("rotation %d (automorphism %d): a %s b %s" % (rot, a, (k1 == r1).all(), (k0 == r0).all()))
        n_keys += 1
    for i, a in enumerate(autos):
        if a in done:
            continue
        done.add(a)
        k0, k1 = ctx.export_switch_key(True, a)
        r0, r1 = ref.swk_auto(a)
        if not ((k1 == r1).all() and (k0 == r0).all()):
            bad.append("bootstrap key #%d (automorphism %d): a %s b %s" % (i, a, (k1 == r1).all(), (k0 == r0).all()))
        n_keys += 1
    assert not bad, "%d of %d switch keys differ:\n" % (len(bad), n_keys) + "\n".join(bad[:12])
    print("secret, public, relinearisation and %d rotation keys identical" % n_keys, flush=True)
    # an encryption after the keys: three more triangle draws on the same stream
    rng = np.random.default_rng(5)
    vals = rng.uniform(-1, 1, N // 2)
    lvl = ref.L
    ct = ref.encrypt(vals, lvl, N // 2)
    pt = ctx.encode(vals, lvl, N // 2, 1)
    enc = ctx.encrypt(pt, lvl, seed=1)
    got = enc.get()
    assert (got[:lvl] == ct.c0).all() and (got[lvl:] == ct.c1).all(), "encryption differs"
    print("encryption identical")
    # key file round trip (evaluation side + secret) into a fresh context
    path = "/tmp/ace_b200_keys_%d.bin" % os.getpid()
    ctx.save_keys(path, with_secret=True)
    size = os.path.getsize(path)
    ctx2 = ace.Context(N, depth, 51, SF, 3, hamming_weight=hw)
    ctx2.load_keys(path)
    os.unlink(path)
    assert (ctx2.export_secret_key() == ref.sk()).all()
    a0, a1 = ctx2.export_switch_key(True, ctx.auto_index(rots[-1]))
    b0, b1 = ref.swk(True, rots[-1])
    assert (a0 == b0).all() and (a1 == b1).all()
    # ciphertext file round trip, decrypted by the context that loaded the keys
    cpath = "/tmp/ace_b200_ct_%d.bin" % os.getpid()
    NB = N * 8
    ctx.save_ct(cpath, enc.ptr, enc.ptr + lvl * NB, lvl, N // 2, 1, float(2 ** SF))
    d, lv, sl, sfd, sc = ctx2.load_ct(cpath, lvl)
    os.unlink(cpath)
    assert (lv, sl, sfd, sc) == (lvl, N // 2, 1, float(2 ** SF))
    assert (d.get() == got).all()
    dec = ctx2.decrypt_decode(d.ptr, d.ptr + lvl * NB, lvl, N // 2, sc)
    assert np.abs(dec.real - vals).max() < 1e-6
    print("key file (%.1f MB) and ciphertext file round trips OK" % (size / 1e6))
    ctx.close()
    ctx2.close()
    print("KEYGEN PARITY OK")


if __name__ == "__main__":
    main()
