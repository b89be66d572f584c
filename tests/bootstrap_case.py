"""One bootstrap parity case, run in its own process (the reference keeps one global context).

    python tests/bootstrap_case.py N depth hamming_weight slots level_in level_after [q0_bits sf_bits]

The reference library (oracle/_ref/libace_ref.so) generates the keys -- including the bootstrap
rotation keys and the conjugation key of Bootstrap_keygen -- encrypts a message and bootstraps it
on the CPU; the GPU runtime imports the same keys, bootstraps the same ciphertext through the
C ABI (ace_bootstrap) and must return identical limbs, level, scale and scale degree.
TEST INFRASTRUCTURE: the product never loads anything under oracle/."""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)


def main():
    N, depth, hw, slots, level_in, level_after = (int(x) for x in sys.argv[1:7])
    q0, sf = (int(sys.argv[7]), int(sys.argv[8])) if len(sys.argv) > 8 else (51, 50)
    import ace_compiler_b200 as ace
    from oracle_bindings import RefLib, build_oracles
    build_oracles()
    t = time.time()
    ref = RefLib(N, depth, q0, sf, 3, hw, [1], with_bootstrap=True)
    ctx = ace.Context(N, depth, q0, sf, 3, hamming_weight=hw)
    rots = ctx.bootstrap_rot_indices(slots)
    print("reference init %.1fs; %d bootstrap rotation keys, depth %d" %
          (time.time() - t, len(rots), ctx.bootstrap_depth()), flush=True)
    if slots != N // 2:
        # the reference builds tables/keys for non-default slot counts on first use: trigger it
        dummy = ref.encrypt(np.zeros(slots), level_in, slots)
        ref.ct_bootstrap(dummy, level_after)
    for r in rots:
        k0, k1 = ref.swk(True, r)
        ctx.import_switch_key(True, r, k0, k1)
    k0, k1 = ref.swk_auto(2 * N - 1)
    ctx.import_switch_key(True, 2 * N - 1, k0, k1)
    k0, k1 = ref.swk(False, 0)
    ctx.import_switch_key(False, 0, k0, k1)

    rng = np.random.default_rng(1234)
    vals = rng.uniform(-0.5, 0.5, slots)
    ct = ref.encrypt(vals, level_in, slots)
    t = time.time()
    exp = ref.ct_bootstrap(ct, level_after)
    t_ref = time.time() - t
    ctx.bootstrap_setup(slots)
    t = time.time()
    g0, g1, sc, sfd = ctx.bootstrap(ct.c0, ct.c1, slots, ct.scale, ct.sf_degree, level_after)
    t_gpu = time.time() - t
    print("reference %.2fs, gpu (incl. copies) %.2fs; level %d -> %d (ref %d)" %
          (t_ref, t_gpu, level_in, g0.shape[0], exp.level), flush=True)
    assert g0.shape[0] == exp.level, (g0.shape, exp.level)
    assert sc == exp.scale and sfd == exp.sf_degree, (sc, exp.scale, sfd, exp.sf_degree)
    bad0 = int((g0 != exp.c0).sum())
    bad1 = int((g1 != exp.c1).sum())
    assert bad0 == 0 and bad1 == 0, "limb mismatch: c0 %d, c1 %d coefficients differ" % (bad0, bad1)
    dec = ref.decrypt(exp)
    err = np.abs(dec[:slots] - vals).max()
    print("bit-exact; decrypted max error vs message %.3e" % err)
    assert err < 1e-2
    ctx.close()
    print("BOOTSTRAP PARITY OK")


if __name__ == "__main__":
    main()
