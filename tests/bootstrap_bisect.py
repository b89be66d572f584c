"""Debug helper: compares the bootstrap diagonal plaintexts and the C2S / S2C transforms of the
GPU runtime with the reference one stage at a time.  python tests/bootstrap_bisect.py N depth"""
import os, sys
import numpy as np
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE)); sys.path.insert(0, HERE)

def main():
    N, depth = int(sys.argv[1]), int(sys.argv[2])
    slots = N // 2
    import ace_compiler_b200 as ace
    from oracle_bindings import RefLib, build_oracles
    build_oracles()
    ref = RefLib(N, depth, 51, 50, 3, 192, [1], with_bootstrap=True)
    ctx = ace.Context(N, depth, 51, 50, 3, hamming_weight=192)
    ctx.bootstrap_setup(slots)
    for enc in (1, 0):
        for step in range(3):
            bad, cnt = [], 0
            for idx in range(64):
                a = ctx.bootstrap_plain(slots, enc, step, idx)
                b = ref.bts_plain(slots, enc, step, idx)
                if a is None and b is None:
                    continue
                cnt += 1
                if a is None or b is None or a.shape != b.shape or (a != b).any():
                    bad.append((idx, None if a is None else a.shape, None if b is None else b.shape,
                                -1 if (a is None or b is None or a.shape != b.shape) else int((a != b).sum())))
            print("plain enc=%d step=%d: %d entries, bad: %s" % (enc, step, cnt, bad[:6]), flush=True)
    for r in ctx.bootstrap_rot_indices(slots):
        k0, k1 = ref.swk(True, r)
        ctx.import_switch_key(True, r, k0, k1)
    rng = np.random.default_rng(5)
    vals = rng.uniform(-0.5, 0.5, slots)
    for enc in (1, 0):
        ct = ref.encrypt(vals, ref.L if enc else ref.L - ctx.bootstrap_depth() + 3, slots)
        exp = ref.bts_linear(ct, enc)
        g0, g1, sc, sfd = ctx.bootstrap_linear(ct.c0, ct.c1, slots, ct.scale, ct.sf_degree, enc)
        print("linear enc=%d: level %d/%d sfd %d/%d bad limbs c0 %s c1 %s" % (
            enc, g0.shape[0], exp.level, sfd, exp.sf_degree,
            [int((g0[i] != exp.c0[i]).any()) for i in range(min(g0.shape[0], exp.level))],
            [int((g1[i] != exp.c1[i]).any()) for i in range(min(g1.shape[0], exp.level))]), flush=True)

main()
