"""Per-primitive device timings (CUDA events on the context stream) at the ResNet-20
parameter set.  Usage: python tools/microbench.py [level ...]"""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ace_compiler_b200 as ace

N, DEPTH, Q0, SF, PARTS = 65536, 33, 51, 50, 3


def main():
    levels = [int(a) for a in sys.argv[1:]] or [34, 17]
    ctx = ace.Context(N, DEPTH, Q0, SF, PARTS)
    lib, h = ctx.lib, ctx.h
    rng = np.random.default_rng(0)
    mods = np.concatenate([ctx.q, ctx.p])
    G = ctx.L + ctx.K
    NB = N * 8

    def rand(gs):
        return np.stack([rng.integers(0, mods[g], N, dtype=np.int64) for g in gs])

    for is_rot, rot in [(False, 0), (True, 1)]:
        k0 = np.stack([rand(range(G)) for _ in range(PARTS)])
        k1 = np.stack([rand(range(G)) for _ in range(PARTS)])
        ctx.import_switch_key(is_rot, rot, k0, k1)

    def timeit(fn, reps=20, warm=3):
        for _ in range(warm):
            fn()
        ctx.sync()
        lib.ace_timer_start(h)
        for _ in range(reps):
            fn()
        ms = C.c_float()
        lib.ace_timer_stop_ms(h, C.byref(ms))
        return ms.value / reps * 1e3  # us

    full = ctx.put(rand(list(range(G)) * 3))
    res = {}
    res["ntt x%d limbs" % G] = timeit(lambda: lib.ace_ntt(h, full.ptr, 0, G))
    res["intt x%d limbs" % G] = timeit(lambda: lib.ace_intt(h, full.ptr, 0, G))
    res["ntt x%d limbs" % (3 * G)] = timeit(lambda: [lib.ace_ntt(h, full.ptr + k * G * NB, 0, G) for k in range(3)])
    for lv in levels:
        a = ctx.put(rand(list(range(lv)) * 2))
        b = ctx.put(rand(list(range(lv)) * 2))
        o = ctx.empty(2 * lv)
        ext = ctx.empty(lv + ctx.K, zero=True)
        a0, a1 = a.ptr, a.ptr + lv * NB
        b0, b1 = b.ptr, b.ptr + lv * NB
        o0, o1 = o.ptr, o.ptr + lv * NB
        res["L=%d hw_modmul x%d" % (lv, lv)] = timeit(lambda: lib.ace_hw_modmul(h, o0, a0, b0, 0, lv))
        res["L=%d hw_modadd x%d" % (lv, lv)] = timeit(lambda: lib.ace_hw_modadd(h, o0, a0, b0, 0, lv))
        res["L=%d decomp_modup part0" % lv] = timeit(lambda: lib.ace_decomp_modup(h, ext.ptr, a0, lv, 0))
        e2 = ctx.put(rand(list(range(lv)) + [ctx.L + i for i in range(ctx.K)]))
        res["L=%d mod_down" % lv] = timeit(lambda: lib.ace_mod_down(h, o0, e2.ptr, lv))
        res["L=%d rescale (1 poly)" % lv] = timeit(lambda: lib.ace_rescale(h, o0, a0, lv))
        res["L=%d key_switch" % lv] = timeit(lambda: lib.ace_key_switch(h, o0, o1, a1, lv, 0, 0))
        res["L=%d ct_rotate" % lv] = timeit(lambda: lib.ace_ct_rotate(h, o0, o1, a0, a1, lv, 1))
        res["L=%d ct_mul_relin" % lv] = timeit(lambda: lib.ace_ct_mul_relin(h, o0, o1, a0, a1, b0, b1, lv))
        res["L=%d ct_rescale" % lv] = timeit(lambda: lib.ace_ct_rescale(h, o0, o1, a0, a1, lv))
        # 9 rotations of one ciphertext (the input rotations of a 3x3 convolution): one call with a
        # shared ModUp against nine calls (the same key every time: only the timing matters here)
        outs = [ctx.empty(2 * lv) for _ in range(9)]
        p0 = (C.c_void_p * 9)(*[x.ptr for x in outs])
        p1 = (C.c_void_p * 9)(*[x.ptr + lv * NB for x in outs])
        r9 = (C.c_int32 * 9)(*([1] * 9))
        res["L=%d 9 x ct_rotate" % lv] = timeit(lambda: [lib.ace_ct_rotate(h, p0[k], p1[k], a0, a1, lv, 1) for k in range(9)], reps=5)
        res["L=%d ct_rotate_hoisted x9" % lv] = timeit(lambda: lib.ace_ct_rotate_hoisted(h, p0, p1, a0, a1, lv, r9, 9), reps=5)
        res["L=%d ct_mul_plain_acc" % lv] = timeit(lambda: lib.ace_ct_mul_plain_acc(h, o0, o1, a0, a1, b0, lv, 0))
    for k, v in res.items():
        print("%-32s %10.1f us" % (k, v))
    ctx.close()


if __name__ == "__main__":
    main()
