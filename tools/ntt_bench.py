"""Times the batched forward / inverse NTT at the ResNet parameter set (N = 2^16, 45 limbs) with
CUDA events, checks NTT(INTT(x)) == x, and prints us per launch.  Usage: python tools/ntt_bench.py [reps]"""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ace_compiler_b200 as ace

N, DEPTH, Q0, SF, PARTS = 65536, 33, 51, 50, 3


def main():
    reps = int(sys.argv[1]) if len(sys.argv) > 1 else 50
    ctx = ace.Context(N, DEPTH, Q0, SF, PARTS)
    lib, h = ctx.lib, ctx.h
    rng = np.random.default_rng(0)
    mods = np.concatenate([ctx.q, ctx.p])
    G = ctx.L + ctx.K
    x = np.stack([rng.integers(0, mods[g], N, dtype=np.int64) for g in range(G)])
    d = ctx.put(x)

    def timeit(fn):
        for _ in range(3):
            fn()
        ctx.sync()
        lib.ace_timer_start(h)
        for _ in range(reps):
            fn()
        ms = C.c_float()
        lib.ace_timer_stop_ms(h, C.byref(ms))
        return ms.value / reps * 1e3

    t_f = timeit(lambda: lib.ace_ntt(h, d.ptr, 0, G))
    t_i = timeit(lambda: lib.ace_intt(h, d.ptr, 0, G))
    t_q = timeit(lambda: lib.ace_ntt(h, d.ptr, 0, ctx.L))
    t_p = timeit(lambda: lib.ace_ntt(h, d.ptr + ctx.L * N * 8, ctx.L, ctx.K))
    t_ip = timeit(lambda: lib.ace_intt(h, d.ptr + ctx.L * N * 8, ctx.L, ctx.K))
    t_ip2 = timeit(lambda: lib.ace_intt(h, d.ptr + (ctx.L - 11) * N * 8, ctx.L - 11, 22))
    print("intt P x%d: %.1f us   intt 11 Q + 11 P: %.1f us" % (ctx.K, t_ip, t_ip2))
    small = []
    for n in (1, 2, 4, 8, 9, 10, 16, 18, 19, 24, 27, 28, 30, 33):
        small.append("x%d %.1f/%.1f" % (n, timeit(lambda: lib.ace_ntt(h, d.ptr, 0, n)), timeit(lambda: lib.ace_intt(h, d.ptr, 0, n))))
    print("small batches (ntt/intt us): " + "  ".join(small))
    d2 = ctx.put(x)
    lib.ace_ntt(h, d2.ptr, 0, G)
    lib.ace_intt(h, d2.ptr, 0, G)
    ok = bool((d2.get() == x).all())
    print("ntt x%d: %.1f us   intt x%d: %.1f us   ntt Q x%d: %.1f us   ntt P x%d: %.1f us   roundtrip %s"
          % (G, t_f, G, t_i, ctx.L, t_q, ctx.K, t_p, "OK" if ok else "MISMATCH"))
    for form, name in ((0, "fp64"), (1, "int lazy"), (2, "int csub")):
        print("  radix-16 pass on registers, %-8s: %7.1f G butterflies/s" % (name, lib.ace_ntt_bfly_peak(h, form, 3)))
    ctx.close()


if __name__ == "__main__":
    main()
