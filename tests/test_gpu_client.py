"""GPU tests (-m gpu) for the client-side pieces and the drop-in boundary:
 * run-time CKKS encode (Pt_from_msg path) is bit-exact with the reference / the oracle;
 * decrypt + decode of a reference ciphertext with the imported secret key gives the SAME
   doubles as the reference's Get_msg;
 * the runtime's own keygen / encrypt / evaluate / decrypt round-trips;
 * ACE-emitted example programs (reference sources, unmodified, compiled against our
   include/ tree by tests/build_emitted.py) run on the GPU and pass their own checks."""
import os
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FULL = (65536, 33, 51, 50, 3)


@pytest.fixture(scope="module")
def full():
    import ace_compiler_b200 as ace
    from oracle_bindings import RefLib, build_oracles
    build_oracles()
    N, depth, q0, sf, parts = FULL
    try:
        from test_gpu_parity import FULL_ROTS
        ref = RefLib(N, depth, q0, sf, parts, 192, FULL_ROTS, with_bootstrap=False)
    except RuntimeError as e:
        pytest.skip(str(e))
    ctx = ace.Context(N, depth, q0, sf, parts)
    yield ref, ctx
    ctx.close()


@pytest.mark.parametrize("params", [(1024, 5, 60, 56, 2), (4096, 7, 51, 50, 3)],
                         ids=lambda p: "N%d" % p[0])
def test_encode_vs_oracle_port(params):
    import ace_compiler_b200 as ace
    from oracle_bindings import PortLib, build_oracles
    build_oracles()
    N, depth, q0, sf, parts = params
    port = PortLib(N, depth, q0, sf, parts)
    ctx = ace.Context(N, depth, q0, sf, parts)
    rng = np.random.default_rng(21)
    for slots, length, level, deg in [(N // 2, N // 2, port.L, 1), (N // 2, 100, 3, 2),
                                      (N // 8, N // 8, 2, 1), (N // 2, 7, port.L, 3)]:
        v = rng.uniform(-1, 1, length)
        got = ctx.encode(v, level, slots, deg).get()
        assert (got == port.encode(v, level, slots, deg)).all(), (slots, length, level, deg)
    ctx.close()


def test_full_encode_float_path(full):
    """Encode_plain_from_float as called by Pt_from_msg (len 4096 .. 32768, default slots)"""
    ref, ctx = full
    rng = np.random.default_rng(22)
    for length, level, deg in [(32768, 34, 1), (4096, 20, 1), (16384, 5, 2), (8192, 1, 1)]:
        w = rng.uniform(-0.05, 0.05, length).astype(np.float32)
        exp, scale, slots = ref.encode_float(w, deg, level)
        got = ctx.encode(w.astype(np.float64), level, 0, deg).get()
        assert slots == ctx.N // 2
        assert (got == exp).all(), (length, level, deg)
    # edge cases: zeros, +-max magnitude, single non-zero
    for w in (np.zeros(32768, np.float32), np.full(32768, 0.999, np.float32),
              -np.eye(1, 32768, 5, dtype=np.float32)[0]):
        exp, _, _ = ref.encode_float(w, 1, 7)
        assert (ctx.encode(w.astype(np.float64), 7, 0, 1).get() == exp).all()


def test_full_encode_constant_path(full):
    """len == 1 -> Encode_val_at_level (bias / scalar plaintexts)"""
    ref, ctx = full
    for val, level, deg in [(0.5, 34, 1), (-0.03125, 10, 2), (1.0, 3, 1), (123.456, 20, 1),
                            (-7.0e-3, 6, 3)]:
        exp, _, _ = ref.encode_double(np.array([val]), deg, level)
        got = ctx.encode_value(val, level, deg).get()
        assert (got == exp).all(), (val, level, deg)


def test_full_decrypt_decode_exact(full):
    ref, ctx = full
    ctx.import_secret_key(ref.sk())
    rng = np.random.default_rng(23)
    slots = ref.N // 2
    v = rng.uniform(-1, 1, slots)
    ct = ref.encrypt(v, ref.L, slots)
    for level in (ref.L, 5, 1):
        from oracle_bindings import Ct
        low = Ct(ct.c0[:level], ct.c1[:level], ct.slots, ct.sf_degree, ct.scale)
        exp = ref.decrypt(low)
        d0, d1 = ctx.put(low.c0), ctx.put(low.c1)
        got = ctx.decrypt_decode(d0.ptr, d1.ptr, level, slots, low.scale)
        assert (got.real == exp).all(), level
        assert np.abs(got.real - v).max() < 1e-6


def test_own_keys_roundtrip():
    """keys from the runtime's own generator: encrypt -> HMult+relin -> rescale -> rotate ->
    decrypt is correct to CKKS precision"""
    import ace_compiler_b200 as ace
    N, depth, q0, sf, parts = 8192, 6, 55, 50, 3
    ctx = ace.Context(N, depth, q0, sf, parts)
    ctx.keygen(7, [1, -2, 5])
    rng = np.random.default_rng(24)
    slots = N // 2
    v = rng.uniform(-1, 1, slots)
    L, NB = ctx.L, N * 8
    pt = ctx.encode(v, L, slots, 1)
    ct = ctx.encrypt(pt, L, seed=3)
    c0, c1 = ct.ptr, ct.ptr + L * NB
    scale = float(1 << sf)
    assert np.abs(ctx.decrypt_decode(c0, c1, L, slots, scale).real - v).max() < 1e-7
    m = ctx.empty(2 * L)
    ctx._ck(ctx.lib.ace_ct_mul_relin(ctx.h, m.ptr, m.ptr + L * NB, c0, c1, c0, c1, L))
    r = ctx.empty(2 * (L - 1))
    ctx._ck(ctx.lib.ace_ct_rescale(ctx.h, r.ptr, r.ptr + (L - 1) * NB, m.ptr, m.ptr + L * NB, L))
    for rot in (1, -2, 5):
        o = ctx.empty(2 * (L - 1))
        ctx._ck(ctx.lib.ace_ct_rotate(ctx.h, o.ptr, o.ptr + (L - 1) * NB, r.ptr,
                                      r.ptr + (L - 1) * NB, L - 1, rot))
        got = ctx.decrypt_decode(o.ptr, o.ptr + (L - 1) * NB, L - 1, slots, scale * scale / ctx.q[L - 1] * 1.0 if False else scale)
        # after rescale the scale is Delta^2 / 2^sf (the reference tracks Delta, cipher_eval.c:57-62)
        assert np.abs(got.real - np.roll(v * v, -rot)).max() < 1e-4, rot
    ctx.close()


EXAMPLES = ["add", "add_const", "mul_const", "rotate", "rotate_02", "relin", "relin_02",
            "gemm", "gemm_02", "conv2d", "avg_pool", "relu", "bootstrap", "bootstrap_02"]


@pytest.mark.parametrize("name", EXAMPLES)
def test_emitted_example_program(name):
    exe = os.path.join(ROOT, "tests", "_emitted_bin", "eg_" + name)
    if not os.path.exists(exe):
        pytest.skip("emitted example binaries not built (tests/build_emitted.py needs /root/reference)")
    r = subprocess.run([exe], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "SUCESS!" in r.stdout, r.stdout[-2000:]
