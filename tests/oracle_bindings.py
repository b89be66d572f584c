"""ctypes bindings for the two CPU checkers (TEST INFRASTRUCTURE):

* ``RefLib``  -- oracle/_ref/libace_ref.so: the unmodified reference rtlib compiled from
  /root/reference by oracle/Makefile, driven through oracle/ref_harness.c.
* ``PortLib`` -- oracle/liboracle_port.so: our plain-C restatement (oracle/ckks_oracle.c).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE = os.path.join(ROOT, "oracle")
REF_SO = os.path.join(ORACLE, "_ref", "libace_ref.so")
PORT_SO = os.path.join(ORACLE, "liboracle_port.so")

vp = C.c_void_p
u32, i32, sz, dbl = C.c_uint32, C.c_int32, C.c_size_t, C.c_double


def _p(a):
    return a.ctypes.data_as(vp)


def build_oracles():
    """(re)build the checkers; a no-op for _ref when /root/reference is absent."""
    subprocess.run(["make", "-s", "-C", ORACLE, "all"], check=True,
                   stdout=subprocess.DEVNULL)


class RefCt(C.Structure):
    _fields_ = [("c0", vp), ("c1", vp), ("level", u32), ("slots", u32),
                ("sf_degree", u32), ("scale", dbl)]


class Ct:
    """host-side ciphertext: two (level, N) int64 arrays + CKKS bookkeeping"""

    def __init__(self, c0, c1, slots, sf_degree, scale):
        self.c0, self.c1 = np.ascontiguousarray(c0), np.ascontiguousarray(c1)
        self.slots, self.sf_degree, self.scale = slots, sf_degree, scale

    @property
    def level(self):
        return self.c0.shape[0]


class RefLib:
    """One reference context per process (the reference keeps a global `Context`)."""
    _inited = None

    def __init__(self, N, depth, q0_bits, sf_bits, parts, hw, rots, with_bootstrap=False):
        key = (N, depth, q0_bits, sf_bits, parts, hw, tuple(rots), with_bootstrap)
        if RefLib._inited is not None and RefLib._inited != key:
            raise RuntimeError("reference context already initialised with other parameters "
                               "(run in a fresh process)")
        L = self.lib = C.CDLL(REF_SO)
        L.ref_init.argtypes = [u32, sz, sz, sz, sz, sz, vp, sz, C.c_int]
        for f in ("ref_num_q", "ref_num_p", "ref_num_parts", "ref_part_size"):
            getattr(L, f).restype = sz
        L.ref_default_scale.restype = dbl
        L.ref_psi.restype = C.c_int64
        L.ref_psi.argtypes = [C.c_int, sz]
        for f in ("ref_ntt", "ref_intt"):
            getattr(L, f).argtypes = [C.c_int, sz, vp]
        for f in ("ref_hw_modadd", "ref_hw_modmul", "ref_hw_rotate"):
            getattr(L, f).argtypes = [vp, vp, vp, C.c_int, sz]
        L.ref_auto_order.argtypes = [i32, vp]
        for f in ("ref_decomp_modup", "ref_decomp_then_modup"):
            getattr(L, f).argtypes = [vp, vp, sz, u32]
        L.ref_mod_down.argtypes = [vp, vp, sz]
        L.ref_rescale.argtypes = [vp, vp, sz]
        L.ref_swk_export.argtypes = [C.c_int, i32, u32, C.c_int, vp]
        L.ref_swk_export_auto.argtypes = [u32, u32, C.c_int, vp]
        L.ref_encode.argtypes = [vp, vp, sz, u32, u32, u32, vp]
        L.ref_encode_float.argtypes = [vp, vp, sz, u32, u32, vp, vp]
        L.ref_encode_double.argtypes = [vp, vp, sz, u32, u32, vp, vp]
        L.ref_encrypt.argtypes = [vp, vp, sz, u32, u32]
        L.ref_decrypt.argtypes = [vp, vp]
        L.ref_ct_mul_plain.argtypes = [vp, vp, vp, dbl, u32]
        L.ref_ct_rotate.argtypes = [vp, vp, i32]
        L.ref_ct_bootstrap.argtypes = [vp, vp, u32]
        L.ref_bts_linear.argtypes = [vp, vp, C.c_int]
        L.ref_bts_plain.argtypes = [u32, C.c_int, u32, u32, vp]
        if RefLib._inited is None:
            r = (i32 * max(1, len(rots)))(*rots)
            L.ref_init(N, depth, q0_bits, sf_bits, parts, hw, r, len(rots), int(with_bootstrap))
            RefLib._inited = key
        self.N, self.L, self.K = N, L.ref_num_q(), L.ref_num_p()
        self.parts, self.part_size = L.ref_num_parts(), L.ref_part_size()
        q = np.zeros(self.L, np.int64)
        p = np.zeros(self.K, np.int64)
        L.ref_get_primes(_p(q), _p(p))
        self.q, self.p = q, p

    def psi(self, is_p, idx):
        return self.lib.ref_psi(int(is_p), idx)

    def _split(self, g):
        return (1, g - self.L) if g >= self.L else (0, g)

    def ntt(self, g, a):
        a = np.array(a, dtype=np.int64)
        self.lib.ref_ntt(*self._split(g), _p(a))
        return a

    def intt(self, g, a):
        a = np.array(a, dtype=np.int64)
        self.lib.ref_intt(*self._split(g), _p(a))
        return a

    def hw(self, op, g, a, b):
        r = np.empty(self.N, np.int64)
        a, b = np.ascontiguousarray(a), np.ascontiguousarray(b)
        getattr(self.lib, "ref_hw_" + op)(_p(r), _p(a), _p(b), *self._split(g))
        return r

    def auto_order(self, rot):
        o = np.empty(self.N, np.int64)
        k = self.lib.ref_auto_order(rot, _p(o))
        return k, o

    def decomp_modup(self, a, part, fused=True):
        a = np.ascontiguousarray(a)
        nq = a.shape[0]
        out = np.zeros((nq + self.K, self.N), np.int64)
        f = self.lib.ref_decomp_modup if fused else self.lib.ref_decomp_then_modup
        f(_p(out), _p(a), nq, part)
        return out

    def mod_down(self, a):
        a = np.array(a, dtype=np.int64)  # copy: the reference clobbers the P part
        nq = a.shape[0] - self.K
        out = np.zeros((nq, self.N), np.int64)
        self.lib.ref_mod_down(_p(out), _p(a), nq)
        return out

    def rescale(self, a):
        a = np.ascontiguousarray(a)
        nq = a.shape[0]
        out = np.zeros((nq, self.N), np.int64)
        self.lib.ref_rescale(_p(out), _p(a), nq)
        return out[: nq - 1]

    def swk(self, is_rot, rot):
        """returns key0, key1 of shape (parts, L+K, N)"""
        k0 = np.zeros((self.parts, self.L + self.K, self.N), np.int64)
        k1 = np.zeros_like(k0)
        for part in range(self.parts):
            n = self.lib.ref_swk_export(int(is_rot), rot, part, 0, _p(k0[part]))
            assert n == self.L + self.K, n
            self.lib.ref_swk_export(int(is_rot), rot, part, 1, _p(k1[part]))
        return k0, k1

    def swk_auto(self, auto_idx):
        """switch key looked up by automorphism index (conjugation key: 2N-1)"""
        k0 = np.zeros((self.parts, self.L + self.K, self.N), np.int64)
        k1 = np.zeros_like(k0)
        for part in range(self.parts):
            n = self.lib.ref_swk_export_auto(auto_idx, part, 0, _p(k0[part]))
            assert n == self.L + self.K, (auto_idx, n)
            self.lib.ref_swk_export_auto(auto_idx, part, 1, _p(k1[part]))
        return k0, k1

    def sk(self):
        s = np.zeros((self.L + self.K, self.N), np.int64)
        self.lib.ref_sk_export(_p(s))
        return s

    def encode(self, vals, level, slots, sf_degree=1):
        vals = np.ascontiguousarray(vals, dtype=np.float64)
        out = np.zeros((level, self.N), np.int64)
        sc = dbl()
        self.lib.ref_encode(_p(out), _p(vals), len(vals), level, slots, sf_degree, C.byref(sc))
        return out, sc.value

    def encode_float(self, vals, sc_degree, level):
        vals = np.ascontiguousarray(vals, dtype=np.float32)
        lvl = level if level else self.L
        out = np.zeros((lvl, self.N), np.int64)
        sc, sl = dbl(), u32()
        self.lib.ref_encode_float(_p(out), _p(vals), len(vals), sc_degree, level,
                                  C.byref(sc), C.byref(sl))
        return out, sc.value, sl.value

    def encode_double(self, vals, sc_degree, level):
        vals = np.ascontiguousarray(vals, dtype=np.float64)
        lvl = level if level else self.L
        out = np.zeros((lvl, self.N), np.int64)
        sc, sl = dbl(), u32()
        self.lib.ref_encode_double(_p(out), _p(vals), len(vals), sc_degree, level,
                                   C.byref(sc), C.byref(sl))
        return out, sc.value, sl.value

    # ---- ciphertext level
    def _rc(self, ct):
        return RefCt(_p(ct.c0), _p(ct.c1), ct.level, ct.slots, ct.sf_degree, ct.scale)

    def _new(self, level):
        return Ct(np.zeros((level, self.N), np.int64), np.zeros((level, self.N), np.int64),
                  0, 0, 0.0)

    def _fin(self, out, rc):
        out.c0, out.c1 = out.c0[: rc.level], out.c1[: rc.level]
        out.slots, out.sf_degree, out.scale = rc.slots, rc.sf_degree, rc.scale
        return out

    def encrypt(self, vals, level, slots):
        vals = np.ascontiguousarray(vals, dtype=np.float64)
        out = self._new(level)
        rc = self._rc(out)
        self.lib.ref_encrypt(C.byref(rc), _p(vals), len(vals), level, slots)
        return self._fin(out, rc)

    def decrypt(self, ct):
        out = np.zeros(ct.slots, np.float64)
        rc = self._rc(ct)
        self.lib.ref_decrypt(_p(out), C.byref(rc))
        return out

    def _unary(self, fn, ct, *args, out_level=None):
        out = self._new(out_level or ct.level)
        rc, ra = self._rc(out), self._rc(ct)
        fn(C.byref(rc), C.byref(ra), *args)
        return self._fin(out, rc)

    def ct_rotate(self, ct, rot):
        return self._unary(self.lib.ref_ct_rotate, ct, rot)

    def ct_rescale(self, ct):
        return self._unary(self.lib.ref_ct_rescale, ct)

    def ct_bootstrap(self, ct, level_after):
        return self._unary(self.lib.ref_ct_bootstrap, ct, level_after, out_level=self.L)

    def bts_linear(self, ct, encoding):
        return self._unary(self.lib.ref_bts_linear, ct, int(encoding), out_level=self.L)

    def bts_plain(self, slots, encoding, step, idx):
        """diagonal plaintext (num_q + K limbs) of the bootstrap tables, or None"""
        out = np.zeros((self.L + self.K, self.N), np.int64)
        nq = self.lib.ref_bts_plain(slots, int(encoding), step, idx, _p(out))
        return None if nq <= 0 else out[: nq + self.K]

    def ct_mul_plain(self, ct, pt, pt_scale, pt_sf_degree=1):
        pt = np.ascontiguousarray(pt)
        return self._unary(self.lib.ref_ct_mul_plain, ct, _p(pt), pt_scale, pt_sf_degree)

    def _binary(self, fn, a, b):
        out = self._new(min(a.level, b.level))
        rc, ra, rb = self._rc(out), self._rc(a), self._rc(b)
        fn(C.byref(rc), C.byref(ra), C.byref(rb))
        return self._fin(out, rc)

    def ct_add(self, a, b):
        return self._binary(self.lib.ref_ct_add, a, b)

    def ct_mul(self, a, b):
        return self._binary(self.lib.ref_ct_mul, a, b)


class RefModel:
    """An ACE-emitted model unit (oracle/_ref/<model>_ref.so) running on the reference runtime,
    driven through the reference's own Prepare_context / Prepare_input / Run_main_graph /
    Handle_output.  One per process.  Randomness is pinned by the harness, so keys, the
    encrypted input and every output limb are reproducible."""

    def __init__(self, model, data_file):
        # RTLD_LOCAL + RTLD_NOW: every symbol of the reference pair is bound now and stays private,
        # so the B200 runtime (same function names) can be loaded into the same process later
        self.lib = L = C.CDLL(REF_SO, mode=os.RTLD_LOCAL | os.RTLD_NOW)
        self.unit = C.CDLL(os.path.join(os.path.dirname(REF_SO), model + "_ref.so"),
                           mode=os.RTLD_LOCAL | os.RTLD_NOW)
        L.ref_set_callbacks.argtypes = [vp] * 5
        self.unit.model_register.argtypes = [vp]
        self.unit.model_register(C.cast(L.ref_set_callbacks, vp))
        L.ref_init_emitted.argtypes = [C.c_char_p]
        rc = L.ref_init_emitted(data_file.encode())
        if rc != 0:
            raise RuntimeError("ref_init_emitted failed: %d" % rc)
        RefLib._inited = ("model", model)
        for f in ("ref_num_q", "ref_num_p", "ref_num_parts", "ref_part_size"):
            getattr(L, f).restype = sz
        L.ref_prepare_input.argtypes = [vp, sz, sz, sz, sz, C.c_char_p]
        L.ref_peek_input.argtypes = [C.c_char_p, vp]
        L.ref_peek_output.argtypes = [C.c_char_p, vp]
        L.ref_handle_output.argtypes = [C.c_char_p, vp, sz]
        L.ref_swk_export.argtypes = [C.c_int, i32, u32, C.c_int, vp]
        L.ref_swk_export_auto.argtypes = [u32, u32, C.c_int, vp]
        L.ref_decrypt.argtypes = [vp, vp]
        self.N, self.L, self.K = L.ref_degree(), L.ref_num_q(), L.ref_num_p()
        self.parts = L.ref_num_parts()

    swk = RefLib.swk
    swk_auto = RefLib.swk_auto
    _rc = RefLib._rc
    decrypt = RefLib.decrypt

    def _peek(self, fn, name):
        out = Ct(np.zeros((self.L, self.N), np.int64), np.zeros((self.L, self.N), np.int64), 0, 0, 0.0)
        rc = RefCt(_p(out.c0), _p(out.c1), self.L, 0, 0, 0.0)
        if fn(name.encode(), C.byref(rc)) < 0:
            raise RuntimeError("no ciphertext named " + name)
        out.c0, out.c1 = out.c0[: rc.level], out.c1[: rc.level]
        out.slots, out.sf_degree, out.scale = rc.slots, rc.sf_degree, rc.scale
        return out

    def prepare_input(self, image, name="input"):
        v = np.ascontiguousarray(image, dtype=np.float64)
        self.lib.ref_prepare_input(_p(v), 1, 3, 32, 32, name.encode())

    def peek_input(self, name="input"):
        return self._peek(self.lib.ref_peek_input, name)

    def run(self):
        self.lib.ref_run_main_graph()

    def peek_output(self, name="output"):
        return self._peek(self.lib.ref_peek_output, name)

    def handle_output(self, n, name="output"):
        out = np.zeros(n)
        self.lib.ref_handle_output(name.encode(), _p(out), n)
        return out


class PortLib:
    def __init__(self, N, depth, q0_bits, sf_bits, parts):
        L = self.lib = C.CDLL(PORT_SO)
        L.orc_create.restype = vp
        L.orc_create.argtypes = [u32, sz, sz, sz, sz]
        L.orc_destroy.argtypes = [vp]
        for f in ("orc_num_q", "orc_num_p", "orc_part_size"):
            getattr(L, f).restype = sz
            getattr(L, f).argtypes = [vp]
        L.orc_num_decomp.restype = sz
        L.orc_num_decomp.argtypes = [vp, sz]
        L.orc_get_primes.argtypes = [vp, vp, vp]
        L.orc_psi.restype = C.c_int64
        L.orc_psi.argtypes = [vp, C.c_int, sz]
        L.orc_ntt.argtypes = [vp, sz, vp]
        L.orc_intt.argtypes = [vp, sz, vp]
        for f in ("orc_hw_modadd", "orc_hw_modmul", "orc_hw_rotate"):
            getattr(L, f).argtypes = [vp, vp, vp, vp, sz]
        L.orc_auto_index.restype = u32
        L.orc_auto_index.argtypes = [vp, i32]
        L.orc_auto_order.argtypes = [vp, u32, vp]
        L.orc_decomp_modup.argtypes = [vp, vp, vp, sz, sz]
        L.orc_mod_down.argtypes = [vp, vp, vp, sz]
        L.orc_rescale.argtypes = [vp, vp, vp, sz]
        L.orc_key_switch.argtypes = [vp, vp, vp, vp, sz, vp, vp]
        L.orc_ct_rotate.argtypes = [vp, vp, vp, vp, vp, sz, u32, vp, vp]
        L.orc_ct_mul_relin.argtypes = [vp] * 7 + [sz, vp, vp]
        L.orc_encode.argtypes = [vp, vp, vp, sz, u32, u32, u32]
        self.ctx = L.orc_create(N, depth, q0_bits, sf_bits, parts)
        self.N, self.L, self.K = N, L.orc_num_q(self.ctx), L.orc_num_p(self.ctx)
        self.parts, self.part_size = parts, L.orc_part_size(self.ctx)
        q = np.zeros(self.L, np.int64)
        p = np.zeros(self.K, np.int64)
        L.orc_get_primes(self.ctx, _p(q), _p(p))
        self.q, self.p = q, p

    def __del__(self):
        try:
            self.lib.orc_destroy(self.ctx)
        except Exception:
            pass

    def psi(self, is_p, idx):
        return self.lib.orc_psi(self.ctx, int(is_p), idx)

    def ntt(self, g, a):
        a = np.array(a, dtype=np.int64)
        self.lib.orc_ntt(self.ctx, g, _p(a))
        return a

    def intt(self, g, a):
        a = np.array(a, dtype=np.int64)
        self.lib.orc_intt(self.ctx, g, _p(a))
        return a

    def hw(self, op, g, a, b):
        r = np.empty(self.N, np.int64)
        a, b = np.ascontiguousarray(a), np.ascontiguousarray(b)
        getattr(self.lib, "orc_hw_" + op)(self.ctx, _p(r), _p(a), _p(b), g)
        return r

    def auto_order(self, rot):
        k = self.lib.orc_auto_index(self.ctx, rot)
        o = np.empty(self.N, np.int64)
        self.lib.orc_auto_order(self.ctx, k, _p(o))
        return k, o

    def num_decomp(self, nq):
        return self.lib.orc_num_decomp(self.ctx, nq)

    def decomp_modup(self, a, part):
        a = np.ascontiguousarray(a)
        nq = a.shape[0]
        out = np.zeros((nq + self.K, self.N), np.int64)
        self.lib.orc_decomp_modup(self.ctx, _p(out), _p(a), nq, part)
        return out

    def mod_down(self, a):
        a = np.ascontiguousarray(a)
        nq = a.shape[0] - self.K
        out = np.zeros((nq, self.N), np.int64)
        self.lib.orc_mod_down(self.ctx, _p(out), _p(a), nq)
        return out

    def rescale(self, a):
        a = np.ascontiguousarray(a)
        nq = a.shape[0]
        out = np.zeros((nq - 1, self.N), np.int64)
        self.lib.orc_rescale(self.ctx, _p(out), _p(a), nq)
        return out

    def key_switch(self, d, k0, k1):
        d, k0, k1 = map(np.ascontiguousarray, (d, k0, k1))
        nq = d.shape[0]
        o0, o1 = np.zeros((nq, self.N), np.int64), np.zeros((nq, self.N), np.int64)
        self.lib.orc_key_switch(self.ctx, _p(o0), _p(o1), _p(d), nq, _p(k0), _p(k1))
        return o0, o1

    def ct_rotate(self, c0, c1, auto_idx, k0, k1):
        c0, c1, k0, k1 = map(np.ascontiguousarray, (c0, c1, k0, k1))
        nq = c0.shape[0]
        o0, o1 = np.zeros((nq, self.N), np.int64), np.zeros((nq, self.N), np.int64)
        self.lib.orc_ct_rotate(self.ctx, _p(o0), _p(o1), _p(c0), _p(c1), nq, auto_idx,
                               _p(k0), _p(k1))
        return o0, o1

    def ct_mul_relin(self, a0, a1, b0, b1, k0, k1):
        a0, a1, b0, b1, k0, k1 = map(np.ascontiguousarray, (a0, a1, b0, b1, k0, k1))
        nq = a0.shape[0]
        o0, o1 = np.zeros((nq, self.N), np.int64), np.zeros((nq, self.N), np.int64)
        self.lib.orc_ct_mul_relin(self.ctx, _p(o0), _p(o1), _p(a0), _p(a1), _p(b0), _p(b1),
                                  nq, _p(k0), _p(k1))
        return o0, o1

    def encode(self, vals, level, slots, sf_degree=1):
        vals = np.ascontiguousarray(vals, dtype=np.float64)
        out = np.zeros((level, self.N), np.int64)
        self.lib.orc_encode(self.ctx, _p(out), _p(vals), len(vals), level, slots, sf_degree)
        return out
