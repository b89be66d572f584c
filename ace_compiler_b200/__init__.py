"""ace_compiler_b200 -- B200-native CKKS evaluation runtime for ACE-generated programs.

Python is only a thin ctypes veneer over the C ABI in include/ace_b200.h
(libace_b200.so: hand-written sm_100a CUDA kernels + C++ host runtime).  There is no CPU
fallback: creating a context without a CUDA device raises.
"""
import ctypes as C
import os

import numpy as np

from . import build as _build

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libace_b200.so")

vp, u32, i32, sz = C.c_void_p, C.c_uint32, C.c_int32, C.c_size_t

# name -> (restype, argtypes); must list every symbol include/ace_b200.h declares
SIGNATURES = {
    "ace_ctx_create": (C.c_int, [C.POINTER(vp), u32, sz, sz, sz, sz, sz, C.c_int]),
    "ace_ctx_destroy": (None, [vp]),
    "ace_last_error": (C.c_char_p, []),
    "ace_degree": (u32, [vp]),
    "ace_num_q": (sz, [vp]),
    "ace_num_p": (sz, [vp]),
    "ace_num_q_parts": (sz, [vp]),
    "ace_part_size": (sz, [vp]),
    "ace_get_primes": (C.c_int, [vp, vp, vp]),
    "ace_psi": (C.c_int64, [vp, u32]),
    "ace_num_decomp": (sz, [vp, sz]),
    "ace_launch_count": (C.c_uint64, [vp]),
    "ace_alloc_limbs": (vp, [vp, sz, C.c_int]),
    "ace_free_limbs": (C.c_int, [vp, vp]),
    "ace_upload": (C.c_int, [vp, vp, vp, sz]),
    "ace_download": (C.c_int, [vp, vp, vp, sz]),
    "ace_copy_limbs": (C.c_int, [vp, vp, vp, sz]),
    "ace_zero_limbs": (C.c_int, [vp, vp, sz]),
    "ace_sync": (C.c_int, [vp]),
    "ace_hw_modadd": (C.c_int, [vp, vp, vp, vp, u32, u32]),
    "ace_hw_modsub": (C.c_int, [vp, vp, vp, vp, u32, u32]),
    "ace_hw_modmul": (C.c_int, [vp, vp, vp, vp, u32, u32]),
    "ace_hw_rotate": (C.c_int, [vp, vp, vp, vp, u32, u32]),
    "ace_ntt": (C.c_int, [vp, vp, u32, u32]),
    "ace_intt": (C.c_int, [vp, vp, u32, u32]),
    "ace_decomp_modup": (C.c_int, [vp, vp, vp, u32, u32]),
    "ace_mod_down": (C.c_int, [vp, vp, vp, u32]),
    "ace_rescale": (C.c_int, [vp, vp, vp, u32]),
    "ace_auto_index": (u32, [vp, i32]),
    "ace_auto_order": (vp, [vp, i32]),
    "ace_swk_import": (C.c_int, [vp, C.c_int, i32, u32, C.c_int, vp]),
    "ace_swk_poly": (vp, [vp, C.c_int, i32, u32, C.c_int]),
    "ace_key_switch": (C.c_int, [vp, vp, vp, vp, u32, C.c_int, i32]),
    "ace_ct_rotate": (C.c_int, [vp, vp, vp, vp, vp, u32, i32]),
    "ace_ct_mul_relin": (C.c_int, [vp, vp, vp, vp, vp, vp, vp, u32]),
    "ace_ct_rotate_hoisted": (C.c_int, [vp, vp, vp, vp, vp, u32, vp, sz]),
    "ace_ct_mul_plain_acc": (C.c_int, [vp, vp, vp, vp, vp, vp, u32, C.c_int]),
    "ace_ct_rescale": (C.c_int, [vp, vp, vp, vp, vp, u32]),
    "ace_keygen": (C.c_int, [vp, C.c_uint64, vp, sz]),
    "ace_sk_import": (C.c_int, [vp, vp]),
    "ace_pk_import": (C.c_int, [vp, vp, vp]),
    "ace_encode": (C.c_int, [vp, vp, vp, sz, u32, u32, u32, u32]),
    "ace_encode_value": (C.c_int, [vp, vp, C.c_double, u32, u32]),
    "ace_encrypt": (C.c_int, [vp, vp, vp, vp, u32, C.c_uint64]),
    "ace_decrypt": (C.c_int, [vp, vp, vp, vp, u32]),
    "ace_decode": (C.c_int, [vp, vp, vp, vp, u32, u32, C.c_double]),
    "ace_bootstrap_depth": (C.c_int, [vp]),
    "ace_bootstrap_setup": (C.c_int, [vp, u32]),
    "ace_bootstrap_rot_indices": (C.c_int, [vp, u32, vp, sz]),
    "ace_bootstrap_linear": (C.c_int, [vp, vp, vp, C.POINTER(u32), C.POINTER(C.c_double), C.POINTER(u32),
                                       vp, vp, u32, u32, C.c_double, u32, C.c_int]),
    "ace_bootstrap_plain": (vp, [vp, u32, C.c_int, u32, u32, C.POINTER(u32)]),
    "ace_bootstrap_fft_diagonals": (sz, [u32, u32, C.c_int, C.c_int, vp]),
    "ace_keygen_rotations": (C.c_int, [vp, C.c_uint64, vp, sz]),
    "ace_bootstrap": (C.c_int, [vp, vp, vp, C.POINTER(u32), C.POINTER(C.c_double), C.POINTER(u32),
                                vp, vp, u32, u32, C.c_double, u32, u32]),
    "ace_measure_pipe_peaks": (C.c_int, [C.c_int, C.POINTER(C.c_double), C.c_int]),
    "ace_keygen_reference": (C.c_int, [vp, vp, C.c_uint64, u32, vp, sz]),
    "ace_keygen_reference_stream": (C.c_int, [vp, vp, C.c_uint64, u32, vp, sz, vp, sz]),
    "ace_keygen_autos": (C.c_int, [vp, vp, sz]),
    "ace_sk_export": (C.c_int, [vp, vp]),
    "ace_pk_export": (C.c_int, [vp, vp, vp]),
    "ace_swk_export": (C.c_int, [vp, C.c_int, u32, u32, C.c_int, vp]),
    "ace_keys_save": (C.c_int, [vp, C.c_char_p, C.c_int]),
    "ace_keys_load": (C.c_int, [vp, C.c_char_p]),
    "ace_ct_save": (C.c_int, [vp, C.c_char_p, vp, vp, u32, u32, u32, C.c_double]),
    "ace_ct_load": (C.c_int, [vp, C.c_char_p, vp, vp, u32, C.POINTER(u32), C.POINTER(u32), C.POINTER(u32), C.POINTER(C.c_double)]),
    "ace_refrng_words": (None, [vp, C.c_uint64, vp, sz]),
    "ace_refrng_uniform": (None, [vp, C.c_uint64, vp, sz, C.c_uint64]),
    "ace_refrng_ternary": (None, [vp, C.c_uint64, vp, sz, C.c_int64]),
    "ace_refrng_triangle": (None, [u32, vp, sz]),
    "ace_ntt_bfly_peak": (C.c_double, [vp, C.c_int, C.c_int]),
    "ace_timer_start": (C.c_int, [vp]),
    "ace_timer_stop_ms": (C.c_int, [vp, C.POINTER(C.c_float)]),
}

_lib = None


def load_library():
    """dlopen libace_b200.so (building it first if sources are newer) and bind signatures."""
    global _lib
    if _lib is None:
        _build.build()
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError if the library does not export it
            fn.restype, fn.argtypes = res, args
        _lib = lib
    return _lib


class AceError(RuntimeError):
    pass


def _hp(a):
    return a.ctypes.data_as(vp)


class DevPoly:
    """n_limbs x N device-resident limbs (owned)."""

    def __init__(self, ctx, n_limbs, zero=False):
        self.ctx, self.n_limbs = ctx, n_limbs
        self.ptr = ctx.lib.ace_alloc_limbs(ctx.h, n_limbs, int(zero))
        if not self.ptr:
            raise AceError(ctx.lib.ace_last_error().decode())

    def limb(self, i):
        return self.ptr + i * self.ctx.N * 8

    def get(self):
        out = np.empty((self.n_limbs, self.ctx.N), np.int64)
        self.ctx._ck(self.ctx.lib.ace_download(self.ctx.h, _hp(out), self.ptr, self.n_limbs))
        return out

    def free(self):
        if self.ptr:
            self.ctx.lib.ace_free_limbs(self.ctx.h, self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class Context:
    """Evaluation context on one B200 (mirrors Prepare_context's parameter part,
    fhe-cmplr/rtlib/ant/src/rtlib/context.c:29-47)."""

    def __init__(self, degree, mul_depth, first_mod_size, scaling_mod_size, num_q_parts,
                 hamming_weight=192, device=0):
        self.lib = load_library()
        h = vp()
        rc = self.lib.ace_ctx_create(C.byref(h), degree, mul_depth, first_mod_size,
                                     scaling_mod_size, num_q_parts, hamming_weight, device)
        if rc != 0:
            raise AceError(self.lib.ace_last_error().decode())
        self.h = h
        self.N = self.lib.ace_degree(h)
        self.L, self.K = self.lib.ace_num_q(h), self.lib.ace_num_p(h)
        self.parts, self.part_size = self.lib.ace_num_q_parts(h), self.lib.ace_part_size(h)
        q, p = np.zeros(self.L, np.int64), np.zeros(self.K, np.int64)
        self.lib.ace_get_primes(h, _hp(q), _hp(p))
        self.q, self.p = q, p

    def close(self):
        if self.h:
            self.lib.ace_ctx_destroy(self.h)
            self.h = None

    def _ck(self, rc):
        if rc != 0:
            raise AceError(self.lib.ace_last_error().decode())

    # ---- data movement
    def put(self, arr):
        arr = np.ascontiguousarray(arr, dtype=np.int64).reshape(-1, self.N)
        d = DevPoly(self, arr.shape[0])
        self._ck(self.lib.ace_upload(self.h, d.ptr, _hp(arr), arr.shape[0]))
        self.sync()  # the host array may be a temporary
        return d

    def empty(self, n_limbs, zero=False):
        return DevPoly(self, n_limbs, zero)

    def sync(self):
        self._ck(self.lib.ace_sync(self.h))

    def psi(self, g):
        return self.lib.ace_psi(self.h, g)

    def num_decomp(self, num_q):
        return self.lib.ace_num_decomp(self.h, num_q)

    def launch_count(self):
        return self.lib.ace_launch_count(self.h)

    # ---- per-limb ops on host arrays (convenience for tests)
    def hw(self, op, g, a, b, n_limbs=1):
        da, db = self.put(a), self.put(b)
        r = self.empty(n_limbs)
        self._ck(getattr(self.lib, "ace_hw_" + op)(self.h, r.ptr, da.ptr, db.ptr, g, n_limbs))
        return r.get()

    def ntt(self, g, a, n_limbs=1, inverse=False):
        d = self.put(a)
        f = self.lib.ace_intt if inverse else self.lib.ace_ntt
        self._ck(f(self.h, d.ptr, g, n_limbs))
        return d.get()

    def auto_index(self, rot):
        return self.lib.ace_auto_index(self.h, rot)

    def auto_order(self, rot):
        p = self.lib.ace_auto_order(self.h, rot)
        if not p:
            raise AceError(self.lib.ace_last_error().decode())
        return p

    def rotate_limbs(self, g, a, rot, n_limbs=1):
        d, r = self.put(a), self.empty(n_limbs)
        self._ck(self.lib.ace_hw_rotate(self.h, r.ptr, d.ptr, self.auto_order(rot), g, n_limbs))
        return r.get()

    # ---- polynomial-level ops
    def decomp_modup(self, a, part):
        a = np.ascontiguousarray(a)
        nq = a.shape[0]
        d, out = self.put(a), self.empty(nq + self.K, zero=True)
        self._ck(self.lib.ace_decomp_modup(self.h, out.ptr, d.ptr, nq, part))
        return out.get()

    def mod_down(self, a):
        a = np.ascontiguousarray(a)
        nq = a.shape[0] - self.K
        d, out = self.put(a), self.empty(nq)
        self._ck(self.lib.ace_mod_down(self.h, out.ptr, d.ptr, nq))
        return out.get()

    def rescale(self, a):
        a = np.ascontiguousarray(a)
        nq = a.shape[0]
        d, out = self.put(a), self.empty(nq - 1)
        self._ck(self.lib.ace_rescale(self.h, out.ptr, d.ptr, nq))
        return out.get()

    # ---- keys
    def import_switch_key(self, is_rot, rot, k0, k1):
        """k0/k1: (parts, L+K, N) host arrays as exported by the key generator"""
        for part in range(k0.shape[0]):
            a0 = np.ascontiguousarray(k0[part], dtype=np.int64)
            a1 = np.ascontiguousarray(k1[part], dtype=np.int64)
            self._ck(self.lib.ace_swk_import(self.h, int(is_rot), rot, part, 0, _hp(a0)))
            self._ck(self.lib.ace_swk_import(self.h, int(is_rot), rot, part, 1, _hp(a1)))

    # ---- fused ciphertext-level ops (host in / host out, for tests)
    def key_switch(self, d, is_rot, rot):
        nq = d.shape[0]
        dd, o0, o1 = self.put(d), self.empty(nq), self.empty(nq)
        self._ck(self.lib.ace_key_switch(self.h, o0.ptr, o1.ptr, dd.ptr, nq, int(is_rot), rot))
        return o0.get(), o1.get()

    def ct_rotate(self, c0, c1, rot):
        nq = c0.shape[0]
        d0, d1, r0, r1 = self.put(c0), self.put(c1), self.empty(nq), self.empty(nq)
        self._ck(self.lib.ace_ct_rotate(self.h, r0.ptr, r1.ptr, d0.ptr, d1.ptr, nq, rot))
        return r0.get(), r1.get()

    def ct_rotate_hoisted(self, c0, c1, rots):
        """n rotations of one ciphertext sharing one ModUp; returns [(r0, r1)] per rotation"""
        nq = c0.shape[0]
        d0, d1 = self.put(c0), self.put(c1)
        outs = [(self.empty(nq), self.empty(nq)) for _ in rots]
        p0 = (C.c_void_p * len(rots))(*[o[0].ptr for o in outs])
        p1 = (C.c_void_p * len(rots))(*[o[1].ptr for o in outs])
        r = (C.c_int32 * len(rots))(*rots)
        self._ck(self.lib.ace_ct_rotate_hoisted(self.h, p0, p1, d0.ptr, d1.ptr, nq, r, len(rots)))
        return [(a.get(), b.get()) for a, b in outs]

    def ct_mul_plain_acc(self, acc, c0, c1, pt):
        """acc (+)= ct (.) pt; acc = None starts a sum.  Returns (acc0, acc1) host arrays."""
        nq = c0.shape[0]
        d0, d1, dp = self.put(c0), self.put(c1), self.put(pt)
        if acc is None:
            a0, a1, first = self.empty(nq), self.empty(nq), 1
        else:
            a0, a1, first = self.put(acc[0]), self.put(acc[1]), 0
        self._ck(self.lib.ace_ct_mul_plain_acc(self.h, a0.ptr, a1.ptr, d0.ptr, d1.ptr, dp.ptr, nq, first))
        return a0.get(), a1.get()

    def ct_mul_relin(self, a0, a1, b0, b1):
        nq = a0.shape[0]
        da0, da1, db0, db1 = self.put(a0), self.put(a1), self.put(b0), self.put(b1)
        r0, r1 = self.empty(nq), self.empty(nq)
        self._ck(self.lib.ace_ct_mul_relin(self.h, r0.ptr, r1.ptr, da0.ptr, da1.ptr, db0.ptr,
                                           db1.ptr, nq))
        return r0.get(), r1.get()

    def ct_rescale(self, c0, c1):
        nq = c0.shape[0]
        d0, d1, r0, r1 = self.put(c0), self.put(c1), self.empty(nq - 1), self.empty(nq - 1)
        self._ck(self.lib.ace_ct_rescale(self.h, r0.ptr, r1.ptr, d0.ptr, d1.ptr, nq))
        return r0.get(), r1.get()

    # ---- bootstrap
    def bootstrap_depth(self):
        return self.lib.ace_bootstrap_depth(self.h)

    def bootstrap_rot_indices(self, slots=0):
        buf = (i32 * 4096)()
        n = self.lib.ace_bootstrap_rot_indices(self.h, slots, buf, 4096)
        if n < 0:
            self._ck(n)
        return [int(buf[i]) for i in range(n)]

    def bootstrap_setup(self, slots=0):
        self._ck(self.lib.ace_bootstrap_setup(self.h, slots))

    def keygen_rotations(self, seed, rots):
        r = (i32 * max(1, len(rots)))(*rots)
        self._ck(self.lib.ace_keygen_rotations(self.h, seed, r, len(rots)))

    def bootstrap(self, c0, c1, slots, scale, sf_degree, level_after):
        """host in / host out; returns (r0, r1, scale, sf_degree)"""
        nq = c0.shape[0]
        d0, d1 = self.put(c0), self.put(c1)
        r0, r1 = self.empty(self.L), self.empty(self.L)
        lvl, sfd, sc = u32(0), u32(0), C.c_double(0)
        self._ck(self.lib.ace_bootstrap(self.h, r0.ptr, r1.ptr, C.byref(lvl), C.byref(sc),
                                        C.byref(sfd), d0.ptr, d1.ptr, nq, slots, scale,
                                        sf_degree, level_after))
        return r0.get()[: lvl.value], r1.get()[: lvl.value], sc.value, sfd.value

    def bootstrap_linear(self, c0, c1, slots, scale, sf_degree, encoding):
        nq = c0.shape[0]
        d0, d1 = self.put(c0), self.put(c1)
        r0, r1 = self.empty(self.L), self.empty(self.L)
        lvl, sfd, sc = u32(0), u32(0), C.c_double(0)
        self._ck(self.lib.ace_bootstrap_linear(self.h, r0.ptr, r1.ptr, C.byref(lvl), C.byref(sc),
                                               C.byref(sfd), d0.ptr, d1.ptr, nq, slots, scale,
                                               sf_degree, int(encoding)))
        return r0.get()[: lvl.value], r1.get()[: lvl.value], sc.value, sfd.value

    def bootstrap_plain(self, slots, encoding, step, idx):
        lvl = u32(0)
        ptr = self.lib.ace_bootstrap_plain(self.h, slots, int(encoding), step, idx, C.byref(lvl))
        if not ptr:
            return None
        n = lvl.value + self.K
        out = np.zeros((n, self.N), np.int64)
        self._ck(self.lib.ace_download(self.h, _hp(out), ptr, n))
        return out

    # ---- client side
    def keygen(self, seed, rots):
        r = (i32 * max(1, len(rots)))(*rots)
        self._ck(self.lib.ace_keygen(self.h, seed, r, len(rots)))

    def keygen_reference(self, seed16, counter, tri_base, rots):
        """the reference's own generators, consumed in its order: bit-identical keys
        (ckks_key_generator.c:13-37; csrc/refrng.h)"""
        sd = (C.c_uint32 * 16)(*[int(x) & 0xFFFFFFFF for x in seed16])
        r = (C.c_int32 * max(1, len(rots)))(*rots)
        self._ck(self.lib.ace_keygen_reference(self.h, sd, counter, tri_base, r, len(rots)))

    def keygen_reference_stream(self, seed16, counter, srandom_seed, tri_pos, rots):
        """as keygen_reference, for a reference whose rand() is seeded once: the k-th
        Sample_triangle starts tri_pos[k] draws into the stream (the golden runs' pinning)"""
        sd = (C.c_uint32 * 16)(*[int(x) & 0xFFFFFFFF for x in seed16])
        r = (C.c_int32 * max(1, len(rots)))(*rots)
        tp = (C.c_uint64 * max(1, len(tri_pos)))(*tri_pos)
        self._ck(self.lib.ace_keygen_reference_stream(self.h, sd, counter, srandom_seed, tp, len(tri_pos),
                                                      r, len(rots)))

    def keygen_autos(self, autos):
        """more switch keys by automorphism index on the current stream (Bootstrap_keygen)"""
        a = (C.c_uint32 * max(1, len(autos)))(*autos)
        self._ck(self.lib.ace_keygen_autos(self.h, a, len(autos)))

    def export_secret_key(self):
        s = np.empty((self.L + self.K, self.N), np.int64)
        self._ck(self.lib.ace_sk_export(self.h, _hp(s)))
        return s

    def export_public_key(self):
        p0, p1 = np.empty((self.L, self.N), np.int64), np.empty((self.L, self.N), np.int64)
        self._ck(self.lib.ace_pk_export(self.h, _hp(p0), _hp(p1)))
        return p0, p1

    def export_switch_key(self, is_rot, auto_idx=0):
        """(parts, L+K, N) key0, key1 -- the layout of import_switch_key"""
        k0 = np.empty((self.parts, self.L + self.K, self.N), np.int64)
        k1 = np.empty_like(k0)
        for part in range(self.parts):
            self._ck(self.lib.ace_swk_export(self.h, int(is_rot), auto_idx, part, 0, _hp(k0[part])))
            self._ck(self.lib.ace_swk_export(self.h, int(is_rot), auto_idx, part, 1, _hp(k1[part])))
        return k0, k1

    def save_keys(self, path, with_secret=False):
        self._ck(self.lib.ace_keys_save(self.h, path.encode(), int(with_secret)))

    def load_keys(self, path):
        self._ck(self.lib.ace_keys_load(self.h, path.encode()))

    def save_ct(self, path, c0, c1, level, slots, sf_degree, scale):
        self._ck(self.lib.ace_ct_save(self.h, path.encode(), c0, c1, level, slots, sf_degree, scale))

    def load_ct(self, path, max_level):
        """returns (DevPoly holding c0 | c1 at max_level stride, level, slots, sf_degree, scale)"""
        d = self.empty(2 * max_level)
        lv, sl, sfd, sc = C.c_uint32(), C.c_uint32(), C.c_uint32(), C.c_double()
        self._ck(self.lib.ace_ct_load(self.h, path.encode(), d.ptr, d.ptr + max_level * self.N * 8, max_level,
                                      C.byref(lv), C.byref(sl), C.byref(sfd), C.byref(sc)))
        return d, lv.value, sl.value, sfd.value, sc.value

    def import_secret_key(self, sk):
        sk = np.ascontiguousarray(sk, dtype=np.int64)
        self._ck(self.lib.ace_sk_import(self.h, _hp(sk)))

    def encode(self, vals, level, slots=0, sf_degree=1, p_cnt=0):
        vals = np.ascontiguousarray(vals, dtype=np.float64)
        out = self.empty((level or self.L) + p_cnt)
        self._ck(self.lib.ace_encode(self.h, out.ptr, _hp(vals), len(vals), level, slots,
                                     sf_degree, p_cnt))
        return out

    def encode_value(self, value, level, sf_degree=1):
        out = self.empty(level or self.L)
        self._ck(self.lib.ace_encode_value(self.h, out.ptr, float(value), level, sf_degree))
        return out

    def encrypt(self, pt, level, seed=1):
        c = self.empty(2 * level)
        self._ck(self.lib.ace_encrypt(self.h, c.ptr, c.ptr + level * self.N * 8, pt.ptr, level,
                                      seed))
        return c

    def decrypt_decode(self, c0, c1, level, slots, scale):
        """c0/c1: device pointers; returns the decoded real parts (host)"""
        pt = self.empty(level)
        self._ck(self.lib.ace_decrypt(self.h, pt.ptr, c0, c1, level))
        return self.decode(pt, level, slots, scale)

    def decode(self, pt, level, slots, scale):
        re = np.zeros(slots, np.float64)
        im = np.zeros(slots, np.float64)
        self._ck(self.lib.ace_decode(self.h, _hp(re), _hp(im), pt.ptr, level, slots, scale))
        return re + 1j * im
