"""Helper of tests/test_gpu_ptmgr.py, run in its own process (the reference keeps one global
context per process).

    python tests/ptmgr_case.py make <dir>     messages + DE_PLAINTEXT file made with the reference's
                                              Encode_plain_buffer (plain_eval.c:98-124)
    python tests/ptmgr_case.py layout <dir>   the same, then read back through the REFERENCE's own
                                              Pt_mgr_init / Pt_get (pt_mgr.c:35-159)
TEST INFRASTRUCTURE."""
import ctypes as C
import os
import struct
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libace_ref.so")
PAGE = 4096
N_ENT, MSG_LEN = 5, 2048


def write_plaintext_file(path, bufs, model="pt_get_case"):
    """rt_data_def.h:17-32,90-109: header page, entries aligned to 2^12, look-up table"""
    lut, ofs = [], PAGE
    with open(path, "wb") as f:
        f.write(b"\0" * PAGE)
        for i, b in enumerate(bufs):
            f.write(b)
            lut.append((b"cst_%d" % i, i, len(b), ofs))
            ofs += len(b)
            pad = (-ofs) % PAGE
            f.write(b"\0" * pad)
            ofs += pad
        for name, idx, size, o in lut:
            f.write(struct.pack("<16sIIQ", name, idx, size, o))
        f.seek(0)
        f.write(struct.pack("<8sIHBBQQqq48s40s", b"!ANTFHE\0", 1, 0, 2, 12, len(bufs), ofs, 0, 0,
                            model.encode(), b"XXXXXXXX-XXXX-XXXX-XXXX-XXXXXXXXXXXX"))


def make_inputs(tmp):
    """messages + their PLAINTEXT_BUFFERs from the reference (same parameters as pt_get_case.c)"""
    from oracle_bindings import RefLib
    ref = RefLib(4096, 7, 51, 50, 3, 192, [], with_bootstrap=False)
    L = ref.lib
    L.Encode_plain_buffer.restype = C.c_void_p
    L.Encode_plain_buffer.argtypes = [C.c_void_p, C.c_size_t, C.c_uint32, C.c_uint32]
    L.Free_plain_buffer.argtypes = [C.c_void_p]
    rng = np.random.default_rng(11)
    msgs = rng.uniform(-1, 1, (N_ENT, MSG_LEN)).astype(np.float32)
    bufs = []
    for i in range(N_ENT):
        p = L.Encode_plain_buffer(msgs[i].ctypes.data, MSG_LEN, 1, 0)
        magic, ver, size = struct.unpack("<8sII", C.string_at(p, 16))
        assert magic == b"ANTPLAIN" and ver == 1
        bufs.append(C.string_at(p, 16 + size))
        L.Free_plain_buffer(p)
    msg_path, pt_path = os.path.join(tmp, "msgs.bin"), os.path.join(tmp, "weights.pt")
    msgs.tofile(msg_path)
    write_plaintext_file(pt_path, bufs)
    return msg_path, pt_path



def layout(tmp):
    msg_path, pt_path = make_inputs(tmp)
    L = C.CDLL(REF_SO)
    L.Pt_mgr_init.restype = C.c_bool
    L.Pt_mgr_init.argtypes = [C.c_char_p]
    L.Pt_get.restype = C.c_void_p
    L.Pt_get.argtypes = [C.c_uint32, C.c_size_t, C.c_uint32, C.c_uint32]
    assert L.Pt_mgr_init(pt_path.encode())
    raw = open(pt_path, "rb").read()
    for i in range(N_ENT):
        pt = L.Pt_get(i, MSG_LEN, 1, 0)
        # PLAINTEXT: POLYNOMIAL {u32 degree; size_t alloc, nq, np; bool ntt; int64* data}; slots; sf; sfd
        degree, alloc, nq, np_, ntt, data = struct.unpack("<I4xQQQ?7xQ", C.string_at(pt, 48))
        assert (degree, alloc, nq, np_) == (4096, 8, 8, 0) and data == pt + 72
        limbs = np.frombuffer(C.string_at(data, 8 * 4096 * 8), np.int64)
        ofs = PAGE + i * ((16 + 72 + 8 * 4096 * 8 + PAGE - 1) // PAGE * PAGE)
        assert (limbs == np.frombuffer(raw[ofs + 88: ofs + 88 + 8 * 4096 * 8], np.int64)).all()
    L.Pt_mgr_fini()
    print("LAYOUT OK")


if __name__ == "__main__":
    if sys.argv[1] == "make":
        print(*make_inputs(sys.argv[2]))
    else:
        layout(sys.argv[2])
