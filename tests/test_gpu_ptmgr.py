"""Pre-encoded weights: a DE_PLAINTEXT data file -- entries made by the REFERENCE's
Encode_plain_buffer (plain_eval.c:98-124) on the host, laid out as RT_DATA_WRITER does
(fhe-cmplr/include/fhe/core/rt_data_writer.h:28-106) -- read through Pt_get on the B200 runtime
by tests/emitted/pt_get_case.c (an emitted-style unit: `dest = *(PLAIN)Pt_get(...)`).
Bar: every plaintext limb identical to the run-time encode of the same message (bit-exact),
ring recycling under deferred execution, decrypted result within 1e-6."""
import ctypes as C
import os
import struct
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tests", "_emitted_bin", "pt_get_case")
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


@pytest.mark.gpu
@pytest.mark.parametrize("slots,eager", [(2, "0"), (8, "0"), (2, "1")])
def test_pt_get(tmp_path, slots, eager):
    if not os.path.exists(EXE) or not os.path.exists(REF_SO):
        pytest.skip("pt_get_case / compiled reference not built (need /root/reference at build time)")
    msg_path, pt_path = make_inputs(str(tmp_path))
    env = dict(os.environ, ACE_B200_DATA_FILE=pt_path, PT_CASE_MSGS=msg_path, PT_ENTRY_COUNT=str(slots),
               ACE_B200_EAGER=eager, ACE_B200_SEED="5")
    r = subprocess.run([EXE], capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0 and "SUCESS!" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


@pytest.mark.skipif(not os.path.exists(REF_SO), reason="compiled reference not present")
def test_plaintext_file_layout(tmp_path):
    """CPU: the file the GPU test reads is one the REFERENCE's reader accepts -- Pt_mgr_init +
    Pt_get of the compiled reference (pt_mgr.c:35-159) hand back the buffers that went in"""
    msg_path, pt_path = make_inputs(str(tmp_path))
    L = C.CDLL(REF_SO)
    L.Pt_mgr_init.restype = C.c_bool
    L.Pt_mgr_init.argtypes = [C.c_char_p]
    L.Pt_get.restype = C.c_void_p
    L.Pt_get.argtypes = [C.c_uint32, C.c_size_t, C.c_uint32, C.c_uint32]
    assert L.Pt_mgr_init(pt_path.encode())
    msgs = np.fromfile(msg_path, np.float32).reshape(N_ENT, MSG_LEN)
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
    assert msgs.shape == (N_ENT, MSG_LEN)
