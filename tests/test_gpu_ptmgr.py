"""Pre-encoded weights: a DE_PLAINTEXT data file -- entries made by the REFERENCE's
Encode_plain_buffer (plain_eval.c:98-124) on the host, laid out as RT_DATA_WRITER does
(fhe-cmplr/include/fhe/core/rt_data_writer.h:28-106; tests/ptmgr_case.py) -- read through Pt_get on
the B200 runtime by tests/emitted/pt_get_case.c (an emitted-style unit:
`dest = *(PLAIN)Pt_get(...)`).  Bar: every plaintext limb identical to the run-time encode of the
same message (bit-exact), ring recycling under deferred execution, decrypted result within 1e-6."""
import os
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
EXE = os.path.join(ROOT, "tests", "_emitted_bin", "pt_get_case")
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libace_ref.so")


def _case(what, tmp):
    r = subprocess.run([sys.executable, os.path.join(HERE, "ptmgr_case.py"), what, tmp],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
    return r.stdout.strip().splitlines()[-1]


@pytest.mark.gpu
@pytest.mark.parametrize("slots,eager", [(2, "0"), (8, "0"), (2, "1")])
def test_pt_get(tmp_path, slots, eager):
    if not os.path.exists(EXE) or not os.path.exists(REF_SO):
        pytest.skip("pt_get_case / compiled reference not built (need /root/reference at build time)")
    msg_path, pt_path = _case("make", str(tmp_path)).split()
    env = dict(os.environ, ACE_B200_DATA_FILE=pt_path, PT_CASE_MSGS=msg_path, PT_ENTRY_COUNT=str(slots),
               ACE_B200_EAGER=eager, ACE_B200_SEED="5")
    r = subprocess.run([EXE], capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0 and "SUCESS!" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


@pytest.mark.gpu
@pytest.mark.parametrize("damage,message", [
    ("magic", "Plaintext buffer magic mismatch"),      # plain_eval.c:133-136
    ("size", "Plaintext size mismatch"),               # plain_eval.c:151-154
    ("degree", "Plaintext does not fit the context"),  # (a file made for another parameter set)
    ("lut", "look-up table outside the file"),         # truncated file
])
def test_pt_get_rejects_damaged_files(tmp_path, damage, message):
    """untrusted input: a damaged entry or table aborts with the reference's message (FMT_ASSERT ->
    abort, RT/include/common/error.h:23-29) instead of addressing memory through it"""
    if not os.path.exists(EXE) or not os.path.exists(REF_SO):
        pytest.skip("pt_get_case / compiled reference not built (need /root/reference at build time)")
    msg_path, pt_path = _case("make", str(tmp_path)).split()
    raw = bytearray(open(pt_path, "rb").read())
    ent = 4096  # first entry: PLAINTEXT_BUFFER {magic[8], version, size} + PLAINTEXT {u32 degree; ...}
    if damage == "magic":
        raw[ent:ent + 8] = b"NOTPLAIN"
    elif damage == "size":
        raw[ent + 12:ent + 16] = (int.from_bytes(raw[ent + 12:ent + 16], "little") - 8).to_bytes(4, "little")
    elif damage == "degree":
        raw[ent + 16:ent + 20] = (8192).to_bytes(4, "little")
        # keep the buffer self-consistent: size = sizeof(PLAINTEXT) + 8 * alloc * degree would no
        # longer match, so shrink the number of primes instead (8 limbs of 4096 = 4 limbs of 8192)
        raw[ent + 24:ent + 32] = (4).to_bytes(8, "little")
        raw[ent + 32:ent + 40] = (4).to_bytes(8, "little")
    else:
        raw = raw[:len(raw) - 64]
    open(pt_path, "wb").write(bytes(raw))
    env = dict(os.environ, ACE_B200_DATA_FILE=pt_path, PT_CASE_MSGS=msg_path, ACE_B200_SEED="5")
    r = subprocess.run([EXE], capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode != 0 and "SUCESS!" not in r.stdout
    assert message in r.stdout + r.stderr, (r.stdout[-1000:], r.stderr[-1000:])


@pytest.mark.skipif(not os.path.exists(REF_SO), reason="compiled reference not present")
def test_plaintext_file_layout(tmp_path):
    """CPU: the file the GPU test reads is one the REFERENCE's reader accepts -- Pt_mgr_init +
    Pt_get of the compiled reference (pt_mgr.c:35-159) hand back the buffers that went in"""
    assert _case("layout", str(tmp_path)) == "LAYOUT OK"
