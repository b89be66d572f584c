"""Whole-model parity of an ACE-emitted ResNet on the B200 runtime (run in its own process).

    python tests/model_case.py resnet20_cifar10_pre [exact]

default : the GPU runtime generates its own keys, runs the emitted Main_graph on the synthetic
          image / weight file and must decrypt to the logits the reference produced
          (tests/golden/<model>.json) within 1e-5 -- different keys, same computation.
exact   : oracle/_ref/libace_ref.so regenerates the golden run's keys and input ciphertext (pinned
          randomness, ~3-6 min of host time, ~40 GB of host RAM), the GPU runtime imports them and
          must reproduce the golden OUTPUT CIPHERTEXT bit for bit (SHA-256 of its limbs, level,
          scale) and hence identical decrypted logits.
TEST INFRASTRUCTURE: the product never loads anything under oracle/."""
import ctypes as C
import hashlib
import json
import os
import subprocess
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    model = sys.argv[1]
    exact = len(sys.argv) > 2 and sys.argv[2] == "exact"
    gold = json.load(open(os.path.join(HERE, "golden", model + ".json")))
    os.environ["RTLIB_BTS_EVEN_POLY"] = str(gold["even_poly"])
    msg = "/tmp/%s_amp%s.msg" % (model, gold["amp"])
    if not os.path.exists(msg):
        subprocess.run([sys.executable, os.path.join(ROOT, "tools", "make_weights.py"),
                        os.path.join(HERE, "emitted", model + ".entries.json"), msg,
                        "--amp", str(gold["amp"]), "--seed", str(gold["seed"])], check=True)
    from ace_compiler_b200.model_runner import EmittedModel, synthetic_image
    image = synthetic_image(0)
    n_cls = len(gold["logits"])
    if not exact:
        m = EmittedModel(model, msg)
        m.prepare_input(image)
        t = time.time()
        m.run()
        print("Main_graph %.2f s on the GPU (reference: %.0f s on %d host cpus)" %
              (time.time() - t, gold["main_graph_s"], gold["host_cpus"]))
        logits = m.handle_output(n_cls)
        err = np.abs(logits - np.array(gold["logits"])).max()
        print("logits", logits, "max |diff| vs reference %.3e" % err)
        assert err < 1e-5, err
        m.close()
        print("MODEL PARITY OK (decrypted logits)")
        return

    from oracle_bindings import RefModel, build_oracles
    build_oracles()
    t = time.time()
    ref = RefModel(model, msg)
    print("reference Prepare_context %.0f s" % (time.time() - t), flush=True)
    ref.prepare_input(image)
    cin = ref.peek_input()
    assert cin.level == gold["input"]["level"]
    assert sha(cin.c0) == gold["input"]["sha256_c0"] and sha(cin.c1) == gold["input"]["sha256_c1"], \
        "the reference did not reproduce the golden run's input ciphertext (randomness not pinned?)"
    m = EmittedModel(model, msg, own_keys=False)
    L = m.lib

    class Params(C.Structure):
        _fields_ = [("provider", C.c_int), ("degree", C.c_uint32), ("sec", C.c_size_t),
                    ("depth", C.c_size_t), ("q0", C.c_size_t), ("sf", C.c_size_t),
                    ("parts", C.c_size_t), ("hw", C.c_size_t), ("n_rot", C.c_size_t)]
    L.Get_context_params.restype = C.POINTER(Params)
    p = L.Get_context_params()
    n_rot = p.contents.n_rot
    rots_p = C.cast(C.addressof(p.contents) + C.sizeof(Params), C.POINTER(C.c_int32))
    rots = [int(rots_p[i]) for i in range(n_rot)]
    buf = (C.c_int32 * 4096)()
    nb = L.Ace_bootstrap_rot_indices(0, buf, 4096)
    bts = [int(buf[i]) for i in range(nb)]
    N = ref.N
    L.Ace_import_switch_key.argtypes = [C.c_bool, C.c_int32, C.c_uint32, C.c_int, C.c_void_p]
    t = time.time()

    def imp(is_rot, rot, k0, k1):
        for part in range(k0.shape[0]):
            L.Ace_import_switch_key(is_rot, rot, part, 0, k0[part].ctypes.data)
            L.Ace_import_switch_key(is_rot, rot, part, 1, k1[part].ctypes.data)
    seen = set()
    for r in rots + bts:
        if r in seen:
            continue
        seen.add(r)
        imp(True, r, *ref.swk(True, r))
    imp(True, 2 * N - 1, *ref.swk_auto(2 * N - 1))
    imp(False, 0, *ref.swk(False, 0))
    print("%d switch keys imported in %.0f s" % (len(seen) + 2, time.time() - t), flush=True)
    L.Ace_set_input.argtypes = [C.c_char_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_uint32,
                                C.c_uint32, C.c_double, C.c_uint32]
    L.Ace_set_input(b"input", 0, cin.c0.ctypes.data, cin.c1.ctypes.data, cin.level, cin.slots,
                    cin.scale, cin.sf_degree)
    t = time.time()
    m.run()
    print("Main_graph %.2f s on the GPU (reference: %.0f s)" % (time.time() - t, gold["main_graph_s"]),
          flush=True)

    class Poly(C.Structure):
        _fields_ = [("degree", C.c_uint32), ("alloc", C.c_size_t), ("nq", C.c_size_t),
                    ("np", C.c_size_t), ("ntt", C.c_bool), ("data", C.c_void_p)]

    class Ct(C.Structure):
        _fields_ = [("c0", Poly), ("c1", Poly), ("slots", C.c_uint32), ("scale", C.c_double),
                    ("sfd", C.c_uint32)]
    L.Ace_get_output.restype = C.POINTER(Ct)
    L.Ace_get_output.argtypes = [C.c_char_p, C.c_size_t]
    L.Ace_download_poly.argtypes = [C.c_void_p, C.c_void_p]
    out = L.Ace_get_output(b"output", 0).contents
    lvl = out.c0.nq
    c0, c1 = np.zeros((lvl, N), np.int64), np.zeros((lvl, N), np.int64)
    L.Ace_download_poly(c0.ctypes.data, C.byref(out.c0))
    L.Ace_download_poly(c1.ctypes.data, C.byref(out.c1))
    g = gold["output"]
    assert (lvl, out.slots, out.sfd, out.scale) == (g["level"], g["slots"], g["sf_degree"], g["scale"]), \
        (lvl, out.slots, out.sfd, out.scale, g)
    sums0 = [int(x) for x in c0.view(np.uint64).sum(axis=1, dtype=np.uint64)]
    assert sums0 == g["limb_sums_c0"], (sums0, g["limb_sums_c0"])
    assert sha(c0) == g["sha256_c0"] and sha(c1) == g["sha256_c1"], "output ciphertext differs"
    from oracle_bindings import Ct as HostCt
    dec = ref.decrypt(HostCt(c0, c1, out.slots, out.sfd, out.scale))[:n_cls]
    assert (dec == np.array(gold["logits"])).all(), (dec, gold["logits"])
    m.close()
    print("MODEL PARITY OK (output ciphertext bit-exact, logits identical)")


if __name__ == "__main__":
    main()
