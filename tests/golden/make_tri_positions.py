"""Positions of the reference's Sample_triangle draws in its (pinned, mode 0) rand() stream.

    python tests/golden/make_tri_positions.py resnet20_cifar10_pre

The golden runs (make_model_golden.py) were made in the harness's pin mode 0: srand() swallowed,
rand() running on from srandom(12345).  Where a Sample_triangle starts in that stream depends on
every rand() call before it (Is_prime's trials during set-up, number_theory.c:160-185).  This
script repeats the golden run's Prepare_context + Prepare_input on the compiled reference
(oracle/_ref, ~6 min, ~40 GB; no Main_graph), checks that the input ciphertext is the golden
run's, and stores the stream position of every Sample_triangle call in
tests/golden/<model>.json ("tri_positions": start, count, and the calls that do not follow
their predecessor by exactly N draws).  With them tests/model_case.py lets the B200 runtime
generate the golden run's keys itself (ace_keygen_reference_stream) instead of waiting minutes
for the reference's key generation on the GPU box.  TEST INFRASTRUCTURE."""
import ctypes as C
import hashlib
import json
import os
import subprocess
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, ROOT)


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    model = sys.argv[1]
    path = os.path.join(HERE, model + ".json")
    gold = json.load(open(path))
    os.environ["RTLIB_BTS_EVEN_POLY"] = str(gold["even_poly"])
    msg = "/tmp/%s_amp%s.msg" % (model, gold["amp"])
    if not os.path.exists(msg):
        subprocess.run([sys.executable, os.path.join(ROOT, "tools", "make_weights.py"),
                        os.path.join(ROOT, "tests", "emitted", model + ".entries.json"), msg,
                        "--amp", str(gold["amp"]), "--seed", str(gold["seed"])], check=True)
    from make_model_golden import synthetic_image
    from oracle_bindings import RefModel, build_oracles
    build_oracles()
    t = time.time()
    ref = RefModel(model, msg)
    print("reference Prepare_context %.0f s" % (time.time() - t), flush=True)
    ref.prepare_input(synthetic_image(0))
    cin = ref.peek_input()
    assert sha(cin.c0) == gold["input"]["sha256_c0"] and sha(cin.c1) == gold["input"]["sha256_c1"], \
        "not the golden run's input ciphertext"
    ref.lib.ref_triangle_positions.argtypes = [C.c_void_p, C.c_uint32]
    buf = (C.c_uint64 * 8192)()
    n = ref.lib.ref_triangle_positions(buf, 8192)
    assert 0 < n < 8192
    pos = [int(buf[i]) for i in range(n)]
    N = ref.N
    breaks = [[k, pos[k]] for k in range(1, n) if pos[k] != pos[k - 1] + N]
    gold["tri_positions"] = {"srandom": 12345, "N": N, "start": pos[0], "count": n, "breaks": breaks}
    json.dump(gold, open(path, "w"), indent=1)
    print("wrote", path, gold["tri_positions"]["start"], n, breaks[:8])


if __name__ == "__main__":
    sys.path.insert(0, HERE)
    main()
