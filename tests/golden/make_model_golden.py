"""Golden run of an ACE-emitted ResNet on the UNMODIFIED reference rtlib (CPU, this container).

    python tests/golden/make_model_golden.py resnet20_cifar10_pre [--amp 0.05]

Runs oracle/_ref/<model>_ref.so (the emitted .onnx.inc bound to oracle/_ref/libace_ref.so) with
pinned randomness on the synthetic image / weight file of SURVEY.md 8(d) config 1 and writes
tests/golden/<model>.json: decrypted logits, level/scale of the output ciphertext, SHA-256 of its
limbs and of the input ciphertext's, and per-limb 64-bit sums for quick diffing.  The GPU test
(tests/test_gpu_model.py) regenerates the same keys with the same library and seeds on the GPU
box, runs the same emitted unit on the B200 runtime and must reproduce these values.
Takes tens of minutes and ~45 GB of RAM (227 switch keys + bootstrap tables on the host)."""
import argparse
import ctypes as C
import hashlib
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, ROOT)


def synthetic_image(idx=0):
    x = np.uint32(12345 + idx)
    out = np.zeros(3 * 32 * 32)
    with np.errstate(over="ignore"):
        for i in range(out.size):
            x = np.uint32(x * np.uint32(1664525) + np.uint32(1013904223))
            out[i] = float(int(x) >> 8) / 16777216.0 - 0.5
    return out


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("model")
    ap.add_argument("--amp", type=float, default=0.05)
    ap.add_argument("--classes", type=int, default=10)
    a = ap.parse_args()
    from oracle_bindings import RefModel
    os.environ.setdefault("RTLIB_BTS_EVEN_POLY", "1")
    msg = "/tmp/%s.msg" % a.model
    import subprocess
    subprocess.run([sys.executable, os.path.join(ROOT, "tools", "make_weights.py"),
                    os.path.join(ROOT, "tests", "emitted", a.model + ".entries.json"), msg,
                    "--amp", str(a.amp)], check=True)
    t = time.time()
    m = RefModel(a.model, msg)
    t_ctx = time.time() - t
    print("Prepare_context %.1f s" % t_ctx, flush=True)
    m.prepare_input(synthetic_image(0))
    cin = m.peek_input()
    t = time.time()
    m.run()
    t_run = time.time() - t
    print("Main_graph %.1f s" % t_run, flush=True)
    cout = m.peek_output()
    logits = m.handle_output(a.classes)
    rec = {
        "model": a.model, "amp": a.amp, "seed": 1, "even_poly": os.environ["RTLIB_BTS_EVEN_POLY"],
        "prepare_context_s": t_ctx, "main_graph_s": t_run, "host_cpus": os.cpu_count(),
        "input": {"level": cin.level, "sha256_c0": sha(cin.c0), "sha256_c1": sha(cin.c1)},
        "output": {"level": cout.level, "slots": cout.slots, "sf_degree": cout.sf_degree,
                   "scale": cout.scale, "sha256_c0": sha(cout.c0), "sha256_c1": sha(cout.c1),
                   "limb_sums_c0": [int(x) for x in cout.c0.view(np.uint64).sum(axis=1, dtype=np.uint64)],
                   "limb_sums_c1": [int(x) for x in cout.c1.view(np.uint64).sum(axis=1, dtype=np.uint64)]},
        "logits": [float(x) for x in logits],
    }
    path = os.path.join(HERE, a.model + ".json")
    json.dump(rec, open(path, "w"), indent=1)
    print("wrote", path, rec["logits"])


if __name__ == "__main__":
    main()
