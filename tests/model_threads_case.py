"""Several images in flight on one GPU: S host threads call the reference's driver API
(Prepare_input / Run_main_graph / Handle_output, fhe-cmplr/rtlib/include/common/rt_api.h) on the
same prepared context at the same time, the way the reference's OpenMP loop over images does
(ant/dataset/resnet_cifar.main.inc:81).  Every thread gets its own worker context inside the
runtime (stream, limb allocator, scheduler); tables and keys are shared.  Each result must match
the reference's golden logits (tests/golden/<model>.json) like a single-threaded run does.

    python tests/model_threads_case.py resnet20_cifar10_pre [threads] [images per thread]
"""
import json
import os
import subprocess
import sys
import threading

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)


def main():
    model = sys.argv[1]
    S = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    per = int(sys.argv[3]) if len(sys.argv) > 3 else 2
    gold = json.load(open(os.path.join(HERE, "golden", model + ".json")))
    os.environ["RTLIB_BTS_EVEN_POLY"] = str(gold["even_poly"])
    msg = "/tmp/%s_amp%s.msg" % (model, gold["amp"])
    if not os.path.exists(msg):
        subprocess.run([sys.executable, os.path.join(ROOT, "tools", "make_weights.py"),
                        os.path.join(HERE, "emitted", model + ".entries.json"), msg,
                        "--amp", str(gold["amp"]), "--seed", str(gold["seed"])], check=True)
    from ace_compiler_b200.model_runner import EmittedModel, synthetic_image
    m = EmittedModel(model, msg)
    image = synthetic_image(0)
    want = np.array(gold["logits"])
    errs, fails = [[] for _ in range(S)], []
    start = threading.Barrier(S)

    def body(t):
        try:
            start.wait()
            for _ in range(per):
                m.prepare_input(image)
                m.run()
                errs[t].append(float(np.abs(m.handle_output(len(want)) - want).max()))
        except Exception as e:  # noqa: BLE001
            fails.append((t, repr(e)))

    threads = [threading.Thread(target=body, args=(t,)) for t in range(1, S)]
    for th in threads:
        th.start()
    body(0)
    for th in threads:
        th.join()
    m.close()
    assert not fails, fails
    print("max |logit - reference| per thread:", [max(e) for e in errs])
    assert all(len(e) == per and max(e) < 1e-5 for e in errs), errs
    print("THREADS PARITY OK (%d threads x %d images)" % (S, per))


if __name__ == "__main__":
    main()
