"""Calibration of bench.py's composed CPU baseline: a REAL end-to-end run of the emitted unit on the
unmodified reference rtlib (oracle/_ref), on the host it is started on, next to the composed
estimate (unit costs x op trace) taken on the same host.

    python tools/reference_real_run.py [model] [--out profiles/r2_reference_real_run.json]

Prepare_context (key generation, minutes) + one Main_graph (tens of minutes, one thread, ~40 GB of
host RAM).  Appends {host, cpus, real_main_graph_s, composed_s, composed_over_real} to the JSON
file.  TEST / MEASUREMENT INFRASTRUCTURE: loads oracle/_ref, never the product library."""
import argparse
import json
import os
import platform
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("model", nargs="?", default="resnet20_cifar10_pre")
    ap.add_argument("--out", default=os.path.join(ROOT, "profiles", "r2_reference_real_run.json"))
    a = ap.parse_args()
    import bench
    bench.select_model(a.model)
    os.environ.setdefault("RTLIB_BTS_EVEN_POLY", "1")
    from oracle_bindings import RefModel
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    from make_model_golden import synthetic_image
    msg = bench.weight_file(a.model)
    t = time.time()
    m = RefModel(a.model, msg)
    t_ctx = time.time() - t
    print("Prepare_context %.1f s" % t_ctx, flush=True)
    m.prepare_input(synthetic_image(0))
    t = time.time()
    m.run()
    t_run = time.time() - t
    print("Main_graph %.1f s" % t_run, flush=True)
    logits = m.handle_output(bench.CLASSES)
    print("logits", [float(x) for x in logits[:4]], flush=True)
    del m
    # the composed estimate on the same host, in a fresh process (the reference owns one global context)
    import subprocess
    r = subprocess.run([sys.executable, "-c",
                        "import json,sys; sys.path.insert(0,%r); import bench; bench.select_model(%r); "
                        "b=bench.cpu_baseline(bench.load_trace(%r),1); print('COMPOSED',json.dumps(b['s_per_image_1thread']))"
                        % (ROOT, a.model, a.model)], capture_output=True, text=True)
    composed = None
    for line in r.stdout.splitlines():
        if line.startswith("COMPOSED"):
            composed = float(line.split()[1])
    rec = {"model": a.model, "host": platform.node(), "cpu": platform.processor() or platform.machine(),
           "cpus": os.cpu_count(), "prepare_context_s": round(t_ctx, 1),
           "real_main_graph_s": round(t_run, 1), "composed_s": composed,
           "composed_over_real": round(composed / t_run, 4) if composed else None,
           "when": time.strftime("%Y-%m-%dT%H:%M:%SZ", time.gmtime())}
    try:
        runs = json.load(open(a.out))
    except Exception:
        runs = []
    runs.append(rec)
    json.dump(runs, open(a.out, "w"), indent=1)
    print("wrote", a.out, rec)


if __name__ == "__main__":
    main()
