#!/usr/bin/env python
"""bench.py -- ResNet-20 / CIFAR-10 encrypted inference (BASELINE.json headline metric).

Workload (BASELINE.json configs[0]): the reference's checked-in,
unmodified ACE output `resnet20_cifar10_pre.onnx.inc` (N = 2^16, 34 Q limbs of 51/50 bit, 11 P
limbs of 60 bit, dnum = 3, 19 bootstraps) compiled against this repo's rt_ant header tree and
executed by the B200 runtime; synthetic weight file and image (no network): tools/make_weights.py
and ace_compiler_b200.model_runner.synthetic_image.  One *step* = one encrypted image through
Main_graph on every GPU (images are independent -> one image per rank, no collective).

  value   : images/s, whole job; timed region = Run_main_graph() with the encrypted input already
            resident in HBM (the reference's RTM_MAIN_GRAPH region), CUDA events, max over ranks
  e2e     : images/s through the reference's driver API with HOST data: Prepare_input (host image
            -> encode + encrypt on the GPU) + Run_main_graph + Handle_output (decrypt, decode,
            logits back on the host) all inside the timed region
  roofline: batched forward NTT (the dominant kernel family), timed live
  cpu_baseline / --impl reference: the unmodified reference rtlib (oracle/_ref/libace_ref.so)
            on the host cores.  One reference image takes ~20 min and its context ~6 min to
            build, so the CPU figure is composed: the reference's unit cost of every primitive
            the image consists of (Decomp_modup, Mod_down, Rescale, limb mul/add/rotate, NTT,
            encode; measured live at several levels, a bounded sample) times the model's op trace
            (tests/emitted/<model>.trace.json, recorded by the GPU runtime; control flow of an
            emitted program is data independent).  DESIGN.md section 5 compares this estimate
            with a real end-to-end reference run.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# BASELINE.json configs 1 and 3-5: the reference's checked-in emitted ResNets (all N = 2^16, 34 Q
# limbs, 11 P limbs, dnum = 3).  The default (headline) is ResNet-20; --model selects the others.
MODELS = {
    "resnet20_cifar10_pre": dict(sf=50, classes=10, bootstraps=19, title="ResNet-20 CIFAR-10"),
    "resnet32_cifar100_pre": dict(sf=50, classes=100, bootstraps=31, title="ResNet-32 CIFAR-100"),
    "resnet56_cifar10_pre": dict(sf=50, classes=10, bootstraps=55, title="ResNet-56 CIFAR-10"),
    "resnet110_cifar10_train": dict(sf=48, classes=10, bootstraps=109, title="ResNet-110 CIFAR-10"),
}
MODEL = "resnet20_cifar10_pre"
N, DEPTH, Q0, SF, PARTS, HW = 65536, 33, 51, 50, 3, 192
METRIC = "resnet20_cifar10_encrypted_inference_throughput"
UNIT = "images/s"
WORKLOAD = ""
CLASSES = 10


def select_model(name):
    global MODEL, SF, METRIC, WORKLOAD, CLASSES
    cfg = MODELS[name]
    MODEL, SF, CLASSES = name, cfg["sf"], cfg["classes"]
    METRIC = name.replace("_pre", "").replace("_train", "") + "_encrypted_inference_throughput"
    WORKLOAD = ("%s single-image encrypted inference: ACE-emitted %s.onnx.inc (N=2^16, L=34, K=11, "
                "dnum=3, Delta=2^%d, %d bootstraps), synthetic weights/image"
                % (cfg["title"], name, cfg["sf"], cfg["bootstraps"]))


select_model(MODEL)
TRACE_CLASSES = ["modup_digit", "moddown_poly", "rescale_poly", "encode", "limb_mul", "limb_add",
                 "limb_rot", "limb_ntt"]
TRACE_LEVELS = 72
# images in flight per GPU: 3 measured best that fits comfortably (1: 0.93, 2: 1.15, 3: 1.21
# images/s on one B200); every image in flight has its own stream, allocator cache and deferred frees
DEFAULT_STREAMS = 3
# dram__bytes_read.sum + dram__bytes_write.sum of one 45-limb forward NTT: taken from the ncu
# --set full capture summarised in this file (written by tools/ncu_traffic.py from the .ncu-rep);
# null when no capture of the current kernels has been committed
NTT_TRAFFIC_FILE = os.path.join(ROOT, "profiles", "r2_ncu_ntt_traffic.json")


def read_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


def weight_file(model, amp=0.05):
    path = "/tmp/ace_b200_%s%s.msg" % (model, "" if amp == 0.05 else "_amp%g" % amp)
    if not os.path.exists(path):
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        import make_weights
        ent = [tuple(e) for e in json.load(open(os.path.join(ROOT, "tests", "emitted",
                                                             model + ".entries.json")))]
        make_weights.write_file(path + ".tmp%d" % os.getpid(), ent, amp, 1)
        os.replace(path + ".tmp%d" % os.getpid(), path)
    return path


class ClockSampler(threading.Thread):
    """samples SM clocks and throttle reasons while the timed region runs"""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag = index, [], set(), False
        self.max_mhz = None

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                      "--format=csv,noheader,nounits"], capture_output=True,
                                     text=True, timeout=5).stdout.strip().split(",")
                self.samples.append(float(out[0]))
                self.max_mhz = float(out[1])
                for n, v in zip(names, out[2:]):
                    if v.strip().lower() == "active":
                        self.reasons.add(n)
            except Exception:
                pass
            time.sleep(0.5)

    def summary(self):
        return {"sm_mhz": float(np.median(self.samples)) if self.samples else None,
                "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


def dist_setup():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist_
        torch.cuda.set_device(local)
        dist_.init_process_group("nccl", device_id=torch.device("cuda", local))
        dist = dist_
    return rank, world, local, dist


def barrier(dist, local):
    if dist is not None:
        import torch
        dist.barrier(device_ids=[local])
        torch.cuda.synchronize(local)


def max_over_ranks(dist, local, x):
    if dist is None:
        return x
    import torch
    t = torch.tensor([x], dtype=torch.float64, device="cuda:%d" % local)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def read_trace(model_obj):
    buf = (C.c_uint64 * (len(TRACE_CLASSES) * TRACE_LEVELS))()
    n = model_obj.lib.Ace_trace(buf, len(buf))
    assert n == len(buf)
    return np.array(buf, dtype=np.uint64).reshape(len(TRACE_CLASSES), TRACE_LEVELS)


def trace_to_json(t):
    return {c: {str(l): int(t[i, l]) for l in range(TRACE_LEVELS) if t[i, l]}
            for i, c in enumerate(TRACE_CLASSES)}


def load_trace(model):
    with open(os.path.join(ROOT, "tests", "emitted", model + ".trace.json")) as f:
        return json.load(f)["per_image"]


# ----------------------------------------------------------------------- kernel rooflines
def kernel_rooflines(device):
    """Device timings (CUDA events on the context stream, inputs resident) of the primitives of
    BASELINE.json config 2 at the model's parameter set, each against the roofline that bounds it:
    algorithmic bytes / time against the measured HBM peak, and for the
    integer-bound ones butterflies/s against the measured arithmetic ceiling of the radix-16
    register pass (ace_ntt_bfly_peak) resp. 64x64-bit MACs/s against the measured IMAD.WIDE rate
    (ace_measure_pipe_peaks).  Returns (roofline of the dominant kernel, table)."""
    import ace_compiler_b200 as ace
    ctx = ace.Context(N, DEPTH, Q0, SF, PARTS, device=device)
    lib, h = ctx.lib, ctx.h
    L, K, G = ctx.L, ctx.K, ctx.L + ctx.K
    NB, LIMB = N * 8, N * 8.0
    peak, how = read_peaks()
    rng = np.random.default_rng(7)
    mods = np.concatenate([ctx.q, ctx.p])

    def rand(gs):
        return np.stack([rng.integers(0, mods[g], N, dtype=np.int64) for g in gs])

    for is_rot, rot in [(False, 0), (True, 1)]:
        k0 = np.stack([rand(range(G)) for _ in range(PARTS)])
        k1 = np.stack([rand(range(G)) for _ in range(PARTS)])
        ctx.import_switch_key(is_rot, rot, k0, k1)

    def timeit(fn, reps=20, warm=3):
        for _ in range(warm):
            fn()
        ctx.sync()
        lib.ace_timer_start(h)
        for _ in range(reps):
            fn()
        ms = C.c_float()
        lib.ace_timer_stop_ms(h, C.byref(ms))
        return ms.value / reps * 1e3  # us

    pipes = (C.c_double * 7)()
    lib.ace_measure_pipe_peaks(device, pipes, 7)
    imad_wide, imad, dfma = pipes[0], pipes[1], pipes[3]
    ceil = {f: lib.ace_ntt_bfly_peak(h, f, 3) for f in (0, 1, 2)}  # G butterflies/s: fp64, int, int+csub
    n_fp64 = sum(1 for g in range(L) if ceil[0] > 0 and mods[g] < 1500000000000000)
    bfly_per_limb = N // 2 * 16

    def bfly_ceiling_s(n_q, n_p):  # seconds the butterflies alone would take at the ceilings
        fq = min(n_fp64, max(0, n_q - 1)) if n_q else 0  # q_0 is a 51-bit prime: integer form
        return bfly_per_limb * (fq / (ceil[0] * 1e9) if fq else 0.0) + bfly_per_limb * (
            (n_q - fq) / (ceil[1] * 1e9) + n_p / (ceil[2] * 1e9))

    # three copies of every operand, used round robin: 3 x 47 MB of limbs + tables do not stay in L2
    full = ctx.put(rand(list(range(G)) * 3))
    rows = []

    def add(name, us, alg_limbs, bound, extra=None):
        gbs = alg_limbs * LIMB / (us * 1e-6) / 1e9
        row = {"name": name, "us": round(us, 2), "alg_MiB": round(alg_limbs * LIMB / 2**20, 1),
               "GBps": round(gbs, 1), "hbm_frac": round(gbs / peak, 4), "bound": bound}
        if extra:
            row.update(extra)
        rows.append(row)
        return row

    cnt = [0]

    def rr():
        cnt[0] += 1
        return full.ptr + (cnt[0] % 3) * G * NB

    def int_ntt(us, nq, np_):
        return {"butterflies_per_s_G": round((nq + np_) * bfly_per_limb / us / 1e3, 1),
                "arith_ceiling_frac": round(bfly_ceiling_s(nq, np_) / (us * 1e-6), 4)}

    us = timeit(lambda: lib.ace_ntt(h, rr(), 0, G), reps=30, warm=6)
    ntt_row = add("ntt x%d" % G, us, 2 * G, "int/fp64 pipes", int_ntt(us, L, K))
    us = timeit(lambda: lib.ace_intt(h, rr(), 0, G), reps=30, warm=6)
    add("intt x%d" % G, us, 2 * G, "int/fp64 pipes", int_ntt(us, L, K))
    us = timeit(lambda: lib.ace_ntt(h, rr(), 0, L))
    add("ntt x%d (Q)" % L, us, 2 * L, "int/fp64 pipes", int_ntt(us, L, 0))
    us = timeit(lambda: lib.ace_ntt(h, rr() + L * NB, L, K))
    add("ntt x%d (P)" % K, us, 2 * K, "int pipe", int_ntt(us, 0, K))
    us = timeit(lambda: lib.ace_ntt(h, rr(), 0, 1))
    add("ntt x1", us, 2, "latency", int_ntt(us, 1, 0))
    for lv in (L, (L + 1) // 2):
        beta = -(-lv // ctx.part_size)
        a = ctx.put(rand(list(range(lv)) * 2))
        b = ctx.put(rand(list(range(lv)) * 2))
        o = ctx.empty(2 * lv)
        ext = ctx.empty(lv + K, zero=True)
        e2 = ctx.put(rand(list(range(lv)) + [L + i for i in range(K)]))
        a0, a1, b0, b1 = a.ptr, a.ptr + lv * NB, b.ptr, b.ptr + lv * NB
        o0, o1 = o.ptr, o.ptr + lv * NB
        tag = " L=%d" % lv
        add("limb_mul x%d%s" % (lv, tag), timeit(lambda: lib.ace_hw_modmul(h, o0, a0, b0, 0, lv)), 3 * lv, "hbm")
        add("limb_add x%d%s" % (lv, tag), timeit(lambda: lib.ace_hw_modadd(h, o0, a0, b0, 0, lv)), 3 * lv, "hbm")
        alpha = min(ctx.part_size, lv)
        us = timeit(lambda: lib.ace_decomp_modup(h, ext.ptr, a0, lv, 0))
        macs = N * alpha * (lv - alpha + K)
        add("modup_digit" + tag, us, alpha + lv + K, "int pipe",
            {"macs_per_s_G": round(macs / us / 1e3, 1), "transforms": alpha + lv - alpha + K})
        us = timeit(lambda: lib.ace_mod_down(h, o0, e2.ptr, lv))
        add("moddown_poly" + tag, us, lv + K + lv, "int pipe",
            {"macs_per_s_G": round(N * K * lv / us / 1e3, 1), "transforms": K + lv})
        add("rescale_poly" + tag, timeit(lambda: lib.ace_rescale(h, o0, a0, lv)), 2 * lv - 1, "int pipe",
            {"transforms": lv})
        ks_limbs = 4 * lv + 2 * beta * (lv + K)
        add("key_switch" + tag, timeit(lambda: lib.ace_key_switch(h, o0, o1, a1, lv, 0, 0)), ks_limbs,
            "composite", {"transforms": beta * (lv + K) + 2 * K + 2 * lv})
        add("ct_rotate" + tag, timeit(lambda: lib.ace_ct_rotate(h, o0, o1, a0, a1, lv, 1)), ks_limbs, "composite")
        add("ct_mul_relin" + tag, timeit(lambda: lib.ace_ct_mul_relin(h, o0, o1, a0, a1, b0, b1, lv)),
            ks_limbs + 7 * lv, "composite")
        add("ct_rescale" + tag, timeit(lambda: lib.ace_ct_rescale(h, o0, o1, a0, a1, lv)), 4 * lv - 2, "int pipe")
        vals = rng.uniform(-0.05, 0.05, N // 2)
        pt = ctx.empty(lv)
        add("encode" + tag, timeit(lambda: lib.ace_encode(h, pt.ptr, vals.ctypes.data_as(C.c_void_p), N // 2, lv,
                                                            N // 2, 1, 0), reps=5),
            lv + 0.25, "fp64 + h2d (host message)")
        for x in (a, b, o, ext, e2, pt):
            x.free()
    traffic = None
    try:
        traffic = json.load(open(NTT_TRAFFIC_FILE))["dram_bytes_per_launch"]
    except Exception:
        pass
    alg_bytes = G * N * 8 * 2
    roof = {"bound": "hbm", "kernel": "ntt16 forward, K1 cols + K2 rows (%d limbs/launch pair)" % G,
            "achieved": ntt_row["GBps"], "peak": peak, "peak_source": how, "unit": "GB/s",
            "frac": ntt_row["hbm_frac"], "traffic": traffic, "alg_bytes_per_launch": alg_bytes,
            "launch_us": ntt_row["us"],
            "arith": {"butterflies_per_s_G": ntt_row["butterflies_per_s_G"],
                      "ceiling_frac": ntt_row["arith_ceiling_frac"],
                      "ceiling_G_bfly_per_s": {"fp64": round(ceil[0], 1), "int64_lazy": round(ceil[1], 1),
                                               "int64_csub": round(ceil[2], 1)},
                      "limbs": {"fp64": n_fp64, "int64_lazy": L - n_fp64, "int64_csub": K},
                      "pipe_peaks_G_instr_per_s": {"imad_wide_u32": round(imad_wide, 1), "imad": round(imad, 1),
                                                   "dfma": round(dfma, 1)}},
            "note": "the transform is bound by the multiply pipes, not by HBM: frac (algorithmic "
                    "bytes against the measured HBM peak) is reported as the contract asks, "
                    "arith.ceiling_frac is the fraction of the measured butterfly rate of the "
                    "radix-16 register pass with no memory traffic (all measured in this run)"}
    full.free()
    ctx.close()
    return roof, rows


# --------------------------------------------------------------------------------- ours
def run_ours(args):
    from ace_compiler_b200.model_runner import EmittedModel, synthetic_image
    rank, world, local, dist = dist_setup()
    msg = weight_file(MODEL)
    t0 = time.time()
    m = EmittedModel(MODEL, msg, device=local)
    t_ctx = time.time() - t0
    import torch
    S = max(1, args.streams)
    images = [torch.from_numpy(synthetic_image(rank * 1000 + i)).pin_memory() for i in range(4 * S)]
    warm = max(3, args.warmup)

    def step_e2e(i):
        m.prepare_input(images[i % len(images)].numpy())
        m.run()
        return m.handle_output(CLASSES)

    t0 = time.time()
    logits = step_e2e(0)
    t_first = time.time() - t0
    # single-image latency on the primary thread alone (the s/image half of the metric)
    m.prepare_input(images[1].numpy())
    m.timer_start()
    m.run()
    ms_single = m.timer_stop_ms()
    m.handle_output(CLASSES)

    # ---- S images in flight per GPU: S host threads, each with its own stream / allocator /
    # scheduler inside the runtime (the reference's OpenMP-over-images driver); thread 0 is this one
    sync = threading.Barrier(S)
    res = [dict() for _ in range(S)]
    sampler = ClockSampler(local)

    def body(t):
        out = res[t]
        for i in range(warm - 1 if t == 0 else warm):
            out["logits"] = step_e2e(t * 4 + i)
        sync.wait()
        if t == 0:
            sampler.start()
            barrier(dist, local)
        sync.wait()
        # value: Main_graph with the input ciphertext resident, device-timed per step
        tr0, l0 = read_trace(m), m.launch_count()
        ms = 0.0
        for i in range(args.steps):
            m.prepare_input(images[(t * 4 + i) % len(images)].numpy())
            m.timer_start()
            m.run()
            ms += m.timer_stop_ms()
        out["ms"] = ms
        out["launches"] = m.launch_count() - l0
        out["trace"] = (read_trace(m) - tr0) // max(1, args.steps)
        sync.wait()
        if t == 0:
            barrier(dist, local)
            sampler.stop_flag = True
            barrier(dist, local)
        sync.wait()
        # e2e: host image in, logits out, everything inside the timed region
        m.timer_start()
        for i in range(args.steps):
            out["logits"] = step_e2e(t * 4 + i)
        out["ms_e2e"] = m.timer_stop_ms()
        sync.wait()

    threads = [threading.Thread(target=body, args=(t,)) for t in range(1, S)]
    for th in threads:
        th.start()
    body(0)
    for th in threads:
        th.join()
    barrier(dist, local)
    logits = res[0]["logits"]
    ms = max_over_ranks(dist, local, max(r["ms"] for r in res))
    ms_e2e = max_over_ranks(dist, local, max(r["ms_e2e"] for r in res))
    launches = sum(r["launches"] for r in res)
    trace = res[0]["trace"]
    out_level_bytes = 2 * N * 8  # Handle_output downloads the decrypted plaintext (2 limbs)

    if args.record_trace and rank == 0:
        path = os.path.join(ROOT, "tests", "emitted", MODEL + ".trace.json")
        json.dump({"model": MODEL, "classes": TRACE_CLASSES, "per_image": trace_to_json(trace)},
                  open(path, "w"), indent=0)
        print("trace written to", path, file=sys.stderr)
        out_dir = os.path.join(ROOT, "gpurun_out")  # the GPU box only brings this directory back
        if os.path.isdir(out_dir):
            json.dump({"model": MODEL, "classes": TRACE_CLASSES, "per_image": trace_to_json(trace)},
                      open(os.path.join(out_dir, MODEL + ".trace.json"), "w"), indent=0)
    m.close()

    # ---- rooflines, all measured live: the dominant kernel family (batched forward NTT over all
    # L+K limbs) and a table over the primitives of BASELINE.json config 2
    roof, by_kernel = None, None
    if rank == 0:
        roof, by_kernel = kernel_rooflines(local)
    sampler.join(timeout=2)

    total = world * args.steps * S
    line = {
        "metric": METRIC, "value": round(total / (ms / 1e3), 4), "unit": UNIT,
        "n_gpus": world, "steps": args.steps, "warmup": warm,
        "ms_per_step": round(ms / args.steps, 2), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "int64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "model": MODEL, "N": N, "L": DEPTH + 1, "dnum": PARTS,
                   "images_per_step_per_gpu": S, "streams_per_gpu": S,
                   "s_per_image": round(ms_single / 1e3, 4),
                   "s_per_image_note": "latency of one image alone on the GPU; value = throughput "
                                       "with %d image(s) in flight per GPU" % S,
                   "first_image_s": round(t_first, 3), "prepare_context_s": round(t_ctx, 2),
                   "l2": "working set per image (30 GB of switch keys, 225 MB of weights, "
                         "ciphertexts of 34 MB) exceeds the 126 MB L2",
                   "parallelism": "%d image(s) in flight per GPU (one host thread, stream and limb "
                                  "allocator each; tables and keys shared), full key replica per "
                                  "GPU, no collective" % S,
                   "logits0": [float(x) for x in logits[:3]]},
        "clocks": sampler.summary(),
        "e2e": {"value": round(total / (ms_e2e / 1e3), 4), "unit": UNIT,
                "h2d_bytes_per_step": S * 3 * 32 * 32 * 8, "d2h_bytes_per_step": S * out_level_bytes},
        "gpu_launches": int(launches),
    }
    if rank == 0:
        line["roofline"] = roof
        line["roofline_by_kernel"] = by_kernel
        if world == 1 and not args.no_cpu:
            base = cpu_baseline(trace_to_json(trace), threads=1)
            unit_us = base.pop("unit_us")
            for row in by_kernel:  # the reference's cost of the same primitive on one host core
                if row["name"] in unit_us:
                    row["cpu_us"] = round(unit_us[row["name"]], 1)
                    row["speedup_vs_1_core"] = round(unit_us[row["name"]] / row["us"], 1)
            line["cpu_baseline"] = base
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


# ---------------------------------------------------------------------------- reference
def reference_calibration():
    """composed / real: the composed CPU figure against REAL end-to-end runs of the same emitted unit
    on the unmodified reference (profiles/r2_reference_real_run.json, written by
    tools/reference_real_run.py; one entry per host it was run on)"""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "r2_reference_real_run.json")))
    except Exception:
        return None


def _interp(levels, costs, l):
    return float(np.interp(l, levels, costs))


def cpu_baseline(trace, threads):
    """seconds per image of the reference rtlib on the host = sum over the op trace of
    count(class, level) * unit cost(class, level) measured live on oracle/_ref (1 thread);
    with threads > 1 the images/s figure is scaled by the measured parallel efficiency of
    `threads` independent ciphertext chains (the reference parallelises over images with
    OpenMP, ant/dataset/resnet_cifar.main.inc:81)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from oracle_bindings import RefLib, C as _C, u32, i32
    t_start = time.time()
    ref = RefLib(N, DEPTH, Q0, SF, PARTS, HW, [1], with_bootstrap=False)
    rng = np.random.default_rng(3)
    mods = np.concatenate([ref.q, ref.p])
    L, K = ref.L, ref.K

    def poly(nq, ext=False):
        idx = list(range(nq)) + ([L + i for i in range(K)] if ext else [])
        return np.stack([rng.integers(0, mods[g], N, dtype=np.int64) for g in idx])

    def clock(fn, reps=1):
        best = 1e30
        for _ in range(reps):
            t = time.perf_counter()
            fn()
            best = min(best, time.perf_counter() - t)
        return best

    levels = [2, 12, 23, 34]
    unit = {c: [] for c in ("modup_digit", "moddown_poly", "rescale_poly", "encode")}
    msg = rng.uniform(-0.05, 0.05, 16384).astype(np.float32)
    for l in levels:
        a = poly(l)
        beta = min(PARTS, -(-l // ref.part_size))
        unit["modup_digit"].append(sum(clock(lambda j=j: ref.decomp_modup(a, j)) for j in range(beta)) / beta)
        e = poly(l, ext=True)
        unit["moddown_poly"].append(clock(lambda: ref.mod_down(e)))
        unit["rescale_poly"].append(clock(lambda: ref.rescale(a)))
        unit["encode"].append(clock(lambda: ref.encode_float(msg, 1, l)))
    x, y = poly(1)[0], poly(1)[0]
    order = ref.auto_order(1)[1] if isinstance(ref.auto_order(1), tuple) else ref.auto_order(1)
    limb = {"limb_mul": clock(lambda: ref.hw("modmul", 0, x, y), 5),
            "limb_add": clock(lambda: ref.hw("modadd", 0, x, y), 5),
            "limb_rot": clock(lambda: ref.hw("rotate", 0, x, order), 5),
            "limb_ntt": clock(lambda: ref.ntt(0, x), 5)}
    secs, parts = 0.0, {}
    for cls, by_level in trace.items():
        s = 0.0
        for lvl, cnt in by_level.items():
            s += cnt * (limb[cls] if cls in limb else _interp(levels, unit[cls], int(lvl)))
        parts[cls] = round(s, 1)
        secs += s
    eff = 1.0
    if threads > 1:
        f = ref.lib.ref_bench_chain
        f.restype, f.argtypes = _C.c_double, [_C.c_int, _C.c_int, u32, i32]
        t1 = f(1, 5, 17, 1)        # 5 iterations per thread: a single one moved the figure by
        tn = f(threads, 5, 17, 1)  # +-12 % from run to run
        eff = min(1.0, t1 / tn)
    top = len(levels) - 1
    unit_us = {"modup_digit L=%d" % levels[top]: 1e6 * unit["modup_digit"][top],
               "moddown_poly L=%d" % levels[top]: 1e6 * unit["moddown_poly"][top],
               "rescale_poly L=%d" % levels[top]: 1e6 * unit["rescale_poly"][top],
               "encode L=%d" % levels[top]: 1e6 * unit["encode"][top],
               "limb_mul x%d L=%d" % (levels[top], levels[top]): 1e6 * limb["limb_mul"] * levels[top],
               "limb_add x%d L=%d" % (levels[top], levels[top]): 1e6 * limb["limb_add"] * levels[top],
               "ntt x%d" % (L + K): 1e6 * limb["limb_ntt"] * (L + K), "ntt x1": 1e6 * limb["limb_ntt"]}
    calib = reference_calibration()
    return {"value": round(threads * eff / secs, 6), "unit": UNIT, "cores": threads,
            "kind": "reference", "unit_us": unit_us, "composed_over_real": calib,
            "s_per_image_1thread": round(secs, 1), "parallel_efficiency": round(eff, 3),
            "sample": "unit costs of Decomp_modup/Mod_down/Rescale/encode at levels %s and of a limb "
                      "mul/add/rotate/NTT on oracle/_ref (-O3), x the op trace of one image "
                      "(tests/emitted/%s.trace.json)" % (levels, MODEL),
            "seconds_by_class": parts, "sample_seconds": round(time.time() - t_start, 1)}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = max(1, min(os.cpu_count() or 1, 64))
    # every concurrent image needs its own working set (~8 GB) next to the shared 35 GB of keys
    try:
        mem_gb = os.sysconf("SC_PAGE_SIZE") * os.sysconf("SC_PHYS_PAGES") / 2**30
        threads = max(1, min(threads, int((mem_gb - 40) // 8)))
    except Exception:
        pass
    t0 = time.time()
    base = cpu_baseline(load_trace(MODEL), threads)
    base.pop("unit_us", None)
    s_img = base["s_per_image_1thread"]
    line = {"impl": "reference", "metric": METRIC, "value": base["value"], "unit": UNIT,
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": round(1e3 / base["value"], 1), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "int64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "model": MODEL, "N": N, "L": DEPTH + 1, "dnum": PARTS,
                       "s_per_image_1thread": s_img, "threads": threads},
            "cpu_baseline": base,
            "e2e": {"value": base["value"], "unit": UNIT, "h2d_bytes_per_step": 0,
                    "d2h_bytes_per_step": 0},
            "wall_s": round(time.time() - t0, 1)}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--streams", type=int, default=DEFAULT_STREAMS,
                    help="images in flight per GPU (host threads, one stream each)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--model", default=MODEL, choices=sorted(MODELS),
                    help="emitted model (BASELINE.json configs 1, 3-5); default: the headline ResNet-20")
    ap.add_argument("--record-trace", action="store_true",
                    help="write tests/emitted/<model>.trace.json from this run")
    args = ap.parse_args()
    select_model(args.model)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
