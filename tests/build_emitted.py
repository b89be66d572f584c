"""Compiles ACE-emitted C (the checked-in example programs under
/root/reference/fhe-cmplr/rtlib/ant/example, used UNMODIFIED, where they lie) against OUR
header tree (include/) and libace_b200.so.  This is the drop-in check of the boundary: the
same translation units the reference links against libFHErt_ant.a.
Binaries go to tests/_emitted_bin/ (git-ignored, shipped to the GPU box by gpurun).
Run in the container that has /root/reference; a no-op elsewhere."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_EX = "/root/reference/fhe-cmplr/rtlib/ant/example"
OUT = os.path.join(ROOT, "tests", "_emitted_bin")
REF_DS = "/root/reference/fhe-cmplr/rtlib/ant/dataset"
# ACE-emitted CIFAR ResNets checked in by the reference
MODELS = ["resnet20_cifar10_pre", "resnet32_cifar100_pre", "resnet56_cifar10_pre",
          "resnet110_cifar10_train"]
EXAMPLES = ["add", "add_const", "mul_const", "rotate", "rotate_02", "relin", "relin_02",
            "gemm", "gemm_02", "conv2d", "avg_pool", "relu", "bootstrap", "bootstrap_02"]


def build_all(verbose=False):
    if not os.path.isdir(REF_EX):
        return []
    os.makedirs(OUT, exist_ok=True)
    built = []
    for name in EXAMPLES:
        src = os.path.join(REF_EX, "eg_fhertlib_%s.c" % name)
        exe = os.path.join(OUT, "eg_" + name)
        cmd = ["gcc", "-O2", "-w", "-std=gnu11", "-I", os.path.join(ROOT, "include"), "-I", REF_EX,
               src, "-o", exe, "-L", os.path.join(ROOT, "ace_compiler_b200"), "-lace_b200",
               "-Wl,-rpath,$ORIGIN/../../ace_compiler_b200", "-lm"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            print("FAILED", name, r.stderr[:2000])
        else:
            built.append(exe)
            if verbose:
                print("built", exe)
    # our own emitted-style unit for the pre-encoded weight path (no reference program uses it)
    src = os.path.join(ROOT, "tests", "emitted", "pt_get_case.c")
    exe = os.path.join(OUT, "pt_get_case")
    r = subprocess.run(["gcc", "-O2", "-w", "-std=gnu11", "-I", os.path.join(ROOT, "include"), src, "-o", exe,
                        "-L", os.path.join(ROOT, "ace_compiler_b200"), "-lace_b200",
                        "-Wl,-rpath,$ORIGIN/../../ace_compiler_b200", "-lm"], capture_output=True, text=True)
    if r.returncode != 0:
        print("FAILED pt_get_case", r.stderr[:2000])
    else:
        built.append(exe)
    built += build_models(verbose)
    return built


def build_models(verbose=False, models=MODELS):
    """the emitted ResNet translation units, #included unmodified by tests/emitted/resnet_driver.c"""
    built = []
    drv = os.path.join(ROOT, "tests", "emitted", "resnet_driver.c")
    for m in models:
        inc = os.path.join(REF_DS, m + ".onnx.inc")
        if not os.path.exists(inc):
            continue
        exe = os.path.join(OUT, m)
        if os.path.exists(exe) and os.path.getmtime(exe) > max(os.path.getmtime(drv), os.path.getmtime(inc)):
            built.append(exe)
            continue
        cmd = ["gcc", "-O1", "-w", "-std=gnu11", "-DMODEL_INC=\"%s\"" % inc, "-I",
               os.path.join(ROOT, "include"), drv, "-o", exe, "-L",
               os.path.join(ROOT, "ace_compiler_b200"), "-lace_b200",
               "-Wl,-rpath,$ORIGIN/../../ace_compiler_b200", "-lm"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            print("FAILED", m, r.stderr[:2000])
            continue
        built.append(exe)
        # the same unit as a shared object (ace_compiler_b200/models/lib<model>.so) for in-process
        # drivers: bench.py and the whole-model parity test load it with RTLD_GLOBAL
        mdir = os.path.join(ROOT, "ace_compiler_b200", "models")
        os.makedirs(mdir, exist_ok=True)
        so = os.path.join(mdir, "lib%s.so" % m)
        cmd = ["gcc", "-O1", "-w", "-std=gnu11", "-fPIC", "-shared", "-include", "common/rtlib.h",
               "-x", "c", inc, "-I", os.path.join(ROOT, "include"), "-o", so, "-L",
               os.path.join(ROOT, "ace_compiler_b200"), "-lace_b200", "-Wl,-rpath,$ORIGIN/..",
               "-lm"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            print("FAILED", so, r.stderr[:2000])
        else:
            built.append(so)
        if verbose:
            print("built", exe, so)
    return built


if __name__ == "__main__":
    b = build_all(verbose=True)
    print(len(b), "of", len(EXAMPLES), "example programs built")
