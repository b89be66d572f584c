"""Instruction mix of the hot kernels from cuobjdump -sass (no GPU needed).

    python tools/sass_summary.py > profiles/r2_sass_summary.md

For ntt16_kernel<FWD_ROWS, LimbBatch, false> (ntt16.cu) the three arithmetic bodies are separated
at their EXIT instructions (FP64 body, integer lazy body, integer csub body); 64 butterflies and 16
canonicalisations per thread each.  For base_conv_kernel<12> (kernels.cu) the whole kernel."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OBJ = os.path.join(ROOT, "ace_compiler_b200", "build")


def sass(obj, needle):
    out = subprocess.run(["cuobjdump", "-sass", os.path.join(OBJ, obj)], capture_output=True, text=True).stdout
    fn, lines = None, []
    for l in out.splitlines():
        if "Function :" in l:
            fn = l.split("Function :")[1].strip()
        elif fn and needle in fn:
            m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(.*?);", l)
            if m:
                lines.append(m.group(1).strip())
    return lines


def mix(lines):
    c = collections.Counter()
    for l in lines:
        op = l.split()[1] if l.startswith("@") else l.split()[0]
        c[op] += 1
    return c


def table(title, c, per, unit):
    print("### %s\n" % title)
    print("| instruction | count | per %s |\n|---|---|---|" % unit)
    for op, n in c.most_common(14):
        print("| `%s` | %d | %.2f |" % (op, n, n / per))
    fma = sum(n for op, n in c.items() if op.startswith(("IMAD", "HFMA2", "FFMA")))
    wide = sum(n for op, n in c.items() if op.startswith(("IMAD.WIDE", "IMAD.HI")))
    dp = sum(n for op, n in c.items() if op.startswith(("DFMA", "DADD", "DMUL", "DSETP")))
    alu = sum(n for op, n in c.items() if op.startswith(("IADD3", "VIADD", "LOP3", "SHF", "SEL", "ISETP", "MOV", "LEA", "PRMT", "FSEL")))
    mem = sum(n for op, n in c.items() if op.startswith(("LDG", "STG", "LDS", "STS")))
    print("\nFMA-pipe instructions %d (of which wide/hi multiplies %d), FP64 %d, ALU %d, memory %d, total %d\n"
          % (fma, wide, dp, alu, mem, sum(c.values())))


def main():
    print("# SASS instruction mix of the hot kernels (round 2)\n")
    print("`cuobjdump -sass ace_compiler_b200/build/*.o`, sm_100a, nvcc 12.9, `-O3 --fmad=false`.\n")
    rows = sass("ntt16.o", "ntt16_kernelILi1ENS_9LimbBatchELb0")
    bodies, cur = [], []
    for l in rows:
        cur.append(l)
        if l.startswith("EXIT") or " EXIT" in l:
            bodies.append(cur)
            cur = []
    names = ["FP64 body (q < 2^50.4)", "64-bit integer lazy body (q < 2^57)", "64-bit integer body with conditional subtraction (q < 2^61)"]
    print("## `ntt16_kernel<FWD_ROWS>`: stages 8-15, 64 butterflies + 16 canonicalisations per thread\n")
    def which(b):  # the compiler orders the bodies as it likes: recognise them by their content
        c = mix(b)
        if c.get("DFMA", 0) > 50:
            return 0
        return 2 if c.get("SEL", 0) > 100 else 1
    bodies = [b for b in bodies if len(b) > 500]
    bodies.sort(key=which)
    for b in bodies:
        table(names[which(b)], mix(b), 64.0, "butterfly")
    rows = sass("kernels.o", "base_conv_kernelILi12")
    print("## `base_conv_kernel<12>` (all unrolled input counts 1..12)\n")
    table("whole kernel", mix(rows), 1.0, "kernel")
    # an excerpt: the first FP64 butterflies
    dp = [l for l in bodies[0]] if bodies else []
    start = next((i for i, l in enumerate(dp) if l.startswith("DMUL")), 0)
    print("## Excerpt: FP64 butterflies (h = v w; l = fma(v, w, -h); c = fma(h, 1/q, 1.5 2^52) - 1.5 2^52; r = fma(-c, q, h); t = r + l)\n\n```")
    for l in dp[start:start + 48]:
        print(l)
    print("```")


if __name__ == "__main__":
    main()
