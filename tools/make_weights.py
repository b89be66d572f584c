"""Synthetic weight ("message") file for an ACE-emitted model, derived from the emitted C alone.

The emitted translation unit reads its weights through Pt_from_msg(pt, index, len, ...) from an
external data file (reference format: fhe-cmplr/include/fhe/core/rt_data_def.h:90-109, reader
fhe-cmplr/rtlib/common/src/rt_data_file.c:26-126): a 4 KiB header page, float32 entries, then a
lookup table of {name[16], index, size, offset}.  Real weight files are a compiler output that
cannot be regenerated offline, so this tool lays out a file with the same entry table --
entry ranges and lengths are recovered from the Pt_from_msg call sites and the
`extern float32_t _cst_N[size]` declarations -- and fills it with small seeded values
(U(-a, a), numpy default_rng(seed)) so that activations stay inside the bootstrap's (-1, 1).

    python tools/make_weights.py <model.onnx.inc | entries.json> <out.msg> [--amp 0.05] [--seed 1]
    python tools/make_weights.py <model.onnx.inc> --table entries.json     # only the entry table

The entry tables of the checked-in reference models are committed under tests/emitted/ so that
the file can be regenerated on a machine that does not have the reference tree.
"""
import argparse
import json
import re
import struct

import numpy as np

PAGE = 4096


def parse_entries(inc_text):
    """returns [(first_index, count, len, name)] sorted by first_index"""
    sizes = {m.group(1): int(m.group(2))
             for m in re.finditer(r"extern float32_t (_cst_\d+)\[(\d+)\];", inc_text)}
    calls = {}
    for m in re.finditer(r"Pt_from_msg\(&\w+, (.*?)/\* (cst_\d+)(?:_\d+)? \*/, (\d+),", inc_text):
        expr, name, ln = m.group(1), m.group(2), int(m.group(3))
        base = int(re.findall(r"(\d+)\s*$", expr.strip())[0])
        if base in calls and calls[base][1] != ln:
            raise ValueError("inconsistent entry length at index %d" % base)
        calls[base] = (name, ln)
    bases = sorted(calls)
    out = []
    for i, b in enumerate(bases):
        name, ln = calls[b]
        total = sizes["_" + name]
        cnt_decl = max(1, total // ln)
        cnt = (bases[i + 1] - b) if i + 1 < len(bases) else cnt_decl
        out.append((b, cnt, ln, name))
    return out


def write_file(path, entries, amp, seed, model="synthetic"):
    rng = np.random.default_rng(seed)
    n_ent = sum(c for _, c, _, _ in entries)
    lut = []
    ofs = PAGE
    with open(path, "wb") as f:
        f.write(b"\0" * PAGE)
        for first, cnt, ln, name in entries:
            for k in range(cnt):
                data = rng.uniform(-amp, amp, ln).astype(np.float32)
                f.write(data.tobytes())
                lut.append((name.encode()[:15], first + k, ln * 4, ofs))
                ofs += ln * 4
                pad = (-ofs) % 32  # 32-byte aligned entries (ent_align = 5)
                f.write(b"\0" * pad)
                ofs += pad
        lut_ofs = ofs
        for name, idx, size, o in sorted(lut, key=lambda t: t[1]):
            f.write(struct.pack("<16sIIQ", name, idx, size, o))
        hdr = struct.pack("<8sIHBBQQqq48s40s", b"!ANTFHE\0", 0, 0, 0, 5, n_ent, lut_ofs,
                          0, 0, model.encode()[:47], b"XXXXXXXX-XXXX-XXXX-XXXX-XXXXXXXXXXXX")
        f.seek(0)
        f.write(hdr)
    return n_ent, lut_ofs


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("inc")
    ap.add_argument("out", nargs="?")
    ap.add_argument("--table")
    ap.add_argument("--amp", type=float, default=0.05)
    ap.add_argument("--seed", type=int, default=1)
    a = ap.parse_args()
    if a.inc.endswith(".json"):
        ent = [tuple(e) for e in json.load(open(a.inc))]
    else:
        ent = parse_entries(open(a.inc).read())
    idx = 0
    for first, cnt, ln, name in ent:
        assert first == idx, "entry table has a gap at %d (next constant starts at %d)" % (idx, first)
        idx += cnt
    if a.table:
        json.dump(ent, open(a.table, "w"))
        print("%d constants, %d entries -> %s" % (len(ent), idx, a.table))
    if a.out:
        n, lut = write_file(a.out, ent, a.amp, a.seed)
        print("%d entries, %.1f MB, lut at %d" % (n, lut / 1e6, lut))
