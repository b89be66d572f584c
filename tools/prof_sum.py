"""Sums the ACE_B200_PROF reports of one run by section (emitted code vs Bootstrap) and scope.
usage: python tools/prof_sum.py gpurun_out/prof_x.log [image_index]"""
import collections
import re
import sys

img = int(sys.argv[2]) if len(sys.argv) > 2 else 1
sec, cur, image = collections.defaultdict(lambda: collections.defaultdict(lambda: [0, 0.0])), None, 0
for line in open(sys.argv[1]):
    if line.startswith("[driver] image"):
        image += 1
        continue
    m = re.match(r"\[ace_b200 prof\] (\w+): device time", line)
    if m:
        cur = m.group(1)
        continue
    m = re.match(r"\[ace_b200 prof\] (\S.*?)\s+(\d+) scopes\s+([\d.]+) ms", line)
    if m and cur and image == img:
        a = sec[cur][m.group(1).strip()]
        a[0] += int(m.group(2))
        a[1] += float(m.group(3))
for s, d in sec.items():
    tot = sum(v[1] for k, v in d.items() if not k.startswith("api.") and k != "encode(total)")
    n = sum(v[0] for k, v in d.items() if not k.startswith("api.") and k != "encode(total)")
    print("== %s: %.1f ms in %d scopes (event pairs add ~4 us per scope)" % (s, tot, n))
    for k, v in sorted(d.items(), key=lambda kv: -kv[1][1]):
        print("   %-22s %7d scopes %9.2f ms %8.2f us/scope" % (k, v[0], v[1], 1e3 * v[1] / max(1, v[0])))
