"""DRAM traffic of the NTT kernels from an ncu --set full capture:

    ncu -i gpurun_out/<capture>.ncu-rep --page raw --csv > /tmp/raw.csv
    python tools/ncu_traffic.py /tmp/raw.csv profiles/r2_ncu_ntt_traffic.json

Sums dram__bytes_read.sum + dram__bytes_write.sum over one forward launch pair (K1 cols + K2 rows)
and writes the JSON bench.py reads for roofline.traffic, with the per-kernel counters beside it."""
import csv
import json
import sys


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}

    def val(r, name):
        v = float(r[col[name]])
        u = units[col[name]].lower()
        return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)

    want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
            "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
            "smsp__issue_active.avg.pct_of_peak_sustained_active",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
            "lts__throughput.avg.pct_of_peak_sustained_elapsed"]
    fp64 = [h for h in hdr if "pipe_fp64" in h and "pct_of_peak_sustained_active" in h and h.startswith("sm__inst_executed")]
    kernels, seen = [], set()
    for r in data:
        name = r[col["Kernel Name"]]
        short = name[:40]
        for i, n in enumerate(("fwd_cols", "fwd_rows", "inv_rows", "inv_cols")):
            if "ntt16_kernel<%d" % i in name or "ntt16_kernelILi%dE" % i in name:
                short = n
        if short in seen:
            continue
        seen.add(short)
        k = {"kernel": short, "grid": r[col["Grid Size"]] if "Grid Size" in col else None}
        for w in want + fp64:
            if w in col:
                k[w] = val(r, w) if "bytes" in w else float(r[col[w]])
        kernels.append(k)
    fwd = [k for k in kernels if k["kernel"].startswith("fwd")]
    total = sum(k["dram__bytes_read.sum"] + k["dram__bytes_write.sum"] for k in fwd)
    json.dump({"dram_bytes_per_launch": total, "what": "forward NTT, one launch pair over 45 limbs "
               "(dram__bytes_read.sum + dram__bytes_write.sum of K1 + K2, ncu --set full)",
               "kernels": kernels}, open(sys.argv[2], "w"), indent=1)
    print(total, [k["kernel"] for k in kernels])


if __name__ == "__main__":
    main()
