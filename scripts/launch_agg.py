"""Aggregates an `ncu --csv` launch list (gpu__time_duration.sum + optional instruction metrics) per kernel."""
import collections
import csv
import sys


def agg(path):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr, data = rows[hi], rows[hi + 1:]
    kn, mn, mv, idc = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("ID")
    per = collections.defaultdict(dict)
    for r in data:
        if len(r) > mv:
            per[(r[idc], r[kn])][r[mn]] = float(r[mv].replace(",", ""))
    out = collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0])
    for (_i, k), m in per.items():
        a = out[k.split("(")[0]]
        a[0] += 1
        a[1] += m.get("gpu__time_duration.sum", 0) / 1e3
        a[2] += m.get("smsp__inst_executed.sum", 0)
        a[3] += m.get("smsp__thread_inst_executed_per_inst_executed.ratio", 0)
    return out


if __name__ == "__main__":
    for p in sys.argv[1:]:
        out = agg(p)
        tot = sum(a[1] for a in out.values())
        print(p, "total %.1f us" % tot)
        for k, a in sorted(out.items(), key=lambda x: -x[1][1]):
            print("  %-34s n=%3d  %9.1f us (%.3f)  mean %8.1f us  inst %8.1f M  active thr/inst %.1f"
                  % (k[:34], a[0], a[1], a[1] / tot, a[1] / a[0], a[2] / 1e6, a[3] / a[0]))
