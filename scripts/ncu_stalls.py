"""Per CUDA source line: instructions executed, stall samples and the dominant stall reasons, shared-memory wavefronts,
for one kernel of an .ncu-rep (needs -lineinfo, --import-source on, --set full)."""
import csv, io, subprocess, sys, collections
rep, pat = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", "regex:" + pat,
                      "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
fname, hdr = None, None
lines = {}
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        fname = r[1].split("/")[-1]; continue
    if r[0] == "Line No":
        hdr = r; continue
    if hdr is None or len(r) < len(hdr) or not r[0].isdigit():
        continue
    d = dict(zip(hdr, r))
    def num(k):
        try:
            return int(d.get(k, "0") or 0)
        except ValueError:
            return 0
    key = (fname, int(r[0]))
    e = lines.setdefault(key, collections.Counter())
    e["inst"] += num("Instructions Executed"); e["samples"] += num("# Samples")
    e["wave"] += num("L1 Wavefronts Shared"); e["ideal"] += num("L1 Wavefronts Shared Ideal")
    for k in hdr:
        if k.startswith("stall_") and "Not Issued" not in k:
            e[k] += num(k)
    e.src = r[1]
    lines[key] = e
    if not hasattr(e, "src"):
        pass
srcs = {}
for r in rows:
    if r and r[0].isdigit() and len(r) > 1:
        pass
ti = sum(e["inst"] for e in lines.values()); ts = sum(e["samples"] for e in lines.values())
print("total inst %d samples %d" % (ti, ts))
tot = collections.Counter()
for e in lines.values():
    for k, v in e.items():
        if k.startswith("stall_"):
            tot[k] += v
print("stall totals:", ", ".join("%s %.3f" % (k[6:], v / ts) for k, v in tot.most_common(8)))
for key, e in sorted(lines.items(), key=lambda x: -x[1]["samples"])[:top]:
    st = ", ".join("%s %d" % (k[6:], v) for k, v in sorted(((k, v) for k, v in e.items() if k.startswith("stall_")), key=lambda x: -x[1])[:3])
    print("%-14s %5d inst %8d %.3f samples %6d %.3f wave %8d/%8d | %s" % (key[0][:14], key[1], e["inst"], e["inst"] / ti, e["samples"], e["samples"] / ts,
                                                                   e["wave"], e["ideal"], st))
