"""Instructions executed per CUDA source line for one kernel of an .ncu-rep (needs -lineinfo and --import-source on)."""
import csv, io, subprocess, sys, collections
rep, pat = sys.argv[1], sys.argv[2]
skip = sys.argv[3] if len(sys.argv) > 3 else "0"
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", "regex:" + pat,
                      "--launch-skip", skip, "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
agg = collections.OrderedDict()
fname, hdr = None, None
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        fname = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        hdr = r
        I = hdr.index("Instructions Executed")
        continue
    if hdr is None or len(r) < len(hdr):
        continue
    if r[0].isdigit():       # a CUDA line with its aggregate
        try:
            agg[(fname, int(r[0]))] = agg.get((fname, int(r[0])), 0) + int(r[I])
            src = r[1]
            agg.setdefault(("src", fname, int(r[0])), src)
        except ValueError:
            pass
tot = sum(v for k, v in agg.items() if len(k) == 2)
print("total", tot)
for k, v in sorted(((k, v) for k, v in agg.items() if len(k) == 2), key=lambda x: -x[1])[:int(sys.argv[4]) if len(sys.argv) > 4 else 40]:
    print("%-16s %5d %10d %.3f  %s" % (k[0], k[1], v, v / tot, agg[("src",) + k].strip()[:100]))
