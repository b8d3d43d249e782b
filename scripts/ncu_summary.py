"""Prints the key metrics and the stall / instruction hot spots of the kernels in an .ncu-rep file."""
import csv, subprocess, sys, collections, io
rep = sys.argv[1]
pat = sys.argv[2] if len(sys.argv) > 2 else "."
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
want = ['Kernel Name', 'gpu__time_duration.sum', 'smsp__inst_executed.sum', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__grid_size', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'sm__cycles_elapsed.max', 'smsp__cycles_active.avg',
        'dram__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed']
for r in rows[2:]:
    print('---')
    for w in want:
        if w in idx:
            print(w, r[idx[w]], units[idx[w]])
    for h in hdr:
        if 'issue_stalled' in h and 'per_issue_active' in h:
            try:
                v = float(r[idx[h]])
            except ValueError:
                continue
            if v > 0.3:
                print('   ', h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''), v)
if len(sys.argv) > 3:
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass", "--kernel-name", "regex:" + pat,
                          "--launch-count", "1"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    hdr = rows[1]
    I, S, A, SRC = hdr.index('Instructions Executed'), hdr.index('# Samples'), hdr.index('Address'), hdr.index('Source')
    data, seen = [], set()
    for r in rows[2:]:
        try:
            a = int(r[A], 16)
        except (ValueError, IndexError):
            continue
        if a in seen:
            continue
        seen.add(a)
        data.append((a, r[SRC].strip(), int(r[I]), int(r[S])))
    base = data[0][0]
    tot, ts = sum(d[2] for d in data), sum(d[3] for d in data)
    print('total warp inst', tot, 'samples', ts)
    b = collections.OrderedDict()
    for a, s, i, sm in data:
        k = (a - base) // 0x200
        b.setdefault(k, [0, 0, s])
        b[k][0] += i
        b[k][1] += sm
    for k, v in b.items():
        if v[0] / tot > 0.02 or v[1] / ts > 0.02:
            print(hex(k * 0x200), 'inst %.3f' % (v[0] / tot), 'stall %.3f' % (v[1] / ts), v[2][:60])
    print('top stall lines:')
    for a, s, i, sm in sorted(data, key=lambda d: -d[3])[:14]:
        print(hex(a - base), sm, i, s[:80])
