#!/usr/bin/env python
"""Benchmark of the GenoFix hot path on B200 (contract: see the task statement / DESIGN.md section 6).

  python bench.py [--config c2] --gpus N --steps K --warmup W   # CUDA path (this repo)
  python bench.py --impl reference --gpus N --steps K ...         # the reference's CPU algorithm (oracle port) on host cores

Workloads (BASELINE.json `configs`; c2 is the default and the one the metric is quoted on):
  c1  examples/simulate pedigree (4 560 animals) x 1 000 SNPs, 1 % errors, 1 000-SNP chunks (simulate.py -c 1000)
  c2  synthetic 10k-animal 6-generation pedigree x 50k SNPs, 1 % errors, 10 000-SNP chunks (genofix.py -C 10000)
  c3  100k animals x 600k SNPs in 60 chunks of 10 000 SNPs dealt round-robin to the ranks (STRONG scaling: the problem is
      fixed, every rank generates and corrects only its own chunks; one all-reduce of the per-animal tallies)
  c4  deep 15-generation pedigree with inbreeding loops, 10 % unknown parent pairs, 5 % missing calls x 50k SNPs
  c5  trio consistency scan, 1M animals x 50k SNPs, 2-bit packed (k_trio_scan)
A step = one pass of the path over the whole matrix.  With --gpus N and c1/c2/c4 every rank owns the SNP blocks
[rank * S, (rank + 1) * S) of ONE N*S-SNP matrix over the replicated pedigree (weak scaling: SNP blocks are independent
given the pedigree schedule); the single collective is the all-reduce of the per-animal tallies.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

PARAMS = dict(threshold_pair=0.5, threshold_single=0.64, init_filter_p=0.95, tiethreshold=0.4, err_thresh=1.0)  # genofix.py:95-112
BACK = 2
RANK_BYTES_PER_CELL = 2.25       # 0.25 B packed code read + 2 B float16 score written
PIPELINE_BYTES_PER_CELL = 4.75   # rank (2.25) + correction (0.25 + 2 read, 0.25 written), SURVEY 8(d)
SCAN_BYTES_PER_CELL = 0.25
METRIC = "individual x SNP genotype cells corrected/sec"

CONFIGS = {
    "c1": dict(workload="examples/simulate pedigree (4 560 animals) x 1 000 SNPs, 1% errors, 1 000-SNP chunks like "
                        "simulate.py -c 1000 (BASELINE configs[0])",
               pedigree="example_fam", animals=4560, generations=8, snps=1000, chunk=1000),
    "c2": dict(workload="10k-animal 6-generation synthetic pedigree x 50k SNPs, 1% errors (BASELINE configs[1]), "
                        "corrected in 10k-SNP chunks like genofix.py",
               animals=10000, generations=6, snps=50000, chunk=10000),
    "c3": dict(workload="100k-animal 10-generation synthetic pedigree x 600k SNPs, 1% errors (BASELINE configs[2]): 60 chunks "
                        "of 10k SNPs dealt round-robin to the ranks, inputs dropped on the device per chunk",
               animals=100000, generations=10, snps=600000, chunk=10000, sharded=True),
    "c4": dict(workload="deep 15-generation pedigree (1 000 per generation, 2% sires / 30% dams, related matings 0.25, 10% of "
                        "the non-founders with both parents unknown) x 50k SNPs, 1% errors, 5% missing calls "
                        "(BASELINE configs[3])",
               animals=15000, generations=15, snps=50000, chunk=10000, sire_frac=0.02, dam_frac=0.3, related=0.25,
               unknown=0.10, missing=0.05),
    "c5": dict(workload="trio consistency scan + founder extraction, 1M animals x 50k SNPs, 2-bit packed "
                        "(BASELINE configs[4])",
               animals=1000000, generations=10, snps=50000, chunk=50000, trio_scan=True),
}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="c2", choices=sorted(CONFIGS))
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--animals", type=int, default=None)
    ap.add_argument("--snps", type=int, default=None)
    ap.add_argument("--chunk", type=int, default=None)
    ap.add_argument("--generations", type=int, default=None)
    ap.add_argument("--lanes", type=int, default=3, help="chunks corrected concurrently per GPU (streams + host threads)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=20.0, help="target CPU time of one baseline sample")
    ap.add_argument("--e2e-chunks", type=int, default=8, help="c3: chunks per rank of the end-to-end (host buffer) leg")
    a = ap.parse_args()
    cfg = dict(CONFIGS[a.config])
    for k in ("animals", "snps", "chunk", "generations"):
        if getattr(a, k) is not None:
            cfg[k] = getattr(a, k)
            cfg["workload"] = "%s [overridden: %s=%d]" % (cfg["workload"], k, cfg[k])
    a.cfg = cfg
    return a


def make_pedigree(cfg):
    from genofix_b200 import synth
    if cfg.get("pedigree") == "example_fam":
        g = np.load(os.path.join(ROOT, "tests", "golden", "example_fam.npz"))
        return g["ped_rows"], synth.complete_rows(g["ped_rows"])          # (pedigree table, gene-drop rows)
    rows = synth.make_pedigree(cfg["animals"], cfg["generations"], sire_frac=cfg.get("sire_frac", 0.05),
                               dam_frac=cfg.get("dam_frac", 0.5), seed=1234, related_mating_p=cfg.get("related", 0.0),
                               unknown_parents_frac=cfg.get("unknown", 0.0))
    return rows, rows


def host_matrix(cfg, drop_rows, n_snps, seed):
    from genofix_b200 import synth
    obs = synth.inject_errors(synth.gene_drop(drop_rows, n_snps, seed=seed), 0.01, seed=seed)
    if cfg.get("missing", 0) > 0:
        obs = synth.inject_missing(obs, cfg["missing"], seed=seed + 1)
    return obs


# --------------------------------------------------------------------------------------------------
# CPU baseline: the reference's algorithm (oracle port) on the host cores, bounded sample
# --------------------------------------------------------------------------------------------------
def _cpu_worker(args):
    rows, ids, sub = args
    from oracle.genofix_oracle import Oracle
    import io, contextlib
    with contextlib.redirect_stdout(io.StringIO()):
        Oracle(rows, ids).correct_matrix(sub, PARAMS["threshold_pair"], PARAMS["threshold_single"], back=BACK,
                                         init_filter_p=PARAMS["init_filter_p"], tiethreshold=PARAMS["tiethreshold"],
                                         err_thresh=PARAMS["err_thresh"])
    return sub.size


def cpu_baseline(rows, ids, obs, seconds, cores=None, rate_hint=None):
    """Times the oracle port on `cores` processes, each correcting its own block of SNP columns of the same matrix
    (full pedigree): ONE sample of about `seconds` of wall time, at least 13 SNPs per process (>= 200 SNPs on 16
    cores) unless the matrix is so tall that this would take minutes."""
    import multiprocessing as mp
    cores = cores or os.cpu_count() or 1
    n = obs.shape[0]
    if rate_hint is None:
        t0 = time.time()
        _cpu_worker((rows, ids, np.ascontiguousarray(obs[:, :2])))
        rate_hint = 2 * n / max(time.time() - t0, 1e-3)          # cells/s of one process (pessimistic: set-up included)
    per_proc = int(round(seconds * rate_hint / n))
    per_proc = max(13 if 13 * n / rate_hint <= 4 * seconds else 1, per_proc)
    per_proc = int(max(1, min(obs.shape[1] // cores, per_proc)))
    blocks = [np.ascontiguousarray(obs[:, i * per_proc:(i + 1) * per_proc]) for i in range(cores)]
    ctx = mp.get_context("fork")
    t0 = time.time()
    with ctx.Pool(cores) as pool:
        cells = sum(pool.map(_cpu_worker, [(rows, ids, b) for b in blocks]))
    dt = time.time() - t0
    return dict(value=cells / dt, unit="cells/s", cores=cores, kind="port",
                sample="%d animals x %d SNPs (%d per process x %d processes, full pedigree, columns of the bench matrix), %.1f s"
                       % (n, per_proc * cores, per_proc, cores, dt)), rate_hint


def clocks_sampler(stop, out, gpu_index):
    q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
    while not stop.is_set():
        try:
            r = subprocess.run(["nvidia-smi", "-i", str(gpu_index), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                               capture_output=True, text=True, timeout=5)
            f = [x.strip() for x in r.stdout.strip().split(",")]
            if len(f) >= 6:
                out.append(f)
        except Exception:
            pass
        stop.wait(0.2)


def clocks_summary(samples):
    sm = [float(s[0]) for s in samples if s[0].replace(".", "").isdigit()]
    reasons = []
    for name, idx in (("hw_slowdown", 2), ("hw_thermal_slowdown", 3), ("sw_thermal_slowdown", 4), ("sw_power_cap", 5)):
        if any(s[idx].lower().startswith("active") for s in samples):
            reasons.append(name)
    return dict(sm_mhz=float(np.median(sm)) if sm else None, sm_max_mhz=float(samples[0][1]) if samples else None,
                reasons=reasons, samples=len(samples))


def config_block(args, world, lanes, extra=None):
    cfg = args.cfg
    c = dict(workload=cfg["workload"], config=args.config, animals=cfg["animals"], snps=cfg["snps"], chunk_snps=cfg["chunk"],
             back=BACK, lanes=lanes, **PARAMS)
    per_chunk = cfg["animals"] * min(cfg["chunk"], cfg["snps"]) * 2.25
    if per_chunk > 2 * 126e6:
        c["l2"] = "inputs larger than L2: every chunk streams %.0f MB of packed codes + %.0f MB of float16 scores (126 MB L2)" \
                  % (cfg["animals"] * cfg["chunk"] / 4 / 1e6, cfg["animals"] * cfg["chunk"] * 2 / 1e6)
    else:
        c["l2"] = "working set of a chunk (%.0f MB) fits in the 126 MB L2: a 256 MB buffer is written before every timed step " \
                  "(flush, inside the timed region)" % (per_chunk / 1e6)
    if extra:
        c.update(extra)
    return c


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cfg = args.cfg
    ped_rows, drop_rows = make_pedigree(cfg)
    cores = os.cpu_count() or 1
    obs = host_matrix(cfg, drop_rows, max(64, min(cfg["snps"], 16 * max(13, cores))), seed=1234)
    ids = drop_rows[:, 0]
    vals, cb, hint = [], None, None
    for i in range(args.warmup + args.steps):
        cb, hint = cpu_baseline(ped_rows, ids, obs, args.cpu_seconds, rate_hint=hint)
        if i >= args.warmup:
            vals.append(cb["value"])
    v = float(np.mean(vals)) if vals else 0.0
    cb["value"] = v
    cb["spread"] = dict(min=float(np.min(vals)), max=float(np.max(vals)), samples=len(vals)) if vals else None
    line = dict(metric=METRIC, value=v, unit="cells/s", n_gpus=args.gpus, steps=args.steps, warmup=args.warmup, ms_per_step=None,
                higher_is_better=True, scaling="strong" if cfg.get("sharded") else "weak", vs_baseline=None, dtype="f64",
                data="synthetic", impl="reference",
                config=config_block(args, 1, 0, dict(note="CPU arm: every step times one bounded SNP sample of the workload")),
                cpu_baseline=cb, e2e=dict(value=v, unit="cells/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    print(json.dumps(line))


# --------------------------------------------------------------------------------------------------
# config 5: trio consistency scan
# --------------------------------------------------------------------------------------------------
def run_trio_scan(args, torch, dist, world, rank, local):
    from genofix_b200 import _lib, synth
    from genofix_b200.engine import Schedule
    from genofix_b200.pedigree.pedigree_dag import PedigreeDAG
    from genofix_b200.simulate_device import DeviceSimulator
    cfg = args.cfg
    L = _lib.lib()
    dev = torch.device("cuda", local)
    n = cfg["animals"]
    s_rank = cfg["snps"] // world if world > 1 else cfg["snps"]      # SNP blocks of the one matrix, dealt to the ranks
    rows = synth.make_pedigree(n, cfg["generations"], seed=1234)
    sched = Schedule(PedigreeDAG.from_table(rows), rows[:, 0], BACK)
    sim = DeviceSimulator(rows, s_rank, seed=1234 + rank)
    packed = sim.inject_errors(sim.gene_drop(), 0.01)
    del sim.hap
    wpr = packed.shape[1]
    inc = torch.empty(n, dtype=torch.int64, device=dev)
    tst = torch.empty(n, dtype=torch.int64, device=dev)
    pin_in = torch.empty(packed.shape, dtype=torch.int32, pin_memory=True)
    pin_in.copy_(packed)
    pin_out = torch.empty((2, n), dtype=torch.int64, pin_memory=True)
    staged = torch.empty_like(packed)

    def step():
        _lib.check(L.gfx_trio_scan(sched.handle, packed.data_ptr(), s_rank, wpr, inc.data_ptr(), tst.data_ptr(),
                                   torch.cuda.current_stream().cuda_stream))
        if world > 1:
            dist.all_reduce(inc)
            dist.all_reduce(tst)

    # the same matrix as column chunks of 2048 SNPs (one 512-byte tile per row and chunk), the layout SURVEY 8(d)'s tile rule
    # describes: consecutive rows of a chunk are adjacent in HBM, so a litter is one contiguous read
    CW = 2048
    chunks = [(packed[:, a // 16:min(wpr, (a + CW) // 16)].contiguous(), min(CW, s_rank - a)) for a in range(0, s_rank, CW)]
    inc_c = torch.empty(n, dtype=torch.int64, device=dev)
    tst_c = torch.empty(n, dtype=torch.int64, device=dev)

    def step_chunked():
        inc_c.zero_()
        tst_c.zero_()
        for p_c, s_c in chunks:
            _lib.check(L.gfx_trio_scan_add(sched.handle, p_c.data_ptr(), s_c, p_c.shape[1], inc_c.data_ptr(), tst_c.data_ptr(),
                                           torch.cuda.current_stream().cuda_stream))
        if world > 1:
            dist.all_reduce(inc_c)
            dist.all_reduce(tst_c)

    def step_e2e():
        staged.copy_(pin_in, non_blocking=True)
        _lib.check(L.gfx_trio_scan(sched.handle, staged.data_ptr(), s_rank, wpr, inc.data_ptr(), tst.data_ptr(),
                                   torch.cuda.current_stream().cuda_stream))
        if world > 1:
            dist.all_reduce(inc)
            dist.all_reduce(tst)
        pin_out[0].copy_(inc, non_blocking=True)
        pin_out[1].copy_(tst, non_blocking=True)
        torch.cuda.synchronize()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    stop, samples = threading.Event(), []
    th = threading.Thread(target=clocks_sampler, args=(stop, samples, local), daemon=True)
    th.start()
    l0 = L.gfx_launch_count()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(args.steps):
        step()
    t1.record()
    barrier()
    launches = int(L.gfx_launch_count() - l0)
    ms = t0.elapsed_time(t1)
    stop.set()
    th.join(timeout=2)
    for _ in range(3):
        step_chunked()
    barrier()
    c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    c0.record()
    for _ in range(args.steps):
        step_chunked()
    c1.record()
    barrier()
    ms_chunked = c0.elapsed_time(c1)
    if not (torch.equal(inc_c, inc) and torch.equal(tst_c, tst)):
        raise SystemExit("bench.py: the chunked scan disagrees with the scan of the row-major matrix")
    step_e2e()
    barrier()
    w0 = time.perf_counter()
    for _ in range(args.steps):
        step_e2e()
    barrier()
    e2e_ms = (time.perf_counter() - w0) * 1e3
    tm = torch.tensor([ms, e2e_ms, ms_chunked], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
    ms, e2e_ms, ms_chunked = float(tm[0]), float(tm[1]), float(tm[2])
    # the same scan against a torch restatement on a sample of trios (torch is the checker here, not the product)
    k = sched._keep
    has = np.nonzero((k["sire"] >= 0) & (k["dam"] >= 0))[0]
    sel = has[np.random.default_rng(0).choice(len(has), 64, replace=False)]
    T = torch.tensor([[1, .5, 0, .5, .25, 0, 0, 0, 0], [0, .5, 1, .5, .5, .5, 1, .5, 0], [0, 0, 0, 0, .25, .5, 0, .5, 1]],
                     device=dev).reshape(3, 3, 3)
    sh = torch.arange(16, device=dev) * 2

    def unpack(r):
        w = packed[int(r)].to(torch.int64) & 0xFFFFFFFF
        return ((w[:, None] >> sh[None, :]) & 3).reshape(-1)[:s_rank]
    if world == 1:
        ok = True
        inc_h, tst_h = inc.cpu().numpy(), tst.cpu().numpy()
        for a in sel:
            kk, ss, dd = unpack(k["row"][a]), unpack(k["row"][k["sire"][a]]), unpack(k["row"][k["dam"][a]])
            pres = (kk < 3) & (ss < 3) & (dd < 3)
            bad = pres & (T[kk.clamp(max=2), ss.clamp(max=2), dd.clamp(max=2)] == 0)
            ok = ok and int(bad.sum()) == int(inc_h[k["row"][a]]) and int(pres.sum()) == int(tst_h[k["row"][a]])
        if not ok:
            raise SystemExit("bench.py: k_trio_scan disagrees with the torch restatement on the sampled trios")
    peaks = load_peaks()
    peak = float(peaks.get("hbm_gbs", 6650.0))
    cells = n * s_rank * world
    ach = n * s_rank * SCAN_BYTES_PER_CELL / (ms / args.steps / 1e3) / 1e9
    if rank == 0:
        founders = int(len(sched.founders()))
        line = dict(metric="individual x SNP genotype cells scanned/sec (trio consistency scan)", value=cells * args.steps / (ms / 1e3),
                    unit="cells/s", n_gpus=world, steps=args.steps, warmup=max(args.warmup, 3), ms_per_step=ms / args.steps,
                    higher_is_better=True, scaling="strong", vs_baseline=None, dtype="u8 (2-bit codes, popcounts)",
                    data="synthetic",
                    config=config_block(args, world, 1, dict(per_gpu_matrix="%d x %d" % (n, s_rank), founders=founders,
                                                             l2="%.1f GB of packed codes per step (126 MB L2)" % (n * wpr * 4 / 1e9))),
                    e2e=dict(value=cells * args.steps / (e2e_ms / 1e3), unit="cells/s", ms_per_step=e2e_ms / args.steps,
                             h2d_bytes_per_step=int(packed.numel() * 4) * world, d2h_bytes_per_step=16 * n * world),
                    gpu_launches=launches,
                    roofline=dict(bound="hbm", achieved=ach, peak=peak, unit="GB/s", frac=ach / peak, traffic=None,
                                  kernel="k_trio_scan (one warp per family)", algorithmic_bytes_per_cell=SCAN_BYTES_PER_CELL,
                                  peak_source="MEASURED_PEAKS.json" if peaks else "fallback 6.65 TB/s",
                                  timed="inside the step: the scan is the step",
                                  chunked=dict(layout="%d column chunks of %d SNPs, each [n, %d words] contiguous" % (len(chunks), CW, CW // 16),
                                               ms_per_step=ms_chunked / args.steps,
                                               achieved=n * s_rank * SCAN_BYTES_PER_CELL / (ms_chunked / args.steps / 1e3) / 1e9,
                                               frac=n * s_rank * SCAN_BYTES_PER_CELL / (ms_chunked / args.steps / 1e3) / 1e9 / peak,
                                               launches_per_step=len(chunks))),
                    clocks=clocks_summary(samples), verified="64 sampled trios equal a torch restatement" if world == 1 else None)
        print(json.dumps(line))


def load_peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        return {}


# --------------------------------------------------------------------------------------------------
# correction pipeline (c1 .. c4)
# --------------------------------------------------------------------------------------------------
def run_cuda(args):
    import torch
    import torch.distributed as dist
    from genofix_b200 import _lib
    from genofix_b200.chunking import chunks_for_rank
    from genofix_b200.engine import Engine, MultiLane, pinned_array
    from genofix_b200.gfutils.packedgens import pack_codes
    from genofix_b200.pedigree.pedigree_dag import PedigreeDAG
    cfg = args.cfg
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if cfg.get("trio_scan"):
        run_trio_scan(args, torch, dist, world, rank, local)
        if world > 1:
            dist.destroy_process_group()
        return
    L = _lib.lib()
    dev = torch.device("cuda", local)
    ped_rows, drop_rows = make_pedigree(cfg)
    ids = drop_rows[:, 0]
    n = len(ids)
    sharded = bool(cfg.get("sharded"))
    chunk = cfg["chunk"]
    n_chunks_total = (cfg["snps"] + chunk - 1) // chunk
    ped = PedigreeDAG.from_table(ped_rows)
    ml = MultiLane(ped, ids, BACK, lanes=max(1, args.lanes))
    eng = ml.engines[0]
    lanes = len(ml.engines)
    lane_streams = [torch.cuda.Stream() for _ in range(lanes)]
    # ---- inputs: this rank's chunks, packed and resident in HBM; pinned host copies for the end-to-end leg ----
    packed, widths, host_chunks = [], [], []
    if sharded:
        # one fixed problem: global chunk g belongs to rank g % world; its genotypes are dropped on the device from a seed
        # that depends on g only, so every number of ranks corrects the same 100k x 600k matrix
        from genofix_b200.simulate_device import DeviceSimulator
        mine = chunks_for_rank(n_chunks_total, rank, world)
        for g in mine:
            w = min(chunk, cfg["snps"] - g * chunk)
            sim = DeviceSimulator(drop_rows, w, seed=1234 + 7919 * g)
            packed.append(sim.inject_errors(sim.gene_drop(), 0.01))
            widths.append(w)
            del sim
        e2e_idx = list(range(min(len(packed), args.e2e_chunks)))
    else:
        # weak scaling over SNP blocks: rank r owns columns [r * S, (r + 1) * S) of one (world * S)-SNP matrix
        obs = host_matrix(cfg, drop_rows, cfg["snps"], seed=1234 + rank)
        mine = list(range(n_chunks_total))
        for c in mine:
            a, b = c * chunk, min(cfg["snps"], (c + 1) * chunk)
            sub = pinned_array((n, b - a))       # e2e inputs: pinned host memory (bench contract)
            sub[...] = obs[:, a:b]
            host_chunks.append(sub)
            eng._alloc(b - a)
            hin = eng.host_in.numpy()
            hin[:, :eng.s] = sub
            hin[:, eng.s:] = 9
            eng.codes.copy_(eng.host_in)
            p = torch.empty((eng.n, eng.wpr), dtype=torch.int32, device=dev)
            _lib.check(L.gfx_pack2(eng.codes.data_ptr(), eng.n, eng.s, eng.codes.stride(0), p.data_ptr(), eng.wpr, eng._stream()))
            packed.append(p)
            widths.append(b - a)
        e2e_idx = list(range(len(packed)))
    torch.cuda.synchronize()
    cells_rank = n * int(sum(widths))
    tall = torch.zeros(n, dtype=torch.int32, device=dev)
    lane_tall = [torch.zeros(n, dtype=torch.int32, device=dev) for _ in range(lanes)]
    rank_events, phase_events = [], []

    def step_resident(record, keep=None):
        # chunk k on lane k mod L: one host thread + one stream per lane (chunks are independent)
        cur = torch.cuda.current_stream()

        def lane(l, e):
            st = lane_streams[l]
            st.wait_stream(cur)
            with torch.cuda.stream(st):
                lane_tall[l].zero_()
                for ci in range(l, len(packed), lanes):
                    e.use_packed(packed[ci], widths[ci])
                    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)] if record else None
                    if record:
                        ev[0].record()
                    e.mendel_rank_device()
                    if record:
                        ev[1].record()
                    e.thresholds_device(PARAMS["init_filter_p"])
                    if record:
                        ev[2].record()
                    e.correct_device(PARAMS["threshold_pair"], PARAMS["threshold_single"], PARAMS["tiethreshold"], PARAMS["err_thresh"])
                    if record:
                        ev[3].record()
                        rank_events.append((ev[0], ev[1], widths[ci]))
                        phase_events.append(("thresholds_host_roundtrip", ev[1], ev[2]))
                        phase_events.append(("correct_pairs_singles", ev[2], ev[3]))
                    lane_tall[l].add_(e.tallies[:e.n])
                    if keep is not None and ci == keep[0]:
                        keep[1].copy_(e.packed_live)
        ml._run_lanes(lane)
        for st in lane_streams:
            cur.wait_stream(st)
        tall.zero_()
        for x in lane_tall:
            tall.add_(x)
        if world > 1:
            dist.all_reduce(tall)   # the one collective: per-animal error tallies

    # end-to-end: the public multi-chunk call (genofix.py's chunk loop) on pinned host buffers, every step
    if sharded:
        packed_chunks = []
        for i in e2e_idx:
            pin = pinned_array(tuple(packed[i].shape), np.uint32)
            pin.view(np.int32)[...] = packed[i].cpu().numpy()
            packed_chunks.append((pin, widths[i]))
        packed_outs = [pinned_array(p.shape, np.uint32) for p, _ in packed_chunks]
        host_outs = None
    else:
        host_outs = [pinned_array(c.shape) for c in host_chunks]   # caller-owned, pinned result matrices
        packed_chunks, packed_outs = [], []
        for c in host_chunks:
            pc = pack_codes(c)
            pin = pinned_array(pc.shape, np.uint32)
            pin[...] = pc
            packed_chunks.append((pin, c.shape[1]))
            packed_outs.append(pinned_array(pc.shape, np.uint32))

    def step_e2e(packed_io):
        total = 0
        src = packed_chunks if packed_io else host_chunks
        for out, stats in ml.correct_chunks(iter(src), PARAMS["threshold_pair"], PARAMS["threshold_single"],
                                            init_filter_p=PARAMS["init_filter_p"], tiethreshold=PARAMS["tiethreshold"],
                                            err_thresh=PARAMS["err_thresh"], outs=packed_outs if packed_io else host_outs,
                                            packed_io=packed_io):
            total += int(stats["corrected_per_animal"].sum())
        if world > 1:
            dist.all_reduce(tall)
        return total

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    small = n * min(chunk, cfg["snps"]) * 2.25 <= 2 * 126e6
    flush_buf = torch.empty(64 << 20, dtype=torch.int32, device=dev) if small else None     # 256 MB > L2
    for _ in range(max(args.warmup, 3)):
        step_resident(False)
    barrier()
    stop, samples = threading.Event(), []
    th = threading.Thread(target=clocks_sampler, args=(stop, samples, local), daemon=True)
    th.start()
    l0 = L.gfx_launch_count()
    t0 = torch.cuda.Event(enable_timing=True)
    t1 = torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(args.steps):
        if small:
            flush_buf.zero_()
        step_resident(True)
    t1.record()
    barrier()
    launches = int(L.gfx_launch_count() - l0)
    ms = t0.elapsed_time(t1)
    stop.set()
    th.join(timeout=2)
    corrected_total = int(tall.sum().item())
    # ---- outside the timed region: the step's output of chunk 0 equals the public single-chunk call on the same data ----
    kept = torch.empty_like(packed[0])
    step_resident(False, keep=(0, kept))
    torch.cuda.synchronize()
    chk = Engine(ped, ids, BACK, schedule=ml.schedule)
    chk._alloc(widths[0])
    out_codes = torch.empty((n, chk.ldc), dtype=torch.uint8, device=dev)
    _lib.check(L.gfx_unpack2(kept.data_ptr(), n, widths[0], chk.wpr, out_codes.data_ptr(), out_codes.stride(0), chk._stream()))
    _lib.check(L.gfx_unpack2(packed[0].data_ptr(), n, widths[0], chk.wpr, chk.codes.data_ptr(), chk.codes.stride(0), chk._stream()))
    torch.cuda.synchronize()
    step_out = out_codes[:, :widths[0]].cpu().numpy()
    inp0 = chk.codes[:, :widths[0]].cpu().numpy()
    api_out = chk.correct(inp0, PARAMS["threshold_pair"], PARAMS["threshold_single"], init_filter_p=PARAMS["init_filter_p"],
                          tiethreshold=PARAMS["tiethreshold"], err_thresh=PARAMS["err_thresh"])
    if not np.array_equal(step_out, api_out):
        raise SystemExit("bench.py: the timed step's result of chunk 0 differs from Engine.correct() on the same chunk")
    verified = "chunk 0 of the timed step equals Engine.correct() on the same input (%d cells changed)" % int((api_out != inp0).sum())
    del chk, out_codes
    # ---- e2e: host buffers in, host buffers out, every step ----
    e2e_cells_rank = n * int(sum(widths[i] for i in e2e_idx))
    e2e_ms = e2e_packed_ms = float("nan")
    if not sharded:
        for _ in range(2):
            step_e2e(False)
        barrier()
        w0 = time.perf_counter()
        for _ in range(args.steps):
            step_e2e(False)
        barrier()
        e2e_ms = (time.perf_counter() - w0) * 1e3
    step_e2e(True)
    barrier()
    w0 = time.perf_counter()
    for _ in range(args.steps):
        step_e2e(True)
    barrier()
    e2e_packed_ms = (time.perf_counter() - w0) * 1e3
    if sharded:
        e2e_ms = e2e_packed_ms
    tm = torch.tensor([ms, e2e_ms, e2e_packed_ms, float(cells_rank), float(e2e_cells_rank)], dtype=torch.float64, device=dev)
    per_rank_ms = None
    if world > 1:
        gathered = [torch.zeros_like(tm) for _ in range(world)]
        dist.all_gather(gathered, tm)
        per_rank_ms = [float(g[0]) / args.steps for g in gathered]
        cells_step = float(sum(g[3] for g in gathered))
        e2e_cells_step = float(sum(g[4] for g in gathered))
        tm = torch.stack(gathered).max(dim=0).values
    else:
        cells_step, e2e_cells_step = float(cells_rank), float(e2e_cells_rank)
    ms, e2e_ms, e2e_packed_ms = float(tm[0]), float(tm[1]), float(tm[2])
    value = cells_step * args.steps / (ms / 1e3)
    e2e_value = e2e_cells_step * args.steps / (e2e_ms / 1e3)
    # roofline of the HBM-bound peeling kernel: timed alone on the bench inputs right after the timed region (inside the
    # step it shares the SMs with the other lanes' correction kernels, which says nothing about the kernel itself)
    eng.use_packed(packed[0], widths[0])
    for _ in range(3):
        eng.mendel_rank_device()
    torch.cuda.synchronize()
    iso = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(10)]
    flush = torch.empty(64 << 20, dtype=torch.int32, device=dev)     # 256 MB > L2: every timed launch starts cold
    # the library's own pair of events sits right around the dominant kernel (k_mendel_rank3); the pair here is around the
    # whole gfx_mendel_rank_hist call = histogram memset + that kernel + the consumer of the loopy rows
    import ctypes as _C
    _lib.check(eng.L.gfx_rank_timing(eng.schedule.handle, 1))
    kern_ms = []
    for e0, e1 in iso:
        flush.zero_()
        e0.record()
        eng.mendel_rank_device()
        e1.record()
        kms = _C.c_float(0.0)
        _lib.check(eng.L.gfx_rank_last_kernel_ms(eng.schedule.handle, _C.byref(kms)))
        kern_ms.append(float(kms.value))
    torch.cuda.synchronize()
    _lib.check(eng.L.gfx_rank_timing(eng.schedule.handle, 0))
    call_ms = float(np.mean([e0.elapsed_time(e1) for e0, e1 in iso]))
    iso_ms = float(np.mean(kern_ms))
    iso_ach = n * widths[0] * RANK_BYTES_PER_CELL / (iso_ms / 1e3) / 1e9
    rk_ms = [e0.elapsed_time(e1) for e0, e1, _ in rank_events]
    rk_cells = [n * w for _, _, w in rank_events]
    peaks = load_peaks()
    peak = float(peaks.get("hbm_gbs", 6650.0))
    ach = sum(rk_cells) * RANK_BYTES_PER_CELL / (sum(rk_ms) / 1e3) / 1e9
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "rank_kernel_traffic.json"))).get("dram_bytes_per_launch")
    except Exception:
        pass
    phases = {}
    for name, e0, e1 in phase_events:
        phases[name] = phases.get(name, 0.0) + e0.elapsed_time(e1)
    phases["mendel_rank"] = sum(rk_ms)
    phase_share = {k: v / ms / lanes for k, v in phases.items()}   # lanes run concurrently: share of lane time
    kernel_block = dict(bound="hbm", achieved=iso_ach, peak=peak, unit="GB/s", frac=iso_ach / peak, traffic=traffic,
                        name=("k_mendel_rank<HIST> (second form: the engine picks it for matrices with more than 0.5 % missing calls)"
                              if getattr(eng.schedule, "_rank_form", 0) == 2 else "k_mendel_rank3<HIST>") +
                             " (Mendel rank fused with the score histogram): the HBM-bound peeling kernel BASELINE.json's "
                             "north_star names",
                        timed="alone: 10 launches on the bench inputs right after the timed region, L2 flushed (256 MB written) "
                              "before each (burst peak applies); CUDA events recorded by the library on the launching stream "
                              "right around this kernel (gfx_rank_timing); us_per_call is the whole gfx_mendel_rank_hist call "
                              "(histogram memset + this kernel + the loopy-row consumer k_mendel_rank_generic[_wl])",
                        us_per_launch=iso_ms * 1e3, us_per_call=call_ms * 1e3, achieved_per_call=n * widths[0] * RANK_BYTES_PER_CELL /
                        (call_ms / 1e3) / 1e9, algorithmic_bytes_per_cell=RANK_BYTES_PER_CELL,
                        cells_per_launch=n * widths[0], achieved_in_step=ach, frac_in_step=ach / peak,
                        share_of_step=sum(rk_ms) / ms / lanes)
    pipe_gbs = cells_rank * args.steps * PIPELINE_BYTES_PER_CELL / (ms / 1e3) / 1e9
    roofline = dict(kernel_block)
    roofline.update(
        kernel=kernel_block,
        pipeline=dict(bytes_per_cell=PIPELINE_BYTES_PER_CELL, achieved_gbs=pipe_gbs, frac=pipe_gbs / peak, per="GPU",
                      note="whole step against the HBM roofline of rank + correction with the float16 rank matrix materialised: "
                           "the correction kernels are instruction / latency bound (exact inference on pedigree blankets), "
                           "not bandwidth bound"),
        peak_source="MEASURED_PEAKS.json" if peaks else "fallback 6.65 TB/s",
        correct_kernels=dict(share_of_step=phase_share.get("correct_pairs_singles"), algorithmic_bytes_per_cell=2.5,
                             achieved_gbs=cells_rank * args.steps * 2.5 /
                             (phases.get("correct_pairs_singles", float("nan")) / lanes / 1e3) / 1e9))
    if rank == 0:
        cb = None
        if world == 1 and not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            want = 16 * max(13, cores)
            sample = inp0[:, :min(widths[0], want)]          # the bench matrix's own first columns (chunk 0)
            cb, _ = cpu_baseline(ped_rows, ids, np.ascontiguousarray(sample), args.cpu_seconds)
        e2e = dict(value=e2e_value, unit="cells/s", ms_per_step=e2e_ms / args.steps)
        if sharded:
            io_bytes = int(sum(p.nbytes for p, _ in packed_chunks)) * world
            e2e.update(h2d_bytes_per_step=io_bytes, d2h_bytes_per_step=io_bytes,
                       sample="the first %d chunks of every rank through correct_chunks(packed_io=True): 2-bit blocks in pinned "
                              "host memory in and out" % len(e2e_idx))
        else:
            e2e.update(h2d_bytes_per_step=int(cells_rank) * world, d2h_bytes_per_step=int(cells_rank) * world)
        line = dict(metric=METRIC, value=value, unit="cells/s", n_gpus=world, steps=args.steps, warmup=max(args.warmup, 3),
                    ms_per_step=ms / args.steps, higher_is_better=True, scaling="strong" if sharded else "weak", vs_baseline=None,
                    dtype="u8/f16/f64 (2-bit codes, float16 scores, integer + float64 posteriors)", data="synthetic",
                    config=config_block(args, world, lanes, dict(
                        per_gpu_matrix="%d x %d" % (n, int(sum(widths))),
                        sharding=("%d chunks of the one %d x %d matrix dealt round-robin (chunk g -> rank g %% %d)"
                                  % (n_chunks_total, n, cfg["snps"], world)) if sharded else
                        "rank r owns SNP columns [r*S, (r+1)*S) of one %d x %d matrix, S = %d" % (n, cfg["snps"] * world, cfg["snps"]))),
                    e2e=e2e, gpu_launches=launches, roofline=roofline, phase_share=phase_share, clocks=clocks_summary(samples),
                    corrected_cells_per_step=corrected_total, verified=verified)
        if not sharded:
            line["e2e_packed_io"] = dict(value=e2e_cells_step * args.steps / (e2e_packed_ms / 1e3), unit="cells/s",
                                         ms_per_step=e2e_packed_ms / args.steps,
                                         h2d_bytes_per_step=int(sum(p.nbytes for p, _ in packed_chunks)) * world,
                                         d2h_bytes_per_step=int(sum(p.nbytes for p in packed_outs)) * world,
                                         note="same call with packed_io=True: 2-bit blocks as stored by gfutils.packedgens (GFX2BIT1)")
        if per_rank_ms is not None:
            line["per_rank_ms_per_step"] = per_rank_ms
            line["rank_imbalance"] = max(per_rank_ms) / (sum(per_rank_ms) / len(per_rank_ms))
        if cb is not None:
            line["cpu_baseline"] = cb
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    # stdout carries exactly one JSON line: everything libraries write to fd 1 (NCCL prints its version there) goes to stderr
    sys.stdout.flush()
    _json_fd = os.dup(1)
    os.dup2(2, 1)
    _json_out = os.fdopen(_json_fd, "w")
    _print = print

    def print(*args, **kw):  # noqa: A001 -- the JSON line is the only print of this module
        kw.setdefault("file", _json_out)
        _print(*args, **kw)
        _json_out.flush()
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_cuda(a)
