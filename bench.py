#!/usr/bin/env python
"""Benchmark of the GenoFix hot path on B200 (contract: see the task statement / DESIGN.md section 6).

  python bench.py --gpus N --steps K --warmup W            # CUDA path (this repo)
  python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU algorithm (oracle port) on host cores

Workload (BASELINE.json configs[1]): synthetic 10k-animal 6-generation pedigree x 50k SNPs, 1 % injected
errors, processed like genofix.py does: chunks of 10 000 SNPs (genofix.py:111), each chunk through
correctMatrix (rank -> transform/quantile -> per-generation pair correction -> singles).
A step = one pass over the whole matrix of one GPU.  Multi-GPU: every rank owns its own matrix of the
same shape (weak scaling; SNP blocks are independent given the replicated pedigree schedule) and one
NCCL all-reduce of the per-animal tallies closes the step.
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

N_ANIMALS, N_GENERATIONS, N_SNPS, CHUNK = 10000, 6, 50000, 10000
PARAMS = dict(threshold_pair=0.5, threshold_single=0.64, init_filter_p=0.95, tiethreshold=0.4, err_thresh=1.0)  # genofix.py:95-112
BACK = 2
RANK_BYTES_PER_CELL = 2.25   # 0.25 B packed code read + 2 B float16 score written


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--animals", type=int, default=N_ANIMALS)
    ap.add_argument("--snps", type=int, default=N_SNPS)
    ap.add_argument("--chunk", type=int, default=CHUNK)
    ap.add_argument("--generations", type=int, default=N_GENERATIONS)
    ap.add_argument("--lanes", type=int, default=3, help="chunks corrected concurrently per GPU (streams + host threads)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=20.0, help="target CPU time of the baseline sample")
    return ap.parse_args()


def make_inputs(n_animals, n_snps, seed, generations=N_GENERATIONS):
    from genofix_b200 import synth
    rows = synth.make_pedigree(n_animals, generations, sire_frac=0.05, dam_frac=0.5, seed=1234)
    truth = synth.gene_drop(rows, n_snps, seed=seed)
    obs = synth.inject_errors(truth, 0.01, seed=seed)
    return rows, obs


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


def cpu_baseline(rows, obs, seconds, cores=None):
    """Times the oracle port on `cores` processes, each correcting its own block of SNP columns of the
    same matrix (full pedigree).  Sample size is scaled to ~`seconds` of wall time from a 1-SNP probe."""
    import multiprocessing as mp
    cores = cores or os.cpu_count() or 1
    ids = rows[:, 0]
    t0 = time.time()
    _cpu_worker((rows, ids, np.ascontiguousarray(obs[:, :1])))
    probe = time.time() - t0
    per_proc = int(max(1, min(obs.shape[1] // cores, round(seconds / max(probe, 1e-3)))))
    blocks = [np.ascontiguousarray(obs[:, i * per_proc:(i + 1) * per_proc]) for i in range(cores)]
    ctx = mp.get_context("fork")
    t0 = time.time()
    with ctx.Pool(cores) as pool:
        cells = sum(pool.map(_cpu_worker, [(rows, ids, b) for b in blocks]))
    dt = time.time() - t0
    return dict(value=cells / dt, unit="cells/s", cores=cores, kind="port",
                sample="%d animals x %d SNPs per process x %d processes (full pedigree, %d-SNP blocks of the bench matrix), %.1f s"
                       % (obs.shape[0], per_proc, cores, per_proc, dt))


# --------------------------------------------------------------------------------------------------
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


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    rows, obs = make_inputs(args.animals, max(64, min(args.snps, 512)), seed=1234)
    vals = []
    cb = None
    for i in range(args.warmup + args.steps):
        cb = cpu_baseline(rows, obs, args.cpu_seconds / max(1, args.steps))
        if i >= args.warmup:
            vals.append(cb["value"])
    v = float(np.mean(vals)) if vals else 0.0
    cb["value"] = v
    line = dict(metric="individual x SNP genotype cells corrected/sec", value=v, unit="cells/s", n_gpus=args.gpus,
                steps=args.steps, warmup=args.warmup, ms_per_step=None, higher_is_better=True, scaling="weak",
                vs_baseline=None, dtype="f64", data="synthetic", impl="reference",
                config=dict(workload="10k-animal 6-generation synthetic pedigree x 50k SNPs, 1% errors (BASELINE configs[1]); "
                                     "CPU arm times a bounded SNP sample of it", animals=args.animals, snps=args.snps,
                            chunk_snps=args.chunk, back=BACK, **PARAMS),
                cpu_baseline=cb, e2e=dict(value=v, unit="cells/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    print(json.dumps(line))


def run_cuda(args):
    import torch
    import torch.distributed as dist
    from genofix_b200 import _lib
    from genofix_b200.engine import MultiLane, pinned_array
    from genofix_b200.pedigree.pedigree_dag import PedigreeDAG
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    L = _lib.lib()
    rows, obs = make_inputs(args.animals, args.snps, seed=1234 + rank, generations=args.generations)
    ped = PedigreeDAG.from_table(rows)
    n_chunks = (args.snps + args.chunk - 1) // args.chunk
    bounds = [(c * args.chunk, min(args.snps, (c + 1) * args.chunk)) for c in range(n_chunks)]
    ml = MultiLane(ped, rows[:, 0], BACK, lanes=max(1, args.lanes))
    eng = ml.engines[0]
    lanes = len(ml.engines)
    lane_streams = [torch.cuda.Stream() for _ in range(lanes)]
    dev = torch.device("cuda", local)
    # resident inputs: packed chunks in HBM
    packed = []
    host_chunks = []
    for (a, b) in bounds:
        eng._alloc(b - a)
        sub = pinned_array((obs.shape[0], b - a))       # e2e inputs: pinned host memory (bench contract)
        sub[...] = obs[:, a:b]
        host_chunks.append(sub)
        hin = eng.host_in.numpy()
        hin[:, :eng.s] = sub
        hin[:, eng.s:] = 9
        eng.codes.copy_(eng.host_in)
        p = torch.empty((eng.n, eng.wpr), dtype=torch.int32, device=dev)
        _lib.check(L.gfx_pack2(eng.codes.data_ptr(), eng.n, eng.s, eng.codes.stride(0), p.data_ptr(), eng.wpr, eng._stream()))
        packed.append(p)
    torch.cuda.synchronize()
    tall = torch.zeros(args.animals, dtype=torch.int32, device=dev)
    rank_events = []
    phase_events = []   # (name, start, end) CUDA events on the launching stream

    lane_tall = [torch.zeros(args.animals, dtype=torch.int32, device=dev) for _ in range(lanes)]

    def step_resident(record):
        # chunk k on lane k mod L: one host thread + one stream per lane (chunks are independent)
        cur = torch.cuda.current_stream()

        def lane(l, e):
            st = lane_streams[l]
            st.wait_stream(cur)
            with torch.cuda.stream(st):
                lane_tall[l].zero_()
                for ci in range(l, len(bounds), lanes):
                    a, b = bounds[ci]
                    e.use_packed(packed[ci], b - a)
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
                        rank_events.append((ev[0], ev[1], (b - a)))
                        phase_events.append(("thresholds_host_roundtrip", ev[1], ev[2]))
                        phase_events.append(("correct_pairs_singles", ev[2], ev[3]))
                    lane_tall[l].add_(e.tallies[:e.n])
        ml._run_lanes(lane)
        for st in lane_streams:
            cur.wait_stream(st)
        tall.zero_()
        for x in lane_tall:
            tall.add_(x)
        if world > 1:
            dist.all_reduce(tall)   # the one collective: per-animal error tallies

    host_outs = [pinned_array(c.shape) for c in host_chunks]   # caller-owned, pinned result matrices

    def step_e2e():
        # the public multi-chunk call (genofix.py's chunk loop): host uint8 in, host uint8 out, every step
        total = 0
        for out, stats in ml.correct_chunks(iter(host_chunks), PARAMS["threshold_pair"], PARAMS["threshold_single"],
                                             init_filter_p=PARAMS["init_filter_p"], tiethreshold=PARAMS["tiethreshold"],
                                             err_thresh=PARAMS["err_thresh"], outs=host_outs):
            total += int(stats["corrected_per_animal"].sum())
        return total

    # the same step fed with 2-bit packed blocks (what gfutils.packedgens.PackedGens streams from disk): extra key
    from genofix_b200.gfutils.packedgens import pack_codes
    packed_chunks, packed_outs = [], []
    for c in host_chunks:
        pc = pack_codes(c)
        pin = pinned_array(pc.shape, np.uint32)
        pin[...] = pc
        packed_chunks.append((pin, c.shape[1]))
        packed_outs.append(pinned_array(pc.shape, np.uint32))

    def step_e2e_packed():
        total = 0
        for out, stats in ml.correct_chunks(iter(packed_chunks), PARAMS["threshold_pair"], PARAMS["threshold_single"],
                                            init_filter_p=PARAMS["init_filter_p"], tiethreshold=PARAMS["tiethreshold"],
                                            err_thresh=PARAMS["err_thresh"], outs=packed_outs, packed_io=True):
            total += int(stats["corrected_per_animal"].sum())
        return total

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

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
        step_resident(True)
    t1.record()
    barrier()
    launches = int(L.gfx_launch_count() - l0)
    ms = t0.elapsed_time(t1)
    stop.set()
    th.join(timeout=2)
    corrected_total = int(tall.sum().item())
    # e2e: host buffers in, host buffers out, every step
    for _ in range(2):
        step_e2e()
    barrier()
    w0 = time.perf_counter()
    for _ in range(args.steps):
        step_e2e()
    barrier()
    e2e_s = time.perf_counter() - w0
    step_e2e_packed()
    barrier()
    w0 = time.perf_counter()
    for _ in range(args.steps):
        step_e2e_packed()
    barrier()
    e2e_packed_s = time.perf_counter() - w0
    tm = torch.tensor([ms, e2e_s * 1e3, e2e_packed_s * 1e3], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
    ms, e2e_ms, e2e_packed_ms = float(tm[0]), float(tm[1]), float(tm[2])
    cells_step = args.animals * args.snps * world
    value = cells_step * args.steps / (ms / 1e3)
    e2e_value = cells_step * args.steps / (e2e_ms / 1e3)
    # roofline of the dominant kernel (Mendel rank), CUDA events on the launching stream
    # the rank kernel alone (same inputs, GPU otherwise idle): inside the step it shares the SMs with the other lane's
    # correction kernels, which says nothing about the kernel itself
    eng.use_packed(packed[0], bounds[0][1] - bounds[0][0])
    for _ in range(3):
        eng.mendel_rank_device()
    torch.cuda.synchronize()
    iso = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(10)]
    for e0, e1 in iso:
        e0.record()
        eng.mendel_rank_device()
        e1.record()
    torch.cuda.synchronize()
    iso_ms = float(np.mean([e0.elapsed_time(e1) for e0, e1 in iso]))
    iso_ach = args.animals * (bounds[0][1] - bounds[0][0]) * RANK_BYTES_PER_CELL / (iso_ms / 1e3) / 1e9
    rk_ms = [e0.elapsed_time(e1) for e0, e1, _ in rank_events]
    rk_cells = [args.animals * w for _, _, w in rank_events]
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
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
    roofline = dict(bound="hbm", achieved=iso_ach, peak=peak, unit="GB/s", frac=iso_ach / peak, traffic=traffic,
                    timed="alone: 10 launches on the bench inputs right after the timed region (burst peak applies)",
                    us_per_launch=iso_ms * 1e3, achieved_in_step=ach, frac_in_step=ach / peak,
                    in_step_note="CUDA events around the same call inside the timed region, where it overlaps the other lane's kernels",
                    kernel="k_mendel_rank<HIST> (Mendel rank fused with the score histogram; + k_mendel_rank_generic for loopy "
                           "blankets): the HBM-bound peeling kernel BASELINE.json's north_star names",
                    algorithmic_bytes_per_cell=RANK_BYTES_PER_CELL,
                    peak_source="MEASURED_PEAKS.json" if peaks else "fallback 6.65 TB/s",
                    share_of_step=sum(rk_ms) / ms / lanes,
                    note="the step is dominated by the correction kernels (k_correct, k_correct_jobs), which are "
                         "instruction/latency bound: phase shares in phase_share, launch list in profiles/",
                    correct_kernels=dict(share_of_step=phase_share.get("correct_pairs_singles"),
                                         algorithmic_bytes_per_cell=2.5,
                                         achieved_gbs=args.animals * args.snps * args.steps * 2.5 /
                                         (phases.get("correct_pairs_singles", float("nan")) / 1e3) / 1e9))
    sm = [float(s[0]) for s in samples if s[0].replace(".", "").isdigit()]
    reasons = []
    for name, idx in (("hw_slowdown", 2), ("hw_thermal_slowdown", 3), ("sw_thermal_slowdown", 4), ("sw_power_cap", 5)):
        if any(s[idx].lower().startswith("active") for s in samples):
            reasons.append(name)
    clocks = dict(sm_mhz=float(np.median(sm)) if sm else None,
                  sm_max_mhz=float(samples[0][1]) if samples else None, reasons=reasons, samples=len(samples))
    if rank == 0:
        cb = None
        if world == 1 and not args.no_cpu_baseline:
            cb = cpu_baseline(rows, obs, args.cpu_seconds)
        line = dict(metric="individual x SNP genotype cells corrected/sec", value=value, unit="cells/s", n_gpus=world,
                    steps=args.steps, warmup=max(args.warmup, 3), ms_per_step=ms / args.steps, higher_is_better=True,
                    scaling="weak", vs_baseline=None, dtype="u8/f16/f64 (2-bit codes, float16 scores, float64 posteriors)",
                    data="synthetic",
                    config=dict(workload=("10k-animal 6-generation synthetic pedigree x 50k SNPs, 1% errors (BASELINE configs[1]), "
                                          "corrected in 10k-SNP chunks like genofix.py") if (args.animals, args.snps, args.generations) ==
                                (N_ANIMALS, N_SNPS, N_GENERATIONS) else "%d-animal %d-generation synthetic pedigree x %d SNPs, 1%% errors"
                                % (args.animals, args.generations, args.snps), animals=args.animals, snps=args.snps,
                                chunk_snps=args.chunk, back=BACK, lanes=lanes, per_gpu_matrix="%d x %d" % (args.animals, args.snps),
                                l2="each step streams %.0f MB of packed codes + %.0f MB of float16 scores per chunk (> 126 MB L2)"
                                   % (args.animals * args.chunk / 4 / 1e6, args.animals * args.chunk * 2 / 1e6), **PARAMS),
                    e2e=dict(value=e2e_value, unit="cells/s", h2d_bytes_per_step=int(args.animals * args.snps) * world,
                             d2h_bytes_per_step=int(args.animals * args.snps) * world, ms_per_step=e2e_ms / args.steps),
                    e2e_packed_io=dict(value=cells_step * args.steps / (e2e_packed_ms / 1e3), unit="cells/s",
                                       ms_per_step=e2e_packed_ms / args.steps,
                                       h2d_bytes_per_step=int(sum(p.nbytes for p, _ in packed_chunks)) * world,
                                       d2h_bytes_per_step=int(sum(p.nbytes for p in packed_outs)) * world,
                                       note="same call with packed_io=True: 2-bit blocks as stored by gfutils.packedgens (GFX2BIT1)"),
                    gpu_launches=launches, roofline=roofline, phase_share=phase_share, clocks=clocks,
                    corrected_cells_per_step=corrected_total)
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
