"""Multi-GPU driver: whole SNP chunks are dealt round-robin to ranks (one process per GPU), the pedigree
schedule is replicated, and the single collective is an all-reduce of the per-animal tallies
(SURVEY section 8(e)).  Every chunk-level scalar (rank maximum, quantile threshold) is chunk-local, so the
corrected matrix does not depend on the number of ranks.

The compute callback is injected so that the same logic runs on NCCL with the CUDA engine and, in the CPU
tests, on gloo with a stand-in.
"""
import os

import numpy as np

from .chunking import chunk, chunks_for_rank


def init_from_env(backend=None):
    """torchrun-style rendezvous (RANK / WORLD_SIZE / MASTER_ADDR / MASTER_PORT).  Returns (rank, world)."""
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            local = int(os.environ.get("LOCAL_RANK", "0"))
            torch.cuda.set_device(local)
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(backend)
    return rank, world


def allreduce_tallies(tallies):
    """Sum of the per-animal tallies over ranks (in place; tensor on the backend's device)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(tallies)
    return tallies


def correct_sharded(n_snps, chunk_snps, correct_chunk, n_animals, rank=0, world=1, overlap=0, device="cpu"):
    """Run `correct_chunk(col_indices) -> (corrected uint8 [n, len(cols)], tallies int [n])` on this rank's
    chunks.  Returns ({chunk index: (cols, corrected)}, all-reduced tallies tensor)."""
    import torch
    cols = list(range(n_snps))
    chunks = chunk(cols, chunk_snps, overlap, overlap) if overlap else \
        [cols[i:i + chunk_snps] for i in range(0, n_snps, chunk_snps)]
    mine = chunks_for_rank(len(chunks), rank, world)
    out = {}
    tall = torch.zeros(n_animals, dtype=torch.int64, device=device)
    for ci in mine:
        corrected, t = correct_chunk(chunks[ci])
        out[ci] = (chunks[ci], corrected)
        tall += torch.as_tensor(np.asarray(t, dtype=np.int64), device=device)
    allreduce_tallies(tall)
    return out, tall


def mendel_tally_sharded(n_chunks, hist_of_chunk, rowsum_of_chunk, table_from_hist, n_animals, rank=0, world=1, device="cpu"):
    """Per-animal Mendel tally (error_checker.py:193-204) of a matrix whose SNP chunks are dealt round-robin to ranks.
    Unlike the correction path, the rank transform here is normalised by the maximum over the WHOLE matrix, so the
    chunk-local state is exchanged twice ("if chunks had to be split"):
      1. all-reduce (sum) of the 65 537-bin histogram of score patterns -> every rank builds the same transform table
         (`table_from_hist(hist)`, exact: the table depends on the set of patterns present only);
      2. all-reduce (sum) of the float64 partial row sums (`rowsum_of_chunk(ci, table)` -> float64 [n_animals]).
    The float16 accumulation chain of the single-GPU form (`gfx_rank_rowsum_f16`) is order dependent and cannot be
    sharded; the sharded tally is the float64 sum, identical on every rank and independent of the number of ranks up to
    the rounding of a sum of `world` partial sums (compare to 1e-12 relative).
    Returns (global histogram int64 [65537] numpy, row sums float64 [n_animals] numpy)."""
    import torch
    mine = chunks_for_rank(n_chunks, rank, world)
    hist = torch.zeros(65537, dtype=torch.int64, device=device)
    for ci in mine:
        hist += torch.as_tensor(np.asarray(hist_of_chunk(ci)).astype(np.int64), device=device)
    allreduce_tallies(hist)
    hist_np = hist.cpu().numpy()
    table = table_from_hist(hist_np)
    sums = torch.zeros(n_animals, dtype=torch.float64, device=device)
    for ci in mine:
        sums += torch.as_tensor(np.asarray(rowsum_of_chunk(ci, table), dtype=np.float64), device=device)
    allreduce_tallies(sums)
    return hist_np, sums.cpu().numpy()
