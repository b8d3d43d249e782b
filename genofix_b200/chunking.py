"""SNP chunking of the entry points (genofix.py:207-210, simulate.py:399-403) and the chunk -> GPU map."""


def chunk(l, n, b=0, m=2):
    """Consecutive chunks of n SNPs, each extended by b SNPs of overlap; the tail is merged into the
    previous chunk when fewer than m SNPs would remain after it, and chunks of <= m SNPs are dropped."""
    n = max(1, n)
    out = []
    for i in range(0, len(l), n):
        c = l[i:] if len(l) - (i + n + b) < m else l[i:i + n + b]
        if len(c) > m:
            out.append(c)
    return out


def chunks_for_rank(n_chunks, rank, world_size):
    """Whole reference chunks are dealt round-robin to GPUs, so every chunk-level scalar (max, quantile)
    stays GPU-local and the result does not depend on the number of GPUs."""
    return list(range(rank, n_chunks, world_size))
