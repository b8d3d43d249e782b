"""2-bit packed genotype container + reader.

The reference keeps genotypes as a space-separated text matrix and re-reads lines through `linecache`
(gfutils/gensrandaccess.py:10-50); at 100 k animals x 600 k SNPs that file is ~120 GB and parsing it dwarfs the
kernels.  This container stores exactly the layout the kernels read in HBM -- 16 codes per little-endian uint32,
code 3 = missing (the reference's 9), rows padded to a multiple of 4 words -- so a SNP block goes from the page cache
to the GPU with one strided copy and no parsing, and comes back the same way.  `PackedGens` keeps the
`GensCache(file, header).getMatrix(ids)` surface of the reference.

File layout (little endian):
  0   8 bytes  magic  b"GFX2BIT1"
  8   int64    n_animals
  16  int64    n_snps
  24  int64    wpr          words per row = 4 * ceil(ceil(n_snps / 16) / 4)
  32  int64    names_bytes  length of the SNP-name block
  40  int64[n_animals]      animal ids
  ..  names_bytes           SNP names, '\\n' separated, UTF-8, zero padded to a multiple of 16 bytes
  ..  uint32[n_animals, wpr] packed rows (offset is 16-byte aligned)
"""
import numpy as np

MAGIC = b"GFX2BIT1"


def words_per_row(n_snps):
    return ((int(n_snps) + 15) // 16 + 3) // 4 * 4


def pack_codes(G):
    """uint8 [n, s] with 0, 1, 2 (anything else = missing) -> uint32 [n, words_per_row(s)]; padding cells are 3."""
    G = np.asarray(G, dtype=np.uint8)
    n, s = G.shape
    wpr = words_per_row(s)
    full = np.full((n, wpr * 16), 3, dtype=np.uint8)
    full[:, :s] = np.where(G > 2, 3, G)
    full = full.reshape(n, wpr, 16).astype(np.uint32)
    shifts = (2 * np.arange(16, dtype=np.uint32)).reshape(1, 1, 16)
    return np.bitwise_or.reduce(full << shifts, axis=2).astype(np.uint32)


def unpack_codes(P, n_snps):
    """inverse of pack_codes: uint8 [n, n_snps] with 0, 1, 2, 9"""
    P = np.asarray(P, dtype=np.uint32)
    n = P.shape[0]
    shifts = (2 * np.arange(16, dtype=np.uint32)).reshape(1, 1, 16)
    codes = ((P[:, :, None] >> shifts) & np.uint32(3)).astype(np.uint8).reshape(n, -1)[:, :n_snps]
    return np.where(codes == 3, 9, codes).astype(np.uint8)


class PackedGens(object):
    """Reader (memory mapped).  Same surface as GensCache: `all_ids`, `snps`, `kid2index`, `getMatrix(ids)`;
    plus `packed_columns(a, b)` for the SNP blocks the pipeline streams."""

    def __init__(self, path, mode="r"):
        self.gensfile = path
        with open(path, "rb") as f:
            head = f.read(40)
        if head[:8] != MAGIC:
            raise ValueError("%s is not a GFX2BIT1 container" % path)
        n, s, wpr, nb = (int(x) for x in np.frombuffer(head[8:40], dtype="<i8"))
        if wpr != words_per_row(s):
            raise ValueError("corrupt header: wpr %d for %d snps" % (wpr, s))
        self.n_animals, self.n_snps, self.wpr = n, s, wpr
        off = 40
        self.all_ids = [int(x) for x in np.fromfile(path, dtype="<i8", count=n, offset=off)]
        off += 8 * n
        names = np.fromfile(path, dtype=np.uint8, count=nb, offset=off).tobytes().rstrip(b"\0").decode("utf-8")
        self.snps = names.split("\n") if names else []
        if len(self.snps) != s:
            raise ValueError("corrupt header: %d SNP names for %d snps" % (len(self.snps), s))
        off += nb
        self._offset = off
        self.kid2index = {k: i for i, k in enumerate(self.all_ids)}
        self.packed = np.memmap(path, dtype="<u4", mode=mode, offset=off, shape=(n, wpr))

    # ---- writers ---------------------------------------------------------------------------------------
    @staticmethod
    def write(path, ids, snps, G=None, packed=None):
        """Write a container from a uint8 matrix `G` [n, s] (or an already packed uint32 [n, wpr] matrix)."""
        ids = np.asarray(ids, dtype="<i8")
        snps = [str(x) for x in snps]
        if any("\n" in x for x in snps):
            raise ValueError("SNP names must not contain newlines")
        if packed is None:
            packed = pack_codes(G)
        packed = np.ascontiguousarray(packed, dtype="<u4")
        n, s = len(ids), len(snps)
        if packed.shape != (n, words_per_row(s)):
            raise ValueError("matrix shape does not match ids / snps")
        names = "\n".join(snps).encode("utf-8")
        pad = (-(40 + 8 * n + len(names))) % 16
        names += b"\0" * pad
        with open(path, "wb") as f:
            f.write(MAGIC)
            f.write(np.array([n, s, packed.shape[1], len(names)], dtype="<i8").tobytes())
            f.write(ids.tobytes())
            f.write(names)
            f.write(packed.tobytes())
        return path

    @staticmethod
    def from_text(textfile, path, header=True, delimiter=" "):
        """Convert the reference's text matrix (gfutils/gensrandaccess.py format) once."""
        import os
        import sys
        src = os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "src")
        if src not in sys.path:
            sys.path.insert(0, src)
        from gfutils.gensrandaccess import GensCache
        g = GensCache(textfile, header=header, delimiter=delimiter)
        M = g.getMatrix(g.all_ids)
        snps = g.snps if g.snps is not None else [str(i) for i in range(M.shape[1])]
        PackedGens.write(path, g.all_ids, snps, M)
        return PackedGens(path)

    # ---- readers ---------------------------------------------------------------------------------------
    def getMatrix(self, ids):
        """uint8 [len(ids), n_snps] with 0, 1, 2, 9 (ids that are not in the file are skipped, like the reference)"""
        idx = [self.kid2index[int(x)] for x in ids if int(x) in self.kid2index]
        return unpack_codes(np.asarray(self.packed[idx]), self.n_snps)

    def packed_columns(self, a, b, out=None):
        """Packed block of SNP columns [a, b) for all animals: uint32 [n, words_per_row(b - a)], ready for
        Engine.correct_chunks(..., packed_io=True).  Only the words of these columns are read from the file.  A start
        that is not a multiple of 16 is handled by shifting across word boundaries.
        `out` may be a pinned array (genofix_b200.engine.pinned_array(shape, np.uint32))."""
        if not (0 <= a < b <= self.n_snps):
            raise ValueError("need 0 <= a < b <= n_snps")
        s = b - a
        w0, nw, wpr = a // 16, (s + 15) // 16, words_per_row(s)
        if out is None:
            out = np.empty((self.n_animals, wpr), dtype=np.uint32)
        if out.shape != (self.n_animals, wpr) or out.dtype != np.uint32:
            raise ValueError("out must be uint32 [%d, %d]" % (self.n_animals, wpr))
        sh = 2 * (a % 16)
        if sh == 0:
            out[:, :nw] = self.packed[:, w0:w0 + nw]
        else:
            last = min(self.wpr, w0 + nw + 1)
            src = np.asarray(self.packed[:, w0:last], dtype=np.uint32)
            lo = src[:, :nw] >> np.uint32(sh)
            hi = np.zeros_like(lo)
            k = src.shape[1] - 1                     # words that have a successor inside the row
            hi[:, :min(nw, k)] = src[:, 1:1 + min(nw, k)] << np.uint32(32 - sh)
            out[:, :nw] = lo | hi
        out[:, nw:] = 0xFFFFFFFF
        if s % 16:   # cells of the last word beyond b belong to the next block: mark them missing (padding)
            out[:, nw - 1] |= np.uint32((0xFFFFFFFF << (2 * (s % 16))) & 0xFFFFFFFF)
        return out

    def store_columns(self, a, b, block):
        """Write a corrected packed block back (file opened with mode='r+'); padding cells are not stored."""
        if not (0 <= a < b <= self.n_snps):
            raise ValueError("need 0 <= a < b <= n_snps")
        s = b - a
        blk = np.asarray(block, dtype=np.uint32)
        if a % 16:
            # unaligned start: through the codes of the covering words (single writer only: neighbouring blocks share words)
            w0, w1 = a // 16, (b + 15) // 16
            codes = unpack_codes(np.asarray(self.packed[:, w0:w1]), (w1 - w0) * 16)
            codes[:, a - 16 * w0:b - 16 * w0] = unpack_codes(blk, s)
            self.packed[:, w0:w1] = pack_codes(codes)[:, :w1 - w0]
            return
        w0, nw = a // 16, (s + 15) // 16
        if s % 16 and b < self.n_snps:
            keep = np.uint32((1 << (2 * (s % 16))) - 1)
            self.packed[:, w0 + nw - 1] = (self.packed[:, w0 + nw - 1] & ~keep) | (blk[:, nw - 1] & keep)
            nw -= 1
        self.packed[:, w0:w0 + nw] = blk[:, :nw]
