"""Row sharding of the VGA step across the GPUs of one box.

Entries are independent given EL[i,], EF[j,], sigma2; the only couplings are the global
update count (MAX over ranks) and the column statistics (SUM over ranks).  Rank r owns the
contiguous rows ``shard_rows(n, world)[r]`` of Y / M / EL -- what ``Y[b,]`` does for a batch in
R/ebpmf_log.R:199.  Pure index arithmetic, no compute.
"""
from __future__ import annotations

import numpy as np


def shard_rows(n, world, align=256):
    """[(row0, nrows)] for ranks 0..world-1: contiguous, sizes differ by at most one aligned block
    (alignment to the kernels' 256-row blocks keeps every shard's last block the only ragged one)."""
    if world < 1 or n < 0:
        raise ValueError("bad shard request")
    nblk = -(-n // align)
    base, extra = divmod(nblk, world)
    out, row0 = [], 0
    for r in range(world):
        rows = (base + (1 if r < extra else 0)) * align
        rows = max(0, min(rows, n - row0))
        out.append((row0, rows))
        row0 += rows
    return out


def shard_csc(colptr, rowidx, x, row0, nrows):
    """Row slice [row0, row0+nrows) of a dgCMatrix given by its slots -> slots of the slice
    (row indices re-based to the shard), keeping the within-column order."""
    colptr = np.asarray(colptr); rowidx = np.asarray(rowidx); x = np.asarray(x)
    keep = (rowidx >= row0) & (rowidx < row0 + nrows)
    csum = np.concatenate([[0], np.cumsum(keep)])
    new_ptr = csum[colptr].astype(np.int32)
    return new_ptr, (rowidx[keep] - row0).astype(np.int32), x[keep].astype(np.float64)
