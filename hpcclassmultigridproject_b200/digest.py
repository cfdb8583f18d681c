"""Field digest that is independent of how many ranks hold the field.

The final iterate of a sharded run lives as row slabs on N ranks.  To let a 1-, 2-, 4- and 8-rank
run print the SAME hash for the same field, the hash is defined on a fixed partition: the dense
row-major (n+1) x (n+1) float64 field is cut into NSLAB = 8 row slabs (slab g = rows g*n/8 ..
(g+1)*n/8 - 1, the last one also holds row n -- the sharded solver's own cut for 8 ranks), every
slab is hashed with SHA-256 over its bytes, and the field digest is the SHA-256 of the eight slab
digests in slab order.  A rank that owns rows [lo, hi] computes the digests of the slabs inside
its rows; rank 0 concatenates them in rank order.
"""
from __future__ import annotations

import hashlib

NSLAB = 8


def slab_rows(n: int, g: int):
    """first and last row of slab g of a level with n+1 rows"""
    per = n // NSLAB
    return g * per, (g + 1) * per - 1 + (1 if g == NSLAB - 1 else 0)


def slab_digests(u, n: int, row_lo: int = 0, row_hi: int | None = None, base_row: int = 0):
    """{slab index: 32-byte digest} for the slabs lying inside rows row_lo..row_hi of the dense float64 field.
    `u` (numpy, C order, n+1 columns) holds rows base_row .. base_row + len(u) - 1 of it: the whole field
    (base_row = 0) or just the rows a rank owns."""
    import numpy as np
    row_hi = n if row_hi is None else row_hi
    assert u.dtype == np.float64 and u.ndim == 2 and u.shape[1] == n + 1 and u.flags.c_contiguous
    assert base_row <= row_lo and row_hi <= base_row + u.shape[0] - 1
    assert n % NSLAB == 0
    out = {}
    for g in range(NSLAB):
        lo, hi = slab_rows(n, g)
        if lo >= row_lo and hi <= row_hi:
            out[g] = hashlib.sha256(memoryview(u[lo - base_row:hi + 1 - base_row]).cast("B")).digest()
    return out


def combine(digests) -> str:
    """field digest from the eight slab digests ({g: bytes} or a list in slab order)"""
    if isinstance(digests, dict):
        assert sorted(digests) == list(range(NSLAB)), sorted(digests)
        digests = [digests[g] for g in range(NSLAB)]
    assert len(digests) == NSLAB
    return hashlib.sha256(b"".join(digests)).hexdigest()


def field_digest(u, n: int) -> str:
    return combine(slab_digests(u, n))
