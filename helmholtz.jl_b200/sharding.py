"""Right-hand sides are independent linear systems (columns of B, ShiftedLaplacianMultigridSolver.jl:46-47):
they are sharded over ranks / devices as contiguous column ranges with no data-path collective."""
from __future__ import annotations


def column_range(nrhs: int, nparts: int, part: int):
    """Contiguous range [c0, c1) of columns owned by `part` of `nparts` (same rule as the C library)."""
    if nparts < 1 or not (0 <= part < nparts):
        raise ValueError("bad partition")
    base, rem = divmod(int(nrhs), int(nparts))
    c0 = part * base + min(part, rem)
    c1 = c0 + base + (1 if part < rem else 0)
    return c0, c1


def all_ranges(nrhs: int, nparts: int):
    return [column_range(nrhs, nparts, p) for p in range(nparts)]
