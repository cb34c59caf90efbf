"""Slab decomposition plumbing (one process per GPU, launched by torchrun).

The grid is cut along j (rows): rank r owns interior rows [j_begin, j_end) plus one halo row on each side.
Rows are contiguous in memory (id = i + j*nx), so a halo message is one row segment per field.  The data path
never goes through torch: each rank exports the CUDA-IPC handle of its exchange segment (ifx_ipc_export), the
launcher all-gathers the 128-byte blobs with torch.distributed (any backend — gloo on CPU in the tests, NCCL on
the GPU box), and ifx_ipc_connect maps the neighbours so the sweep kernels can store boundary rows and residual
partials straight into peer memory over NVLink.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

IPC_BLOB_BYTES = 128


def partition_rows(ny: int, nranks: int) -> List[Tuple[int, int]]:
    """Split the interior rows 1 .. ny-2 (ghost-inclusive ny) into `nranks` contiguous slabs, sizes differing by
    at most one row (larger slabs first).  Returns [(j_begin, j_end)], j_end exclusive."""
    n = ny - 2
    if nranks < 1 or nranks > 8:
        raise ValueError("1..8 slabs")
    if n < nranks:
        raise ValueError("fewer interior rows than ranks")
    base, extra = divmod(n, nranks)
    out, j = [], 1
    for r in range(nranks):
        rows = base + (1 if r < extra else 0)
        out.append((j, j + rows))
        j += rows
    assert j == ny - 1
    return out


def partition_rows_weighted(ny: int, nranks: int, row_cost: Sequence[float], min_rows: int = 8) -> List[Tuple[int, int]]:
    """Contiguous slabs of (nearly) equal COST instead of equal height: `row_cost[k]` is the relative cost of interior row
    k + 1 (k = 0 .. ny-3).  Rows inside immersed bodies are cheaper for the Poisson sweep (a cell that is not fluid is
    carried over, not relaxed), so equal-height slabs leave the ranks without a body as the slowest ones and every sweep
    runs at their pace.  Any contiguous partition gives bit-identical results (tests/test_gpu_slabs.py runs uneven
    slabs); every slab gets at least `min_rows` rows (ghost-cell stencils reach 4 rows into a neighbour)."""
    n = ny - 2
    if nranks < 1 or nranks > 8:
        raise ValueError("1..8 slabs")
    if len(row_cost) != n:
        raise ValueError("one cost per interior row")
    if n < nranks * min_rows:
        return partition_rows(ny, nranks)
    total = float(sum(row_cost))
    if not total > 0.0:
        return partition_rows(ny, nranks)
    out, j, acc, k = [], 1, 0.0, 0
    for r in range(nranks):
        target = total * (r + 1) / nranks
        j0 = j
        last_possible = ny - 1 - (nranks - 1 - r) * min_rows           # leave min_rows for every slab still to come
        while j < last_possible and (j - j0 < min_rows or acc + 0.5 * row_cost[k] < target):
            acc += row_cost[k]
            j += 1
            k += 1
        if r == nranks - 1:
            while j < ny - 1:
                acc += row_cost[k]; j += 1; k += 1
        out.append((j0, j))
    assert j == ny - 1 and all(e - b >= min_rows for b, e in out)
    return out


def local_rows(ny: int, nranks: int, rank: int) -> Tuple[int, int]:
    """Stored global rows [lo, hi) of a rank: its owned rows plus one halo / ghost row on each side."""
    jb, je = partition_rows(ny, nranks)[rank]
    return jb - 1, je + 1


def gather_blobs(blob: bytes, dist=None) -> bytes:
    """All-gather one fixed-size blob per rank, in rank order, with torch.distributed (already initialised)."""
    if len(blob) != IPC_BLOB_BYTES:
        raise ValueError("blob size")
    if dist is None:
        import torch.distributed as dist   # noqa: PLC0415
    world = dist.get_world_size()
    out: List[bytes] = [b""] * world
    dist.all_gather_object(out, bytes(blob))
    if any(len(b) != IPC_BLOB_BYTES for b in out):
        raise RuntimeError("a rank sent a malformed blob")
    return b"".join(out)


def connect(solver, dist=None) -> None:
    """Export this rank's exchange segment, all-gather, map the peers."""
    if dist is None:
        import torch.distributed as dist   # noqa: PLC0415
    blobs = gather_blobs(solver.ipc_export(), dist)
    solver.ipc_connect(blobs, dist.get_world_size())
    dist.barrier()


def scatter_rows(field_global, nx: int, ny: int, nranks: int, rank: int):
    """Rows of a global reference-layout field that rank `rank` stores (owned + halo rows)."""
    lo, hi = local_rows(ny, nranks, rank)
    return field_global.reshape(ny, nx)[lo:hi].reshape(-1).copy()


def assemble_rows(parts: Sequence, nx: int, ny: int):
    """Inverse of scatter_rows for a full set of per-rank arrays: owned rows from each rank, grid ghost rows from
    the first / last rank."""
    import numpy as np   # noqa: PLC0415
    nranks = len(parts)
    out = np.empty((ny, nx))
    for r, part in enumerate(parts):
        jb, je = partition_rows(ny, nranks)[r]
        lo, _ = local_rows(ny, nranks, r)
        rows = np.asarray(part).reshape(-1, nx)
        out[jb:je] = rows[jb - lo:je - lo]
        if r == 0:
            out[0] = rows[0]
        if r == nranks - 1:
            out[ny - 1] = rows[-1]
    return out.reshape(-1)
