"""Host-side logic of the slab decomposition, on CPU: row partition, scatter/assemble round trip, and the
blob all-gather over torch.distributed with the gloo backend at world_size 2 (the same call the GPU launcher
makes over NCCL)."""
import os
import socket
import subprocess
import sys
import textwrap

import numpy as np
import pytest

from immerseflow_b200 import slabs
from conftest import ROOT


@pytest.mark.parametrize("ny,n", [(52, 1), (52, 2), (53, 4), (16386, 8), (11, 8), (10, 8)])
def test_partition_rows_covers_interior_contiguously(ny, n):
    parts = slabs.partition_rows(ny, n)
    assert parts[0][0] == 1 and parts[-1][1] == ny - 1
    assert all(a[1] == b[0] for a, b in zip(parts, parts[1:]))
    sizes = [b - a for a, b in parts]
    assert max(sizes) - min(sizes) <= 1 and min(sizes) >= 1


def test_partition_rejects_impossible_splits():
    with pytest.raises(ValueError):
        slabs.partition_rows(9, 8)
    with pytest.raises(ValueError):
        slabs.partition_rows(100, 9)


@pytest.mark.parametrize("n", [1, 2, 3, 8])
def test_scatter_assemble_roundtrip(n):
    nx, ny = 37, 41
    f = np.random.default_rng(n).standard_normal(nx * ny)
    parts = [slabs.scatter_rows(f, nx, ny, n, r) for r in range(n)]
    for r, p in enumerate(parts):
        lo, hi = slabs.local_rows(ny, n, r)
        assert p.size == (hi - lo) * nx
    assert np.array_equal(slabs.assemble_rows(parts, nx, ny), f)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_blob_allgather_gloo_world2(tmp_path):
    script = tmp_path / "w.py"
    script.write_text(textwrap.dedent(f"""
        import os, sys
        sys.path.insert(0, {ROOT!r})
        import torch.distributed as dist
        from immerseflow_b200 import slabs
        dist.init_process_group("gloo")
        r, w = dist.get_rank(), dist.get_world_size()
        blob = bytes([r + 1]) * slabs.IPC_BLOB_BYTES
        allb = slabs.gather_blobs(blob, dist)
        assert len(allb) == w * slabs.IPC_BLOB_BYTES
        for q in range(w):
            assert allb[q * slabs.IPC_BLOB_BYTES:(q + 1) * slabs.IPC_BLOB_BYTES] == bytes([q + 1]) * slabs.IPC_BLOB_BYTES
        jb, je = slabs.partition_rows(52, w)[r]
        print("rank", r, "rows", jb, je, "ok")
        dist.destroy_process_group()
    """))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT=str(_free_port()))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", env["MASTER_PORT"], str(script)],
                       capture_output=True, text=True, timeout=300, env=env)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert r.stdout.count("ok") == 2


def test_weighted_partition_balances_cost_and_keeps_every_slab_tall_enough():
    import numpy as np
    ny = 1026
    w = np.ones(ny - 2)
    w[100:400] = 0.6            # rows through a body: cheaper
    for n in (1, 2, 3, 4, 8):
        part = slabs.partition_rows_weighted(ny, n, w)
        assert part[0][0] == 1 and part[-1][1] == ny - 1
        assert all(a[1] == b[0] for a, b in zip(part, part[1:]))
        assert all(e - b >= 8 for b, e in part)
        costs = [w[b - 1:e - 1].sum() for b, e in part]
        assert max(costs) - min(costs) <= 2.0 + 1e-9, costs          # within a couple of rows of each other
    # equal costs -> the equal-height partition (up to a row)
    eq, pw = slabs.partition_rows(ny, 4), slabs.partition_rows_weighted(ny, 4, np.ones(ny - 2))
    assert all(abs((e1 - b1) - (e2 - b2)) <= 1 for (b1, e1), (b2, e2) in zip(eq, pw))
    # too few rows for the minimum height: falls back to equal heights
    assert slabs.partition_rows_weighted(20, 4, np.ones(18)) == slabs.partition_rows(20, 4)
