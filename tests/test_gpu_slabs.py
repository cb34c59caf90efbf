"""Slab decomposition on >= 2 GPUs of one box: every field of the slab run is bit-identical to the single-GPU
run, with identical iteration counts (reference mode and full mode, uneven slabs)."""
import os
import subprocess
import sys

import pytest

import immerseflow_b200 as ifx
from conftest import ROOT

pytestmark = pytest.mark.gpu


def _ngpus():
    try:
        return ifx.load_library().ifx_device_count()
    except Exception:
        return 0


@pytest.mark.parametrize("world", [2, 4])
def test_slab_run_matches_single_gpu(world):
    if _ngpus() < world:
        pytest.skip(f"needs {world} GPUs")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
                        "--master-addr", "127.0.0.1", "--master-port", "29517", os.path.join(ROOT, "tests", "mgpu_worker.py")],
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and "MGPU_OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
