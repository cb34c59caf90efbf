"""pytest configuration: `gpu` marker + shared fixtures.

`-m "not gpu"` covers the oracle against the golden vectors, the host logic and the C-ABI surface;
`-m gpu` is the parity suite proper (CUDA path through the C-ABI vs the oracle / the reference
binary / the golden files).
"""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _have_gpu() -> bool:
    try:
        import immerseflow_b200 as ifx
        return ifx.load_library().ifx_device_count() > 0
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _have_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture(scope="session")
def ref_case(golden_dir):
    """The reference's shipped case: inputs.txt + the uniform 51-face grids its code actually opens."""
    d = os.path.join(golden_dir, "reference_case")
    xf = np.loadtxt(os.path.join(d, "inputs", "xgrid.dat2"))[:, 1].copy()
    yf = np.loadtxt(os.path.join(d, "inputs", "ygrid.dat2"))[:, 1].copy()
    return {"dir": d, "xf": xf, "yf": yf, "nx": 52, "ny": 52, "dt": 1e-3, "Re": 150.0, "AD_itermax": 25,
            "PPE_itermax": 100000}


def load_tecplot(path):
    """Reader side of the output contract (reference results/plot.py:19-28)."""
    return np.genfromtxt(path, skip_header=3, delimiter=",")


def fmt6(a):
    """What "%f" keeps of a value."""
    return np.array([float("%f" % x) for x in np.asarray(a).reshape(-1)])
