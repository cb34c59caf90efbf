"""tools/make_case.py writes BASELINE.json's configurations in the reference's file formats: the files must read back
through the C-ABI's own readers (host only), the stretched grids must look like the shipped inputs/xgrid.dat (uniform
core over the body, smooth geometric growth), the bodies must sit inside the core."""
import os
import sys

import numpy as np
import pytest

import immerseflow_b200 as ifx
from conftest import ROOT

sys.path.insert(0, os.path.join(ROOT, "tools"))
import make_case  # noqa: E402


@pytest.mark.parametrize("case,scale", [("cavity", 1 / 16), ("cylinder", 1 / 16), ("airfoil", 1 / 32), ("bodies", 1 / 64)])
def test_case_files_read_back(tmp_path, case, scale):
    r = make_case.build(case, str(tmp_path), scale, steps=7)
    inp = ifx.read_input_file(str(tmp_path / "inputs" / "inputs.txt"))
    ncx, ncy = r["cells"]
    assert (inp.nx, inp.ny, inp.nxf, inp.nyf) == (ncx + 2, ncy + 2, ncx + 1, ncy + 1)
    assert ncx % 2 == 0 and ncy % 2 == 0
    assert (inp.tmax, inp.AD_itermax, inp.PPE_itermax) == (7.0, 25, 50) and inp.Re == r["Re_file"] and abs(inp.dt - r["dt"]) < 1e-6 * r["dt"]
    assert inp.PPE_solver == (4 if case in ("cavity", "bodies") else 5)
    xf = ifx.read_grid_file(str(tmp_path / "inputs" / "xgrid.dat2"), inp.nxf)
    yf = ifx.read_grid_file(str(tmp_path / "inputs" / "ygrid.dat2"), inp.nyf)
    for f, g in ((xf, r["xf"]), (yf, r["yf"])):
        d = np.diff(f)
        assert np.all(d > 0) and np.allclose(f, g, rtol=1e-6, atol=1e-9)
        ratio = d[1:] / d[:-1]
        assert ratio.max() < 1.25 and ratio.min() > 0.8            # smooth (the shipped grid jumps by up to 1.7)
    if case in ("cylinder", "airfoil"):
        dx, dy = np.diff(xf), np.diff(yf)
        assert dx.max() / dx.min() > 5 and dy.max() / dy.min() > 5
        m = r["bodies"][0][0]
        for f, c in ((xf, m[:, 0]), (yf, m[:, 1])):                # the body lies in the uniform core
            d = np.diff(f)
            k0, k1 = np.searchsorted(f, c.min()) - 1, np.searchsorted(f, c.max())
            assert np.allclose(d[k0:k1], d.min(), rtol=1e-5)
        txt = open(tmp_path / "inputs" / "bodies.txt").read().split()
        assert int(txt[0]) == 1 and int(txt[1]) == len(m)
    assert "immerseflow --mode full" in r["command"]


def test_growth_ratio_fits_the_length():
    r = make_case._growth_ratio(0.01, 50, 3.0)
    assert abs(0.01 * r * (r ** 50 - 1) / (r - 1) - 3.0) < 1e-9 and r > 1
    assert make_case._growth_ratio(0.1, 50, 3.0) == 1.0
