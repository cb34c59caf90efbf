"""The oracle's UNPINNED stages against their frozen outputs (tests/golden/full_mode/full_mode.npz, written by
make_fixtures.py next to it): a change in what the oracle defines shows up here and has to be re-frozen on purpose."""
import os
import sys

import numpy as np
import pytest

from conftest import GOLDEN

sys.path.insert(0, os.path.join(GOLDEN, "full_mode"))
import make_fixtures as mf  # noqa: E402


@pytest.fixture(scope="module")
def frozen():
    return np.load(os.path.join(GOLDEN, "full_mode", "full_mode.npz"))


def test_inputs_are_what_the_script_generates(frozen):
    xf, yf, bodies, vel = mf.inputs()
    assert np.array_equal(frozen["xf"], xf) and np.array_equal(frozen["yf"], yf) and np.array_equal(frozen["vel"], vel)
    # the markers come from numpy's cos / sin: allow the last bit (the frozen ones are what every comparison below uses)
    for k, m in enumerate(bodies):
        assert np.allclose(frozen[f"markers{k}"], m, rtol=0, atol=1e-15)


@pytest.mark.parametrize("name", list(mf.CASES))
def test_oracle_reproduces_its_frozen_outputs(frozen, name):
    bodies = [frozen[f"markers{k}"] for k in range(int(frozen["nbodies"]))]
    out = mf.run(frozen["xf"], frozen["yf"], bodies, frozen["vel"], *mf.CASES[name])
    assert out["counts"].tolist() == frozen[f"{name}/counts"].tolist()
    for key in ("celltype", "ghost_cells", "u", "v", "p", "forces"):
        assert np.array_equal(out[key], frozen[f"{name}/{key}"]), key
