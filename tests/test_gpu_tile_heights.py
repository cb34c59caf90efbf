"""IFX_ROWS_PER_CTA (rows per tile of the sweep kernels; normally chosen from the grid size): results do not depend on the
tile geometry — fields and iteration counts are bit-identical for every tile height, in both modes, with a body, through the
single sweeps and the pair sweep (the fused residual sums change their summation order with the geometry; the certified stop
rule keeps the counts)."""
import os

import numpy as np
import pytest

import immerseflow_b200 as ifx
import _oracle as orc

pytestmark = pytest.mark.gpu


def _full(pairs=0):
    ncx, ncy = 520, 301
    xf, yf = orc.stretched_faces(ncx, 4.0, 1.01), orc.stretched_faces(ncy, 2.0, 1.012)
    inp = ifx.make_input(ncx, ncy, 2e-3, 200.0, AD_itermax=9, PPE_itermax=21)
    with ifx.ImmerseFlow(inp, xf, yf, compat=ifx.IFX_COMPAT_FULL, ppe_abs_residual=1, ppe_pairs=pairs) as s:
        s.initializeData()
        s.set_bodies([orc.ellipse_markers(1.7, 1.0, 0.5, 0.2, 0.3, 96)])
        cnt = []
        for _ in range(2):
            st = s.step()
            cnt.append((st.ad_iters, st.ppe_sweeps))
        return cnt, {k: s.get(k) for k in ("u", "v", "p")}


def _reference():
    ncx, ncy = 300, 420
    xf, yf = orc.stretched_faces(ncx, 2.0, 1.01), orc.stretched_faces(ncy, 3.0, 1.01)
    inp = ifx.make_input(ncx, ncy, 1e-3, 150.0, AD_itermax=8, PPE_itermax=30)
    with ifx.ImmerseFlow(inp, xf, yf) as s:
        s.initializeData()
        a = s.ADsolver(); b = s.PPESolver()
        return [(a.ad_iters, b.ppe_sweeps)], {k: s.get(k) for k in ("u", "v", "p")}


@pytest.mark.parametrize("case", ["full", "full_pairs", "reference"])
def test_results_do_not_depend_on_the_tile_height(case, monkeypatch):
    run = {"full": _full, "full_pairs": lambda: _full(1), "reference": _reference}[case]
    monkeypatch.delenv("IFX_ROWS_PER_CTA", raising=False)
    want_counts, want = run()
    for rows in (4, 7, 33, 256):
        monkeypatch.setenv("IFX_ROWS_PER_CTA", str(rows))
        counts, got = run()
        assert counts == want_counts, (rows, counts, want_counts)
        for k in want:
            assert np.array_equal(got[k], want[k]), (case, rows, k)
