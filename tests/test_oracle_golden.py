"""The oracle is only trusted after it reproduces every golden vector the reference ships
(SURVEY §4 / App. B): results/uc.dat, vc.dat (20 predictor steps), p.dat (Laplace-Jacobi at
convergence) and final_results.dat (iBlank == 1 + cell-centre coordinates), all to the six decimals
the files carry, in ALL 2704 cells including ghosts."""
import os

import numpy as np

import _oracle as orc
from conftest import fmt6, load_tecplot


def test_grid_metrics_match_final_results(ref_case):
    g = orc.Grid(ref_case["xf"], ref_case["yf"])
    gold = load_tecplot(os.path.join(ref_case["dir"], "results", "final_results.dat"))
    assert gold.shape == (52 * 52, 3)
    assert np.array_equal(fmt6(np.tile(g.xc, g.ny)), gold[:, 0])
    assert np.array_equal(fmt6(np.repeat(g.yc, g.nx)), gold[:, 1])
    ib = np.zeros(g.nx * g.ny)
    orc.lib().orc_iBlankComputeKernel(g.nx, g.ny, orc.P(g.xc), orc.P(g.yc), orc.P(ib))
    assert np.array_equal(fmt6(ib), gold[:, 2])          # iBlank == 1.000000 everywhere
    # ghost-centre rule (preSim.cu:304-307): -0.01 and 1.01 on the 50-cell unit square
    assert abs(g.xc[0] + 0.01) < 1e-12 and abs(g.xc[-1] - 1.01) < 1e-12


def test_predictor_20_steps_match_uc_vc(ref_case):
    g = orc.Grid(ref_case["xf"], ref_case["yf"])
    u, v, _ = orc.initial_condition(g)
    pr = orc.Predictor(g, u, v, ref_case["dt"], ref_case["Re"], ref_case["AD_itermax"])
    for step in range(20):
        k, hist = pr.step()
        assert k == 5                                     # 5 Jacobi iterations per step
        assert hist[-2] + hist[-1] <= 1e-6 < hist[-4] + hist[-3]
    gu = load_tecplot(os.path.join(ref_case["dir"], "results", "uc.dat"))
    gv = load_tecplot(os.path.join(ref_case["dir"], "results", "vc.dat"))
    assert np.array_equal(fmt6(pr.u), gu[:, 2])
    assert np.array_equal(fmt6(pr.v), gv[:, 2])
    # the goldens were written at step 20, not 19 or 21 (App. B)
    pr.step()
    assert np.abs(fmt6(pr.u) - gu[:, 2]).max() > 1e-5


def test_vf_bug_is_required_by_goldens(ref_case):
    """With vf computed "as intended" the goldens do NOT match: the as-written bug is the spec."""
    g = orc.Grid(ref_case["xf"], ref_case["yf"])
    u, v, _ = orc.initial_condition(g)
    pr = orc.Predictor(g, u, v, ref_case["dt"], ref_case["Re"], ref_case["AD_itermax"], vf_mode=1)
    for _ in range(20):
        pr.step()
    gu = load_tecplot(os.path.join(ref_case["dir"], "results", "uc.dat"))
    assert np.abs(fmt6(pr.u) - gu[:, 2]).max() > 1e-4


def test_laplace_ppe_matches_p_dat(ref_case):
    g = orc.Grid(ref_case["xf"], ref_case["yf"])
    k, p, res = orc.ppe_solve(g, np.zeros(g.nx * g.ny), ref_case["PPE_itermax"])
    gp = load_tecplot(os.path.join(ref_case["dir"], "results", "p.dat"))
    assert np.array_equal(fmt6(p), gp[:, 2])
    assert 14000 < k < 15500 and res <= 1e-6              # ~14.8 k sweeps
    P2 = p.reshape(g.ny, g.nx)
    assert P2[0, 0] == 100.0 and P2[0, -1] == 100.0 and P2[-1, 0] == 100.0 and P2[-1, -1] == 0.0


def test_reduction_order_is_reference_order():
    """orc_Reduction follows reduce6<256>'s pairing, not numpy's: check against a literal emulation."""
    rng = np.random.default_rng(7)
    for n in (1, 255, 256, 257, 511, 512, 513, 2704, 52 * 52 + 3, 70000):
        x = rng.standard_normal(n) * 10.0 ** rng.integers(-8, 8, n)
        B = (n + 255) // 256

        def block(vals, b, nb):
            s = np.zeros(256)
            for t in range(256):
                i = b * 512 + t
                while i < vals.size:
                    s[t] += (vals[i] + vals[i + 256]) if i + 256 < vals.size else vals[i]
                    i += 512 * nb
            off = 128
            while off >= 1:
                s[:off] = s[:off] + s[off:2 * off]
                off //= 2
            return s[0]

        part = np.array([block(x, b, B) for b in range(B)])
        expect = block(part, 0, 1)
        assert orc.reduction(x) == expect


def test_tecplot_writer_byte_format(tmp_path, ref_case):
    g = orc.Grid(ref_case["xf"], ref_case["yf"])
    ones = np.ones(g.nx * g.ny)
    out = tmp_path / "final_results.dat"
    assert orc.lib().orc_write_results_to_file(orc.P(g.xc), orc.P(g.yc), orc.P(ones), g.nx, g.ny, str(out).encode()) == 0
    ref = open(os.path.join(ref_case["dir"], "results", "final_results.dat"), "rb").read().replace(b"\r\n", b"\n")
    assert out.read_bytes() == ref
