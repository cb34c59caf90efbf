"""Grid widths on and around the column-tile widths of the sweep kernels (256 for the predictor and the slab geometry of the
general Poisson sweep, 512 for its single-GPU geometry and for the Laplace sweep, 240 for the pair sweep): the last tile column then holds exactly one, all but one, or no inactive
column — the guarded edge paths — and the result must equal the CPU oracle's bit for bit."""
import numpy as np
import pytest

import immerseflow_b200 as ifx
import _oracle as orc

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("ncx", [239, 240, 241, 255, 256, 257, 479, 481, 511, 512, 513, 767, 769])
@pytest.mark.parametrize("pairs", [0, 1])
def test_full_step_on_widths_around_the_tile_widths(ncx, pairs):
    ncy = 45
    xf, yf = orc.stretched_faces(ncx, 6.0, 1.004), orc.stretched_faces(ncy, 1.5, 1.02)
    dt, Re, ad_it, ppe_it = 2e-3, 150.0, 7, 13
    inp = ifx.make_input(ncx, ncy, dt, Re, AD_itermax=ad_it, PPE_itermax=ppe_it)
    # two bodies: one near the west wall, one straddling the boundary between the last two column tiles
    xe = xf[min(ncx - 20, (ncx // 240) * 240)] if ncx > 260 else xf[ncx // 2]
    bodies = [orc.circle_markers(0.6, 0.75, 0.21, 40), orc.ellipse_markers(float(xe), 0.7, 0.3, 0.16, 0.4, 48)]
    o = orc.FullSolver(xf, yf, dt, Re, ad_it, ppe_it, ppe_abs=1)
    with ifx.ImmerseFlow(inp, xf, yf, compat=ifx.IFX_COMPAT_FULL, ppe_abs_residual=1, ppe_pairs=pairs) as s:
        s.initializeData()
        n = inp.nx * inp.ny
        rng = np.random.default_rng(ncx)
        u0, v0 = 1.0 + 0.05 * rng.standard_normal(n), 0.05 * rng.standard_normal(n)
        s.set("u", u0); s.set("v", v0); o.set("u", u0); o.set("v", v0)
        s.set_bodies(bodies); o.set_bodies(bodies); o.update_ib()
        inner = np.zeros((inp.ny, inp.nx), bool); inner[1:-1, 1:-1] = True
        inner = inner.reshape(-1)
        for step in range(2):
            st = s.step(); so = o.step()
            assert (st.ad_iters, st.ppe_sweeps) == (int(so[0]), int(so[3]))
            for k in ("u", "v", "p"):
                got, want = s.get(k), o.get(k)
                assert np.array_equal(got[inner], want[inner]), (ncx, pairs, step, k, int(np.count_nonzero(got[inner] != want[inner])))
    o.close()


@pytest.mark.parametrize("ncx", [255, 256, 257, 511, 512, 513, 1023, 1025])
def test_reference_mode_on_widths_around_the_tile_widths(ncx):
    ncy = ncx + 3                                          # reference mode needs nx <= ny (App. A Q2)
    xf, yf = orc.stretched_faces(ncx, 2.0, 1.003), orc.stretched_faces(ncy, 2.0, 1.003)
    inp = ifx.make_input(ncx, ncy, 1e-3, 150.0, AD_itermax=6, PPE_itermax=9)
    g = orc.Grid(xf, yf)
    with ifx.ImmerseFlow(inp, xf, yf) as s:
        s.initializeData()
        pr = orc.Predictor(g, s.get("u"), s.get("v"), inp.dt, inp.Re, inp.AD_itermax)
        st = s.ADsolver()
        k, _ = pr.step()
        assert st.ad_iters == k
        m = np.ones((g.ny, g.nx), bool)
        m[0, 0] = m[0, -1] = m[-1, 0] = m[-1, -1] = False          # the reference's racy corner ghosts (App. A Q6)
        m = m.reshape(-1)
        assert np.array_equal(s.get("u")[m], pr.u[m]) and np.array_equal(s.get("v")[m], pr.v[m])
        k_o, p_o, _ = orc.ppe_solve(g, np.zeros(g.nx * g.ny), inp.PPE_itermax)
        st = s.PPESolver()
        assert st.ppe_sweeps == k_o and np.array_equal(s.get("p"), p_o)
