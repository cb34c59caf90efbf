"""IFX_COMPAT_FULL on the GPU (through the C-ABI) against oracle/ifx_oracle_full.c: iBlank and ghost-cell index maps
bit-exact (north_star), interpolation weights, u, v, p, face velocities and the PPE source term bit-exact per step,
identical iteration counts.  PARITY UNPINNED with respect to the reference (it has no code for these stages); the
oracle defines them."""
import numpy as np
import pytest

import immerseflow_b200 as ifx
import _oracle as orc

pytestmark = pytest.mark.gpu


def pair(xf, yf, dt, Re, ad_it, ppe_it, bc_u=(1.0, 1.0, 1.0, 1.0), bc_v=(0.0, 0.0, 0.0, 0.0), ppe_tol=1e-6,
         reduce_mode=ifx.IFX_REDUCE_FUSED, **kw):
    inp = ifx.make_input(len(xf) - 1, len(yf) - 1, dt, Re, AD_itermax=ad_it, PPE_itermax=ppe_it)
    bc = {"u_bc_w": bc_u[0], "u_bc_e": bc_u[1], "u_bc_s": bc_u[2], "u_bc_n": bc_u[3],
          "v_bc_w": bc_v[0], "v_bc_e": bc_v[1], "v_bc_s": bc_v[2], "v_bc_n": bc_v[3]}
    g = ifx.ImmerseFlow(inp, xf, yf, compat=ifx.IFX_COMPAT_FULL, bc=bc, ppe_abs_residual=1, ppe_tol=ppe_tol,
                        reduce_mode=reduce_mode, **kw)
    o = orc.FullSolver(xf, yf, dt, Re, ad_it, ppe_it, ppe_tol=ppe_tol, ppe_abs=1, bc_u=bc_u, bc_v=bc_v)
    return g, o


def assert_same_fields(g, o, names=("u", "v", "p", "uf", "vf"), tag=""):
    for n in names:
        a, b = g.get(n), o.get(n)
        assert np.array_equal(a, b), f"{tag}: {n} differs, max |diff| = {np.abs(a - b).max():.3e} at {np.abs(a - b).argmax()}"


BODIES = {
    "circle": lambda: [orc.circle_markers(1.5, 1.0, 0.3, 64)],
    "ellipse": lambda: [orc.ellipse_markers(1.6, 1.05, 0.45, 0.12, 0.5, 80)],
    "three": lambda: [orc.circle_markers(1.2, 0.7, 0.2, 40), orc.ellipse_markers(2.2, 1.2, 0.3, 0.15, -0.3, 50),
                      np.array([[2.9, 0.5], [3.3, 0.55], [3.35, 0.9], [3.1, 1.1], [2.85, 0.8]])],
}


@pytest.mark.parametrize("body", ["circle", "ellipse", "three"])
@pytest.mark.parametrize("ncx,ncy", [(96, 64), (257, 130)])
def test_iblank_and_ghost_cell_maps_bit_exact(body, ncx, ncy):
    xf, yf = orc.stretched_faces(ncx, 4.0, 1.02), orc.stretched_faces(ncy, 2.0, 1.02)
    g, o = pair(xf, yf, 1e-3, 100.0, 5, 10)
    with g:
        bodies = BODIES[body]()
        g.set_bodies(bodies); o.set_bodies(bodies)
        g.initializeData()
        g.iblank_update()
        n = o.update_ib()
        assert np.array_equal(g.get("celltype"), o.get("celltype"))          # 0 solid / 1 fluid / 2 ghost
        assert np.array_equal(g.get("iblank"), o.get("iblank"))              # the reference's 1.0 / 0.0 array
        a, b = g.ghost_cells(), o.ghost_cells()
        assert len(a["cell"]) == n > 0
        assert np.array_equal(a["cell"], b["cell"])                          # ghost-cell index map
        assert np.array_equal(a["stencil"], b["stencil"])                    # interpolation-stencil index map
        assert np.array_equal(a["weights"], b["weights"])                    # Dirichlet + Neumann closures, owner body
        assert np.array_equal(a["bi"], b["bi"]) and np.array_equal(a["ip"], b["ip"])
    o.close()


@pytest.mark.parametrize("mode", [ifx.IFX_REDUCE_FUSED, ifx.IFX_REDUCE_REFERENCE])
def test_full_step_cylinder_bit_exact(mode):
    xf, yf = orc.stretched_faces(96, 4.0, 1.02), orc.stretched_faces(64, 2.0, 1.02)
    g, o = pair(xf, yf, 2e-3, 100.0, 25, 300, reduce_mode=mode, sweeps_per_batch=64)
    with g:
        bodies = BODIES["circle"]()
        g.set_bodies(bodies); o.set_bodies(bodies)
        g.initializeData()
        n = g.field_size("u")
        u0, v0 = np.ones(n), np.zeros(n)
        g.set("u", u0); g.set("v", v0); g.set("p", np.zeros(n))
        o.set("u", u0); o.set("v", v0); o.update_ib()
        for step in range(4):
            st = g.step()
            so = o.step()
            assert (st.ad_iters, st.ppe_sweeps) == (int(so[0]), int(so[3])), f"step {step}"
            assert_same_fields(g, o, ("u", "v", "p", "uf", "vf", "ppe_rhs", "sx", "sy"), f"step {step}")
            if mode == ifx.IFX_REDUCE_REFERENCE:
                assert (st.ad_ures, st.ad_vres, st.ppe_residual) == (so[1], so[2], so[4])
    o.close()


def test_full_stages_separately_and_convergence():
    """predictor / Poisson / projection one at a time; Poisson run to its tolerance: counts must match exactly."""
    xf, yf = orc.stretched_faces(48, 4.0, 1.02), orc.stretched_faces(32, 2.0, 1.02)
    g, o = pair(xf, yf, 2e-3, 100.0, 25, 60000, ppe_tol=1e-6, sweeps_per_batch=512)
    with g:
        bodies = BODIES["circle"]()
        g.set_bodies(bodies); o.set_bodies(bodies)
        g.initializeData()
        n = g.field_size("u")
        g.set("u", np.ones(n)); g.set("v", np.zeros(n)); g.set("p", np.zeros(n))
        o.set("u", np.ones(n)); o.set("v", np.zeros(n)); o.update_ib()
        a = g.ADsolver(); b = o.predictor()
        assert a.ad_iters == int(b[0])
        assert_same_fields(g, o, ("u", "v", "sx", "sy", "uf", "vf"), "predictor")
        a = g.PPESolver(); b = o.poisson()
        assert a.ppe_sweeps == int(b[3]) and 1000 < a.ppe_sweeps < 60000 and a.ppe_residual <= 1e-6
        assert_same_fields(g, o, ("p", "ppe_rhs"), "poisson")
        g.correct(); o.correct()
        assert_same_fields(g, o, ("u", "v", "uf", "vf"), "projection")
        # the projected face field is divergence-free on fluid cells
        gr = orc.Grid(xf, yf)
        dx, dy = gr.dx.reshape(gr.ny, gr.nx), gr.dy.reshape(gr.ny, gr.nx)
        uf = g.get("uf").reshape(gr.ny - 2, gr.nx - 1); vf = g.get("vf").reshape(gr.ny - 1, gr.nx - 2)
        ct = g.get("celltype").reshape(gr.ny, gr.nx)
        div = (uf[:, 1:] - uf[:, :-1]) / dx[1:-1, 1:-1] + (vf[1:, :] - vf[:-1, :]) / dy[1:-1, 1:-1]
        assert np.abs(div[ct[1:-1, 1:-1] == 1]).max() < 1e-9
    o.close()


@pytest.mark.parametrize("ncx,ncy", [(64, 64), (130, 70), (300, 257)])
def test_full_step_no_body_cavity_and_wide_grids(ncx, ncy):
    """No immersed body (lid-driven cavity BCs); also nx > ny, which the reference-compat mode cannot run."""
    xf, yf = ifx.uniform_faces(ncx, 1.0), ifx.uniform_faces(ncy, 1.0)
    g, o = pair(xf, yf, 1e-3, 100.0, 10, 80, bc_u=(0.0, 0.0, 0.0, 1.0), bc_v=(0.0, 0.0, 0.0, 0.0))
    with g:
        g.initializeData()
        n = g.field_size("u")
        rng = np.random.default_rng(ncx)
        u0, v0 = 0.05 * rng.standard_normal(n), 0.05 * rng.standard_normal(n)
        g.set("u", u0); g.set("v", v0); g.set("p", np.zeros(n))
        o.set("u", u0); o.set("v", v0); o.update_ib()
        for step in range(3):
            st = g.step(); so = o.step()
            assert (st.ad_iters, st.ppe_sweeps) == (int(so[0]), int(so[3]))
            assert_same_fields(g, o, tag=f"step {step}")
    o.close()


def test_moving_bodies_iblank_recomputed_every_step():
    """Config 5 of BASELINE.json in miniature: several bodies, iBlank + ghost cells rebuilt each step."""
    xf, yf = orc.stretched_faces(128, 4.0, 1.015), orc.stretched_faces(96, 2.0, 1.015)
    g, o = pair(xf, yf, 2e-3, 200.0, 15, 60)
    with g:
        g.initializeData()
        n = g.field_size("u")
        g.set("u", np.ones(n)); g.set("v", np.zeros(n)); g.set("p", np.zeros(n))
        o.set("u", np.ones(n)); o.set("v", np.zeros(n))
        for step in range(4):
            shift = 0.013 * step
            bodies = [orc.circle_markers(1.2 + shift, 0.8, 0.22, 48), orc.ellipse_markers(2.3, 1.1 - shift, 0.3, 0.14, 0.3 + shift, 64)]
            vel = [(6.5, 0.0), (0.0, -6.5)]
            g.set_bodies(bodies, vel); o.set_bodies(bodies, vel)
            o.update_ib()
            st = g.step(); so = o.step()
            assert np.array_equal(g.get("celltype"), o.get("celltype")), f"step {step}"
            a, b = g.ghost_cells(), o.ghost_cells()
            assert np.array_equal(a["cell"], b["cell"]) and np.array_equal(a["stencil"], b["stencil"])
            assert_same_fields(g, o, ("sx", "sy"), tag=f"step {step}")
            assert (st.ad_iters, st.ppe_sweeps) == (int(so[0]), int(so[3])), f"step {step}"
            assert_same_fields(g, o, tag=f"step {step}")
    o.close()


def test_comb_body_overflows_the_row_crossing_lists():
    """A comb with 450 teeth: 900 polygon edges straddle every grid row through the teeth, more than the classifier's
    per-row crossing list holds, so those rows take the cell-by-cell fallback.  Same bytes either way."""
    xf, yf = orc.stretched_faces(420, 4.0, 1.01), orc.stretched_faces(90, 2.0, 1.01)
    g, o = pair(xf, yf, 1e-3, 100.0, 5, 10)
    teeth = 450
    xs = np.linspace(3.2, 0.8, 2 * teeth + 1)
    top = np.stack([xs, np.where(np.arange(xs.size) % 2 == 0, 1.4, 0.9)], axis=1)
    comb = np.concatenate([np.array([[0.8, 0.6], [3.2, 0.6]]), top[:-1]])          # counter-clockwise
    with g:
        g.set_bodies([comb]); o.set_bodies([comb])
        g.initializeData()
        g.iblank_update()
        n = o.update_ib()
        assert n > 0
        assert np.array_equal(g.get("celltype"), o.get("celltype"))
        a, b = g.ghost_cells(), o.ghost_cells()
        for k in ("cell", "stencil", "weights", "bi", "ip"):
            assert np.array_equal(a[k], b[k]), k
    o.close()


@pytest.mark.parametrize("reduce_mode", [ifx.IFX_REDUCE_FUSED, ifx.IFX_REDUCE_REFERENCE])
def test_red_black_sor_matches_oracle_and_beats_jacobi(reduce_mode):
    """PPE_Solver = 3, red-black SOR with w-PPE.  Same bits and the same iteration counts as the oracle's
    half-sweeps (vortex flow past a small body), and it reaches a tolerance point Jacobi does not get near in more
    than twice the iterations."""
    xf, yf = ifx.uniform_faces(96, 1.0), ifx.uniform_faces(64, 1.0)
    u0, v0, _ = orc.initial_condition(orc.Grid(xf, yf))
    counts = {}
    for solver, omega in ((3, 1.9), (1, 1.0)):
        g, o = pair(xf, yf, 1e-3, 150.0, 25, 6000, ppe_tol=1e-4, reduce_mode=reduce_mode, ppe_solver=solver, ppe_omega=omega)
        o.set_ppe_solver(solver, omega)
        with g:
            bodies = [orc.circle_markers(0.3, 0.6, 0.08, 40)]
            g.set_bodies(bodies); o.set_bodies(bodies)
            g.initializeData()
            g.set("u", u0); g.set("v", v0); g.set("p", np.zeros_like(u0))
            o.set("u", u0); o.set("v", v0)
            o.update_ib()
            for step in range(2):
                st = g.step(); so = o.step()
                assert (st.ad_iters, st.ppe_sweeps) == (int(so[0]), int(so[3])), f"solver {solver} step {step}"
                assert_same_fields(g, o, tag=f"solver {solver} step {step}")
            counts[solver] = (st.ppe_sweeps, st.ppe_residual)
        o.close()
    assert counts[3][0] < 3000 and counts[3][1] <= 1e-4, counts          # SOR converged ...
    assert counts[1][0] == 6000 and counts[1][1] > 1.0, counts           # ... Jacobi is nowhere near after 6000 sweeps


def test_sor_is_refused_in_reference_mode(ref_case):
    inp = ifx.make_input(50, 50, 1e-3, 150.0)
    with pytest.raises(ifx.IfxError, match="IFX_COMPAT_FULL"):
        ifx.ImmerseFlow(inp, ref_case["xf"], ref_case["yf"], ppe_solver=3, ppe_omega=1.5)


def test_reference_mode_rejects_full_only_calls(ref_case):
    inp = ifx.make_input(50, 50, 1e-3, 150.0)
    with ifx.ImmerseFlow(inp, ref_case["xf"], ref_case["yf"]) as s:
        s.initializeData()
        with pytest.raises(ifx.IfxError, match="IFX_COMPAT_FULL"):
            s.correct()
        with pytest.raises(ifx.IfxError, match="IFX_COMPAT_FULL"):
            s.set_bodies([orc.circle_markers(0.5, 0.5, 0.1, 16)])
