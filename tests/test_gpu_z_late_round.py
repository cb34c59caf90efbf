"""PPE_Solver 4 — geometric multigrid on the GPU through the C-ABI, against oracle/ifx_oracle_mg.c:
pressure bit-exact after every step, identical V-cycle counts.  PARITY UNPINNED (the reference has no multigrid);
tests/test_mg_shim.py runs the same kernel source on the CPU.  (The file name sorts after the pinned-parity suites on
purpose: `pytest -x` reaches the reference-parity and slab tests first.)"""
import numpy as np
import pytest

import immerseflow_b200 as ifx
import _oracle as orc
from test_gpu_full_parity import pair, assert_same_fields

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("reduce_mode", [ifx.IFX_REDUCE_FUSED, ifx.IFX_REDUCE_REFERENCE])
@pytest.mark.parametrize("ncx,ncy,body", [(64, 64, False), (96, 64, True), (200, 120, True), (181, 129, True)])
def test_multigrid_matches_oracle(ncx, ncy, body, reduce_mode):
    xf, yf = ifx.uniform_faces(ncx, ncx / 64.0), ifx.uniform_faces(ncy, ncy / 64.0)
    u0, v0, _ = orc.initial_condition(orc.Grid(xf, yf))
    g, o = pair(xf, yf, 1e-3, 150.0, 25, 40, ppe_tol=1e-5, reduce_mode=reduce_mode, ppe_solver=4, ppe_omega=1.0)
    o.set_ppe_solver(4, 1.0)
    with g:
        if body:
            bodies = [orc.circle_markers(0.3, 0.6, 0.11, 40), orc.ellipse_markers(0.9, 0.4, 0.2, 0.07, 0.4, 48)]
            g.set_bodies(bodies); o.set_bodies(bodies)
        g.initializeData()
        g.set("u", u0); g.set("v", v0); g.set("p", np.zeros_like(u0))
        o.set("u", u0); o.set("v", v0)
        o.update_ib()
        for step in range(3):
            st = g.step(); so = o.step()
            assert (st.ad_iters, st.ppe_sweeps) == (int(so[0]), int(so[3])), f"step {step}"
            assert_same_fields(g, o, tag=f"step {step}")
            assert st.ppe_sweeps < 40 and st.ppe_residual <= 1e-5, (st.ppe_sweeps, st.ppe_residual)
            if reduce_mode == ifx.IFX_REDUCE_REFERENCE:
                assert st.ppe_residual == so[4]
    o.close()


def test_multigrid_moving_body_rebuilds_the_hierarchy():
    xf, yf = ifx.uniform_faces(128, 2.0), ifx.uniform_faces(64, 1.0)
    g, o = pair(xf, yf, 2e-3, 100.0, 25, 30, ppe_tol=1e-5, ppe_solver=4, ppe_omega=1.0)
    o.set_ppe_solver(4, 1.0)
    with g:
        g.initializeData()
        n = g.field_size("u")
        g.set("u", np.ones(n)); g.set("v", np.zeros(n)); g.set("p", np.zeros(n))
        o.set("u", np.ones(n)); o.set("v", np.zeros(n))
        for step in range(3):
            bodies = [orc.circle_markers(0.6 + 0.05 * step, 0.5, 0.15, 48)]
            g.set_bodies(bodies, [(0.5, 0.0)]); o.set_bodies(bodies, [(0.5, 0.0)])
            o.update_ib()
            st = g.step(); so = o.step()
            assert (st.ad_iters, st.ppe_sweeps) == (int(so[0]), int(so[3])), f"step {step}"
            assert_same_fields(g, o, tag=f"step {step}")
    o.close()


def test_multigrid_needs_three_cells_and_full_mode():
    inp = ifx.make_input(2, 64, 1e-3, 150.0)
    with pytest.raises(ifx.IfxError, match="at least 3 cells"):
        ifx.ImmerseFlow(inp, ifx.uniform_faces(2, 1.0), ifx.uniform_faces(64, 1.0), compat=ifx.IFX_COMPAT_FULL, ppe_solver=4)
    inp = ifx.make_input(64, 64, 1e-3, 150.0)
    with pytest.raises(ifx.IfxError, match="IFX_COMPAT_FULL"):
        ifx.ImmerseFlow(inp, ifx.uniform_faces(64, 1.0), ifx.uniform_faces(64, 1.0), ppe_solver=4)


@pytest.mark.parametrize("reduce_mode", [ifx.IFX_REDUCE_FUSED, ifx.IFX_REDUCE_REFERENCE])
@pytest.mark.parametrize("solver,omega,ncx,ncy,itermax", [(5, 1.0, 96, 64, 40), (5, 1.0, 180, 128, 40), (5, 1.0, 75, 51, 40), (2, 1.8, 32, 24, 3000)])
def test_line_relaxation_and_line_smoothed_multigrid_match_oracle(solver, omega, ncx, ncy, itermax, reduce_mode):
    """PPE_Solver 2 (zebra line SOR, the input file's own "2. Line SOR") and 5 (V-cycle smoothed by it) on a stretched grid
    with two bodies: pressure bit-exact after every step, identical iteration counts."""
    xf, yf = orc.stretched_faces(ncx, 4.0, 1.03), orc.stretched_faces(ncy, 2.0, 1.03)
    g, o = pair(xf, yf, 2e-3, 100.0, 25, itermax, ppe_tol=1e-5, reduce_mode=reduce_mode, ppe_solver=solver, ppe_omega=omega)
    o.set_ppe_solver(solver, omega)
    with g:
        bodies = [orc.circle_markers(1.5, 1.0, 0.3, 64), orc.ellipse_markers(2.6, 0.9, 0.35, 0.12, 0.5, 50)]
        g.set_bodies(bodies); o.set_bodies(bodies)
        g.initializeData()
        n = g.field_size("u")
        g.set("u", np.ones(n)); g.set("v", np.zeros(n)); g.set("p", np.zeros(n))
        o.set("u", np.ones(n)); o.set("v", np.zeros(n)); o.update_ib()
        for step in range(2):
            st = g.step(); so = o.step()
            assert (st.ad_iters, st.ppe_sweeps) == (int(so[0]), int(so[3])), f"step {step}"
            assert_same_fields(g, o, tag=f"step {step}")
            assert st.ppe_sweeps < itermax and st.ppe_residual <= 1e-5, (st.ppe_sweeps, st.ppe_residual)
            if reduce_mode == ifx.IFX_REDUCE_REFERENCE:
                assert st.ppe_residual == so[4]
    o.close()


@pytest.mark.xfail(strict=False, reason="graph replay of the coarse cycle was written after the round's GPU budget was spent: "
                                        "not yet run on a GPU (the default path does not use it)")
def test_multigrid_graph_replay_matches_plain_launches():
    """ifx_options.use_graphs = 1: same bits, same counts, fewer launches on the stream."""
    xf, yf = ifx.uniform_faces(128, 2.0), ifx.uniform_faces(64, 1.0)
    u0, v0, _ = orc.initial_condition(orc.Grid(xf, yf))
    out = []
    for graphs in (0, 1):
        g, o = pair(xf, yf, 1e-3, 150.0, 25, 40, ppe_tol=1e-5, ppe_solver=4, ppe_omega=1.0, use_graphs=graphs)
        o.close()
        with g:
            g.set_bodies([orc.circle_markers(0.6, 0.5, 0.15, 48)])
            g.initializeData()
            g.set("u", u0); g.set("v", v0); g.set("p", np.zeros_like(u0))
            sts = [g.step() for _ in range(3)]
            out.append(([s.ppe_sweeps for s in sts], g.get("p"), g.get("u")))
    assert out[0][0] == out[1][0]
    assert np.array_equal(out[0][1], out[1][1]) and np.array_equal(out[0][2], out[1][2])


def test_bodies_must_lie_inside_the_grid():
    inp = ifx.make_input(96, 64, 2e-3, 100.0)
    xf, yf = orc.stretched_faces(96, 4.0, 1.02), orc.stretched_faces(64, 2.0, 1.02)
    with ifx.ImmerseFlow(inp, xf, yf, compat=ifx.IFX_COMPAT_FULL) as s:
        s.initializeData()
        for bad in (orc.circle_markers(0.1, 1.0, 0.3, 64), orc.ellipse_markers(2.0, 1.97, 0.4, 0.1, 0.0, 48),
                    orc.circle_markers(5.0, 1.0, 0.3, 32)):
            with pytest.raises(ifx.IfxError, match="outermost cells"):
                s.set_bodies([orc.circle_markers(1.5, 1.0, 0.3, 32), bad])
        s.set_bodies([orc.circle_markers(1.5, 1.0, 0.3, 32)])
        s.iblank_update()
        assert len(s.ghost_cells()["cell"]) > 0


@pytest.mark.parametrize("name", ["jacobi", "line_sor", "rb_sor", "multigrid_point", "multigrid_line"])
def test_gpu_reproduces_the_frozen_full_mode_outputs(name, golden_dir):
    """The CUDA path against a file (tests/golden/full_mode/full_mode.npz: inputs + the oracle's outputs, frozen), for
    every Poisson solver: cell types, ghost-cell list, u, v, p, body forces, iteration counts."""
    import os
    import sys
    sys.path.insert(0, os.path.join(golden_dir, "full_mode"))
    import make_fixtures as mf
    z = np.load(os.path.join(golden_dir, "full_mode", "full_mode.npz"))
    solver, omega, itermax, tol = mf.CASES[name]
    xf, yf = z["xf"], z["yf"]
    inp = ifx.make_input(mf.NCX, mf.NCY, mf.DT, mf.RE, AD_itermax=mf.AD_ITERMAX, PPE_itermax=itermax)
    bodies = [z[f"markers{k}"] for k in range(int(z["nbodies"]))]
    with ifx.ImmerseFlow(inp, xf, yf, compat=ifx.IFX_COMPAT_FULL, ppe_abs_residual=1, ppe_tol=tol, ppe_solver=solver,
                         ppe_omega=omega) as s:
        s.set_bodies(bodies, [tuple(v) for v in z["vel"]])
        s.initializeData()
        n = s.field_size("u")
        s.set("u", np.ones(n)); s.set("v", np.zeros(n)); s.set("p", np.zeros(n))
        counts = []
        for _ in range(mf.STEPS):
            st = s.step()
            counts.append([st.ad_iters, st.ppe_sweeps])
        assert counts == z[f"{name}/counts"].tolist()
        assert np.array_equal(s.get("celltype").astype(np.uint8), z[f"{name}/celltype"])
        assert np.array_equal(s.ghost_cells()["cell"], z[f"{name}/ghost_cells"])
        for f in ("u", "v", "p"):
            assert np.array_equal(s.get(f), z[f"{name}/{f}"]), f
        assert np.array_equal(s.body_forces(len(bodies)), z[f"{name}/forces"])


@pytest.mark.parametrize("compat", [ifx.IFX_COMPAT_REFERENCE, ifx.IFX_COMPAT_FULL])
def test_nan_in_the_state_is_reported_not_declared_converged(compat, ref_case):
    """Failure detection: NaN > tol is false, so every stop rule of the reference would end "converged" on a
    poisoned state; the library returns IFX_ERR_STATE instead."""
    inp = ifx.make_input(50, 50, 1e-3, 150.0)
    with ifx.ImmerseFlow(inp, ref_case["xf"], ref_case["yf"], compat=compat) as s:
        s.initializeData()
        u = s.get("u")
        u[52 * 20 + 17] = np.nan
        s.set("u", u)
        with pytest.raises(ifx.IfxError, match="not finite"):
            s.step()


def test_zero_copy_control_and_async_transfers_change_nothing():
    """ifx_options.zero_copy_control moves the step's few-byte control traffic through host-mapped memory instead of
    cudaMemcpyAsync, ifx_set/get_field_async drop the wait: same bits, same counts (bench.py's end-to-end pipeline
    relies on both)."""
    xf, yf = orc.stretched_faces(96, 4.0, 1.02), orc.stretched_faces(64, 2.0, 1.02)
    inp = ifx.make_input(96, 64, 2e-3, 100.0, AD_itermax=25, PPE_itermax=120)
    out = []
    for zc in (0, 1):
        with ifx.ImmerseFlow(inp, xf, yf, compat=ifx.IFX_COMPAT_FULL, ppe_abs_residual=1, zero_copy_control=zc) as s:
            s.initializeData()
            n = s.field_size("u")
            u0, v0, p0 = np.ones(n), np.zeros(n), np.zeros(n)
            if zc:      # upload without waiting; the step that follows is ordered behind the copies on the handle's stream
                for name, a in (("u", u0), ("v", v0), ("p", p0)):
                    s.set_async(name, a)
            else:
                s.set("u", u0); s.set("v", v0); s.set("p", p0)
            counts = []
            for step in range(3):
                s.set_bodies([orc.circle_markers(1.5 + 0.02 * step, 1.0, 0.3, 64)], [(0.2, 0.0)])
                st = s.step()
                counts.append((st.ad_iters, st.ppe_sweeps))
            if zc:
                got = {k: np.empty(n) for k in ("u", "v", "p")}
                for k in got:
                    s.get_async(k, got[k])
                s.synchronize()
            else:
                got = {k: s.get(k) for k in ("u", "v", "p")}
            out.append((counts, got, s.ghost_cells()["cell"]))
    assert out[0][0] == out[1][0]
    assert np.array_equal(out[0][2], out[1][2])
    for k in ("u", "v", "p"):
        assert np.array_equal(out[0][1][k], out[1][1][k]), k
