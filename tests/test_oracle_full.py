"""Properties of the UNPINNED stages as the oracle defines them (CPU only).  Nothing in the reference pins these —
its PPE source is a stub, its projection file is empty, its iBlank kernel writes 1.0 everywhere — so the checks
here are the domain's own invariants: ghost cells are exactly the solid cells with a fluid neighbour, the
interpolation closures reproduce constants, the projection removes the divergence the Poisson solve was given."""
import numpy as np

import _oracle as orc


def make(ncx=48, ncy=32, **kw):
    xf, yf = orc.stretched_faces(ncx, 4.0, 1.02), orc.stretched_faces(ncy, 2.0, 1.02)
    return orc.FullSolver(xf, yf, 2e-3, 100.0, 25, kw.pop("ppe_itermax", 60000), **kw), xf, yf


def test_classification_and_ghost_cell_definition():
    s, xf, yf = make()
    s.set_bodies([orc.circle_markers(1.5, 1.0, 0.3, 64), orc.ellipse_markers(2.6, 0.9, 0.35, 0.15, 0.4, 48)])
    n = s.update_ib()
    ct = s.get("celltype").reshape(s.ny, s.nx).astype(int)
    g = orc.Grid(xf, yf)
    X, Y = np.meshgrid(g.xc, g.yc)
    inside = ((X - 1.5) ** 2 + (Y - 1.0) ** 2 < 0.29 ** 2)          # well inside the polygonal circle
    assert np.all(ct[inside] != 1)
    assert np.all(ct[[0, -1], :] == 1) and np.all(ct[:, [0, -1]] == 1)   # grid ghost ring stays fluid
    solid = ct != 1
    fluid_nb = np.zeros_like(solid)
    fluid_nb[1:-1, 1:-1] = (ct[1:-1, :-2] == 1) | (ct[1:-1, 2:] == 1) | (ct[:-2, 1:-1] == 1) | (ct[2:, 1:-1] == 1)
    assert np.array_equal(ct == 2, solid & fluid_nb)
    gc = s.ghost_cells()
    assert n == (ct == 2).sum() == len(gc["cell"])
    assert np.all(np.diff(gc["cell"]) > 0)                              # increasing reference id
    assert np.array_equal(np.sort(gc["cell"]), np.flatnonzero(ct.reshape(-1) == 2))
    ib = s.get("iblank")
    assert np.array_equal(ib, (ct.reshape(-1) == 1).astype(float))      # the reference's double iBlank convention
    s.close()


def test_interpolation_closures_reproduce_constants_and_reflect():
    s, xf, yf = make()
    s.set_bodies([orc.circle_markers(1.5, 1.0, 0.3, 96)])
    s.update_ib()
    gc = s.ghost_cells()
    w = gc["weights"]
    wd, cd, wn = w[:, :4], w[:, 4], w[:, 5:9]
    assert np.allclose(cd + wd.sum(1), 1.0, atol=1e-14)      # phi == c everywhere (incl. the surface) -> phi_GC == c
    assert np.allclose(wn.sum(1), 1.0, atol=1e-14)           # Neumann: phi_GC is an average
    assert np.all(wn >= 0) and np.all(wd <= 0) and np.all(cd > 0)
    # image point is the mirror of the ghost-cell centre through the body intercept, and lies outside the circle
    g = orc.Grid(xf, yf)
    xg, yg = g.xc[gc["cell"] % s.nx], g.yc[gc["cell"] // s.nx]
    assert np.allclose(gc["ip"][:, 0], 2 * gc["bi"][:, 0] - xg) and np.allclose(gc["ip"][:, 1], 2 * gc["bi"][:, 1] - yg)
    assert np.all(np.hypot(gc["ip"][:, 0] - 1.5, gc["ip"][:, 1] - 1.0) >= 0.3 * np.cos(np.pi / 96) - 1e-12)
    assert np.all(np.abs(np.hypot(gc["bi"][:, 0] - 1.5, gc["bi"][:, 1] - 1.0) - 0.3) < 0.3 * (1 - np.cos(np.pi / 96)) + 1e-12)
    s.close()


def test_projection_removes_divergence_and_poisson_converges():
    s, xf, yf = make(ppe_tol=1e-6)
    s.set_bodies([orc.circle_markers(1.5, 1.0, 0.3, 64)])
    s.update_ib()
    s.set("u", np.ones(s.nx * s.ny)); s.set("v", np.zeros(s.nx * s.ny))
    g = orc.Grid(xf, yf)
    dx, dy = g.dx.reshape(s.ny, s.nx), g.dy.reshape(s.ny, s.nx)
    ct = s.get("celltype").reshape(s.ny, s.nx)
    for _ in range(2):
        st = s.step()
        assert st[3] < 60000 and st[4] <= 1e-6            # point-Jacobi converged on the singular-but-compatible system
        uf = s.get("uf").reshape(s.ny - 2, s.nx - 1); vf = s.get("vf").reshape(s.ny - 1, s.nx - 2)
        div = (uf[:, 1:] - uf[:, :-1]) / dx[1:-1, 1:-1] + (vf[1:, :] - vf[:-1, :]) / dy[1:-1, 1:-1]
        assert np.abs(div[ct[1:-1, 1:-1] == 1]).max() < 1e-9
        # closed faces carry the (zero) body velocity
        closed_u = (ct[1:-1, :-1] != 1) | (ct[1:-1, 1:] != 1)
        assert np.all(uf[closed_u] == 0.0)
    s.close()


def test_no_body_lid_driven_cavity_step_is_finite_and_divergence_free():
    xf = yf = np.linspace(0, 1, 33)
    s = orc.FullSolver(xf, yf, 1e-3, 100.0, 25, 40000, bc_u=(0, 0, 0, 1.0), bc_v=(0, 0, 0, 0), ppe_tol=1e-7)
    assert s.update_ib() == 0
    for _ in range(3):
        st = s.step()
    u = s.get("u").reshape(s.ny, s.nx)
    assert np.isfinite(u).all() and u[-2, 5:-5].mean() > 0.01 and abs(u[1, 5:-5]).max() < 0.2   # lid drags the top rows
    uf = s.get("uf").reshape(s.ny - 2, s.nx - 1); vf = s.get("vf").reshape(s.ny - 1, s.nx - 2)
    h = 1.0 / 32
    div = (uf[:, 1:] - uf[:, :-1]) / h + (vf[1:, :] - vf[:-1, :]) / h
    assert np.abs(div).max() < 1e-8
    s.close()
