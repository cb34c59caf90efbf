"""ctypes binding of the CPU oracle (oracle/libifx_oracle.so).  TEST INFRASTRUCTURE ONLY."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
LIB = os.path.join(ORACLE_DIR, "libifx_oracle.so")

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)


def P(a):
    assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(_dp)


def PI(a):
    assert a.dtype == np.int32 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(_ip)


_lib = None


def lib():
    global _lib
    if _lib is None:
        srcs = [os.path.join(ORACLE_DIR, f) for f in ("ifx_oracle.c", "ifx_oracle_full.c", "ifx_oracle_mg.c", "ifx_oracle_diag.c", "ifx_oracle.h")]
        if not os.path.exists(LIB) or any(os.path.getmtime(s) > os.path.getmtime(LIB) for s in srcs):
            subprocess.check_call(["make", "-C", ORACLE_DIR, "libifx_oracle.so"], stdout=subprocess.DEVNULL)
        _lib = C.CDLL(LIB)
        _lib.orc_Reduction.restype = C.c_double
        _lib.orc_ADsolver.restype = C.c_int
        _lib.orc_PPESolver.restype = C.c_int
        _lib.orc_ADsolver_tol.restype = C.c_int
        _lib.orc_PPESolver_tol.restype = C.c_int
    return _lib


class Grid:
    """Metrics exactly as readGridData builds them (2-D dx, dy like the reference)."""

    def __init__(self, xf, yf):
        self.xf = np.ascontiguousarray(xf, dtype=np.float64)
        self.yf = np.ascontiguousarray(yf, dtype=np.float64)
        self.nx, self.ny = self.xf.size + 1, self.yf.size + 1
        N = self.nx * self.ny
        self.xc, self.yc = np.zeros(self.nx), np.zeros(self.ny)
        self.dx, self.dy = np.zeros(N), np.zeros(N)
        lib().orc_grid_metrics(self.nx, self.ny, P(self.xf), P(self.yf), P(self.xc), P(self.yc), P(self.dx), P(self.dy))


def initial_condition(g: Grid):
    N = g.nx * g.ny
    u, v, p = np.zeros(N), np.zeros(N), np.zeros(N)
    lib().orc_initializeKernel(g.nx, g.ny, P(g.xc), P(g.yc), P(u), P(v), P(p))
    return u, v, p


class Predictor:
    """State of the reference's predictor across time steps (u, v, face arrays)."""

    def __init__(self, g: Grid, u, v, dt, Re, itermax, iblank=None, vf_mode=0, tol=10.0 ** -6.0):
        self.g, self.dt, self.Re, self.itermax, self.vf_mode, self.tol = g, dt, Re, itermax, vf_mode, tol
        self.u, self.v = u.copy(), v.copy()
        self.uf = np.zeros((g.nx - 1) * (g.ny - 2))
        self.vf = np.zeros((g.nx - 2) * (g.ny - 1))
        self.iblank = np.ones(g.nx * g.ny) if iblank is None else iblank.copy()
        self.hist = np.zeros(2 * max(itermax, 1))

    def step(self):
        g = self.g
        k = lib().orc_ADsolver_tol(g.nx, g.ny, P(g.dx), P(g.dy), C.c_double(self.dt), C.c_double(self.Re), self.itermax,
                                   P(self.iblank), P(self.u), P(self.v), P(self.uf), P(self.vf), self.vf_mode,
                                   P(self.hist), C.c_double(self.tol))
        return k, self.hist[:2 * k].copy()


def ppe_solve(g: Grid, p, itermax, tol=10.0 ** -6.0):
    p = p.copy()
    res = C.c_double()
    k = lib().orc_PPESolver_tol(g.nx, g.ny, P(g.dx), P(g.dy), itermax, P(p), C.byref(res), C.c_double(tol))
    return k, p, res.value


def reduction(values, threads=256):
    v = np.ascontiguousarray(values, dtype=np.float64).reshape(-1)
    blocks = (v.size + threads - 1) // threads
    return lib().orc_Reduction(P(v), v.size, threads, blocks)


def stretched_faces(n_cells, length, ratio=1.04, seed=None):
    """Deterministic stretched grid in the spirit of inputs/xgrid.dat: fine uniform core, geometric growth outward."""
    core = max(2, n_cells // 3)
    side = n_cells - core
    left, right = side // 2, side - side // 2
    d = np.concatenate([ratio ** np.arange(left, 0, -1), np.ones(core), ratio ** np.arange(1, right + 1)])
    f = np.concatenate([[0.0], np.cumsum(d)])
    f *= length / f[-1]
    # 7 significant digits like the shipped grid files, so text round-trips exactly
    return np.array([float(f"{v:.7E}") for v in f])


# ------------------------------------------------------------------------------------------------
# UNPINNED stages (oracle/ifx_oracle_full.c): full fractional step with immersed bodies
# ------------------------------------------------------------------------------------------------
FIELD = {"u": 0, "v": 1, "p": 2, "iblank": 3, "uf": 4, "vf": 5, "sx": 6, "sy": 7, "ppe_rhs": 8, "celltype": 11}


def circle_markers(cx, cy, r, n):
    """Counter-clockwise polygon approximating a circle; generated once on the host and handed, identically, to the
    oracle and to the CUDA path (so libm never enters the CPU/GPU comparison)."""
    t = 2.0 * np.pi * np.arange(n) / n
    return np.ascontiguousarray(np.stack([cx + r * np.cos(t), cy + r * np.sin(t)], axis=1))


def ellipse_markers(cx, cy, a, b, angle, n):
    t = 2.0 * np.pi * np.arange(n) / n
    x, y = a * np.cos(t), b * np.sin(t)
    ca, sa = np.cos(angle), np.sin(angle)
    return np.ascontiguousarray(np.stack([cx + ca * x - sa * y, cy + sa * x + ca * y], axis=1))


class FullSolver:
    def __init__(self, xf, yf, dt, Re, ad_itermax, ppe_itermax, ad_tol=1e-6, ppe_tol=1e-6, ppe_abs=1,
                 bc_u=(1.0, 1.0, 1.0, 1.0), bc_v=(0.0, 0.0, 0.0, 0.0)):
        L = lib()
        L.orc_full_create.restype = C.c_void_p
        self.xf = np.ascontiguousarray(xf, dtype=np.float64)
        self.yf = np.ascontiguousarray(yf, dtype=np.float64)
        self.nx, self.ny = self.xf.size + 1, self.yf.size + 1
        bu = np.ascontiguousarray(bc_u, dtype=np.float64)      # W, E, S, N
        bv = np.ascontiguousarray(bc_v, dtype=np.float64)
        self.h = C.c_void_p(L.orc_full_create(self.nx, self.ny, P(self.xf), P(self.yf), C.c_double(dt), C.c_double(Re),
                                              ad_itermax, ppe_itermax, C.c_double(ad_tol), C.c_double(ppe_tol), ppe_abs,
                                              P(bu), P(bv)))
        self.stats = np.zeros(8)

    def close(self):
        if self.h:
            lib().orc_full_destroy(self.h)
            self.h = None

    def set_bodies(self, bodies, velocities=None):
        offs = np.zeros(len(bodies) + 1, dtype=np.int32)
        for b, m in enumerate(bodies):
            offs[b + 1] = offs[b] + len(m)
        xm = np.ascontiguousarray(np.concatenate([np.asarray(m, dtype=np.float64)[:, 0] for m in bodies]))
        ym = np.ascontiguousarray(np.concatenate([np.asarray(m, dtype=np.float64)[:, 1] for m in bodies]))
        ub = vb = None
        if velocities is not None:
            ub = np.ascontiguousarray([v[0] for v in velocities], dtype=np.float64)
            vb = np.ascontiguousarray([v[1] for v in velocities], dtype=np.float64)
        lib().orc_full_set_bodies(self.h, len(bodies), PI(offs), P(xm), P(ym), P(ub) if ub is not None else None,
                                  P(vb) if vb is not None else None)

    def update_ib(self):
        return lib().orc_full_update_ib(self.h)

    def get(self, name):
        n = {"uf": (self.nx - 1) * (self.ny - 2), "vf": (self.nx - 2) * (self.ny - 1)}.get(name, self.nx * self.ny)
        out = np.zeros(n)
        assert lib().orc_full_get(self.h, FIELD[name], P(out)) == 0
        return out

    def set(self, name, a):
        a = np.ascontiguousarray(a, dtype=np.float64).reshape(-1)
        assert lib().orc_full_set(self.h, FIELD[name], P(a)) == 0

    def set_ppe_solver(self, solver, omega):
        lib().orc_full_set_ppe_solver.argtypes = [C.c_void_p, C.c_int, C.c_double]
        lib().orc_full_set_ppe_solver(self.h, int(solver), float(omega))

    def set_mg(self, nu1, nu2, ncoarse):
        lib().orc_full_set_mg(self.h, int(nu1), int(nu2), int(ncoarse))

    def body_forces(self, nbodies):
        F = np.zeros((nbodies, 4))
        lib().orc_full_body_forces(self.h, P(F.reshape(-1)))
        return F

    def probe(self, x, y):
        x = np.ascontiguousarray(x, dtype=np.float64); y = np.ascontiguousarray(y, dtype=np.float64)
        u, v, p = np.zeros(x.size), np.zeros(x.size), np.zeros(x.size)
        lib().orc_full_probe(self.h, x.size, P(x), P(y), P(u), P(v), P(p))
        return u, v, p

    def predictor(self):
        lib().orc_full_predictor(self.h, P(self.stats)); return self.stats.copy()

    def poisson(self):
        lib().orc_full_poisson(self.h, P(self.stats)); return self.stats.copy()

    def correct(self):
        lib().orc_full_correct(self.h)

    def step(self):
        lib().orc_full_step(self.h, P(self.stats)); return self.stats.copy()

    def ghost_cells(self):
        n = lib().orc_full_ghost_cells(self.h, None, None, None, None, None)
        cell = np.zeros(n, dtype=np.int32); sten = np.zeros((n, 4), dtype=np.int32)
        w = np.zeros((n, 10)); bi = np.zeros((n, 2)); ip = np.zeros((n, 2))
        if n:
            lib().orc_full_ghost_cells(self.h, PI(cell), PI(sten.reshape(-1)), P(w.reshape(-1)), P(bi.reshape(-1)), P(ip.reshape(-1)))
        return {"cell": cell, "stencil": sten, "weights": w, "bi": bi, "ip": ip}
