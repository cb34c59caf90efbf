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
        srcs = [os.path.join(ORACLE_DIR, f) for f in ("ifx_oracle.c", "ifx_oracle_full.c", "ifx_oracle.h")]
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
