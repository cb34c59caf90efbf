"""Multigrid transfer / coarse-level kernels (immerseflow_b200/csrc/kernels_mg.cu) compiled as plain C++ through
tests/shim/cuda_host_shim.h and run, launch grid and all, on the CPU — every kernel bit for bit against the oracle
(oracle/ifx_oracle_mg.c).  The GPU run of the same source is tests/test_gpu_z_late_round.py."""
import ctypes as C
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import _oracle as orc  # noqa: E402
from _oracle import P  # noqa: E402
from shim.build import build  # noqa: E402

PADL = 15


@pytest.fixture(scope="module")
def shim():
    return C.CDLL(build("mg", deps=("kernels_mg.cu", "multigrid.cuh", "common.cuh", "stencil_math.cuh")))


def pitch_of(nx):
    return (PADL + nx + 1 + 15) // 16 * 16


def pad(a, nx, ny, dtype=np.float64, fill=0):
    out = np.full((ny, pitch_of(nx)), fill, dtype=dtype)
    out[:, PADL:PADL + nx] = a.reshape(ny, nx)
    return np.ascontiguousarray(out.reshape(-1))


def unpad(a, nx, ny):
    return np.ascontiguousarray(a.reshape(ny, pitch_of(nx))[:, PADL:PADL + nx].reshape(-1))


def tables_1d(g):
    """the product's separable tables (capi.cu build_metrics), same IEEE expressions"""
    dx, dy = g.dx[:g.nx].copy(), g.dy[::g.nx].copy()

    def co(d):
        n = d.size
        a_p, a_m = np.ones(n), np.ones(n)
        a_p[1:-1] = 2.0 / (d[1:-1] * (d[1:-1] + d[2:]))
        a_m[1:-1] = 2.0 / (d[1:-1] * (d[1:-1] + d[:-2]))
        s = np.ones(n); s[1:-1] = a_p[1:-1] + a_m[1:-1]
        return a_p, a_m, s

    cE, cW, sx = co(dx)
    cN, cS, sy = co(dy)
    t = [np.ascontiguousarray(a) for a in (dx, dy, cE, cW, sx, cN, cS, sy)]
    arr = (C.POINTER(C.c_double) * 8)(*[P(a) for a in t])
    return t, arr


def case(ncx, ncy, stretched, bodies, seed=None):
    if stretched:
        xf, yf = orc.stretched_faces(ncx, 10.0, ratio=1.03), orc.stretched_faces(ncy, 5.0, ratio=1.02)
    else:
        xf, yf = np.linspace(0, 10, ncx + 1), np.linspace(0, 5, ncy + 1)
    g = orc.Grid(xf, yf)
    N = g.nx * g.ny
    ct = np.ones(N, dtype=np.uint8)
    if bodies:
        offs = np.array([0, 48, 48 + 40], dtype=np.int32)
        m = np.concatenate([orc.circle_markers(4.0, 2.5, 0.9, 48), orc.ellipse_markers(7.0, 1.6, 1.1, 0.35, 0.5, 40)])
        if seed is not None:          # random bodies, large against small grids: thin fluid gaps, isolated cells, blocked lines
            r = np.random.default_rng(seed)
            m = np.concatenate([orc.circle_markers(r.uniform(2, 8), r.uniform(1.5, 3.5), r.uniform(0.5, 1.4), 48),
                                orc.ellipse_markers(r.uniform(2, 8), r.uniform(1.5, 3.5), r.uniform(0.8, 2.5), r.uniform(0.2, 0.8),
                                                    r.uniform(0, 3), 40)])
        xm, ym = np.ascontiguousarray(m[:, 0]), np.ascontiguousarray(m[:, 1])
        body_of = np.zeros(N, dtype=np.int32)
        orc.lib().orc_iblank_classify(g.nx, g.ny, P(g.xc), P(g.yc), 2, orc.PI(offs), P(xm), P(ym),
                                      ct.ctypes.data_as(C.POINTER(C.c_ubyte)), orc.PI(body_of))
        assert seed is not None or (ct != 1).sum() > 20
    return g, ct


def u8(a):
    return a.ctypes.data_as(C.POINTER(C.c_ubyte))


@pytest.mark.parametrize("lines", [0, 1])
@pytest.mark.parametrize("ncx,ncy,stretched,bodies", [(64, 32, False, False), (96, 48, True, True), (40, 72, True, True),
                                                      (50, 50, False, True), (61, 35, True, True)])
def test_mg_kernels_match_oracle_bit_for_bit(shim, ncx, ncy, stretched, bodies, lines):
    check_all_kernels(shim, ncx, ncy, stretched, bodies, lines)


def random_shapes():
    r = np.random.default_rng(2024)
    out = [(3, 3, 0), (3, 9, 1), (4, 3, 2), (5, 5, 3), (7, 4, 4)]              # the smallest grids multigrid accepts
    out += [(int(r.integers(5, 45)), int(r.integers(5, 45)), 10 + k) for k in range(19)]
    return out


@pytest.mark.parametrize("ncx,ncy,seed", random_shapes())
def test_mg_kernels_on_random_shapes_and_bodies(shim, ncx, ncy, seed):
    """Odd and even cell counts from 3 up, random bodies that are large against the grid (blocked lines, isolated fluid
    cells, coarse cells with one child): every kernel of the file, both hierarchies, bit for bit against the oracle."""
    check_all_kernels(shim, ncx, ncy, bool(seed % 2), True, seed % 2, seed=seed)


def check_all_kernels(shim, ncx, ncy, stretched, bodies, lines, seed=None):
    L = orc.lib()
    L.orc_mg_create2.restype = C.c_void_p
    g, ct = case(ncx, ncy, stretched, bodies, seed)
    nx, ny, N = g.nx, g.ny, g.nx * g.ny
    lx, ly = (C.c_int * 16)(), (C.c_int * 16)()
    nlev = L.orc_mg_plan(ncx, ncy, lx, ly)
    lx2, ly2 = (C.c_int * 16)(), (C.c_int * 16)()
    assert shim.shim_mg_plan(ncx, ncy, lx2, ly2) == nlev and list(lx) == list(lx2) and list(ly) == list(ly2)
    assert nlev >= 2
    mg = C.c_void_p(L.orc_mg_create2(nx, ny, P(g.dx), P(g.dy), u8(ct), lines))
    tabs, tarr = tables_1d(g)
    ctp = pad(ct, nx, ny, dtype=np.uint8, fill=1)
    pitch = pitch_of(nx)

    def olevel(l, which):
        out = np.zeros((lx[l] + 2) * (ly[l] + 2))
        assert L.orc_mg_get(mg, l, which, P(out), None, None) == 0
        return out

    # ---- hierarchy of conductances
    GE = [None] * nlev; GN = [None] * nlev
    for l in range(1, nlev):
        n = (lx[l] + 2) * (ly[l] + 2)
        GE[l], GN[l] = np.zeros(n), np.zeros(n)
        if l == 1:
            shim.shim_mg_build1(nx, ny, pitch, tarr, u8(ctp), lx[1], ly[1], P(GE[1]), P(GN[1]), lines)
        else:
            shim.shim_mg_coarsen(lx[l - 1], ly[l - 1], P(GE[l - 1]), P(GN[l - 1]), lx[l], ly[l], P(GE[l]), P(GN[l]), lines)
        assert np.array_equal(GE[l], olevel(l, 0)) and np.array_equal(GN[l], olevel(l, 1)), f"level {l}"
        if bodies and l == 1 and seed is None:
            assert (GE[1].reshape(ly[1] + 2, lx[1] + 2)[1:-1, 1:-2] == 0).any()      # closed faces made it to level 1

    # ---- residual restriction from the fine level
    rng = np.random.default_rng(5)
    p = rng.standard_normal(N); rhs = rng.standard_normal(N) * 3.0
    cP, cxm, cxp, cym, cyp = (np.zeros(N) for _ in range(5))
    L.orc_calculatePPECoefficients(nx, ny, P(g.dx), P(g.dy), P(cP), P(cxm), P(cxp), P(cym), P(cyp))
    n1 = (lx[1] + 2) * (ly[1] + 2)
    R_o, R_s = np.zeros(n1), np.zeros(n1)
    L.orc_mg_restrict_fine(nx, ny, P(g.dx), P(g.dy), P(cP), P(cxm), P(cxp), P(cym), P(cyp), u8(ct), P(rhs), P(p),
                           lx[1] + 2, ly[1] + 2, P(R_o))
    pp, rp = pad(p, nx, ny), pad(rhs, nx, ny)
    shim.shim_mg_restrict_fine(nx, ny, pitch, tarr, u8(ctp), P(rp), P(pp), lx[1], ly[1], P(R_s))
    assert np.array_equal(R_o, R_s)
    assert np.abs(R_o).max() > 0

    # ---- smoothing, restriction, prolongation on every coarse level
    omega = C.c_double(1.15)
    for l in range(1, nlev):
        n = (lx[l] + 2) * (ly[l] + 2)
        NX, NY = lx[l] + 2, ly[l] + 2
        R = rng.standard_normal(n)
        e_o = rng.standard_normal(n); e_o.reshape(NY, NX)[[0, -1], :] = 0; e_o.reshape(NY, NX)[:, [0, -1]] = 0
        e_s = e_o.copy()
        for colour in (0, 1, 0):
            L.orc_mg_smooth(NX, NY, P(GE[l]), P(GN[l]), P(R), colour, omega, P(e_o))
            shim.shim_mg_smooth(lx[l], ly[l], P(GE[l]), P(GN[l]), P(e_s), P(R), colour, omega)
            assert np.array_equal(e_o, e_s), f"smooth level {l} colour {colour}"
        sc_o = [np.zeros(n), np.zeros(n)]; sc_s = np.zeros(5 * n)
        for d, par in ((0, 0), (0, 1), (1, 0), (1, 1), (0, 0)):      # zebra line relaxation
            L.orc_mg_line_pass(NX, NY, P(GE[l]), P(GN[l]), P(R), d, par, omega, P(e_o), P(sc_o[0]), P(sc_o[1]))
            shim.shim_mg_line_pass(lx[l], ly[l], P(GE[l]), P(GN[l]), P(e_s), P(R), P(sc_s), d, par, omega)
            assert np.array_equal(e_o, e_s), f"line pass level {l} dir {d} parity {par}"
        if l + 1 < nlev:
            nc = (lx[l + 1] + 2) * (ly[l + 1] + 2)
            Rc_o, Rc_s = np.zeros(nc), np.zeros(nc)
            L.orc_mg_restrict(NX, P(GE[l]), P(GN[l]), P(R), P(e_o), lx[l + 1] + 2, ly[l + 1] + 2, P(Rc_o))
            shim.shim_mg_restrict(lx[l], ly[l], P(GE[l]), P(GN[l]), P(e_s), P(R), lx[l + 1], ly[l + 1], P(Rc_s))
            assert np.array_equal(Rc_o, Rc_s), f"restrict level {l}"
            ec = rng.standard_normal(nc)
            L.orc_mg_prolong(NX, NY, P(GE[l]), P(GN[l]), lx[l + 1] + 2, P(ec), P(e_o))
            shim.shim_mg_prolong(lx[l + 1], ly[l + 1], None, None, P(ec), lx[l], ly[l], P(GE[l]), P(GN[l]), P(e_s), 0)
            assert np.array_equal(e_o, e_s), f"prolong level {l}"
            GEc, GNc = olevel(l + 1, 0), olevel(l + 1, 1)           # bilinear, coarse conductances as connectivity
            L.orc_mg_prolong2(NX, NY, P(GE[l]), P(GN[l]), lx[l + 1] + 2, P(ec), P(GEc), P(GNc), P(e_o))
            shim.shim_mg_prolong(lx[l + 1], ly[l + 1], P(GEc), P(GNc), P(ec), lx[l], ly[l], P(GE[l]), P(GN[l]), P(e_s), 1)
            assert np.array_equal(e_o, e_s), f"bilinear prolong level {l}"

    # ---- zebra line relaxation on the fine level (in place; scratch in the layout of p)
    p_o = p.copy(); pp2 = pad(p, nx, ny)
    so = [np.zeros(N), np.zeros(N)]; ss = [np.zeros(pp2.size), np.zeros(pp2.size), np.zeros(pp2.size)]
    for d, par in ((0, 0), (0, 1), (1, 0), (1, 1), (1, 0), (0, 1)):
        L.orc_ppe_line_pass(nx, ny, P(cP), P(cxm), P(cxp), P(cym), P(cyp), u8(ct), P(rhs), d, par, omega, P(p_o), P(so[0]), P(so[1]))
        shim.shim_line_pass(nx, ny, pitch, tarr, u8(ctp), P(rp), P(pp2), P(ss[0]), P(ss[1]), P(ss[2]), d, par, omega)
        assert np.array_equal(unpad(pp2, nx, ny), p_o), f"fine line pass dir {d} parity {par}"
    assert seed is not None or not np.array_equal(p_o, p)
    if bodies and seed is None:      # a line solve is exact along the line: the residual of the rows just relaxed with omega = 1 vanishes there
        q_o = p.copy()
        L.orc_ppe_line_pass(nx, ny, P(cP), P(cxm), P(cxp), P(cym), P(cyp), u8(ct), P(rhs), 0, 1, C.c_double(1.0), P(q_o), P(so[0]), P(so[1]))
        res = np.zeros(N); scratch = np.zeros(N)
        L.orc_ppe_sweep_general(nx, ny, P(cP), P(cxm), P(cxp), P(cym), P(cyp), u8(ct), P(rhs), P(q_o), P(scratch), P(res))
        r2 = res.reshape(ny, nx)
        assert np.abs(r2[1:-1:2, :]).max() < 1e-9 * np.abs(cP).max() and np.abs(r2[2:-1:2, :]).max() > 1e-3

    # ---- prolongation to the fine level
    e1 = rng.standard_normal(n1)
    p_o = p.copy()
    L.orc_mg_prolong_fine(nx, ny, u8(ct), lx[1] + 2, P(e1), P(p_o))
    shim.shim_mg_prolong_fine(nx, ny, pitch, u8(ctp), lx[1], ly[1], None, None, P(e1), P(pp), 0)
    assert np.array_equal(unpad(pp, nx, ny), p_o)
    assert not np.array_equal(p_o, p)
    L.orc_mg_prolong_fine2(nx, ny, u8(ct), lx[1] + 2, P(e1), P(GE[1]), P(GN[1]), P(p_o))
    shim.shim_mg_prolong_fine(nx, ny, pitch, u8(ctp), lx[1], ly[1], P(GE[1]), P(GN[1]), P(e1), P(pp), 1)
    assert np.array_equal(unpad(pp, nx, ny), p_o)
    # a constant coarse correction is reproduced exactly by both transfers, bodies or not
    one = np.ones(n1); q_o = np.zeros(N)
    L.orc_mg_prolong_fine2(nx, ny, u8(ct), lx[1] + 2, P(one), P(GE[1]), P(GN[1]), P(q_o))
    assert set(np.unique(q_o)) <= {0.0, 1.0} and q_o.sum() == (ct.reshape(ny, nx)[1:-1, 1:-1] == 1).sum()
    L.orc_mg_destroy(mg)
