"""Diagnostics: the probe kernel (immerseflow_b200/csrc/kernels_diag.cu) and the host arithmetic of the
surface forces (csrc/diag.cuh) compiled through tests/shim/cuda_host_shim.h and compared bit for bit with
oracle/ifx_oracle_diag.c; and the oracle's force definition checked against the divergence theorem."""
import ctypes as C
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import _oracle as orc  # noqa: E402
from _oracle import P, PI  # noqa: E402
from shim.build import build  # noqa: E402
from test_mg_shim import pad, pitch_of, u8  # noqa: E402


@pytest.fixture(scope="module")
def shim():
    return C.CDLL(build("diag", deps=("kernels_diag.cu", "diag.cuh", "common.cuh")))


def bodies_case():
    xf, yf = orc.stretched_faces(96, 4.0, 1.02), orc.stretched_faces(64, 2.0, 1.02)
    g = orc.Grid(xf, yf)
    bodies = [orc.circle_markers(1.5, 1.0, 0.3, 64), orc.ellipse_markers(2.6, 0.9, 0.35, 0.12, 0.5, 50)]
    offs = np.array([0, 64, 114], dtype=np.int32)
    m = np.concatenate(bodies)
    xm, ym = np.ascontiguousarray(m[:, 0]), np.ascontiguousarray(m[:, 1])
    N = g.nx * g.ny
    ct = np.ones(N, dtype=np.uint8); body_of = np.zeros(N, dtype=np.int32)
    orc.lib().orc_iblank_classify(g.nx, g.ny, P(g.xc), P(g.yc), 2, PI(offs), P(xm), P(ym), u8(ct), PI(body_of))
    return g, ct, offs, xm, ym


def test_probe_kernel_and_force_arithmetic_match_oracle(shim):
    L = orc.lib()
    g, ct, offs, xm, ym = bodies_case()
    nx, ny, N = g.nx, g.ny, g.nx * g.ny
    rng = np.random.default_rng(11)
    u, v, p = rng.standard_normal(N), rng.standard_normal(N), rng.standard_normal(N)
    # points everywhere: inside the domain, hugging the bodies, outside the domain (clamped boxes)
    n = 4000
    px = np.concatenate([rng.uniform(-0.2, 4.2, n), 1.5 + 0.33 * np.cos(np.linspace(0, 6.28, 200))])
    py = np.concatenate([rng.uniform(-0.2, 2.2, n), 1.0 + 0.33 * np.sin(np.linspace(0, 6.28, 200))])
    n = px.size
    out_o = [np.zeros(n) for _ in range(3)]; out_s = [np.zeros(n) for _ in range(3)]
    L.orc_probe(nx, ny, P(g.xc), P(g.yc), u8(ct), P(u), P(v), P(p), n, P(px), P(py), *[P(a) for a in out_o])
    ctp = pad(ct, nx, ny, dtype=np.uint8, fill=1)
    shim.shim_probe(nx, ny, pitch_of(nx), P(g.xc), P(g.yc), u8(ctp), P(pad(u, nx, ny)), P(pad(v, nx, ny)), P(pad(p, nx, ny)),
                    n, P(px), P(py), *[P(a) for a in out_s])
    for a, b in zip(out_o, out_s):
        assert np.array_equal(a, b)
    # force geometry + sums
    ns = int(offs[-1])
    geo_o, geo_s = np.zeros(8 * ns), np.zeros(8 * ns)
    L.orc_force_geometry(nx, ny, P(g.xc), P(g.yc), ns, PI(offs), 2, P(xm), P(ym), P(geo_o))
    shim.shim_force_geometry(nx, ny, P(g.xc), P(g.yc), 2, PI(offs), P(xm), P(ym), P(geo_s))
    assert np.array_equal(geo_o, geo_s)
    pu, pv, pp = rng.standard_normal(2 * ns), rng.standard_normal(2 * ns), rng.standard_normal(2 * ns)
    ub, vb = np.array([0.3, -0.1]), np.array([0.0, 0.2])
    F_o, F_s = np.zeros(8), np.zeros(8)
    L.orc_force_sum(2, PI(offs), P(geo_o), P(pu), P(pv), P(pp), P(ub), P(vb), C.c_double(150.0), P(F_o))
    shim.shim_force_sum(2, PI(offs), P(geo_s), P(pu), P(pv), P(pp), P(ub), P(vb), C.c_double(150.0), P(F_s))
    assert np.array_equal(F_o, F_s) and np.abs(F_o).min() > 0


def polygon_area(x, y):
    return 0.5 * abs(np.dot(x, np.roll(y, -1)) - np.dot(y, np.roll(x, -1)))


def test_pressure_force_obeys_the_divergence_theorem():
    """p = a x + b y is reproduced exactly by bilinear interpolation, and  -closed-integral p n ds = -(a, b) * area  for any
    polygon; a uniform flow past a body at rest has no velocity gradient, hence no viscous force."""
    g, ct, offs, xm, ym = bodies_case()
    X, Y = np.meshgrid(g.xc, g.yc)
    a, b = 0.7, -1.3
    p = np.ascontiguousarray((a * X + b * Y).reshape(-1))
    u = np.zeros_like(p); v = np.zeros_like(p)
    F = np.zeros(8)
    orc.lib().orc_body_forces(g.nx, g.ny, P(g.xc), P(g.yc), u8(ct), C.c_double(100.0), 2, PI(offs), P(xm), P(ym), None, None,
                              P(u), P(v), P(p), P(F))
    for k in range(2):
        area = polygon_area(xm[offs[k]:offs[k + 1]], ym[offs[k]:offs[k + 1]])
        assert np.allclose(F[4 * k:4 * k + 2], [-a * area, -b * area], rtol=0, atol=1e-12)
        assert np.array_equal(F[4 * k + 2:4 * k + 4], [0.0, 0.0])
    # a body moving with a uniform stream: no velocity gradient, no force at all
    u = np.full_like(p, 0.8)
    ub = np.array([0.8, 0.8]); vb = np.zeros(2)
    orc.lib().orc_body_forces(g.nx, g.ny, P(g.xc), P(g.yc), u8(ct), C.c_double(100.0), 2, PI(offs), P(xm), P(ym), P(ub), P(vb),
                              P(u), P(v), P(np.zeros_like(p)), P(F))
    assert np.abs(F).max() < 1e-12
