#!/usr/bin/env python
"""Drag of a circular cylinder at low Reynolds number with the CPU oracle (test infrastructure): physics anchor for the
UNPINNED immersed-boundary stages and the force diagnostic (the reference has no code for either).

    python tools/cylinder_drag.py [Re] [T] [scale]         # defaults 20, 40, 1/16

The `cylinder` case of tools/make_case.py (stretched 40 x 20 domain, D = 1 in a uniform core; scale 1/16 = 256 x 128
cells, 16 cells per diameter), uniform inflow u = 1 imposed on all four sides like the reference's BCs, line-smoothed
multigrid for the Poisson equation.  The predictor keeps the reference's factor 1/2 on the convective fluxes, i.e. it
integrates the Navier-Stokes equation for w = u/2 at half the input file's Reynolds number; in terms of the force F
computed from (u, p) the drag coefficient of that flow is  Cd = 2 (F/2) / ((1/2)^2 D) = 4 F  (tests/test_oracle_physics.py).

Measured (Re = 20, t = 40, 256 x 128 cells):  Cd = 2.049 (pressure 1.223 + friction 0.825).
Literature, steady flow at Re = 20: Cd = 2.045 (Dennis & Chang, J. Fluid Mech. 42, 1970: 1.233 + 0.812); 2.09 (Tritton 1959, exp.).

Unsteady wake (Re = 100, t = 160, same grid, started with a small asymmetric kick: run(100, 160, 1/16, kick=0.1)): periodic
vortex shedding with period 11.72 in the solver's time, i.e. St = f D / (U/2) = 0.171; mean Cd = 1.275 (0.973 + 0.301),
lift amplitude 0.288, drag amplitude 0.0075.  Literature at Re = 100: St = 0.164-0.166 (Williamson 1989), mean
Cd = 1.33-1.38, lift amplitude 0.30-0.34 — within 4 % / 5 % / 10 % at 16 cells per diameter with u = 1 imposed on walls
10 diameters away.
"""
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import _oracle as orc  # noqa: E402
import make_case  # noqa: E402


def run(re_phys=20.0, t_end=40.0, scale=1.0 / 16, report=None, report_every=100, kick=0.0):
    with tempfile.TemporaryDirectory() as d:
        r = make_case.build("cylinder", d, scale, steps=1)
    xf = np.array([float(f"{v:.7E}") for v in r["xf"]]); yf = np.array([float(f"{v:.7E}") for v in r["yf"]])
    s = orc.FullSolver(xf, yf, r["dt"], 2.0 * re_phys, 25, 50, ppe_tol=1e-6)
    n = (len(xf) + 1) * (len(yf) + 1)
    v0 = np.zeros(n)
    if kick:                            # a small asymmetric disturbance to start vortex shedding without waiting for round-off to grow
        g = orc.Grid(xf, yf)
        X, Y = np.meshgrid(g.xc, g.yc)
        v0 = np.ascontiguousarray((kick * np.exp(-((X - 11.5) ** 2 + (Y - 10.0) ** 2))).reshape(-1))
    s.set("u", np.ones(n)); s.set("v", v0)
    s.set_bodies([r["bodies"][0][0]]); s.update_ib(); s.set_ppe_solver(5, 1.0)
    nsteps = int(t_end / r["dt"])
    hist = []
    for k in range(nsteps):
        st = s.step()
        if k % report_every == report_every - 1 or k == nsteps - 1:
            F = s.body_forces(1)[0]
            hist.append(((k + 1) * r["dt"], 4 * (F[0] + F[2]), 4 * F[0], 4 * F[2], 4 * (F[1] + F[3]), int(st[3])))
            if report:
                report(hist[-1])
    s.close()
    return hist


if __name__ == "__main__":
    a = sys.argv[1:]
    t0 = time.time()
    run(float(a[0]) if a else 20.0, float(a[1]) if len(a) > 1 else 40.0, float(a[2]) if len(a) > 2 else 1.0 / 16,
        report=lambda h: print("t = %6.2f  Cd = %.4f (pressure %.4f + friction %.4f)  Cl = %+.1e  V-cycles %d  [%.0f s]"
                               % (h + (time.time() - t0,)), flush=True))
