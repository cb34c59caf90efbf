#!/usr/bin/env python
"""Regression fixtures for the UNPINNED stages (IFX_COMPAT_FULL): inputs and the oracle's outputs of a small case per Poisson
solver, frozen to a file.  The reference has no code for these stages, so these are NOT reference vectors — they freeze
what oracle/ifx_oracle_full.c + ifx_oracle_mg.c define today, so that (a) a change of the oracle's semantics between rounds
is a visible, deliberate act (this script is re-run and the diff committed) and (b) the GPU path is also held to a file,
not only to whatever the oracle computes at test time.  Inputs (grid faces, body markers) are stored with the outputs:
nothing in the comparison depends on libm.

    python tests/golden/full_mode/make_fixtures.py        # rewrites tests/golden/full_mode/full_mode.npz
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
import _oracle as orc  # noqa: E402

CASES = {                       # name: (PPE_Solver, w-PPE, PPE_itermax, ppe_tol)
    "jacobi": (1, 1.0, 60, 1e-5),
    "line_sor": (2, 1.6, 40, 1e-5),
    "rb_sor": (3, 1.7, 80, 1e-5),
    "multigrid_point": (4, 1.0, 25, 1e-5),
    "multigrid_line": (5, 1.0, 25, 1e-5),
}
NCX, NCY, STEPS, DT, RE, AD_ITERMAX = 75, 50, 3, 2e-3, 120.0, 25


def inputs():
    xf, yf = orc.stretched_faces(NCX, 3.0, 1.03), orc.stretched_faces(NCY, 2.0, 1.03)
    bodies = [orc.circle_markers(1.1, 1.0, 0.28, 40), orc.ellipse_markers(2.1, 0.8, 0.3, 0.11, 0.5, 36)]
    vel = np.array([[0.0, 0.0], [0.15, -0.05]])
    return xf, yf, bodies, vel


def run(xf, yf, bodies, vel, solver, omega, itermax, tol):
    s = orc.FullSolver(xf, yf, DT, RE, AD_ITERMAX, itermax, ppe_tol=tol, ppe_abs=1, bc_u=(1.0, 1.0, 1.0, 1.0))
    s.set_ppe_solver(solver, omega)
    n = (NCX + 2) * (NCY + 2)
    s.set("u", np.ones(n)); s.set("v", np.zeros(n))
    s.set_bodies(bodies, [tuple(v) for v in vel])
    s.update_ib()
    counts = []
    for _ in range(STEPS):
        st = s.step()
        counts.append((int(st[0]), int(st[3])))
    out = {"u": s.get("u"), "v": s.get("v"), "p": s.get("p"), "celltype": s.get("celltype").astype(np.uint8),
           "counts": np.array(counts, dtype=np.int32), "ghost_cells": s.ghost_cells()["cell"],
           "forces": s.body_forces(len(bodies))}
    s.close()
    return out


def build():
    xf, yf, bodies, vel = inputs()
    d = {"xf": xf, "yf": yf, "vel": vel, "nbodies": np.array(len(bodies))}
    for k, m in enumerate(bodies):
        d[f"markers{k}"] = m
    for name, (solver, omega, itermax, tol) in CASES.items():
        for key, val in run(xf, yf, bodies, vel, solver, omega, itermax, tol).items():
            d[f"{name}/{key}"] = val
    return d


if __name__ == "__main__":
    d = build()
    np.savez_compressed(os.path.join(HERE, "full_mode.npz"), **d)
    print({k: (v.shape if hasattr(v, "shape") else v) for k, v in d.items() if k.endswith("counts")},
          {k: d[k].tolist() for k in d if k.endswith("counts")})
