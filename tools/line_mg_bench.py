"""Line-smoothed multigrid (PPE_Solver 5) against red-black SOR (3) and the point-smoothed cycle (4) on a stretched grid
with a cylinder (BASELINE.json configs[2] in small): iterations, residual reached, CUDA-event time of the Poisson stage."""
import json
import sys

import numpy as np

import immerseflow_b200 as ifx


def stretched_faces(n_cells, length, total_ratio):
    core = max(2, n_cells // 3)
    side = n_cells - core
    left, right = side // 2, side - side // 2
    ratio = total_ratio ** (3.0 / n_cells)
    d = np.concatenate([ratio ** np.arange(left, 0, -1), np.ones(core), ratio ** np.arange(1, right + 1)])
    f = np.concatenate([[0.0], np.cumsum(d)])
    f *= length / f[-1]
    return np.array([float(f"{v:.7E}") for v in f])


def run(ncx, ncy, solver, omega, itermax, tol, steps):
    xf, yf = stretched_faces(ncx, 10.0, 25.0), stretched_faces(ncy, 5.0, 12.0)
    inp = ifx.make_input(ncx, ncy, 2e-3, 300.0, AD_itermax=25, PPE_itermax=itermax)
    out = []
    with ifx.ImmerseFlow(inp, xf, yf, compat=ifx.IFX_COMPAT_FULL, ppe_abs_residual=1, ppe_tol=tol, ppe_solver=solver,
                         ppe_omega=omega) as s:
        t = 2 * np.pi * np.arange(256) / 256
        s.set_bodies([np.ascontiguousarray(np.stack([5.0 + 0.5 * np.cos(t), 2.5 + 0.5 * np.sin(t)], axis=1))])
        s.initializeData()
        n = s.field_size("u")
        s.set("u", np.ones(n)); s.set("v", np.zeros(n)); s.set("p", np.zeros(n))
        for _ in range(steps):
            st = s.step()
            out.append({"grid": [ncx, ncy], "solver": solver, "omega": omega, "iterations": st.ppe_sweeps,
                        "residual": st.ppe_residual, "ms_ppe": st.ms_ppe})
    return out


if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
    cases = [(n, n // 2, 5, 1.0, 30, 1e-3, 2), (n, n // 2, 4, 1.0, 30, 1e-3, 1), (n, n // 2, 3, 1.9, 2000, 1e-3, 1)]
    if len(sys.argv) > 2:
        cases = [c for c in cases if c[2] == int(sys.argv[2])]
    for c in cases:
        for r in run(*c):
            print(json.dumps(r), flush=True)
