"""Multigrid (PPE_Solver 4) against red-black SOR (PPE_Solver 3) on the lid-driven cavity, uniform grids: V-cycles /
iterations, residual reached and CUDA-event time of the Poisson stage per step.  Writes one JSON line per case."""
import json
import os
import sys

import numpy as np

import immerseflow_b200 as ifx


def run(n, solver, omega, itermax, tol, steps):
    xf = yf = ifx.uniform_faces(n, 1.0)
    inp = ifx.make_input(n, n, 1.0 / n / 4, 1000.0, AD_itermax=25, PPE_itermax=itermax)
    bc = {"u_bc_w": 0.0, "u_bc_e": 0.0, "u_bc_s": 0.0, "u_bc_n": 1.0}
    out = []
    with ifx.ImmerseFlow(inp, xf, yf, compat=ifx.IFX_COMPAT_FULL, bc=bc, ppe_abs_residual=1, ppe_tol=tol, ppe_solver=solver,
                         ppe_omega=omega, use_graphs=int(os.environ.get("IFX_MG_BENCH_GRAPHS", "0"))) as s:
        s.initializeData()
        z = np.zeros(s.field_size("u"))
        s.set("u", z); s.set("v", z); s.set("p", z)
        for _ in range(steps):
            st = s.step()
            out.append({"n": n, "solver": solver, "omega": omega, "iterations": st.ppe_sweeps, "residual": st.ppe_residual,
                        "ms_ppe": st.ms_ppe, "ms_ad": st.ms_ad, "ad_iters": st.ad_iters,
                        "graphs": int(os.environ.get("IFX_MG_BENCH_GRAPHS", "0"))})
    return out


if __name__ == "__main__":
    cases = [(1024, 4, 1.0, 60, 1e-3, 3), (1024, 3, 1.99, 4000, 1e-3, 2), (4096, 4, 1.0, 60, 1e-1, 2)]
    for c in cases:
        for r in run(*c):
            print(json.dumps(r), flush=True)
