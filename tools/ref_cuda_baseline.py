#!/usr/bin/env python
"""Time the REFERENCE's own CUDA build (oracle/_ref/immerseFlow_ref: its unmodified translation units, file output
stubbed) on one GPU — the "reference CUDA build on one B200" baseline of north_star.  Reported, not a target.

Per-step time = the spacing of the reference's own `iter = <AD_itermax> ...` lines (the last Jacobi iteration of every
time step, ADSolver.cu:369), time-stamped as they arrive on a line-buffered pipe (`stdbuf -oL`): process start-up,
allocation and grid I/O are outside.  (Differencing the wall time of two runs does not work: start-up takes 5-6 s at
16384 x 16384 and varies by more than a step from run to run.)
The reference runs only its predictor per time step (src/main.cu:93-96), with its own per-step cudaMalloc/cudaFree
(ADSolver.cu:275-287) and a cudaDeviceSynchronize after every launch — all of that is part of what is timed.
"""
import argparse
import json
import os
import shutil
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref", "immerseFlow_ref")

INPUTS = """===============================| INPUT FILE |====================================
Restart     Restart_Time
0           9

___________________________| Domain Information |________________________________
nx      ny 
{nx}      {ny}

Lx      Ly 
10      5                                      


________________________| Iterative Solver Settings |____________________________
w-AD    w-PPE   AD-itermax  PPE-itermax  AD_Solver PPE_Solver(1. Point GS, 2. Line SOR)
1       1       {itmax}          100000          1         1


___________________________| Simulation Settings |_______________________________
ErrorMax    tmax    dt       Re      mu
1E-6        {tmax}       {dt}    {Re}    0.01

___________________________| Data Write |_______________________________
Write Interval(t/dt)
1000
"""


def write_grid(path, n_cells):
    mesh = np.linspace(0, 1, n_cells + 1)
    with open(path, "w") as f:                      # inputs/uniformGrid.py:10
        for i, v in enumerate(mesh):
            f.write(f"{i + 1:>10} {v:.7E}\n")


def run(nx, ny, itmax, dt, Re, tmax, workdir):
    """-> (wall seconds, number of `iter = ` lines, arrival times of the `iter = <itmax>` lines = end of every step)"""
    with open(os.path.join(workdir, "inputs", "inputs.txt"), "w") as f:
        f.write(INPUTS.format(nx=nx, ny=ny, itmax=itmax, tmax=tmax, dt=dt, Re=Re))
    cmd = [REF]
    if shutil.which("stdbuf"):
        cmd = ["stdbuf", "-oL"] + cmd
    t0 = time.perf_counter()
    p = subprocess.Popen(cmd, cwd=os.path.join(workdir, "src"), env=dict(os.environ, IFX_REF_SAVE="none"),
                         stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, bufsize=1)
    iters, ends = 0, []
    last = f"iter = {itmax} "
    for line in p.stdout:
        if line.startswith("iter = "):
            iters += 1
            if line.startswith(last):
                ends.append(time.perf_counter() - t0)
    err = p.stderr.read()
    rc = p.wait()
    t = time.perf_counter() - t0
    if rc != 0:
        raise RuntimeError(err[-500:])
    return t, iters, ends


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nx", type=int, default=16384)
    ap.add_argument("--ny", type=int, default=16384)
    ap.add_argument("--ad-itermax", type=int, default=25)
    ap.add_argument("--dt", type=float, default=1e-3)
    ap.add_argument("--Re", type=float, default=150.0)
    ap.add_argument("--steps-a", type=int, default=4)
    ap.add_argument("--steps-b", type=int, default=1)
    a = ap.parse_args()
    if not os.path.exists(REF):
        print(json.dumps({"unavailable": "oracle/_ref/immerseFlow_ref not built (make -C oracle ref, needs /root/reference)"}))
        return
    w = tempfile.mkdtemp(prefix="ifx_ref_")
    try:
        for d in ("src", "inputs", "results"):
            os.makedirs(os.path.join(w, d))
        write_grid(os.path.join(w, "inputs", "xgrid.dat2"), a.nx)
        write_grid(os.path.join(w, "inputs", "ygrid.dat2"), a.ny)
        ta, ia, ends = run(a.nx, a.ny, a.ad_itermax, a.dt, a.Re, a.steps_a, w)
        if len(ends) < 2:
            raise RuntimeError("the reference did not report the end of at least two steps (converged before AD_itermax?)")
        per_step = (ends[-1] - ends[0]) / (len(ends) - 1)
        k = ia / a.steps_a
        print(json.dumps({"impl": "reference CUDA build (unmodified TUs, nvcc -arch=sm_100, file output stubbed)",
                          "grid": [a.nx, a.ny], "predictor_iterations_per_step": k, "s_per_step": per_step,
                          "Mcell_steps_per_s": a.nx * a.ny / per_step / 1e6,
                          "note": "the reference's time step is the predictor only (src/main.cu:93-96)",
                          "step_end_times_s": ends, "wall_s": ta}))
    finally:
        shutil.rmtree(w, ignore_errors=True)


if __name__ == "__main__":
    main()
