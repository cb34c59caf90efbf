#!/usr/bin/env python
"""Write the input files of BASELINE.json's configurations in the reference's own formats, and print the driver command.

    python tools/make_case.py cavity   --out /tmp/cavity   [--scale 0.25]
    python tools/make_case.py cylinder --out /tmp/cyl
    python tools/make_case.py airfoil  --out /tmp/foil
    python tools/make_case.py bodies   --out /tmp/bodies

A case directory is laid out like the reference tree (`inputs/`, `results/`, `src/` = the working directory):
`inputs/inputs.txt` in the keyword / value-line grammar of `src/main.cu:17-52`, `inputs/xgrid.dat2`, `inputs/ygrid.dat2`
as `index value` pairs (`src/include/preSim.cu:275-290`, written like `inputs/uniformGrid.py:10`), `inputs/bodies.txt`
for the driver's `--bodies`.  `--scale s` multiplies the cell counts (to try a case on a small grid first).

| case | BASELINE.json config | grid | bodies |
|---|---|---|---|
| cavity   | configs[1] lid-driven cavity Re = 1000 | uniform 1024 x 1024 on 1 x 1 | — (lid u = 1 on the north wall) |
| cylinder | configs[2] circular cylinder Re = 300 | stretched 4096 x 2048 on 40 x 20, uniform core around the body | circle, D = 1 |
| airfoil  | configs[3] elliptic airfoil Re = 300 | stretched 8192 x 8192 on 40 x 40 | ellipse 1 x 0.12 at 10 degrees |
| bodies   | configs[4] several complex bodies Re = 1000, re-classified every step | uniform 16384 x 16384 on 1 x 1 | 8 lobed bodies, each translating |

Reynolds numbers: the predictor keeps the reference's factor 1/2 on the convective fluxes (ADSolver.cu:66-73), which makes
the input file's `Re` twice the physical Reynolds number of the computed flow (tests/test_oracle_physics.py); the files
written here carry 2 x the number in the table above.
"""
from __future__ import annotations

import argparse
import os

import numpy as np


def uniform_faces(n: int, length: float) -> np.ndarray:
    return np.linspace(0.0, length, n + 1)


def _growth_ratio(first: float, n: int, length: float) -> float:
    """r >= 1 with first * (r + r^2 + ... + r^n) = length"""
    if n <= 0:
        raise ValueError("no cells to stretch over")
    if first * n >= length:
        return 1.0
    lo, hi = 1.0, 4.0
    for _ in range(200):
        r = 0.5 * (lo + hi)
        try:
            tot = first * n if abs(r - 1.0) < 1e-14 else first * r * (r ** n - 1.0) / (r - 1.0)
        except OverflowError:                      # r ** n beyond the double range: far too long
            tot = float("inf")
        lo, hi = (r, hi) if tot < length else (lo, r)
    return 0.5 * (lo + hi)


def stretched_faces(n: int, length: float, core_lo: float, core_hi: float, n_core: int) -> np.ndarray:
    """In the spirit of the shipped inputs/xgrid.dat: `n_core` uniform cells on [core_lo, core_hi] (the body), spacing growing
    geometrically towards both ends of [0, length]; the remaining cells go to the two sides in proportion to the
    logarithm of their lengths, which keeps the two growth ratios close."""
    if not (0.0 < core_lo < core_hi < length) or not (2 <= n_core < n - 1):
        raise ValueError("bad core")
    d0 = (core_hi - core_lo) / n_core
    rest = n - n_core
    wl, wr = np.log1p(core_lo / d0), np.log1p((length - core_hi) / d0)
    nl = int(np.clip(round(rest * wl / (wl + wr)), 1, rest - 1))
    nr = rest - nl
    rl, rr = _growth_ratio(d0, nl, core_lo), _growth_ratio(d0, nr, length - core_hi)
    left = d0 * rl ** np.arange(nl, 0, -1)
    right = d0 * rr ** np.arange(1, nr + 1)
    left *= core_lo / left.sum(); right *= (length - core_hi) / right.sum()
    d = np.concatenate([left, np.full(n_core, d0), right])
    f = np.concatenate([[0.0], np.cumsum(d)])
    f[-1] = length
    return f


def lobed_bodies(nb: int):
    """the moving bodies of bench.py's default workload (same formula)"""
    cols = int(np.ceil(np.sqrt(nb)))
    rows = (nb + cols - 1) // cols
    th = 2.0 * np.pi * np.arange(256) / 256
    out = []
    for b in range(nb):
        cx, cy = (b % cols + 0.5) / cols, (b // cols + 0.5) / rows
        r0 = 0.22 / max(cols, rows)
        ub, vb = 0.05 * (1 if b % 2 == 0 else -1), 0.03 * (1 if (b // 2) % 2 == 0 else -1)
        r = r0 * (1.0 + 0.25 * np.cos((3 + b % 4) * th + 0.3 * b))
        out.append((np.stack([cx + r * np.cos(th), cy + r * np.sin(th)], axis=1), (ub, vb, 0.0, 0.0, 0.0)))
    return out


def write_grid(path: str, faces: np.ndarray) -> None:
    with open(path, "w") as f:
        f.writelines(f"{k + 1:>10} {v:.7E}\n" for k, v in enumerate(faces))      # inputs/uniformGrid.py:10


def write_inputs(path: str, ncx: int, ncy: int, lx: float, ly: float, dt: float, re_file: float, steps: int, ad_itermax: int,
                 ppe_itermax: int, ppe_solver: int, w_ppe: int, write_interval: int) -> None:
    with open(path, "w") as f:
        f.write(f"""===============================| INPUT FILE |====================================
Restart     Restart_Time
0           0

___________________________| Domain Information |________________________________
nx      ny 
{ncx}      {ncy}

Lx      Ly 
{lx:g}      {ly:g}


________________________| Iterative Solver Settings |____________________________
w-AD    w-PPE   AD-itermax  PPE-itermax  AD_Solver PPE_Solver(1. Point GS, 2. Line SOR)
1       {w_ppe}       {ad_itermax}          {ppe_itermax}          1         {ppe_solver}


___________________________| Simulation Settings |_______________________________
ErrorMax    tmax    dt       Re      mu
1E-6        {steps}       {dt:g}    {re_file:g}    0.01

___________________________| Data Write |_______________________________
Write Interval(t/dt)
{write_interval}
""")


def write_bodies(path: str, bodies) -> None:
    with open(path, "w") as f:
        f.write(f"{len(bodies)}\n")
        for m, (ub, vb, ax, ay, fr) in bodies:
            f.write(f"{len(m)} {ub!r} {vb!r} {ax!r} {ay!r} {fr!r}\n")
            f.writelines(f"{x:.17g} {y:.17g}\n" for x, y in m)


def build(case: str, out: str, scale: float = 1.0, steps: int = 100):
    n = lambda c: max(8, int(round(c * scale / 2.0)) * 2)          # even cell counts (any count works; even ones keep every coarse cell 2 x 2)
    os.makedirs(os.path.join(out, "inputs"), exist_ok=True)
    os.makedirs(os.path.join(out, "results"), exist_ok=True)
    os.makedirs(os.path.join(out, "src"), exist_ok=True)
    bodies, extra = [], []
    if case == "cavity":
        ncx = ncy = n(1024)
        lx = ly = 1.0
        xf, yf = uniform_faces(ncx, lx), uniform_faces(ncy, ly)
        dt, re_phys, solver = 0.25 * lx / ncx, 1000.0, 4          # CFL 0.25 under the lid; uniform cells: point-smoothed cycle
        extra = ["--bc-u", "0,0,0,1", "--ic", "zero"]
    elif case == "cylinder":
        ncx, ncy, lx, ly = n(4096), n(2048), 40.0, 20.0
        xf = stretched_faces(ncx, lx, 9.0, 13.0, n(1024))          # D/dx = 256 at full size
        yf = stretched_faces(ncy, ly, 8.5, 11.5, n(768))
        t = 2.0 * np.pi * np.arange(512) / 512
        bodies = [(np.stack([10.0 + 0.5 * np.cos(t), 10.0 + 0.5 * np.sin(t)], axis=1), (0.0,) * 5)]
        dt, re_phys, solver = 0.25 * (xf[1:] - xf[:-1]).min(), 300.0, 5
        extra = ["--ic", "uniform:1,0"]
    elif case == "airfoil":
        ncx = ncy = n(8192)
        lx = ly = 40.0
        xf = stretched_faces(ncx, lx, 9.0, 12.0, n(3072))
        yf = stretched_faces(ncy, ly, 19.0, 21.0, n(2048))
        t = 2.0 * np.pi * np.arange(1024) / 1024
        a, b, ang = 0.5, 0.06, -np.deg2rad(10.0)
        x, y = a * np.cos(t), b * np.sin(t)
        bodies = [(np.stack([10.0 + np.cos(ang) * x - np.sin(ang) * y, 20.0 + np.sin(ang) * x + np.cos(ang) * y], axis=1), (0.0,) * 5)]
        dt, re_phys, solver = 0.25 * (xf[1:] - xf[:-1]).min(), 300.0, 5
        extra = ["--ic", "uniform:1,0"]
    elif case == "bodies":
        ncx = ncy = n(16384)
        lx = ly = 1.0
        xf, yf = uniform_faces(ncx, lx), uniform_faces(ncy, ly)
        bodies = lobed_bodies(8)
        dt, re_phys, solver = 0.25 * lx / ncx, 1000.0, 4
    else:
        raise ValueError(case)
    write_grid(os.path.join(out, "inputs", "xgrid.dat2"), xf)
    write_grid(os.path.join(out, "inputs", "ygrid.dat2"), yf)
    write_inputs(os.path.join(out, "inputs", "inputs.txt"), ncx, ncy, lx, ly, dt, 2.0 * re_phys, steps, 25, 50, solver, 1, steps)
    # the reference's residuals are un-normalised sums over all cells: a tolerance per cell, times the cell count
    # the reference's residuals are un-normalised sums and its loops start from a residual of 1.0 (PPESolver.cu:170-172,
    # ADSolver.cu:313-315): a tolerance of 1 (2 for the predictor) or more would mean "no iteration", so the per-cell scaling
    # of the reference's 1e-6 on its 52 x 52 case is capped below that
    cmd = ["immerseflow", "--mode", "full", "--ppe-tol", f"{min(0.5, 1e-9 * ncx * ncy / dt):.3g}",
           "--ad-tol", f"{min(0.5, 4e-10 * ncx * ncy):.3g}"] + extra
    if bodies:
        write_bodies(os.path.join(out, "inputs", "bodies.txt"), bodies)
        cmd += ["--bodies", "../inputs/bodies.txt", "--forces", "../results/forces.dat"]
    return {"cells": (ncx, ncy), "dt": dt, "Re_file": 2.0 * re_phys, "xf": xf, "yf": yf, "bodies": bodies,
            "command": f"cd {os.path.join(out, 'src')} && " + " ".join(cmd)}


if __name__ == "__main__":
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("case", choices=["cavity", "cylinder", "airfoil", "bodies"])
    ap.add_argument("--out", required=True)
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--steps", type=int, default=100)
    a = ap.parse_args()
    r = build(a.case, a.out, a.scale, a.steps)
    dx = r["xf"][1:] - r["xf"][:-1]; dy = r["yf"][1:] - r["yf"][:-1]
    print(f"{a.case}: {r['cells'][0]} x {r['cells'][1]} cells, dx {dx.min():.3g} .. {dx.max():.3g}, dy {dy.min():.3g} .. {dy.max():.3g}, "
          f"dt = {r['dt']:.3g}, Re (file) = {r['Re_file']:g}, {len(r['bodies'])} bodies")
    print(r["command"])
