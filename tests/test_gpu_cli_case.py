"""The C++ driver's case options: wall values, initial condition, moving bodies, force and probe
files, multigrid — checked against the oracle run with the same settings (the Tecplot text of bit-identical fields is
identical; the diagnostics files carry 13 significant digits)."""
import math
import os
import subprocess

import numpy as np
import pytest

import _oracle as orc
import immerseflow_b200 as ifx
from conftest import ROOT, fmt6, load_tecplot

pytestmark = pytest.mark.gpu

CLI = os.path.join(ROOT, "immerseflow_b200", "bin", "immerseflow")


def write_case(w, ncx, ncy, Lx, Ly, dt, Re, ppe_itermax, steps):
    (w / "src").mkdir(parents=True); (w / "results").mkdir(); (w / "inputs").mkdir()
    ref = open(os.path.join(ROOT, "tests", "golden", "reference_case", "inputs", "inputs.txt")).read()
    inp = ifx.read_input_file(os.path.join(ROOT, "tests", "golden", "reference_case", "inputs", "inputs.txt"))
    txt = ref.replace("50      50", f"{ncx}     {ncy}").replace("100000", f"{ppe_itermax}")
    (w / "inputs" / "inputs.txt").write_text(txt)
    got = ifx.read_input_file(str(w / "inputs" / "inputs.txt"))
    assert (got.nx, got.ny, got.PPE_itermax) == (ncx + 2, ncy + 2, ppe_itermax)
    xf, yf = ifx.uniform_faces(ncx, Lx), ifx.uniform_faces(ncy, Ly)
    for name, f in (("xgrid.dat2", xf), ("ygrid.dat2", yf)):
        with open(w / "inputs" / name, "w") as fh:
            fh.writelines(f"{k + 1:>10} {v:.7E}\n" for k, v in enumerate(f))
    return xf, yf, got


@pytest.mark.skipif(not os.path.exists(CLI), reason="CLI not built")
def test_cli_lid_driven_cavity_with_multigrid(tmp_path):
    """BASELINE.json configs[1] in small: lid u = 1 on the north wall, fluid at rest, multigrid Poisson."""
    w = tmp_path / "cavity"
    xf, yf, inp = write_case(w, 64, 64, 1.0, 1.0, None, None, 30, 5)
    r = subprocess.run([CLI, "--mode", "full", "--bc-u", "0,0,0,1", "--ic", "zero", "--ppe-solver", "4", "--ppe-omega", "1",
                        "--steps", "5"], cwd=w / "src", capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    o = orc.FullSolver(xf, yf, inp.dt, inp.Re, inp.AD_itermax, 30, ppe_tol=1e-6, ppe_abs=1, bc_u=(0.0, 0.0, 0.0, 1.0))
    o.set_ppe_solver(4, 1.0)
    o.update_ib()
    for _ in range(5):
        st = o.step()
    assert f"Poisson {int(st[3])} sweeps" in r.stdout.strip().splitlines()[-2]
    for name, f in (("uc.dat", "u"), ("vc.dat", "v"), ("p.dat", "p")):
        d = load_tecplot(w / "results" / name)
        assert np.array_equal(d[:, 2], fmt6(o.get(f))), name
    assert np.abs(load_tecplot(w / "results" / "uc.dat")[:, 2]).max() > 0.5         # the lid drags the fluid
    o.close()


@pytest.mark.skipif(not os.path.exists(CLI), reason="CLI not built")
def test_cli_moving_body_forces_and_probes(tmp_path):
    w = tmp_path / "moving"
    xf, yf, inp = write_case(w, 128, 64, 2.0, 1.0, None, None, 200, 4)
    m = orc.circle_markers(0.6, 0.5, 0.15, 48)
    ub, vb, ax, ay, fr = 0.3, 0.0, 0.0, 0.05, 2.0
    with open(w / "inputs" / "bodies.txt", "w") as f:
        f.write(f"1\n48 {ub} {vb} {ax} {ay} {fr}\n" + "".join(f"{x:.17g} {y:.17g}\n" for x, y in m))
    pts = np.array([[0.2, 0.5], [1.0, 0.52], [1.5, 0.3], [0.61, 0.7]])
    np.savetxt(w / "inputs" / "probes.txt", pts, fmt="%.17g")
    r = subprocess.run([CLI, "--mode", "full", "--ic", "uniform:1,0", "--bodies", "../inputs/bodies.txt", "--ppe-solver", "3",
                        "--ppe-omega", "1.7", "--steps", "4", "--forces", "../results/forces.dat", "--probes",
                        "../inputs/probes.txt", "--probe-out", "../results/probes.dat"],
                       cwd=w / "src", capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    o = orc.FullSolver(xf, yf, inp.dt, inp.Re, inp.AD_itermax, 200, ppe_tol=1e-6, ppe_abs=1)
    o.set_ppe_solver(3, 1.7)
    n = (128 + 2) * (64 + 2)
    o.set("u", np.ones(n)); o.set("v", np.zeros(n))
    F, PR = [], []
    two_pi = 6.283185307179586476925286766559
    # the driver classifies the bodies where they are at t = 0 (final_results.dat) before the first step
    for step in range(4):
        t = (step + 1) * inp.dt
        sn, cs = math.sin(two_pi * fr * t), math.cos(two_pi * fr * t)      # libm, like the driver
        dxb, dyb = ub * t + ax * sn, vb * t + ay * sn
        mk = np.ascontiguousarray(np.stack([m[:, 0] + dxb, m[:, 1] + dyb], axis=1))
        o.set_bodies([mk], [(ub + ax * two_pi * fr * cs, vb + ay * two_pi * fr * cs)])
        o.update_ib()
        o.step()
        F.append(o.body_forces(1)[0]); PR.append(np.stack(o.probe(pts[:, 0], pts[:, 1]), axis=1))
    got_f = np.loadtxt(w / "results" / "forces.dat")
    assert got_f.shape == (4, 7) and np.array_equal(got_f[:, 0], [1, 2, 3, 4])
    assert np.allclose(got_f[:, 3:], np.array(F), rtol=1e-11, atol=1e-13)
    got_p = np.loadtxt(w / "results" / "probes.dat")
    assert got_p.shape == (16, 6)
    assert np.allclose(got_p[:, 3:], np.concatenate(PR), rtol=1e-11, atol=1e-13)
    for name, f in (("uc.dat", "u"), ("vc.dat", "v"), ("p.dat", "p")):
        assert np.array_equal(load_tecplot(w / "results" / name)[:, 2], fmt6(o.get(f))), name
    o.close()
