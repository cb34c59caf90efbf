"""Probes and body forces on the GPU (ifx_probe / ifx_body_forces) bit for bit against oracle/ifx_oracle_diag.c after a
few steps of flow past two bodies.  PARITY UNPINNED (the reference has no diagnostics)."""
import numpy as np
import pytest

import immerseflow_b200 as ifx
import _oracle as orc
from test_gpu_full_parity import pair

pytestmark = pytest.mark.gpu


def test_probes_and_forces_match_oracle():
    xf, yf = orc.stretched_faces(96, 4.0, 1.02), orc.stretched_faces(64, 2.0, 1.02)
    g, o = pair(xf, yf, 2e-3, 100.0, 25, 200, ppe_solver=3, ppe_omega=1.7)
    o.set_ppe_solver(3, 1.7)
    with g:
        bodies = [orc.circle_markers(1.5, 1.0, 0.3, 64), orc.ellipse_markers(2.6, 0.9, 0.35, 0.12, 0.5, 50)]
        vel = [(0.0, 0.0), (0.2, -0.1)]
        g.set_bodies(bodies, vel); o.set_bodies(bodies, vel)
        g.initializeData()
        n = g.field_size("u")
        g.set("u", np.ones(n)); g.set("v", np.zeros(n)); g.set("p", np.zeros(n))
        o.set("u", np.ones(n)); o.set("v", np.zeros(n)); o.update_ib()
        rng = np.random.default_rng(3)
        px = np.concatenate([rng.uniform(-0.1, 4.1, 3000), 1.5 + 0.32 * np.cos(np.linspace(0, 6.28, 100))])
        py = np.concatenate([rng.uniform(-0.1, 2.1, 3000), 1.0 + 0.32 * np.sin(np.linspace(0, 6.28, 100))])
        for step in range(3):
            g.step(); o.step()
            for a, b in zip(g.probe(px, py), o.probe(px, py)):
                assert np.array_equal(a, b), f"step {step}"
            Fg, Fo = g.body_forces(2), o.body_forces(2)
            assert np.array_equal(Fg, Fo), f"step {step}: {Fg} vs {Fo}"
        assert Fg[0, 0] + Fg[0, 2] > 0          # the flow pushes the resting cylinder downstream
    o.close()


def test_diagnostics_are_refused_in_reference_mode(ref_case):
    inp = ifx.make_input(50, 50, 1e-3, 150.0)
    with ifx.ImmerseFlow(inp, ref_case["xf"], ref_case["yf"]) as s:
        s.initializeData()
        with pytest.raises(ifx.IfxError, match="IFX_COMPAT_FULL"):
            s.probe([0.5], [0.5])
