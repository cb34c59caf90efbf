"""Parity of the CUDA path (through the C-ABI) with the oracle, the golden files and — when it was
built — the reference's own CUDA binary, for the stages the reference implements (IFX_COMPAT_REFERENCE):
predictor, reduction order, Laplace-Jacobi PPE.

Bar: BIT-EXACT (np.array_equal on fp64) for fields and, in IFX_REDUCE_REFERENCE mode, for residuals;
identical iteration counts always.  The only cells excluded anywhere are the 4 corner ghosts when
comparing against the reference binary, where the reference itself has a data race (App. A Q6).
"""
import os
import re
import shutil
import subprocess

import numpy as np
import pytest

import immerseflow_b200 as ifx
import _oracle as orc
from conftest import ROOT, fmt6, load_tecplot

pytestmark = pytest.mark.gpu


def make_solver(xf, yf, dt, Re, ad_itermax=25, ppe_itermax=100000, **kw):
    inp = ifx.make_input(len(xf) - 1, len(yf) - 1, dt, Re, AD_itermax=ad_itermax, PPE_itermax=ppe_itermax)
    return ifx.ImmerseFlow(inp, xf, yf, **kw)


def corners_mask(nx, ny):
    m = np.ones((ny, nx), dtype=bool)
    m[0, 0] = m[0, -1] = m[-1, 0] = m[-1, -1] = False
    return m.reshape(-1)


# ------------------------------------------------------------------------------------------------
def test_shipped_case_20_steps_bit_exact_and_golden(ref_case):
    """inputs.txt on xgrid.dat2/ygrid.dat2: 20 predictor steps == oracle bit for bit, == uc.dat/vc.dat at 6 decimals."""
    inp = ifx.read_input_file(os.path.join(ref_case["dir"], "inputs", "inputs.txt"))
    with ifx.ImmerseFlow(inp, ref_case["xf"], ref_case["yf"]) as s:
        s.initializeData()
        g = orc.Grid(ref_case["xf"], ref_case["yf"])
        assert np.array_equal(s.get("xc"), g.xc) and np.array_equal(s.get("yc"), g.yc)
        u0, v0 = s.get("u"), s.get("v")
        # device IC (libdevice) vs oracle IC (glibc): within 2 ulp; the oracle run starts from the device IC
        uo, vo, _ = orc.initial_condition(g)
        assert np.max(np.abs(u0 - uo)) <= 4.5e-16 and np.max(np.abs(v0 - vo)) <= 1e-16
        pr = orc.Predictor(g, u0, v0, inp.dt, inp.Re, inp.AD_itermax)
        for step in range(20):
            st = s.step()
            k, hist = pr.step()
            assert st.ad_iters == k == 5
            assert np.array_equal(s.get("u"), pr.u), f"u differs at step {step}"
            assert np.array_equal(s.get("v"), pr.v), f"v differs at step {step}"
            assert st.ad_ures == pytest.approx(hist[-2], rel=1e-12) and st.ad_vres == pytest.approx(hist[-1], rel=1e-12)
            assert st.exact_fallbacks == 0
        gu = load_tecplot(os.path.join(ref_case["dir"], "results", "uc.dat"))
        gv = load_tecplot(os.path.join(ref_case["dir"], "results", "vc.dat"))
        assert np.array_equal(fmt6(s.get("u")), gu[:, 2])
        assert np.array_equal(fmt6(s.get("v")), gv[:, 2])
        assert np.array_equal(fmt6(s.get("iblank")), np.ones(52 * 52))     # final_results.dat


@pytest.mark.parametrize("ncx,ncy,stretch", [(50, 50, False), (37, 64, True), (65, 65, True), (255, 256, False),
                                             (256, 257, True), (300, 301, True), (513, 700, True), (1024, 1024, False)])
def test_predictor_bit_exact_various_grids(ncx, ncy, stretch):
    """Odd/even widths, widths around the 64/256-column warp/CTA tiles, stretched metrics, nx <= ny."""
    xf = orc.stretched_faces(ncx, 2.0, 1.03) if stretch else ifx.uniform_faces(ncx, 1.0)
    yf = orc.stretched_faces(ncy, 1.5, 1.02) if stretch else ifx.uniform_faces(ncy, 1.0)
    g = orc.Grid(xf, yf)
    rng = np.random.default_rng(ncx * 1000 + ncy)
    u0 = 1.0 + 0.1 * rng.standard_normal(g.nx * g.ny)
    v0 = 0.1 * rng.standard_normal(g.nx * g.ny)
    dt, Re, itmax = 2e-4, 100.0, 12
    with make_solver(xf, yf, dt, Re, ad_itermax=itmax) as s:
        s.initializeData()
        s.set("u", u0); s.set("v", v0)
        assert np.array_equal(s.get("u"), u0)
        pr = orc.Predictor(g, u0, v0, dt, Re, itmax)
        for step in range(3):
            st = s.step()
            k, hist = pr.step()
            assert st.ad_iters == k
            assert np.array_equal(s.get("u"), pr.u) and np.array_equal(s.get("v"), pr.v), f"step {step}"


def _oracle_source(g, pr, dt):
    """sx, sy of the NEXT step as the oracle would compute them (for a direct check of k_ad_source)."""
    N = g.nx * g.ny
    u, v = pr.u.copy(), pr.v.copy()
    uf, vf = np.zeros_like(pr.uf), np.zeros_like(pr.vf)
    sx, sy = np.zeros(N), np.zeros(N)
    L = orc.lib()
    L.orc_Compute_velf(g.nx, g.ny, orc.P(g.dx), orc.P(g.dy), orc.P(u), orc.P(v), orc.P(uf), orc.P(vf), 0)
    L.orc_set_velocity_BC(g.nx, g.ny, orc.P(u), orc.P(v))
    import ctypes as C
    L.orc_ADSource(g.nx, g.ny, orc.P(g.dx), orc.P(g.dy), C.c_double(dt), orc.P(u), orc.P(v), orc.P(uf), orc.P(vf),
                   orc.P(sx), orc.P(sy))
    return sx, sy


def test_source_term_bit_exact():
    """k_ad_source (velf-before-BC + BC + ADSource fused) against the oracle's three separate passes."""
    xf, yf = orc.stretched_faces(90, 3.0, 1.05), orc.stretched_faces(131, 2.0, 1.04)
    g = orc.Grid(xf, yf)
    rng = np.random.default_rng(11)
    u0 = 1.0 + 0.2 * rng.standard_normal(g.nx * g.ny)
    v0 = 0.2 * rng.standard_normal(g.nx * g.ny)
    dt, Re = 1e-3, 50.0
    with make_solver(xf, yf, dt, Re, ad_itermax=1) as s:
        s.initializeData()
        s.set("u", u0); s.set("v", v0)
        pr = orc.Predictor(g, u0, v0, dt, Re, 1)
        sx, sy = _oracle_source(g, pr, dt)
        s.step()
        assert np.array_equal(s.get("sx"), sx) and np.array_equal(s.get("sy"), sy)


def test_reference_order_residuals_bit_exact(ref_case):
    """IFX_REDUCE_REFERENCE: every residual the reference would print is reproduced to the last bit."""
    g = orc.Grid(ref_case["xf"], ref_case["yf"])
    with make_solver(ref_case["xf"], ref_case["yf"], 1e-3, 150.0, reduce_mode=ifx.IFX_REDUCE_REFERENCE) as s:
        s.initializeData()
        pr = orc.Predictor(g, s.get("u"), s.get("v"), 1e-3, 150.0, 25)
        for _ in range(5):
            st = s.step()
            k, hist = pr.step()
            assert st.ad_iters == k
            assert st.ad_ures == hist[-2] and st.ad_vres == hist[-1]
            assert np.array_equal(s.get("u"), pr.u)


@pytest.mark.parametrize("n", [1, 255, 256, 257, 511, 512, 513, 2704, 131072, 1052676, 3_000_001])
def test_reduction_bit_exact(n, ref_case):
    rng = np.random.default_rng(n)
    x = rng.standard_normal(n) * 10.0 ** rng.integers(-10, 10, n)
    with make_solver(ref_case["xf"], ref_case["yf"], 1e-3, 150.0) as s:
        assert s.Reduction(x) == orc.reduction(x)


def test_stop_rule_on_the_rounding_boundary(ref_case):
    """Tolerance placed exactly on (and one ulp below) a reference-order residual: the fused sum cannot decide,
    the certified fallback re-evaluates in reference order and lands on the oracle's iteration count."""
    g = orc.Grid(ref_case["xf"], ref_case["yf"])
    xf, yf = ref_case["xf"], ref_case["yf"]
    with make_solver(xf, yf, 1e-3, 150.0) as probe:
        probe.initializeData()
        u0, v0 = probe.get("u"), probe.get("v")
    _, hist = orc.Predictor(g, u0, v0, 1e-3, 150.0, 25).step()
    s3 = hist[4] + hist[5]                                   # uRes + vRes after iteration 3
    for tol, expect_k in ((s3, 3), (np.nextafter(s3, 0.0), 4)):
        pr = orc.Predictor(g, u0, v0, 1e-3, 150.0, 25, tol=tol)
        k, _ = pr.step()
        assert k == expect_k
        with make_solver(xf, yf, 1e-3, 150.0, ad_tol=tol) as s:
            s.initializeData()
            st = s.step()
            assert st.ad_iters == expect_k and st.exact_fallbacks == 1
            assert np.array_equal(s.get("u"), pr.u) and np.array_equal(s.get("v"), pr.v)


@pytest.mark.parametrize("itermax", [0, 1, 2, 3])
def test_small_itermax_edge_cases(ref_case, itermax):
    """AD_itermax 0 (no iteration), 1 (final buffer's ghost ring never written by the solver), 2, 3."""
    g = orc.Grid(ref_case["xf"], ref_case["yf"])
    with make_solver(ref_case["xf"], ref_case["yf"], 1e-3, 150.0, ad_itermax=itermax) as s:
        s.initializeData()
        pr = orc.Predictor(g, s.get("u"), s.get("v"), 1e-3, 150.0, itermax)
        for _ in range(3):
            st = s.step()
            k, _ = pr.step()
            assert st.ad_iters == k == itermax
            if itermax > 0:
                assert np.array_equal(s.get("u"), pr.u) and np.array_equal(s.get("v"), pr.v)
            else:   # the oracle applies set_velocity_BC once even with zero iterations (ADSolver.cu:304)
                m = np.zeros((g.ny, g.nx), bool); m[1:-1, 1:-1] = True
                assert np.array_equal(s.get("u")[m.reshape(-1)], pr.u[m.reshape(-1)])


# ------------------------------------------------------------------------------------------------
def test_laplace_ppe_shipped_case_bit_exact_and_golden(ref_case):
    g = orc.Grid(ref_case["xf"], ref_case["yf"])
    k_o, p_o, res_o = orc.ppe_solve(g, np.zeros(g.nx * g.ny), 100000)
    for mode in (ifx.IFX_REDUCE_FUSED, ifx.IFX_REDUCE_REFERENCE):
        with make_solver(ref_case["xf"], ref_case["yf"], 1e-3, 150.0, reduce_mode=mode) as s:
            s.initializeData()
            st = s.PPESolver()
            assert st.ppe_sweeps == k_o
            p = s.get("p")
            assert np.array_equal(p, p_o)
            if mode == ifx.IFX_REDUCE_REFERENCE:
                assert st.ppe_residual == res_o
            else:
                assert st.ppe_residual == pytest.approx(res_o, rel=1e-9)
            gp = load_tecplot(os.path.join(ref_case["dir"], "results", "p.dat"))
            assert np.array_equal(fmt6(p), gp[:, 2])


@pytest.mark.parametrize("ncx,ncy,itermax", [(40, 61, 500), (129, 130, 300), (300, 257, 64), (64, 64, 0), (64, 64, 1)])
def test_laplace_ppe_various(ncx, ncy, itermax):
    """Stretched grids, sweep cap reached (iter == PPE_itermax), nx > ny, zero/one sweep, nonzero start field."""
    xf, yf = orc.stretched_faces(ncx, 2.0, 1.03), orc.stretched_faces(ncy, 1.0, 1.02)
    g = orc.Grid(xf, yf)
    rng = np.random.default_rng(5)
    p0 = rng.standard_normal(g.nx * g.ny)
    k_o, p_o, res_o = orc.ppe_solve(g, p0, itermax)
    with make_solver(xf, yf, 1e-3, 100.0, ppe_itermax=itermax, sweeps_per_batch=37) as s:
        s.initializeData()
        s.set("p", p0)
        st = s.PPESolver()
        assert st.ppe_sweeps == k_o == itermax
        assert np.array_equal(s.get("p"), p_o)


def test_set_get_roundtrip_and_save(tmp_path):
    xf, yf = orc.stretched_faces(33, 1.0), orc.stretched_faces(47, 1.0)
    g = orc.Grid(xf, yf)
    rng = np.random.default_rng(1)
    with make_solver(xf, yf, 1e-3, 100.0) as s:
        s.initializeData()
        for name in ("u", "v", "p"):
            a = rng.standard_normal(g.nx * g.ny)
            s.set(name, a)
            assert np.array_equal(s.get(name), a)
            out = tmp_path / f"{name}.dat"
            s.saveDataToFile(name, str(out))
            ref = tmp_path / f"{name}_ref.dat"
            orc.lib().orc_write_results_to_file(orc.P(g.xc), orc.P(g.yc), orc.P(a), g.nx, g.ny, str(ref).encode())
            assert out.read_bytes() == ref.read_bytes()
        with pytest.raises(ifx.IfxError):
            s.set("u", np.zeros(5))


# ------------------------------------------------------------------------------------------------
# Against the reference's own CUDA solver (oracle/_ref, built here from /root/reference's sources;
# the binary travels to the GPU box, the sources do not).
# ------------------------------------------------------------------------------------------------
REF_BIN = os.path.join(ROOT, "oracle", "_ref", "immerseFlow_ref")
REF_PPE_BIN = os.path.join(ROOT, "oracle", "_ref", "immerseFlow_ref_ppe")


def stage_reference_run(tmp_path, ref_case, binary, tmax=None, env_save="all"):
    w = tmp_path / "w"
    (w / "src").mkdir(parents=True); (w / "inputs").mkdir(); (w / "results").mkdir()
    txt = open(os.path.join(ref_case["dir"], "inputs", "inputs.txt")).read()
    if tmax is not None:
        new = re.sub(r"^1E-6(\s+)100(\s)", lambda m: f"1E-6{m.group(1)}{tmax}{m.group(2)}", txt, flags=re.M)
        assert new != txt or tmax == 100
        txt = new
    (w / "inputs" / "inputs.txt").write_text(txt)
    for f in ("xgrid.dat2", "ygrid.dat2"):
        shutil.copy(os.path.join(ref_case["dir"], "inputs", f), w / "inputs" / f)
    env = dict(os.environ, IFX_REF_SAVE=env_save)
    r = subprocess.run([binary], cwd=w / "src", env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    return w, r.stdout


@pytest.mark.skipif(not os.path.exists(REF_BIN), reason="reference binary not built (oracle/_ref)")
def test_against_reference_cuda_binary_predictor(tmp_path, ref_case):
    """The reference's unmodified ADSolver/preSim/main TUs running on this GPU, 25 steps: our u, v are
    BIT-IDENTICAL after every step (corner ghosts excluded: racy in the reference) and the iteration counts
    printed by the reference (ADSolver.cu:369) match."""
    nsteps = 25
    w, out = stage_reference_run(tmp_path, ref_case, REF_BIN, tmax=nsteps)
    iters = [int(m.group(1)) for m in re.finditer(r"^iter = (\d+) ", out, re.M)]
    per_step, cur = [], 0
    for it in iters:
        if it == 1 and cur:
            per_step.append(cur)
        cur = it
    per_step.append(cur)
    assert len(per_step) == nsteps
    inp = ifx.read_input_file(str(w / "inputs" / "inputs.txt"))
    mask = corners_mask(inp.nx, inp.ny)
    with ifx.ImmerseFlow(inp, ref_case["xf"], ref_case["yf"]) as s:
        s.initializeData()
        worst = 0.0
        for step in range(nsteps):
            st = s.step()
            ru = np.fromfile(w / "results" / f"uc.dat.{step}.f64")
            rv = np.fromfile(w / "results" / f"vc.dat.{step}.f64")
            assert st.ad_iters == per_step[step]
            u, v = s.get("u"), s.get("v")
            worst = max(worst, np.linalg.norm(u[mask] - ru[mask]) / np.linalg.norm(ru[mask]))
            assert np.array_equal(u[mask], ru[mask]), f"step {step}: rel-L2 {worst:.3e}"
            assert np.array_equal(v[mask], rv[mask]), f"step {step}"
        ib = np.fromfile(w / "results" / "final_results.dat.0.f64")
        assert np.array_equal(s.get("iblank"), ib)


@pytest.mark.skipif(not os.path.exists(REF_PPE_BIN), reason="reference PPE binary not built (oracle/_ref)")
def test_against_reference_cuda_binary_ppe(tmp_path, ref_case):
    """The reference's Poisson kernels (PPESolver.cu with the documented compile repair) on this GPU."""
    w, _ = stage_reference_run(tmp_path, ref_case, REF_PPE_BIN)
    files = sorted((w / "results").glob("p.dat.*.f64"))
    assert files
    rp = np.fromfile(files[-1])
    inp = ifx.read_input_file(str(w / "inputs" / "inputs.txt"))
    with ifx.ImmerseFlow(inp, ref_case["xf"], ref_case["yf"]) as s:
        s.initializeData()
        s.PPESolver()
        assert np.array_equal(s.get("p"), rp)
