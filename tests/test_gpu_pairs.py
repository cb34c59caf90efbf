"""ifx_options.ppe_pairs: two point-Jacobi sweeps of the general Poisson operator per pass over memory
(kernels_pair.cu).  The iterate between the two sweeps exists in shared memory only, yet nothing observable may change:
pressure, velocities and iteration counts equal those of single sweeps and of the CPU oracle bit for bit — when the stop
rule fires on a stored iterate, on an unstored one (re-created by one single sweep), at PPE_itermax (odd and even), and
when the fused sum lands on the tolerance (certified re-evaluation)."""
import numpy as np
import pytest

import immerseflow_b200 as ifx
import _oracle as orc

pytestmark = pytest.mark.gpu


def _case(ncx=300, ncy=171):
    xf, yf = orc.stretched_faces(ncx, 4.0, 1.015), orc.stretched_faces(ncy, 2.0, 1.015)
    bodies = [orc.circle_markers(1.3, 1.0, 0.31, 64), orc.ellipse_markers(2.6, 0.8, 0.35, 0.12, 0.4, 80)]
    return xf, yf, bodies


U0 = 1e-8     # a slow stream: the un-normalised residual sums stay below 1 — the reference's loop starts from a residual of 1.0
              # (PPESolver.cu:170-172), so only tolerances below 1 leave anything to iterate on
BC = dict(u_bc_w=U0, u_bc_e=U0, u_bc_s=U0, u_bc_n=U0)


def _run(xf, yf, bodies, itermax, tol, pairs, steps=2, ncx=300, ncy=171, reduce_mode=ifx.IFX_REDUCE_FUSED):
    inp = ifx.make_input(ncx, ncy, 2e-3, 200.0, AD_itermax=10, PPE_itermax=itermax)
    with ifx.ImmerseFlow(inp, xf, yf, compat=ifx.IFX_COMPAT_FULL, ppe_abs_residual=1, ppe_tol=tol, ppe_pairs=pairs,
                         sweeps_per_batch=7, bc=BC, reduce_mode=reduce_mode) as s:
        s.initializeData()
        s.set("u", np.full(inp.nx * inp.ny, U0)); s.set("v", np.zeros(inp.nx * inp.ny))
        s.set_bodies(bodies)
        counts, fb = [], 0
        for _ in range(steps):
            st = s.step()
            counts.append((st.ad_iters, st.ppe_sweeps))
            fb += st.exact_fallbacks
        return counts, {k: s.get(k) for k in ("u", "v", "p")}, fb, s.launch_count, st.ppe_residual


def _oracle(xf, yf, bodies, itermax, tol, steps=2):
    o = orc.FullSolver(xf, yf, 2e-3, 200.0, 10, itermax, ppe_tol=tol, ppe_abs=1, bc_u=(U0,) * 4)
    n = (len(xf) + 1) * (len(yf) + 1)
    o.set("u", np.full(n, U0)); o.set("v", np.zeros(n))
    o.set_bodies(bodies); o.update_ib()
    counts = []
    for _ in range(steps):
        st = o.step()
        counts.append((int(st[0]), int(st[3])))
    f = {k: o.get(k) for k in ("u", "v", "p")}
    o.close()
    return counts, f


@pytest.mark.parametrize("itermax", [1, 2, 7, 8, 30])
def test_pairs_equal_single_sweeps_when_the_solve_ends_at_itermax(itermax):
    xf, yf, bodies = _case()
    c1, f1, _, l1, _ = _run(xf, yf, bodies, itermax, 0.0, 0)
    c2, f2, _, l2, _ = _run(xf, yf, bodies, itermax, 0.0, 1)
    co, fo = _oracle(xf, yf, bodies, itermax, 0.0)
    assert c1 == c2 == co and c2[-1][1] == itermax
    for k in f1:
        assert np.array_equal(f1[k], f2[k]), k
    inner = np.zeros((len(yf) + 1, len(xf) + 1), bool); inner[1:-1, 1:-1] = True
    for k in fo:
        assert np.array_equal(f2[k][inner.reshape(-1)], fo[k][inner.reshape(-1)]), k
    if itermax >= 7:
        assert l2 < l1, "pairs must need fewer launches"


def test_pairs_equal_single_sweeps_whatever_iterate_the_stop_rule_fires_on():
    """tolerances chosen from the residual history so that the rule fires after 1 .. 12 sweeps: odd counts end on an
    iterate a pair never stored"""
    xf, yf, bodies = _case()
    # residual after k sweeps of the first step: run single sweeps with itermax = k and read the residual
    seen = set()
    for k in range(1, 13):
        res_k = _run(xf, yf, bodies, k, 0.0, 0, steps=1)[4]       # residual of iterate k (single sweeps, itermax = k)
        assert res_k < 1.0, "the case must stay below the loop's starting residual of 1.0"
        tol = res_k * (1.0 + 1e-9)            # just above the residual of iterate k: the rule fires there (or earlier)
        c1, f1, _, _, _ = _run(xf, yf, bodies, 40, tol, 0, steps=1)
        c2, f2, _, _, _ = _run(xf, yf, bodies, 40, tol, 1, steps=1)
        co, fo = _oracle(xf, yf, bodies, 40, tol, steps=1)
        assert c1 == c2 == co, (k, c1, c2, co)
        for name in f1:
            assert np.array_equal(f1[name], f2[name]), (k, name)
        seen.add(c2[0][1])
    assert any(q % 2 for q in seen) and any(q % 2 == 0 for q in seen), seen


def test_a_tolerance_of_one_or_more_means_no_sweep_like_the_reference_loop():
    """PPESolver.cu:170-172: `res = 1.0; while (res > tol && ...)` — with tol >= 1 the loop body never runs"""
    xf, yf, bodies = _case()
    for pairs in (0, 1):
        c, f, _, _, _ = _run(xf, yf, bodies, 40, 1.0, pairs, steps=1)
        co, fo = _oracle(xf, yf, bodies, 40, 1.0, steps=1)
        assert c == co and c[0][1] == 0
        inner = np.zeros((len(yf) + 1, len(xf) + 1), bool); inner[1:-1, 1:-1] = True
        for k in fo:
            assert np.array_equal(f[k][inner.reshape(-1)], fo[k][inner.reshape(-1)]), k


def test_pairs_on_the_rounding_boundary_take_the_certified_path():
    """the tolerance placed exactly on a residual: the fused sum is ambiguous, the reference-order re-evaluation decides
    (on a stored or an unstored iterate alike) and the counts still equal the oracle's"""
    xf, yf, bodies = _case(200, 121)
    for k in (3, 4):
        # the reference-order sum of iterate k itself
        tol = _run(xf, yf, bodies, k, 0.0, 0, steps=1, ncx=200, ncy=121, reduce_mode=ifx.IFX_REDUCE_REFERENCE)[4]
        assert tol < 1.0
        c2, f2, fb, _, _ = _run(xf, yf, bodies, 40, tol, 1, steps=1, ncx=200, ncy=121)
        co, fo = _oracle(xf, yf, bodies, 40, tol, steps=1)
        assert c2 == co, (k, c2, co)
        assert fb >= 1, "the decision on the tolerance must have been certified"
        inner = np.zeros((len(yf) + 1, len(xf) + 1), bool); inner[1:-1, 1:-1] = True
        assert np.array_equal(f2["p"][inner.reshape(-1)], fo["p"][inner.reshape(-1)])
