"""Restart files: a run continued from a checkpoint is bit-identical to an uninterrupted run, in both
compat modes, with bodies, through the Python binding and through the CLI's `Restart` / `Write Interval` handling."""
import os
import re
import shutil
import subprocess

import numpy as np
import pytest

import immerseflow_b200 as ifx
import _oracle as orc
from conftest import ROOT

pytestmark = pytest.mark.gpu


def _run(mode, tmp_path, split):
    full = mode == "full"
    ncx, ncy = (140, 90) if full else (90, 140)        # reference mode needs nx <= ny (App. A Q2: vf == 0 there)
    xf, yf = orc.stretched_faces(ncx, 4.0, 1.02), orc.stretched_faces(ncy, 2.0, 1.02)
    inp = ifx.make_input(ncx, ncy, 1e-3, 150.0, AD_itermax=12, PPE_itermax=40)
    kw = dict(compat=ifx.IFX_COMPAT_FULL, ppe_abs_residual=1) if full else {}
    bodies = [orc.circle_markers(1.5, 1.0, 0.3, 48)]

    def fresh():
        s = ifx.ImmerseFlow(inp, xf, yf, **kw)
        s.initializeData()
        if full:
            s.set_bodies(bodies)
        return s

    nsteps = 5
    with fresh() as a:
        for _ in range(nsteps):
            a.step() if full else (a.ADsolver(), a.PPESolver())
        want = {k: a.get(k) for k in ("u", "v", "p") + (("uf", "vf") if full else ())}
    ck = str(tmp_path / f"ck_{mode}.ifx")
    with fresh() as b:
        for _ in range(split):
            b.step() if full else (b.ADsolver(), b.PPESolver())
        b.save_checkpoint(ck, step=split, time=split * 1e-3)
    with fresh() as c:
        step, t = c.load_checkpoint(ck)
        assert (step, t) == (split, split * 1e-3)
        for _ in range(nsteps - split):
            c.step() if full else (c.ADsolver(), c.PPESolver())
        for k, w in want.items():
            assert np.array_equal(c.get(k), w), f"{mode}: {k} differs after restart at step {split}"


@pytest.mark.parametrize("mode", ["reference", "full"])
@pytest.mark.parametrize("split", [0, 1, 3])      # 0: saved before the first step — the face velocities are not state yet
def test_restart_is_bit_identical(mode, split, tmp_path):
    _run(mode, tmp_path, split)


def test_checkpoint_taken_after_the_bodies_moved_ignores_the_stale_faces(tmp_path):
    """After ifx_iblank_update with moved bodies the stored uf, vf are stale (closed faces moved): a file written then
    must not restore them as valid — the continued run rebuilds them from the cells, like the unbroken run."""
    ncx, ncy = 140, 90
    xf, yf = orc.stretched_faces(ncx, 4.0, 1.02), orc.stretched_faces(ncy, 2.0, 1.02)
    inp = ifx.make_input(ncx, ncy, 1e-3, 150.0, AD_itermax=12, PPE_itermax=40)
    kw = dict(compat=ifx.IFX_COMPAT_FULL, ppe_abs_residual=1)
    b0, b1 = [orc.circle_markers(1.5, 1.0, 0.3, 48)], [orc.circle_markers(1.56, 0.97, 0.3, 48)]
    ck = str(tmp_path / "moved.ifx")
    with ifx.ImmerseFlow(inp, xf, yf, **kw) as a:
        a.initializeData(); a.set_bodies(b0)
        a.step(); a.step()
        a.set_bodies(b1); a.iblank_update()
        a.save_checkpoint(ck, 2, 2e-3)
        a.step()
        want = {k: a.get(k) for k in ("u", "v", "p", "uf", "vf")}
    with ifx.ImmerseFlow(inp, xf, yf, **kw) as c:
        c.initializeData(); c.set_bodies(b1)
        c.load_checkpoint(ck)
        c.step()
        for k, w in want.items():
            assert np.array_equal(c.get(k), w), k


def test_checkpoint_rejects_a_different_grid(tmp_path):
    xf, yf = ifx.uniform_faces(40, 1.0), ifx.uniform_faces(30, 1.0)
    ck = str(tmp_path / "a.ifx")
    with ifx.ImmerseFlow(ifx.make_input(40, 30, 1e-3, 100.0), xf, yf) as s:
        s.initializeData()
        s.save_checkpoint(ck, 7, 0.007)
    with ifx.ImmerseFlow(ifx.make_input(30, 40, 1e-3, 100.0), yf, xf) as s:
        s.initializeData()
        with pytest.raises(ifx.IfxError, match="different grid"):
            s.load_checkpoint(ck)
    with open(ck, "r+b") as f:
        f.write(b"garbage!")
    with ifx.ImmerseFlow(ifx.make_input(40, 30, 1e-3, 100.0), xf, yf) as s:
        s.initializeData()
        with pytest.raises(ifx.IfxError, match="not a checkpoint"):
            s.load_checkpoint(ck)


def test_cli_restart_and_write_interval(ref_case, tmp_path):
    """inputs.txt `Restart 1 T` + `Write Interval`: 20 steps in one go == 12 steps, restart file, 8 more steps."""
    exe = os.path.join(ROOT, "immerseflow_b200", "bin", "immerseflow")
    if not os.path.exists(exe):
        pytest.skip("CLI not built")

    def stage(name, restart, tmax, interval):
        w = tmp_path / name
        (w / "src").mkdir(parents=True); (w / "results").mkdir()
        shutil.copytree(os.path.join(ref_case["dir"], "inputs"), w / "inputs")
        txt = (w / "inputs" / "inputs.txt").read_text()
        txt = re.sub(r"^0([ \t]+)9[ \t]*$", f"{restart[0]}\\g<1>{restart[1]}", txt, count=1, flags=re.M)
        txt = re.sub(r"^1E-6(\s+)100(\s)", f"1E-6\\g<1>{tmax}\\2", txt, flags=re.M)
        txt = re.sub(r"^1000[ \t]*$", str(interval), txt, flags=re.M)
        (w / "inputs" / "inputs.txt").write_text(txt)
        return w

    one = stage("one", (0, 9), 20, 1000)
    subprocess.run([exe], cwd=one / "src", check=True, capture_output=True, timeout=300)
    two = stage("two", (0, 9), 12, 4)
    subprocess.run([exe, "--checkpoints"], cwd=two / "src", check=True, capture_output=True, text=True, timeout=300)
    assert sorted(p.name for p in (two / "results").glob("restart.*.ifx")) == [f"restart.{k:07d}.ifx" for k in (4, 8, 12)]
    keep = two / "results"
    again = stage("three", (1, 12), 20, 1000)
    shutil.copy(keep / "restart.0000012.ifx", again / "results" / "restart.0000012.ifx")
    r = subprocess.run([exe], cwd=again / "src", check=True, capture_output=True, text=True, timeout=300)
    assert "restarted from" in r.stdout and "step 13:" in r.stdout and "step 12:" not in r.stdout, r.stdout
    for f in ("uc.dat", "vc.dat"):
        assert (one / "results" / f).read_bytes() == (again / "results" / f).read_bytes(), f
