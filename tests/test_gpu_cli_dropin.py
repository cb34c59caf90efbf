"""Drop-in check of the C++ driver (immerseflow_b200/bin/immerseflow): staged exactly like the reference tree
(cwd = src/, ../inputs, ../results), it must leave the reference's own result files behind — BYTE-identical to the
golden results/uc.dat, vc.dat, final_results.dat the reference ships (20 steps, App. B of SURVEY.md) — and, with
--reference-log, print the same "iter = k uRes vRes" lines as the reference binary."""
import os
import re
import shutil
import subprocess

import numpy as np
import pytest

from conftest import ROOT, load_tecplot

pytestmark = pytest.mark.gpu

CLI = os.path.join(ROOT, "immerseflow_b200", "bin", "immerseflow")
REF_BIN = os.path.join(ROOT, "oracle", "_ref", "immerseFlow_ref")


def stage(tmp_path, ref_case):
    w = tmp_path / "tree"
    (w / "src").mkdir(parents=True); (w / "results").mkdir()
    shutil.copytree(os.path.join(ref_case["dir"], "inputs"), w / "inputs")
    return w


@pytest.mark.skipif(not os.path.exists(CLI), reason="CLI not built")
def test_cli_reproduces_shipped_result_files_byte_for_byte(tmp_path, ref_case):
    w = stage(tmp_path, ref_case)
    r = subprocess.run([CLI, "--steps", "20", "--write-every-step"], cwd=w / "src", capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    for name in ("uc.dat", "vc.dat", "final_results.dat"):
        got = (w / "results" / name).read_bytes()
        want = open(os.path.join(ref_case["dir"], "results", name), "rb").read().replace(b"\r\n", b"\n")
        assert got == want, name
    # reader side of the contract (reference results/plot.py:19-28)
    d = load_tecplot(w / "results" / "uc.dat")
    assert d.shape == (52 * 52, 3)


@pytest.mark.skipif(not os.path.exists(CLI), reason="CLI not built")
def test_cli_error_behaviour_matches_reference(tmp_path, ref_case):
    w = stage(tmp_path, ref_case)
    os.remove(w / "inputs" / "inputs.txt")
    r = subprocess.run([CLI], cwd=w / "src", capture_output=True, text=True, timeout=60)
    assert r.returncode == 1 and "Unable to open file: ../inputs/inputs.txt" in r.stderr     # main.cu:12-15
    w2 = stage(tmp_path / "b", ref_case)
    os.remove(w2 / "inputs" / "xgrid.dat2")
    r = subprocess.run([CLI], cwd=w2 / "src", capture_output=True, text=True, timeout=60)
    assert r.returncode == 1 and "Error opening xgrid.dat" in r.stderr                       # preSim.cu:270


@pytest.mark.skipif(not (os.path.exists(CLI) and os.path.exists(REF_BIN)), reason="needs CLI and reference binary")
def test_cli_reference_log_equals_reference_stdout(tmp_path, ref_case):
    w = stage(tmp_path, ref_case)
    txt = (w / "inputs" / "inputs.txt").read_text()
    (w / "inputs" / "inputs.txt").write_text(re.sub(r"^1E-6(\s+)100(\s)", r"1E-6\g<1>8\2", txt, flags=re.M))
    ours = subprocess.run([CLI, "--reference-log", "--exact-reduction"], cwd=w / "src", capture_output=True, text=True, timeout=300)
    ref = subprocess.run([REF_BIN], cwd=w / "src", capture_output=True, text=True, timeout=300,
                         env=dict(os.environ, IFX_REF_SAVE="none"))
    assert ours.returncode == 0 and ref.returncode == 0
    pick = lambda s: [l for l in s.splitlines() if l.startswith("iter = ")]
    assert pick(ours.stdout) == pick(ref.stdout) and len(pick(ref.stdout)) == 8 * 5


@pytest.mark.skipif(not os.path.exists(CLI), reason="CLI not built")
def test_cli_full_mode_with_bodies_on_the_stretched_grid(tmp_path, ref_case):
    """The case config 1 of BASELINE.json means: the shipped inputs.txt sized for the stretched xgrid.dat/ygrid.dat
    (180 x 128 cells, 10 x 5 domain) with the reference's cylinder (centre (3, 2.5), r = 0.5, preSim.cu:123)."""
    w = stage(tmp_path, ref_case)
    txt = (w / "inputs" / "inputs.txt").read_text().replace("50      50", "180     128")
    (w / "inputs" / "inputs.txt").write_text(txt.replace("100000", "300   "))
    t = 2 * np.pi * np.arange(96) / 96
    with open(w / "inputs" / "bodies.txt", "w") as f:
        f.write("1\n96 0 0\n" + "".join(f"{3 + 0.5 * np.cos(a):.17g} {2.5 + 0.5 * np.sin(a):.17g}\n" for a in t))
    r = subprocess.run([CLI, "--mode", "full", "--stretched", "--bodies", "../inputs/bodies.txt", "--steps", "3"],
                       cwd=w / "src", capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    ib = load_tecplot(w / "results" / "final_results.dat")
    assert ib.shape == (182 * 130, 3)
    inside = np.hypot(ib[:, 0] - 3.0, ib[:, 1] - 2.5) < 0.45
    assert np.all(ib[inside, 2] == 0.0) and ib[~inside, 2].mean() > 0.98     # iBlank: 0 in the cylinder, 1 in the fluid
    for name in ("uc.dat", "vc.dat", "p.dat"):
        d = load_tecplot(w / "results" / name)
        assert d.shape == (182 * 130, 3) and np.isfinite(d).all()
