"""The C++ driver's host-side behaviour that needs no GPU: usage text, option validation, and the reference's error
behaviour for missing input files (message on stderr + exit status 1, src/main.cu:12-15, preSim.cu:270,283) — all of
which happen before the first CUDA call."""
import os
import shutil
import subprocess

import pytest

from conftest import ROOT

CLI = os.path.join(ROOT, "immerseflow_b200", "bin", "immerseflow")
pytestmark = pytest.mark.skipif(not os.path.exists(CLI), reason="CLI not built")


def run(args, cwd):
    return subprocess.run([CLI] + args, cwd=cwd, capture_output=True, text=True, timeout=60)


def stage(tmp_path, ref_case):
    w = tmp_path / "tree"
    (w / "src").mkdir(parents=True); (w / "results").mkdir()
    shutil.copytree(os.path.join(ref_case["dir"], "inputs"), w / "inputs")
    return w


def test_help_lists_every_option(tmp_path):
    r = run(["--help"], tmp_path)
    assert r.returncode == 0
    for opt in ("--input", "--xgrid", "--ygrid", "--stretched", "--results", "--mode", "--bodies", "--steps", "--write-every-step",
                "--reference-log", "--exact-reduction", "--checkpoints", "--restart", "--device", "--ppe-solver", "--ppe-omega",
                "--bc-u", "--bc-v", "--ic", "--forces", "--probes", "--probe-out", "--ad-tol", "--ppe-tol"):
        assert opt in r.stdout, opt


@pytest.mark.parametrize("args,msg", [(["--frobnicate"], "unknown option"), (["--mode", "fast"], "--mode must be"),
                                      (["--ic", "random"], "--ic must be"), (["--bc-u", "1,2,3"], "--bc-u needs W,E,S,N"),
                                      (["--forces", "f.dat"], "need --mode full"), (["--steps"], "missing value"),
                                      (["--mode", "full", "--probes", "p.txt"], "go together")])
def test_bad_options_exit_with_a_message(tmp_path, args, msg):
    r = run(args, tmp_path)
    assert r.returncode == 1 and msg in r.stderr, r.stderr


def test_missing_files_fail_like_the_reference(tmp_path, ref_case):
    w = stage(tmp_path, ref_case)
    os.remove(w / "inputs" / "ygrid.dat2")
    r = run([], w / "src")
    assert r.returncode == 1 and "Error opening ygrid.dat" in r.stderr                         # preSim.cu:283
    os.remove(w / "inputs" / "inputs.txt")
    r = run([], w / "src")
    assert r.returncode == 1 and "Unable to open file: ../inputs/inputs.txt" in r.stderr       # main.cu:12-15


def test_without_a_gpu_the_driver_says_so(tmp_path, ref_case):
    import immerseflow_b200 as ifx
    if ifx.load_library().ifx_device_count() > 0:
        pytest.skip("a GPU is present")
    w = stage(tmp_path, ref_case)
    r = run([], w / "src")
    assert r.returncode == 1 and "no CUDA device" in r.stderr and "no CPU fallback" in r.stderr
