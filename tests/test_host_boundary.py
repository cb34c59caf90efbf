"""Drop-in boundary, host side (no GPU): the C-ABI library loads and exports every symbol
include/immerseflow_c.h declares; the input parser, grid reader and Tecplot writer honour the
reference's file contract (src/main.cu:10-59, src/include/preSim.cu:268-291, src/include/postSim.cu:41-66)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import immerseflow_b200 as ifx
import _oracle as orc
from conftest import ROOT


def test_library_exports_every_declared_symbol():
    lib = ifx.load_library()
    header = open(os.path.join(ROOT, "include", "immerseflow_c.h")).read()
    declared = set(re.findall(r"\b(ifx_[a-z0-9_]+)\s*\(", header))
    assert declared == set(ifx.C_ABI_SYMBOLS), declared ^ set(ifx.C_ABI_SYMBOLS)
    for name in declared:
        assert hasattr(lib, name), f"{name} not exported"
    assert lib.ifx_abi_version() == ifx.IFX_ABI_VERSION


def test_default_options_are_the_reference_hard_codes():
    lib = ifx.load_library()
    o = ifx.IfxOptions()
    lib.ifx_default_options(C.byref(o))
    assert (o.bc.u_bc_w, o.bc.u_bc_e, o.bc.u_bc_n, o.bc.u_bc_s) == (1.0, 1.0, 1.0, 1.0)   # ADSolver.cu:200-216
    assert (o.bc.v_bc_w, o.bc.v_bc_e, o.bc.v_bc_n, o.bc.v_bc_s) == (0.0, 0.0, 0.0, 0.0)
    assert o.ad_tol == 10.0 ** -6.0 and o.ppe_tol == 10.0 ** -6.0                        # :315, PPESolver.cu:172
    assert o.compat == ifx.IFX_COMPAT_REFERENCE


def test_parse_shipped_inputs_txt(ref_case):
    inp = ifx.read_input_file(os.path.join(ref_case["dir"], "inputs", "inputs.txt"))
    assert (inp.Restart, inp.Restart_Time) == (0, 9)
    assert (inp.nx, inp.ny, inp.nxf, inp.nyf) == (52, 52, 51, 51)        # main.cu:55-58
    assert (inp.Lx, inp.Ly) == (10.0, 5.0)
    assert (inp.w_AD, inp.w_PPE, inp.AD_itermax, inp.PPE_itermax, inp.AD_solver, inp.PPE_solver) == (1, 1, 25, 100000, 1, 1)
    assert (inp.ErrorMax, inp.tmax, inp.dt, inp.Re, inp.mu) == (1e-6, 100.0, 0.001, 150.0, 0.01)
    assert inp.Write_Interval == 1000


def test_parser_grammar_edge_cases(tmp_path):
    # separators starting with '=' or '_' and empty lines are skipped; keyword order is free; unknown
    # keyword lines are ignored (main.cu:17-52)
    f = tmp_path / "in.txt"
    f.write_text("=====\n\n___| x |___\nErrorMax tmax dt Re mu\n1E-3 7 0.5 42 0.1\nsomething else\n9 9 9\n"
                 "nx ny\n8 6\nWrite Interval(t/dt)\n3\n")
    inp = ifx.read_input_file(str(f))
    assert (inp.nx, inp.ny, inp.nxf, inp.nyf) == (10, 8, 9, 7)
    assert (inp.ErrorMax, inp.tmax, inp.dt, inp.Re, inp.mu) == (1e-3, 7.0, 0.5, 42.0, 0.1)
    assert inp.Write_Interval == 3 and inp.AD_itermax == 0
    with pytest.raises(ifx.IfxError):
        ifx.read_input_file(str(tmp_path / "missing.txt"))


def test_grid_reader_both_number_formats(ref_case, tmp_path):
    d = os.path.join(ref_case["dir"], "inputs")
    x2 = ifx.read_grid_file(os.path.join(d, "xgrid.dat2"), 51)          # "%.7E" written by uniformGrid.py
    assert np.array_equal(x2, ref_case["xf"]) and x2[0] == 0.0 and x2[-1] == 1.0
    xs = ifx.read_grid_file(os.path.join(d, "xgrid.dat"), 181)          # Fortran list-directed
    assert xs[0] == 0.0 and xs[1] == 0.5485451 and abs(xs[-1] - 10.0) < 1e-6
    ys = ifx.read_grid_file(os.path.join(d, "ygrid.dat"), 129)
    assert np.all(np.diff(ys) > 0) and abs(ys[-1] - 5.0) < 1e-6
    with pytest.raises(ifx.IfxError):                                    # fewer entries than requested
        ifx.read_grid_file(os.path.join(d, "xgrid.dat2"), 52)
    with pytest.raises(ifx.IfxError):
        ifx.read_grid_file(str(tmp_path / "nope.dat"), 3)


def test_tecplot_writer_bytes_match_reference_file(ref_case, tmp_path):
    g = orc.Grid(ref_case["xf"], ref_case["yf"])
    out = tmp_path / "final_results.dat"
    ifx.write_results_to_file(g.xc, g.yc, np.ones(g.nx * g.ny), g.nx, g.ny, str(out))
    ref = open(os.path.join(ref_case["dir"], "results", "final_results.dat"), "rb").read().replace(b"\r\n", b"\n")
    assert out.read_bytes() == ref
    # and byte-identical to the oracle's writer on awkward values (negative zero, rounding at 6 decimals)
    rng = np.random.default_rng(3)
    data = rng.standard_normal(g.nx * g.ny) * 10.0 ** rng.integers(-7, 4, g.nx * g.ny)
    data[:4] = [-0.0, 0.0000005, -0.0000005, 123456.7890125]
    a, b = tmp_path / "a.dat", tmp_path / "b.dat"
    ifx.write_results_to_file(g.xc, g.yc, data, g.nx, g.ny, str(a))
    assert orc.lib().orc_write_results_to_file(orc.P(g.xc), orc.P(g.yc), orc.P(data), g.nx, g.ny, str(b).encode()) == 0
    assert a.read_bytes() == b.read_bytes()
    with pytest.raises(ifx.IfxError):
        ifx.write_results_to_file(g.xc, g.yc, data, g.nx, g.ny, str(tmp_path / "no_dir" / "x.dat"))


def test_tecplot_writer_matches_printf_on_adversarial_values(tmp_path, monkeypatch):
    """The writer formats "%f" itself (one FMA gives the exact product v*1e6) on every host thread.  It must print
    what printf prints: exact ties at the sixth decimal (k/2^m values such as 1/128), negative zero, values that
    round to zero, large magnitudes, inf/nan, and the row/thread chunking must not reorder anything."""
    rng = np.random.default_rng(7)
    ties = np.array([(2 * n + 1) * 15625 / 2.0 ** m for n in range(40) for m in (7, 8, 10, 13)])      # exact .5 at 1e-6
    near = np.concatenate([np.nextafter(ties, np.inf), np.nextafter(ties, -np.inf)])
    special = np.array([0.0, -0.0, 1e-7, -1e-7, 4.9999995e-7, 5e-7, -5e-7, 0.9999995, 0.99999949999, 123456789.1234565,
                        3.9999e9, 4.0e9, 1e15, -2.5e18, 1e300, np.inf, -np.inf, np.nan, 2.0 ** -1074, 100.0, 1.0000005])
    vals = np.concatenate([ties, -ties, near, special, rng.standard_normal(3000) * 10.0 ** rng.integers(-8, 9, 3000)])
    ni = 37
    nj = (vals.size + ni - 1) // ni
    data = np.resize(vals, ni * nj)
    x = rng.standard_normal(ni) * 3.0
    y = np.concatenate([[-0.0, 1 / 128], rng.standard_normal(nj - 2)])
    want = ['TITLE = "Post Processing Tecplot"', 'VARIABLES = "X","Y","T"',
            f'ZONE T="BIG ZONE", I={ni}, J={nj}, DATAPACKING=POINT']
    want += ["%f,%f,%f" % (x[i], y[j], data[i + j * ni]) for j in range(nj) for i in range(ni)]
    for threads in ("1", "3", "16"):
        monkeypatch.setenv("IFX_IO_THREADS", threads)
        out = tmp_path / f"adv_{threads}.dat"
        ifx.write_results_to_file(x, y, data, ni, nj, str(out))
        got = out.read_text().split("\n")
        assert got[-1] == "" and got[:-1] == want, threads


def test_tecplot_writer_many_rows_in_order(tmp_path):
    ni, nj = 257, 1500                       # several batches of row chunks
    x = np.linspace(-1.0, 2.0, ni); y = np.linspace(0.0, 5.0, nj)
    data = np.arange(ni * nj, dtype=np.float64) * 0.001
    out = tmp_path / "rows.dat"
    ifx.write_results_to_file(x, y, data, ni, nj, str(out))
    lines = out.read_text().split("\n")
    assert len(lines) == 3 + ni * nj + 1
    for k in (0, 1, ni - 1, ni, 12345, ni * nj - 1):
        assert lines[3 + k] == "%f,%f,%f" % (x[k % ni], y[k // ni], data[k])


def test_create_without_gpu_fails_loudly(ref_case):
    """No silent CPU path: on a box without a CUDA device the solver refuses to exist."""
    lib = ifx.load_library()
    if lib.ifx_device_count() > 0:
        pytest.skip("GPU present")
    inp = ifx.make_input(50, 50, 1e-3, 150.0)
    with pytest.raises(ifx.IfxError, match="no CUDA device"):
        ifx.ImmerseFlow(inp, ref_case["xf"], ref_case["yf"])


def test_product_never_imports_the_oracle():
    """The oracle is test infrastructure; the shipped package must not reference it."""
    pkg = os.path.join(ROOT, "immerseflow_b200")
    for dirpath, _, files in os.walk(pkg):
        if "_build" in dirpath:
            continue
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                txt = open(os.path.join(dirpath, fn), errors="ignore").read()
                for line in txt.splitlines():
                    s = line.strip()
                    if s.startswith(("#include", "import ", "from ")) or "dlopen" in s or "CDLL" in s:
                        assert "oracle" not in s.replace("oracle/ and is never imported", ""), f"{fn}: {s}"
