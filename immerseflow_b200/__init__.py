"""immerseflow_b200 — B200-native (sm_100a) fractional-step path of ImmerseFlow++.

The product is the C-ABI shared library ``libimmerseflow_b200.so`` (include/immerseflow_c.h) built
from hand-written CUDA in ``csrc/``; this module is the thin ctypes binding plus a host-side mirror
of the reference's ``struct ImmerseFlow`` interface (reference src/header/globalVariables.cuh:70-88,
driven by src/main.cu:75-101) so the parity tests read like the reference's own driver.

There is NO CPU fallback: if the library is missing, importing the binding raises; if no GPU is
present, creating a solver raises.  The CPU oracle lives in ``oracle/`` and is never imported here.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass
from typing import Optional, Sequence, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# IFX_LIBRARY: explicit path of a build of the same C-ABI (kernel experiments, tools/build_variant.sh)
LIB_PATH = os.environ.get("IFX_LIBRARY") or os.path.join(_HERE, "libimmerseflow_b200.so")

IFX_ABI_VERSION = 2
IFX_COMPAT_REFERENCE, IFX_COMPAT_FULL = 0, 1
IFX_REDUCE_FUSED, IFX_REDUCE_REFERENCE = 0, 1

FIELD = {"u": 0, "v": 1, "p": 2, "iblank": 3, "uf": 4, "vf": 5, "sx": 6, "sy": 7, "ppe_rhs": 8,
         "xc": 9, "yc": 10, "celltype": 11}

# every symbol include/immerseflow_c.h declares (tests check the library exports all of them)
C_ABI_SYMBOLS = [
    "ifx_abi_version", "ifx_last_error", "ifx_device_count", "ifx_default_options",
    "ifx_read_input_file", "ifx_read_grid_file", "ifx_write_results_to_file",
    "ifx_create", "ifx_destroy", "ifx_initialize",
    "ifx_field_size", "ifx_set_field", "ifx_get_field", "ifx_save_field",
    "ifx_checkpoint_write", "ifx_checkpoint_read",
    "ifx_ad_solve", "ifx_ppe_solve", "ifx_correct", "ifx_step", "ifx_reduce_sum", "ifx_get_residual_history",
    "ifx_set_bodies", "ifx_iblank_update", "ifx_ghost_cell_count", "ifx_get_ghost_cells",
    "ifx_probe", "ifx_body_forces", "ifx_set_field_async", "ifx_get_field_async",
    "ifx_ipc_export", "ifx_ipc_connect",
    "ifx_set_stream", "ifx_synchronize", "ifx_launch_count",
]


class IfxInput(C.Structure):
    """Mirror of ``ifx_input`` == reference ``struct CFDInput`` (globalVariables.cuh:15-33)."""
    _fields_ = [("Restart", C.c_int), ("Restart_Time", C.c_int),
                ("nx", C.c_int), ("ny", C.c_int), ("nxf", C.c_int), ("nyf", C.c_int),
                ("Lx", C.c_double), ("Ly", C.c_double),
                ("w_AD", C.c_int), ("w_PPE", C.c_int), ("AD_itermax", C.c_int), ("PPE_itermax", C.c_int),
                ("AD_solver", C.c_int), ("PPE_solver", C.c_int),
                ("ErrorMax", C.c_double), ("tmax", C.c_double), ("dt", C.c_double), ("Re", C.c_double),
                ("mu", C.c_double), ("Write_Interval", C.c_int)]


class IfxBC(C.Structure):
    """Mirror of reference ``struct BC`` (globalVariables.cuh:35-39)."""
    _fields_ = [(n, C.c_double) for n in
                ("u_bc_w", "u_bc_e", "u_bc_n", "u_bc_s", "v_bc_w", "v_bc_e", "v_bc_n", "v_bc_s",
                 "p_bc_w", "p_bc_e", "p_bc_n", "p_bc_s")]


class IfxOptions(C.Structure):
    _fields_ = [("abi_version", C.c_int), ("device", C.c_int), ("compat", C.c_int), ("reduce_mode", C.c_int),
                ("bc", IfxBC), ("ad_tol", C.c_double), ("ppe_tol", C.c_double), ("ppe_abs_residual", C.c_int),
                ("rank", C.c_int), ("nranks", C.c_int), ("j_begin", C.c_int), ("j_end", C.c_int),
                ("sweeps_per_batch", C.c_int), ("use_graphs", C.c_int), ("ppe_solver", C.c_int), ("ppe_omega", C.c_double),
                ("zero_copy_control", C.c_int), ("ppe_pairs", C.c_int), ("reserved", C.c_int * 3)]


class IfxStepStats(C.Structure):
    _fields_ = [("ad_iters", C.c_int), ("ad_ures", C.c_double), ("ad_vres", C.c_double),
                ("ppe_sweeps", C.c_int), ("ppe_residual", C.c_double), ("exact_fallbacks", C.c_int),
                ("ms_ad", C.c_float), ("ms_ppe", C.c_float), ("ms_correct", C.c_float), ("ms_ib", C.c_float),
                ("ms_total", C.c_float), ("ms_ad_sweeps", C.c_float)]


class IfxError(RuntimeError):
    pass


_lib = None


def load_library() -> C.CDLL:
    """Load the CUDA C-ABI library.  Fails loudly: there is no other implementation."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise IfxError(f"{LIB_PATH} is missing — build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                       "(nvcc, sm_100a).  There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    dp = C.POINTER(C.c_double)
    ip = C.POINTER(C.c_int)
    vp = C.c_void_p
    lib.ifx_abi_version.restype = C.c_int
    lib.ifx_last_error.restype = C.c_char_p
    lib.ifx_last_error.argtypes = [vp]
    lib.ifx_device_count.restype = C.c_int
    lib.ifx_default_options.argtypes = [C.POINTER(IfxOptions)]
    lib.ifx_default_options.restype = None
    lib.ifx_read_input_file.argtypes = [C.c_char_p, C.POINTER(IfxInput)]
    lib.ifx_read_grid_file.argtypes = [C.c_char_p, C.c_int, dp]
    lib.ifx_write_results_to_file.argtypes = [dp, dp, dp, C.c_int, C.c_int, C.c_char_p]
    lib.ifx_create.argtypes = [C.POINTER(IfxInput), dp, dp, C.POINTER(IfxOptions), C.POINTER(vp)]
    lib.ifx_destroy.argtypes = [vp]
    lib.ifx_initialize.argtypes = [vp]
    lib.ifx_field_size.argtypes = [vp, C.c_int]
    lib.ifx_field_size.restype = C.c_size_t
    lib.ifx_set_field.argtypes = [vp, C.c_int, dp, C.c_size_t]
    lib.ifx_get_field.argtypes = [vp, C.c_int, dp, C.c_size_t]
    lib.ifx_set_field_async.argtypes = [vp, C.c_int, dp, C.c_size_t]
    lib.ifx_get_field_async.argtypes = [vp, C.c_int, dp, C.c_size_t]
    lib.ifx_save_field.argtypes = [vp, C.c_int, C.c_char_p]
    lib.ifx_checkpoint_write.argtypes = [vp, C.c_char_p, C.c_longlong, C.c_double]
    lib.ifx_checkpoint_read.argtypes = [vp, C.c_char_p, C.POINTER(C.c_longlong), C.POINTER(C.c_double)]
    for name in ("ifx_ad_solve", "ifx_ppe_solve", "ifx_correct", "ifx_step", "ifx_iblank_update"):
        getattr(lib, name).argtypes = [vp, C.POINTER(IfxStepStats)]
    lib.ifx_reduce_sum.argtypes = [vp, dp, C.c_size_t, dp]
    lib.ifx_get_residual_history.argtypes = [vp, dp, C.c_int]
    lib.ifx_set_bodies.argtypes = [vp, C.c_int, ip, dp, dp, dp, dp]
    lib.ifx_ghost_cell_count.argtypes = [vp]
    lib.ifx_get_ghost_cells.argtypes = [vp, ip, ip, dp, dp, dp, C.c_int]
    lib.ifx_probe.argtypes = [vp, C.c_int, dp, dp, dp, dp, dp]
    lib.ifx_body_forces.argtypes = [vp, dp, C.c_int]
    lib.ifx_ipc_export.argtypes = [vp, C.c_char_p]
    lib.ifx_ipc_connect.argtypes = [vp, C.c_char_p, C.c_int]
    lib.ifx_set_stream.argtypes = [vp, vp]
    lib.ifx_synchronize.argtypes = [vp]
    lib.ifx_launch_count.argtypes = [vp]
    lib.ifx_launch_count.restype = C.c_longlong
    if lib.ifx_abi_version() != IFX_ABI_VERSION:
        raise IfxError("C-ABI version mismatch between the binding and the library")
    _lib = lib
    return lib


def _dp(a: np.ndarray):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def read_input_file(path: str) -> IfxInput:
    """readInputFile (reference src/main.cu:10-59)."""
    lib = load_library()
    inp = IfxInput()
    rc = lib.ifx_read_input_file(os.fsencode(path), C.byref(inp))
    if rc != 0:
        raise IfxError(f"Unable to open file: {path} (status {rc})")
    return inp


def read_grid_file(path: str, n: int) -> np.ndarray:
    """Grid loops of readGridData (reference src/include/preSim.cu:268-291)."""
    lib = load_library()
    out = np.zeros(n)
    rc = lib.ifx_read_grid_file(os.fsencode(path), n, _dp(out))
    if rc != 0:
        raise IfxError(f"Error opening {os.path.basename(path)} (status {rc})")
    return out


def write_results_to_file(x, y, data, ni: int, nj: int, filename: str) -> None:
    """write_results_to_file (reference src/include/postSim.cu:41-66)."""
    lib = load_library()
    x = np.ascontiguousarray(x, dtype=np.float64)
    y = np.ascontiguousarray(y, dtype=np.float64)
    d = np.ascontiguousarray(data, dtype=np.float64).reshape(-1)
    rc = lib.ifx_write_results_to_file(_dp(x), _dp(y), _dp(d), ni, nj, os.fsencode(filename))
    if rc != 0:
        raise IfxError(f"Error opening file: {filename}")


def make_input(nx_cells: int, ny_cells: int, dt: float, Re: float, AD_itermax: int = 25,
               PPE_itermax: int = 100000, tmax: float = 1.0, Lx: float = 1.0, Ly: float = 1.0) -> IfxInput:
    """Build a CFDInput the way readInputFile leaves it (ghost-inclusive nx, ny; main.cu:55-58)."""
    inp = IfxInput()
    inp.nx, inp.ny = nx_cells + 2, ny_cells + 2
    inp.nxf, inp.nyf = nx_cells + 1, ny_cells + 1
    inp.Lx, inp.Ly = Lx, Ly
    inp.w_AD = inp.w_PPE = 1
    inp.AD_itermax, inp.PPE_itermax = AD_itermax, PPE_itermax
    inp.AD_solver = inp.PPE_solver = 1
    inp.ErrorMax, inp.tmax, inp.dt, inp.Re, inp.mu = 1e-6, tmax, dt, Re, 0.01
    inp.Write_Interval = 1000
    return inp


@dataclass
class StepStats:
    ad_iters: int = 0
    ad_ures: float = 0.0
    ad_vres: float = 0.0
    ppe_sweeps: int = 0
    ppe_residual: float = 0.0
    exact_fallbacks: int = 0
    ms_ad: float = 0.0
    ms_ppe: float = 0.0
    ms_correct: float = 0.0
    ms_ib: float = 0.0
    ms_total: float = 0.0
    ms_ad_sweeps: float = 0.0

    @classmethod
    def from_c(cls, s: IfxStepStats) -> "StepStats":
        return cls(**{f[0]: getattr(s, f[0]) for f in IfxStepStats._fields_})


class ImmerseFlow:
    """Host-side mirror of the reference's ``struct ImmerseFlow``.

    The reference driver (src/main.cu:75-101) reads::

        readInputFile("../inputs/inputs.txt", Solver);
        Solver.CUDAQuery(); Solver.allocation(); Solver.readGridData(); Solver.initializeData();
        for (...) Solver.ADsolver();
        Solver.freeAllocation();

    The same call sequence works here (CUDAQuery/allocation/readGridData collapse into the
    constructor); fields cross the boundary as numpy arrays in the reference's id = i + j*nx layout.
    """

    def __init__(self, inp: IfxInput, xf: Sequence[float], yf: Sequence[float], *, compat: int = IFX_COMPAT_REFERENCE,
                 reduce_mode: int = IFX_REDUCE_FUSED, device: int = 0, bc: Optional[dict] = None,
                 rank: int = 0, nranks: int = 1, j_begin: int = 0, j_end: int = 0,
                 sweeps_per_batch: int = 64, ppe_abs_residual: int = 0,
                 ad_tol: Optional[float] = None, ppe_tol: Optional[float] = None,
                 ppe_solver: int = 0, ppe_omega: float = 0.0, zero_copy_control: int = 0, use_graphs: Optional[int] = None,
                 ppe_pairs: int = 0):
        self.lib = load_library()
        self.Input = inp
        xf = np.ascontiguousarray(xf, dtype=np.float64)
        yf = np.ascontiguousarray(yf, dtype=np.float64)
        if xf.size != inp.nxf or yf.size != inp.nyf:
            raise IfxError(f"expected {inp.nxf} x-faces and {inp.nyf} y-faces")
        opt = IfxOptions()
        self.lib.ifx_default_options(C.byref(opt))
        opt.compat, opt.reduce_mode, opt.device = compat, reduce_mode, device
        opt.rank, opt.nranks, opt.j_begin, opt.j_end = rank, nranks, j_begin, j_end
        opt.sweeps_per_batch = sweeps_per_batch
        opt.ppe_abs_residual = ppe_abs_residual
        opt.ppe_solver, opt.ppe_omega = ppe_solver, ppe_omega      # 2 line SOR, 3 red-black SOR, 4 / 5 multigrid (full mode)
        opt.zero_copy_control = zero_copy_control
        opt.ppe_pairs = ppe_pairs                                  # two Jacobi sweeps per pass (opt-in, see immerseflow_c.h)
        if use_graphs is not None:                                 # library default: on
            opt.use_graphs = use_graphs
        if ad_tol is not None:
            opt.ad_tol = ad_tol
        if ppe_tol is not None:
            opt.ppe_tol = ppe_tol
        for k, v in (bc or {}).items():
            setattr(opt.bc, k, v)
        self.options = opt
        h = C.c_void_p()
        rc = self.lib.ifx_create(C.byref(inp), _dp(xf), _dp(yf), C.byref(opt), C.byref(h))
        if rc != 0:
            raise IfxError(f"ifx_create failed ({rc}): {self.lib.ifx_last_error(None).decode()}")
        self._h = h
        self.nx, self.ny = inp.nx, inp.ny

    # ---- lifetime ---------------------------------------------------------------------------
    def close(self) -> None:
        if getattr(self, "_h", None):
            self.lib.ifx_destroy(self._h)
            self._h = None

    freeAllocation = close   # preSim.cu:164-179

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc: int, what: str) -> None:
        if rc != 0:
            raise IfxError(f"{what} failed ({rc}): {self.lib.ifx_last_error(self._h).decode()}")

    # ---- reference-named entry points -----------------------------------------------------
    def initializeData(self) -> None:                       # preSim.cu:201-217
        self._check(self.lib.ifx_initialize(self._h), "ifx_initialize")

    def ADsolver(self) -> StepStats:                        # ADSolver.cu:268-395
        st = IfxStepStats()
        self._check(self.lib.ifx_ad_solve(self._h, C.byref(st)), "ifx_ad_solve")
        return StepStats.from_c(st)

    def PPESolver(self) -> StepStats:                       # PPESolver.cu:137-205
        st = IfxStepStats()
        self._check(self.lib.ifx_ppe_solve(self._h, C.byref(st)), "ifx_ppe_solve")
        return StepStats.from_c(st)

    def correct(self) -> StepStats:                         # AD_PPE_Correction.cu (empty in the reference)
        st = IfxStepStats()
        self._check(self.lib.ifx_correct(self._h, C.byref(st)), "ifx_correct")
        return StepStats.from_c(st)

    def step(self) -> StepStats:                            # main.cu:93-96
        st = IfxStepStats()
        self._check(self.lib.ifx_step(self._h, C.byref(st)), "ifx_step")
        return StepStats.from_c(st)

    def Reduction(self, values: np.ndarray) -> float:       # preSim.cu:376-445
        v = np.ascontiguousarray(values, dtype=np.float64).reshape(-1)
        out = C.c_double()
        self._check(self.lib.ifx_reduce_sum(self._h, _dp(v), v.size, C.byref(out)), "ifx_reduce_sum")
        return out.value

    def residual_history(self) -> np.ndarray:               # the reference's "iter = %d %f %f" lines, ADSolver.cu:369
        buf = np.zeros(128)
        n = self.lib.ifx_get_residual_history(self._h, _dp(buf), 64)
        return buf[:2 * n].reshape(n, 2).copy()

    def saveDataToFile(self, field: str, filename: str) -> None:   # postSim.cu:10-39
        self._check(self.lib.ifx_save_field(self._h, FIELD[field], os.fsencode(filename)), "ifx_save_field")

    # ---- immersed boundary --------------------------------------------------------------------
    def save_checkpoint(self, path: str, step: int = 0, time: float = 0.0) -> None:
        """Restart file: raw fp64 of u, v, p (+ uf, vf in full mode), see include/immerseflow_c.h."""
        self._check(self.lib.ifx_checkpoint_write(self._h, os.fsencode(path), C.c_longlong(step), C.c_double(time)),
                    "ifx_checkpoint_write")

    def load_checkpoint(self, path: str) -> Tuple[int, float]:
        step, time = C.c_longlong(0), C.c_double(0.0)
        self._check(self.lib.ifx_checkpoint_read(self._h, os.fsencode(path), C.byref(step), C.byref(time)),
                    "ifx_checkpoint_read")
        return int(step.value), float(time.value)

    def set_bodies(self, bodies: Sequence[np.ndarray], velocities: Optional[Sequence[Sequence[float]]] = None) -> None:
        offs = np.zeros(len(bodies) + 1, dtype=np.int32)
        for b, m in enumerate(bodies):
            offs[b + 1] = offs[b] + len(m)
        xm = np.ascontiguousarray(np.concatenate([np.asarray(m, dtype=np.float64)[:, 0] for m in bodies]))
        ym = np.ascontiguousarray(np.concatenate([np.asarray(m, dtype=np.float64)[:, 1] for m in bodies]))
        ub = vb = None
        if velocities is not None:
            ub = np.ascontiguousarray([v[0] for v in velocities], dtype=np.float64)
            vb = np.ascontiguousarray([v[1] for v in velocities], dtype=np.float64)
        self._check(self.lib.ifx_set_bodies(self._h, len(bodies), offs.ctypes.data_as(C.POINTER(C.c_int)), _dp(xm), _dp(ym),
                                            _dp(ub) if ub is not None else None, _dp(vb) if vb is not None else None),
                    "ifx_set_bodies")

    def iblank_update(self) -> StepStats:
        st = IfxStepStats()
        self._check(self.lib.ifx_iblank_update(self._h, C.byref(st)), "ifx_iblank_update")
        return StepStats.from_c(st)

    def ghost_cells(self) -> dict:
        n = self.lib.ifx_ghost_cell_count(self._h)
        cell = np.zeros(n, dtype=np.int32)
        sten = np.zeros((n, 4), dtype=np.int32)
        w = np.zeros((n, 10))
        bi = np.zeros((n, 2))
        ip = np.zeros((n, 2))
        if n:
            self._check(self.lib.ifx_get_ghost_cells(self._h, cell.ctypes.data_as(C.POINTER(C.c_int)),
                                                     sten.ctypes.data_as(C.POINTER(C.c_int)), _dp(w), _dp(bi), _dp(ip), n),
                        "ifx_get_ghost_cells")
        return {"cell": cell, "stencil": sten, "weights": w, "bi": bi, "ip": ip}

    # ---- diagnostics ----------------------------------------------------------------
    def probe(self, x, y):
        """u, v, p interpolated at the points (x[k], y[k])."""
        x = np.ascontiguousarray(x, dtype=np.float64).reshape(-1)
        y = np.ascontiguousarray(y, dtype=np.float64).reshape(-1)
        u, v, p = np.zeros(x.size), np.zeros(x.size), np.zeros(x.size)
        self._check(self.lib.ifx_probe(self._h, x.size, _dp(x), _dp(y), _dp(u), _dp(v), _dp(p)), "ifx_probe")
        return u, v, p

    def body_forces(self, nbodies: int) -> np.ndarray:
        """(nbodies, 4): pressure force x, y and viscous force x, y on every body."""
        F = np.zeros((max(nbodies, 1), 4))
        self._check(self.lib.ifx_body_forces(self._h, _dp(F.reshape(-1)), max(nbodies, 1)), "ifx_body_forces")
        return F[:nbodies]

    # ---- state ------------------------------------------------------------------------------------
    def field_size(self, name: str) -> int:
        return int(self.lib.ifx_field_size(self._h, FIELD[name]))

    def get(self, name: str, out: Optional[np.ndarray] = None) -> np.ndarray:
        n = self.field_size(name)
        if out is None:
            out = np.empty(n)
        self._check(self.lib.ifx_get_field(self._h, FIELD[name], _dp(out), n), f"ifx_get_field({name})")
        return out

    def set(self, name: str, values: np.ndarray) -> None:
        v = np.ascontiguousarray(values, dtype=np.float64).reshape(-1)
        self._check(self.lib.ifx_set_field(self._h, FIELD[name], _dp(v), v.size), f"ifx_set_field({name})")

    def set_async(self, name: str, values: np.ndarray) -> None:
        """Enqueue the upload on this handle's stream; `values` must be a contiguous float64 array in page-locked memory
        that stays untouched until synchronize() or the next blocking call on this handle."""
        if values.dtype != np.float64 or not values.flags["C_CONTIGUOUS"]:
            raise IfxError("set_async needs a contiguous float64 array (no hidden copy may be made)")
        self._check(self.lib.ifx_set_field_async(self._h, FIELD[name], _dp(values), values.size), f"ifx_set_field_async({name})")

    def get_async(self, name: str, out: np.ndarray) -> None:
        if out.dtype != np.float64 or not out.flags["C_CONTIGUOUS"]:
            raise IfxError("get_async needs a contiguous float64 array")
        self._check(self.lib.ifx_get_field_async(self._h, FIELD[name], _dp(out), out.size), f"ifx_get_field_async({name})")

    def set_stream(self, cuda_stream: int) -> None:
        self._check(self.lib.ifx_set_stream(self._h, C.c_void_p(cuda_stream)), "ifx_set_stream")

    def synchronize(self) -> None:
        self._check(self.lib.ifx_synchronize(self._h), "ifx_synchronize")

    @property
    def launch_count(self) -> int:
        return int(self.lib.ifx_launch_count(self._h))

    # ---- multi-GPU ------------------------------------------------------------------------------
    def ipc_export(self) -> bytes:
        buf = C.create_string_buffer(128)
        self._check(self.lib.ifx_ipc_export(self._h, buf), "ifx_ipc_export")
        return buf.raw

    def ipc_connect(self, all_handles: bytes, nranks: int) -> None:
        self._check(self.lib.ifx_ipc_connect(self._h, all_handles, nranks), "ifx_ipc_connect")


def uniform_faces(n_cells: int, length: float = 1.0) -> np.ndarray:
    """Faces as inputs/uniformGrid.py writes and the reference parses them (7 significant digits)."""
    mesh = np.linspace(0, length, n_cells + 1)
    return np.array([float(f"{v:.7E}") for v in mesh])
