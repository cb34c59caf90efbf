#!/usr/bin/env python
"""bench.py — Mcell-steps/s of the fractional-step hot path on N B200s (BASELINE.json's metric).

    python bench.py --gpus 1 --steps K --warmup W             # our CUDA path
    python bench.py --impl reference ...                      # CPU arm: OpenMP transcription of the reference loops
    torchrun --nproc-per-node N bench.py --gpus N ...         # slab-decomposed, one rank per GPU

A "step" is one pass of the hot path over the whole grid: predictor (source + Jacobi iterations to the
reference's stop rule / AD_itermax) followed by the Poisson solve (sweeps to the stop rule / PPE_itermax) and, in
full mode, the projection.  Prints ONE JSON line (rank 0).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "Mcell-steps/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=6)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--nx", type=int, default=16384, help="cells in x (workload: 16384x16384, north_star's scaling grid)")
    ap.add_argument("--ny", type=int, default=16384)
    ap.add_argument("--ad-itermax", type=int, default=25, help="AD_itermax of the shipped inputs.txt")
    ap.add_argument("--ppe-sweeps", type=int, default=50, help="PPE_itermax for the bench step (bounded Poisson solve)")
    ap.add_argument("--dt", type=float, default=1e-3)
    ap.add_argument("--Re", type=float, default=150.0)
    ap.add_argument("--cpu-sample-rows", type=int, default=2048,
                    help="rows of the workload the CPU arm runs on (a row slab through the first lattice row of bodies; >= 1/8 of the grid)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity-check", action="store_true",
                    help="skip the bit-for-bit check of the timed workload's first step against the CPU oracle (rank 0) and, on "
                         "more than one GPU, of a slab run against a single-GPU run")
    ap.add_argument("--ppe-pairs", type=int, default=0,
                    help="1: two Jacobi sweeps per pass over memory in the Poisson solve (ifx_options.ppe_pairs; single GPU; measured "
                         "bit-identical and NOT faster, profiles/r2_pair_kernel.md)")
    ap.add_argument("--slab-balance", default="rows", choices=["rows", "cost"],
                    help="more than one GPU: slabs of equal height (default) or of equal estimated cost (rows inside bodies counted as "
                         "cheaper for the Poisson sweep — measured at 8 GPUs: slower, 20.4 against 19.7 ms per step, the partly solid "
                         "rows are not cheaper in practice; kept as an experiment switch)")
    ap.add_argument("--no-secondary", action="store_true",
                    help="skip the short runs of BASELINE.json's other configurations (cavity 1024^2, cylinder 4096x2048, converged "
                         "pressure solves) that the default single-GPU run appends under `configs`")
    ap.add_argument("--no-ref-cuda", action="store_true",
                    help="skip timing the reference's own CUDA build (oracle/_ref/immerseFlow_ref) on this GPU")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--e2e-handles", type=int, default=3,
                    help="simulations in flight in the end-to-end measurement: 3 = one uploading its next state, one stepping, "
                         "one downloading its result (PCIe both ways overlapped with the kernels); 1 = upload, step, download in series")
    ap.add_argument("--mode", default="full", choices=["full", "reference"],
                    help="full: complete fractional step (predictor + Poisson with source + projection); reference: what the "
                         "reference binary's time step runs (predictor) + its Laplace-Jacobi Poisson kernels")
    ap.add_argument("--bodies", type=int, default=8,
                    help="moving immersed bodies (BASELINE.json configs[4]: multiple complex-shaped bodies on 16384x16384, iBlank "
                         "and ghost cells recomputed every step); 0 = no body")
    ap.add_argument("--emulate-slab-of", type=int, default=0,
                    help="diagnostic: run ONE rank's slab of an N-way decomposition on one GPU, no exchange (timing only)")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clocks + throttle reasons DURING the timed region (B200_PROFILING.md's clocks line), read through NVML
    from a sampling thread every 10 ms (the solver calls release the GIL); `nvidia-smi -lms` takes longer to start
    than a short multi-GPU timed region lasts, so it is only the fallback."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.proc = None
        self.path = None
        self.nvml = None
        self.handle = None
        self.thread = None
        self.stop_flag = False
        self.sm, self.reasons, self.mx = [], set(), None
        try:
            import pynvml
            pynvml.nvmlInit()
            h = None
            try:
                import torch
                pr = torch.cuda.get_device_properties(gpu_index)
                bus = "%08x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
                h = pynvml.nvmlDeviceGetHandleByPciBusId(bus.encode())
            except Exception:
                h = pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
            self.nvml, self.handle = pynvml, h
            self.mx = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nvml = None

    def _sample(self):
        n = self.nvml
        try:
            self.sm.append(float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)))
            r = n.nvmlDeviceGetCurrentClocksEventReasons(self.handle) if hasattr(n, "nvmlDeviceGetCurrentClocksEventReasons") \
                else n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
            for name, bit in (("hw_slowdown", 0x8), ("sw_power_cap", 0x4), ("sw_thermal_slowdown", 0x20),
                              ("hw_thermal_slowdown", 0x40)):
                if r & bit:
                    self.reasons.add(name)
        except Exception:
            pass

    def _loop(self):
        while not self.stop_flag:
            self._sample()
            time.sleep(0.01)

    def start(self):
        if self.nvml:
            import threading
            self.stop_flag = False
            self.thread = threading.Thread(target=self._loop, daemon=True)
            self.thread.start()
            return
        try:
            f = tempfile.NamedTemporaryFile(prefix="ifx_clocks_", suffix=".csv", delete=False)
            self.path = f.name
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.gpu)], stdout=f, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self) -> dict:
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.nvml:
            self.stop_flag = True
            if self.thread:
                self.thread.join(timeout=2)
            if self.sm:
                s = sorted(self.sm)
                out.update(sm_mhz=float(np.median(s[len(s) // 2:])), sm_max_mhz=self.mx, samples=len(s), source="nvml")
            out["reasons"] = sorted(self.reasons)
            return out
        if not self.proc:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        try:
            for line in open(self.path):
                c = [x.strip() for x in line.split(",")]
                if len(c) < 9:
                    continue
                try:
                    sm.append(float(c[1])); mx.append(float(c[2]))
                except ValueError:
                    continue
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            # median over the samples under load (top half), the idle ones before/after are not the kernel's clocks
            s = sorted(sm)
            out["sm_mhz"] = float(np.median(s[len(s) // 2:]))
            out["sm_max_mhz"] = max(mx)
            out["samples"] = len(sm)
            out["source"] = "nvidia-smi"
        out["reasons"] = sorted(reasons)
        return out


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def bodies_at(nb: int, step: int, dt: float):
    """nb lobed closed polygons (256 markers each) on a lattice over the unit square, translating with a constant
    velocity each: marker arrays + body velocities.  The same arrays go to the CUDA path and to the CPU arm."""
    if nb <= 0:
        return [], []
    cols = int(np.ceil(np.sqrt(nb)))
    rows = (nb + cols - 1) // cols
    th = 2.0 * np.pi * np.arange(256) / 256
    out, vel = [], []
    for b in range(nb):
        cx, cy = (b % cols + 0.5) / cols, (b // cols + 0.5) / rows
        r0 = 0.22 / max(cols, rows)
        ub, vb = 0.05 * (1 if b % 2 == 0 else -1), 0.03 * (1 if (b // 2) % 2 == 0 else -1)
        r = r0 * (1.0 + 0.25 * np.cos((3 + b % 4) * th + 0.3 * b))
        out.append(np.ascontiguousarray(np.stack([cx + ub * dt * step + r * np.cos(th), cy + vb * dt * step + r * np.sin(th)], axis=1)))
        vel.append((ub, vb))
    return out, vel


def row_costs(nb: int, ncx: int, ncy: int, dt: float) -> np.ndarray:
    """Relative cost of every interior row of the workload for the slab partition: the Poisson sweep carries a cell that
    is not fluid over instead of relaxing it (about a third of the work), the predictor sweep costs the same everywhere.
    Solid fraction of a row = total chord length of the body polygons (step 0) along the row's cell centres."""
    b, _ = bodies_at(nb, 0, dt)
    yc = (np.arange(ncy) + 0.5) / ncy
    solid = np.zeros(ncy)
    for m in b:
        x0, y0 = m[:, 0], m[:, 1]
        x1, y1 = np.roll(x0, -1), np.roll(y0, -1)
        lo, hi = np.searchsorted(yc, m[:, 1].min()), np.searchsorted(yc, m[:, 1].max())
        if hi <= lo:
            continue
        Y = yc[lo:hi, None]
        cross = (y0[None, :] > Y) != (y1[None, :] > Y)
        dy = np.where(y1 == y0, 1.0, y1 - y0)
        xi = np.where(cross, x0[None, :] + (Y - y0[None, :]) * ((x1 - x0) / dy)[None, :], np.inf)
        xi.sort(axis=1)
        xi = np.where(np.isfinite(xi), xi, 0.0)
        solid[lo:hi] += np.clip((xi[:, 1::2] - xi[:, 0::2]).sum(axis=1), 0.0, 1.0)
    s = np.clip(solid, 0.0, 1.0)
    share_ad, share_ppe, share_rest, c_solid = 0.41, 0.47, 0.12, 0.35       # measured stage shares of the single-GPU step
    return share_ad + share_rest + share_ppe * (1.0 - s * (1.0 - c_solid))


# ------------------------------------------------------------------------------------------------
PARITY_MARGIN = 200     # rows.  One step moves information by at most 25 predictor iterations x 4 cells (a ghost cell's
                        # closure reads nodes up to 4 cells away) + 50 Poisson sweeps x 1 cell + a few cells (faces, source
                        # term, projection): < 200, so the oracle's artificial slab walls cannot reach further in


def cpu_slab_rows(args, rows: int):
    """Row slab [r0, r0 + ncy) (cell rows) of the workload the CPU arm runs on: through the first lattice row of bodies."""
    ncy = min(rows, args.ny)
    r0 = 0
    if args.bodies > 0 and args.mode == "full":
        cols = int(np.ceil(np.sqrt(args.bodies)))
        nrow = (args.bodies + cols - 1) // cols
        r0 = max(0, min(args.ny - ncy, int(0.5 / nrow * args.ny) - ncy // 2))
    return r0, ncy


def initial_pressure(nx: int, ny_global: int, row0: int, nrows: int) -> np.ndarray:
    """The bench's smooth non-zero starting pressure on rows [row0, row0 + nrows) of the ghost-inclusive grid (the
    reference starts from p == 0, which turns most quotients of the first sweeps into exact zeros — an arithmetic special
    case).  A function of the GLOBAL cell index, so every slab of every decomposition (and the CPU arm) sees the same field."""
    jj, ii = np.divmod(np.arange(nrows * nx, dtype=np.float64), float(nx))
    jj += row0
    return 50.0 + 40.0 * np.sin(ii * (6.283185307179586 / nx)) * np.cos(jj * (6.283185307179586 / ny_global))


def host_threads() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def cpu_reference_arm(args, rows: int, steps: int, warmup: int, keep_first_step: bool = False, ic=None):
    """The reference has no CPU path; its reported CPU baseline is the OpenMP transcription of its loops
    (oracle/, `kind: port`), run with every host thread (set explicitly: torchrun exports OMP_NUM_THREADS=1) on a
    bounded row slab of the same workload (same grid width, same parameters, same iteration caps, `rows` of the rows).
    keep_first_step: also return the slab's u, v, p after the first step from the initial condition (parity check).
    ic: the slab's rows of the initial u, v as the CUDA path holds them (its vortex kernel uses CUDA's exp / pow, glibc's
    differ in the last bit: a parity check of the STEP needs the same input on both sides)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import _oracle as orc
    import immerseflow_b200 as ifx
    r0, ncy = cpu_slab_rows(args, rows)
    ncx = args.nx
    xf = ifx.uniform_faces(ncx, 1.0)
    yf = ifx.uniform_faces(args.ny, 1.0)[r0: r0 + ncy + 1]
    orc.lib().orc_set_num_threads(host_threads())
    cores = orc.lib().orc_num_threads()
    times, k_ad, k_ppe, first = [], 0, 0, None
    if args.mode == "full":
        fs = orc.FullSolver(xf, yf, args.dt, args.Re, args.ad_itermax, args.ppe_sweeps, ppe_tol=0.0, ppe_abs=1)
        g = orc.Grid(xf, yf)
        u, v, p = orc.initial_condition(g)
        if ic is not None:
            u, v = ic["u"].reshape(-1), ic["v"].reshape(-1)
        fs.set("u", u); fs.set("v", v)
        fs.set("p", initial_pressure(g.nx, args.ny + 2, r0, g.ny))
        fs.update_ib()
        for it in range(warmup + steps):
            t0 = time.perf_counter()
            if args.bodies > 0:
                b, vel = bodies_at(args.bodies, it, args.dt)
                fs.set_bodies(b, vel)
                fs.update_ib()
            st = fs.step()
            if it >= warmup:
                times.append(time.perf_counter() - t0)
            k_ad, k_ppe = int(st[0]), int(st[3])
            if it == 0 and keep_first_step:
                first = {k: fs.get(k) for k in ("u", "v", "p")}
        fs.close()
    else:
        g = orc.Grid(xf, yf)
        u, v, p = orc.initial_condition(g)
        if ic is not None:
            u, v = ic["u"].reshape(-1), ic["v"].reshape(-1)
        p = initial_pressure(g.nx, args.ny + 2, r0, g.ny)
        pr = orc.Predictor(g, u, v, args.dt, args.Re, args.ad_itermax)
        for it in range(warmup + steps):
            t0 = time.perf_counter()
            k_ad, _ = pr.step()
            k_ppe, p, _ = orc.ppe_solve(g, p, args.ppe_sweeps)
            if it >= warmup:
                times.append(time.perf_counter() - t0)
            if it == 0 and keep_first_step:
                first = {"u": pr.u.copy(), "v": pr.v.copy(), "p": p.copy()}
    cells = ncx * ncy
    t = float(np.mean(times)) if times else float("nan")
    out = {"value": cells / t / 1e6, "unit": METRIC, "cores": cores, "kind": "port",
           "sample": f"{ncx}x{ncy}-cell row slab (rows {r0}..{r0 + ncy} of {args.ny}, {100.0 * ncy / args.ny:.1f} % of the grid; throughput "
                     f"per cell of the slab) of the workload ({args.mode} step), {steps} step(s) after {warmup} warm-up "
                     f"(K_AD={k_ad}, {k_ppe} Poisson sweeps), OpenMP x{cores}", "ms_per_step": t * 1e3,
           "sample_rows": [r0, r0 + ncy], "sample_fraction": ncy / args.ny}
    if keep_first_step:
        out["first_step"] = first
        out["grid"] = (g.nx, g.ny)
    return out


def row_digests(field: np.ndarray, nx: int, rows) -> list:
    """One 16-byte digest per row of a reference-layout field: what the ranks exchange instead of gigabytes."""
    import hashlib
    f = field.reshape(-1, nx)
    return [hashlib.blake2b(np.ascontiguousarray(f[r, 1:-1]).tobytes(), digest_size=16).digest() for r in rows]


def ref_cuda_baseline(args):
    """The reference's OWN CUDA build (oracle/_ref/immerseFlow_ref: its unmodified translation units, file output
    stubbed) timed on this GPU — north_star's second reported baseline.  Its time step is the predictor only
    (src/main.cu:93-96 -> ADSolver.cu:268-395), all cells fluid.  Per-step time = difference of two runs (start-up,
    allocation and grid I/O cancel).  Bounded: one run of 4 steps."""
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    try:
        import ref_cuda_baseline as rcb
        import shutil
        if not os.path.exists(rcb.REF):
            return {"unavailable": "oracle/_ref/immerseFlow_ref not built here (needs /root/reference at build time)"}
        if args.nx > args.ny:
            return {"unavailable": "the reference overruns its vf allocation for nx > ny (preSim.cu:153)"}
        w = tempfile.mkdtemp(prefix="ifx_ref_")
        try:
            for d in ("src", "inputs", "results"):
                os.makedirs(os.path.join(w, d))
            rcb.write_grid(os.path.join(w, "inputs", "xgrid.dat2"), args.nx)
            rcb.write_grid(os.path.join(w, "inputs", "ygrid.dat2"), args.ny)
            nsteps = 4
            ta, ia, ends = rcb.run(args.nx, args.ny, args.ad_itermax, args.dt, args.Re, nsteps, w)
        finally:
            shutil.rmtree(w, ignore_errors=True)
        if len(ends) < 2:
            return {"unavailable": "the reference did not report the end of two steps (converged before AD_itermax?)"}
        per = (ends[-1] - ends[0]) / (len(ends) - 1)
        return {"impl": "reference CUDA build (unmodified TUs of /root/reference/src, nvcc -arch=sm_100, file output stubbed)",
                "work": "the reference's time step = predictor only (src/main.cu:93-96), no immersed body",
                "ms_per_step": per * 1e3, "value": args.nx * args.ny / per / 1e6, "unit": METRIC,
                "predictor_iterations_per_step": ia / nsteps, "wall_s": ta, "step_end_times_s": ends,
                "method": "spacing of the reference's own `iter = <AD_itermax>` lines (end of every time step) on a line-buffered "
                          "pipe: start-up, allocation and grid I/O are outside"}
    except Exception as ex:       # noqa: BLE001 — reported, never fatal
        return {"unavailable": f"{type(ex).__name__}: {ex}"[:300]}


def measured_traffic(kernel: str, nx: int, ny: int):
    """DRAM bytes per launch of a sweep kernel from the committed ncu capture (profiles/r2_traffic.json), valid
    for the grid it was captured on; None otherwise."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "r2_traffic.json")))
        e = t.get(kernel)
        if e and e["grid"] == [nx, ny]:
            return e["dram_bytes_per_launch"]
    except Exception:
        pass
    return None


def secondary_configs(ifx, torch, dev, args):
    """BASELINE.json's other configurations on ONE GPU, each for a few steps (bounded: well under a minute in total), so that
    the driver's record carries them too: configs[1] (lid-driven cavity, uniform 1024 x 1024 — L2-resident, launch-latency
    bound), configs[2] (cylinder on the stretched 4096 x 2048 grid, checked against the CPU oracle in every cell), and steps
    whose pressure solve actually CONVERGES (multigrid; point Jacobi never gets there) — to the reference's criterion with the
    tolerance scaled per cell and capped at 0.5: its residual is an un-normalised sum and its loop starts from a residual of
    1.0 (PPESolver.cu:170-172), so a tolerance of 1 or more means "no sweep"."""
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import make_case
    out = []

    def digits(f):
        return np.array([float(f"{v:.7E}") for v in f])          # the grid files carry 7 significant digits

    def run(name, what, ncx, ncy, xf, yf, dt, Re, *, bodies=None, vel=None, bc=None, ic="vortex", ppe_solver=1, ppe_itermax=50,
            ppe_tol=0.0, ad_tol=None, steps=5, warmup=2, moving=0, lx=1.0, ly=1.0, extra=None):
        inp = ifx.make_input(ncx, ncy, dt, Re, AD_itermax=args.ad_itermax, PPE_itermax=ppe_itermax, Lx=lx, Ly=ly)
        rec = {"name": name, "what": what, "grid": [ncx, ncy], "ppe_solver": ppe_solver}
        try:
            with ifx.ImmerseFlow(inp, xf, yf, device=dev, compat=ifx.IFX_COMPAT_FULL, ppe_abs_residual=1, ppe_solver=ppe_solver,
                                 ppe_tol=ppe_tol, ad_tol=ad_tol, bc=bc, sweeps_per_batch=ppe_itermax + 1) as h:
                h.initializeData()
                n = inp.nx * inp.ny
                if ic == "zero":
                    h.set("u", np.zeros(n)); h.set("v", np.zeros(n))
                elif ic == "uniform":
                    h.set("u", np.ones(n)); h.set("v", np.zeros(n))
                stepno = 0

                def one():
                    nonlocal stepno
                    if bodies is not None and (moving or stepno == 0):
                        b, v_ = bodies(stepno) if callable(bodies) else (bodies, vel)
                        h.set_bodies(b, v_)
                    stepno += 1
                    return h.step()
                for _ in range(warmup):
                    one()
                torch.cuda.synchronize()
                l0, t0 = h.launch_count, time.perf_counter()
                sts = [one() for _ in range(steps)]
                torch.cuda.synchronize()
                t = (time.perf_counter() - t0) / steps
                k_ppe = float(np.mean([q.ppe_sweeps for q in sts]))
                rec.update({"ms_per_step": t * 1e3, "value": ncx * ncy / t / 1e6, "unit": METRIC, "steps": steps, "warmup": warmup,
                            "gpu_launches_per_step": (h.launch_count - l0) / steps,
                            "predictor_iterations": float(np.mean([q.ad_iters for q in sts])),
                            "poisson_iterations": k_ppe, "poisson_residual": float(sts[-1].ppe_residual), "poisson_tolerance": ppe_tol,
                            "poisson_converged": bool(ppe_tol > 0 and sts[-1].ppe_residual <= ppe_tol),
                            "ms_poisson": float(np.mean([q.ms_ppe for q in sts])), "ms_predictor": float(np.mean([q.ms_ad for q in sts])),
                            "ghost_cells": int(h.lib.ifx_ghost_cell_count(h._h))})
                if ppe_solver == 1:
                    rec["us_per_poisson_sweep"] = rec["ms_poisson"] * 1e3 / (k_ppe + 1)
                    rec["us_per_predictor_sweep"] = float(np.mean([q.ms_ad_sweeps for q in sts])) * 1e3 / max(rec["predictor_iterations"], 1)
                if extra:
                    rec.update(extra)
        except Exception as ex:       # noqa: BLE001 — a secondary line never takes the headline down
            rec["error"] = f"{type(ex).__name__}: {ex}"[:300]
        out.append(rec)
        return rec

    # ---- configs[1]: lid-driven cavity Re = 1000 (input-file Re = 2000: the predictor keeps the reference's factor 1/2)
    n1 = 1024
    xf1 = ifx.uniform_faces(n1, 1.0)
    dt1 = 0.25 / n1
    lid = dict(u_bc_w=0.0, u_bc_e=0.0, u_bc_s=0.0, u_bc_n=1.0)
    l2 = {"l2": "8 fields x 8.4 MB = 67 MB: the working set is L2-resident (126 MB), the sweeps run from L2 and the step is "
                "launch-latency bound — HBM roofline fractions do not apply to this configuration"}
    run("cavity1024", "BASELINE configs[1]: lid-driven cavity Re=1000, uniform 1024x1024, no body; 25 predictor iterations + 50 Jacobi "
        "Poisson sweeps per step (the bench step)", n1, n1, xf1, xf1, dt1, 2000.0, bc=lid, ic="zero", steps=20, warmup=3, extra=l2)
    run("cavity1024_converged", "configs[1], pressure solved to a grid-scaled tolerance by multigrid (PPE_Solver 4, V(2,2) cycles)",
        n1, n1, xf1, xf1, dt1, 2000.0, bc=lid, ic="zero", ppe_solver=4, ppe_tol=min(0.5, 1e-9 * n1 * n1 / dt1), ad_tol=min(0.5, 4e-10 * n1 * n1),
        steps=10, warmup=3, extra=l2)

    # ---- configs[2]: circular cylinder Re = 300 on the stretched 4096 x 2048 grid (tools/make_case.py)
    try:
        tmp = tempfile.mkdtemp(prefix="ifx_case_")
        case = make_case.build("cylinder", tmp, 1.0, 2)
        import shutil
        shutil.rmtree(tmp, ignore_errors=True)
        ncx, ncy = case["cells"]
        xf2, yf2 = digits(case["xf"]), digits(case["yf"])
        dt2, re2 = case["dt"], case["Re_file"]
        cyl = [np.ascontiguousarray(m) for m, _ in case["bodies"]]
        par = None
        if not args.no_parity_check:
            # every interior cell of u, v, p after two steps + the iteration counts + the ghost-cell maps, vs the CPU oracle
            import _oracle as orc
            orc.lib().orc_set_num_threads(host_threads())
            inp = ifx.make_input(ncx, ncy, dt2, re2, AD_itermax=args.ad_itermax, PPE_itermax=args.ppe_sweeps, Lx=40.0, Ly=20.0)
            o = orc.FullSolver(xf2, yf2, dt2, re2, args.ad_itermax, args.ppe_sweeps, ppe_abs=1)
            nn = inp.nx * inp.ny
            with ifx.ImmerseFlow(inp, xf2, yf2, device=dev, compat=ifx.IFX_COMPAT_FULL, ppe_abs_residual=1) as h:
                h.initializeData()
                u0, v0 = np.ones(nn), np.zeros(nn)
                h.set("u", u0); h.set("v", v0); o.set("u", u0); o.set("v", v0)
                h.set_bodies(cyl, [(0.0, 0.0)]); o.set_bodies(cyl, [(0.0, 0.0)]); o.update_ib()
                inner = np.zeros((inp.ny, inp.nx), bool); inner[1:-1, 1:-1] = True
                inner = inner.reshape(-1)
                bad, counts_ok = 0, True
                for _ in range(2):
                    st = h.step(); so = o.step()
                    counts_ok = counts_ok and (st.ad_iters, st.ppe_sweeps) == (int(so[0]), int(so[3]))
                    for k in ("u", "v", "p"):
                        bad += int(np.count_nonzero(h.get(k)[inner] != o.get(k)[inner]))
                g1, go = h.ghost_cells(), o.ghost_cells()
                maps_ok = all(np.array_equal(g1[k], go[k]) for k in g1)
            o.close()
            par = {"bit_exact": bad == 0 and counts_ok and maps_ok, "cells": int(ncx * ncy), "steps": 2, "mismatching_values": bad,
                   "iteration_counts_equal": counts_ok, "ghost_cell_maps_equal": maps_ok,
                   "what": "u, v, p in every interior cell after each of two steps, iteration counts, ghost-cell index maps and weights "
                           "vs the CPU oracle"}
        run("cylinder4096x2048", "BASELINE configs[2]: circular cylinder Re=300 (D/dx = 256), stretched 4096x2048 grid on 40x20, "
            "sharp-interface body; 25 predictor iterations + 50 Jacobi Poisson sweeps per step", ncx, ncy, xf2, yf2, dt2, re2,
            bodies=cyl, vel=[(0.0, 0.0)], ic="uniform", steps=10, warmup=3, lx=40.0, ly=20.0,
            extra={"parity_check": par, "l2": "8 fields x 67 MB = 0.54 GB: streams from HBM"})
        run("cylinder4096x2048_converged", "configs[2], pressure solved to a grid-scaled tolerance by line-smoothed multigrid "
            "(PPE_Solver 5: the point smoother stalls on cell aspect ratios of 30)", ncx, ncy, xf2, yf2, dt2, re2, bodies=cyl,
            vel=[(0.0, 0.0)], ic="uniform", ppe_solver=5, ppe_itermax=60, ppe_tol=min(0.5, 1e-9 * ncx * ncy / dt2), ad_tol=min(0.5, 4e-10 * ncx * ncy),
            steps=2, warmup=1, lx=40.0, ly=20.0)
    except Exception as ex:       # noqa: BLE001
        out.append({"name": "cylinder4096x2048", "error": f"{type(ex).__name__}: {ex}"[:300]})

    # ---- the headline workload with a CONVERGED pressure solve (multigrid, bodies re-classified every step)
    nxm, nym = args.nx, args.ny
    if args.mode == "full" and nxm * nym <= 16384 * 16384:
        dtm = args.dt
        xfm, yfm = ifx.uniform_faces(nxm, 1.0), ifx.uniform_faces(nym, 1.0)
        # At 268 M cells the reference's criterion is out of reach in fp64: the un-normalised residual sum has a round-off
        # floor of order 0.1-1 (2^-53 x the sum of the magnitudes of 268 M x 5 terms of order 1e7), the same order as the
        # largest tolerance its loop allows (< 1).  So the 16384^2 line runs a FIXED number of V-cycles and reports the
        # residual they reach next to the one 50 Jacobi sweeps leave.
        run("workload_multigrid", f"the headline workload ({nxm}x{nym}, {args.bodies} moving bodies) with 16 multigrid V-cycles "
            "(PPE_Solver 4) per step instead of 50 Jacobi sweeps: residual sum 1e7 -> order 10 (Jacobi-50 leaves it at order 1e7-1e8; "
            "the reference's absolute criterion, tolerance < 1 on the un-normalised sum, is at the fp64 round-off floor of a "
            "268 M-cell sum and is not attainable at this size)", nxm, nym, xfm, yfm, dtm, args.Re,
            bodies=(lambda k: bodies_at(args.bodies, k, dtm)) if args.bodies else None, moving=1, ppe_solver=4, ppe_itermax=16,
            ppe_tol=0.5, steps=2, warmup=1)
    return out


def slab_parity_check(ifx, slabs, dist, torch, rank, world, dev):
    """world > 1: a 2048 x 2048 full-mode run with moving bodies that straddle the slab boundaries, slab-decomposed
    over all ranks, against the same run on ONE GPU (rank 0): every owned row of u, v, p, bit for bit (row digests),
    and the iteration counts.  What tests/test_gpu_slabs.py checks, inside the benchmark's own process group."""
    ncx = ncy = 2048
    dt, Re = 1e-3, 150.0
    inp = ifx.make_input(ncx, ncy, dt, Re, AD_itermax=12, PPE_itermax=40)
    xf, yf = ifx.uniform_faces(ncx, 1.0), ifx.uniform_faces(ncy, 1.0)
    nsteps, nb = 2, 4
    kw = dict(compat=ifx.IFX_COMPAT_FULL, ppe_abs_residual=1, sweeps_per_batch=16)

    def run(h, lo_row):
        h.initializeData()
        h.set("p", initial_pressure(inp.nx, inp.ny, lo_row, h.field_size("p") // inp.nx))
        return h

    jb, je = slabs.partition_rows(inp.ny, world)[rank]
    h = run(ifx.ImmerseFlow(inp, xf, yf, device=dev, rank=rank, nranks=world, j_begin=jb, j_end=je, **kw), jb - 1)
    slabs.connect(h, dist)
    counts = []
    for it in range(nsteps):
        b, vel = bodies_at(nb, 40 * it, dt)          # 40 steps' worth of motion per step: the cell types change
        h.set_bodies(b, vel)
        st = h.step()
        counts.append((st.ad_iters, st.ppe_sweeps))
    mine = {k: row_digests(h.get(k), inp.nx, range(1, je - jb + 1)) for k in ("u", "v", "p")}
    gathered = [None] * world
    dist.all_gather_object(gathered, (counts, mine))
    dist.barrier()
    h.close()
    out = None
    if rank == 0:
        r = run(ifx.ImmerseFlow(inp, xf, yf, device=dev, **kw), 0)
        rc = []
        for it in range(nsteps):
            b, vel = bodies_at(nb, 40 * it, dt)
            r.set_bodies(b, vel)
            st = r.step()
            rc.append((st.ad_iters, st.ppe_sweeps))
        ref = {k: row_digests(r.get(k), inp.nx, range(1, inp.ny - 1)) for k in ("u", "v", "p")}
        ghost = int(r.lib.ifx_ghost_cell_count(r._h))
        r.close()
        bad = 0
        for q, (cq, dq) in enumerate(gathered):
            qb, qe = slabs.partition_rows(inp.ny, world)[q]
            for k in ("u", "v", "p"):
                bad += sum(1 for a_, b_ in zip(dq[k], ref[k][qb - 1:qe - 1]) if a_ != b_)
            if cq != rc:
                bad += 1
        out = {"bit_exact": bad == 0, "grid": [ncx, ncy], "steps": nsteps, "ranks": world, "bodies": nb, "ghost_cells": ghost,
               "iteration_counts": rc, "mismatching_rows": bad,
               "what": "u, v, p of every owned row (digests) and the iteration counts of the slab run == the single-GPU run"}
    return out


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    nb = args.bodies if args.mode == "full" else 0
    body_txt = (f"{nb} moving lobed bodies (256 markers each), iBlank + ghost cells recomputed every step" if nb
                else "no immersed body")
    config = {"mode": args.mode,
              "workload": f"vortex IC on uniform {args.nx}x{args.ny} cells, reference BCs (u=1,v=0), {body_txt}, "
                          f"dt={args.dt}, Re={args.Re}, AD_itermax={args.ad_itermax}, {args.ppe_sweeps} Poisson sweeps/step from a smooth "
                          f"non-zero pressure field",
              "grid": [args.nx, args.ny],
              "l2": "working set %.1f GB of fields vs 126 MB L2 (%s)" % (
                  8 * 8.0 * (args.nx + 2) * (args.ny + 2) / 1e9,
                  "inputs larger than L2, no flush needed" if 8 * 8.0 * (args.nx + 2) * (args.ny + 2) > 4 * 126e6
                  else "NOT larger than L2: an HBM fraction measured on this grid is an L2 figure"),
              "decomposition": f"{world} row slab(s), halo rows by in-kernel NVLink P2P stores, residual sums by P2P mailboxes "
                               f"(stop decision one sweep behind)" + (f", slab heights balanced by {args.slab_balance}" if world > 1 else "")}

    if args.impl == "reference":
        if rank != 0:
            return
        steps, warmup = max(1, args.steps), max(0, args.warmup)
        cb = cpu_reference_arm(args, args.cpu_sample_rows, steps, warmup)
        line = {"metric": METRIC, "value": cb["value"], "unit": METRIC, "n_gpus": args.gpus, "steps": steps,
                "warmup": warmup, "ms_per_step": cb["ms_per_step"], "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic", "impl": "reference", "config": config,
                "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample", "sample_fraction")},
                "e2e": {"value": cb["value"], "unit": METRIC, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return

    import torch
    import immerseflow_b200 as ifx

    dist = None
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG", "WARN")          # keeps NCCL's version banner off stdout (one JSON line)
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl")
    dev = local_rank if world > 1 else 0
    torch.cuda.set_device(dev)

    ncx, ncy = args.nx, args.ny
    inp = ifx.make_input(ncx, ncy, args.dt, args.Re, AD_itermax=args.ad_itermax, PPE_itermax=args.ppe_sweeps)
    xf, yf = ifx.uniform_faces(ncx, 1.0), ifx.uniform_faces(ncy, 1.0)
    from immerseflow_b200 import slabs
    part = slabs.partition_rows(inp.ny, world)
    if world > 1 and nb > 0 and args.slab_balance == "cost":
        part = slabs.partition_rows_weighted(inp.ny, world, row_costs(nb, ncx, ncy, args.dt))
    jb, je = part[rank]
    if args.emulate_slab_of and world == 1:
        jb, je = slabs.partition_rows(inp.ny, args.emulate_slab_of)[0]
    full = args.mode == "full"

    slab_parity = None
    if world > 1 and not args.no_parity_check:
        slab_parity = slab_parity_check(ifx, slabs, dist, torch, rank, world, dev)

    p0 = []

    def make_solver(zero_copy_control=0):
        """one simulation: handle, initial state, slab connections"""
        h = ifx.ImmerseFlow(inp, xf, yf, device=dev, sweeps_per_batch=args.ppe_sweeps + 1, rank=rank, nranks=world,
                            j_begin=jb, j_end=je, compat=ifx.IFX_COMPAT_FULL if full else ifx.IFX_COMPAT_REFERENCE,
                            ppe_abs_residual=1 if full else 0, zero_copy_control=zero_copy_control,
                            ppe_pairs=args.ppe_pairs if world == 1 else 0)
        h.initializeData()
        if not p0:                         # computed once, shared by the handles of the end-to-end pipeline
            p0.append(initial_pressure(inp.nx, inp.ny, jb - 1, h.field_size("p") // inp.nx))
        h.set("p", p0[0])
        if world > 1:
            slabs.connect(h, dist)
        return h

    s = make_solver()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    step_no = [0]

    def one_step(h=None):
        h = h or s
        if full:
            if nb:
                b, vel = bodies_at(nb, step_no[0], args.dt)
                h.set_bodies(b, vel)
                step_no[0] += 1
            st = h.step()
            return st, st
        a = h.ADsolver()
        b = h.PPESolver()
        return a, b

    # ---- parity of THIS workload, before anything is timed: the first step from the initial condition, the rows well
    # inside the CPU arm's row slab (the oracle runs that slab with walls where the grid goes on), bit for bit
    parity_gpu = None
    do_parity = not args.no_parity_check and not args.emulate_slab_of
    warm_done = 0
    if do_parity:
        r0, nrows_cpu = cpu_slab_rows(args, args.cpu_sample_rows)
        lo, hi = r0 + PARITY_MARGIN + 1, r0 + nrows_cpu - PARITY_MARGIN + 1       # ghost-inclusive global rows [lo, hi)
        # the oracle starts from the SAME initial velocity: the slab's rows (its two ghost rows included) of u, v as the
        # vortex kernel left them, collected on rank 0
        s_lo, s_hi = r0, r0 + nrows_cpu + 2                                       # ghost-inclusive global rows of the slab
        o_lo = jb if rank > 0 else 0                                               # rows this rank answers for
        o_hi = je if rank < world - 1 else inp.ny
        c_lo, c_hi = max(s_lo, o_lo), min(s_hi, o_hi)
        ic_part = None
        if c_hi > c_lo:
            ic_part = {"rows": (c_lo, c_hi)}
            for k in ("u", "v"):
                f = s.get(k).reshape(-1, inp.nx)
                ic_part[k] = f[c_lo - (jb - 1): c_hi - (jb - 1)].copy()
                del f
        if world > 1:
            ic_all = [None] * world
            dist.gather_object(ic_part, ic_all if rank == 0 else None, dst=0)
        else:
            ic_all = [ic_part]
        parity_ic = None
        if rank == 0:
            parity_ic = {k: np.empty((s_hi - s_lo, inp.nx)) for k in ("u", "v")}
            for part in ic_all:
                if part:
                    a_, b_ = part["rows"]
                    for k in ("u", "v"):
                        parity_ic[k][a_ - s_lo: b_ - s_lo] = part[k]
        del ic_part, ic_all
        a0, b0 = one_step()
        warm_done = 1
        mine = {}
        rows = [j for j in range(max(lo, jb), min(hi, je))]
        for k in ("u", "v", "p"):
            f = s.get(k)
            mine[k] = dict(zip(rows, row_digests(f, inp.nx, [j - (jb - 1) for j in rows]))) if rows else {}
            del f
        mine["counts"] = (a0.ad_iters, b0.ppe_sweeps)
        if world > 1:
            allm = [None] * world
            dist.gather_object(mine, allm if rank == 0 else None, dst=0)
        else:
            allm = [mine]
        if rank == 0:
            parity_gpu = {k: {} for k in ("u", "v", "p")}
            for m in allm:
                for k in ("u", "v", "p"):
                    parity_gpu[k].update(m[k])
            parity_gpu["counts"] = [m["counts"] for m in allm]
            parity_gpu["rows"] = (lo, hi, r0)

    for _ in range(max(0, args.warmup - warm_done)):
        one_step()
    sampler = ClockSampler(dev)
    barrier()
    sampler.start()
    l0 = s.launch_count
    t0 = time.perf_counter()
    ad_ms, ppe_ms, cor_ms, ib_ms, sweep_ms, k_ad, k_ppe = [], [], [], [], [], 0, 0
    for _ in range(args.steps):
        a, b = one_step()
        ad_ms.append(a.ms_ad); ppe_ms.append(b.ms_ppe); cor_ms.append(b.ms_correct); ib_ms.append(b.ms_ib)
        sweep_ms.append(a.ms_ad_sweeps)
        k_ad, k_ppe = a.ad_iters, b.ppe_sweeps
    barrier()
    wall = time.perf_counter() - t0
    launches = s.launch_count - l0
    clocks = sampler.stop()
    # device time of the stages = CUDA-event stage timers (the source-term / face kernels between the stages are not inside)
    dev_ms = float(np.mean(ad_ms) + np.mean(ppe_ms) + np.mean(cor_ms) + np.mean(ib_ms))
    if world > 1:
        t = torch.tensor([dev_ms, wall * 1e3 / args.steps], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_ms, wall_ms = t.tolist()
    else:
        wall_ms = wall * 1e3 / args.steps
    cells = ncx * ncy if not args.emulate_slab_of else ncx * (je - jb)
    value = cells / (wall_ms * 1e-3) / 1e6

    # ---- roofline of the dominant kernel BY TIME: the Poisson sweep (k_ppe + 1 launches of 25 B/cell per step against
    # k_ad predictor launches of 49 B/cell).  achieved = algorithmic bytes / CUDA-event time per launch (events around
    # the Poisson stage / around the predictor's sweep launches, ifx_step_stats).  Both fractions: against the copy
    # bandwidth measured on this pool's B200s (MEASURED_PEAKS.json) and against north_star's nominal 8 TB/s.
    peak, peak_src = peaks()
    cells_local = ncx * (je - jb)
    ad_launch_ms = float(np.mean(sweep_ms)) / k_ad if k_ad else None
    jac_bytes = 49.0 * cells_local               # read u,v,sx,sy + write u',v' (48 B/cell) + 1 B cell type
    ppe_launch_ms = float(np.mean(ppe_ms)) / (k_ppe + 1)
    ppe_b_cell = 25 if full else 16              # full: read p, rhs, face mask; write p' — reference Laplace: read p, write p'
    ppe_bytes = float(ppe_b_cell) * cells_local
    ach_ad = jac_bytes / (ad_launch_ms * 1e-3) / 1e9 if ad_launch_ms else 0.0
    ach_ppe = ppe_bytes / (ppe_launch_ms * 1e-3) / 1e9
    ppe_share = float(np.mean(ppe_ms)) / wall_ms
    ad_share = float(np.mean(sweep_ms)) / wall_ms
    roofline = {"bound": "hbm", "kernel": "k_sweep_v4<poisson %s>" % ("general" if full else "laplace"),
                "achieved": ach_ppe, "peak": peak, "unit": "GB/s", "frac": ach_ppe / peak,
                "frac_nominal_8tbs": ach_ppe / 8000.0,
                "traffic": measured_traffic("k_sweep_v4_ppe_general" if full else "k_sweep_v4_ppe_laplace", ncx, ncy) if world == 1 else None,
                "peak_source": peak_src, "algorithmic_bytes_per_cell": ppe_b_cell, "ms_per_launch": ppe_launch_ms,
                "launches_per_step": k_ppe + 1, "share_of_step": ppe_share, "sweeps_per_s": 1e3 / ppe_launch_ms,
                "predictor": {"kernel": "k_sweep_v4<predictor Jacobi>", "achieved": ach_ad, "frac": ach_ad / peak,
                              "frac_nominal_8tbs": ach_ad / 8000.0,
                              "traffic": measured_traffic("k_sweep_v4_ad", ncx, ncy) if world == 1 else None,
                              "algorithmic_bytes_per_cell": 49, "ms_per_launch": ad_launch_ms, "launches_per_step": k_ad,
                              "share_of_step": ad_share},
                "projection_ms": float(np.mean(cor_ms)), "iblank_ghost_cells_ms": float(np.mean(ib_ms)),
                "ghost_cells": int(s.lib.ifx_ghost_cell_count(s._h))}
    if full and world == 1 and roofline["traffic"] is not None:
        roofline["traffic_note"] = ("ncu capture of the kernel's two-columns-per-thread geometry (256-column tiles, what slabs run); a "
                                    "single GPU runs the four-column geometry since the end of round 2: 516 instead of 2 x 260 "
                                    "columns loaded per 512 cells of a row, same rhs / mask / store bytes — not more traffic")

    # ---- e2e: the same step through the C-ABI with HOST buffers (pinned): every step uploads its u, v, p and downloads
    # the u, v, p it produced.  One handle does that in series (H2D, kernels, D2H: the PCIe time of 2 x 3 fields is
    # longer than the step).  The default keeps `--e2e-handles` simulations in flight, each on its own stream: while
    # one steps, the next one's state is uploading and the previous one's result is downloading, so both PCIe
    # directions overlap the kernels; pipeline fill and drain are inside the timed region.
    e2e = None
    if not args.no_e2e:
        n = s.field_size("u")
        names = ("u", "v", "p")
        host_in = {k: torch.empty(n, dtype=torch.float64, pin_memory=True).numpy() for k in names}
        for k in names:
            s.get(k, host_in[k])
        nh = max(1, min(args.e2e_handles, 3))
        reps = max(1, args.steps)
        sims, host_out, fallback = [], None, None
        if nh > 1:
            # handles whose control traffic does not queue behind the bulk transfers (ifx_options.zero_copy_control).
            # If the box cannot hold them (device memory, page-locked host memory), every rank falls back to one handle.
            try:
                sims = [make_solver(1) for _ in range(nh)]
                host_out = [{k: torch.empty(n, dtype=torch.float64, pin_memory=True).numpy() for k in names} for _ in range(2)]
            except Exception as ex:       # noqa: BLE001 — reported in the JSON line
                fallback = f"{type(ex).__name__}: {ex}"[:200]
            ok = torch.tensor([0.0 if fallback else 1.0], device="cuda")
            if world > 1:
                dist.all_reduce(ok, op=dist.ReduceOp.MIN)
            if ok.item() < 1.0:
                fallback = fallback or "another rank could not set up its handles"
                for h in sims:
                    h.close()
                sims, host_out, nh = [], None, 1
        if nh == 1:
            barrier()
            t0 = time.perf_counter()
            for _ in range(reps):
                for k in names:
                    s.set(k, host_in[k])
                one_step()
                for k in names:
                    s.get(k, host_in[k])
            barrier()
        else:
            for h in sims:                         # every handle allocates its lazily sized buffers before the clock starts
                one_step(h)
            barrier()
            for h in sims:
                h.synchronize()
            t0 = time.perf_counter()
            for k in names:                        # fill: the first simulation's state
                sims[0].set_async(k, host_in[k])
            for it in range(reps):
                cur, nxt, prv = sims[it % nh], sims[(it + 1) % nh], sims[(it - 1) % nh]
                cur.synchronize()                  # its upload has landed (and its older download left host_out[it % 2] free)
                if it >= 1:                        # (download before upload: with 2 handles they share a stream)
                    for k in names:
                        prv.get_async(k, host_out[it % 2][k])
                if it + 1 < reps:
                    for k in names:
                        nxt.set_async(k, host_in[k])
                one_step(cur)
            for k in names:                        # drain: the last result
                sims[(reps - 1) % nh].get_async(k, host_out[reps % 2][k])
            for h in sims:
                h.synchronize()
            barrier()
        te = (time.perf_counter() - t0) / reps
        if world > 1:
            t = torch.tensor([te], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            te = float(t.item())
        e2e = {"value": cells / te / 1e6, "unit": METRIC, "h2d_bytes_per_step": 3 * 8 * n, "d2h_bytes_per_step": 3 * 8 * n,
               "ms_per_step": te * 1e3, "steps": reps,
               "pipeline": (f"{nh} simulations in flight (upload | step | download overlapped), fill + drain timed" if nh > 1
                            else "upload, step, download in series" + (f" (fallback: {fallback})" if fallback else ""))}
        if nh > 1:
            if world > 1:
                dist.barrier()                     # nobody frees a segment a neighbour may still be storing into
            for h in sims:
                h.close()
        del host_in, host_out

    line = {"metric": METRIC, "value": value, "unit": METRIC, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": wall_ms, "device_ms_per_step": dev_ms, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
            "iters": {"ad": k_ad, "ppe_sweeps": k_ppe}, "roofline": roofline, "gpu_launches": launches,
            "clocks": clocks, "e2e": e2e}
    if slab_parity is not None or world > 1:
        line["slab_parity"] = slab_parity["bit_exact"] if slab_parity else None
        line["slab_parity_detail"] = slab_parity
    s.close()
    if world > 1:
        dist.barrier()

    # ---- rank 0: the CPU arm (N = 1: timed as `cpu_baseline`; its first step doubles as the parity oracle; N > 1: the
    # first step only) and the reference's own CUDA binary (N = 1)
    if rank == 0:
        want_cb = world == 1 and not args.no_cpu_baseline
        cb = None
        if want_cb or parity_gpu is not None:
            cb = cpu_reference_arm(args, args.cpu_sample_rows, 2 if want_cb else 0, 1, keep_first_step=parity_gpu is not None,
                                   ic=parity_ic if parity_gpu is not None else None)
        if want_cb:
            line["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample", "sample_fraction")}
        if parity_gpu is not None:
            lo, hi, r0 = parity_gpu["rows"]
            gnx, gny = cb["grid"]
            bad, checked = {}, 0
            for k in ("u", "v", "p"):
                want = dict(zip(range(lo, hi), row_digests(cb["first_step"][k], gnx, [j - r0 for j in range(lo, hi)])))
                bad[k] = sum(1 for j in range(lo, hi) if parity_gpu[k].get(j) != want[j])
                checked += hi - lo
            k_or = (args.ad_itermax, args.ppe_sweeps)
            counts_ok = all(tuple(c) == k_or for c in parity_gpu["counts"])
            line["parity_check"] = {
                "bit_exact": all(v == 0 for v in bad.values()) and counts_ok,
                "rows": hi - lo, "global_rows": [lo, hi], "cells": (hi - lo) * ncx, "fields": ["u", "v", "p"],
                "mismatching_rows": bad, "iteration_counts_equal": counts_ok,
                "what": f"first step of the timed workload (all {world} rank(s)) vs the CPU oracle run on cell rows {r0}..{r0 + cb['grid'][1] - 2}: "
                        f"rows at least {PARITY_MARGIN} cells inside the oracle's slab (its artificial walls cannot reach them in "
                        f"one step), every interior column, compared as per-row digests"}
        if world == 1 and not args.no_ref_cuda and not args.emulate_slab_of:
            rcu = ref_cuda_baseline(args)
            if "unavailable" not in rcu and args.nx <= args.ny:
                # the same work on our path: predictor only, no body, reference-compatible arithmetic (bit-identical results,
                # tests/test_gpu_reference_parity.py)
                inp_r = ifx.make_input(ncx, ncy, args.dt, args.Re, AD_itermax=args.ad_itermax, PPE_itermax=1)
                with ifx.ImmerseFlow(inp_r, xf, yf, device=dev, compat=ifx.IFX_COMPAT_REFERENCE) as hr:
                    hr.initializeData()
                    hr.ADsolver()
                    torch.cuda.synchronize()
                    t0 = time.perf_counter()
                    for _ in range(3):
                        hr.ADsolver()
                    torch.cuda.synchronize()
                    tr = (time.perf_counter() - t0) / 3
                rcu["ours_same_work"] = {"ms_per_step": tr * 1e3, "value": ncx * ncy / tr / 1e6, "unit": METRIC,
                                         "what": "ifx_ad_solve in IFX_COMPAT_REFERENCE (predictor only, same bits)"}
            line["ref_cuda_baseline"] = rcu
        if world == 1 and not args.no_secondary and not args.emulate_slab_of and full:
            line["configs"] = secondary_configs(ifx, torch, dev, args)
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
