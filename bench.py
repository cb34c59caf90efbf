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
    ap.add_argument("--cpu-sample-rows", type=int, default=192, help="rows of the workload the CPU baseline runs on")
    ap.add_argument("--no-cpu-baseline", action="store_true")
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


# ------------------------------------------------------------------------------------------------
def cpu_reference_arm(args, rows: int, steps: int, warmup: int):
    """The reference has no CPU path; its reported CPU baseline is the OpenMP transcription of its loops
    (oracle/, `kind: port`), run with every host thread on a bounded row-slab of the same workload (same grid
    width, same parameters, same iteration caps, `rows` of the 16384 rows)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import _oracle as orc
    import immerseflow_b200 as ifx
    ncx, ncy = args.nx, min(rows, args.ny)
    xf = ifx.uniform_faces(ncx, 1.0)
    r0 = 0
    if args.bodies > 0 and args.mode == "full":      # the slab that cuts through the first lattice row of bodies
        cols = int(np.ceil(np.sqrt(args.bodies)))
        nrow = (args.bodies + cols - 1) // cols
        r0 = max(0, min(args.ny - ncy, int(0.5 / nrow * args.ny) - ncy // 2))
    yf = ifx.uniform_faces(args.ny, 1.0)[r0: r0 + ncy + 1]
    cores = orc.lib().orc_num_threads()
    times, k_ad, k_ppe = [], 0, 0
    if args.mode == "full":
        fs = orc.FullSolver(xf, yf, args.dt, args.Re, args.ad_itermax, args.ppe_sweeps, ppe_tol=0.0, ppe_abs=1)
        g = orc.Grid(xf, yf)
        u, v, p = orc.initial_condition(g)
        fs.set("u", u); fs.set("v", v)
        jj, ii = np.divmod(np.arange(g.nx * g.ny, dtype=np.float64), float(g.nx))
        fs.set("p", 50.0 + 40.0 * np.sin(ii * (6.283185307179586 / g.nx)) * np.cos(jj * (6.283185307179586 / (args.ny + 2))))
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
        fs.close()
    else:
        g = orc.Grid(xf, yf)
        u, v, p = orc.initial_condition(g)
        pr = orc.Predictor(g, u, v, args.dt, args.Re, args.ad_itermax)
        for it in range(warmup + steps):
            t0 = time.perf_counter()
            k_ad, _ = pr.step()
            k_ppe, p, _ = orc.ppe_solve(g, p, args.ppe_sweeps)
            if it >= warmup:
                times.append(time.perf_counter() - t0)
    cells = ncx * ncy
    t = float(np.mean(times))
    return {"value": cells / t / 1e6, "unit": METRIC, "cores": cores, "kind": "port",
            "sample": f"{ncx}x{ncy}-cell row slab (rows {r0}..{r0 + ncy}) of the workload ({args.mode} step), {steps} step(s) "
                      f"(K_AD={k_ad}, {k_ppe} Poisson sweeps), OpenMP x{cores}", "ms_per_step": t * 1e3}


def measured_traffic(kernel: str, nx: int, ny: int):
    """DRAM bytes per launch of the dominant kernel from the committed ncu capture (profiles/r1_traffic.json), valid
    for the grid it was captured on; None otherwise."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "r1_traffic.json")))
        e = t.get(kernel)
        if e and e["grid"] == [nx, ny]:
            return e["dram_bytes_per_launch"]
    except Exception:
        pass
    return None


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
              "decomposition": f"{world} row slab(s), halo rows by in-kernel NVLink P2P stores, residual by P2P mailboxes"}

    if args.impl == "reference":
        if rank != 0:
            return
        steps = max(1, min(args.steps, 3))
        cb = cpu_reference_arm(args, max(args.cpu_sample_rows, 768), steps, min(args.warmup, 1))
        line = {"metric": METRIC, "value": cb["value"], "unit": METRIC, "n_gpus": args.gpus, "steps": steps,
                "warmup": min(args.warmup, 1), "ms_per_step": cb["ms_per_step"], "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic", "impl": "reference", "config": config,
                "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")},
                "e2e": {"value": cb["value"], "unit": METRIC, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return

    import torch
    import immerseflow_b200 as ifx

    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl")
    dev = local_rank if world > 1 else 0
    torch.cuda.set_device(dev)

    ncx, ncy = args.nx, args.ny
    inp = ifx.make_input(ncx, ncy, args.dt, args.Re, AD_itermax=args.ad_itermax, PPE_itermax=args.ppe_sweeps)
    xf, yf = ifx.uniform_faces(ncx, 1.0), ifx.uniform_faces(ncy, 1.0)
    from immerseflow_b200 import slabs
    jb, je = slabs.partition_rows(inp.ny, world)[rank]
    if args.emulate_slab_of and world == 1:
        jb, je = slabs.partition_rows(inp.ny, args.emulate_slab_of)[0]
    full = args.mode == "full"

    p0 = []

    def make_solver(zero_copy_control=0):
        """one simulation: handle, initial state, slab connections"""
        h = ifx.ImmerseFlow(inp, xf, yf, device=dev, sweeps_per_batch=args.ppe_sweeps + 1, rank=rank, nranks=world,
                            j_begin=jb, j_end=je, compat=ifx.IFX_COMPAT_FULL if full else ifx.IFX_COMPAT_REFERENCE,
                            ppe_abs_residual=1 if full else 0, zero_copy_control=zero_copy_control)
        h.initializeData()
        # The reference starts the Poisson solve from p == 0 (preSim.cu:67), which makes most quotients exact zeros for
        # the first sweeps — an arithmetic special case.  The bench uses a smooth non-zero pressure field so that every
        # fp64 division takes the normal-operand path a converging solve sees.
        if not p0:                         # computed once, shared by the handles of the end-to-end pipeline
            n_p = h.field_size("p")
            jj, ii = np.divmod(np.arange(n_p, dtype=np.float64), float(inp.nx))
            jj += jb - 1
            p0.append(50.0 + 40.0 * np.sin(ii * (6.283185307179586 / inp.nx)) * np.cos(jj * (6.283185307179586 / inp.ny)))
            del jj, ii
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

    for _ in range(args.warmup):
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
    # device time of the step = CUDA-event stage timers (both stages run on the solver's stream back to back)
    dev_ms = float(np.mean(ad_ms) + np.mean(ppe_ms) + np.mean(cor_ms) + np.mean(ib_ms))
    if world > 1:
        t = torch.tensor([dev_ms, wall * 1e3 / args.steps], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_ms, wall_ms = t.tolist()
    else:
        wall_ms = wall * 1e3 / args.steps
    cells = ncx * ncy if not args.emulate_slab_of else ncx * (je - jb)
    value = cells / (wall_ms * 1e-3) / 1e6

    # ---- roofline of the dominant kernel (k_ad_jacobi: 25 launches/step vs ~51 Poisson launches of 1/3 the bytes)
    peak, peak_src = peaks()
    cells_local = ncx * (je - jb)
    # CUDA events around the k_ad sweep launches alone (ifx_step_stats.ms_ad_sweeps; with bodies the per-iteration
    # ghost-cell kernels, a few microseconds each, are inside)
    ad_launch_ms = float(np.mean(sweep_ms)) / k_ad if k_ad else None
    jac_bytes = 49.0 * cells_local               # read u,v,sx,sy + write u',v' (48 B/cell) + 1 B cell type
    ppe_launch_ms = float(np.mean(ppe_ms)) / (k_ppe + 1)
    ppe_b_cell = 25 if full else 16              # full: read p, rhs, cell type; write p' — reference Laplace: read p, write p'
    ppe_bytes = float(ppe_b_cell) * cells_local
    ach_ad = jac_bytes / (ad_launch_ms * 1e-3) / 1e9 if ad_launch_ms else 0.0
    ach_ppe = ppe_bytes / (ppe_launch_ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "kernel": "k_sweep_v4<predictor Jacobi>", "achieved": ach_ad, "peak": peak, "unit": "GB/s",
                "frac": ach_ad / peak, "traffic": measured_traffic("k_sweep_v4_ad", ncx, ncy) if world == 1 else None,
                "peak_source": peak_src,
                "algorithmic_bytes_per_cell": 49, "ms_per_launch": ad_launch_ms,
                "poisson": {"kernel": "k_sweep_v4<poisson %s>" % ("general" if full else "laplace"), "achieved": ach_ppe,
                            "traffic": measured_traffic("k_sweep_v4_ppe_general" if full else "k_sweep_v4_ppe_laplace", ncx, ncy)
                            if world == 1 else None,
                            "frac": ach_ppe / peak, "algorithmic_bytes_per_cell": ppe_b_cell, "ms_per_launch": ppe_launch_ms,
                            "sweeps_per_s": 1e3 / ppe_launch_ms},
                "projection_ms": float(np.mean(cor_ms)), "iblank_ghost_cells_ms": float(np.mean(ib_ms)),
                "ghost_cells": int(s.lib.ifx_ghost_cell_count(s._h))}

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

    line = {"metric": METRIC, "value": value, "unit": METRIC, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": wall_ms, "device_ms_per_step": dev_ms, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
            "iters": {"ad": k_ad, "ppe_sweeps": k_ppe}, "roofline": roofline, "gpu_launches": launches,
            "clocks": clocks, "e2e": e2e}
    if rank == 0 and not args.no_cpu_baseline:
        cb = cpu_reference_arm(args, args.cpu_sample_rows, 1, 1)
        line["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}
    s.close()
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
