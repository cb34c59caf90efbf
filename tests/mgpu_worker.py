"""torchrun worker: slab-decomposed run on WORLD_SIZE GPUs vs the single-GPU run, bit for bit.
Launched by tests/test_gpu_slabs.py (needs >= 2 GPUs) — not collected by pytest itself."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch
import torch.distributed as dist

import immerseflow_b200 as ifx
from immerseflow_b200 import slabs


def run_case(compat, ncx, ncy, steps, ppe_it, bc=None, stretched=True):
    import _oracle as orc
    rank, world = dist.get_rank(), dist.get_world_size()
    xf = orc.stretched_faces(ncx, 2.0, 1.02) if stretched else ifx.uniform_faces(ncx, 1.0)
    yf = orc.stretched_faces(ncy, 1.5, 1.02) if stretched else ifx.uniform_faces(ncy, 1.0)
    inp = ifx.make_input(ncx, ncy, 5e-4, 100.0, AD_itermax=12, PPE_itermax=ppe_it)
    nx, ny = inp.nx, inp.ny
    rng = np.random.default_rng(1234)
    u0 = 1.0 + 0.1 * rng.standard_normal(nx * ny)
    v0 = 0.1 * rng.standard_normal(nx * ny)
    p0 = rng.standard_normal(nx * ny)
    jb, je = slabs.partition_rows(ny, world)[rank]
    kw = dict(compat=compat, bc=bc, ppe_abs_residual=1 if compat == ifx.IFX_COMPAT_FULL else 0)
    s = ifx.ImmerseFlow(inp, xf, yf, device=torch.cuda.current_device(), rank=rank, nranks=world, j_begin=jb, j_end=je,
                        sweeps_per_batch=16, **kw)
    s.initializeData()
    for name, f in (("u", u0), ("v", v0), ("p", p0)):
        s.set(name, slabs.scatter_rows(f, nx, ny, world, rank))
    slabs.connect(s, dist)
    counts = []
    for _ in range(steps):
        if compat == ifx.IFX_COMPAT_FULL:
            st = s.step()
        else:
            st = s.ADsolver(); st2 = s.PPESolver(); st.ppe_sweeps = st2.ppe_sweeps
        counts.append((st.ad_iters, st.ppe_sweeps))
    parts = {}
    for name in ("u", "v", "p"):
        mine = s.get(name)
        gathered = [None] * world
        dist.all_gather_object(gathered, mine)
        parts[name] = gathered
    dist.barrier()
    s.close()
    ok = True
    if rank == 0:
        ref = ifx.ImmerseFlow(inp, xf, yf, device=torch.cuda.current_device(), sweeps_per_batch=16, **kw)
        ref.initializeData()
        ref.set("u", u0); ref.set("v", v0); ref.set("p", p0)
        rc = []
        for _ in range(steps):
            if compat == ifx.IFX_COMPAT_FULL:
                st = ref.step()
            else:
                st = ref.ADsolver(); st2 = ref.PPESolver(); st.ppe_sweeps = st2.ppe_sweeps
            rc.append((st.ad_iters, st.ppe_sweeps))
        m = np.ones((ny, nx), bool)
        m[0, 0] = m[0, -1] = m[-1, 0] = m[-1, -1] = False
        # ring columns of slab-interior rows are only refreshed lazily in reference mode: compare the cells that matter
        inner = np.zeros((ny, nx), bool); inner[1:-1, 1:-1] = True
        for name in ("u", "v", "p"):
            got = slabs.assemble_rows(parts[name], nx, ny)
            want = ref.get(name)
            same = np.array_equal(got[inner.reshape(-1)], want[inner.reshape(-1)])
            print(f"[{compat}] {ncx}x{ncy} world={world} {name}: interior bit-identical = {same}; counts {counts} vs {rc}")
            ok = ok and same
        ok = ok and counts == rc
        ref.close()
    return ok


def run_bodies_case(ncx, ncy, steps, moving, ppe_solver=1, ppe_omega=1.0, scale=1.0, ad_tol=None, ppe_tol=None):
    """Immersed bodies straddling the slab boundaries (configs 3-5 of BASELINE.json in miniature): cell types,
    ghost-cell maps and weights, u, v, p and the iteration counts of the slab run against the single-GPU run AND the
    CPU oracle, bit for bit."""
    import _oracle as orc
    rank, world = dist.get_rank(), dist.get_world_size()
    xf, yf = orc.stretched_faces(ncx, 4.0, 1.015), orc.stretched_faces(ncy, 2.0, 1.015)
    dt, Re, ad_it, ppe_it = 2e-3, 200.0, 15, 60
    inp = ifx.make_input(ncx, ncy, dt, Re, AD_itermax=ad_it, PPE_itermax=ppe_it)
    nx, ny = inp.nx, inp.ny
    # scale < 1: a slow stream, so that the un-normalised residual sums fall below the tolerances MID-LOOP (the lagged stop
    # decision of the slab runs then has converged iterates to preserve: one extra sweep has run when it fires)
    kw = dict(compat=ifx.IFX_COMPAT_FULL, ppe_abs_residual=1, ppe_solver=ppe_solver, ppe_omega=ppe_omega)
    okw = {}
    if scale != 1.0:
        kw["bc"] = dict(u_bc_w=scale, u_bc_e=scale, u_bc_s=scale, u_bc_n=scale)
        okw["bc_u"] = (scale,) * 4
    if ad_tol is not None:
        kw["ad_tol"] = ad_tol; okw["ad_tol"] = ad_tol
    if ppe_tol is not None:
        kw["ppe_tol"] = ppe_tol; okw["ppe_tol"] = ppe_tol

    def bodies_at(step):
        sh = 0.013 * step if moving else 0.021      # off the grid's symmetry line: stencils cross the slab boundary
        return ([orc.circle_markers(1.2 + sh, 1.0 - sh, 0.27, 48), orc.ellipse_markers(2.3, 0.55 + sh, 0.3, 0.14, 0.3 + sh, 64),
                 orc.ellipse_markers(3.1 - sh, 1.45, 0.16, 0.33, -0.4, 56)],
                [(6.5 * scale, -6.5 * scale), (0.0, 6.5 * scale), (-6.5 * scale, 0.0)] if moving else None)

    jb, je = slabs.partition_rows(ny, world)[rank]
    s = ifx.ImmerseFlow(inp, xf, yf, device=torch.cuda.current_device(), rank=rank, nranks=world, j_begin=jb, j_end=je,
                        sweeps_per_batch=16, **kw)
    s.initializeData()
    u0, v0, p0 = np.full(nx * ny, scale), np.zeros(nx * ny), np.zeros(nx * ny)
    for name, f in (("u", u0), ("v", v0), ("p", p0)):
        s.set(name, slabs.scatter_rows(f, nx, ny, world, rank))
    slabs.connect(s, dist)
    hist = []
    for step in range(steps):
        if moving or step == 0:
            b, vel = bodies_at(step)
            s.set_bodies(b, vel)
        st = s.step()
        snap = {"counts": (st.ad_iters, st.ppe_sweeps), "gc": s.ghost_cells()}
        for name in ("u", "v", "p", "celltype"):
            snap[name] = s.get(name)
        gathered = [None] * world
        dist.all_gather_object(gathered, snap)
        hist.append(gathered)
    dist.barrier()
    s.close()
    ok = True
    if rank == 0:
        ref = ifx.ImmerseFlow(inp, xf, yf, device=torch.cuda.current_device(), sweeps_per_batch=16, **kw)
        ref.initializeData()
        ref.set("u", u0); ref.set("v", v0); ref.set("p", p0)
        o = orc.FullSolver(xf, yf, dt, Re, ad_it, ppe_it, ppe_abs=1, **okw)
        o.set_ppe_solver(ppe_solver, ppe_omega)
        o.set("u", u0); o.set("v", v0)
        inner = np.zeros((ny, nx), bool); inner[1:-1, 1:-1] = True
        inner = inner.reshape(-1)
        for step in range(steps):
            if moving or step == 0:
                b, vel = bodies_at(step)
                ref.set_bodies(b, vel); o.set_bodies(b, vel); o.update_ib()
            st = ref.step(); so = o.step()
            parts = hist[step]
            msgs = []
            counts = parts[0]["counts"]
            if not all(q["counts"] == counts for q in parts): msgs.append(f"ranks disagree on counts {[q['counts'] for q in parts]}")
            if counts != (st.ad_iters, st.ppe_sweeps): msgs.append(f"counts {counts} vs single GPU {(st.ad_iters, st.ppe_sweeps)}")
            if counts != (int(so[0]), int(so[3])): msgs.append(f"counts {counts} vs oracle {(int(so[0]), int(so[3]))}")
            gc = {k: np.concatenate([q["gc"][k] for q in parts]) for k in ("cell", "stencil", "weights", "bi", "ip")}
            g1, go = ref.ghost_cells(), o.ghost_cells()
            for k in gc:
                if not np.array_equal(gc[k], g1[k]): msgs.append(f"ghost-cell {k} differs from single GPU")
                if not np.array_equal(gc[k], go[k]): msgs.append(f"ghost-cell {k} differs from oracle")
            for name in ("celltype", "u", "v", "p"):
                got = slabs.assemble_rows([q[name] for q in parts], nx, ny)
                for tag, want in (("single GPU", ref.get(name)), ("oracle", o.get(name))):
                    if not np.array_equal(got[inner], want[inner]):
                        d = np.flatnonzero(got[inner] != want[inner])
                        msgs.append(f"{name} differs from {tag} in {d.size} cells (first inner index {d[0]})")
            want_ct = ref.get("celltype").reshape(ny, nx)
            for r, q in enumerate(parts):          # halo rows carry the owner's cell types, ghost bit included
                lo, hi = slabs.local_rows(ny, world, r)
                if not np.array_equal(q["celltype"].reshape(-1, nx)[:, 1:-1], want_ct[lo:hi, 1:-1]):
                    msgs.append(f"rank {r}: stored cell types (halo rows included) differ from the single-GPU run")
            print(f"[bodies moving={moving} solver={ppe_solver}] {ncx}x{ncy} world={world} step {step}: counts {counts}, "
                  f"{len(gc['cell'])} ghost cells ({[len(q['gc']['cell']) for q in parts]}): {'OK' if not msgs else msgs}")
            ok = ok and not msgs
        ref.close(); o.close()
    return ok


def main():
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    dist.init_process_group("nccl")
    ok = True
    if os.environ.get("IFX_MGPU_QUICK"):      # tools/sanitize.sh: the two cases that touch every slab code path, small
        ok = run_case(ifx.IFX_COMPAT_REFERENCE, 300, 301, 1, 12) and ok
        ok = run_bodies_case(300, 161, 1, moving=True) and ok
        flag = torch.tensor([1 if ok else 0], device="cuda")
        dist.broadcast(flag, 0)
        dist.destroy_process_group()
        print("MGPU_OK" if int(flag.item()) == 1 else "MGPU_FAIL")
        sys.exit(0 if int(flag.item()) == 1 else 1)
    ok = run_case(ifx.IFX_COMPAT_REFERENCE, 300, 301, 3, 40) and ok
    ok = run_case(ifx.IFX_COMPAT_REFERENCE, 1100, 1200, 2, 24) and ok
    ok = run_case(ifx.IFX_COMPAT_FULL, 260, 130, 3, 40, bc={"u_bc_w": 0.0, "u_bc_e": 0.0, "u_bc_s": 0.0, "u_bc_n": 1.0}) and ok
    ok = run_bodies_case(257, 130, 3, moving=False) and ok
    ok = run_bodies_case(300, 161, 3, moving=True) and ok
    ok = run_bodies_case(257, 130, 2, moving=False, ppe_solver=3, ppe_omega=1.8) and ok      # red-black SOR on slabs
    # both loops converge mid-way (predictor after 9 iterations, Poisson after ~15 / ~25 sweeps): lagged stop decisions
    ok = run_bodies_case(300, 161, 3, moving=True, scale=1e-9, ad_tol=2e-11, ppe_tol=0.1023) and ok
    ok = run_bodies_case(300, 161, 2, moving=True, scale=1e-9, ad_tol=2e-11, ppe_tol=0.1018) and ok
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, 0)
    dist.destroy_process_group()
    if dist.is_initialized() is False and int(flag.item()) != 1:
        sys.exit(1)
    print("MGPU_OK" if int(flag.item()) == 1 else "MGPU_FAIL")
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
