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


def main():
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    dist.init_process_group("nccl")
    ok = True
    ok = run_case(ifx.IFX_COMPAT_REFERENCE, 300, 301, 3, 40) and ok
    ok = run_case(ifx.IFX_COMPAT_REFERENCE, 1100, 1200, 2, 24) and ok
    ok = run_case(ifx.IFX_COMPAT_FULL, 260, 130, 3, 40, bc={"u_bc_w": 0.0, "u_bc_e": 0.0, "u_bc_s": 0.0, "u_bc_n": 1.0}) and ok
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, 0)
    dist.destroy_process_group()
    if dist.is_initialized() is False and int(flag.item()) != 1:
        sys.exit(1)
    print("MGPU_OK" if int(flag.item()) == 1 else "MGPU_FAIL")
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
