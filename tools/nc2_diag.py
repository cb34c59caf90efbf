"""The stage-release race of the TMA row pipeline, GPU against GPU over every cell (profiles/r2_nc2_race.md).

The benchmark's steps (16384 x 16384, 8 moving lobed bodies, smooth non-zero starting pressure) are run in fresh processes,
one library each, and u, v, p are compared through a 64-bit xor fingerprint per (row, 64-column block) over the whole grid.
No torch, no oracle: the reference is the build every oracle comparison of round 2 was made with — two columns per thread
in the general Poisson sweep, stage released right after the loads.  Output: gpurun_out/nc2_diag.json.

    tools/build_variant.sh narrow_early "-DIFX_PPE_NC2_WIDE=1 -DIFX_RELEASE_EARLY=1"     # the reference (round-2 kernels)
    tools/build_variant.sh narrow       "-DIFX_PPE_NC2_WIDE=1"                           # late release, two columns
    tools/build_variant.sh wide_early   "-DIFX_RELEASE_EARLY=1"                          # the race
    gpurun -- python tools/nc2_diag.py                                                   # the shipped library = wide, late

(The runs on record used the same kernels under their earlier names: -DIFX_PPE_NC2=2 for the four-column geometry, which
is now the single-GPU default; the hypothesis builds of r2_nc2_diag_variants.json are tools/experiments/kernels_v4_race_knobs.cu.txt.)
"""
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
N = int(os.environ.get("NC2_DIAG_N", "16384"))
BLK = 64


def child(tag, outdir, steps, mode):
    import bench
    import immerseflow_b200 as ifx
    t0 = time.time()
    mg = mode == "mg"              # 16 multigrid V-cycles (red-black SOR smoother = the same kernel) instead of 50 Jacobi sweeps
    inp = ifx.make_input(N, N, 1e-3, 150.0, AD_itermax=25, PPE_itermax=16 if mg else 50)
    xf = ifx.uniform_faces(N, 1.0)
    kw = dict(ppe_solver=4, ppe_tol=0.5) if mg else dict(ppe_pairs=1 if mode == "pairs" else 0)
    with ifx.ImmerseFlow(inp, xf, xf, sweeps_per_batch=51, compat=ifx.IFX_COMPAT_FULL, ppe_abs_residual=1, **kw) as h:
        h.initializeData()
        h.set("p", np.load(os.path.join(outdir, "p0.npy"), mmap_mode="r"))
        res = {"tag": tag, "steps": steps, "mode": mode, "library": os.environ.get("IFX_LIBRARY", "default"), "ppe_residual": [],
               "ms_ppe": [], "ms_ad": []}
        buf = np.empty(inp.nx * inp.ny)
        for it in range(steps):
            b, vel = bench.bodies_at(8, it, 1e-3)
            h.set_bodies(b, vel)
            st = h.step()
            res["ppe_residual"].append(float(st.ppe_residual))
            res["ms_ppe"].append(round(float(st.ms_ppe), 3)); res["ms_ad"].append(round(float(st.ms_ad), 3))
            if it in (0, steps - 1):
                for k in ("u", "v", "p"):
                    bits = h.get(k, buf).reshape(inp.ny, inp.nx).view(np.uint64)[:, 1:-1]
                    np.save(os.path.join(outdir, f"{tag}_{k}_{it + 1}.npy"), np.bitwise_xor.reduce(bits.reshape(inp.ny, -1, BLK), axis=2))
        res["counts"] = [int(st.ad_iters), int(st.ppe_sweeps)]
    res["seconds"] = time.time() - t0
    json.dump(res, open(os.path.join(outdir, f"{tag}.json"), "w"))


def blobs(bad):
    """connected groups of differing (row, block) cells, as bounding boxes (rows joined across gaps of <= 2)"""
    if not len(bad):
        return []
    cells = {(int(r), int(c)) for r, c in bad}
    seen, out = set(), []
    for c0 in sorted(cells):
        if c0 in seen:
            continue
        stack, box = [c0], [c0[0], c0[0], c0[1], c0[1], 0]
        seen.add(c0)
        while stack:
            r, c = stack.pop()
            box[0], box[1], box[2], box[3], box[4] = min(box[0], r), max(box[1], r), min(box[2], c), max(box[3], c), box[4] + 1
            for dr in (-2, -1, 0, 1, 2):
                for dc in (-1, 0, 1):
                    n = (r + dr, c + dc)
                    if n in cells and n not in seen:
                        seen.add(n); stack.append(n)
        out.append(box)
    return out


def compare(outdir, a, sa, b, sb):
    out = {"pair": [f"{a}@{sa}", f"{b}@{sb}"]}
    per_tile = 512 // BLK
    for k in ("u", "v", "p"):
        fa, fb = np.load(os.path.join(outdir, f"{a}_{k}_{sa}.npy")), np.load(os.path.join(outdir, f"{b}_{k}_{sb}.npy"))
        bad = np.argwhere(fa != fb)
        rec = {"blocks_differing": int(len(bad))}
        if len(bad) and len(bad) < 200000:
            bl = blobs(bad)
            rec["blobs"] = len(bl)
            rec["block_in_512_tile"] = {int(c): int(n) for c, n in zip(*np.unique(bad[:, 1] % per_tile, return_counts=True))}
            # centre of every blob: row within its 64-row tile, 64-column block within its 512-column tile
            rec["blob_boxes"] = [[r0, r1, c0 * BLK + 1, (c1 + 1) * BLK, n] for r0, r1, c0, c1, n in bl[:40]]
            rec["blob_centres"] = [[(r0 + r1) // 2, ((r0 + r1) // 2 - 1) % 64, ((c0 + c1 + 1) * BLK // 2) % 512] for r0, r1, c0, c1, n in bl[:40]]
        out[k] = rec
    return out


def dump(result, t0):
    result["seconds"] = time.time() - t0
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(result, open(os.path.join(ROOT, "gpurun_out", "nc2_diag.json"), "w"), indent=1)


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "child":
        return child(sys.argv[2], sys.argv[3], int(sys.argv[4]), sys.argv[5])
    import shutil
    import tempfile
    t0 = time.time()
    base = "/dev/shm" if os.path.isdir("/dev/shm") and shutil.disk_usage("/dev/shm").free > 12e9 else "/tmp"
    outdir = tempfile.mkdtemp(prefix="nc2diag_", dir=base)
    result = {"n": N, "block_columns": BLK, "scratch": base, "runs": [], "compare": []}
    budget = float(os.environ.get("NC2_DIAG_BUDGET_S", "100"))
    long_steps = int(os.environ.get("NC2_DIAG_LONG_STEPS", "10"))
    bin_ = os.path.join(ROOT, "tools", "_bin")
    # (tag, library, steps, mode, compared with): "ref" = the stage release of rounds 1-2 with two columns per thread, the
    # build every oracle comparison of the round was made with.  (In r2_nc2_diag_release.json the same builds are called
    # ref / new = narrow_late / nc2_late = shipped / nc2_early = wide_early.)
    lib = lambda n: os.path.join(bin_, f"lib_{n}.so")
    plan = [("ref", lib("narrow_early"), long_steps, "jacobi", None),
            ("shipped", "", long_steps, "jacobi", "ref"),                   # late release, four columns on a single GPU
            ("narrow_late", lib("narrow"), long_steps, "jacobi", "ref"),    # late release, two columns (the slab geometry)
            ("wide_early", lib("wide_early"), 1, "jacobi", "ref"),          # control: the race, on this box
            ("shipped_pairs", "", 1, "pairs", "ref"),                       # two sweeps per pass (kernels_pair.cu), same iterates
            ("ref_mg", lib("narrow_early"), 1, "mg", None),
            ("shipped_mg", "", 1, "mg", "ref_mg"),
            ("narrow_late_mg", lib("narrow"), 1, "mg", "ref_mg")]
    only = os.environ.get("NC2_DIAG_ONLY")
    if only:
        plan = [q for q in plan if q[0] in only.split(",")]
    try:
        # bench.initial_pressure's field up to rounding (the outer product takes a second, the bench's formula 20)
        c = 6.283185307179586 / (N + 2)
        np.save(os.path.join(outdir, "p0.npy"), 50.0 + 40.0 * np.outer(np.cos(np.arange(N + 2) * c), np.sin(np.arange(N + 2) * c)))
        done = []
        for tag, path, steps, mode, against in plan:
            if time.time() - t0 > budget:
                result.setdefault("skipped", []).append(tag)
                continue
            env = dict(os.environ)
            env.pop("IFX_LIBRARY", None)
            if path:
                if not os.path.exists(path):
                    result["runs"].append({"tag": tag, "error": f"{path} not built"})
                    continue
                env["IFX_LIBRARY"] = path
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "child", tag, outdir, str(steps), mode], env=env,
                               capture_output=True, text=True)
            if r.returncode != 0:
                result["runs"].append({"tag": tag, "error": r.stderr[-600:]})
                continue
            result["runs"].append(json.load(open(os.path.join(outdir, f"{tag}.json"))))
            done.append(tag)
            if against in done:
                result["compare"].append(compare(outdir, against, 1, tag, 1))
                if steps > 1:
                    result["compare"].append(compare(outdir, against, steps, tag, steps))
            dump(result, t0)
    finally:
        shutil.rmtree(outdir, ignore_errors=True)
    dump(result, t0)
    for c_ in result["compare"]:
        print(c_["pair"], {k: (c_[k]["blocks_differing"], c_[k].get("blobs")) for k in "uvp"})


if __name__ == "__main__":
    main()
