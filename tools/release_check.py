"""Late stage release (kernels_v4.cu: IFX_RELEASE_EARLY) against the rounds 1-2 placement, on the template variants the
16384^2 run of tools/nc2_diag.py does not reach: the reference-compatible mode (Laplace sweep), grids that are not a
multiple of the tile width (edge tiles), the reference-order reduction (residual fields written), red-black SOR, tolerances
that stop the loops early.  Both libraries run the same cases in their own process; u, v, p must be equal in every bit
after every step.  No torch, no oracle.  Output: gpurun_out/release_check.json.

    gpurun -- python tools/release_check.py      (needs tools/_bin/lib_narrow_early.so: tools/build_variant.sh narrow_early
    "-DIFX_PPE_NC2_WIDE=1 -DIFX_RELEASE_EARLY=1"; the run on record compared the two-column geometry both ways)
"""
import hashlib
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

CASES = [
    # name, cells x, cells y, compat full?, options, bodies, steps
    ("reference_mode_700x1000", 700, 1000, 0, {}, 0, 3),
    ("reference_mode_1030x2000", 1030, 2000, 0, {}, 0, 2),
    ("reference_mode_exact_reduce_777", 777, 777, 0, {"reduce_mode": 1}, 0, 2),
    ("full_bodies_1000x1300", 1000, 1300, 1, {"ppe_abs_residual": 1}, 3, 3),
    ("full_exact_reduce_bodies_900", 900, 900, 1, {"ppe_abs_residual": 1, "reduce_mode": 1}, 2, 2),
    ("full_sor_1100x600", 1100, 600, 1, {"ppe_abs_residual": 1, "ppe_solver": 3, "ppe_omega": 1.5}, 2, 2),
    ("full_tolerances_2048", 2048, 2048, 1, {"ppe_abs_residual": 1, "ppe_tol": 0.4, "ad_tol": 1e-7}, 2, 3),
    ("full_pairs_1500x1000", 1500, 1000, 1, {"ppe_abs_residual": 1, "ppe_pairs": 1}, 2, 2),
    ("full_small_tiles_300x40", 300, 40, 1, {"ppe_abs_residual": 1}, 1, 2),
]


def child():
    import bench
    import immerseflow_b200 as ifx
    out = {}
    only = os.environ.get("RELEASE_CHECK_ONLY", "")
    for name, ncx, ncy, full, opts, nb, steps in CASES:
        if only and not name.startswith(only):
            continue
        inp = ifx.make_input(ncx, ncy, 1e-3, 150.0, AD_itermax=25, PPE_itermax=50)
        xf, yf = ifx.uniform_faces(ncx, 1.0), ifx.uniform_faces(ncy, 1.0)
        rec = []
        try:
            with ifx.ImmerseFlow(inp, xf, yf, compat=ifx.IFX_COMPAT_FULL if full else ifx.IFX_COMPAT_REFERENCE, **opts) as h:
                h.initializeData()
                for it in range(steps):
                    if full:
                        if nb:
                            b, vel = bench.bodies_at(nb, it, 1e-3)
                            h.set_bodies(b, vel)
                        st = h.step()
                        cnt = [int(st.ad_iters), int(st.ppe_sweeps)]
                    else:
                        a, p = h.ADsolver(), h.PPESolver()
                        cnt = [int(a.ad_iters), int(p.ppe_sweeps)]
                    rec.append({"counts": cnt, **{k: hashlib.blake2b(h.get(k).tobytes(), digest_size=16).hexdigest() for k in ("u", "v", "p")}})
        except Exception as ex:       # noqa: BLE001
            rec.append({"error": f"{type(ex).__name__}: {ex}"[:300]})
        out[name] = rec
    print("RESULT " + json.dumps(out))


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "child":
        return child()
    t0 = time.time()
    res = {}
    for tag, lib in (("early", os.path.join(ROOT, "tools", "_bin", "lib_narrow_early.so")), ("late", "")):
        env = dict(os.environ)
        env.pop("IFX_LIBRARY", None)
        if lib:
            env["IFX_LIBRARY"] = lib
        r = subprocess.run([sys.executable, os.path.abspath(__file__), "child"], env=env, capture_output=True, text=True)
        line = [l for l in r.stdout.splitlines() if l.startswith("RESULT ")]
        res[tag] = json.loads(line[0][7:]) if line else {"error": (r.stderr or r.stdout)[-500:]}
    summary = {}
    only = os.environ.get("RELEASE_CHECK_ONLY", "")
    for name, *_ in CASES:
        if only and not name.startswith(only):
            continue
        a, b = res["early"].get(name), res["late"].get(name)
        ok = bool(a) and a == b and all("error" not in s for s in a)
        summary[name] = {"equal": ok, "steps": len(a or []), "counts": [s.get("counts") for s in (b or [])]}
        if not ok:
            summary[name]["early"], summary[name]["late"] = a, b
    out = {"all_equal": all(v["equal"] for v in summary.values()), "cases": summary, "seconds": time.time() - t0}
    if "error" in res["early"] or "error" in res["late"]:
        out["errors"] = {k: v.get("error") for k, v in res.items() if "error" in v}
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "release_check%s.json" % ("_" + only if only else "")), "w"), indent=1)
    print(json.dumps(out)[:3000])


if __name__ == "__main__":
    main()
