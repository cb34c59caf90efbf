"""Smallest possible GPU check of the single-GPU (four-column) geometry of the general Poisson sweep through the shipped
host code: two full-mode steps with two bodies on 700 x 500 cells (one full 512-column tile and one edge tile per tile row)
against the CPU oracle, every cell.  A few seconds; no torch.  Output: gpurun_out/wide_host_check.json."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
t0 = time.time()
import immerseflow_b200 as ifx      # noqa: E402
import _oracle as orc               # noqa: E402

ncx, ncy, dt, Re, ad_it, ppe_it = 700, 500, 2e-3, 150.0, 8, 20
xf, yf = orc.stretched_faces(ncx, 6.0, 1.002), orc.stretched_faces(ncy, 4.0, 1.003)
inp = ifx.make_input(ncx, ncy, dt, Re, AD_itermax=ad_it, PPE_itermax=ppe_it)
bodies = [orc.circle_markers(1.5, 2.0, 0.4, 48), orc.ellipse_markers(float(xf[520]), 1.7, 0.5, 0.3, 0.4, 64)]
out = {"grid": [ncx, ncy], "steps": []}
o = orc.FullSolver(xf, yf, dt, Re, ad_it, ppe_it, ppe_abs=1)
with ifx.ImmerseFlow(inp, xf, yf, compat=ifx.IFX_COMPAT_FULL, ppe_abs_residual=1) as s:
    s.initializeData()
    u0, v0 = s.get("u"), s.get("v")
    o.set("u", u0); o.set("v", v0)
    s.set_bodies(bodies); o.set_bodies(bodies); o.update_ib()
    inner = np.zeros((inp.ny, inp.nx), bool); inner[1:-1, 1:-1] = True
    inner = inner.reshape(-1)
    for step in range(2):
        st, so = s.step(), o.step()
        rec = {"counts_equal": (st.ad_iters, st.ppe_sweeps) == (int(so[0]), int(so[3]))}
        for k in ("u", "v", "p"):
            rec[k + "_differing"] = int(np.count_nonzero(s.get(k)[inner] != o.get(k)[inner]))
        out["steps"].append(rec)
o.close()
out["bit_exact"] = all(r["counts_equal"] and not (r["u_differing"] or r["v_differing"] or r["p_differing"]) for r in out["steps"])
out["seconds"] = time.time() - t0
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "wide_host_check.json"), "w"), indent=1)
print(json.dumps(out))
