"""Which of {step kernels, H2D, D2H} overlap on this box?  Three handles on one GPU (8192 x 8192), wall-clock around
combinations of: one step on A, upload of u, v, p into B, download of u, v, p from C.  Diagnostic for bench.py's
end-to-end pipeline."""
import json
import sys
import time

import numpy as np
import torch

import immerseflow_b200 as ifx

n_c = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
inp = ifx.make_input(n_c, n_c, 1e-3, 150.0, AD_itermax=25, PPE_itermax=50)
xf = yf = ifx.uniform_faces(n_c, 1.0)


def make():
    h = ifx.ImmerseFlow(inp, xf, yf, sweeps_per_batch=51, compat=ifx.IFX_COMPAT_FULL, ppe_abs_residual=1,
                        zero_copy_control=int(sys.argv[2]) if len(sys.argv) > 2 else 1)
    h.initializeData()
    h.step()
    return h


A, B, Cc = make(), make(), make()
n = A.field_size("u")
names = ("u", "v", "p")
hin = {k: torch.empty(n, dtype=torch.float64, pin_memory=True).numpy() for k in names}
hout = {k: torch.empty(n, dtype=torch.float64, pin_memory=True).numpy() for k in names}
for k in names:
    A.get(k, hin[k])


def sync():
    for h in (A, B, Cc):
        h.synchronize()


def run(step, up, down):
    sync()
    t0 = time.perf_counter()
    if up:
        for k in names:
            B.set_async(k, hin[k])
    if down:
        for k in names:
            Cc.get_async(k, hout[k])
    t_enq = time.perf_counter() - t0
    if step:
        A.step()
    t_step = time.perf_counter() - t0
    sync()
    return {"enqueue_ms": t_enq * 1e3, "step_returned_ms": t_step * 1e3, "all_done_ms": (time.perf_counter() - t0) * 1e3}


out = {}
for name, cfg in (("step", (1, 0, 0)), ("h2d", (0, 1, 0)), ("d2h", (0, 0, 1)), ("h2d+d2h", (0, 1, 1)), ("step+h2d", (1, 1, 0)),
                  ("step+d2h", (1, 0, 1)), ("step+h2d+d2h", (1, 1, 1))):
    run(*cfg)
    out[name] = run(*cfg)
print(json.dumps(out))
