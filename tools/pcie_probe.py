#!/usr/bin/env python
"""Host <-> device copy bandwidth on this box: contiguous (1-D) against pitched (2-D, the padded HBM layout of a field) copies,
one direction alone and both directions at once — what bounds bench.py's end-to-end number.  Prints one JSON line."""
import json
import time

import torch

nx, ny, pitch = 16386, 16386, 16416
host_a = torch.empty(ny * nx, dtype=torch.float64, pin_memory=True)
host_b = torch.empty(ny * nx, dtype=torch.float64, pin_memory=True)
dev_flat = torch.empty(ny * nx, dtype=torch.float64, device="cuda")
dev_flat2 = torch.empty(ny * nx, dtype=torch.float64, device="cuda")
dev_pad = torch.empty(ny, pitch, dtype=torch.float64, device="cuda")
dev_pad2 = torch.empty(ny, pitch, dtype=torch.float64, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
gb = ny * nx * 8 / 1e9


def timed(fn, reps=3):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps


def h2d_flat():
    with torch.cuda.stream(s1):
        dev_flat.copy_(host_a, non_blocking=True)


def d2h_flat():
    with torch.cuda.stream(s2):
        host_b.copy_(dev_flat2, non_blocking=True)


def h2d_pad():
    with torch.cuda.stream(s1):
        dev_pad[:, :nx].copy_(host_a.view(ny, nx), non_blocking=True)


def d2h_pad():
    with torch.cuda.stream(s2):
        host_b.view(ny, nx).copy_(dev_pad2[:, :nx], non_blocking=True)


out = {"field_GB": gb}
for name, fns in (("h2d_contiguous", [h2d_flat]), ("d2h_contiguous", [d2h_flat]), ("h2d_pitched", [h2d_pad]), ("d2h_pitched", [d2h_pad]),
                  ("both_contiguous", [h2d_flat, d2h_flat]), ("both_pitched", [h2d_pad, d2h_pad])):
    t = timed(lambda: [f() for f in fns])
    out[name + "_GBps_per_direction"] = gb / t
print(json.dumps(out))
