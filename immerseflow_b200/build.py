"""Build the CUDA shared library (C-ABI) and the CLI driver in-tree with nvcc for sm_100a.

No torch extension machinery: the product is a plain C-ABI .so (include/immerseflow_c.h) that the
C++ driver links and that Python loads with ctypes.  Objects land in immerseflow_b200/_build/,
the library in immerseflow_b200/libimmerseflow_b200.so (git-ignored, but it travels to the GPU box).
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "_build")
LIB = os.path.join(HERE, "libimmerseflow_b200.so")
CLI = os.path.join(HERE, "bin", "immerseflow")

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC,-ffp-contract=off,-Wall", "-Xptxas", "-v"]
# solver kernels: no implicit FMA contraction — every fma() in the source is deliberate
STRICT = ["-fmad=false"]

# (source, strict-fp?)
SOURCES = [
    ("capi.cu", True),
    ("capi_full.cu", True),
    ("capi_ckpt.cu", True),
    ("capi_mg.cu", True),
    ("kernels_ad.cu", True),
    ("kernels_ppe.cu", True),
    ("kernels_v4.cu", True),
    ("kernels_pair.cu", True),
    ("kernels_reduce.cu", True),
    ("kernels_misc.cu", True),
    ("kernels_facemask.cu", True),
    ("kernels_halo.cu", True),
    ("kernels_full.cu", True),
    ("kernels_ib.cu", True),
    ("kernels_mg.cu", True),
    ("kernels_diag.cu", True),
    ("capi_diag.cu", True),
    ("kernels_ic.cu", False),   # default flags on purpose: same libdevice expansion as the reference build
    ("io.cpp", True),
]


def _nvcc() -> str:
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: the CUDA library cannot be built (there is no CPU fallback)")
    return nvcc


def _stale(target: str, deps: list[str]) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build(verbose: bool = False, force: bool = False) -> str:
    nvcc = _nvcc()
    os.makedirs(OUT, exist_ok=True)
    os.makedirs(os.path.dirname(CLI), exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    headers.append(os.path.join(HERE, "..", "include", "immerseflow_c.h"))
    headers.append(os.path.abspath(__file__))
    objs = []
    log = []
    for src, strict in SOURCES:
        sp = os.path.join(CSRC, src)
        if not os.path.exists(sp):
            continue
        obj = os.path.join(OUT, src.rsplit(".", 1)[0] + ".o")
        objs.append(obj)
        if force or _stale(obj, [sp] + headers):
            cmd = [nvcc, *ARCH, *COMMON, *(STRICT if strict else []), "-x", "cu", "-c", sp, "-o", obj]
            r = subprocess.run(cmd, capture_output=True, text=True)
            log.append("$ " + " ".join(cmd) + "\n" + r.stdout + r.stderr)
            if r.returncode != 0:
                sys.stderr.write(log[-1])
                raise RuntimeError(f"nvcc failed on {src}")
            if verbose:
                print(log[-1])
    if force or _stale(LIB, objs):
        cmd = [nvcc, *ARCH, "-shared", "-o", LIB, *objs, "-Xcompiler", "-fPIC"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("link failed")
    main_cpp = os.path.join(CSRC, "main.cpp")
    if os.path.exists(main_cpp) and (force or _stale(CLI, [main_cpp, LIB])):
        cmd = [nvcc, "-O2", "-std=c++17", main_cpp, "-o", CLI, "-L" + HERE, "-limmerseflow_b200",
               "-Xlinker", "-rpath", "-Xlinker", "$ORIGIN/.."]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("CLI link failed")
    with open(os.path.join(OUT, "build.log"), "w") as f:
        f.write("\n".join(log))
    return LIB


if __name__ == "__main__":
    print(build(verbose="-v" in sys.argv, force="-f" in sys.argv))
