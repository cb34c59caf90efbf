"""TEST INFRASTRUCTURE: builds the host-shim libraries (thread-per-cell CUDA sources of the product compiled as plain
C++ by g++, see cuda_host_shim.h) into tests/shim/_build/."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "..", "..", "immerseflow_b200", "csrc")
OUT = os.path.join(HERE, "_build")


def build(name: str, deps=()):
    """name: 'mg' -> tests/shim/mg_shim.cpp -> _build/libmg_shim.so"""
    src = os.path.join(HERE, f"{name}_shim.cpp")
    lib = os.path.join(OUT, f"lib{name}_shim.so")
    watch = [src, os.path.join(HERE, "cuda_host_shim.h")] + [os.path.join(CSRC, d) for d in deps]
    if not os.path.exists(lib) or any(os.path.getmtime(w) > os.path.getmtime(lib) for w in watch):
        os.makedirs(OUT, exist_ok=True)
        # -ffp-contract=off mirrors the product's -fmad=false: only the explicit fma() calls fuse
        subprocess.check_call(["g++", "-O1", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared", "-DIFX_HOST_SHIM",
                               "-Wno-unknown-pragmas", "-o", lib, src])
    return lib
