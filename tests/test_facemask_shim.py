"""The face-mask kernel of the general Poisson operator (immerseflow_b200/csrc/kernels_facemask.cu) compiled as plain C++
through tests/shim/ and run on the CPU: its bytes against the definition — bit 4 = the cell is fluid, bits 0-3 = the W / E /
S / N neighbour is a fluid cell inside the grid; everything that is not a fluid interior cell (solid, ghost cell, ghost ring,
padding) is 0 — on cell types classified by the oracle for bodies touching every special case: next to the grid boundary,
across a slab boundary, thin gaps between two bodies."""
import ctypes as C

import numpy as np
import pytest

import _oracle as orc
from shim import build as shim_build

PADL = 15


@pytest.fixture(scope="module")
def shim():
    lib = C.CDLL(shim_build.build("facemask", deps=("kernels_facemask.cu", "facemask.cuh", "common.cuh")))
    lib.shim_build_facemask.restype = None
    return lib


def _celltypes(ncx, ncy):
    xf, yf = orc.stretched_faces(ncx, 3.0, 1.01), orc.stretched_faces(ncy, 2.0, 1.01)
    o = orc.FullSolver(xf, yf, 1e-3, 100.0, 1, 1)
    # a body close to the west wall, two bodies with a gap of one or two cells between them, one near the north wall
    bodies = [orc.circle_markers(0.33, 1.0, 0.25, 40), orc.ellipse_markers(1.5, 0.9, 0.4, 0.2, 0.2, 64),
              orc.ellipse_markers(1.5, 1.36, 0.4, 0.2, -0.1, 64), orc.circle_markers(2.3, 1.72, 0.2, 32)]
    o.set_bodies(bodies); o.update_ib()
    ct = o.get("celltype").astype(np.uint8).reshape(ncy + 2, ncx + 2)     # type code (low 2 bits): 0 solid, 1 fluid, 2 ghost
    o.close()
    return ct


def _expected(ct):
    ny, nx = ct.shape
    fluid = ct == 1
    m = np.zeros_like(ct)
    inner = np.zeros_like(fluid); inner[1:-1, 1:-1] = True
    f = fluid & inner
    w = np.zeros_like(fluid); w[:, 2:] = f[:, 1:-1]          # west neighbour fluid and inside the grid (i > 1)
    e = np.zeros_like(fluid); e[:, :-2] = f[:, 1:-1]
    s = np.zeros_like(fluid); s[2:, :] = f[1:-1, :]
    n = np.zeros_like(fluid); n[:-2, :] = f[1:-1, :]
    m[f] = 16
    m[f & w] |= 1; m[f & e] |= 2; m[f & s] |= 4; m[f & n] |= 8
    return m


def _padded(ct, pitch, fill):
    ny, nx = ct.shape
    out = np.full((ny, pitch), fill, dtype=np.uint8)
    out[:, PADL:PADL + nx] = ct
    return out


@pytest.mark.parametrize("ncx,ncy", [(97, 61), (260, 70)])
def test_facemask_kernel_matches_the_definition(shim, ncx, ncy):
    ct = _celltypes(ncx, ncy)
    assert (ct == 0).any() and (ct == 2).any()
    ny, nx = ct.shape
    pitch = (PADL + nx + 1 + 15) // 16 * 16
    want = _expected(ct)
    # whole grid in one slab
    ctp = np.ascontiguousarray(_padded(ct, pitch, 1))                      # ring and padding are FLUID in the product's array
    fm = np.full((ny, pitch), 0xEE, dtype=np.uint8)
    shim.shim_build_facemask(nx, ny, pitch, ny, 0, ctp.ctypes.data_as(C.c_void_p), fm.ctypes.data_as(C.c_void_p), 1, ny - 1)
    assert np.array_equal(fm[1:-1, PADL:PADL + nx], want[1:-1]), "masks differ from the definition"
    assert not fm[1:-1, :PADL].any() and not fm[1:-1, PADL + nx:].any(), "padding must be 0"
    assert (fm[0] == 0xEE).all() and (fm[-1] == 0xEE).all(), "rows outside [jl_lo, jl_hi) must not be written"
    # the same rows as two slabs: rows 1 .. k and k+1 .. ny-2, each with one halo row either side
    k = ny // 2
    for jb, je in ((1, k + 1), (k + 1, ny - 1)):
        j0, nyl = jb - 1, je - jb + 2
        cts = np.ascontiguousarray(ctp[j0:j0 + nyl])
        fms = np.zeros((nyl, pitch), dtype=np.uint8)
        shim.shim_build_facemask(nx, ny, pitch, nyl, j0, cts.ctypes.data_as(C.c_void_p), fms.ctypes.data_as(C.c_void_p), 1, nyl - 1)
        assert np.array_equal(fms[1:-1, PADL:PADL + nx], want[jb:je]), (jb, je)
