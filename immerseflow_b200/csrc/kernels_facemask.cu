// kernels_facemask.cu — the face mask of the general Poisson operator, derived from the cell types once per classification.
// No shared memory, no intrinsics: also compiles under tests/shim/cuda_host_shim.h (tests/test_facemask_shim.py holds it
// to the definition on the CPU).
#include "facemask.cuh"

#ifdef IFX_HOST_SHIM
#define IFX_KLAUNCH(k, grid, block, st, ...) (shim_launch((grid), (block), [&] { k(__VA_ARGS__); }), cudaSuccess)
#define IFX_LAUNCH_BOUNDS(n)
#else
#define IFX_KLAUNCH(k, grid, block, st, ...) (k<<<(grid), (block), 0, (st)>>>(__VA_ARGS__), cudaGetLastError())
#define IFX_LAUNCH_BOUNDS(n) __launch_bounds__(n)
#endif

namespace ifx {

// Face mask of the general Poisson operator (closed-face rule): one byte per cell, derived from the cell
// types once per classification so that the sweep reads ONE byte per cell instead of the types of the cell and of
// its four neighbours:  bit 0 / 1 / 2 / 3 = the W / E / S / N face is open (the neighbour is a fluid cell inside the
// grid), bit 4 = the cell itself is fluid.  Non-fluid cells, the ghost ring and the padding get 0, so 0x1f is "plain
// interior cell" and 0 is "copy".  16 cells per thread; rows jl_lo .. jl_hi-1 (the rows above and below must be
// classified).
static __global__ void IFX_LAUNCH_BOUNDS(256)
k_build_facemask(Layout L, const uint8_t* __restrict__ ct, uint8_t* __restrict__ fm, int jl_lo, int jl_hi) {
  const int jl = jl_lo + blockIdx.y;
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (jl >= jl_hi || c * 16 >= L.pitch) return;
  const int j = L.j0 + jl;
  const uint8_t* rowC = ct + (ptrdiff_t)jl * L.pitch + 16 * c;
  uint4 out = make_uint4(0u, 0u, 0u, 0u);
  if (j >= 1 && j <= L.ny - 2) {
    const uint4 vC = *reinterpret_cast<const uint4*>(rowC);
    const uint4 vS = *reinterpret_cast<const uint4*>(rowC - L.pitch);
    const uint4 vN = *reinterpret_cast<const uint4*>(rowC + L.pitch);
    const unsigned wC[4] = {vC.x, vC.y, vC.z, vC.w}, wS[4] = {vS.x, vS.y, vS.z, vS.w}, wN[4] = {vN.x, vN.y, vN.z, vN.w};
    const unsigned tW = (c > 0) ? rowC[-1] : 0u, tE = (16 * c + 16 < L.pitch) ? rowC[16] : 0u;
    unsigned o[4] = {0u, 0u, 0u, 0u};
#pragma unroll
    for (int k = 0; k < 16; k++) {
      const int i = 16 * c - IFX_PADL + k;
      const unsigned t = (wC[k >> 2] >> (8 * (k & 3))) & 0xffu;
      if (i < 1 || i > L.nx - 2 || t != IFX_FLUID) continue;
      const unsigned w = (k == 0) ? tW : (wC[(k - 1) >> 2] >> (8 * ((k - 1) & 3))) & 0xffu;
      const unsigned e = (k == 15) ? tE : (wC[(k + 1) >> 2] >> (8 * ((k + 1) & 3))) & 0xffu;
      const unsigned s = (wS[k >> 2] >> (8 * (k & 3))) & 0xffu, n = (wN[k >> 2] >> (8 * (k & 3))) & 0xffu;
      unsigned m = IFX_FM_FLUID;
      if (i > 1 && w == IFX_FLUID) m |= IFX_FM_W;
      if (i < L.nx - 2 && e == IFX_FLUID) m |= IFX_FM_E;
      if (j > 1 && s == IFX_FLUID) m |= IFX_FM_S;
      if (j < L.ny - 2 && n == IFX_FLUID) m |= IFX_FM_N;
      o[k >> 2] |= m << (8 * (k & 3));
    }
    out = make_uint4(o[0], o[1], o[2], o[3]);
  }
  *reinterpret_cast<uint4*>(fm + (ptrdiff_t)jl * L.pitch + 16 * c) = out;
}
cudaError_t launch_build_facemask(const Layout& L, const uint8_t* celltype, uint8_t* facemask, int jl_lo, int jl_hi,
                                  cudaStream_t st) {
  if (jl_hi <= jl_lo) return cudaSuccess;
  dim3 g((L.pitch / 16 + 255) / 256, jl_hi - jl_lo);
  return IFX_KLAUNCH(k_build_facemask, g, dim3(256), st, L, celltype, facemask, jl_lo, jl_hi);
}

}  // namespace ifx
