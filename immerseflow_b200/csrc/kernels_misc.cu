// kernels_misc.cu — layout conversion between the reference's dense id = i + j*nx arrays (C-ABI
// boundary) and the padded HBM layout, plus small fills.
#include "kernels.cuh"

namespace ifx {

static __global__ void k_fill_u8(uint8_t* p, size_t n, uint8_t v) {
  size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; t < n; t += stride) p[t] = v;
}

// dense rows [j_lo, j_hi) (global) <-> padded local rows; dense points at global row j_lo.
static __global__ void k_pack(Layout L, const double* __restrict__ padded, double* __restrict__ dense,
                       int j_lo, int j_hi) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int j = j_lo + blockIdx.y;
  if (i < L.nx && j < j_hi) dense[(size_t)(j - j_lo) * L.nx + i] = padded[lidx(L, i, j - L.j0)];
}
static __global__ void k_unpack(Layout L, const double* __restrict__ dense, double* __restrict__ padded,
                         int j_lo, int j_hi) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int j = j_lo + blockIdx.y;
  if (i < L.nx && j < j_hi) padded[lidx(L, i, j - L.j0)] = dense[(size_t)(j - j_lo) * L.nx + i];
}
// cell types -> doubles.  raw = 0: the reference's iBlank (1.0 fluid, 0.0 otherwise); raw = 1: type code (low 2 bits).
static __global__ void k_pack_u8(Layout L, const uint8_t* __restrict__ padded, double* __restrict__ dense,
                          int j_lo, int j_hi, int raw) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int j = j_lo + blockIdx.y;
  if (i < L.nx && j < j_hi) {
    const uint8_t c = padded[lidx(L, i, j - L.j0)];
    dense[(size_t)(j - j_lo) * L.nx + i] = raw ? (double)(c & 3) : (c == IFX_FLUID ? 1.0 : 0.0);
  }
}

// dst <- src, 4-byte words; either side may be page-locked host memory (unified addressing: same pointer on the
// device), which is how the zero-copy control path moves its few bytes without touching a copy engine
static __global__ void k_copy_words(uint32_t* __restrict__ dst, const uint32_t* __restrict__ src, size_t n) {
  size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; t < n; t += stride) dst[t] = src[t];
}
cudaError_t launch_copy_words(void* dst, const void* src, size_t bytes, cudaStream_t st) {
  const size_t n = (bytes + 3) / 4;
  if (n == 0) return cudaSuccess;
  const unsigned blocks = (unsigned)((n + 255) / 256 > 64 ? 64 : (n + 255) / 256);
  k_copy_words<<<blocks, 256, 0, st>>>(static_cast<uint32_t*>(dst), static_cast<const uint32_t*>(src), n);
  return cudaGetLastError();
}

// Face mask of the general Poisson operator (closed-face rule): one byte per cell, derived from the cell
// types once per classification so that the sweep reads ONE byte per cell instead of the types of the cell and of
// its four neighbours:  bit 0 / 1 / 2 / 3 = the W / E / S / N face is open (the neighbour is a fluid cell inside the
// grid), bit 4 = the cell itself is fluid.  Non-fluid cells, the ghost ring and the padding get 0, so 0x1f is "plain
// interior cell" and 0 is "copy".  16 cells per thread; rows jl_lo .. jl_hi-1 (the rows above and below must be
// classified).
static __global__ void __launch_bounds__(256)
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
      unsigned m = 16u;
      if (i > 1 && w == IFX_FLUID) m |= 1u;
      if (i < L.nx - 2 && e == IFX_FLUID) m |= 2u;
      if (j > 1 && s == IFX_FLUID) m |= 4u;
      if (j < L.ny - 2 && n == IFX_FLUID) m |= 8u;
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
  k_build_facemask<<<g, 256, 0, st>>>(L, celltype, facemask, jl_lo, jl_hi);
  return cudaGetLastError();
}

cudaError_t launch_fill_u8(uint8_t* p, size_t n, uint8_t v, cudaStream_t st) {
  k_fill_u8<<<296, 256, 0, st>>>(p, n, v);
  return cudaGetLastError();
}
cudaError_t launch_pack_u8(const Layout& L, const uint8_t* padded, double* dense, int raw, cudaStream_t st) {
  dim3 g((L.nx + 127) / 128, L.nyl);
  k_pack_u8<<<g, 128, 0, st>>>(L, padded, dense, L.j0, L.j0 + L.nyl, raw);
  return cudaGetLastError();
}

}  // namespace ifx
