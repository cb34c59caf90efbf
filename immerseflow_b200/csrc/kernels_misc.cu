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
