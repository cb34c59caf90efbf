// kernels_diag.cu — probe kernel: bilinear interpolation of u, v, p at arbitrary points, one thread per point
// (oracle: orc_interp_setup + orc_probe, UNPINNED).  Interior-solid nodes are dropped and the weights renormalised,
// exactly like the ghost-cell image-point stencils (kernels_ib.cu); ghost cells take part with their
// boundary-condition values.  No shared memory, no intrinsics: also compiles under tests/shim/cuda_host_shim.h.
#include "diag.cuh"

#ifdef IFX_HOST_SHIM
#define IFX_KLAUNCH(k, grid, block, st, ...) (shim_launch((grid), (block), [&] { k(__VA_ARGS__); }), cudaSuccess)
#else
#define IFX_KLAUNCH(k, grid, block, st, ...) (k<<<(grid), (block), 0, (st)>>>(__VA_ARGS__), cudaGetLastError())
#endif

namespace ifx {

static __global__ void k_probe(Layout L, const double* __restrict__ xc, const double* __restrict__ yc,
                               const uint8_t* __restrict__ ct, const double* __restrict__ u, const double* __restrict__ v,
                               const double* __restrict__ p, int n, const double* __restrict__ px,
                               const double* __restrict__ py, double* __restrict__ ou, double* __restrict__ ov,
                               double* __restrict__ op) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const double x = px[k], y = py[k];
  const int i0 = box_index(xc, L.nx, x), j0 = box_index(yc, L.ny, y);
  const double a = (x - xc[i0]) / (xc[i0 + 1] - xc[i0]);
  const double b = (y - yc[j0]) / (yc[j0 + 1] - yc[j0]);
  const size_t o0 = lidx(L, i0, j0 - L.j0);
  const size_t o[4] = {o0, o0 + 1, o0 + L.pitch, o0 + L.pitch + 1};
  double w[4] = {(1.0 - a) * (1.0 - b), a * (1.0 - b), (1.0 - a) * b, a * b};
  double W = 0.0;
#pragma unroll
  for (int m = 0; m < 4; m++) {
    if ((ct[o[m]] & 3) == IFX_SOLID) w[m] = 0.0;
    W = W + w[m];
  }
#pragma unroll
  for (int m = 0; m < 4; m++) w[m] = (W > 0.0) ? w[m] / W : 0.0;
  auto interp = [&](const double* __restrict__ q) {
    double t = w[0] * q[o[0]];
    t = fma(w[1], q[o[1]], t);
    t = fma(w[2], q[o[2]], t);
    t = fma(w[3], q[o[3]], t);
    return t;
  };
  ou[k] = interp(u); ov[k] = interp(v); op[k] = interp(p);
}

cudaError_t launch_probe(const Layout& L, const double* xc, const double* yc, const uint8_t* celltype, const double* u,
                         const double* v, const double* p, int n, const double* px, const double* py, double* ou,
                         double* ov, double* op, cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  return IFX_KLAUNCH(k_probe, dim3((n + 127) / 128, 1, 1), dim3(128, 1, 1), st, L, xc, yc, celltype, u, v, p, n, px, py, ou, ov, op);
}

}  // namespace ifx
