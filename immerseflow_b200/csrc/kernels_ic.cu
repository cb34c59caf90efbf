// kernels_ic.cu — initial condition of the reference (initializeKernel, preSim.cu:53-76): uniform
// stream plus a Gaussian vortex at (0.5, 0.5), r0 = 0.1; p = 0.
// This translation unit is compiled with nvcc's DEFAULT floating-point flags (the reference's
// Makefile passes none), not with -fmad=false like the solver kernels, so that libdevice's
// pow/exp/sqrt expand exactly as they do in the reference build.
#include "kernels.cuh"

namespace ifx {

static __global__ void k_init_vortex(Layout L, const double* __restrict__ xc, const double* __restrict__ yc,
                              double* __restrict__ u, double* __restrict__ v, double* __restrict__ p) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int jl = blockIdx.y;
  if (i >= L.nx || jl >= L.nyl) return;
  const int j = L.j0 + jl;
  const size_t o = lidx(L, i, jl);
  double r = sqrt(pow(xc[i] - 0.5, 2.0) + pow(yc[j] - 0.5, 2.0));
  double r0 = 0.1;
  u[o] = 1.0 - 0.25 * (yc[j] - 0.5) * exp((1.0 - pow(r / r0, 2.0)) / 2.0);
  v[o] = 0.25 * (xc[i] - 0.5) * exp((1.0 - pow(r / r0, 2.0)) / 2.0);
  p[o] = 0.0;
}

cudaError_t launch_init_vortex(const Layout& L, const double* xc, const double* yc, double* u, double* v, double* p,
                               cudaStream_t st) {
  dim3 g((L.nx + 127) / 128, L.nyl);
  k_init_vortex<<<g, 128, 0, st>>>(L, xc, yc, u, v, p);
  return cudaGetLastError();
}

}  // namespace ifx
