// kernels_ppe.cu — the reference's pressure boundary values (the Poisson sweeps are in kernels_v4.cu).
#include "kernels.cuh"
#include "stencil_math.cuh"

namespace ifx {

// set_pressure_BC (PPESolver.cu:54-71): p = 100 where i == 0 or j == 0, applied to both ping-pong
// buffers (the reference copies boundary cells through every sweep, PPESolver.cu:21).
static __global__ void k_set_pressure_bc_ref(Layout L, double* p0, double* p1) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < L.nyl) { p0[lidx(L, 0, t)] = 100.0; if (p1) p1[lidx(L, 0, t)] = 100.0; }
  if (t < L.nx && L.j0 == 0) { p0[lidx(L, t, 0)] = 100.0; if (p1) p1[lidx(L, t, 0)] = 100.0; }
}

cudaError_t launch_set_pressure_bc_ref(const Layout& L, double* p0, double* p1, cudaStream_t st) {
  const int n1 = L.nyl > L.nx ? L.nyl : L.nx;
  k_set_pressure_bc_ref<<<(n1 + 127) / 128, 128, 0, st>>>(L, p0, p1);
  return cudaGetLastError();
}

}  // namespace ifx
