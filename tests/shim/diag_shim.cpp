// diag_shim.cpp — TEST INFRASTRUCTURE: C entry points over the host-shim build of kernels_diag.cu and the host
// arithmetic of diag.cuh (see cuda_host_shim.h).  Arrays are host memory; fields use the product's padded layout.
#include "cuda_host_shim.h"
thread_local dim3 threadIdx, blockIdx, blockDim, gridDim;
#include "../../immerseflow_b200/csrc/kernels_diag.cu"

using namespace ifx;

extern "C" {
void shim_probe(int nx, int ny, int pitch, const double* xc, const double* yc, const uint8_t* ct, const double* u,
                const double* v, const double* p, int n, const double* px, const double* py, double* ou, double* ov, double* op) {
  launch_probe(Layout{nx, ny, pitch, ny, 0, 1, ny - 1}, xc, yc, ct, u, v, p, n, px, py, ou, ov, op, nullptr);
}
void shim_force_geometry(int nx, int ny, const double* xc, const double* yc, int nbodies, const int* off, const double* xm,
                         const double* ym, double* geo) {
  std::vector<double> g;
  force_geometry(nx, ny, xc, yc, nbodies, off, xm, ym, g);
  std::memcpy(geo, g.data(), sizeof(double) * g.size());
}
void shim_force_sum(int nbodies, const int* off, const double* geo, const double* pu, const double* pv, const double* pp,
                    const double* ub, const double* vb, double Re, double* F) {
  force_sum(nbodies, off, geo, pu, pv, pp, ub, vb, Re, F);
}
}
