// capi_diag.cu — C-ABI of the run-time diagnostics: ifx_probe, ifx_body_forces.
#include "solver.h"
#include "diag.cuh"

using namespace ifx;

static int diag_ready(ifx_solver* s, const char* what) {
  if (s->opt.compat != IFX_COMPAT_FULL)
    return fail(s, IFX_ERR_INVALID, std::string(what) + " requires IFX_COMPAT_FULL (the reference has no diagnostics)");
  if (s->opt.nranks != 1) return fail(s, IFX_ERR_INVALID, std::string(what) + " is single-GPU for now: gather the slabs first");
  if (s->bodies_dirty) return fail(s, IFX_ERR_STATE, std::string(what) + ": the bodies changed since the last classification (ifx_iblank_update / ifx_step)");
  return IFX_OK;
}

static int probe_points(ifx_solver* s, int n, const double* x, const double* y, double* u, double* v, double* p) {
  if (n <= 0) return IFX_OK;
  double* d = nullptr;
  IFX_CUDA(s, cudaMalloc(&d, sizeof(double) * 5 * (size_t)n));
  cudaError_t e = cudaMemcpyAsync(d, x, sizeof(double) * n, cudaMemcpyHostToDevice, s->stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(d + n, y, sizeof(double) * n, cudaMemcpyHostToDevice, s->stream);
  if (e == cudaSuccess) {
    s->launches++;
    e = launch_probe(s->L, s->M.xc, s->M.yc, s->celltype, s->u[s->cur_uv], s->v[s->cur_uv], s->p[s->cur_p], n, d, d + n,
                     d + 2 * (size_t)n, d + 3 * (size_t)n, d + 4 * (size_t)n, s->stream);
  }
  if (e == cudaSuccess) e = cudaMemcpyAsync(u, d + 2 * (size_t)n, sizeof(double) * n, cudaMemcpyDeviceToHost, s->stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(v, d + 3 * (size_t)n, sizeof(double) * n, cudaMemcpyDeviceToHost, s->stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(p, d + 4 * (size_t)n, sizeof(double) * n, cudaMemcpyDeviceToHost, s->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(s->stream);
  cudaFree(d);
  IFX_CUDA(s, e);
  return IFX_OK;
}

extern "C" int ifx_probe(ifx_solver* s, int n, const double* x, const double* y, double* u, double* v, double* p) {
  if (!s || n < 0 || (n > 0 && (!x || !y || !u || !v || !p))) return IFX_ERR_INVALID;
  int rc = diag_ready(s, "ifx_probe");
  if (rc != IFX_OK) return rc;
  IFX_CUDA(s, cudaSetDevice(s->device));
  return probe_points(s, n, x, y, u, v, p);
}

extern "C" int ifx_body_forces(ifx_solver* s, double* forces, int capacity_bodies) {
  if (!s || !forces) return IFX_ERR_INVALID;
  int rc = diag_ready(s, "ifx_body_forces");
  if (rc != IFX_OK) return rc;
  if (capacity_bodies < s->nbodies) return fail(s, IFX_ERR_INVALID, "capacity too small");
  if (s->nbodies == 0) return IFX_OK;
  IFX_CUDA(s, cudaSetDevice(s->device));
  const int ns = s->h_body_off[s->nbodies];
  std::vector<double> geo;
  force_geometry(s->L.nx, s->L.ny, s->h_xc.data(), s->h_yc.data(), s->nbodies, s->h_body_off.data(), s->h_xm.data(),
                 s->h_ym.data(), geo);
  std::vector<double> q(10 * (size_t)ns);         // the ns points P1, then the ns points P2
  double *px = q.data(), *py = px + 2 * ns, *pu = py + 2 * ns, *pv = pu + 2 * ns, *pp = pv + 2 * ns;
  for (int k = 0; k < ns; k++) {
    px[k] = geo[8 * (size_t)k]; py[k] = geo[8 * (size_t)k + 1];
    px[ns + k] = geo[8 * (size_t)k + 2]; py[ns + k] = geo[8 * (size_t)k + 3];
  }
  if ((rc = probe_points(s, 2 * ns, px, py, pu, pv, pp)) != IFX_OK) return rc;
  force_sum(s->nbodies, s->h_body_off.data(), geo.data(), pu, pv, pp, s->h_ub.data(), s->h_vb.data(), s->in.Re, forces);
  return IFX_OK;
}
