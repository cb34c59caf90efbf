// diag.cuh — run-time diagnostics: probes and surface forces on the immersed bodies (the reference
// has neither, the predecessor had both: test/UTIL_PRE_SIM.f90:172-201).  Semantics: oracle/ifx_oracle_diag.c
// (PARITY UNPINNED).  The device part is one interpolation kernel; the per-segment geometry and the force sums are
// O(markers) host arithmetic in the oracle's operation order (this directory's host code is compiled
// -ffp-contract=off), kept in this header so that the CPU test suite can run it against the oracle.
#pragma once
#include "common.cuh"

#include <cmath>
#include <vector>

namespace ifx {

// largest i in [0, n-2] with c[i] <= x, clamped (the box of cell centres that contains x)
__host__ __device__ __forceinline__ int box_index(const double* c, int n, double x) {
  int lo = 0, hi = n - 2;
  if (x < c[0]) return 0;
  while (lo < hi) {
    const int mid = (lo + hi + 1) / 2;
    if (c[mid] <= x) lo = mid; else hi = mid - 1;
  }
  return lo;
}

// per marker segment: probe points P1 = M + delta n, P2 = M + 2 delta n
// (8 doubles: P1x, P1y, P2x, P2y, nx, ny, ds, delta); oracle: orc_force_geometry
inline void force_geometry(int nx, int ny, const double* xc, const double* yc, int nbodies, const int* off, const double* xm,
                           const double* ym, std::vector<double>& geo) {
  geo.assign(8 * (size_t)(nbodies ? off[nbodies] : 0), 0.0);
  for (int b = 0; b < nbodies; b++) {
    const int n = off[b + 1] - off[b];
    for (int k = 0; k < n; k++) {
      const int ka = off[b] + k, kb = off[b] + ((k + 1 == n) ? 0 : k + 1);
      const double ex = xm[kb] - xm[ka], ey = ym[kb] - ym[ka];
      const double len = std::sqrt(ex * ex + ey * ey);
      double* g = geo.data() + 8 * (size_t)ka;
      if (!(len > 0.0)) { g[0] = g[2] = xm[ka]; g[1] = g[3] = ym[ka]; g[4] = g[5] = g[6] = 0.0; g[7] = 1.0; continue; }
      const double nxo = ey / len, nyo = -ex / len;                  // counter-clockwise polygon: outward
      const double mx = xm[ka] + 0.5 * ex, my = ym[ka] + 0.5 * ey;
      const int i0 = box_index(xc, nx, mx), j0 = box_index(yc, ny, my);
      const double hx = xc[i0 + 1] - xc[i0], hy = yc[j0 + 1] - yc[j0];
      const double delta = 1.5 * std::sqrt(hx * hx + hy * hy);
      g[0] = mx + delta * nxo; g[1] = my + delta * nyo;
      g[2] = mx + (2.0 * delta) * nxo; g[3] = my + (2.0 * delta) * nyo;
      g[4] = nxo; g[5] = nyo; g[6] = len; g[7] = delta;
    }
  }
}

// F: 4 per body (pressure x, y; viscous x, y) from the values probed at the segments' points (the ns values at P1,
// then the ns at P2): wall pressure 2 p1 - p2, wall-normal derivative (4 u1 - u2 - 3 u_body) / (2 delta).
// oracle: orc_force_sum
inline void force_sum(int nbodies, const int* off, const double* geo, const double* pu, const double* pv, const double* pp,
                      const double* ub, const double* vb, double Re, double* F) {
  const int ns = nbodies ? off[nbodies] : 0;
  for (int b = 0; b < nbodies; b++) {
    double fpx = 0.0, fpy = 0.0, fvx = 0.0, fvy = 0.0;
    for (int k = off[b]; k < off[b + 1]; k++) {
      const double* g = geo + 8 * (size_t)k;
      const double pw = 2.0 * pp[k] - pp[ns + k];
      const double dudn = ((4.0 * pu[k] - pu[ns + k]) - 3.0 * ub[b]) / (2.0 * g[7]);
      const double dvdn = ((4.0 * pv[k] - pv[ns + k]) - 3.0 * vb[b]) / (2.0 * g[7]);
      fpx = fpx + (-(pw * g[4]) * g[6]);
      fpy = fpy + (-(pw * g[5]) * g[6]);
      fvx = fvx + ((dudn * g[6]) / Re);
      fvy = fvy + ((dvdn * g[6]) / Re);
    }
    F[4 * b] = fpx; F[4 * b + 1] = fpy; F[4 * b + 2] = fvx; F[4 * b + 3] = fvy;
  }
}

// kernels_diag.cu: u, v, p interpolated at n points (device arrays)
cudaError_t launch_probe(const Layout& L, const double* xc, const double* yc, const uint8_t* celltype, const double* u,
                         const double* v, const double* p, int n, const double* px, const double* py, double* ou,
                         double* ov, double* op, cudaStream_t st);

}  // namespace ifx
