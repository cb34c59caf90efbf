/*
 * ifx_oracle_diag.c — CPU definition of the run-time diagnostics ("probes and force coefficients as
 * in the predecessor", test/UTIL_PRE_SIM.f90:172-201).  TEST INFRASTRUCTURE ONLY (see ifx_oracle.h).
 *
 * **PARITY UNPINNED.**  The reference has neither probes nor forces (its BC struct and everything around it is
 * unused, globalVariables.cuh:35-39); this file defines what ifx_probe / ifx_body_forces compute.
 *
 * Probe: bilinear interpolation of the cell-centred u, v, p in the box of cell centres that contains the point —
 * the same box, weights and drop-and-renormalise rule for interior-solid nodes as the ghost-cell image points
 * (orc_ghost_cells); ghost cells take part with their boundary-condition values, which is what makes the
 * interpolation valid right up to the surface.
 * Surface force on a body: two probes per marker segment, P1 and P2 at distances delta and 2 delta (delta = 1.5 cell
 * diagonals) along the outward normal from the segment midpoint M:
 *     F_p = - sum (2 p(P1) - p(P2)) n ds                          (wall pressure by linear extrapolation)
 *     F_v = 1/Re sum (4 u(P1) - u(P2) - 3 u_body)/(2 delta) ds    (second-order one-sided normal derivative)
 * summed over the segments in marker order.
 */
#include "ifx_oracle.h"

#include <math.h>
#include <stdlib.h>

#define ID(i, j, nx) ((i) + (j) * (nx))
#define CT_TYPE(c) ((c) & 3)

static int lower_index(const double* c, int n, double x) {
  int lo = 0, hi = n - 2;
  if (x < c[0]) return 0;
  while (lo < hi) {
    const int mid = (lo + hi + 1) / 2;
    if (c[mid] <= x) lo = mid; else hi = mid - 1;
  }
  return lo;
}

void orc_interp_setup(int nx, int ny, const double* xc, const double* yc, const unsigned char* ct, double x, double y,
                      int* sten, double* w) {
  const int i0 = lower_index(xc, nx, x), j0 = lower_index(yc, ny, y);
  const double a = (x - xc[i0]) / (xc[i0 + 1] - xc[i0]);
  const double b = (y - yc[j0]) / (yc[j0 + 1] - yc[j0]);
  sten[0] = ID(i0, j0, nx); sten[1] = ID(i0 + 1, j0, nx); sten[2] = ID(i0, j0 + 1, nx); sten[3] = ID(i0 + 1, j0 + 1, nx);
  w[0] = (1.0 - a) * (1.0 - b); w[1] = a * (1.0 - b); w[2] = (1.0 - a) * b; w[3] = a * b;
  double W = 0.0;
  for (int m = 0; m < 4; m++) {
    if (CT_TYPE(ct[sten[m]]) == 0) w[m] = 0.0;
    W = W + w[m];
  }
  for (int m = 0; m < 4; m++) w[m] = (W > 0.0) ? w[m] / W : 0.0;
}

static double interp(const double* q, const int* sten, const double* w) {
  double t = w[0] * q[sten[0]];
  t = fma(w[1], q[sten[1]], t);
  t = fma(w[2], q[sten[2]], t);
  t = fma(w[3], q[sten[3]], t);
  return t;
}

void orc_probe(int nx, int ny, const double* xc, const double* yc, const unsigned char* ct, const double* u,
               const double* v, const double* p, int npts, const double* px, const double* py, double* ou, double* ov,
               double* op) {
  for (int k = 0; k < npts; k++) {
    int sten[4]; double w[4];
    orc_interp_setup(nx, ny, xc, yc, ct, px[k], py[k], sten, w);
    ou[k] = interp(u, sten, w); ov[k] = interp(v, sten, w); op[k] = interp(p, sten, w);
  }
}

/* per segment 8 doubles: P1x, P1y, P2x, P2y, nx, ny, ds, delta */
void orc_force_geometry(int nx, int ny, const double* xc, const double* yc, int nseg_total, const int* off, int nbodies,
                        const double* xm, const double* ym, double* geo) {
  (void)nseg_total;
  for (int b = 0; b < nbodies; b++) {
    const int n = off[b + 1] - off[b];
    for (int k = 0; k < n; k++) {
      const int ka = off[b] + k, kb = off[b] + ((k + 1 == n) ? 0 : k + 1);
      const double ex = xm[kb] - xm[ka], ey = ym[kb] - ym[ka];
      const double len = sqrt(ex * ex + ey * ey);
      double* g = geo + 8 * (size_t)ka;
      if (!(len > 0.0)) { g[0] = g[2] = xm[ka]; g[1] = g[3] = ym[ka]; g[4] = g[5] = g[6] = 0.0; g[7] = 1.0; continue; }
      const double nxo = ey / len, nyo = -ex / len;                /* counter-clockwise polygon: outward */
      const double mx = xm[ka] + 0.5 * ex, my = ym[ka] + 0.5 * ey;
      const int i0 = lower_index(xc, nx, mx), j0 = lower_index(yc, ny, my);
      const double hx = xc[i0 + 1] - xc[i0], hy = yc[j0 + 1] - yc[j0];
      const double delta = 1.5 * sqrt(hx * hx + hy * hy);
      g[0] = mx + delta * nxo; g[1] = my + delta * nyo;
      g[2] = mx + (2.0 * delta) * nxo; g[3] = my + (2.0 * delta) * nyo;
      g[4] = nxo; g[5] = nyo; g[6] = len; g[7] = delta;
    }
  }
}

/* F: 4 per body (Fpx, Fpy, Fvx, Fvy) from the probed values; pu, pv, pp hold the ns values at P1 then the ns at P2 */
void orc_force_sum(int nbodies, const int* off, const double* geo, const double* pu, const double* pv, const double* pp,
                   const double* ub, const double* vb, double Re, double* F) {
  const int ns = nbodies ? off[nbodies] : 0;
  for (int b = 0; b < nbodies; b++) {
    double fpx = 0.0, fpy = 0.0, fvx = 0.0, fvy = 0.0;
    for (int k = off[b]; k < off[b + 1]; k++) {
      const double* g = geo + 8 * (size_t)k;
      const double pw = 2.0 * pp[k] - pp[ns + k];
      const double dudn = ((4.0 * pu[k] - pu[ns + k]) - 3.0 * (ub ? ub[b] : 0.0)) / (2.0 * g[7]);
      const double dvdn = ((4.0 * pv[k] - pv[ns + k]) - 3.0 * (vb ? vb[b] : 0.0)) / (2.0 * g[7]);
      fpx = fpx + (-(pw * g[4]) * g[6]);
      fpy = fpy + (-(pw * g[5]) * g[6]);
      fvx = fvx + ((dudn * g[6]) / Re);
      fvy = fvy + ((dvdn * g[6]) / Re);
    }
    F[4 * b] = fpx; F[4 * b + 1] = fpy; F[4 * b + 2] = fvx; F[4 * b + 3] = fvy;
  }
}

void orc_body_forces(int nx, int ny, const double* xc, const double* yc, const unsigned char* ct, double Re, int nbodies,
                     const int* off, const double* xm, const double* ym, const double* ub, const double* vb,
                     const double* u, const double* v, const double* p, double* F) {
  const int ns = nbodies ? off[nbodies] : 0;
  double* geo = (double*)malloc(8 * 8 * (size_t)(ns + 1));
  double* q = (double*)malloc(8 * 10 * (size_t)(ns + 1));
  double *sx = q, *sy = q + 2 * ns, *pu = q + 4 * ns, *pv = q + 6 * ns, *pp = q + 8 * ns;
  orc_force_geometry(nx, ny, xc, yc, ns, off, nbodies, xm, ym, geo);
  for (int k = 0; k < ns; k++) {
    sx[k] = geo[8 * k]; sy[k] = geo[8 * k + 1];
    sx[ns + k] = geo[8 * k + 2]; sy[ns + k] = geo[8 * k + 3];
  }
  orc_probe(nx, ny, xc, yc, ct, u, v, p, 2 * ns, sx, sy, pu, pv, pp);
  orc_force_sum(nbodies, off, geo, pu, pv, pp, ub, vb, Re, F);
  free(geo); free(q);
}
