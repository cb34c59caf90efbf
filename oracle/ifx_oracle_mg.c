/*
 * ifx_oracle_mg.c — CPU definition of the geometric multigrid V-cycle for the pressure-Poisson equation
 * ("better Poisson iterations selected by PPE_Solver ... then geometric multigrid").
 * TEST INFRASTRUCTURE ONLY (see ifx_oracle.h).
 *
 * **PARITY UNPINNED.**  The reference always runs point Jacobi (PPESolver.cu:13-31) and has no multigrid; this
 * file DEFINES what PPE_Solver = 4 computes.  It solves the SAME discrete system as orc_ppe_sweep_general (the
 * reference's coefficients PPESolver.cu:93-99, zero normal gradient on the grid boundary and on closed faces),
 * so the converged pressure agrees with point Jacobi / red-black SOR to the tolerance; only the path differs.
 *
 * Level 0 is the solver's grid.  Level l+1 merges 2x2 cells of level l, down to 2 cells in one direction; where a
 * count is odd the last coarse cell of that direction has a single child (1x2, 2x1 or 1x1 cells).  On the
 * coarse levels the error equation is kept in VOLUME form,
 *       sum_faces G_f (e_nb - e_C) = R_C,
 *   G of a fine face      g_e(i,j) = 2 dy_j / (dx_i + dx_ip1)  (= cE * dx_i * dy_j, symmetric), 0 when the face
 *                         is closed (a non-fluid cell on either side) or on the grid boundary,
 *   G of a coarse face    = a * (sum of the two finer faces it is made of), a = 1/2 [rediscretisation on uniform
 *                         grids; a partly blocked coarse face keeps the open fraction] or 1 (see dir_scale),
 *   R_C                   = sum of the residuals (times cell volume on level 0) of the four children,
 * with piecewise-constant prolongation.  A coarse cell with no open face (sum G = 0) is inactive (e = 0).
 * Smoother on every level: red-black SOR, colour 0 = (i + j) even first; on level 0 it is
 * orc_ppe_sor_halfsweep itself.  All sums are written out in a fixed order so the CUDA kernels
 * (immerseflow_b200/csrc/kernels_mg.cu) reproduce every level bit for bit.
 */
#include "ifx_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#define ID(i, j, nx) ((i) + (j) * (nx))
#define IS_FLUID(c) ((c) == 1)

/* Scale of a coarse face: 1/2 of the sum of its two finer faces (the centre distance doubles) — but only in a
 * direction the point smoother actually smooths, i.e. whose conductance is at least a quarter of the other
 * direction's in this cell; across a weakly coupled direction the error after smoothing is not smooth, the factor
 * 1/2 would over-correct (the cycle diverges on the reference's stretched grids, cell aspect ratios up to 27),
 * so the plain sum (Galerkin for piecewise-constant transfer) is kept there. */
static void dir_scale(int lines, double sx, double sy, double* ax, double* ay) {
  if (lines) { *ax = 0.5; *ay = 0.5; return; }      /* alternating line relaxation smooths both directions everywhere */
  *ax = (sx >= 0.25 * sy) ? 0.5 : 1.0;
  *ay = (sy >= 0.25 * sx) ? 0.5 : 1.0;
}

struct orc_mg {
  int nlevels;               /* including level 0 */
  int ncx[ORC_MG_MAX_LEVELS], ncy[ORC_MG_MAX_LEVELS];     /* cells per level */
  /* levels >= 1: ghost-inclusive arrays (ncx+2) x (ncy+2); GE(I,J) = conductance of the east face of (I,J),
   * GN(I,J) = north face; faces on or outside the grid boundary are 0 */
  double *GE[ORC_MG_MAX_LEVELS], *GN[ORC_MG_MAX_LEVELS], *e[ORC_MG_MAX_LEVELS], *R[ORC_MG_MAX_LEVELS];
  double *cpw[ORC_MG_MAX_LEVELS], *dpw[ORC_MG_MAX_LEVELS];      /* Thomas scratch of the line smoother */
};
orc_mg* orc_mg_create2(int nx, int ny, const double* dx, const double* dy, const unsigned char* celltype, int lines);
void orc_mg_prolong2(int nxl, int nyl, const double* GE, const double* GN, int NX, const double* ec, const double* GEc,
                     const double* GNc, double* e);
void orc_mg_prolong_fine2(int nx, int ny, const unsigned char* ct, int NX, const double* e1, const double* GE1,
                          const double* GN1, double* p);

int orc_mg_plan(int ncx, int ncy, int* lx, int* ly) {
  int n = 1;
  lx[0] = ncx; ly[0] = ncy;
  while (n < ORC_MG_MAX_LEVELS && lx[n - 1] >= 3 && ly[n - 1] >= 3) {
    lx[n] = (lx[n - 1] + 1) / 2; ly[n] = (ly[n - 1] + 1) / 2;      /* odd count: the last coarse cell has one child */
    n++;
  }
  return n;
}

/* red-black iterations on the coarsest level: it is only 2 x 2 .. 3 x 3 cells when the cell counts are powers of
 * two times a small number, but e.g. 200 x 120 stops at 25 x 15, which 32 iterations do not solve */
int orc_mg_ncoarse(int ncx, int ncy) {
  int n = ncx * ncy;
  if (n < 32) n = 32;
  if (n > 2048) n = 2048;
  return n;
}

/* line-relaxation iterations on the coarsest level of PPE_Solver 5: an eighth of its cell count, within [8, 512]
 * (8 for the 2 x 2 .. 7 x 7 levels that even-friendly grids end on; 180 for the 45 x 32 of the shipped 180 x 128 case) */
int orc_mg_ncoarse_lines(int ncx, int ncy) {
  int n = ncx * ncy / 8;
  if (n < 8) n = 8;
  if (n > 512) n = 512;
  return n;
}

/* level-1 conductances from the fine geometry and cell types */
static void build_level1(int lines, int nx, int ny, const double* dx, const double* dy, const unsigned char* ct, int NX,
                         int NY, double* GE, double* GN) {
  memset(GE, 0, sizeof(double) * (size_t)NX * NY);
  memset(GN, 0, sizeof(double) * (size_t)NX * NY);
#pragma omp parallel for
  for (int J = 1; J < NY - 1; J++)
    for (int I = 1; I < NX - 1; I++) {
      const int i = 2 * I, j = 2 * J;          /* children: columns i-1, i; rows j-1, j */
      double ge[2] = {0.0, 0.0}, gn[2] = {0.0, 0.0};
      for (int k = 0; k < 2; k++) {
        const int jj = j - 1 + k, ii = i - 1 + k;          /* the second child row / column may not exist (odd count) */
        if (i <= nx - 3 && jj <= ny - 2 && IS_FLUID(ct[ID(i, jj, nx)]) && IS_FLUID(ct[ID(i + 1, jj, nx)]))
          ge[k] = (2.0 * dy[ID(i, jj, nx)]) / (dx[ID(i, jj, nx)] + dx[ID(i + 1, jj, nx)]);
        if (j <= ny - 3 && ii <= nx - 2 && IS_FLUID(ct[ID(ii, j, nx)]) && IS_FLUID(ct[ID(ii, j + 1, nx)]))
          gn[k] = (2.0 * dx[ID(ii, j, nx)]) / (dy[ID(ii, j, nx)] + dy[ID(ii, j + 1, nx)]);
      }
      double ax, ay; dir_scale(lines, ge[0] + ge[1], gn[0] + gn[1], &ax, &ay);
      GE[ID(I, J, NX)] = ax * (ge[0] + ge[1]);
      GN[ID(I, J, NX)] = ay * (gn[0] + gn[1]);
    }
}

static void build_coarser(int lines, int nxl, const double* GEf, const double* GNf, int NX, int NY, double* GE, double* GN) {
  memset(GE, 0, sizeof(double) * (size_t)NX * NY);
  memset(GN, 0, sizeof(double) * (size_t)NX * NY);
#pragma omp parallel for
  for (int J = 1; J < NY - 1; J++)
    for (int I = 1; I < NX - 1; I++) {
      const int i = 2 * I, j = 2 * J;
      const double sx = GEf[ID(i, j - 1, nxl)] + GEf[ID(i, j, nxl)], sy = GNf[ID(i - 1, j, nxl)] + GNf[ID(i, j, nxl)];
      double ax, ay; dir_scale(lines, sx, sy, &ax, &ay);
      GE[ID(I, J, NX)] = ax * sx;
      GN[ID(I, J, NX)] = ay * sy;
    }
}

orc_mg* orc_mg_create(int nx, int ny, const double* dx, const double* dy, const unsigned char* celltype) {
  return orc_mg_create2(nx, ny, dx, dy, celltype, 0);
}
/* lines = 1: hierarchy for the line-smoothed cycle (coarse faces always scaled by 1/2) + line scratch per level */
orc_mg* orc_mg_create2(int nx, int ny, const double* dx, const double* dy, const unsigned char* celltype, int lines) {
  orc_mg* m = (orc_mg*)calloc(1, sizeof(orc_mg));
  m->nlevels = orc_mg_plan(nx - 2, ny - 2, m->ncx, m->ncy);
  for (int l = 1; l < m->nlevels; l++) {
    const size_t n = (size_t)(m->ncx[l] + 2) * (m->ncy[l] + 2);
    m->GE[l] = (double*)calloc(n, 8); m->GN[l] = (double*)calloc(n, 8);
    m->e[l] = (double*)calloc(n, 8); m->R[l] = (double*)calloc(n, 8);
    m->cpw[l] = (double*)calloc(n, 8); m->dpw[l] = (double*)calloc(n, 8);
    if (l == 1) build_level1(lines, nx, ny, dx, dy, celltype, m->ncx[1] + 2, m->ncy[1] + 2, m->GE[1], m->GN[1]);
    else build_coarser(lines, m->ncx[l - 1] + 2, m->GE[l - 1], m->GN[l - 1], m->ncx[l] + 2, m->ncy[l] + 2, m->GE[l], m->GN[l]);
  }
  return m;
}

void orc_mg_destroy(orc_mg* m) {
  if (!m) return;
  for (int l = 1; l < m->nlevels; l++) { free(m->GE[l]); free(m->GN[l]); free(m->e[l]); free(m->R[l]); free(m->cpw[l]); free(m->dpw[l]); }
  free(m);
}

int orc_mg_levels(const orc_mg* m) { return m->nlevels; }
/* which: 0 GE, 1 GN, 2 e, 3 R; copies the ghost-inclusive (ncx+2)(ncy+2) array of level l >= 1 */
int orc_mg_get(const orc_mg* m, int l, int which, double* out, int* ncx, int* ncy) {
  if (l < 0 || l >= m->nlevels) return -1;
  if (ncx) *ncx = m->ncx[l];
  if (ncy) *ncy = m->ncy[l];
  if (l == 0 || !out) return 0;
  const double* src = which == 0 ? m->GE[l] : which == 1 ? m->GN[l] : which == 2 ? m->e[l] : m->R[l];
  memcpy(out, src, 8 * (size_t)(m->ncx[l] + 2) * (m->ncy[l] + 2));
  return 0;
}

/* R1 = sum over the four children of (rhs - A p) * dx_i * dy_j, fluid children only.  (A p) exactly as the
 * residual of orc_ppe_sweep_general. */
void orc_mg_restrict_fine(int nx, int ny, const double* dx, const double* dy, const double* cP, const double* cxm,
                          const double* cxp, const double* cym, const double* cyp, const unsigned char* ct,
                          const double* rhs, const double* p, int NX, int NY, double* R1) {
#pragma omp parallel for
  for (int J = 1; J < NY - 1; J++)
    for (int I = 1; I < NX - 1; I++) {
      double r[4];
      for (int k = 0; k < 4; k++) {
        const int i = 2 * I - 1 + (k & 1), j = 2 * J - 1 + (k >> 1);
        const int id = ID(i, j, nx);
        r[k] = 0.0;
        if (i > nx - 2 || j > ny - 2 || !IS_FLUID(ct[id])) continue;      /* child beyond the grid (odd count) */
        const double pc = p[id];
        const double pw = (i == 1 || !IS_FLUID(ct[id - 1])) ? pc : p[id - 1];
        const double pe = (i == nx - 2 || !IS_FLUID(ct[id + 1])) ? pc : p[id + 1];
        const double ps = (j == 1 || !IS_FLUID(ct[id - nx])) ? pc : p[id - nx];
        const double pn = (j == ny - 2 || !IS_FLUID(ct[id + nx])) ? pc : p[id + nx];
        double q = pe * cxp[id];
        q = fma(pc, cP[id], q);
        q = fma(pw, cxm[id], q);
        q = fma(pn, cyp[id], q);
        q = fma(ps, cym[id], q);
        r[k] = (rhs[id] - q) * (dx[id] * dy[id]);
      }
      R1[ID(I, J, NX)] = (r[0] + r[1]) + (r[2] + r[3]);
    }
}

/* one colour of red-black SOR on a coarse level, in place (cells of one colour are not neighbours) */
void orc_mg_smooth(int NX, int NY, const double* GE, const double* GN, const double* R, int colour, double omega,
                   double* e) {
#pragma omp parallel for
  for (int J = 1; J < NY - 1; J++)
    for (int I = 1; I < NX - 1; I++) {
      if ((I + J + colour) & 1) continue;
      const int id = ID(I, J, NX);
      const double ge = GE[id], gw = GE[id - 1], gn = GN[id], gs = GN[id - NX];
      const double D = (ge + gw) + (gn + gs);
      if (!(D > 0.0)) continue;
      double t = ge * e[id + 1];
      t = fma(gw, e[id - 1], t);
      t = fma(gn, e[id + NX], t);
      t = fma(gs, e[id - NX], t);
      const double ej = (t - R[id]) / D;
      e[id] = e[id] + omega * (ej - e[id]);
    }
}

/* R_{l+1} = sum over the four children of (R - (sum G e_nb - D e)) on level l */
void orc_mg_restrict(int nxl, const double* GE, const double* GN, const double* R, const double* e, int NX, int NY,
                     double* Rc) {
#pragma omp parallel for
  for (int J = 1; J < NY - 1; J++)
    for (int I = 1; I < NX - 1; I++) {
      double r[4];
      for (int k = 0; k < 4; k++) {
        const int id = ID(2 * I - 1 + (k & 1), 2 * J - 1 + (k >> 1), nxl);
        const double ge = GE[id], gw = GE[id - 1], gn = GN[id], gs = GN[id - nxl];
        const double D = (ge + gw) + (gn + gs);
        r[k] = 0.0;
        if (!(D > 0.0)) continue;
        double t = ge * e[id + 1];
        t = fma(gw, e[id - 1], t);
        t = fma(gn, e[id + nxl], t);
        t = fma(gs, e[id - nxl], t);
        r[k] = R[id] - fma(-D, e[id], t);
      }
      Rc[ID(I, J, NX)] = (r[0] + r[1]) + (r[2] + r[3]);
    }
}

/* Bilinear prolongation (the line-smoothed cycle; the point-smoothed one converges faster with the piecewise-constant
 * transfer, 0.13 against 0.18 per cycle, the line-smoothed one with this, 0.17 against 0.36).  Value of the coarse
 * correction at fine cell (i, j): parent 9/16, the two nearer coarse neighbours 3/16 each, the diagonal one 1/16; a
 * neighbour that is not connected to the parent through open coarse faces (inactive cell, body in between, grid
 * boundary) is replaced by the parent. */
static double prolong_value(int i, int j, int NX, const double* ec, const double* GEc, const double* GNc) {
  const int I = (i + 1) / 2, J = (j + 1) / 2;
  const int di = (i & 1) ? -1 : 1, dj = (j & 1) ? -1 : 1;
  const int P = ID(I, J, NX), A = P + di, B = P + dj * NX, C = A + dj * NX;
  const double eP = ec[P];
  /* conductance of the face between two horizontally / vertically adjacent coarse cells */
  const double gPA = GEc[di > 0 ? P : A], gPB = GNc[dj > 0 ? P : B];
  const double gAC = GNc[dj > 0 ? A : C], gBC = GEc[di > 0 ? B : C];
  const double eA = (gPA > 0.0) ? ec[A] : eP;
  const double eB = (gPB > 0.0) ? ec[B] : eP;
  const double eC = ((gPA > 0.0 && gAC > 0.0) || (gPB > 0.0 && gBC > 0.0)) ? ec[C] : eP;
  return 0.0625 * ((9.0 * eP + 3.0 * eA) + (3.0 * eB + eC));
}

/* e_l += e_{l+1}(parent) on active cells of level l; prolong2 with GEc != NULL: bilinear (see prolong_value) */
void orc_mg_prolong(int nxl, int nyl, const double* GE, const double* GN, int NX, const double* ec, double* e) {
  orc_mg_prolong2(nxl, nyl, GE, GN, NX, ec, NULL, NULL, e);
}
void orc_mg_prolong2(int nxl, int nyl, const double* GE, const double* GN, int NX, const double* ec, const double* GEc,
                     const double* GNc, double* e) {
#pragma omp parallel for
  for (int j = 1; j < nyl - 1; j++)
    for (int i = 1; i < nxl - 1; i++) {
      const int id = ID(i, j, nxl);
      const double D = (GE[id] + GE[id - 1]) + (GN[id] + GN[id - nxl]);
      if (!(D > 0.0)) continue;
      e[id] = e[id] + (GEc ? prolong_value(i, j, NX, ec, GEc, GNc) : ec[ID((i + 1) / 2, (j + 1) / 2, NX)]);
    }
}

/* p += e_1(parent) on fluid cells */
void orc_mg_prolong_fine(int nx, int ny, const unsigned char* ct, int NX, const double* e1, double* p) {
  orc_mg_prolong_fine2(nx, ny, ct, NX, e1, NULL, NULL, p);
}
void orc_mg_prolong_fine2(int nx, int ny, const unsigned char* ct, int NX, const double* e1, const double* GE1,
                          const double* GN1, double* p) {
#pragma omp parallel for
  for (int j = 1; j < ny - 1; j++)
    for (int i = 1; i < nx - 1; i++) {
      const int id = ID(i, j, nx);
      if (IS_FLUID(ct[id])) p[id] = p[id] + (GE1 ? prolong_value(i, j, NX, e1, GE1, GN1) : e1[ID((i + 1) / 2, (j + 1) / 2, NX)]);
    }
}

/* the coarse part of a V-cycle: R[1] is set; on return e[1] holds the correction for level 0 */
void orc_mg_coarse_cycle(orc_mg* m, int nu1, int nu2, int ncoarse, double omega) {
  const int L = m->nlevels;
  for (int l = 1; l < L; l++) {
    const int NX = m->ncx[l] + 2, NY = m->ncy[l] + 2;
    memset(m->e[l], 0, 8 * (size_t)NX * NY);
    const int its = (l == L - 1) ? (ncoarse > 0 ? ncoarse : orc_mg_ncoarse(m->ncx[l], m->ncy[l])) : nu1;
    for (int k = 0; k < its; k++) {
      orc_mg_smooth(NX, NY, m->GE[l], m->GN[l], m->R[l], 0, omega, m->e[l]);
      orc_mg_smooth(NX, NY, m->GE[l], m->GN[l], m->R[l], 1, omega, m->e[l]);
    }
    if (l < L - 1)
      orc_mg_restrict(NX, m->GE[l], m->GN[l], m->R[l], m->e[l], m->ncx[l + 1] + 2, m->ncy[l + 1] + 2, m->R[l + 1]);
  }
  for (int l = L - 2; l >= 1; l--) {
    const int NX = m->ncx[l] + 2, NY = m->ncy[l] + 2;
    orc_mg_prolong(NX, NY, m->GE[l], m->GN[l], m->ncx[l + 1] + 2, m->e[l + 1], m->e[l]);
    for (int k = 0; k < nu2; k++) {
      orc_mg_smooth(NX, NY, m->GE[l], m->GN[l], m->R[l], 0, omega, m->e[l]);
      orc_mg_smooth(NX, NY, m->GE[l], m->GN[l], m->R[l], 1, omega, m->e[l]);
    }
  }
}

/* One V(nu1, nu2) cycle on p (in place; pT is scratch of the same size whose ring is never read). */
void orc_mg_vcycle(orc_mg* m, int nx, int ny, const double* dx, const double* dy, const double* cP, const double* cxm,
                   const double* cxp, const double* cym, const double* cyp, const unsigned char* ct, const double* rhs,
                   int nu1, int nu2, int ncoarse, double omega, double* p, double* pT) {
  for (int k = 0; k < nu1; k++) {
    orc_ppe_sor_halfsweep(nx, ny, cP, cxm, cxp, cym, cyp, ct, rhs, 0, omega, p, pT);
    orc_ppe_sor_halfsweep(nx, ny, cP, cxm, cxp, cym, cyp, ct, rhs, 1, omega, pT, p);
  }
  orc_mg_restrict_fine(nx, ny, dx, dy, cP, cxm, cxp, cym, cyp, ct, rhs, p, m->ncx[1] + 2, m->ncy[1] + 2, m->R[1]);
  orc_mg_coarse_cycle(m, nu1, nu2, ncoarse, omega);
  orc_mg_prolong_fine(nx, ny, ct, m->ncx[1] + 2, m->e[1], p);
  for (int k = 0; k < nu2; k++) {
    orc_ppe_sor_halfsweep(nx, ny, cP, cxm, cxp, cym, cyp, ct, rhs, 0, omega, p, pT);
    orc_ppe_sor_halfsweep(nx, ny, cP, cxm, cxp, cym, cyp, ct, rhs, 1, omega, pT, p);
  }
}

/* =================================================================================================
 * Zebra line relaxation ("line-SOR via batched Thomas solves"; the input file's own
 * "PPE_Solver ... 2. Line SOR", parsed at main.cu:42 and never used by the reference).  UNPINNED.
 *
 * One pass relaxes every second grid line (lines of index parity `parity`) simultaneously: along the line the
 * Poisson equation of orc_ppe_sweep_general is solved exactly for all its cells at once (tridiagonal system,
 * Thomas algorithm in increasing index order), the two neighbouring lines — of the other parity, untouched by the
 * pass — on the right-hand side; then p <- p + omega (p* - p).  dir 0: lines along x (rows j), dir 1: lines along y
 * (columns i).  Closed faces (grid boundary, non-fluid neighbour) drop out of the stencil, which moves their
 * coefficient onto the diagonal (p_nb := p_C); non-fluid cells are identity rows, so a line falls into
 * independent fluid segments by itself.  Row of cell k of the line:   lo_k x_{k-1} + dg_k x_k + up_k x_{k+1} = d_k.
 * Elimination with one reciprocal per cell:
 *     inv_k = 1 / (dg_k - lo_k cp_{k-1}),  cp_k = up_k inv_k,  dp_k = (d_k - lo_k dp_{k-1}) inv_k,
 *     x_k = dp_k - cp_k x_{k+1}.
 * A cell whose pivot is not negative (an isolated fluid cell: dg = 0) keeps its value.
 * cpw / dpw: scratch of nx*ny doubles each (every line uses its own cells' slots).
 * An iteration of PPE_Solver 2 is four passes: x-lines even, x-lines odd, y-lines even, y-lines odd.
 * ================================================================================================= */
void orc_ppe_line_pass(int nx, int ny, const double* cP, const double* cxm, const double* cxp, const double* cym,
                       const double* cyp, const unsigned char* ct, const double* rhs, int dir, int parity, double omega,
                       double* p, double* cpw, double* dpw) {
  const int nline = dir == 0 ? ny : nx, len = dir == 0 ? nx : ny;
  const int sk = dir == 0 ? 1 : nx;            /* stride along the line */
  const int sl = dir == 0 ? nx : 1;            /* stride across lines */
#pragma omp parallel for
  for (int l = 1; l < nline - 1; l++) {
    if ((l & 1) != parity) continue;
    double cprev = 0.0, dprev = 0.0;
    for (int k = 1; k < len - 1; k++) {
      const int id = l * sl + k * sk;
      double lo = 0.0, up = 0.0, dg = 1.0, d = p[id];
      if (IS_FLUID(ct[id])) {
        const int i = id % nx, j = id / nx;
        const int oW = !(i == 1 || !IS_FLUID(ct[id - 1])), oE = !(i == nx - 2 || !IS_FLUID(ct[id + 1]));
        const int oS = !(j == 1 || !IS_FLUID(ct[id - nx])), oN = !(j == ny - 2 || !IS_FLUID(ct[id + nx]));
        /* diagonal = cP + coefficients of the closed faces (p_nb := p_C), added in the order W, E, S, N */
        dg = cP[id];
        if (!oW) dg = dg + cxm[id];
        if (!oE) dg = dg + cxp[id];
        if (!oS) dg = dg + cym[id];
        if (!oN) dg = dg + cyp[id];
        d = rhs[id];
        if (dir == 0) {
          lo = oW ? cxm[id] : 0.0; up = oE ? cxp[id] : 0.0;
          if (oN) d = fma(-cyp[id], p[id + nx], d);
          if (oS) d = fma(-cym[id], p[id - nx], d);
        } else {
          lo = oS ? cym[id] : 0.0; up = oN ? cyp[id] : 0.0;
          if (oE) d = fma(-cxp[id], p[id + 1], d);
          if (oW) d = fma(-cxm[id], p[id - 1], d);
        }
        const double piv = fma(-lo, cprev, dg);
        if (!(piv < 0.0)) { lo = 0.0; up = 0.0; dg = 1.0; d = p[id]; }      /* isolated: identity row */
      }
      const double inv = 1.0 / fma(-lo, cprev, dg);
      cprev = up * inv;
      dprev = fma(-lo, dprev, d) * inv;
      cpw[id] = cprev; dpw[id] = dprev;
    }
    double xnext = 0.0;
    for (int k = len - 2; k >= 1; k--) {
      const int id = l * sl + k * sk;
      const double x = fma(-cpw[id], xnext, dpw[id]);
      xnext = x;
      if (IS_FLUID(ct[id])) p[id] = p[id] + omega * (x - p[id]);
    }
  }
}

void orc_ppe_line_iteration(int nx, int ny, const double* cP, const double* cxm, const double* cxp, const double* cym,
                            const double* cyp, const unsigned char* ct, const double* rhs, double omega, double* p,
                            double* cpw, double* dpw) {
  for (int dir = 0; dir < 2; dir++)
    for (int parity = 0; parity < 2; parity++)
      orc_ppe_line_pass(nx, ny, cP, cxm, cxp, cym, cyp, ct, rhs, dir, parity, omega, p, cpw, dpw);
}

/* the same on a coarse level: row of cell k:  G_lo e_{k-1} - D e_k + G_up e_{k+1} = R - (other direction's G e_nb) */
void orc_mg_line_pass(int NX, int NY, const double* GE, const double* GN, const double* R, int dir, int parity,
                      double omega, double* e, double* cpw, double* dpw) {
  const int nline = dir == 0 ? NY : NX, len = dir == 0 ? NX : NY;
  const int sk = dir == 0 ? 1 : NX, sl = dir == 0 ? NX : 1;
#pragma omp parallel for
  for (int l = 1; l < nline - 1; l++) {
    if ((l & 1) != parity) continue;
    double cprev = 0.0, dprev = 0.0;
    for (int k = 1; k < len - 1; k++) {
      const int id = l * sl + k * sk;
      const double ge = GE[id], gw = GE[id - 1], gn = GN[id], gs = GN[id - NX];
      const double D = (ge + gw) + (gn + gs);
      double lo = 0.0, up = 0.0, dg = 1.0, d = e[id];
      if (D > 0.0) {
        dg = -D;
        d = R[id];
        if (dir == 0) { lo = gw; up = ge; d = fma(-gn, e[id + NX], d); d = fma(-gs, e[id - NX], d); }
        else { lo = gs; up = gn; d = fma(-ge, e[id + 1], d); d = fma(-gw, e[id - 1], d); }
        const double piv = fma(-lo, cprev, dg);
        if (!(piv < 0.0)) { lo = 0.0; up = 0.0; dg = 1.0; d = e[id]; }
      }
      const double inv = 1.0 / fma(-lo, cprev, dg);
      cprev = up * inv;
      dprev = fma(-lo, dprev, d) * inv;
      cpw[id] = cprev; dpw[id] = dprev;
    }
    double xnext = 0.0;
    for (int k = len - 2; k >= 1; k--) {
      const int id = l * sl + k * sk;
      const double x = fma(-cpw[id], xnext, dpw[id]);
      xnext = x;
      const double D = (GE[id] + GE[id - 1]) + (GN[id] + GN[id - NX]);
      if (D > 0.0) e[id] = e[id] + omega * (x - e[id]);
    }
  }
}

/* V(nu1, nu2) cycle smoothed by alternating zebra line relaxation on every level (PPE_Solver 5).  Hierarchy from
 * orc_mg_create2(..., lines = 1).  cpw / dpw: fine-level scratch of nx*ny doubles each. */
static void mg_lines(orc_mg* m, int l, int its, double omega) {
  const int NX = m->ncx[l] + 2, NY = m->ncy[l] + 2;
  for (int k = 0; k < its; k++)
    for (int dir = 0; dir < 2; dir++)
      for (int parity = 0; parity < 2; parity++)
        orc_mg_line_pass(NX, NY, m->GE[l], m->GN[l], m->R[l], dir, parity, omega, m->e[l], m->cpw[l], m->dpw[l]);
}

void orc_mg_vcycle_lines(orc_mg* m, int nx, int ny, const double* dx, const double* dy, const double* cP, const double* cxm,
                         const double* cxp, const double* cym, const double* cyp, const unsigned char* ct, const double* rhs,
                         int nu1, int nu2, int ncoarse, double omega, double* p, double* cpw, double* dpw) {
  const int L = m->nlevels;
  for (int k = 0; k < nu1; k++) orc_ppe_line_iteration(nx, ny, cP, cxm, cxp, cym, cyp, ct, rhs, omega, p, cpw, dpw);
  orc_mg_restrict_fine(nx, ny, dx, dy, cP, cxm, cxp, cym, cyp, ct, rhs, p, m->ncx[1] + 2, m->ncy[1] + 2, m->R[1]);
  for (int l = 1; l < L; l++) {
    const int NX = m->ncx[l] + 2, NY = m->ncy[l] + 2;
    memset(m->e[l], 0, 8 * (size_t)NX * NY);
    mg_lines(m, l, (l == L - 1) ? (ncoarse > 0 ? ncoarse : orc_mg_ncoarse_lines(m->ncx[l], m->ncy[l])) : nu1, omega);
    if (l < L - 1)
      orc_mg_restrict(NX, m->GE[l], m->GN[l], m->R[l], m->e[l], m->ncx[l + 1] + 2, m->ncy[l + 1] + 2, m->R[l + 1]);
  }
  for (int l = L - 2; l >= 1; l--) {
    orc_mg_prolong2(m->ncx[l] + 2, m->ncy[l] + 2, m->GE[l], m->GN[l], m->ncx[l + 1] + 2, m->e[l + 1], m->GE[l + 1],
                    m->GN[l + 1], m->e[l]);
    mg_lines(m, l, nu2, omega);
  }
  orc_mg_prolong_fine2(nx, ny, ct, m->ncx[1] + 2, m->e[1], m->GE[1], m->GN[1], p);
  for (int k = 0; k < nu2; k++) orc_ppe_line_iteration(nx, ny, cP, cxm, cxp, cym, cyp, ct, rhs, omega, p, cpw, dpw);
}
