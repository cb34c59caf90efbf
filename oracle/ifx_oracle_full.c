/*
 * ifx_oracle_full.c — CPU definition of the stages the reference does NOT implement.
 * TEST INFRASTRUCTURE ONLY (see ifx_oracle.h).
 *
 * **PARITY UNPINNED.**  The reference has no code for any of this: its PPE source kernel is a
 * copy-paste stub (PPESolver.cu:106-135), its projection file is empty (AD_PPE_Correction.cu:1-12), its iBlank
 * kernel assigns 1.0 to every cell (preSim.cu:123-133) and there is no ghost-cell code at all.  This file
 * therefore DEFINES the semantics of IFX_COMPAT_FULL, following north_star and the sharp-interface
 * ghost-cell method of Mittal et al. (J. Comput. Phys. 227, 2008) that the reference README cites; the CUDA
 * path is tested bit-for-bit against it, but there is nothing of the reference's to pin either of them to.
 *
 * Conventions shared with the pinned part: reference layout id = i + j*nx, 2-D dx/dy arrays, cell types as
 * unsigned char: low two bits 0 = solid, 1 = fluid, 2 = ghost cell, upper six bits = index of the owning body for
 * non-fluid cells (at most 63 bodies); the reference's double iBlank is 1.0 for fluid, else 0.0.
 *
 * Pressure treatment (why the ghost-cell value of p is diagnostic only): a face between a fluid cell and a
 * non-fluid cell is CLOSED — its velocity is the body's, and the pressure gradient across it is zero (the
 * neighbour's p is replaced by the cell's own, exactly like the homogeneous-Neumann ring).  The discrete Poisson
 * operator is then the volume-symmetric div-grad on the fluid cells, whose only null vector is the constant and
 * whose compatibility condition is the net flux through the grid boundary — so point-Jacobi converges.  Closing
 * the stencil through interpolated ghost-cell pressures instead makes the singular system incompatible (tried:
 * the residual stalls).  This is also what the Fortran predecessor did (test/UTIL_PRE_SIM.f90:154-164 masks face
 * velocities with iblank_fcu/iblank_fcv).  p at ghost cells is still filled from the image point (Neumann) after
 * the solve, for output.
 * All geometry arithmetic is plain IEEE +,-,*,/ in a fixed order (no FMA; -ffp-contract=off here, -fmad=false
 * on the device), so integer maps AND weights agree bit for bit between CPU and GPU.
 */
#include "ifx_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#define ID(i, j, nx) ((i) + (j) * (nx))
#define CT_TYPE(c) ((c) & 3)
#define CT_BODY(c) ((c) >> 2)
#define IS_FLUID(c) ((c) == 1)

/* -----------------------------------------------------------------------------------------------
 * a16 — cell classification.  A cell is SOLID when its centre lies inside any body polygon
 * (crossing-number test, half-open edge rule), FLUID otherwise; the ghost ring of the grid is FLUID.
 * Bodies: polygon b has markers [off[b], off[b+1]), counter-clockwise.
 * --------------------------------------------------------------------------------------------- */
static int point_in_polygon(double x, double y, const double* xm, const double* ym, int n) {
  int inside = 0;
  for (int k = 0; k < n; k++) {
    const int k2 = (k + 1 == n) ? 0 : k + 1;
    const double xa = xm[k], ya = ym[k], xb = xm[k2], yb = ym[k2];
    if ((ya > y) != (yb > y)) {
      const double xi = xa + (y - ya) * (xb - xa) / (yb - ya);
      if (x < xi) inside = !inside;
    }
  }
  return inside;
}

void orc_iblank_classify(int nx, int ny, const double* xc, const double* yc, int nbodies, const int* off,
                         const double* xm, const double* ym, unsigned char* celltype, int* body_of) {
  /* bounding boxes: a centre outside the box of a polygon cannot be inside it, so skipping the crossing test
   * there changes no answer (keeps the CPU baseline honest: O(cells x markers) only inside the boxes) */
  double* bbox = (double*)malloc(sizeof(double) * 4 * (size_t)(nbodies + 1));
  for (int b = 0; b < nbodies; b++) {
    double x0 = xm[off[b]], x1 = x0, y0 = ym[off[b]], y1 = y0;
    for (int k = off[b]; k < off[b + 1]; k++) {
      if (xm[k] < x0) x0 = xm[k];
      if (xm[k] > x1) x1 = xm[k];
      if (ym[k] < y0) y0 = ym[k];
      if (ym[k] > y1) y1 = ym[k];
    }
    bbox[4 * b] = x0; bbox[4 * b + 1] = x1; bbox[4 * b + 2] = y0; bbox[4 * b + 3] = y1;
  }
#pragma omp parallel for
  for (int id = 0; id < nx * ny; id++) {
    const int i = id % nx, j = id / nx;
    celltype[id] = 1;
    if (body_of) body_of[id] = -1;
    if (i > 0 && i < nx - 1 && j > 0 && j < ny - 1) {
      for (int b = 0; b < nbodies; b++) {
        if (xc[i] < bbox[4 * b] || xc[i] > bbox[4 * b + 1] || yc[j] < bbox[4 * b + 2] || yc[j] > bbox[4 * b + 3]) continue;
        if (point_in_polygon(xc[i], yc[j], xm + off[b], ym + off[b], off[b + 1] - off[b])) {
          celltype[id] = (unsigned char)(b << 2);
          if (body_of) body_of[id] = b;
          break;
        }
      }
    }
  }
  free(bbox);
  /* ghost cells: solid with a fluid 4-neighbour (second pass: needs the complete solid/fluid map) */
  unsigned char* tmp = (unsigned char*)malloc((size_t)nx * ny);
  memcpy(tmp, celltype, (size_t)nx * ny);
#pragma omp parallel for
  for (int id = 0; id < nx * ny; id++) {
    const int i = id % nx, j = id / nx;
    if (i > 0 && i < nx - 1 && j > 0 && j < ny - 1 && tmp[id] != 1) {
      if (tmp[id - 1] == 1 || tmp[id + 1] == 1 || tmp[id - nx] == 1 || tmp[id + nx] == 1) celltype[id] = tmp[id] | 2;
    }
  }
  free(tmp);
}

/* -----------------------------------------------------------------------------------------------
 * a17 — ghost-cell records.  For every ghost cell in increasing id order:
 *   BI  = closest point of the owning body's boundary to the cell centre (first minimum in edge order),
 *   IP  = GC + 2 (BI - GC),
 *   (i0, j0) = lower-left node of the cell-centre box containing IP,
 *   bilinear weights; solid non-ghost nodes are dropped and the rest renormalised; the ghost cell itself,
 *   if it is a node, is eliminated algebraically:
 *     Dirichlet  phi_GC = cd*phi_BI + sum_m wd[m]*phi_m,  cd = 2/(1+ws), wd[m] = -w[m]/(1+ws)
 *     Neumann    phi_GC =             sum_m wn[m]*phi_m,  wn[m] = w[m]/(1-ws)
 * Outputs per ghost cell g: cell[g] (id), stencil[4g..] (ids), wd[5g..] = {wd0..3, cd}, wn[4g..],
 * bi[2g..], ip[2g..], body[g].  Returns the ghost-cell count (fills at most `capacity`).
 * --------------------------------------------------------------------------------------------- */
static void closest_on_polygon(double x, double y, const double* xm, const double* ym, int n, double* bx, double* by) {
  double best = INFINITY;
  *bx = x; *by = y;
  for (int k = 0; k < n; k++) {
    const int k2 = (k + 1 == n) ? 0 : k + 1;
    const double ax = xm[k], ay = ym[k], ex = xm[k2] - ax, ey = ym[k2] - ay;
    const double len2 = ex * ex + ey * ey;
    double t = 0.0;
    if (len2 > 0.0) {
      t = ((x - ax) * ex + (y - ay) * ey) / len2;
      if (t < 0.0) t = 0.0;
      if (t > 1.0) t = 1.0;
    }
    const double px = ax + t * ex, py = ay + t * ey;
    const double d2 = (x - px) * (x - px) + (y - py) * (y - py);
    if (d2 < best) { best = d2; *bx = px; *by = py; }
  }
}

static int lower_index(const double* c, int n, double x) {
  /* largest i in [0, n-2] with c[i] <= x (clamped) */
  int lo = 0, hi = n - 2;
  if (x < c[0]) return 0;
  while (lo < hi) {
    const int mid = (lo + hi + 1) / 2;
    if (c[mid] <= x) lo = mid; else hi = mid - 1;
  }
  return lo;
}

int orc_ghost_cells(int nx, int ny, const double* xc, const double* yc, int nbodies, const int* off,
                    const double* xm, const double* ym, const unsigned char* celltype, const int* body_of,
                    int capacity, int* cell, int* stencil, double* wd, double* wn, double* bi, double* ip, int* body) {
  (void)nbodies;
  int g = 0;
  for (int id = 0; id < nx * ny; id++) {
    if (CT_TYPE(celltype[id]) != 2) continue;
    if (g < capacity) {
      const int i = id % nx, j = id / nx;
      const int b = body_of[id];
      const double xg = xc[i], yg = yc[j];
      double bx, by;
      closest_on_polygon(xg, yg, xm + off[b], ym + off[b], off[b + 1] - off[b], &bx, &by);
      const double xi = xg + 2.0 * (bx - xg), yi = yg + 2.0 * (by - yg);
      const int i0 = lower_index(xc, nx, xi), j0 = lower_index(yc, ny, yi);
      const double a = (xi - xc[i0]) / (xc[i0 + 1] - xc[i0]);
      const double bb = (yi - yc[j0]) / (yc[j0 + 1] - yc[j0]);
      const int nid[4] = {ID(i0, j0, nx), ID(i0 + 1, j0, nx), ID(i0, j0 + 1, nx), ID(i0 + 1, j0 + 1, nx)};
      double w[4] = {(1.0 - a) * (1.0 - bb), a * (1.0 - bb), (1.0 - a) * bb, a * bb};
      double W = 0.0, ws = 0.0;
      for (int m = 0; m < 4; m++) {
        if (nid[m] != id && CT_TYPE(celltype[nid[m]]) == 0) w[m] = 0.0;      /* interior solid node: dropped */
        W = W + w[m];
      }
      int kept = 0;
      for (int m = 0; m < 4; m++) {
        w[m] = (W > 0.0) ? w[m] / W : 0.0;
        if (nid[m] == id) { ws = w[m]; w[m] = 0.0; }
        else if (w[m] != 0.0) kept++;
      }
      cell[g] = id;
      body[g] = b;
      bi[2 * g] = bx; bi[2 * g + 1] = by;
      ip[2 * g] = xi; ip[2 * g + 1] = yi;
      for (int m = 0; m < 4; m++) stencil[4 * g + m] = nid[m];
      if (kept == 0 || (1.0 - ws) < 1e-12) {       /* degenerate: GC centre (numerically) on the surface */
        for (int m = 0; m < 4; m++) { wd[5 * g + m] = 0.0; wn[4 * g + m] = 0.0; }
        wd[5 * g + 4] = 1.0;
      } else {
        for (int m = 0; m < 4; m++) {
          wd[5 * g + m] = -w[m] / (1.0 + ws);
          wn[4 * g + m] = w[m] / (1.0 - ws);
        }
        wd[5 * g + 4] = 2.0 / (1.0 + ws);
      }
    }
    g++;
  }
  return g;
}

/* value of one ghost cell from field q (any layout with index map `sten`) */
static double gc_dirichlet(const double* q, const int* sten, const double* wd, double phi_bi) {
  double t = wd[4] * phi_bi;
  t = fma(wd[0], q[sten[0]], t);
  t = fma(wd[1], q[sten[1]], t);
  t = fma(wd[2], q[sten[2]], t);
  t = fma(wd[3], q[sten[3]], t);
  return t;
}
static double gc_neumann(const double* q, const int* sten, const double* wn) {
  double t = wn[0] * q[sten[0]];
  t = fma(wn[1], q[sten[1]], t);
  t = fma(wn[2], q[sten[2]], t);
  t = fma(wn[3], q[sten[3]], t);
  return t;
}

/* ghost-cell values of (u, v) or p evaluated from `src` and written into `dst` (may alias only if the
 * caller accepts Gauss-Seidel ordering; the solver always passes distinct buffers or a gathered copy) */
void orc_gc_update_velocity(int ngc, const int* cell, const int* stencil, const double* wd, const int* body,
                            const double* ub, const double* vb, const double* usrc, const double* vsrc,
                            double* udst, double* vdst) {
  for (int g = 0; g < ngc; g++) {
    udst[cell[g]] = gc_dirichlet(usrc, stencil + 4 * g, wd + 5 * g, ub ? ub[body[g]] : 0.0);
    vdst[cell[g]] = gc_dirichlet(vsrc, stencil + 4 * g, wd + 5 * g, vb ? vb[body[g]] : 0.0);
  }
}
void orc_gc_update_pressure(int ngc, const int* cell, const int* stencil, const double* wn, const double* psrc, double* pdst) {
  for (int g = 0; g < ngc; g++) pdst[cell[g]] = gc_neumann(psrc, stencil + 4 * g, wn + 4 * g);
}

/* -----------------------------------------------------------------------------------------------
 * Boundary conditions with arbitrary values (the reference hard-codes u = 1, v = 0; set_velocity_BC,
 * ADSolver.cu:199-217).  two_bc = 2*bc for W, E, S, N.  Pass order W, S, E, N as in the pinned oracle
 * (corners = 2bc - (2bc - diagonal)).
 * --------------------------------------------------------------------------------------------- */
void orc_set_dirichlet_ring(int nx, int ny, double* q, const double* two_bc) {
  for (int j = 0; j < ny; j++) q[ID(0, j, nx)] = two_bc[0] - q[ID(1, j, nx)];
  for (int i = 0; i < nx; i++) q[ID(i, 0, nx)] = two_bc[2] - q[ID(i, 1, nx)];
  for (int j = 0; j < ny; j++) q[ID(nx - 1, j, nx)] = two_bc[1] - q[ID(nx - 2, j, nx)];
  for (int i = 0; i < nx; i++) q[ID(i, ny - 1, nx)] = two_bc[3] - q[ID(i, ny - 2, nx)];
}
/* homogeneous Neumann ring for p: ghost = interior (same pass order) */
void orc_set_neumann_ring(int nx, int ny, double* q) {
  for (int j = 0; j < ny; j++) q[ID(0, j, nx)] = q[ID(1, j, nx)];
  for (int i = 0; i < nx; i++) q[ID(i, 0, nx)] = q[ID(i, 1, nx)];
  for (int j = 0; j < ny; j++) q[ID(nx - 1, j, nx)] = q[ID(nx - 2, j, nx)];
  for (int i = 0; i < nx; i++) q[ID(i, ny - 1, nx)] = q[ID(i, ny - 2, nx)];
}

/* -----------------------------------------------------------------------------------------------
 * Face velocities by the reference's interpolation (Compute_velf, ADSolver.cu:173-174,184-185) on ALL faces,
 * then the closed-face rule: a face with a non-fluid cell on either side carries that body's velocity
 * (the east/north cell's body if it is non-fluid, else the west/south cell's).  ub/vb may be NULL (bodies at rest).
 * uf[i + (j-1)(nx-1)], i = 0..nx-2, j = 1..ny-2;  vf[(i-1) + j(nx-2)], i = 1..nx-2, j = 0..ny-2.
 * --------------------------------------------------------------------------------------------- */
void orc_faces_from_cells(int nx, int ny, const double* dx, const double* dy, const double* u, const double* v,
                          const unsigned char* celltype, const double* ub, const double* vb, double* uf, double* vf) {
  orc_Compute_velf(nx, ny, dx, dy, u, v, uf, vf, 1);
  if (!celltype) return;
#pragma omp parallel for
  for (int j = 1; j < ny - 1; j++)
    for (int i = 0; i < nx - 1; i++) {
      const unsigned char cw = celltype[ID(i, j, nx)], ce = celltype[ID(i + 1, j, nx)];
      if (!IS_FLUID(ce)) uf[i + (j - 1) * (nx - 1)] = ub ? ub[CT_BODY(ce)] : 0.0;
      else if (!IS_FLUID(cw)) uf[i + (j - 1) * (nx - 1)] = ub ? ub[CT_BODY(cw)] : 0.0;
    }
#pragma omp parallel for
  for (int j = 0; j < ny - 1; j++)
    for (int i = 1; i < nx - 1; i++) {
      const unsigned char cs = celltype[ID(i, j, nx)], cn = celltype[ID(i, j + 1, nx)];
      if (!IS_FLUID(cn)) vf[(i - 1) + j * (nx - 2)] = vb ? vb[CT_BODY(cn)] : 0.0;
      else if (!IS_FLUID(cs)) vf[(i - 1) + j * (nx - 2)] = vb ? vb[CT_BODY(cs)] : 0.0;
    }
}

/* -----------------------------------------------------------------------------------------------
 * a15 — PPE source term: rhs = (1/dt) div(uf*, vf*) on fluid cells, 0 elsewhere:
 *   rhs = ((uf_e - uf_w)/dx_i + (vf_n - vf_s)/dy_j) / dt          (plain IEEE ops, this order)
 * --------------------------------------------------------------------------------------------- */
void orc_ppe_source(int nx, int ny, const double* dx, const double* dy, double dt, const unsigned char* celltype,
                    const double* uf, const double* vf, double* rhs) {
#pragma omp parallel for
  for (int id = 0; id < nx * ny; id++) {
    const int i = id % nx, j = id / nx;
    rhs[id] = 0.0;
    if (i > 0 && i < nx - 1 && j > 0 && j < ny - 1 && IS_FLUID(celltype[id])) {
      const double ufe = uf[i + (j - 1) * (nx - 1)], ufw = uf[i - 1 + (j - 1) * (nx - 1)];
      const double vfn = vf[i - 1 + j * (nx - 2)], vfs = vf[i - 1 + (j - 1) * (nx - 2)];
      rhs[id] = ((ufe - ufw) / dx[id] + (vfn - vfs) / dy[id]) / dt;
    }
  }
}

/* -----------------------------------------------------------------------------------------------
 * General Poisson sweep: coefficients of PPESolver.cu:93-99; zero normal gradient on the grid boundary AND on
 * closed faces (the neighbour's p is replaced by the cell's own when the neighbour is a ring cell or not fluid);
 * non-fluid cells are copied through.  p_new = (rhs - t)/cP with t as in jacobiIteration; residual of the
 * INPUT iterate r = rhs - (A p) with (A p) as in Compute_Residual, on fluid cells.
 * --------------------------------------------------------------------------------------------- */
void orc_ppe_sweep_general(int nx, int ny, const double* cP, const double* cxm, const double* cxp, const double* cym,
                           const double* cyp, const unsigned char* celltype, const double* rhs, const double* p,
                           double* p_new, double* residual) {
#pragma omp parallel for
  for (int id = 0; id < nx * ny; id++) {
    const int i = id % nx, j = id / nx;
    if (residual) residual[id] = 0.0;
    if (!(i > 0 && i < nx - 1 && j > 0 && j < ny - 1)) continue;
    const double pc = p[id];
    const double pw = (i == 1 || !IS_FLUID(celltype[id - 1])) ? pc : p[id - 1];
    const double pe = (i == nx - 2 || !IS_FLUID(celltype[id + 1])) ? pc : p[id + 1];
    const double ps = (j == 1 || !IS_FLUID(celltype[id - nx])) ? pc : p[id - nx];
    const double pn = (j == ny - 2 || !IS_FLUID(celltype[id + nx])) ? pc : p[id + nx];
    double t = pw * cxm[id];
    t = fma(pe, cxp[id], t);
    t = fma(pn, cyp[id], t);
    t = fma(ps, cym[id], t);
    double q = pe * cxp[id];
    q = fma(pc, cP[id], q);
    q = fma(pw, cxm[id], q);
    q = fma(pn, cyp[id], q);
    q = fma(ps, cym[id], q);
    if (IS_FLUID(celltype[id])) {
      p_new[id] = (rhs[id] - t) / cP[id];
      if (residual) residual[id] = rhs[id] - q;
    } else {
      p_new[id] = pc;
    }
  }
}

/* -----------------------------------------------------------------------------------------------
 * SURVEY 8(f)-1 — red-black successive over-relaxation, the first of the "better Poisson iterations" the input
 * file's PPE_Solver / w-PPE fields ask for (parsed at main.cu:42, unused by the reference).  UNPINNED.
 * One half-sweep: cells of one colour ((i + j + colour) even) move to  p + omega ((rhs - t)/cP - p)  with t and
 * the closed-face substitutions exactly as in the Jacobi sweep above; everything else is copied.  An iteration
 * is the colour-0 half-sweep followed by the colour-1 half-sweep on its result.
 * --------------------------------------------------------------------------------------------- */
void orc_ppe_sor_halfsweep(int nx, int ny, const double* cP, const double* cxm, const double* cxp, const double* cym,
                           const double* cyp, const unsigned char* celltype, const double* rhs, int colour, double omega,
                           const double* p, double* p_new) {
#pragma omp parallel for
  for (int id = 0; id < nx * ny; id++) {
    const int i = id % nx, j = id / nx;
    if (!(i > 0 && i < nx - 1 && j > 0 && j < ny - 1)) continue;
    const double pc = p[id];
    p_new[id] = pc;
    if (!IS_FLUID(celltype[id]) || ((i + j + colour) & 1)) continue;
    const double pw = (i == 1 || !IS_FLUID(celltype[id - 1])) ? pc : p[id - 1];
    const double pe = (i == nx - 2 || !IS_FLUID(celltype[id + 1])) ? pc : p[id + 1];
    const double ps = (j == 1 || !IS_FLUID(celltype[id - nx])) ? pc : p[id - nx];
    const double pn = (j == ny - 2 || !IS_FLUID(celltype[id + nx])) ? pc : p[id + nx];
    double t = pw * cxm[id];
    t = fma(pe, cxp[id], t);
    t = fma(pn, cyp[id], t);
    t = fma(ps, cym[id], t);
    const double pj = (rhs[id] - t) / cP[id];
    p_new[id] = pc + omega * (pj - pc);
  }
}

/* -----------------------------------------------------------------------------------------------
 * a18 — projection.  With PN(nb) = p of the neighbour, or the cell's own p when the neighbour is a ring cell or
 * not fluid (zero normal gradient), on fluid cells
 *   u = u* - dt (pe - pw)/dx_i,  pe = rcp(dx_i+dx_ip1) * fma(PN(E), dx_i, p_C*dx_ip1)   (face value, velf form)
 * and on every OPEN face
 *   uf = uf* - dt (p_E - p_C) (2 rcp(dx_i+dx_ip1)),   uf* = velf(u*)
 * closed faces carry the body velocity, grid-boundary faces keep uf* (same in y).  u*, v* of non-fluid cells
 * are copied through.
 * --------------------------------------------------------------------------------------------- */
void orc_correct(int nx, int ny, const double* dx, const double* dy, double dt, const unsigned char* celltype,
                 const double* ub, const double* vb,
                 const double* us, const double* vs, const double* p, double* u, double* v, double* uf, double* vf) {
#pragma omp parallel for
  for (int id = 0; id < nx * ny; id++) {
    const int i = id % nx, j = id / nx;
    u[id] = us[id];
    v[id] = vs[id];
    if (i > 0 && i < nx - 1 && j > 0 && j < ny - 1 && IS_FLUID(celltype[id])) {
      const double dx_i = dx[id], dx_ip1 = dx[id + 1], dx_im1 = dx[id - 1];
      const double dy_j = dy[id], dy_jp1 = dy[id + nx], dy_jm1 = dy[id - nx];
      const double pc = p[id];
      const double pW = (i == 1 || !IS_FLUID(celltype[id - 1])) ? pc : p[id - 1];
      const double pE = (i == nx - 2 || !IS_FLUID(celltype[id + 1])) ? pc : p[id + 1];
      const double pS = (j == 1 || !IS_FLUID(celltype[id - nx])) ? pc : p[id - nx];
      const double pN = (j == ny - 2 || !IS_FLUID(celltype[id + nx])) ? pc : p[id + nx];
      const double pe = (1.0 / (dx_i + dx_ip1)) * fma(pE, dx_i, pc * dx_ip1);
      const double pw = (1.0 / (dx_im1 + dx_i)) * fma(pc, dx_im1, pW * dx_i);
      const double pn = (1.0 / (dy_j + dy_jp1)) * fma(pN, dy_j, pc * dy_jp1);
      const double ps = (1.0 / (dy_jm1 + dy_j)) * fma(pc, dy_jm1, pS * dy_j);
      u[id] = us[id] - dt * ((pe - pw) / dx_i);
      v[id] = vs[id] - dt * ((pn - ps) / dy_j);
    }
  }
  /* faces: uf(i,j) east of cell (i,j), i = 0..nx-2, j = 1..ny-2 */
#pragma omp parallel for
  for (int j = 1; j < ny - 1; j++)
    for (int i = 0; i < nx - 1; i++) {
      const int id = ID(i, j, nx);
      const unsigned char cw = celltype[id], ce = celltype[id + 1];
      double val;
      if (!IS_FLUID(ce)) val = ub ? ub[CT_BODY(ce)] : 0.0;
      else if (!IS_FLUID(cw)) val = ub ? ub[CT_BODY(cw)] : 0.0;
      else {
        const double dxL = dx[id], dxR = dx[id + 1];
        const double rc = 1.0 / (dxL + dxR);
        val = rc * fma(us[id + 1], dxL, us[id] * dxR);
        if (i >= 1 && i <= nx - 3) val = val - dt * ((p[id + 1] - p[id]) * (2.0 * rc));
      }
      uf[i + (j - 1) * (nx - 1)] = val;
    }
#pragma omp parallel for
  for (int j = 0; j < ny - 1; j++)
    for (int i = 1; i < nx - 1; i++) {
      const int id = ID(i, j, nx);
      const unsigned char cs = celltype[id], cn = celltype[id + nx];
      double val;
      if (!IS_FLUID(cn)) val = vb ? vb[CT_BODY(cn)] : 0.0;
      else if (!IS_FLUID(cs)) val = vb ? vb[CT_BODY(cs)] : 0.0;
      else {
        const double dyL = dy[id], dyU = dy[id + nx];
        const double rc = 1.0 / (dyU + dyL);
        val = rc * fma(dyL, vs[id + nx], vs[id] * dyU);
        if (j >= 1 && j <= ny - 3) val = val - dt * ((p[id + nx] - p[id]) * (2.0 * rc));
      }
      vf[(i - 1) + j * (nx - 2)] = val;
    }
}

/* =================================================================================================
 * Full fractional step — the sequence IFX_COMPAT_FULL implements.  Opaque-handle API so the
 * tests drive it exactly like the C-ABI of the product.
 * ================================================================================================= */
struct orc_full {
  int nx, ny, AD_itermax, PPE_itermax, ppe_abs;
  double dt, Re, ad_tol, ppe_tol;
  double two_bc_u[4], two_bc_v[4];
  double *xc, *yc, *dx, *dy;
  double *u, *v, *p, *uf, *vf, *rhs, *sx, *sy;
  unsigned char* celltype;
  double* iblank;                     /* 1.0 fluid / 0.0 otherwise, what the reference kernels multiply by */
  int* body_of;
  int nbodies; int* off; double *xm, *ym, *ub, *vb;
  int ngc, *cell, *stencil, *body; double *wd, *wn, *bi, *ip;
  int faces_valid;
  int ppe_solver;                     /* 1: point Jacobi (the reference's sweep), 3: red-black SOR, 4: multigrid */
  double w_ppe;
  int mg_nu1, mg_nu2, mg_ncoarse;
  orc_mg* mg;                         /* hierarchy of the last multigrid solve (rebuilt per solve: bodies may move) */
};

orc_full* orc_full_create(int nx, int ny, const double* xf, const double* yf, double dt, double Re, int AD_itermax,
                          int PPE_itermax, double ad_tol, double ppe_tol, int ppe_abs, const double* bc_u, const double* bc_v) {
  orc_full* s = (orc_full*)calloc(1, sizeof(orc_full));
  const size_t N = (size_t)nx * ny;
  s->nx = nx; s->ny = ny; s->dt = dt; s->Re = Re; s->AD_itermax = AD_itermax; s->PPE_itermax = PPE_itermax;
  s->ad_tol = ad_tol; s->ppe_tol = ppe_tol; s->ppe_abs = ppe_abs;
  s->ppe_solver = 1; s->w_ppe = 1.0;
  s->mg_nu1 = ORC_MG_NU1; s->mg_nu2 = ORC_MG_NU2; s->mg_ncoarse = ORC_MG_NCOARSE;
  for (int q = 0; q < 4; q++) { s->two_bc_u[q] = bc_u[q] * 2.0; s->two_bc_v[q] = bc_v[q] * 2.0; }   /* W, E, S, N */
  s->xc = (double*)calloc(nx, 8); s->yc = (double*)calloc(ny, 8);
  s->dx = (double*)calloc(N, 8); s->dy = (double*)calloc(N, 8);
  orc_grid_metrics(nx, ny, xf, yf, s->xc, s->yc, s->dx, s->dy);
  s->u = (double*)calloc(N, 8); s->v = (double*)calloc(N, 8); s->p = (double*)calloc(N, 8);
  s->uf = (double*)calloc(N, 8); s->vf = (double*)calloc(N, 8); s->rhs = (double*)calloc(N, 8);
  s->sx = (double*)calloc(N, 8); s->sy = (double*)calloc(N, 8);
  s->celltype = (unsigned char*)malloc(N); memset(s->celltype, 1, N);
  s->iblank = (double*)malloc(N * 8); for (size_t k = 0; k < N; k++) s->iblank[k] = 1.0;
  s->body_of = (int*)malloc(N * sizeof(int));
  return s;
}

static void free_gc(orc_full* s) {
  free(s->cell); free(s->stencil); free(s->body); free(s->wd); free(s->wn); free(s->bi); free(s->ip);
  s->cell = s->stencil = s->body = NULL; s->wd = s->wn = s->bi = s->ip = NULL; s->ngc = 0;
}

void orc_full_destroy(orc_full* s) {
  if (!s) return;
  free_gc(s);
  orc_mg_destroy(s->mg);
  free(s->xc); free(s->yc); free(s->dx); free(s->dy); free(s->u); free(s->v); free(s->p); free(s->uf); free(s->vf);
  free(s->rhs); free(s->sx); free(s->sy); free(s->celltype); free(s->iblank); free(s->body_of);
  free(s->off); free(s->xm); free(s->ym); free(s->ub); free(s->vb);
  free(s);
}

void orc_full_set_bodies(orc_full* s, int nbodies, const int* off, const double* xm, const double* ym,
                         const double* ub, const double* vb) {
  free(s->off); free(s->xm); free(s->ym); free(s->ub); free(s->vb);
  const int nm = nbodies ? off[nbodies] : 0;
  s->nbodies = nbodies;
  s->off = (int*)malloc(sizeof(int) * (nbodies + 1)); memcpy(s->off, off, sizeof(int) * (nbodies + 1));
  s->xm = (double*)malloc(8 * (nm + 1)); s->ym = (double*)malloc(8 * (nm + 1));
  memcpy(s->xm, xm, 8 * (size_t)nm); memcpy(s->ym, ym, 8 * (size_t)nm);
  s->ub = (double*)calloc(nbodies + 1, 8); s->vb = (double*)calloc(nbodies + 1, 8);
  if (ub) memcpy(s->ub, ub, 8 * (size_t)nbodies);
  if (vb) memcpy(s->vb, vb, 8 * (size_t)nbodies);
}

/* a16 + a17: classification, ghost-cell list, stencils */
int orc_full_update_ib(orc_full* s) {
  const size_t N = (size_t)s->nx * s->ny;
  orc_iblank_classify(s->nx, s->ny, s->xc, s->yc, s->nbodies, s->off, s->xm, s->ym, s->celltype, s->body_of);
  for (size_t k = 0; k < N; k++) s->iblank[k] = IS_FLUID(s->celltype[k]) ? 1.0 : 0.0;
  free_gc(s);
  int n = 0;
  for (size_t k = 0; k < N; k++) n += (CT_TYPE(s->celltype[k]) == 2);
  s->ngc = n;
  s->cell = (int*)malloc(sizeof(int) * (n + 1)); s->stencil = (int*)malloc(sizeof(int) * 4 * (n + 1));
  s->body = (int*)malloc(sizeof(int) * (n + 1));
  s->wd = (double*)malloc(8 * 5 * (size_t)(n + 1)); s->wn = (double*)malloc(8 * 4 * (size_t)(n + 1));
  s->bi = (double*)malloc(8 * 2 * (size_t)(n + 1)); s->ip = (double*)malloc(8 * 2 * (size_t)(n + 1));
  orc_ghost_cells(s->nx, s->ny, s->xc, s->yc, s->nbodies, s->off, s->xm, s->ym, s->celltype, s->body_of, n, s->cell,
                  s->stencil, s->wd, s->wn, s->bi, s->ip, s->body);
  /* the set of closed faces moved with the bodies: face velocities are re-initialised from the cell-centred
   * velocities (interpolation + closed-face rule) at the start of the next step */
  s->faces_valid = 0;
  return n;
}

/* ring + ghost cells of (u, v), ghost cells gathered from the field BEFORE any of them is overwritten */
static void apply_velocity_bc(orc_full* s, double* u, double* v) {
  orc_set_dirichlet_ring(s->nx, s->ny, u, s->two_bc_u);
  orc_set_dirichlet_ring(s->nx, s->ny, v, s->two_bc_v);
  if (s->ngc) {
    double* tu = (double*)malloc(8 * (size_t)s->ngc); double* tv = (double*)malloc(8 * (size_t)s->ngc);
    for (int g = 0; g < s->ngc; g++) {
      tu[g] = gc_dirichlet(u, s->stencil + 4 * g, s->wd + 5 * g, s->ub[s->body[g]]);
      tv[g] = gc_dirichlet(v, s->stencil + 4 * g, s->wd + 5 * g, s->vb[s->body[g]]);
    }
    for (int g = 0; g < s->ngc; g++) { u[s->cell[g]] = tu[g]; v[s->cell[g]] = tv[g]; }
    free(tu); free(tv);
  }
}
static void apply_pressure_bc(orc_full* s, double* p) {
  orc_set_neumann_ring(s->nx, s->ny, p);
  if (s->ngc) {
    double* t = (double*)malloc(8 * (size_t)s->ngc);
    for (int g = 0; g < s->ngc; g++) t[g] = gc_neumann(p, s->stencil + 4 * g, s->wn + 4 * g);
    for (int g = 0; g < s->ngc; g++) p[s->cell[g]] = t[g];
    free(t);
  }
}

/* stats: [0] K_AD, [1] uRes, [2] vRes, [3] K_PPE, [4] PPE residual (signed or abs sum per ppe_abs) */
int orc_full_predictor(orc_full* s, double* stats) {
  const int nx = s->nx, ny = s->ny;
  const size_t N = (size_t)nx * ny;
  const int tpb = 256, bpg = ((int)N + tpb - 1) / tpb;
  double* mem = (double*)calloc(N * 9, 8);
  double *uT = mem, *vT = mem + N, *res = mem + 2 * N, *c = mem + 3 * N, *cxp = mem + 4 * N, *cxm = mem + 5 * N,
         *cyp = mem + 6 * N, *cym = mem + 7 * N, *res2 = mem + 8 * N;
  double *uc = s->u, *vc = s->v;
  apply_velocity_bc(s, uc, vc);
  if (!s->faces_valid) {
    orc_faces_from_cells(nx, ny, s->dx, s->dy, uc, vc, s->celltype, s->ub, s->vb, s->uf, s->vf);
    s->faces_valid = 1;
  }
  orc_calculateADCoefficients(nx, ny, s->dx, s->dy, s->dt, s->Re, c, cxm, cxp, cym, cyp);
  orc_ADSource(nx, ny, s->dx, s->dy, s->dt, uc, vc, s->uf, s->vf, s->sx, s->sy);
  memcpy(uT, uc, N * 8); memcpy(vT, vc, N * 8);
  double uRes = 1.0, vRes = 1.0;
  int iter = 0;
  while (uRes + vRes > s->ad_tol && iter < s->AD_itermax) {
    orc_set_dirichlet_ring(nx, ny, uc, s->two_bc_u);
    orc_set_dirichlet_ring(nx, ny, vc, s->two_bc_v);
    orc_ADsolver_kernel(nx, ny, c, cxm, cxp, cym, cyp, s->iblank, uc, uT, s->sx);
    orc_ADsolver_kernel(nx, ny, c, cxm, cxp, cym, cyp, s->iblank, vc, vT, s->sy);
    if (s->ngc) orc_gc_update_velocity(s->ngc, s->cell, s->stencil, s->wd, s->body, s->ub, s->vb, uc, vc, uT, vT);
    double* t = uc; uc = uT; uT = t;
    t = vc; vc = vT; vT = t;
    orc_Compute_Residual_AD(nx, ny, s->iblank, uc, uT, res);
    orc_Compute_Residual_AD(nx, ny, s->iblank, vc, vT, res2);
    uRes = orc_Reduction(res, (int)N, tpb, bpg);
    vRes = orc_Reduction(res2, (int)N, tpb, bpg);
    iter++;
  }
  if (uc != s->u) { memcpy(s->u, uc, N * 8); memcpy(s->v, vc, N * 8); }
  apply_velocity_bc(s, s->u, s->v);
  if (stats) { stats[0] = iter; stats[1] = uRes; stats[2] = vRes; }
  free(mem);
  return iter;
}

int orc_full_poisson(orc_full* s, double* stats) {
  const int nx = s->nx, ny = s->ny;
  const size_t N = (size_t)nx * ny;
  const int tpb = 256, bpg = ((int)N + tpb - 1) / tpb;
  double* mem = (double*)calloc(N * 10, 8);
  double *pT = mem, *res = mem + N, *cP = mem + 2 * N, *cxp = mem + 3 * N, *cxm = mem + 4 * N, *cyp = mem + 5 * N,
         *cym = mem + 6 * N, *ufs = mem + 7 * N, *vfs = mem + 8 * N, *scratch = mem + 9 * N;
  /* source term from the predicted velocities (ring + ghost cells are consistent after the predictor) */
  orc_faces_from_cells(nx, ny, s->dx, s->dy, s->u, s->v, s->celltype, s->ub, s->vb, ufs, vfs);
  orc_ppe_source(nx, ny, s->dx, s->dy, s->dt, s->celltype, ufs, vfs, s->rhs);
  orc_calculatePPECoefficients(nx, ny, s->dx, s->dy, cP, cxm, cxp, cym, cyp);
  double* pc = s->p;
  memcpy(pT, pc, N * 8);
  double R = 1.0, Rabs = 1.0;
  int iter = 0;
  if (s->ppe_solver == 4 || s->ppe_solver == 5) {          /* an "iteration" is one V-cycle */
    orc_mg_destroy(s->mg);
    s->mg = orc_mg_create2(nx, ny, s->dx, s->dy, s->celltype, s->ppe_solver == 5);
  }
  while ((s->ppe_abs ? Rabs : R) > s->ppe_tol && iter < s->PPE_itermax) {
    if (s->ppe_solver == 2) {          /* zebra line relaxation; scratch: pT and res */
      orc_ppe_line_iteration(nx, ny, cP, cxm, cxp, cym, cyp, s->celltype, s->rhs, s->w_ppe, pc, pT, res);
    } else if (s->ppe_solver == 5) {
      orc_mg_vcycle_lines(s->mg, nx, ny, s->dx, s->dy, cP, cxm, cxp, cym, cyp, s->celltype, s->rhs, s->mg_nu1, s->mg_nu2,
                          s->mg_ncoarse, s->w_ppe, pc, pT, res);
    } else if (s->ppe_solver == 4) {
      orc_mg_vcycle(s->mg, nx, ny, s->dx, s->dy, cP, cxm, cxp, cym, cyp, s->celltype, s->rhs, s->mg_nu1, s->mg_nu2,
                    s->mg_ncoarse, s->w_ppe, pc, pT);
    } else if (s->ppe_solver == 3) {       /* red-black SOR: two half-sweeps, the iterate ends up where it started */
      orc_ppe_sor_halfsweep(nx, ny, cP, cxm, cxp, cym, cyp, s->celltype, s->rhs, 0, s->w_ppe, pc, pT);
      orc_ppe_sor_halfsweep(nx, ny, cP, cxm, cxp, cym, cyp, s->celltype, s->rhs, 1, s->w_ppe, pT, pc);
    } else {
      orc_ppe_sweep_general(nx, ny, cP, cxm, cxp, cym, cyp, s->celltype, s->rhs, pc, pT, NULL);
      double* t = pc; pc = pT; pT = t;
    }
    /* residual of the new iterate (the reference evaluates it in a second kernel too, PPESolver.cu:182);
     * `scratch` only absorbs the sweep's p_new output */
    orc_ppe_sweep_general(nx, ny, cP, cxm, cxp, cym, cyp, s->celltype, s->rhs, pc, scratch, res);
    R = orc_Reduction(res, (int)N, tpb, bpg);
#pragma omp parallel for
    for (long k = 0; k < (long)N; k++) res[k] = fabs(res[k]);
    Rabs = orc_Reduction(res, (int)N, tpb, bpg);
    iter++;
  }
  if (pc != s->p) memcpy(s->p, pc, N * 8);
  apply_pressure_bc(s, s->p);
  if (stats) { stats[3] = iter; stats[4] = s->ppe_abs ? Rabs : R; }
  free(mem);
  return iter;
}

void orc_full_set_ppe_solver(orc_full* s, int solver, double omega) { s->ppe_solver = solver; s->w_ppe = omega; }
void orc_full_set_mg(orc_full* s, int nu1, int nu2, int ncoarse) { s->mg_nu1 = nu1; s->mg_nu2 = nu2; s->mg_ncoarse = ncoarse; }
const orc_mg* orc_full_mg(const orc_full* s) { return s->mg; }

void orc_full_correct(orc_full* s) {
  const size_t N = (size_t)s->nx * s->ny;
  double* un = (double*)malloc(N * 8); double* vn = (double*)malloc(N * 8);
  orc_correct(s->nx, s->ny, s->dx, s->dy, s->dt, s->celltype, s->ub, s->vb, s->u, s->v, s->p, un, vn, s->uf, s->vf);
  memcpy(s->u, un, N * 8); memcpy(s->v, vn, N * 8);
  free(un); free(vn);
  apply_velocity_bc(s, s->u, s->v);
  s->faces_valid = 1;
}

void orc_full_step(orc_full* s, double* stats) {
  orc_full_predictor(s, stats);
  orc_full_poisson(s, stats);
  orc_full_correct(s);
}

void orc_full_body_forces(orc_full* s, double* F) {
  orc_body_forces(s->nx, s->ny, s->xc, s->yc, s->celltype, s->Re, s->nbodies, s->off, s->xm, s->ym, s->ub, s->vb, s->u,
                  s->v, s->p, F);
}
void orc_full_probe(orc_full* s, int npts, const double* px, const double* py, double* ou, double* ov, double* op) {
  orc_probe(s->nx, s->ny, s->xc, s->yc, s->celltype, s->u, s->v, s->p, npts, px, py, ou, ov, op);
}

/* field ids follow include/immerseflow_c.h's ifx_field */
static double* full_field(orc_full* s, int f) {
  switch (f) {
    case 0: return s->u; case 1: return s->v; case 2: return s->p; case 4: return s->uf; case 5: return s->vf;
    case 6: return s->sx; case 7: return s->sy; case 8: return s->rhs; default: return NULL;
  }
}
static size_t full_field_n(orc_full* s, int f) {
  if (f == 4) return (size_t)(s->nx - 1) * (s->ny - 2);
  if (f == 5) return (size_t)(s->nx - 2) * (s->ny - 1);
  return (size_t)s->nx * s->ny;
}
int orc_full_get(orc_full* s, int f, double* out) {
  const size_t N = (size_t)s->nx * s->ny;
  if (f == 3) { for (size_t k = 0; k < N; k++) out[k] = s->iblank[k]; return 0; }
  if (f == 11) { for (size_t k = 0; k < N; k++) out[k] = CT_TYPE(s->celltype[k]); return 0; }
  double* p = full_field(s, f);
  if (!p) return -1;
  memcpy(out, p, 8 * full_field_n(s, f));
  return 0;
}
int orc_full_set(orc_full* s, int f, const double* in) {
  double* p = full_field(s, f);
  if (!p) return -1;
  memcpy(p, in, 8 * full_field_n(s, f));
  if (f == 0 || f == 1) s->faces_valid = 0;
  if (f == 4 || f == 5) s->faces_valid = 1;
  return 0;
}
int orc_full_ghost_cells(orc_full* s, int* cell, int* stencil, double* w10, double* bi, double* ip) {
  for (int g = 0; g < s->ngc; g++) {
    if (cell) cell[g] = s->cell[g];
    if (stencil) memcpy(stencil + 4 * g, s->stencil + 4 * g, 4 * sizeof(int));
    if (w10) { memcpy(w10 + 10 * g, s->wd + 5 * g, 5 * 8); memcpy(w10 + 10 * g + 5, s->wn + 4 * g, 4 * 8); w10[10 * g + 9] = s->body[g]; }
    if (bi) memcpy(bi + 2 * g, s->bi + 2 * g, 16);
    if (ip) memcpy(ip + 2 * g, s->ip + 2 * g, 16);
  }
  return s->ngc;
}
