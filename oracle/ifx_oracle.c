/*
 * ifx_oracle.c — CPU restatement (plain C + OpenMP) of the ImmerseFlow++ hot path,
 * reference-pinned part.  TEST INFRASTRUCTURE ONLY — see ifx_oracle.h.
 *
 * Build: gcc -O2 -ffp-contract=off -fopenmp -shared -fPIC (oracle/Makefile).
 * -ffp-contract=off is REQUIRED: every FMA below is deliberate and mirrors the contraction
 * nvcc 12.9 emits for the reference (-arch=sm_100, default flags); see DESIGN.md §"bit parity".
 *
 * Layout is the reference's own: row-major id = i + j*nx, ghost-inclusive nx*ny arrays,
 * 2-D dx/dy/coefficient arrays (globalVariables.cuh:41-68), iBlank as double.
 */
#include "ifx_oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ID(i, j, nx) ((i) + (j) * (nx))

/* bench.py sets the thread count explicitly: torchrun exports OMP_NUM_THREADS=1 to its workers */
void orc_set_num_threads(int n) {
  if (n > 0) omp_set_num_threads(n);
}
int orc_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* ------------------------------------------------------------------------------------------
 * a2 — readGridData(), preSim.cu:294-355.  xf has nx-1 faces, yf has ny-1.
 * ---------------------------------------------------------------------------------------- */
void orc_grid_metrics(int nx, int ny, const double* xf, const double* yf,
                      double* xc, double* yc, double* dx, double* dy) {
  const int nxf = nx - 1, nyf = ny - 1;
  for (int i = 1; i < nx - 1; i++) xc[i] = (xf[i - 1] + xf[i]) / 2.0;      /* :294-296 */
  for (int i = 1; i < ny - 1; i++) yc[i] = (yf[i - 1] + yf[i]) / 2.0;      /* :299-301 */
  xc[0] = -1 * xc[1];                                                      /* :304 */
  yc[0] = -1 * yc[1];                                                      /* :305 */
  xc[nx - 1] = xf[nxf - 1] + (xf[nxf - 1] - xc[nx - 2]);                   /* :306 */
  yc[ny - 1] = yf[nyf - 1] + (yf[nyf - 1] - yc[ny - 2]);                   /* :307 */

  for (int j = 1; j < ny - 1; j++)                                         /* :329-334 */
    for (int i = 1; i < nx - 1; i++) {
      dx[ID(i, j, nx)] = xf[i] - xf[i - 1];
      dy[ID(i, j, nx)] = yf[j] - yf[j - 1];
    }
  /* ghost extrapolation, in the reference's loop order (:337-355) */
  for (int j = 0; j < ny; j++) {
    dx[ID(0, j, nx)] = dx[ID(1, j, nx)];
    dx[ID(nx - 1, j, nx)] = dx[ID(nx - 2, j, nx)];
  }
  for (int i = 0; i < nx; i++) {
    dx[ID(i, 0, nx)] = dx[ID(i, 1, nx)];
    dx[ID(i, ny - 1, nx)] = dx[ID(i, ny - 2, nx)];
  }
  for (int j = 0; j < ny; j++) {
    dy[ID(0, j, nx)] = dy[ID(1, j, nx)];
    dy[ID(nx - 1, j, nx)] = dy[ID(nx - 2, j, nx)];
  }
  for (int i = 0; i < nx; i++) {
    dy[ID(i, 0, nx)] = dy[ID(i, 1, nx)];
    dy[ID(i, ny - 1, nx)] = dy[ID(i, ny - 2, nx)];
  }
}

/* ------------------------------------------------------------------------------------------
 * a19 — initializeKernel, preSim.cu:53-76 (uniform stream + Gaussian vortex).  Host libm
 * here vs CUDA libdevice in the reference: agreement is <= 1 ulp, so parity tests pass the
 * initial state across explicitly instead of regenerating it on each side.
 * ---------------------------------------------------------------------------------------- */
void orc_initializeKernel(int nx, int ny, const double* xc, const double* yc,
                          double* u, double* v, double* p) {
#pragma omp parallel for
  for (int id = 0; id < nx * ny; id++) {
    int idx = id % nx, idy = id / nx;
    double r = sqrt(pow(xc[idx] - 0.5, 2.0) + pow(yc[idy] - 0.5, 2.0));
    double r0 = 0.1;
    u[id] = 1.0 - 0.25 * (yc[idy] - 0.5) * exp((1.0 - pow(r / r0, 2.0)) / 2.0);
    v[id] = 0.25 * (xc[idx] - 0.5) * exp((1.0 - pow(r / r0, 2.0)) / 2.0);
    p[id] = 0.0;
  }
}

/* a16 as written — preSim.cu:110-136: the classification is commented out, iBlank = 1.0. */
void orc_iBlankComputeKernel(int nx, int ny, const double* xc, const double* yc, double* iblank) {
  (void)xc; (void)yc;
  for (int id = 0; id < nx * ny; id++) iblank[id] = 1.0;
}

/* ------------------------------------------------------------------------------------------
 * a3 — calculateADCoefficients, ADSolver.cu:12-44.
 * nvcc: k = dt/Re (div.rn); cP = fma(k, ay_p+ay_m, fma(k, ax_p+ax_m, 1.0)); others k*a (mul).
 * ---------------------------------------------------------------------------------------- */
void orc_calculateADCoefficients(int nx, int ny, const double* dx, const double* dy,
                                 double dt, double Re,
                                 double* coeff, double* dx2_m1, double* dx2_p1,
                                 double* dy2_m1, double* dy2_p1) {
  const double k = dt / Re;
#pragma omp parallel for
  for (int id = 0; id < nx * ny; id++) {
    int i = id % nx, j = id / nx;
    coeff[id] = 1; dx2_m1[id] = 1; dx2_p1[id] = 1; dy2_m1[id] = 1; dy2_p1[id] = 1;
    if (i > 0 && i < nx - 1 && j > 0 && j < ny - 1) {
      double dx_i = dx[id], dx_ip1 = dx[id + 1], dx_im1 = dx[id - 1];
      double dy_j = dy[id], dy_jp1 = dy[id + nx], dy_jm1 = dy[id - nx];
      double ax_p = 2.0 / (dx_i * (dx_i + dx_ip1));
      double ax_m = 2.0 / (dx_i * (dx_i + dx_im1));
      double ay_p = 2.0 / (dy_j * (dy_j + dy_jp1));
      double ay_m = 2.0 / (dy_j * (dy_j + dy_jm1));
      coeff[id] = fma(k, ay_p + ay_m, fma(k, ax_p + ax_m, 1.0));
      dx2_m1[id] = k * ax_m;
      dx2_p1[id] = k * ax_p;
      dy2_m1[id] = k * ay_m;
      dy2_p1[id] = k * ay_p;
    }
  }
}

/* ------------------------------------------------------------------------------------------
 * a4 — Compute_velf, ADSolver.cu:162-189.
 * uf[id] = rcp(dxL+dxR) * fma(uR, dxL, uL*dxR)   (rcp.rn == 1.0/x correctly rounded)
 * BUG-AS-WRITTEN (vf_mode 0): `id` is not reset before the second while loop, so vf is only
 * written for thread ids in [(nx-1)(ny-2), (nx-2)(ny-1)) — empty when nx <= ny.  For nx > ny
 * the reference overruns its vf allocation (preSim.cu:153); here vf has its logical extent.
 * ---------------------------------------------------------------------------------------- */
void orc_Compute_velf(int nx, int ny, const double* dx, const double* dy,
                      const double* u, const double* v, double* uf, double* vf, int vf_mode) {
  const int nuf = (nx - 1) * (ny - 2), nvf = (nx - 2) * (ny - 1);
#pragma omp parallel for
  for (int id = 0; id < nuf; id++) {
    int i = id % (nx - 1), j = id / (nx - 1);
    int idc = j * nx + i;
    double dxL = dx[idc + nx], dxR = dx[idc + nx + 1];
    uf[id] = (1.0 / (dxL + dxR)) * fma(u[idc + nx + 1], dxL, u[idc + nx] * dxR);
  }
  int first = vf_mode ? 0 : nuf;
#pragma omp parallel for
  for (int id = first; id < nvf; id++) {
    int i = id % (nx - 2), j = id / (nx - 2);
    int idc = j * nx + i;
    double dyU = dy[idc + nx + 1], dyL = dy[idc + 1];
    vf[id] = (1.0 / (dyU + dyL)) * fma(dyL, v[idc + nx + 1], v[idc + 1] * dyU);
  }
}

/* ------------------------------------------------------------------------------------------
 * a5 — set_velocity_BC, ADSolver.cu:191-219.  ghost = -interior + bc*2 with bc_u = 1, bc_v = 0
 * (:200-216).  Edges are exact.  The four corner cells are a data race in the reference
 * (thread (0,0) reads a neighbour ghost another thread is writing); they are never read by
 * any stencil.  Deterministic choice here (one of the race's legal outcomes): the rules are
 * applied as whole passes in source order W, S, E, N, so every corner ends up as
 * 2bc - (2bc - diagonal interior neighbour).
 * ---------------------------------------------------------------------------------------- */
void orc_set_velocity_BC(int nx, int ny, double* u, double* v) {
  for (int j = 0; j < ny; j++) {            /* i == 0 */
    int id = ID(0, j, nx);
    u[id] = -u[id + 1] + 1.0 * 2.0;
    v[id] = -v[id + 1] + 0.0 * 2.0;
  }
  for (int i = 0; i < nx; i++) {            /* j == 0 */
    int id = ID(i, 0, nx);
    u[id] = -u[id + nx] + 1.0 * 2.0;
    v[id] = -v[id + nx] + 0.0 * 2.0;
  }
  for (int j = 0; j < ny; j++) {            /* i == nx-1 */
    int id = ID(nx - 1, j, nx);
    u[id] = -u[id - 1] + 1.0 * 2.0;
    v[id] = -v[id - 1] + 0.0 * 2.0;
  }
  for (int i = 0; i < nx; i++) {            /* j == ny-1 */
    int id = ID(i, ny - 1, nx);
    u[id] = -u[id - nx] + 1.0 * 2.0;
    v[id] = -v[id - nx] + 0.0 * 2.0;
  }
}

/* ------------------------------------------------------------------------------------------
 * a6 — ADSource, ADSolver.cu:46-79.  Contraction read from the sm_100 SASS (ptxas fuses the
 * subtractions that nvvm left as mul+sub):
 *   kx = (dt/dx_i)*0.5, ky = (dt/dy_j)*0.5, rcp_* = 1/(d + d_nb)
 *   sx: ue = fma(dx_i ,u_E,dx_ip1*u_C)*rcp_e   uw = fma(dx_im1,u_C,dx_i*u_W)*rcp_w
 *       un = fma(dy_j ,u_N,dy_jp1*u_C)*rcp_n   us = fma(dy_jm1,u_C,dy_j*u_S)*rcp_s
 *       sx = fma(-ky, fma(us,-vf_s, un*vf_n), fma(-kx, fma(uw,-uf_w, ue*uf_e), u_C))
 *   sy: ve = fma(dx_ip1,v_C,dx_i*v_E)*rcp_e    (note: operands swapped relative to ue)
 *       vw, vn, vs as for u
 *       sy = fma(-ky, fma(vn,vf_n, -(vs*vf_s)), fma(-kx, fma(ve,uf_e, -(vw*uf_w)), v_C))
 * ---------------------------------------------------------------------------------------- */
void orc_ADSource(int nx, int ny, const double* dx, const double* dy, double dt,
                  const double* u, const double* v, const double* uf, const double* vf,
                  double* sx, double* sy) {
#pragma omp parallel for
  for (int id = 0; id < nx * ny; id++) {
    int i = id % nx, j = id / nx;
    sx[id] = 0.0;
    sy[id] = 0.0;
    if (i > 0 && i < nx - 1 && j > 0 && j < ny - 1) {
      double dx_i = dx[id], dx_ip1 = dx[id + 1], dx_im1 = dx[id - 1];
      double dy_j = dy[id], dy_jp1 = dy[id + nx], dy_jm1 = dy[id - nx];
      double kx = (dt / dx_i) * 0.5, ky = (dt / dy_j) * 0.5;
      double rcp_e = 1.0 / (dx_i + dx_ip1), rcp_w = 1.0 / (dx_i + dx_im1);
      double rcp_n = 1.0 / (dy_j + dy_jp1), rcp_s = 1.0 / (dy_j + dy_jm1);
      double uf_e = uf[i + (j - 1) * (nx - 1)], uf_w = uf[i - 1 + (j - 1) * (nx - 1)];
      double vf_n = vf[i - 1 + j * (nx - 2)], vf_s = vf[i - 1 + (j - 1) * (nx - 2)];

      double uc = u[id];
      double ue = fma(dx_i, u[id + 1], dx_ip1 * uc) * rcp_e;
      double uw = fma(dx_im1, uc, dx_i * u[id - 1]) * rcp_w;
      double un = fma(dy_j, u[id + nx], dy_jp1 * uc) * rcp_n;
      double us = fma(dy_jm1, uc, dy_j * u[id - nx]) * rcp_s;
      double dfx = fma(uw, -uf_w, ue * uf_e);
      double dfy = fma(us, -vf_s, un * vf_n);
      sx[id] = fma(-ky, dfy, fma(-kx, dfx, uc));

      double vc = v[id];
      double ve = fma(dx_ip1, vc, dx_i * v[id + 1]) * rcp_e;
      double vw = fma(dx_im1, vc, dx_i * v[id - 1]) * rcp_w;
      double vn = fma(dy_j, v[id + nx], dy_jp1 * vc) * rcp_n;
      double vs = fma(dy_jm1, vc, dy_j * v[id - nx]) * rcp_s;
      double gfx = fma(ve, uf_e, -(vw * uf_w));
      double gfy = fma(vn, vf_n, -(vs * vf_s));
      sy[id] = fma(-ky, gfy, fma(-kx, gfx, vc));
    }
  }
}

/* ------------------------------------------------------------------------------------------
 * a7 — ADusolver_kernel / ADvsolver_kernel, ADSolver.cu:81-120: interior cells only;
 * qnew = iBlank * fma(cS,q_S, fma(cN,q_N, fma(cW,q_W, fma(cE,q_E, s)))) / cP.
 * Boundary cells of qnew are NOT written (they keep whatever the buffer held).
 * ---------------------------------------------------------------------------------------- */
void orc_ADsolver_kernel(int nx, int ny,
                         const double* coeff, const double* dx2_m1, const double* dx2_p1,
                         const double* dy2_m1, const double* dy2_p1, const double* iblank,
                         const double* q, double* qnew, const double* s) {
#pragma omp parallel for
  for (int id = 0; id < nx * ny; id++) {
    int i = id % nx, j = id / nx;
    if (i > 0 && i < nx - 1 && j > 0 && j < ny - 1) {
      double t = fma(dx2_p1[id], q[id + 1], s[id]);
      t = fma(dx2_m1[id], q[id - 1], t);
      t = fma(dy2_p1[id], q[id + nx], t);
      t = fma(dy2_m1[id], q[id - nx], t);
      qnew[id] = (iblank[id] * t) / coeff[id];
    }
  }
}

/* a8 — Compute_{u,v}Residual_AD, ADSolver.cu:122-160: |qnew - q| on interior cells with
 * iBlank == 1, zero elsewhere. */
void orc_Compute_Residual_AD(int nx, int ny, const double* iblank,
                             const double* q, const double* qnew, double* res) {
#pragma omp parallel for
  for (int id = 0; id < nx * ny; id++) {
    int i = id % nx, j = id / nx;
    res[id] = 0.0;
    if (i > 0 && i < nx - 1 && j > 0 && j < ny - 1 && iblank[id] == 1)
      res[id] = fabs(qnew[id] - q[id]);
  }
}

/* ------------------------------------------------------------------------------------------
 * a9 — reduce6<bs> + warpReduce<bs> (preSim.cu:12-50) driven by ImmerseFlow::Reduction
 * (preSim.cu:376-441): level 1 with blocksPerGrid blocks over n values, level 2 with one block
 * over the blocksPerGrid partials.  The summation ORDER is reproduced exactly so the
 * result is bit-identical to the reference (no atomics there => deterministic).
 * ---------------------------------------------------------------------------------------- */
static double reduce6_block(const double* g, unsigned n, unsigned bs, unsigned block,
                            unsigned nblocks, double* sdata) {
  const unsigned gridSize = bs * 2 * nblocks;
  for (unsigned tid = 0; tid < bs; tid++) {
    unsigned i = block * (bs * 2) + tid;
    double s = 0;
    while (i < n) {
      if (i + bs < n) s += g[i] + g[i + bs];
      else s += g[i];
      i += gridSize;
    }
    sdata[tid] = s;
  }
  /* :44-47 tree with __syncthreads, then :12-20 warp tail (only lanes feeding sdata[0] matter) */
  for (unsigned off = bs / 2; off >= 1; off >>= 1) {
    for (unsigned tid = 0; tid < off; tid++) sdata[tid] += sdata[tid + off];
  }
  return sdata[0];
}

double orc_Reduction(const double* in, int n, int threadsPerBlock, int blocksPerGrid) {
  const unsigned bs = (unsigned)threadsPerBlock, nb = (unsigned)blocksPerGrid;
  double* partial = (double*)malloc(sizeof(double) * nb);
#pragma omp parallel
  {
    double* sdata = (double*)malloc(sizeof(double) * bs);
#pragma omp for
    for (int b = 0; b < (int)nb; b++)
      partial[b] = reduce6_block(in, (unsigned)n, bs, (unsigned)b, nb, sdata);
    free(sdata);
  }
  double* sdata = (double*)malloc(sizeof(double) * bs);
  double out = reduce6_block(partial, nb, bs, 0, 1, sdata);
  free(sdata);
  free(partial);
  return out;
}

/* ------------------------------------------------------------------------------------------
 * a10 — ImmerseFlow::ADsolver(), ADSolver.cu:268-395: one predictor step.
 * Order of operations is the reference's, including:
 *   - Compute_velf BEFORE the BC refresh (:298 then :304) — boundary faces see the ghost
 *     values left over from the previous step;
 *   - BC applied to the *current* buffer at :317 and again at :337 just before it is
 *     swapped out, so the buffer that ends up current has ghosts lagging two iterates;
 *   - the dead in-loop Compute_velf (:321);
 *   - stop rule uRes + vRes > pow(10,-6), L1 un-normalised (:315).
 * K == 1 note: the reference's final buffer is then the fresh cudaMalloc'ed temp whose ghost
 * ring was never written (indeterminate).  Deterministic choice here: its ghost ring is a copy
 * of the other buffer's (i.e. BC of the step's starting field).
 * ---------------------------------------------------------------------------------------- */
int orc_ADsolver(int nx, int ny, const double* dx, const double* dy, double dt, double Re,
                 int AD_itermax, const double* iblank,
                 double* u, double* v, double* uf, double* vf,
                 int vf_mode, double* res_hist) {
  return orc_ADsolver_tol(nx, ny, dx, dy, dt, Re, AD_itermax, iblank, u, v, uf, vf, vf_mode, res_hist,
                          pow(10.0, -6.0));                                    /* :315 */
}

/* Same loop with the tolerance as a parameter (the reference hard-codes pow(10,-6)); lets the tests
 * place the tolerance exactly on a residual to exercise the stop rule at the rounding boundary. */
int orc_ADsolver_tol(int nx, int ny, const double* dx, const double* dy, double dt, double Re,
                     int AD_itermax, const double* iblank,
                     double* u, double* v, double* uf, double* vf,
                     int vf_mode, double* res_hist, double tol) {
  const size_t N = (size_t)nx * ny;
  const int tpb = 256, bpg = ((int)N + tpb - 1) / tpb;               /* preSim.cu:190-193 */
  double* mem = (double*)malloc(sizeof(double) * N * 11);
  double *uTemp = mem, *uResidue = mem + N, *vTemp = mem + 2 * N, *vResidue = mem + 3 * N;
  double *c = mem + 4 * N, *cxp = mem + 5 * N, *cxm = mem + 6 * N, *cyp = mem + 7 * N,
         *cym = mem + 8 * N, *sx = mem + 9 * N, *sy = mem + 10 * N;
  double *uc = u, *vc = v;                                           /* Data.{u,v}.velc */

  orc_calculateADCoefficients(nx, ny, dx, dy, dt, Re, c, cxm, cxp, cym, cyp);   /* :289 */
  double uResidual = 1.0, vResidual = 1.0;
  int iter = 0;
  orc_Compute_velf(nx, ny, dx, dy, uc, vc, uf, vf, vf_mode);                    /* :298 */
  orc_set_velocity_BC(nx, ny, uc, vc);                                          /* :304 */
  for (int j = 0; j < ny; j++)          /* K==1 note: temp ghost ring = BC(start field) */
    for (int i = 0; i < nx; i++)
      if (i == 0 || j == 0 || i == nx - 1 || j == ny - 1) {
        uTemp[ID(i, j, nx)] = uc[ID(i, j, nx)];
        vTemp[ID(i, j, nx)] = vc[ID(i, j, nx)];
      }
  orc_ADSource(nx, ny, dx, dy, dt, uc, vc, uf, vf, sx, sy);                     /* :308 */

  while (uResidual + vResidual > tol && iter < AD_itermax) {                     /* :315 */
    orc_set_velocity_BC(nx, ny, uc, vc);                                        /* :317 */
    orc_Compute_velf(nx, ny, dx, dy, uc, vc, uf, vf, vf_mode);                  /* :321 */
    orc_ADsolver_kernel(nx, ny, c, cxm, cxp, cym, cyp, iblank, uc, uTemp, sx);  /* :326 */
    orc_ADsolver_kernel(nx, ny, c, cxm, cxp, cym, cyp, iblank, vc, vTemp, sy);  /* :331 */
    orc_set_velocity_BC(nx, ny, uc, vc);                                        /* :337 */
    double* t = uc; uc = uTemp; uTemp = t;                                      /* :342-348 */
    t = vc; vc = vTemp; vTemp = t;
    orc_Compute_Residual_AD(nx, ny, iblank, uc, uTemp, uResidue);               /* :350 */
    orc_Compute_Residual_AD(nx, ny, iblank, vc, vTemp, vResidue);               /* :355 */
    uResidual = orc_Reduction(uResidue, (int)N, tpb, bpg);                      /* :360 */
    vResidual = orc_Reduction(vResidue, (int)N, tpb, bpg);                      /* :364 */
    if (res_hist) { res_hist[2 * iter] = uResidual; res_hist[2 * iter + 1] = vResidual; }
    iter += 1;
  }
  if (uc != u) {            /* pointer identity changed (odd K): hand the result back in u,v */
    memcpy(u, uc, sizeof(double) * N);
    memcpy(v, vc, sizeof(double) * N);
  }
  free(mem);
  return iter;
}

/* ------------------------------------------------------------------------------------------
 * a11 — calculatePPECoefficients, PPESolver.cu:73-104 (neighbour spacings declared as in
 * ADSolver.cu:27-32 / the shipped sm_52 PTX; the file does not compile as shipped).
 *   cP = -((ax_p + ax_m) + (ay_p + ay_m)), no FMA possible.
 * ---------------------------------------------------------------------------------------- */
void orc_calculatePPECoefficients(int nx, int ny, const double* dx, const double* dy,
                                  double* coeff_ppe, double* dx2_m1, double* dx2_p1,
                                  double* dy2_m1, double* dy2_p1) {
#pragma omp parallel for
  for (int id = 0; id < nx * ny; id++) {
    int i = id % nx, j = id / nx;
    coeff_ppe[id] = 1; dx2_m1[id] = 1; dx2_p1[id] = 1; dy2_m1[id] = 1; dy2_p1[id] = 1;
    if (i > 0 && i < nx - 1 && j > 0 && j < ny - 1) {
      double dx_i = dx[id], dx_ip1 = dx[id + 1], dx_im1 = dx[id - 1];
      double dy_j = dy[id], dy_jp1 = dy[id + nx], dy_jm1 = dy[id - nx];
      double ax_p = 2.0 / (dx_i * (dx_i + dx_ip1));
      double ax_m = 2.0 / (dx_i * (dx_i + dx_im1));
      double ay_p = 2.0 / (dy_j * (dy_j + dy_jp1));
      double ay_m = 2.0 / (dy_j * (dy_j + dy_jm1));
      coeff_ppe[id] = -((ax_p + ax_m) + (ay_p + ay_m));
      dx2_m1[id] = ax_m;
      dx2_p1[id] = ax_p;
      dy2_m1[id] = ay_m;
      dy2_p1[id] = ay_p;
    }
  }
}

/* a14 — set_pressure_BC, PPESolver.cu:54-71: p = 100 on i == 0 and on j == 0. */
void orc_set_pressure_BC(int nx, int ny, double* p) {
  for (int id = 0; id < nx * ny; id++) {
    int i = id % nx, j = id / nx;
    if (i == 0) p[id] = 100.0;
    if (j == 0) p[id] = 100.0;
  }
}

/* a12 — jacobiIteration, PPESolver.cu:13-31.
 * nvcc: t = pW*cW; t = fma(pE,cE,t); t = fma(pN,cN,t); t = fma(pS,cS,t); p_new = (-t)/cP. */
void orc_jacobiIteration(int nx, int ny,
                         const double* coeff_ppe, const double* dx2_m1, const double* dx2_p1,
                         const double* dy2_m1, const double* dy2_p1,
                         const double* p, double* p_new) {
#pragma omp parallel for
  for (int id = 0; id < nx * ny; id++) {
    int i = id % nx, j = id / nx;
    p_new[id] = p[id];
    if (i > 0 && i < nx - 1 && j > 0 && j < ny - 1) {
      double t = p[id - 1] * dx2_m1[id];
      t = fma(p[id + 1], dx2_p1[id], t);
      t = fma(p[id + nx], dy2_p1[id], t);
      t = fma(p[id - nx], dy2_m1[id], t);
      p_new[id] = (-t) / coeff_ppe[id];
    }
  }
}

/* a13 — Compute_Residual, PPESolver.cu:33-50: signed (A p)_ij on interior cells; boundary
 * entries are never written by the reference (cudaMallocManaged => zero): zero here.
 * nvcc: t = pE*cE; t = fma(p,cP,t); fma(pW,cW,t); fma(pN,cN,t); fma(pS,cS,t). */
void orc_Compute_Residual(int nx, int ny,
                          const double* coeff_ppe, const double* dx2_m1, const double* dx2_p1,
                          const double* dy2_m1, const double* dy2_p1,
                          const double* p, double* residual) {
#pragma omp parallel for
  for (int id = 0; id < nx * ny; id++) {
    int i = id % nx, j = id / nx;
    if (i > 0 && i < nx - 1 && j > 0 && j < ny - 1) {
      double t = p[id + 1] * dx2_p1[id];
      t = fma(p[id], coeff_ppe[id], t);
      t = fma(p[id - 1], dx2_m1[id], t);
      t = fma(p[id + nx], dy2_p1[id], t);
      t = fma(p[id - nx], dy2_m1[id], t);
      residual[id] = t;
    }
  }
}

/* ------------------------------------------------------------------------------------------
 * ImmerseFlow::PPESolver(), PPESolver.cu:137-205 (Laplace; the stub source-term call at :155
 * dropped).  BC before the loop (:164) and after (:195); stop rule is the SIGNED sum
 * Residual > pow(10,-6) && iter < PPE_itermax (:172).
 * ---------------------------------------------------------------------------------------- */
int orc_PPESolver(int nx, int ny, const double* dx, const double* dy,
                  int PPE_itermax, double* p, double* final_residual) {
  return orc_PPESolver_tol(nx, ny, dx, dy, PPE_itermax, p, final_residual, pow(10.0, -6.0));   /* :172 */
}

int orc_PPESolver_tol(int nx, int ny, const double* dx, const double* dy,
                      int PPE_itermax, double* p, double* final_residual, double tol) {
  const size_t N = (size_t)nx * ny;
  const int tpb = 256, bpg = ((int)N + tpb - 1) / tpb;
  double* mem = (double*)calloc(N * 7, sizeof(double));
  double *pTemp = mem, *pResidue = mem + N, *cP = mem + 2 * N, *cxp = mem + 3 * N,
         *cxm = mem + 4 * N, *cyp = mem + 5 * N, *cym = mem + 6 * N;
  double* pc = p;
  orc_calculatePPECoefficients(nx, ny, dx, dy, cP, cxm, cxp, cym, cyp);    /* :158 */
  orc_set_pressure_BC(nx, ny, pc);                                         /* :164 */
  double Residual = 1.0;
  int iter = 0;
  while (Residual > tol && iter < PPE_itermax) {                           /* :172 */
    orc_jacobiIteration(nx, ny, cP, cxm, cxp, cym, cyp, pc, pTemp);        /* :174 */
    double* t = pc; pc = pTemp; pTemp = t;                                 /* :178-180 */
    orc_Compute_Residual(nx, ny, cP, cxm, cxp, cym, cyp, pc, pResidue);    /* :182 */
    Residual = orc_Reduction(pResidue, (int)N, tpb, bpg);                  /* :185 */
    iter += 1;
  }
  orc_set_pressure_BC(nx, ny, pc);                                         /* :195 */
  if (pc != p) memcpy(p, pc, sizeof(double) * N);
  if (final_residual) *final_residual = Residual;
  free(mem);
  return iter;
}

/* b — write_results_to_file, postSim.cu:41-66 (Tecplot ASCII POINT, "%f,%f,%f"). */
int orc_write_results_to_file(const double* x, const double* y, const double* data,
                              int ni, int nj, const char* filename) {
  FILE* fp = fopen(filename, "w");
  if (fp == NULL) {
    printf("Error opening file: %s\n", filename);
    return 1;
  }
  fprintf(fp, "TITLE = \"Post Processing Tecplot\"\n");
  fprintf(fp, "VARIABLES = \"X\",\"Y\",\"T\"\n");
  fprintf(fp, "ZONE T=\"BIG ZONE\", I=%d, J=%d, DATAPACKING=POINT\n", ni, nj);
  for (int j = 0; j < nj; j++)
    for (int i = 0; i < ni; i++)
      fprintf(fp, "%f,%f,%f\n", x[i], y[j], data[i + j * ni]);
  fclose(fp);
  return 0;
}
