// multigrid.cuh — geometric multigrid for the pressure-Poisson equation (PPE_Solver 4).
// Semantics: oracle/ifx_oracle_mg.c (PARITY UNPINNED — the reference always runs point Jacobi, PPESolver.cu:13-31).
//
// Level 0 is the solver's grid (padded layout, smoothed by the red-black SOR instantiation of the bulk-copy sweep
// kernel, kernels_v4.cu).  Level l >= 1 merges 2x2 cells of level l-1 (an odd count leaves the last coarse cell of that
// direction with one child) and lives in four dense ghost-inclusive arrays
// of (ncx+2) x (ncy+2) doubles: GE / GN = conductance of a cell's east / north face, e = correction, R = right-hand
// side of the volume-form error equation  sum_f G_f (e_nb - e_C) = R_C.
#pragma once
#include "common.cuh"

#define IFX_MG_MAX_LEVELS 16
#define IFX_MG_NU1 2          // pre-smoothing red-black SOR iterations per level
#define IFX_MG_NU2 2          // post-smoothing
// iterations on the coarsest level: max(32, its cell count), at most 2048 (a grid like 200 x 120 stops at 25 x 15)
inline int ifx_mg_ncoarse(int ncx, int ncy) { const int n = ncx * ncy; return n < 32 ? 32 : (n > 2048 ? 2048 : n); }

// line-relaxation iterations on the coarsest level of the line-smoothed cycle: an eighth of its cells, within [8, 512]
inline int ifx_mg_ncoarse_lines(int ncx, int ncy) { const int n = ncx * ncy / 8; return n < 8 ? 8 : (n > 512 ? 512 : n); }

namespace ifx {

struct MgLevel {
  int ncx, ncy;               // cells (without the ghost ring)
  double* GE; double* GN; double* e; double* R;
  // line smoother (PPE_Solver 5 only, else null): stored elimination of the x-lines and of the y-lines, scratch
  double* inv_x; double* cp_x; double* inv_y; double* cp_y; double* dp;
};

// levels for an ncx x ncy grid: halve (rounding up: where a count is odd the last coarse cell has a single child) down to
// 2 cells in one direction.  Returns the number of levels incl. level 0.
inline int mg_plan(int ncx, int ncy, int* lx, int* ly) {
  int n = 1;
  lx[0] = ncx; ly[0] = ncy;
  while (n < IFX_MG_MAX_LEVELS && lx[n - 1] >= 3 && ly[n - 1] >= 3) {
    lx[n] = (lx[n - 1] + 1) / 2; ly[n] = (ly[n - 1] + 1) / 2;
    n++;
  }
  return n;
}

// kernels_mg.cu.  The GE / GN arrays of the target level must be zero-filled before the two build launches
// (their ghost ring is never written).
// lines = 1: hierarchy of the line-smoothed cycle (every coarse face scaled by 1/2, see the oracle's dir_scale)
cudaError_t launch_mg_build1(const Layout& L, const Metrics& M, const uint8_t* celltype, MgLevel c, int lines, cudaStream_t st);
cudaError_t launch_mg_coarsen(MgLevel f, MgLevel c, int lines, cudaStream_t st);
cudaError_t launch_mg_restrict_fine(const Layout& L, const Metrics& M, const uint8_t* celltype, const double* rhs,
                                    const double* p, MgLevel c, cudaStream_t st);
cudaError_t launch_mg_smooth(MgLevel l, int colour, double omega, cudaStream_t st);
cudaError_t launch_mg_restrict(MgLevel f, MgLevel c, cudaStream_t st);
// bilinear = 0: piecewise-constant prolongation (point-smoothed cycle); 1: bilinear, coarse conductances as connectivity
cudaError_t launch_mg_prolong(MgLevel c, MgLevel f, int bilinear, cudaStream_t st);
cudaError_t launch_mg_prolong_fine(const Layout& L, const uint8_t* celltype, MgLevel c, double* p, int bilinear, cudaStream_t st);
// zebra line relaxation (PPE_Solver 2; smoother of PPE_Solver 5).  dir 0 = lines along x (rows), dir 1 = along y
// (columns).  factor: elimination of the matrix of every line of one direction (depends on geometry and cell types
// only) into inv_a / cp_a, fields in the layout of p; solve: one pass over the lines of one parity, dpw = scratch field.
cudaError_t launch_line_factor(const Layout& L, const Metrics& M, const uint8_t* celltype, int dir, double* inv_a, double* cp_a,
                               cudaStream_t st);
cudaError_t launch_line_solve(const Layout& L, const Metrics& M, const uint8_t* celltype, const double* rhs, double* p,
                              const double* inv_a, const double* cp_a, double* dpw, int dir, int parity, double omega,
                              cudaStream_t st);
cudaError_t launch_mg_line_factor(MgLevel l, int dir, cudaStream_t st);
cudaError_t launch_mg_line_solve(MgLevel l, int dir, int parity, double omega, cudaStream_t st);

}  // namespace ifx
