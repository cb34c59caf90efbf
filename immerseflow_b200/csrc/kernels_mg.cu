// kernels_mg.cu — transfer and coarse-level kernels of the geometric multigrid V-cycle (PPE_Solver 4).
// Semantics and operation order: oracle/ifx_oracle_mg.c (PARITY UNPINNED); compiled -fmad=false, every fused
// multiply-add is explicit, so each level is bit-identical to the oracle.
//
// One thread per (coarse) cell, no shared memory, no inter-thread communication: the whole file also compiles as
// plain C++ under tests/cuda_host_shim.h, where the CPU test suite runs every kernel against the oracle.
// HBM traffic per V-cycle beyond the fine-level smoothing (25 B/cell per half-sweep, kernels_v4.cu):
// residual + restriction 17 B per fine cell read + 2 B written, prolongation 17 B read + 8 B written; all coarse
// levels together hold 1/3 of the fine cells.
#include "multigrid.cuh"
#include "stencil_math.cuh"

#ifdef IFX_HOST_SHIM
#define IFX_KLAUNCH(k, grid, block, st, ...) (shim_launch((grid), (block), [&] { k(__VA_ARGS__); }), cudaSuccess)
#else
#define IFX_KLAUNCH(k, grid, block, st, ...) (k<<<(grid), (block), 0, (st)>>>(__VA_ARGS__), cudaGetLastError())
#endif

namespace ifx {

constexpr int MG_BX = 32, MG_BY = 8;

static inline dim3 mg_grid(int ncx, int ncy) { return dim3((ncx + MG_BX - 1) / MG_BX, (ncy + MG_BY - 1) / MG_BY, 1); }

// index into a level's ghost-inclusive array
__device__ __forceinline__ size_t cidx(int I, int J, int NX) { return (size_t)I + (size_t)J * NX; }

// see dir_scale in the oracle: the factor 1/2 of a coarse face only in a direction the point smoother smooths
__device__ __forceinline__ void mg_dir_scale(int lines, double sx, double sy, double& ax, double& ay) {
  if (lines) { ax = 0.5; ay = 0.5; return; }       // alternating line relaxation smooths both directions everywhere
  ax = (sx >= 0.25 * sy) ? 0.5 : 1.0;
  ay = (sy >= 0.25 * sx) ? 0.5 : 1.0;
}

// ---------------------------------------------------------------------------------------------
// level-1 conductances from the grid metrics and the cell types (oracle: build_level1)
// ---------------------------------------------------------------------------------------------
static __global__ void k_mg_build1(Layout L, Metrics M, const uint8_t* __restrict__ ct, MgLevel c, int lines) {
  const int I = 1 + blockIdx.x * blockDim.x + threadIdx.x, J = 1 + blockIdx.y * blockDim.y + threadIdx.y;
  if (I > c.ncx || J > c.ncy) return;
  const int i = 2 * I, j = 2 * J;               // children: columns i-1, i; rows j-1, j
  double ge[2] = {0.0, 0.0}, gn[2] = {0.0, 0.0};
#pragma unroll
  for (int k = 0; k < 2; k++) {
    const int jj = j - 1 + k, ii = i - 1 + k;       // the second child row / column may not exist (odd count)
    const size_t oe = lidx(L, i, jj - L.j0), on = lidx(L, ii, j - L.j0);
    if (i <= L.nx - 3 && jj <= L.ny - 2 && ct[oe] == IFX_FLUID && ct[oe + 1] == IFX_FLUID)
      ge[k] = (2.0 * M.dy[jj]) / (M.dx[i] + M.dx[i + 1]);
    if (j <= L.ny - 3 && ii <= L.nx - 2 && ct[on] == IFX_FLUID && ct[on + L.pitch] == IFX_FLUID)
      gn[k] = (2.0 * M.dx[ii]) / (M.dy[j] + M.dy[j + 1]);
  }
  const double sx = ge[0] + ge[1], sy = gn[0] + gn[1];
  double ax, ay;
  mg_dir_scale(lines, sx, sy, ax, ay);
  const size_t o = cidx(I, J, c.ncx + 2);
  c.GE[o] = ax * sx;
  c.GN[o] = ay * sy;
}

// level l+1 conductances from level l (oracle: build_coarser)
static __global__ void k_mg_coarsen(MgLevel f, MgLevel c, int lines) {
  const int I = 1 + blockIdx.x * blockDim.x + threadIdx.x, J = 1 + blockIdx.y * blockDim.y + threadIdx.y;
  if (I > c.ncx || J > c.ncy) return;
  const int i = 2 * I, j = 2 * J, nxl = f.ncx + 2;
  const double sx = f.GE[cidx(i, j - 1, nxl)] + f.GE[cidx(i, j, nxl)];
  const double sy = f.GN[cidx(i - 1, j, nxl)] + f.GN[cidx(i, j, nxl)];
  double ax, ay;
  mg_dir_scale(lines, sx, sy, ax, ay);
  const size_t o = cidx(I, J, c.ncx + 2);
  c.GE[o] = ax * sx;
  c.GN[o] = ay * sy;
}

// ---------------------------------------------------------------------------------------------
// R1 = sum over the four children of (rhs - A p) dx_i dy_j, fluid children only (oracle: orc_mg_restrict_fine);
// (A p) with the closed-face substitutions of the general Poisson sweep
// ---------------------------------------------------------------------------------------------
static __global__ void k_mg_restrict_fine(Layout L, Metrics M, const uint8_t* __restrict__ ct, const double* __restrict__ rhs,
                                          const double* __restrict__ p, MgLevel c) {
  const int I = 1 + blockIdx.x * blockDim.x + threadIdx.x, J = 1 + blockIdx.y * blockDim.y + threadIdx.y;
  if (I > c.ncx || J > c.ncy) return;
  double r[4];
#pragma unroll
  for (int k = 0; k < 4; k++) {
    const int i = 2 * I - 1 + (k & 1), j = 2 * J - 1 + (k >> 1);
    const size_t o = lidx(L, i, j - L.j0);
    r[k] = 0.0;
    if (i > L.nx - 2 || j > L.ny - 2 || ct[o] != IFX_FLUID) continue;      // child beyond the grid (odd count)
    const double pc = p[o];
    const double pw = (i == 1 || ct[o - 1] != IFX_FLUID) ? pc : p[o - 1];
    const double pe = (i == L.nx - 2 || ct[o + 1] != IFX_FLUID) ? pc : p[o + 1];
    const double ps = (j == 1 || ct[o - L.pitch] != IFX_FLUID) ? pc : p[o - L.pitch];
    const double pn = (j == L.ny - 2 || ct[o + L.pitch] != IFX_FLUID) ? pc : p[o + L.pitch];
    const double cP = -(M.pp_sx[i] + M.pp_sy[j]);                                       // PPESolver.cu:93-94
    const double q = ppe_apply(pc, cP, pw, M.pp_cW[i], pe, M.pp_cE[i], pn, M.pp_cN[j], ps, M.pp_cS[j]);
    r[k] = (rhs[o] - q) * (M.dx[i] * M.dy[j]);
  }
  c.R[cidx(I, J, c.ncx + 2)] = (r[0] + r[1]) + (r[2] + r[3]);
}

// ---------------------------------------------------------------------------------------------
// one colour of red-black SOR on a coarse level, in place: cells of one colour are not neighbours
// (oracle: orc_mg_smooth)
// ---------------------------------------------------------------------------------------------
static __global__ void k_mg_smooth(MgLevel l, int colour, double omega) {
  const int I = 1 + blockIdx.x * blockDim.x + threadIdx.x, J = 1 + blockIdx.y * blockDim.y + threadIdx.y;
  if (I > l.ncx || J > l.ncy || ((I + J + colour) & 1)) return;
  const int NX = l.ncx + 2;
  const size_t o = cidx(I, J, NX);
  const double ge = l.GE[o], gw = l.GE[o - 1], gn = l.GN[o], gs = l.GN[o - NX];
  const double D = (ge + gw) + (gn + gs);
  if (!(D > 0.0)) return;
  double t = ge * l.e[o + 1];
  t = fma(gw, l.e[o - 1], t);
  t = fma(gn, l.e[o + NX], t);
  t = fma(gs, l.e[o - NX], t);
  const double ej = (t - l.R[o]) / D;
  const double ec = l.e[o];
  l.e[o] = ec + omega * (ej - ec);
}

// R_{l+1} = sum over the four children of (R - (sum G e_nb - D e)) (oracle: orc_mg_restrict)
static __global__ void k_mg_restrict(MgLevel f, MgLevel c) {
  const int I = 1 + blockIdx.x * blockDim.x + threadIdx.x, J = 1 + blockIdx.y * blockDim.y + threadIdx.y;
  if (I > c.ncx || J > c.ncy) return;
  const int nxl = f.ncx + 2;
  double r[4];
#pragma unroll
  for (int k = 0; k < 4; k++) {
    const size_t o = cidx(2 * I - 1 + (k & 1), 2 * J - 1 + (k >> 1), nxl);
    const double ge = f.GE[o], gw = f.GE[o - 1], gn = f.GN[o], gs = f.GN[o - nxl];
    const double D = (ge + gw) + (gn + gs);
    r[k] = 0.0;
    if (!(D > 0.0)) continue;
    double t = ge * f.e[o + 1];
    t = fma(gw, f.e[o - 1], t);
    t = fma(gn, f.e[o + nxl], t);
    t = fma(gs, f.e[o - nxl], t);
    r[k] = f.R[o] - fma(-D, f.e[o], t);
  }
  c.R[cidx(I, J, c.ncx + 2)] = (r[0] + r[1]) + (r[2] + r[3]);
}

// Bilinear prolongation of the line-smoothed cycle (oracle: prolong_value): parent 9/16, the two nearer coarse neighbours
// 3/16 each, the diagonal one 1/16; a neighbour not connected to the parent through open coarse faces counts as the parent.
__device__ __forceinline__ double mg_prolong_value(int i, int j, const MgLevel& c, int bilinear) {
  const int NX = c.ncx + 2;
  const int I = (i + 1) / 2, J = (j + 1) / 2;
  const size_t P = cidx(I, J, NX);
  const double eP = c.e[P];
  if (!bilinear) return eP;
  const int di = (i & 1) ? -1 : 1, dj = (j & 1) ? -1 : 1;
  const size_t A = P + di, B = P + (ptrdiff_t)dj * NX, C = A + (ptrdiff_t)dj * NX;
  const double gPA = c.GE[di > 0 ? P : A], gPB = c.GN[dj > 0 ? P : B];
  const double gAC = c.GN[dj > 0 ? A : C], gBC = c.GE[di > 0 ? B : C];
  const double eA = (gPA > 0.0) ? c.e[A] : eP;
  const double eB = (gPB > 0.0) ? c.e[B] : eP;
  const double eC = ((gPA > 0.0 && gAC > 0.0) || (gPB > 0.0 && gBC > 0.0)) ? c.e[C] : eP;
  return 0.0625 * ((9.0 * eP + 3.0 * eA) + (3.0 * eB + eC));
}

// e_l += P e_{l+1} on the active cells of level l (oracle: orc_mg_prolong / orc_mg_prolong2)
static __global__ void k_mg_prolong(MgLevel c, MgLevel f, int bilinear) {
  const int i = 1 + blockIdx.x * blockDim.x + threadIdx.x, j = 1 + blockIdx.y * blockDim.y + threadIdx.y;
  if (i > f.ncx || j > f.ncy) return;
  const int nxl = f.ncx + 2;
  const size_t o = cidx(i, j, nxl);
  const double D = (f.GE[o] + f.GE[o - 1]) + (f.GN[o] + f.GN[o - nxl]);
  if (!(D > 0.0)) return;
  f.e[o] = f.e[o] + mg_prolong_value(i, j, c, bilinear);
}

// p += P e_1 on fluid cells (oracle: orc_mg_prolong_fine / orc_mg_prolong_fine2)
static __global__ void k_mg_prolong_fine(Layout L, const uint8_t* __restrict__ ct, MgLevel c, double* __restrict__ p, int bilinear) {
  const int i = 1 + blockIdx.x * blockDim.x + threadIdx.x, j = 1 + blockIdx.y * blockDim.y + threadIdx.y;
  if (i > L.nx - 2 || j > L.ny - 2) return;
  const size_t o = lidx(L, i, j - L.j0);
  if (ct[o] != IFX_FLUID) return;
  p[o] = p[o] + mg_prolong_value(i, j, c, bilinear);
}

// ---------------------------------------------------------------------------------------------
// Zebra line relaxation (oracle: orc_ppe_line_pass / orc_mg_line_pass).  One thread per line, Thomas elimination in
// increasing index order with one reciprocal per cell, back substitution, relaxation — the oracle's operations in the
// oracle's order, split in two kernels:
//   * FACTOR (once per set of cell types): the elimination of the matrix itself — inv_k = 1/(dg_k - lo_k cp_{k-1}),
//     cp_k = up_k inv_k — depends on the geometry and the cell types only, so it is stored per direction and
//     reused by every pass of every iteration (same values, computed once: the loop-carried division, ~200 cycles
//     per cell, leaves the hot loop).  A real row has inv < 0 (the pivot is negative); identity rows (non-fluid
//     cells, isolated fluid cells) are stored as inv = 1, cp = 0, which is exactly what the oracle's arithmetic
//     makes of them — so the sign of inv tells the solve kernel which kind of row it is.
//   * SOLVE (every pass): right-hand side, dp_k = (d_k - lo_k dp_{k-1}) inv_k, back substitution, relaxation.  The
//     loads of LINE_U consecutive cells are issued before the dependent chain of those cells runs (the first cut,
//     one load-use round trip per cell with a warp or two per SM, was latency-bound at ~1 us per cell:
//     profiles/r1_line_mg_bench_first_cut.jsonl).
// Lines along y put consecutive threads on consecutive columns (coalesced); lines along x stride by a row per thread
// and rely on each 128-byte line serving 16 consecutive cells of the same thread from L1/L2.
// ---------------------------------------------------------------------------------------------
constexpr int LINE_THREADS = 32;
constexpr int LINE_U = 8;
constexpr int LINE_AHEAD = 4 * LINE_U;      // cells ahead of the chain whose cache lines are requested early

// a hint, not a load: the line is on its way into L1/L2 when the loads of LINE_AHEAD cells later ask for it
__device__ __forceinline__ void line_prefetch(const void* a) {
#ifndef IFX_HOST_SHIM
  asm volatile("prefetch.global.L1 [%0];" ::"l"(a));
#else
  (void)a;
#endif
}

template <int DIR>
static __global__ void k_line_factor(Layout L, Metrics M, const uint8_t* __restrict__ ct, double* __restrict__ inv_a,
                                     double* __restrict__ cp_a) {
  const int nline = DIR == 0 ? L.ny : L.nx, len = DIR == 0 ? L.nx : L.ny;
  const int l = 1 + (int)(blockIdx.x * blockDim.x + threadIdx.x);
  if (l > nline - 2) return;
  const size_t sk = DIR == 0 ? 1 : (size_t)L.pitch;
  const size_t o0 = DIR == 0 ? lidx(L, 0, l - L.j0) : lidx(L, l, 0 - L.j0);
  const int nxm2 = L.nx - 2, nym2 = L.ny - 2;
  double cprev = 0.0;
  for (int k = 1; k < len - 1; k++) {
    const size_t o = o0 + (size_t)k * sk;
    const int i = DIR == 0 ? k : l, j = DIR == 0 ? l : k;
    double lo = 0.0, up = 0.0, dg = 1.0;
    if (ct[o] == IFX_FLUID) {
      const bool oW = !(i == 1 || ct[o - 1] != IFX_FLUID), oE = !(i == nxm2 || ct[o + 1] != IFX_FLUID);
      const bool oS = !(j == 1 || ct[o - L.pitch] != IFX_FLUID), oN = !(j == nym2 || ct[o + L.pitch] != IFX_FLUID);
      const double cW = M.pp_cW[i], cE = M.pp_cE[i], cS = M.pp_cS[j], cN = M.pp_cN[j];
      dg = -(M.pp_sx[i] + M.pp_sy[j]);                                                   // cP, PPESolver.cu:93-94
      if (!oW) dg = dg + cW;
      if (!oE) dg = dg + cE;
      if (!oS) dg = dg + cS;
      if (!oN) dg = dg + cN;
      if (DIR == 0) { lo = oW ? cW : 0.0; up = oE ? cE : 0.0; }
      else { lo = oS ? cS : 0.0; up = oN ? cN : 0.0; }
      const double piv = fma(-lo, cprev, dg);
      if (!(piv < 0.0)) { lo = 0.0; up = 0.0; dg = 1.0; }                                 // isolated cell: identity row
    }
    const double inv = 1.0 / fma(-lo, cprev, dg);
    cprev = up * inv;
    inv_a[o] = inv; cp_a[o] = cprev;
  }
}

template <int DIR>
static __global__ void k_line_solve(Layout L, Metrics M, const uint8_t* __restrict__ ct, const double* __restrict__ rhs,
                                    double* p, const double* __restrict__ inv_a, const double* __restrict__ cp_a,
                                    double* __restrict__ dpw, int parity, double omega) {
  const int nline = DIR == 0 ? L.ny : L.nx, len = DIR == 0 ? L.nx : L.ny;
  const int l = (parity ? 1 : 2) + 2 * (int)(blockIdx.x * blockDim.x + threadIdx.x);
  if (l > nline - 2) return;
  const size_t sk = DIR == 0 ? 1 : (size_t)L.pitch;       // along the line
  const size_t sc = DIR == 0 ? (size_t)L.pitch : 1;       // across: the neighbouring lines (other parity, untouched)
  const size_t o0 = DIR == 0 ? lidx(L, 0, l - L.j0) : lidx(L, l, 0 - L.j0);
  // the cross-direction stencil of the whole line: north / south for x-lines, east / west for y-lines
  const bool hi_inside = DIR == 0 ? (l != L.ny - 2) : (l != L.nx - 2), lo_inside = (l != 1);
  const double c_hi = DIR == 0 ? M.pp_cN[l] : M.pp_cE[l], c_lo = DIR == 0 ? M.pp_cS[l] : M.pp_cW[l];
  const double* __restrict__ c_along = DIR == 0 ? M.pp_cW : M.pp_cS;      // the sub-diagonal coefficient, by cell index
  double dprev = 0.0;
  bool prev_fluid = false;                                 // cell k-1 of the line (k = 1: the ring, a closed face)
  for (int k0 = 1; k0 < len - 1; k0 += LINE_U) {
    const int n = min(LINE_U, len - 1 - k0);
    double v_inv[LINE_U], v_rhs[LINE_U], v_phi[LINE_U], v_plo[LINE_U], v_pc[LINE_U], v_ca[LINE_U];
    uint8_t v_ct[LINE_U], v_chi[LINE_U], v_clo[LINE_U];
    if (k0 + LINE_AHEAD + LINE_U <= len - 1) {             // one thread owns the whole line: too few threads to hide DRAM latency
#pragma unroll
      for (int u = 0; u < LINE_U; u += (DIR == 0 ? 4 : 1)) {   // x-lines: 4 doubles per 32-byte sector; y-lines: a line per cell
        const size_t o = o0 + (size_t)(k0 + LINE_AHEAD + u) * sk;
        line_prefetch(inv_a + o); line_prefetch(rhs + o); line_prefetch(p + o); line_prefetch(p + o + sc); line_prefetch(p + o - sc);
      }
    }
#pragma unroll
    for (int u = 0; u < LINE_U; u++) {
      if (u < n) {
        const size_t o = o0 + (size_t)(k0 + u) * sk;
        v_inv[u] = inv_a[o]; v_rhs[u] = rhs[o]; v_pc[u] = p[o];
        v_phi[u] = p[o + sc]; v_plo[u] = p[o - sc];
        v_ct[u] = ct[o]; v_chi[u] = ct[o + sc]; v_clo[u] = ct[o - sc];
        v_ca[u] = c_along[k0 + u];
      }
    }
#pragma unroll
    for (int u = 0; u < LINE_U; u++) {
      if (u < n) {
        const int k = k0 + u;
        const bool fluid = v_ct[u] == IFX_FLUID;
        const bool real_row = v_inv[u] < 0.0;              // fluid and not isolated (see FACTOR)
        double lo = 0.0, d = v_pc[u];
        if (real_row) {
          lo = (k != 1 && prev_fluid) ? v_ca[u] : 0.0;
          d = v_rhs[u];
          if (hi_inside && v_chi[u] == IFX_FLUID) d = fma(-c_hi, v_phi[u], d);
          if (lo_inside && v_clo[u] == IFX_FLUID) d = fma(-c_lo, v_plo[u], d);
        }
        dprev = fma(-lo, dprev, d) * v_inv[u];
        dpw[o0 + (size_t)k * sk] = dprev;
        prev_fluid = fluid;
      }
    }
  }
  double xnext = 0.0;
  for (int k1 = len - 2; k1 >= 1; k1 -= LINE_U) {
    const int n = min(LINE_U, k1);
    double v_cp[LINE_U], v_dp[LINE_U], v_pc[LINE_U];
    uint8_t v_ct[LINE_U];
    if (k1 - LINE_AHEAD - LINE_U >= 0) {
#pragma unroll
      for (int u = 0; u < LINE_U; u += (DIR == 0 ? 4 : 1)) {
        const size_t o = o0 + (size_t)(k1 - LINE_AHEAD - u) * sk;
        line_prefetch(cp_a + o); line_prefetch(dpw + o); line_prefetch(p + o);
      }
    }
#pragma unroll
    for (int u = 0; u < LINE_U; u++) {
      if (u < n) {
        const size_t o = o0 + (size_t)(k1 - u) * sk;
        v_cp[u] = cp_a[o]; v_dp[u] = dpw[o]; v_pc[u] = p[o]; v_ct[u] = ct[o];
      }
    }
#pragma unroll
    for (int u = 0; u < LINE_U; u++) {
      if (u < n) {
        const double x = fma(-v_cp[u], xnext, v_dp[u]);
        xnext = x;
        if (v_ct[u] == IFX_FLUID) p[o0 + (size_t)(k1 - u) * sk] = v_pc[u] + omega * (x - v_pc[u]);
      }
    }
  }
}

template <int DIR>
static __global__ void k_mg_line_factor(MgLevel lv) {
  const int NX = lv.ncx + 2, NY = lv.ncy + 2;
  const int nline = DIR == 0 ? NY : NX, len = DIR == 0 ? NX : NY;
  const int l = 1 + (int)(blockIdx.x * blockDim.x + threadIdx.x);
  if (l > nline - 2) return;
  const size_t sk = DIR == 0 ? 1 : (size_t)NX, sl = DIR == 0 ? (size_t)NX : 1;
  double* __restrict__ inv_a = DIR == 0 ? lv.inv_x : lv.inv_y;
  double* __restrict__ cp_a = DIR == 0 ? lv.cp_x : lv.cp_y;
  double cprev = 0.0;
  for (int k = 1; k < len - 1; k++) {
    const size_t o = (size_t)l * sl + (size_t)k * sk;
    const double ge = lv.GE[o], gw = lv.GE[o - 1], gn = lv.GN[o], gs = lv.GN[o - NX];
    const double D = (ge + gw) + (gn + gs);
    double lo = 0.0, up = 0.0, dg = 1.0;
    if (D > 0.0) {
      dg = -D;
      if (DIR == 0) { lo = gw; up = ge; } else { lo = gs; up = gn; }
      const double piv = fma(-lo, cprev, dg);
      if (!(piv < 0.0)) { lo = 0.0; up = 0.0; dg = 1.0; }
    }
    const double inv = 1.0 / fma(-lo, cprev, dg);
    cprev = up * inv;
    inv_a[o] = inv; cp_a[o] = cprev;
  }
}

template <int DIR>
static __global__ void k_mg_line_solve(MgLevel lv, int parity, double omega) {
  const int NX = lv.ncx + 2, NY = lv.ncy + 2;
  const int nline = DIR == 0 ? NY : NX, len = DIR == 0 ? NX : NY;
  const int l = (parity ? 1 : 2) + 2 * (int)(blockIdx.x * blockDim.x + threadIdx.x);
  if (l > nline - 2) return;
  const size_t sk = DIR == 0 ? 1 : (size_t)NX, sl = DIR == 0 ? (size_t)NX : 1, sc = sl;
  const double* __restrict__ inv_a = DIR == 0 ? lv.inv_x : lv.inv_y;
  const double* __restrict__ cp_a = DIR == 0 ? lv.cp_x : lv.cp_y;
  // conductances towards the next / previous cell ACROSS the line and towards the previous cell ALONG it:
  //   x-lines: across = GN(o), GN(o - NX), along = GE(o - 1);   y-lines: across = GE(o), GE(o - 1), along = GN(o - NX)
  const double* __restrict__ g_across = DIR == 0 ? lv.GN : lv.GE;
  const double* __restrict__ g_along = DIR == 0 ? lv.GE : lv.GN;
  double dprev = 0.0;
  for (int k0 = 1; k0 < len - 1; k0 += LINE_U) {
    const int n = min(LINE_U, len - 1 - k0);
    double v_inv[LINE_U], v_R[LINE_U], v_ehi[LINE_U], v_elo[LINE_U], v_ec[LINE_U], v_ghi[LINE_U], v_glo[LINE_U], v_ga[LINE_U];
    if (k0 + LINE_AHEAD + LINE_U <= len - 1) {
#pragma unroll
      for (int u = 0; u < LINE_U; u += (DIR == 0 ? 4 : 1)) {
        const size_t o = (size_t)l * sl + (size_t)(k0 + LINE_AHEAD + u) * sk;
        line_prefetch(inv_a + o); line_prefetch(lv.R + o); line_prefetch(lv.e + o); line_prefetch(lv.e + o + sc); line_prefetch(lv.e + o - sc);
        line_prefetch(g_across + o); line_prefetch(g_across + o - sc); line_prefetch(g_along + o - sk);
      }
    }
#pragma unroll
    for (int u = 0; u < LINE_U; u++) {
      if (u < n) {
        const size_t o = (size_t)l * sl + (size_t)(k0 + u) * sk;
        v_inv[u] = inv_a[o]; v_R[u] = lv.R[o]; v_ec[u] = lv.e[o];
        v_ehi[u] = lv.e[o + sc]; v_elo[u] = lv.e[o - sc];
        v_ghi[u] = g_across[o]; v_glo[u] = g_across[o - sc]; v_ga[u] = g_along[o - sk];
      }
    }
#pragma unroll
    for (int u = 0; u < LINE_U; u++) {
      if (u < n) {
        double lo = 0.0, d = v_ec[u];
        if (v_inv[u] < 0.0) {                             // active cell that is not isolated along the line
          lo = v_ga[u];
          d = fma(-v_ghi[u], v_ehi[u], v_R[u]);
          d = fma(-v_glo[u], v_elo[u], d);
        }
        dprev = fma(-lo, dprev, d) * v_inv[u];
        lv.dp[(size_t)l * sl + (size_t)(k0 + u) * sk] = dprev;
      }
    }
  }
  double xnext = 0.0;
  for (int k1 = len - 2; k1 >= 1; k1 -= LINE_U) {
    const int n = min(LINE_U, k1);
    double v_cp[LINE_U], v_dp[LINE_U], v_ec[LINE_U], v_ge[LINE_U], v_gw[LINE_U], v_gn[LINE_U], v_gs[LINE_U];
    if (k1 - LINE_AHEAD - LINE_U >= 0) {
#pragma unroll
      for (int u = 0; u < LINE_U; u += (DIR == 0 ? 4 : 1)) {
        const size_t o = (size_t)l * sl + (size_t)(k1 - LINE_AHEAD - u) * sk;
        line_prefetch(cp_a + o); line_prefetch(lv.dp + o); line_prefetch(lv.e + o); line_prefetch(lv.GE + o); line_prefetch(lv.GN + o);
      }
    }
#pragma unroll
    for (int u = 0; u < LINE_U; u++) {
      if (u < n) {
        const size_t o = (size_t)l * sl + (size_t)(k1 - u) * sk;
        v_cp[u] = cp_a[o]; v_dp[u] = lv.dp[o]; v_ec[u] = lv.e[o];
        v_ge[u] = lv.GE[o]; v_gw[u] = lv.GE[o - 1]; v_gn[u] = lv.GN[o]; v_gs[u] = lv.GN[o - NX];
      }
    }
#pragma unroll
    for (int u = 0; u < LINE_U; u++) {
      if (u < n) {
        const double x = fma(-v_cp[u], xnext, v_dp[u]);
        xnext = x;
        const double D = (v_ge[u] + v_gw[u]) + (v_gn[u] + v_gs[u]);
        if (D > 0.0) lv.e[(size_t)l * sl + (size_t)(k1 - u) * sk] = v_ec[u] + omega * (x - v_ec[u]);
      }
    }
  }
}

static inline dim3 line_grid(int nline) {      // lines of one parity among 1 .. nline-2
  const int n = (nline - 2 + 1) / 2;
  return dim3((n + LINE_THREADS - 1) / LINE_THREADS, 1, 1);
}

// ---------------------------------------------------------------------------------------------
static inline dim3 all_lines_grid(int nline) { return dim3((nline - 2 + LINE_THREADS - 1) / LINE_THREADS, 1, 1); }

cudaError_t launch_line_factor(const Layout& L, const Metrics& M, const uint8_t* celltype, int dir, double* inv_a, double* cp_a,
                               cudaStream_t st) {
  if (dir == 0) return IFX_KLAUNCH(k_line_factor<0>, all_lines_grid(L.ny), dim3(LINE_THREADS, 1, 1), st, L, M, celltype, inv_a, cp_a);
  return IFX_KLAUNCH(k_line_factor<1>, all_lines_grid(L.nx), dim3(LINE_THREADS, 1, 1), st, L, M, celltype, inv_a, cp_a);
}
cudaError_t launch_line_solve(const Layout& L, const Metrics& M, const uint8_t* celltype, const double* rhs, double* p,
                              const double* inv_a, const double* cp_a, double* dpw, int dir, int parity, double omega,
                              cudaStream_t st) {
  if (dir == 0)
    return IFX_KLAUNCH(k_line_solve<0>, line_grid(L.ny), dim3(LINE_THREADS, 1, 1), st, L, M, celltype, rhs, p, inv_a, cp_a, dpw, parity, omega);
  return IFX_KLAUNCH(k_line_solve<1>, line_grid(L.nx), dim3(LINE_THREADS, 1, 1), st, L, M, celltype, rhs, p, inv_a, cp_a, dpw, parity, omega);
}
cudaError_t launch_mg_line_factor(MgLevel l, int dir, cudaStream_t st) {
  if (dir == 0) return IFX_KLAUNCH(k_mg_line_factor<0>, all_lines_grid(l.ncy + 2), dim3(LINE_THREADS, 1, 1), st, l);
  return IFX_KLAUNCH(k_mg_line_factor<1>, all_lines_grid(l.ncx + 2), dim3(LINE_THREADS, 1, 1), st, l);
}
cudaError_t launch_mg_line_solve(MgLevel l, int dir, int parity, double omega, cudaStream_t st) {
  if (dir == 0) return IFX_KLAUNCH(k_mg_line_solve<0>, line_grid(l.ncy + 2), dim3(LINE_THREADS, 1, 1), st, l, parity, omega);
  return IFX_KLAUNCH(k_mg_line_solve<1>, line_grid(l.ncx + 2), dim3(LINE_THREADS, 1, 1), st, l, parity, omega);
}
cudaError_t launch_mg_build1(const Layout& L, const Metrics& M, const uint8_t* celltype, MgLevel c, int lines, cudaStream_t st) {
  return IFX_KLAUNCH(k_mg_build1, mg_grid(c.ncx, c.ncy), dim3(MG_BX, MG_BY, 1), st, L, M, celltype, c, lines);
}
cudaError_t launch_mg_coarsen(MgLevel f, MgLevel c, int lines, cudaStream_t st) {
  return IFX_KLAUNCH(k_mg_coarsen, mg_grid(c.ncx, c.ncy), dim3(MG_BX, MG_BY, 1), st, f, c, lines);
}
cudaError_t launch_mg_restrict_fine(const Layout& L, const Metrics& M, const uint8_t* celltype, const double* rhs,
                                    const double* p, MgLevel c, cudaStream_t st) {
  return IFX_KLAUNCH(k_mg_restrict_fine, mg_grid(c.ncx, c.ncy), dim3(MG_BX, MG_BY, 1), st, L, M, celltype, rhs, p, c);
}
cudaError_t launch_mg_smooth(MgLevel l, int colour, double omega, cudaStream_t st) {
  return IFX_KLAUNCH(k_mg_smooth, mg_grid(l.ncx, l.ncy), dim3(MG_BX, MG_BY, 1), st, l, colour, omega);
}
cudaError_t launch_mg_restrict(MgLevel f, MgLevel c, cudaStream_t st) {
  return IFX_KLAUNCH(k_mg_restrict, mg_grid(c.ncx, c.ncy), dim3(MG_BX, MG_BY, 1), st, f, c);
}
cudaError_t launch_mg_prolong(MgLevel c, MgLevel f, int bilinear, cudaStream_t st) {
  return IFX_KLAUNCH(k_mg_prolong, mg_grid(f.ncx, f.ncy), dim3(MG_BX, MG_BY, 1), st, c, f, bilinear);
}
cudaError_t launch_mg_prolong_fine(const Layout& L, const uint8_t* celltype, MgLevel c, double* p, int bilinear, cudaStream_t st) {
  return IFX_KLAUNCH(k_mg_prolong_fine, mg_grid(L.nx - 2, L.ny - 2), dim3(MG_BX, MG_BY, 1), st, L, celltype, c, p, bilinear);
}

}  // namespace ifx
