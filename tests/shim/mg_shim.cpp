// mg_shim.cpp — TEST INFRASTRUCTURE: C entry points over the host-shim build of kernels_mg.cu (see cuda_host_shim.h).
// All arrays are host memory; fine fields use the product's padded layout.
#include "cuda_host_shim.h"
thread_local dim3 threadIdx, blockIdx, blockDim, gridDim;
#include "../../immerseflow_b200/csrc/kernels_mg.cu"

using namespace ifx;

extern "C" {
int shim_mg_plan(int ncx, int ncy, int* lx, int* ly) { return mg_plan(ncx, ncy, lx, ly); }

// metrics: 8 tables in this order: dx, dy (nx / ny entries), pp_cE, pp_cW, pp_sx (nx), pp_cN, pp_cS, pp_sy (ny)
static Metrics metrics_of(const double* const* t) {
  Metrics M{};
  M.dx = t[0]; M.dy = t[1]; M.pp_cE = t[2]; M.pp_cW = t[3]; M.pp_sx = t[4]; M.pp_cN = t[5]; M.pp_cS = t[6]; M.pp_sy = t[7];
  return M;
}
static Layout layout_of(int nx, int ny, int pitch) { return Layout{nx, ny, pitch, ny, 0, 1, ny - 1}; }
static MgLevel level_of(int ncx, int ncy, double* GE, double* GN, double* e, double* R, double* scratch5 = nullptr) {
  MgLevel l{ncx, ncy, GE, GN, e, R, nullptr, nullptr, nullptr, nullptr, nullptr};
  if (scratch5) {
    const size_t n = (size_t)(ncx + 2) * (ncy + 2);
    l.inv_x = scratch5; l.cp_x = scratch5 + n; l.inv_y = scratch5 + 2 * n; l.cp_y = scratch5 + 3 * n; l.dp = scratch5 + 4 * n;
  }
  return l;
}

void shim_mg_build1(int nx, int ny, int pitch, const double* const* tables, const uint8_t* ct, int ncx, int ncy, double* GE, double* GN, int lines) {
  launch_mg_build1(layout_of(nx, ny, pitch), metrics_of(tables), ct, level_of(ncx, ncy, GE, GN, nullptr, nullptr), lines, nullptr);
}
void shim_mg_coarsen(int fx, int fy, double* fGE, double* fGN, int cx, int cy, double* GE, double* GN, int lines) {
  launch_mg_coarsen(level_of(fx, fy, fGE, fGN, nullptr, nullptr), level_of(cx, cy, GE, GN, nullptr, nullptr), lines, nullptr);
}
// one pass = factor of that direction (into scratch) + solve; the product factors once per set of cell types
void shim_line_pass(int nx, int ny, int pitch, const double* const* tables, const uint8_t* ct, const double* rhs, double* p,
                    double* inv_a, double* cp_a, double* dpw, int dir, int parity, double omega) {
  launch_line_factor(layout_of(nx, ny, pitch), metrics_of(tables), ct, dir, inv_a, cp_a, nullptr);
  launch_line_solve(layout_of(nx, ny, pitch), metrics_of(tables), ct, rhs, p, inv_a, cp_a, dpw, dir, parity, omega, nullptr);
}
void shim_mg_line_pass(int ncx, int ncy, double* GE, double* GN, double* e, double* R, double* scratch5, int dir, int parity,
                       double omega) {
  MgLevel l = level_of(ncx, ncy, GE, GN, e, R, scratch5);
  launch_mg_line_factor(l, dir, nullptr);
  launch_mg_line_solve(l, dir, parity, omega, nullptr);
}
void shim_mg_restrict_fine(int nx, int ny, int pitch, const double* const* tables, const uint8_t* ct, const double* rhs,
                           const double* p, int ncx, int ncy, double* R) {
  launch_mg_restrict_fine(layout_of(nx, ny, pitch), metrics_of(tables), ct, rhs, p, level_of(ncx, ncy, nullptr, nullptr, nullptr, R), nullptr);
}
void shim_mg_smooth(int ncx, int ncy, double* GE, double* GN, double* e, double* R, int colour, double omega) {
  launch_mg_smooth(level_of(ncx, ncy, GE, GN, e, R), colour, omega, nullptr);
}
void shim_mg_restrict(int fx, int fy, double* GE, double* GN, double* e, double* R, int cx, int cy, double* Rc) {
  launch_mg_restrict(level_of(fx, fy, GE, GN, e, R), level_of(cx, cy, nullptr, nullptr, nullptr, Rc), nullptr);
}
void shim_mg_prolong(int cx, int cy, double* cGE, double* cGN, double* ec, int fx, int fy, double* GE, double* GN, double* e, int bilinear) {
  launch_mg_prolong(level_of(cx, cy, cGE, cGN, ec, nullptr), level_of(fx, fy, GE, GN, e, nullptr), bilinear, nullptr);
}
void shim_mg_prolong_fine(int nx, int ny, int pitch, const uint8_t* ct, int cx, int cy, double* cGE, double* cGN, double* e1, double* p,
                          int bilinear) {
  launch_mg_prolong_fine(layout_of(nx, ny, pitch), ct, level_of(cx, cy, cGE, cGN, e1, nullptr), p, bilinear, nullptr);
}
}
