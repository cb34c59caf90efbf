// solver.h — host-side state of one slab (one GPU).  Replaces `struct ImmerseFlow`
// (reference src/header/globalVariables.cuh:70-88); owns every device allocation for the life of
// the handle (the reference mallocs/frees 11 N-sized arrays per step, ADSolver.cu:275-287,382-394).
#pragma once
#include "../../include/immerseflow_c.h"
#include "kernels.cuh"
#include "multigrid.cuh"

#include <string>
#include <vector>

struct GhostCells {
  int count = 0, capacity = 0;
  int* cell = nullptr;        // padded-layout offset of each ghost cell (sorted by reference id)
  int* ref_id = nullptr;      // reference id = i + j*nx
  int* stencil = nullptr;     // 4 padded offsets per ghost cell
  int* stencil_ref = nullptr; // 4 reference ids
  double* w_dir = nullptr;    // 4 weights + 1 constant multiplier, Dirichlet closure (u, v)
  double* w_neu = nullptr;    // 4 weights, Neumann closure (p)
  double* bi = nullptr;       // body intercept x,y
  double* ip = nullptr;       // image point x,y
  int* body = nullptr;        // owning body index
};

struct ifx_solver {
  ifx_input in{};
  ifx_options opt{};
  ifx::Layout L{};
  ifx::Metrics M{};
  int device = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  bool initialized = false;

  size_t field_elems = 0;            // doubles per padded field allocation
  double* u[2] = {nullptr, nullptr};
  double* v[2] = {nullptr, nullptr};
  double* p[3] = {nullptr, nullptr, nullptr};   // slab runs rotate THREE pressure buffers (lagged stop decision, capi.cu)
  int np = 2;                        // pressure buffers in use
  int nseg_fields = 6;               // fields in the exchange segment: u[2], v[2], p[np]
  int cur_uv = 0, cur_p = 0;
  double *sx = nullptr, *sy = nullptr, *rhs = nullptr, *uf = nullptr, *vf = nullptr;
  uint8_t* celltype = nullptr;
  uint8_t* facemask = nullptr;       // open faces of the general Poisson operator, derived from the cell types (FULL mode)
  bool facemask_valid = false;
  bool faces_valid = false;          // uf/vf hold projected face velocities (FULL mode)

  // reduction scratch
  double* partials = nullptr;        // 2 doubles per CTA
  size_t partials_cap = 0;
  ifx::LoopCtl* ctl = nullptr;       // device
  ifx::LoopCtl* h_ctl = nullptr;     // pinned host mirror
  int* h_counters = nullptr;         // pinned host mirror of d_counters (4 ints)
  unsigned char* h_stage = nullptr;  // pinned staging for the body arrays (zero-copy control path)
  size_t h_stage_bytes = 0;
  double *res_a = nullptr, *res_b = nullptr;   // reference-layout residual arrays (lazy)
  double* red_partial = nullptr;     // level-1 partials for the reference-order reduction
  double* red_out = nullptr;         // 2 doubles

  std::vector<double*> tables;       // device 1-D metric tables (freed in destroy)
  std::vector<double> h_xc, h_yc, h_dx, h_dy;

  // immersed boundary
  int nbodies = 0;
  std::vector<int> h_body_off;
  std::vector<double> h_xm, h_ym, h_ub, h_vb, h_bbox;
  int* d_body_off = nullptr;
  double *d_xm = nullptr, *d_ym = nullptr, *d_bbox = nullptr;
  size_t markers_cap = 0;
  double *d_ub = nullptr, *d_vb = nullptr;   // 64 entries each, always allocated in FULL mode
  bool bodies_dirty = false;
  bool has_gc = false;
  GhostCells gc;
  int* d_counters = nullptr;         // [0] ghost-cell total, [1] stencil out of reach of the neighbour slab
  int* d_rowcount = nullptr;         // nyl entries
  int* d_rowstart = nullptr;         // nyl + 1 entries
  double *gc_tmp_a = nullptr, *gc_tmp_b = nullptr;   // gather arrays for in-place ghost-cell refresh
  bool state_bc_fresh = false;       // ring + ghost cells of u, v consistent with the interior

  // exchange segment: u[0..1], v[0..1], p[0..1] + XchgSync in ONE allocation (IPC-exported to the other ranks)
  double* seg = nullptr;
  size_t seg_bytes = 0, sync_off = 0;
  ifx::XchgSync* sync = nullptr;     // inside seg
  void* peer_seg[IFX_MAX_RANKS] = {};          // mapped segments (own rank: seg)
  size_t peer_field_elems[IFX_MAX_RANKS] = {};
  int peer_nyl[IFX_MAX_RANKS] = {};
  bool connected = false;
  unsigned seq[IFX_SYNC_GROUPS] = {0, 0, 0};   // launches per sync group (identical on every rank)
  unsigned mseq = 0;                           // residual-mailbox tag

  // geometric multigrid (PPE_Solver 4): coarse levels 1 .. mg_levels-1, allocated on first use
  int mg_levels = 0;
  ifx::MgLevel mg[IFX_MG_MAX_LEVELS] = {};
  bool mg_valid = false;             // the conductances match the current cell types (cleared by ifx_iblank_update)
  cudaGraphExec_t mg_graph = nullptr;   // ifx_options.use_graphs: the coarse part of the point-smoothed V-cycle, captured once
  long long mg_graph_launches = 0;      // kernels one replay stands for
  // fine level of the line relaxation (PPE_Solver 2, 5): stored elimination of the x-lines and of the y-lines
  // (valid for the current cell types: cleared with mg_valid) and the solve's scratch; fields in the layout of p
  double* line_f[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};   // inv_x, cp_x, inv_y, cp_y, dp
  bool line_factor_valid = false;

  int rows_override = 0;             // IFX_ROWS_PER_CTA
  int last_ad_iters = 5;
  double ad_hist[2 * 64] = {};       // (uRes, vRes) per iteration of the last predictor solve
  cudaEvent_t ev[10] = {};           // stage timers; [8], [9]: the Jacobi sweeps of the predictor alone
  long long launches = 0;
  std::string err;
};

namespace ifx {
int fail(ifx_solver* s, int code, const std::string& msg);
// shared by capi.cu and capi_full.cu
int fetch_ctl(ifx_solver* s);
// failure detection: a NaN / Inf residual ends every stop rule as "converged" (NaN > tol is false, in the
// reference too, ADSolver.cu:315); the loops report it instead.  Looks at the control block fetched last.
int check_residual_finite(ifx_solver* s, const char* stage);
int fetch_small(ifx_solver* s, void* host_pinned, const void* dev, size_t bytes);   // D2H of a few bytes + stream sync
int ensure_partials(ifx_solver* s, size_t nblocks);
int ensure_exact_buffers(ifx_solver* s);
int exact_decide(ifx_solver* s, const ReduceCfg& rc, bool two_arrays);
bool ppe_wide_tiles(const ifx_solver* s);
int rows_per_cta_for(const ifx_solver* s, int mode);
dim3 tile_grid(const ifx_solver* s, int ry, int mode);
double rounding_band(const ifx_solver* s, size_t nblocks, int rows_per_cta);
void fill_bc(const ifx_solver* s, double* two_u, double* two_v);
int run_ad_loop(ifx_solver* s, ifx_step_stats* st, bool full);
int run_ppe_loop(ifx_solver* s, ifx_step_stats* st, bool laplace_ref);
int enqueue_ppe_sweep(ifx_solver* s, PpeSweepArgs& a, dim3 grid, bool laplace_ref, bool write_res);   // counts the launch
int ensure_facemask(ifx_solver* s);     // (re)derives the face masks after the cell types changed
int run_ppe_multigrid(ifx_solver* s, ifx_step_stats* st);      // capi_mg.cu
int run_ppe_lines(ifx_solver* s, ifx_step_stats* st);
int full_refresh_velocity_bc(ifx_solver* s, int buf);
int full_refresh_pressure_bc(ifx_solver* s, int buf);
// slabs
void make_halo_ctx(ifx_solver* s, int group, int nfields, const int* out_field_index, HaloCtx* hx);
int halo_exchange(ifx_solver* s, int group, int nfields, const int* field_index, int tile_cols, bool gc_flags = false);
int halo_push(ifx_solver* s, int group, int nfields, const int* field_index, int tile_cols, unsigned seq,
              const LoopCtl* ctl, int iter, bool gc_flags = false);
bool bodies_on_slabs(const ifx_solver* s);
GcPeers gc_peers(ifx_solver* s, int f0, int f1);
int halo_wait(ifx_solver* s, int group, unsigned need, int tile_cols);
}

#define IFX_CUDA(s, call)                                                                      \
  do {                                                                                         \
    cudaError_t e__ = (call);                                                                  \
    if (e__ != cudaSuccess)                                                                    \
      return ifx::fail((s), IFX_ERR_CUDA,                                                      \
                       std::string(#call) + ": " + cudaGetErrorString(e__) + " (" + __FILE__ + \
                           ":" + std::to_string(__LINE__) + ")");                              \
  } while (0)

#define IFX_LAUNCH_CHECK(s)                   \
  do {                                        \
    (s)->launches++;                          \
    IFX_CUDA((s), cudaGetLastError());        \
  } while (0)
