// capi_full.cu — C-ABI entry points of the stages the reference lacks (IFX_COMPAT_FULL) and the dispatch of
// ifx_ad_solve / ifx_ppe_solve / ifx_step between the two compat modes.
#include "solver.h"

#include <nvtx3/nvToolsExt.h>      // header-only: ranges are no-ops unless a profiler injects its library

#include <algorithm>
#include <cstring>

using namespace ifx;

// ------------------------------------------------------------------------------------------------
// boundary-condition refresh of a velocity / pressure buffer pair: ghost ring + ghost cells, the ghost cells
// gathered from the field before any of them is overwritten (oracle: apply_velocity_bc / apply_pressure_bc)
// ------------------------------------------------------------------------------------------------
// Slabs with bodies: a stencil may reach into the neighbour's rows, so the gather is fenced pairwise (halo rows
// delivered and the neighbours' fields complete before they are read; nobody scatters before everybody has
// gathered) and the boundary rows are re-delivered afterwards, ghost cells included.
static int slab_barrier(ifx_solver* s) { return halo_exchange(s, 2, 0, nullptr, 256); }

int ifx::full_refresh_velocity_bc(ifx_solver* s, int buf) {
  double tu[4], tv[4];
  fill_bc(s, tu, tv);
  const bool bs = bodies_on_slabs(s);
  int rc;
  s->launches += 2;
  IFX_CUDA(s, launch_apply_ring(s->L, s->u[buf], tu, 0, s->stream));
  IFX_CUDA(s, launch_apply_ring(s->L, s->v[buf], tv, 0, s->stream));
  const int fi[2] = {buf, 2 + buf};
  if (bs && (rc = halo_exchange(s, 2, 2, fi, 256)) != IFX_OK) return rc;     // fresh halo rows; also the first barrier
  if (s->has_gc) {
    s->launches++;
    IFX_CUDA(s, launch_gc_velocity(s->gc.count, s->gc.cell, s->gc.stencil, s->gc.w_dir, s->gc.body, s->d_ub, s->d_vb,
                                   s->u[buf], s->v[buf], gc_peers(s, buf, 2 + buf), s->gc_tmp_a, s->gc_tmp_b, 1, nullptr, 0,
                                   s->stream));
  }
  if (bs && (rc = slab_barrier(s)) != IFX_OK) return rc;
  if (s->has_gc) {
    s->launches++;
    IFX_CUDA(s, launch_gc_scatter(s->gc.count, s->gc.cell, s->gc_tmp_a, s->u[buf], s->gc_tmp_b, s->v[buf], s->stream));
  }
  if (bs && (rc = halo_exchange(s, 2, 2, fi, 256)) != IFX_OK) return rc;
  return IFX_OK;
}

int ifx::full_refresh_pressure_bc(ifx_solver* s, int buf) {
  const bool bs = bodies_on_slabs(s);
  int rc;
  s->launches++;
  IFX_CUDA(s, launch_apply_ring(s->L, s->p[buf], nullptr, 1, s->stream));
  // diagnostic only: the solver never reads p at ghost cells (closed-face rule)
  if (bs && (rc = slab_barrier(s)) != IFX_OK) return rc;
  if (s->has_gc) {
    s->launches++;
    IFX_CUDA(s, launch_gc_pressure(s->gc.count, s->gc.cell, s->gc.stencil, s->gc.w_neu, s->p[buf], gc_peers(s, 4 + buf, -1),
                                   s->gc_tmp_a, 1, s->stream));
  }
  if (bs && (rc = slab_barrier(s)) != IFX_OK) return rc;
  if (s->has_gc) {
    s->launches++;
    IFX_CUDA(s, launch_gc_scatter(s->gc.count, s->gc.cell, s->gc_tmp_a, s->p[buf], nullptr, nullptr, s->stream));
  }
  return IFX_OK;
}

// NVTX range per stage (the reference has no tracing at all): `ncu --nvtx --nvtx-include "ifx:Poisson/"` etc.
struct StageRange {
  explicit StageRange(const char* name) { nvtxRangePushA(name); }
  ~StageRange() { nvtxRangePop(); }
};

static int require_full(ifx_solver* s, const char* what) {
  if (s->opt.compat != IFX_COMPAT_FULL)
    return fail(s, IFX_ERR_INVALID, std::string(what) + " requires IFX_COMPAT_FULL (the reference has no such stage)");
  return IFX_OK;
}

// ------------------------------------------------------------------------------------------------
// immersed bodies
// ------------------------------------------------------------------------------------------------
extern "C" int ifx_set_bodies(ifx_solver* s, int nbodies, const int* offsets, const double* xm, const double* ym,
                              const double* ubody, const double* vbody) {
  if (!s) return IFX_ERR_INVALID;
  int rc = require_full(s, "ifx_set_bodies");
  if (rc != IFX_OK) return rc;
  if (nbodies < 0 || nbodies > 63) return fail(s, IFX_ERR_INVALID, "at most 63 bodies (6-bit owner field in the cell type)");
  if (nbodies > 0 && (!offsets || !xm || !ym)) return fail(s, IFX_ERR_INVALID, "null body arrays");
  if (nbodies > 0 && s->opt.nranks > 1 && s->L.je - s->L.jb < 2 * IFX_GC_REACH)
    return fail(s, IFX_ERR_INVALID, "immersed bodies on slabs need at least 8 rows per slab (ghost-cell stencils reach up to "
                                    "4 rows into the neighbour)");
  IFX_CUDA(s, cudaSetDevice(s->device));
  const int nm = nbodies ? offsets[nbodies] : 0;
  for (int b = 0; b < nbodies; b++)
    if (offsets[b + 1] - offsets[b] < 3) return fail(s, IFX_ERR_INVALID, "a body needs at least 3 markers");
  s->nbodies = nbodies;
  s->h_body_off.assign(offsets, offsets + nbodies + 1);
  if (nbodies == 0) s->h_body_off.assign(1, 0);
  s->h_xm.assign(xm, xm + nm);
  s->h_ym.assign(ym, ym + nm);
  s->h_ub.assign(64, 0.0);
  s->h_vb.assign(64, 0.0);
  for (int b = 0; b < nbodies; b++) {
    if (ubody) s->h_ub[b] = ubody[b];
    if (vbody) s->h_vb[b] = vbody[b];
  }
  s->h_bbox.assign(4 * std::max(nbodies, 1), 0.0);
  for (int b = 0; b < nbodies; b++) {
    double x0 = xm[offsets[b]], x1 = x0, y0 = ym[offsets[b]], y1 = y0;
    for (int k = offsets[b]; k < offsets[b + 1]; k++) {
      x0 = std::min(x0, xm[k]); x1 = std::max(x1, xm[k]);
      y0 = std::min(y0, ym[k]); y1 = std::max(y1, ym[k]);
    }
    s->h_bbox[4 * b] = x0; s->h_bbox[4 * b + 1] = x1; s->h_bbox[4 * b + 2] = y0; s->h_bbox[4 * b + 3] = y1;
  }
  // A body that reaches the outermost cell centres would put ghost cells next to the grid boundary, whose image points
  // leave the grid: the bilinear closure then extrapolates with weights of order 100 and the predictor blows up
  // (seen on the oracle).  Refuse instead of computing garbage.
  for (int b = 0; b < nbodies; b++) {
    const double* bb = s->h_bbox.data() + 4 * b;
    if (!(bb[0] > s->h_xc[1] && bb[1] < s->h_xc[s->L.nx - 2] && bb[2] > s->h_yc[1] && bb[3] < s->h_yc[s->L.ny - 2])) {
      s->nbodies = 0;
      s->h_body_off.assign(1, 0);
      s->bodies_dirty = true;
      return fail(s, IFX_ERR_INVALID, "body " + std::to_string(b) + " reaches the outermost cells of the grid (or lies outside): immersed "
                                      "bodies must stay inside the first / last interior cell centres");
    }
  }
  if ((size_t)nm + 1 > s->markers_cap) {
    if (s->d_xm) cudaFree(s->d_xm);
    if (s->d_ym) cudaFree(s->d_ym);
    s->d_xm = s->d_ym = nullptr;
    s->markers_cap = (size_t)nm + 1024;
    IFX_CUDA(s, cudaMalloc(&s->d_xm, sizeof(double) * s->markers_cap));
    IFX_CUDA(s, cudaMalloc(&s->d_ym, sizeof(double) * s->markers_cap));
  }
  if (!s->d_body_off) IFX_CUDA(s, cudaMalloc(&s->d_body_off, sizeof(int) * 65));
  if (!s->d_bbox) IFX_CUDA(s, cudaMalloc(&s->d_bbox, sizeof(double) * 4 * 64));
  // uploads: body offsets, markers, bounding boxes, body velocities
  struct Up { void* dst; const void* src; size_t bytes; };
  const Up ups[] = {{s->d_body_off, s->h_body_off.data(), sizeof(int) * s->h_body_off.size()},
                    {s->d_xm, s->h_xm.data(), sizeof(double) * nm},
                    {s->d_ym, s->h_ym.data(), sizeof(double) * nm},
                    {s->d_bbox, s->h_bbox.data(), sizeof(double) * s->h_bbox.size()},
                    {s->d_ub, s->h_ub.data(), sizeof(double) * 64},
                    {s->d_vb, s->h_vb.data(), sizeof(double) * 64}};
  if (s->opt.zero_copy_control) {
    // staged in page-locked memory and pulled by a kernel (no copy engine: see ifx_options.zero_copy_control)
    size_t need = 0;
    for (const Up& u : ups) need += (u.bytes + 15) / 16 * 16;
    if (need > s->h_stage_bytes) {
      if (s->h_stage) cudaFreeHost(s->h_stage);
      s->h_stage = nullptr; s->h_stage_bytes = 0;
      IFX_CUDA(s, cudaMallocHost(&s->h_stage, need + 65536));
      s->h_stage_bytes = need + 65536;
    }
    size_t off = 0;
    for (const Up& u : ups) {
      if (u.bytes == 0) continue;
      std::memcpy(s->h_stage + off, u.src, u.bytes);
      s->launches++;
      IFX_CUDA(s, launch_copy_words(u.dst, s->h_stage + off, u.bytes, s->stream));
      off += (u.bytes + 15) / 16 * 16;
    }
  } else {
    for (const Up& u : ups)
      if (u.bytes) IFX_CUDA(s, cudaMemcpyAsync(u.dst, u.src, u.bytes, cudaMemcpyHostToDevice, s->stream));
  }
  IFX_CUDA(s, cudaStreamSynchronize(s->stream));      // host vectors / the staging buffer may be rewritten by the next call
  s->bodies_dirty = true;
  return IFX_OK;
}

static int ensure_gc_capacity(ifx_solver* s, int n) {
  GhostCells& g = s->gc;
  if (n <= g.capacity) return IFX_OK;
  int* ip[] = {g.cell, g.ref_id, g.stencil, g.stencil_ref, g.body};
  for (int* p : ip) if (p) cudaFree(p);
  double* dp[] = {g.w_dir, g.w_neu, g.bi, g.ip, s->gc_tmp_a, s->gc_tmp_b};
  for (double* p : dp) if (p) cudaFree(p);
  const int cap = n + n / 4 + 1024;
  g = GhostCells();
  s->gc_tmp_a = s->gc_tmp_b = nullptr;
  IFX_CUDA(s, cudaMalloc(&g.cell, sizeof(int) * cap));
  IFX_CUDA(s, cudaMalloc(&g.ref_id, sizeof(int) * cap));
  IFX_CUDA(s, cudaMalloc(&g.body, sizeof(int) * cap));
  IFX_CUDA(s, cudaMalloc(&g.stencil, sizeof(int) * 4 * cap));
  IFX_CUDA(s, cudaMalloc(&g.stencil_ref, sizeof(int) * 4 * cap));
  IFX_CUDA(s, cudaMalloc(&g.w_dir, sizeof(double) * 5 * cap));
  IFX_CUDA(s, cudaMalloc(&g.w_neu, sizeof(double) * 4 * cap));
  IFX_CUDA(s, cudaMalloc(&g.bi, sizeof(double) * 2 * cap));
  IFX_CUDA(s, cudaMalloc(&g.ip, sizeof(double) * 2 * cap));
  IFX_CUDA(s, cudaMalloc(&s->gc_tmp_a, sizeof(double) * cap));
  IFX_CUDA(s, cudaMalloc(&s->gc_tmp_b, sizeof(double) * cap));
  g.capacity = cap;
  return IFX_OK;
}

// replaces iBlankComputeKernel (preSim.cu:110-136) + the ghost-cell machinery the reference lacks
extern "C" int ifx_iblank_update(ifx_solver* s, ifx_step_stats* st) {
  if (!s) return IFX_ERR_INVALID;
  StageRange range("ifx:iblank+ghost-cells");
  int rc = require_full(s, "ifx_iblank_update");
  if (rc != IFX_OK) return rc;
  IFX_CUDA(s, cudaSetDevice(s->device));
  const Layout& L = s->L;
  IFX_CUDA(s, cudaEventRecord(s->ev[6], s->stream));
  if (!s->d_rowcount) {
    IFX_CUDA(s, cudaMalloc(&s->d_rowcount, sizeof(int) * (L.nyl + 1)));
    IFX_CUDA(s, cudaMalloc(&s->d_rowstart, sizeof(int) * (L.nyl + 2)));
    IFX_CUDA(s, cudaMalloc(&s->d_counters, sizeof(int) * 4));
  }
  if (!s->d_body_off) {   // no bodies were ever set: everything is fluid
    int zero[65] = {0};
    IFX_CUDA(s, cudaMalloc(&s->d_body_off, sizeof(int) * 65));
    IFX_CUDA(s, cudaMalloc(&s->d_bbox, sizeof(double) * 4 * 64));
    IFX_CUDA(s, cudaMemcpy(s->d_body_off, zero, sizeof(zero), cudaMemcpyHostToDevice));
  }
  const bool slabs = s->opt.nranks > 1;
  if (slabs && s->nbodies > 0 && !s->connected)
    return fail(s, IFX_ERR_STATE, "slab run: call ifx_ipc_connect before the bodies are classified (the ghost-cell stencils "
                                  "address the neighbours' rows)");
  BodySet B{s->nbodies, s->d_body_off, s->d_xm, s->d_ym, s->d_bbox};
  SlabGeom sg{};
  if (slabs && s->connected) {
    sg.has_lo = s->opt.rank > 0; sg.has_hi = s->opt.rank < s->opt.nranks - 1;
    if (sg.has_lo) sg.nyl_lo = s->peer_nyl[s->opt.rank - 1];
    if (sg.has_hi) sg.nyl_hi = s->peer_nyl[s->opt.rank + 1];
    if ((size_t)std::max(sg.nyl_lo, sg.nyl_hi) * L.pitch >= (1u << 30))
      return fail(s, IFX_ERR_INVALID, "neighbour slab too large for the 30-bit remote stencil index");
  }
  IFX_CUDA(s, cudaMemsetAsync(s->d_counters, 0, sizeof(int) * 4, s->stream));
  s->launches += 2;
  IFX_CUDA(s, launch_classify(L, s->M.xc, s->M.yc, B, s->celltype, s->d_rowcount, s->stream));
  IFX_CUDA(s, launch_gc_count(L, s->d_rowcount, s->d_rowstart, s->d_counters, s->stream));
  if ((rc = fetch_small(s, s->h_counters, s->d_counters, sizeof(int) * 4)) != IFX_OK) return rc;
  const int total = s->h_counters[0];
  if ((rc = ensure_gc_capacity(s, total)) != IFX_OK) return rc;
  s->gc.count = total;
  s->has_gc = total > 0;
  if (total > 0) {
    s->launches += 2;
    IFX_CUDA(s, launch_gc_build(L, s->M.xc, s->M.yc, B, sg, s->celltype, s->d_rowstart, total,
                                s->gc.cell, s->gc.ref_id, s->gc.body, s->gc.stencil, s->gc.stencil_ref, s->gc.w_dir,
                                s->gc.w_neu, s->gc.bi, s->gc.ip, s->d_counters + 1, s->stream));
    if ((rc = fetch_small(s, s->h_counters, s->d_counters, sizeof(int) * 4)) != IFX_OK) return rc;
    const int err = s->h_counters[1];
    if (err) return fail(s, IFX_ERR_INVALID, "a ghost-cell stencil reaches more than 4 rows into (or beyond) the neighbour slab: "
                                             "use fewer ranks for this grid");
  }
  s->bodies_dirty = false;
  s->facemask_valid = false;       // the open faces of the Poisson operator moved with the bodies
  s->mg_valid = false;             // so did the open faces the coarse-level conductances are made of
  s->line_factor_valid = false;    // and the matrices of the line relaxation
  s->faces_valid = false;          // closed faces moved with the bodies
  s->state_bc_fresh = false;
  IFX_CUDA(s, cudaEventRecord(s->ev[7], s->stream));
  IFX_CUDA(s, cudaEventSynchronize(s->ev[7]));
  if (st) cudaEventElapsedTime(&st->ms_ib, s->ev[6], s->ev[7]);
  return IFX_OK;
}

extern "C" int ifx_ghost_cell_count(const ifx_solver* s) { return s ? s->gc.count : 0; }

// weights: 10 doubles per ghost cell = {wd0..wd3, cd, wn0..wn3, body}
extern "C" int ifx_get_ghost_cells(ifx_solver* s, int* cell_id, int* stencil_id, double* weights, double* bi_xy,
                                   double* ip_xy, int capacity) {
  if (!s) return IFX_ERR_INVALID;
  IFX_CUDA(s, cudaSetDevice(s->device));
  const int n = s->gc.count;
  if (capacity < n) return fail(s, IFX_ERR_INVALID, "capacity too small");
  if (n == 0) return IFX_OK;
  IFX_CUDA(s, cudaStreamSynchronize(s->stream));
  if (cell_id) IFX_CUDA(s, cudaMemcpy(cell_id, s->gc.ref_id, sizeof(int) * n, cudaMemcpyDeviceToHost));
  if (stencil_id) IFX_CUDA(s, cudaMemcpy(stencil_id, s->gc.stencil_ref, sizeof(int) * 4 * n, cudaMemcpyDeviceToHost));
  if (bi_xy) IFX_CUDA(s, cudaMemcpy(bi_xy, s->gc.bi, sizeof(double) * 2 * n, cudaMemcpyDeviceToHost));
  if (ip_xy) IFX_CUDA(s, cudaMemcpy(ip_xy, s->gc.ip, sizeof(double) * 2 * n, cudaMemcpyDeviceToHost));
  if (weights) {
    std::vector<double> wd(5 * (size_t)n), wn(4 * (size_t)n);
    std::vector<int> body(n);
    IFX_CUDA(s, cudaMemcpy(wd.data(), s->gc.w_dir, sizeof(double) * 5 * n, cudaMemcpyDeviceToHost));
    IFX_CUDA(s, cudaMemcpy(wn.data(), s->gc.w_neu, sizeof(double) * 4 * n, cudaMemcpyDeviceToHost));
    IFX_CUDA(s, cudaMemcpy(body.data(), s->gc.body, sizeof(int) * n, cudaMemcpyDeviceToHost));
    for (int g = 0; g < n; g++) {
      std::memcpy(weights + 10 * (size_t)g, wd.data() + 5 * (size_t)g, 5 * sizeof(double));
      std::memcpy(weights + 10 * (size_t)g + 5, wn.data() + 4 * (size_t)g, 4 * sizeof(double));
      weights[10 * (size_t)g + 9] = body[g];
    }
  }
  return IFX_OK;
}

// ------------------------------------------------------------------------------------------------
// stage entry points
// ------------------------------------------------------------------------------------------------
static int full_prepare(ifx_solver* s, ifx_step_stats* st = nullptr) {
  int rc;
  if (s->bodies_dirty && (rc = ifx_iblank_update(s, st)) != IFX_OK) return rc;
  return IFX_OK;
}

// replaces ImmerseFlow::ADsolver(), ADSolver.cu:268-395
extern "C" int ifx_ad_solve(ifx_solver* s, ifx_step_stats* st) {
  if (!s) return IFX_ERR_INVALID;
  StageRange range("ifx:predictor");
  IFX_CUDA(s, cudaSetDevice(s->device));
  if (s->opt.compat == IFX_COMPAT_REFERENCE) return run_ad_loop(s, st, false);
  int rc = require_full(s, "predictor");
  if (rc != IFX_OK) return rc;
  if ((rc = full_prepare(s, st)) != IFX_OK) return rc;
  // start-of-step consistency: ring + ghost cells of (u, v); first step: face velocities from the cells
  if (s->connected && s->opt.nranks > 1) {
    const int fi[2] = {s->cur_uv, 2 + s->cur_uv};
    if ((rc = halo_exchange(s, 2, 2, fi, 256)) != IFX_OK) return rc;
  }
  if ((rc = full_refresh_velocity_bc(s, s->cur_uv)) != IFX_OK) return rc;
  if (!s->faces_valid) {
    s->launches++;
    IFX_CUDA(s, launch_faces_init(s->L, s->M, s->celltype, s->d_ub, s->d_vb, s->u[s->cur_uv], s->v[s->cur_uv], s->uf, s->vf,
                                  s->stream));
    s->faces_valid = true;
  }
  if ((rc = run_ad_loop(s, st, true)) != IFX_OK) return rc;
  if ((rc = full_refresh_velocity_bc(s, s->cur_uv)) != IFX_OK) return rc;
  IFX_CUDA(s, cudaStreamSynchronize(s->stream));
  return IFX_OK;
}

// replaces ImmerseFlow::PPESolver(), PPESolver.cu:137-205
extern "C" int ifx_ppe_solve(ifx_solver* s, ifx_step_stats* st) {
  if (!s) return IFX_ERR_INVALID;
  StageRange range("ifx:Poisson");
  IFX_CUDA(s, cudaSetDevice(s->device));
  if (s->opt.compat == IFX_COMPAT_REFERENCE) return run_ppe_loop(s, st, true);
  int rc = require_full(s, "Poisson solve");
  if (rc != IFX_OK) return rc;
  if ((rc = full_prepare(s)) != IFX_OK) return rc;
  // a15: source term from the predicted velocities (their ring and ghost cells were refreshed by the predictor)
  s->launches++;
  IFX_CUDA(s, launch_ppe_rhs(s->L, s->M, s->celltype, s->d_ub, s->d_vb, s->u[s->cur_uv], s->v[s->cur_uv], s->rhs, s->stream));
  const int ps = s->opt.ppe_solver;
  rc = ps == 4 ? run_ppe_multigrid(s, st) : (ps == 2 || ps == 5) ? run_ppe_lines(s, st) : run_ppe_loop(s, st, false);
  if (rc != IFX_OK) return rc;
  if ((rc = full_refresh_pressure_bc(s, s->cur_p)) != IFX_OK) return rc;
  IFX_CUDA(s, cudaStreamSynchronize(s->stream));
  return IFX_OK;
}

// the projection the reference leaves empty, AD_PPE_Correction.cu:1-12
extern "C" int ifx_correct(ifx_solver* s, ifx_step_stats* st) {
  if (!s) return IFX_ERR_INVALID;
  StageRange range("ifx:projection");
  int rc = require_full(s, "ifx_correct");
  if (rc != IFX_OK) return rc;
  IFX_CUDA(s, cudaSetDevice(s->device));
  if ((rc = full_prepare(s)) != IFX_OK) return rc;
  IFX_CUDA(s, cudaEventRecord(s->ev[4], s->stream));
  const int cur = s->cur_uv;
  s->launches++;
  IFX_CUDA(s, launch_correct(s->L, s->M, s->celltype, s->d_ub, s->d_vb, s->u[cur], s->v[cur], s->p[s->cur_p], s->u[cur ^ 1],
                             s->v[cur ^ 1], s->uf, s->vf, s->stream));
  s->cur_uv = cur ^ 1;
  s->faces_valid = true;
  if ((rc = full_refresh_velocity_bc(s, s->cur_uv)) != IFX_OK) return rc;
  IFX_CUDA(s, cudaEventRecord(s->ev[5], s->stream));
  IFX_CUDA(s, cudaEventSynchronize(s->ev[5]));
  if (st) cudaEventElapsedTime(&st->ms_correct, s->ev[4], s->ev[5]);
  return IFX_OK;
}

extern "C" int ifx_get_residual_history(ifx_solver* s, double* pairs, int capacity) {
  if (!s || !pairs || capacity <= 0) return 0;
  const int n = std::min(std::min(s->last_ad_iters, 64), capacity);
  std::memcpy(pairs, s->ad_hist, sizeof(double) * 2 * n);
  return n;
}

// one iteration of main()'s time loop (main.cu:93-96)
extern "C" int ifx_step(ifx_solver* s, ifx_step_stats* st) {
  if (!s) return IFX_ERR_INVALID;
  ifx_step_stats local;
  if (!st) st = &local;
  std::memset(st, 0, sizeof(*st));
  int rc = ifx_ad_solve(s, st);
  if (rc != IFX_OK) return rc;
  if (s->opt.compat == IFX_COMPAT_FULL) {
    if ((rc = ifx_ppe_solve(s, st)) != IFX_OK) return rc;
    if ((rc = ifx_correct(s, st)) != IFX_OK) return rc;
  }
  st->ms_total = st->ms_ad + st->ms_ppe + st->ms_correct + st->ms_ib;
  return IFX_OK;
}

// ------------------------------------------------------------------------------------------------
// slabs: one process per GPU; the launcher all-gathers the blobs (torch.distributed) and hands them back
// blob: [0,64) cudaIpcMemHandle_t of the exchange segment, [64,72) field_elems, [72,76) nyl, [76,80) pitch,
//       [80,84) j_begin, [84,88) j_end, [88,92) rank
// ------------------------------------------------------------------------------------------------
extern "C" int ifx_ipc_export(ifx_solver* s, unsigned char* blob) {
  if (!s || !blob) return IFX_ERR_INVALID;
  IFX_CUDA(s, cudaSetDevice(s->device));
  std::memset(blob, 0, IFX_IPC_HANDLE_BYTES);
  cudaIpcMemHandle_t h;
  IFX_CUDA(s, cudaIpcGetMemHandle(&h, s->seg));
  static_assert(sizeof(h) == 64, "cudaIpcMemHandle_t is 64 bytes");
  std::memcpy(blob, &h, 64);
  const unsigned long long fe = s->field_elems;
  std::memcpy(blob + 64, &fe, 8);
  const int meta[5] = {s->L.nyl, s->L.pitch, s->L.jb, s->L.je, s->opt.rank};
  std::memcpy(blob + 72, meta, sizeof(meta));
  return IFX_OK;
}

extern "C" int ifx_ipc_connect(ifx_solver* s, const unsigned char* all, int nranks) {
  if (!s || !all) return IFX_ERR_INVALID;
  if (nranks != s->opt.nranks) return fail(s, IFX_ERR_INVALID, "nranks does not match the options the solver was created with");
  IFX_CUDA(s, cudaSetDevice(s->device));
  int prev_je = 1;
  for (int r = 0; r < nranks; r++) {
    const unsigned char* b = all + (size_t)r * IFX_IPC_HANDLE_BYTES;
    unsigned long long fe; int meta[5];
    std::memcpy(&fe, b + 64, 8);
    std::memcpy(meta, b + 72, sizeof(meta));
    if (meta[4] != r) return fail(s, IFX_ERR_INVALID, "blobs are not in rank order");
    if (meta[1] != s->L.pitch) return fail(s, IFX_ERR_INVALID, "ranks disagree on the row pitch (different nx?)");
    if (meta[2] != prev_je) return fail(s, IFX_ERR_INVALID, "slabs do not tile the rows contiguously");
    prev_je = meta[3];
    s->peer_field_elems[r] = (size_t)fe;
    s->peer_nyl[r] = meta[0];
    if (r == s->opt.rank) { s->peer_seg[r] = s->seg; continue; }
    cudaIpcMemHandle_t h;
    std::memcpy(&h, b, 64);
    void* p = nullptr;
    IFX_CUDA(s, cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    s->peer_seg[r] = p;
  }
  if (prev_je != s->L.ny - 1) return fail(s, IFX_ERR_INVALID, "slabs do not cover all interior rows");
  s->connected = true;
  return IFX_OK;
}
