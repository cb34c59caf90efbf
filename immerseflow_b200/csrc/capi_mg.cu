// capi_mg.cu — host side of the Poisson solvers beyond the reference's point Jacobi that work on more than one grid
// level or on whole lines: geometric multigrid with the point smoother (PPE_Solver 4), zebra line
// relaxation (2) and the line-smoothed V-cycle (5).  Kernels: kernels_mg.cu (+ the red-black instantiation of the
// bulk-copy sweep kernel, kernels_v4.cu, for fine-level smoothing and for the residual / stop decision).
#include "solver.h"

#include <algorithm>

using namespace ifx;

// ------------------------------------------------------------------------------------------------
// Poisson by geometric multigrid (PPE_Solver 4).  Semantics: oracle/ifx_oracle_mg.c + the solver-4
// branch of orc_full_poisson (UNPINNED).  An "iteration" is one V(NU1, NU2) cycle; the stop rule is the reference's
// (sum of the residual against ppe_tol, at most PPE_itermax iterations) and is evaluated, like everywhere else,
// by the first fine-level half-sweep of the NEXT cycle, which is the only launch of a cycle that can end the loop —
// the host looks at the control block once per cycle, right after it.
// ------------------------------------------------------------------------------------------------
static int mg_ensure(ifx_solver* s) {
  if (s->mg_levels == 0) {
    int lx[IFX_MG_MAX_LEVELS], ly[IFX_MG_MAX_LEVELS];
    const int n = mg_plan(s->L.nx - 2, s->L.ny - 2, lx, ly);
    if (n < 2) return fail(s, IFX_ERR_INVALID, "multigrid needs at least 3 cells in both directions");
    // all or nothing: a hierarchy that stops short of the plan would converge differently from the oracle's
    cudaError_t e = cudaSuccess;
    for (int l = 1; l < n && e == cudaSuccess; l++) {
      const size_t bytes = sizeof(double) * (size_t)(lx[l] + 2) * (ly[l] + 2);
      s->mg[l].ncx = lx[l]; s->mg[l].ncy = ly[l];
      double** a[] = {&s->mg[l].GE, &s->mg[l].GN, &s->mg[l].e, &s->mg[l].R, &s->mg[l].inv_x, &s->mg[l].cp_x, &s->mg[l].inv_y,
                      &s->mg[l].cp_y, &s->mg[l].dp};
      const int narr = s->opt.ppe_solver == 5 ? 9 : 4;       // line-elimination storage only for the line smoother
      for (int q = 0; q < narr && e == cudaSuccess; q++) {
        e = cudaMalloc(a[q], bytes);
        if (e == cudaSuccess) e = cudaMemsetAsync(*a[q], 0, bytes, s->stream);
      }
    }
    if (e != cudaSuccess) {
      cudaGetLastError();
      for (int l = 1; l < n; l++) {
        double** a[] = {&s->mg[l].GE, &s->mg[l].GN, &s->mg[l].e, &s->mg[l].R, &s->mg[l].inv_x, &s->mg[l].cp_x, &s->mg[l].inv_y,
                        &s->mg[l].cp_y, &s->mg[l].dp};
        for (double** q : a) { if (*q) cudaFree(*q); *q = nullptr; }
      }
      return fail(s, IFX_ERR_CUDA, std::string("multigrid hierarchy: ") + cudaGetErrorString(e));
    }
    s->mg_levels = n;
    s->mg_valid = false;
  }
  if (!s->mg_valid) {
    for (int l = 1; l < s->mg_levels; l++) {
      const size_t bytes = sizeof(double) * (size_t)(s->mg[l].ncx + 2) * (s->mg[l].ncy + 2);
      IFX_CUDA(s, cudaMemsetAsync(s->mg[l].GE, 0, bytes, s->stream));
      IFX_CUDA(s, cudaMemsetAsync(s->mg[l].GN, 0, bytes, s->stream));
      s->launches++;
      const int lines = s->opt.ppe_solver == 5;
      if (l == 1) IFX_CUDA(s, launch_mg_build1(s->L, s->M, s->celltype, s->mg[1], lines, s->stream));
      else IFX_CUDA(s, launch_mg_coarsen(s->mg[l - 1], s->mg[l], lines, s->stream));
      if (lines) {                   // eliminate the lines of both directions once for this set of cell types
        s->launches += 2;
        IFX_CUDA(s, launch_mg_line_factor(s->mg[l], 0, s->stream));
        IFX_CUDA(s, launch_mg_line_factor(s->mg[l], 1, s->stream));
      }
    }
    s->mg_valid = true;
  }
  return IFX_OK;
}

// coarse part of a V-cycle: mg[1].R is set; on return mg[1].e is the correction for level 0 (orc_mg_coarse_cycle)
static int mg_coarse_cycle(ifx_solver* s) {
  const int Lv = s->mg_levels;
  const double omega = s->opt.ppe_omega;
  auto smooth = [&](int l, int its) -> int {
    for (int k = 0; k < its; k++)
      for (int colour = 0; colour < 2; colour++) {
        s->launches++;
        IFX_CUDA(s, launch_mg_smooth(s->mg[l], colour, omega, s->stream));
      }
    return IFX_OK;
  };
  int rc;
  for (int l = 1; l < Lv; l++) {
    IFX_CUDA(s, cudaMemsetAsync(s->mg[l].e, 0, sizeof(double) * (size_t)(s->mg[l].ncx + 2) * (s->mg[l].ncy + 2), s->stream));
    if ((rc = smooth(l, l == Lv - 1 ? ifx_mg_ncoarse(s->mg[l].ncx, s->mg[l].ncy) : IFX_MG_NU1)) != IFX_OK) return rc;
    if (l < Lv - 1) {
      s->launches++;
      IFX_CUDA(s, launch_mg_restrict(s->mg[l], s->mg[l + 1], s->stream));
    }
  }
  for (int l = Lv - 2; l >= 1; l--) {
    s->launches++;
    IFX_CUDA(s, launch_mg_prolong(s->mg[l + 1], s->mg[l], 0, s->stream));
    if ((rc = smooth(l, IFX_MG_NU2)) != IFX_OK) return rc;
  }
  return IFX_OK;
}

// ifx_options.use_graphs: the ~130 small launches of the coarse part are captured once into a CUDA graph and replayed per
// cycle (the arguments never change: level pointers and omega; only the contents of the arrays do).  Any failure to
// capture or instantiate switches the option off for the handle and launches directly.
static int mg_coarse_cycle_maybe_graphed(ifx_solver* s) {
  if (!s->opt.use_graphs) return mg_coarse_cycle(s);
  if (!s->mg_graph) {
    const long long before = s->launches;
    cudaError_t e = cudaStreamBeginCapture(s->stream, cudaStreamCaptureModeThreadLocal);
    int rc = IFX_ERR_CUDA;
    cudaGraph_t g = nullptr;
    if (e == cudaSuccess) {
      rc = mg_coarse_cycle(s);
      e = cudaStreamEndCapture(s->stream, &g);          // also when a launch failed: the stream must leave capture mode
      if (rc == IFX_OK && e == cudaSuccess && g) e = cudaGraphInstantiate(&s->mg_graph, g, 0);
      if (g) cudaGraphDestroy(g);
    }
    s->mg_graph_launches = s->launches - before;
    s->launches = before;
    if (rc != IFX_OK || e != cudaSuccess || !s->mg_graph) {
      cudaGetLastError();
      s->mg_graph = nullptr;
      s->opt.use_graphs = 0;
      return mg_coarse_cycle(s);
    }
  }
  IFX_CUDA(s, cudaGraphLaunch(s->mg_graph, s->stream));
  s->launches += s->mg_graph_launches;
  return IFX_OK;
}

int ifx::run_ppe_multigrid(ifx_solver* s, ifx_step_stats* st) {
  const Layout& L = s->L;
  if (s->opt.nranks != 1) return fail(s, IFX_ERR_INVALID, "multigrid is single-GPU for now");
  const bool exact = s->opt.reduce_mode == IFX_REDUCE_REFERENCE;
  const int itermax = (1.0 > s->opt.ppe_tol) ? s->in.PPE_itermax : 0;      // the loop starts from res = 1.0 (PPESolver.cu:170-172)
  const int ry = rows_per_cta_for(s, 1);
  const dim3 grid = tile_grid(s, ry, 1);
  const size_t nblocks = (size_t)grid.x * grid.y;
  int rc = ensure_partials(s, nblocks);
  if (rc != IFX_OK) return rc;
  if (exact && (rc = ensure_exact_buffers(s)) != IFX_OK) return rc;

  IFX_CUDA(s, cudaEventRecord(s->ev[2], s->stream));
  if ((rc = mg_ensure(s)) != IFX_OK) return rc;
  const int base = s->cur_p;          // a red-black iteration leaves the iterate in the buffer it started in
  IFX_CUDA(s, cudaMemsetAsync(s->ctl, 0, sizeof(LoopCtl), s->stream));
  int K = 0, fallbacks = 0;
  if (itermax > 0) {
    PpeSweepArgs pa{};
    pa.L = L; pa.M = s->M;
    if ((rc = ensure_facemask(s)) != IFX_OK) return rc;
    pa.rhs = s->rhs; pa.facemask = s->facemask;
    pa.res = s->res_a; pa.partials = s->partials; pa.ctl = s->ctl;
    pa.rows_per_cta = ry;
    pa.rc.itermax = itermax; pa.rc.tol = s->opt.ppe_tol; pa.rc.use_second = 0;
    pa.rc.test_abs = s->opt.ppe_abs_residual ? 1 : 0;
    pa.rc.certify = exact ? 0 : 1;
    pa.rc.band = rounding_band(s, nblocks, ry);
    pa.sor = 1; pa.sor_omega = s->opt.ppe_omega;
    const int fo[1] = {4 + base};
    make_halo_ctx(s, 1, 1, fo, &pa.hx);           // single GPU: an empty context
    // colour-0 half-sweep base -> partner; eval > 0: it also evaluates the residual of iterate `eval` (the input)
    auto red = [&](int eval, int decide, int force, bool write_res) -> int {
      pa.pC = s->p[base]; pa.pT = s->p[base ^ 1];
      pa.rc.eval_iter = eval; pa.rc.decide = decide;
      pa.sor_colour = 0; pa.force = force;
      return enqueue_ppe_sweep(s, pa, grid, false, write_res);
    };
    auto black = [&]() -> int {                   // colour-1 half-sweep partner -> base, no residual bookkeeping
      pa.pC = s->p[base ^ 1]; pa.pT = s->p[base];
      pa.rc.eval_iter = 0; pa.rc.decide = 0;
      pa.sor_colour = 1; pa.force = 0;
      return enqueue_ppe_sweep(s, pa, grid, false, false);
    };
    for (int c = 1;; c++) {
      // first smoothing iteration of cycle c; its colour-0 half evaluates the residual of iterate c-1 and decides
      if ((rc = red(c - 1, exact ? 0 : 1, 0, exact)) != IFX_OK) return rc;
      if (exact && c > 1 && (rc = exact_decide(s, pa.rc, false)) != IFX_OK) return rc;
      if ((rc = black()) != IFX_OK) return rc;
      if ((rc = fetch_ctl(s)) != IFX_OK) return rc;
      if (s->h_ctl->done && s->h_ctl->ambiguous) {
        // the fused sum is within rounding of the tolerance: re-evaluate in the reference's summation order
        fallbacks++;
        if ((rc = ensure_exact_buffers(s)) != IFX_OK) return rc;
        pa.res = s->res_a;
        if ((rc = red(c - 1, 0, 1, true)) != IFX_OK) return rc;
        if ((rc = exact_decide(s, pa.rc, false)) != IFX_OK) return rc;
        if ((rc = fetch_ctl(s)) != IFX_OK) return rc;
        if (!s->h_ctl->done && (rc = black()) != IFX_OK) return rc;     // not converged after all: second half
      }
      if (s->h_ctl->done) break;
      if (c >= itermax + 1) return fail(s, IFX_ERR_STATE, "multigrid loop ran past PPE_itermax without a decision");
      for (int k = 1; k < IFX_MG_NU1; k++) {
        if ((rc = red(0, 0, 0, false)) != IFX_OK) return rc;
        if ((rc = black()) != IFX_OK) return rc;
      }
      s->launches++;
      IFX_CUDA(s, launch_mg_restrict_fine(L, s->M, s->celltype, s->rhs, s->p[base], s->mg[1], s->stream));
      if ((rc = mg_coarse_cycle_maybe_graphed(s)) != IFX_OK) return rc;
      s->launches++;
      IFX_CUDA(s, launch_mg_prolong_fine(L, s->celltype, s->mg[1], s->p[base], 0, s->stream));
      for (int k = 0; k < IFX_MG_NU2; k++) {
        if ((rc = red(0, 0, 0, false)) != IFX_OK) return rc;
        if ((rc = black()) != IFX_OK) return rc;
      }
    }
    K = s->h_ctl->iter;
  }
  s->cur_p = base;
  IFX_CUDA(s, cudaEventRecord(s->ev[3], s->stream));
  IFX_CUDA(s, cudaEventSynchronize(s->ev[3]));
  if (st) {
    st->ppe_sweeps = K;
    st->ppe_residual = (itermax > 0) ? (s->opt.ppe_abs_residual ? s->h_ctl->res1 : s->h_ctl->res0) : 1.0;
    st->exact_fallbacks += fallbacks;
    cudaEventElapsedTime(&st->ms_ppe, s->ev[2], s->ev[3]);
  }
  return itermax > 0 ? check_residual_finite(s, "Poisson solve") : IFX_OK;
}

// ------------------------------------------------------------------------------------------------
// Poisson by zebra line relaxation (PPE_Solver 2: the input file's own "2. Line SOR", main.cu:42) and by the
// line-smoothed V-cycle (PPE_Solver 5).  Semantics: orc_ppe_line_iteration / orc_mg_vcycle_lines (UNPINNED).
// An iteration = four passes (x-lines even / odd, y-lines even / odd), each solving its lines exactly (batched Thomas,
// one thread per line) — or one V(NU1, NU2) cycle with that iteration as the smoother on every level.  The residual
// of iterate m-1 is evaluated, and the stop decision taken, by one launch of the bulk-copy sweep kernel in its
// red-black instantiation with omega = 0 (it then changes nothing: its output buffer is scratch), so the certified
// stop rule of the other solvers carries over unchanged.
// ------------------------------------------------------------------------------------------------
int ifx::run_ppe_lines(ifx_solver* s, ifx_step_stats* st) {
  const Layout& L = s->L;
  const bool mg = s->opt.ppe_solver == 5;
  if (s->opt.nranks != 1) return fail(s, IFX_ERR_INVALID, "line relaxation is single-GPU for now");
  const bool exact = s->opt.reduce_mode == IFX_REDUCE_REFERENCE;
  const int itermax = (1.0 > s->opt.ppe_tol) ? s->in.PPE_itermax : 0;      // the loop starts from res = 1.0 (PPESolver.cu:170-172)
  const int ry = rows_per_cta_for(s, 1);
  const dim3 grid = tile_grid(s, ry, 1);
  const size_t nblocks = (size_t)grid.x * grid.y;
  int rc = ensure_partials(s, nblocks);
  if (rc != IFX_OK) return rc;
  if (exact && (rc = ensure_exact_buffers(s)) != IFX_OK) return rc;
  if (!s->line_f[4]) {
    for (double*& q : s->line_f) {
      IFX_CUDA(s, cudaMalloc(&q, sizeof(double) * s->field_elems));
      IFX_CUDA(s, cudaMemsetAsync(q, 0, sizeof(double) * s->field_elems, s->stream));
    }
    s->line_factor_valid = false;
  }

  IFX_CUDA(s, cudaEventRecord(s->ev[2], s->stream));
  if (mg && (rc = mg_ensure(s)) != IFX_OK) return rc;
  if (!s->line_factor_valid) {        // the elimination of every line, once per set of cell types
    s->launches += 2;
    IFX_CUDA(s, launch_line_factor(L, s->M, s->celltype, 0, s->line_f[0], s->line_f[1], s->stream));
    IFX_CUDA(s, launch_line_factor(L, s->M, s->celltype, 1, s->line_f[2], s->line_f[3], s->stream));
    s->line_factor_valid = true;
  }
  const int base = s->cur_p;          // the iterate never leaves this buffer; its partner is scratch
  const double omega = s->opt.ppe_omega;
  IFX_CUDA(s, cudaMemsetAsync(s->ctl, 0, sizeof(LoopCtl), s->stream));
  int K = 0, fallbacks = 0;
  if (itermax > 0) {
    PpeSweepArgs pa{};
    pa.L = L; pa.M = s->M;
    if ((rc = ensure_facemask(s)) != IFX_OK) return rc;
    pa.rhs = s->rhs; pa.facemask = s->facemask;
    pa.res = s->res_a; pa.partials = s->partials; pa.ctl = s->ctl;
    pa.rows_per_cta = ry;
    pa.rc.itermax = itermax; pa.rc.tol = s->opt.ppe_tol; pa.rc.use_second = 0;
    pa.rc.test_abs = s->opt.ppe_abs_residual ? 1 : 0;
    pa.rc.certify = exact ? 0 : 1;
    pa.rc.band = rounding_band(s, nblocks, ry);
    pa.sor = 1; pa.sor_omega = 0.0; pa.sor_colour = 0;
    pa.pC = s->p[base]; pa.pT = s->p[base ^ 1];
    const int fo[1] = {4 + (base ^ 1)};
    make_halo_ctx(s, 1, 1, fo, &pa.hx);
    auto evaluate = [&](int eval, int decide, int force, bool write_res) -> int {   // residual of iterate `eval` (+ decision)
      pa.rc.eval_iter = eval; pa.rc.decide = decide; pa.force = force;
      return enqueue_ppe_sweep(s, pa, grid, false, write_res);
    };
    auto fine_lines = [&](int its) -> int {
      for (int k = 0; k < its; k++)
        for (int dir = 0; dir < 2; dir++)
          for (int parity = 0; parity < 2; parity++) {
            s->launches++;
            IFX_CUDA(s, launch_line_solve(L, s->M, s->celltype, s->rhs, s->p[base], s->line_f[2 * dir], s->line_f[2 * dir + 1],
                                          s->line_f[4], dir, parity, omega, s->stream));
          }
      return IFX_OK;
    };
    auto coarse_lines = [&](int l, int its) -> int {
      for (int k = 0; k < its; k++)
        for (int dir = 0; dir < 2; dir++)
          for (int parity = 0; parity < 2; parity++) {
            s->launches++;
            IFX_CUDA(s, launch_mg_line_solve(s->mg[l], dir, parity, omega, s->stream));
          }
      return IFX_OK;
    };
    for (int c = 1;; c++) {
      if ((rc = evaluate(c - 1, exact ? 0 : 1, 0, exact)) != IFX_OK) return rc;
      if (exact && c > 1 && (rc = exact_decide(s, pa.rc, false)) != IFX_OK) return rc;
      if ((rc = fetch_ctl(s)) != IFX_OK) return rc;
      if (s->h_ctl->done && s->h_ctl->ambiguous) {
        fallbacks++;
        if ((rc = ensure_exact_buffers(s)) != IFX_OK) return rc;
        pa.res = s->res_a;
        if ((rc = evaluate(c - 1, 0, 1, true)) != IFX_OK) return rc;
        if ((rc = exact_decide(s, pa.rc, false)) != IFX_OK) return rc;
        if ((rc = fetch_ctl(s)) != IFX_OK) return rc;
      }
      if (s->h_ctl->done) break;
      if (c >= itermax + 1) return fail(s, IFX_ERR_STATE, "line-relaxation loop ran past PPE_itermax without a decision");
      if (!mg) {
        if ((rc = fine_lines(1)) != IFX_OK) return rc;
        continue;
      }
      const int Lv = s->mg_levels;
      if ((rc = fine_lines(IFX_MG_NU1)) != IFX_OK) return rc;
      s->launches++;
      IFX_CUDA(s, launch_mg_restrict_fine(L, s->M, s->celltype, s->rhs, s->p[base], s->mg[1], s->stream));
      for (int l = 1; l < Lv; l++) {
        IFX_CUDA(s, cudaMemsetAsync(s->mg[l].e, 0, sizeof(double) * (size_t)(s->mg[l].ncx + 2) * (s->mg[l].ncy + 2), s->stream));
        if ((rc = coarse_lines(l, l == Lv - 1 ? ifx_mg_ncoarse_lines(s->mg[l].ncx, s->mg[l].ncy) : IFX_MG_NU1)) != IFX_OK) return rc;
        if (l < Lv - 1) {
          s->launches++;
          IFX_CUDA(s, launch_mg_restrict(s->mg[l], s->mg[l + 1], s->stream));
        }
      }
      for (int l = Lv - 2; l >= 1; l--) {
        s->launches++;
        IFX_CUDA(s, launch_mg_prolong(s->mg[l + 1], s->mg[l], 1, s->stream));
        if ((rc = coarse_lines(l, IFX_MG_NU2)) != IFX_OK) return rc;
      }
      s->launches++;
      IFX_CUDA(s, launch_mg_prolong_fine(L, s->celltype, s->mg[1], s->p[base], 1, s->stream));
      if ((rc = fine_lines(IFX_MG_NU2)) != IFX_OK) return rc;
    }
    K = s->h_ctl->iter;
  }
  s->cur_p = base;
  IFX_CUDA(s, cudaEventRecord(s->ev[3], s->stream));
  IFX_CUDA(s, cudaEventSynchronize(s->ev[3]));
  if (st) {
    st->ppe_sweeps = K;
    st->ppe_residual = (itermax > 0) ? (s->opt.ppe_abs_residual ? s->h_ctl->res1 : s->h_ctl->res0) : 1.0;
    st->exact_fallbacks += fallbacks;
    cudaEventElapsedTime(&st->ms_ppe, s->ev[2], s->ev[3]);
  }
  return itermax > 0 ? check_residual_finite(s, "Poisson solve") : IFX_OK;
}
