// capi.cu — C-ABI entry points (include/immerseflow_c.h) and the host-side solver loops.
// Host arithmetic in this file is compiled with -ffp-contract=off and uses std::fma explicitly so
// the 1-D coefficient tables carry the same bits the reference's per-cell kernels produce.
#include "solver.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>

using namespace ifx;

static thread_local std::string g_create_error;

namespace ifx {
int fail(ifx_solver* s, int code, const std::string& msg) {
  if (s) s->err = msg; else g_create_error = msg;
  return code;
}
}  // namespace ifx

extern "C" const char* ifx_last_error(const ifx_solver* s) {
  return s ? s->err.c_str() : g_create_error.c_str();
}

extern "C" int ifx_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

extern "C" void ifx_default_options(ifx_options* o) {
  if (!o) return;
  std::memset(o, 0, sizeof(*o));
  o->abi_version = IFX_ABI_VERSION;
  o->device = 0;
  o->compat = IFX_COMPAT_REFERENCE;
  o->reduce_mode = IFX_REDUCE_FUSED;
  // the reference's hard-coded boundary values (ADSolver.cu:200-216)
  o->bc.u_bc_w = o->bc.u_bc_e = o->bc.u_bc_n = o->bc.u_bc_s = 1.0;
  o->bc.v_bc_w = o->bc.v_bc_e = o->bc.v_bc_n = o->bc.v_bc_s = 0.0;
  o->ad_tol = std::pow(10.0, -6.0);    // ADSolver.cu:315
  o->ppe_tol = std::pow(10.0, -6.0);   // PPESolver.cu:172
  o->ppe_abs_residual = 0;
  o->rank = 0; o->nranks = 1;
  o->j_begin = 0; o->j_end = 0;        // 0,0 = whole grid
  o->sweeps_per_batch = 64;
  o->use_graphs = 1;                   // replay the coarse part of the multigrid cycle from a CUDA graph (measured: r2_mg_bench*.jsonl)
}

// ------------------------------------------------------------------------------------------------
// helpers
// ------------------------------------------------------------------------------------------------
static int upload_table(ifx_solver* s, const std::vector<double>& h, const double** dst) {
  double* d = nullptr;
  IFX_CUDA(s, cudaMalloc(&d, sizeof(double) * h.size()));
  s->tables.push_back(d);
  IFX_CUDA(s, cudaMemcpy(d, h.data(), sizeof(double) * h.size(), cudaMemcpyHostToDevice));
  *dst = d;
  return IFX_OK;
}

// Grid metrics and separable stencil coefficients (a2, a3, a11 of SURVEY §8): the reference's
// per-cell expressions (preSim.cu:294-355, ADSolver.cu:27-39, PPESolver.cu:88-99) evaluated once
// per column / row instead of once per cell per step.
static int build_metrics(ifx_solver* s, const double* xf, const double* yf) {
  const int nx = s->in.nx, ny = s->in.ny;
  const double dt = s->in.dt, Re = s->in.Re;
  auto centres = [](int n, const double* f, std::vector<double>& c, std::vector<double>& d) {
    const int nf = n - 1;
    c.assign(n, 0.0); d.assign(n, 0.0);
    for (int i = 1; i < n - 1; i++) c[i] = (f[i - 1] + f[i]) / 2.0;           // preSim.cu:294-301
    c[0] = -1 * c[1];                                                        // :304-305
    c[n - 1] = f[nf - 1] + (f[nf - 1] - c[n - 2]);                           // :306-307
    for (int i = 1; i < n - 1; i++) d[i] = f[i] - f[i - 1];                  // :331-332
    d[0] = d[1]; d[n - 1] = d[n - 2];                                        // :337-355
  };
  centres(nx, xf, s->h_xc, s->h_dx);
  centres(ny, yf, s->h_yc, s->h_dy);

  const double k = dt / Re;                                                  // ADSolver.cu:34 (div.rn)
  auto coeffs = [&](int n, const std::vector<double>& d, std::vector<double>& rcp, std::vector<double>& kh,
                    std::vector<double>& adP, std::vector<double>& adM, std::vector<double>& adS,
                    std::vector<double>& ppP, std::vector<double>& ppM, std::vector<double>& ppS, bool x_dir) {
    rcp.assign(n, 0.0); kh.assign(n, 0.0);
    adP.assign(n, 1.0); adM.assign(n, 1.0); adS.assign(n, 1.0);
    ppP.assign(n, 1.0); ppM.assign(n, 1.0); ppS.assign(n, 1.0);
    for (int i = 0; i < n - 1; i++) rcp[i] = 1.0 / (d[i] + d[i + 1]);         // ADSolver.cu:66 (rcp.rn)
    for (int i = 0; i < n; i++) kh[i] = (dt / d[i]) * 0.5;                    // ADSolver.cu:66
    for (int i = 1; i < n - 1; i++) {
      const double a_p = 2.0 / (d[i] * (d[i] + d[i + 1]));
      const double a_m = 2.0 / (d[i] * (d[i] + d[i - 1]));
      adP[i] = k * a_p;                                                      // ADSolver.cu:37,39
      adM[i] = k * a_m;                                                      // ADSolver.cu:36,38
      // x: first half of cP, fma(k, ax_p+ax_m, 1.0); y: the bare sum, folded in per cell
      adS[i] = x_dir ? std::fma(k, a_p + a_m, 1.0) : (a_p + a_m);
      ppP[i] = a_p; ppM[i] = a_m; ppS[i] = a_p + a_m;                        // PPESolver.cu:93-99
    }
  };
  std::vector<double> rcpx, kxh, adE, adW, adPx, ppE, ppW, ppSx;
  std::vector<double> rcpy, kyh, adN, adS, adSy, ppN, ppS, ppSy;
  coeffs(nx, s->h_dx, rcpx, kxh, adE, adW, adPx, ppE, ppW, ppSx, true);
  coeffs(ny, s->h_dy, rcpy, kyh, adN, adS, adSy, ppN, ppS, ppSy, false);

  Metrics& M = s->M;
  int rc;
#define UP(vec, field) if ((rc = upload_table(s, vec, &M.field)) != IFX_OK) return rc;
  UP(s->h_dx, dx) UP(s->h_dy, dy) UP(s->h_xc, xc) UP(s->h_yc, yc)
  UP(rcpx, rcpx) UP(rcpy, rcpy) UP(kxh, kxh) UP(kyh, kyh)
  UP(adE, ad_cE) UP(adW, ad_cW) UP(adPx, ad_px) UP(adN, ad_cN) UP(adS, ad_cS) UP(adSy, ad_sy)
  UP(ppE, pp_cE) UP(ppW, pp_cW) UP(ppSx, pp_sx) UP(ppN, pp_cN) UP(ppS, pp_cS) UP(ppSy, pp_sy)
#undef UP
  M.k = k;
  M.dt = dt;
  return IFX_OK;
}

// The general Poisson sweep runs its wide geometry (four columns per thread, kernels_v4.cu: IFX_PPE_NC2_WIDE) on a single
// GPU and the narrow one on slabs.  The ONE place that decides: the grid (tile_grid) and the launch (enqueue_ppe_sweep)
// both ask here.
bool ifx::ppe_wide_tiles(const ifx_solver* s) { return s->opt.nranks == 1; }

// columns per CTA tile of the sweep kernels (mode 0: Laplace sweep, 1: general Poisson sweep, 2: predictor)
static int tile_cols_for(const ifx_solver* s, int mode) {
  return v4_tile_cols(mode, mode == 1 && ppe_wide_tiles(s));
}

int ifx::rows_per_cta_for(const ifx_solver* s, int mode) {
  if (s->rows_override > 0) return s->rows_override;
  // enough CTAs for >= ~6 per SM when the grid allows it, tall tiles (less halo re-read) otherwise
  const int nxi = s->L.nx - 2, nyi = s->L.je - s->L.jb, tw = tile_cols_for(s, mode);
  const int gx = (nxi + tw - 1) / tw;
  int ry = 64;
  // a 2048-row slab of a 16384-wide grid (8 ranks) is only ~4.6 waves of 64-row tiles: 32-row tiles halve the tail
  // (measured on an emulated 1/8 slab: predictor 8.69 -> 8.32 ms per 25 iterations, Poisson unchanged)
  if ((long long)gx * ((nyi + ry - 1) / ry) < 148LL * 24) ry = 32;
  while (ry > 4 && (long long)gx * ((nyi + ry - 1) / ry) < 148LL * 6) ry >>= 1;
  return ry;
}

dim3 ifx::tile_grid(const ifx_solver* s, int ry, int mode) {
  const int nxi = s->L.nx - 2, nyi = s->L.je - s->L.jb, tw = tile_cols_for(s, mode);
  return dim3((nxi + tw - 1) / tw, (nyi + ry - 1) / ry, 1);
}

int ifx::ensure_partials(ifx_solver* s, size_t nblocks) {
  if (nblocks <= s->partials_cap) return IFX_OK;
  if (s->partials) cudaFree(s->partials);
  s->partials = nullptr;
  IFX_CUDA(s, cudaMalloc(&s->partials, sizeof(double) * 2 * nblocks));
  s->partials_cap = nblocks;
  return IFX_OK;
}

int ifx::ensure_exact_buffers(ifx_solver* s) {
  if (s->res_a) return IFX_OK;
  if (s->opt.nranks != 1)
    return fail(s, IFX_ERR_INVALID, "reference-order reduction is single-GPU only");
  const size_t N = (size_t)s->L.nx * s->L.ny;
  IFX_CUDA(s, cudaMalloc(&s->res_a, sizeof(double) * N));
  IFX_CUDA(s, cudaMalloc(&s->res_b, sizeof(double) * N));
  IFX_CUDA(s, cudaMemsetAsync(s->res_a, 0, sizeof(double) * N, s->stream));
  IFX_CUDA(s, cudaMemsetAsync(s->res_b, 0, sizeof(double) * N, s->stream));
  IFX_CUDA(s, cudaMalloc(&s->red_partial, sizeof(double) * ((N + 255) / 256)));
  return IFX_OK;
}

// ImmerseFlow::Reduction (preSim.cu:376-445) on a device array of n doubles -> *out (device)
static int reduce_reference_order(ifx_solver* s, const double* d_in, size_t n, double* d_out, bool abs_values = false) {
  s->launches += 2;
  IFX_CUDA(s, launch_reduce6(d_in, n, s->red_partial, d_out, s->stream, abs_values));
  return IFX_OK;
}

// a few bytes device -> page-locked host, then wait for the stream.  zero_copy_control: a kernel stores them straight
// into the host buffer (no copy engine, so a bulk download of another handle cannot hold them up)
int ifx::fetch_small(ifx_solver* s, void* host_pinned, const void* dev, size_t bytes) {
  if (s->opt.zero_copy_control) {
    s->launches++;
    IFX_CUDA(s, launch_copy_words(host_pinned, dev, bytes, s->stream));
  } else {
    IFX_CUDA(s, cudaMemcpyAsync(host_pinned, dev, bytes, cudaMemcpyDeviceToHost, s->stream));
  }
  IFX_CUDA(s, cudaStreamSynchronize(s->stream));
  return IFX_OK;
}

int ifx::check_residual_finite(ifx_solver* s, const char* stage) {
  if (std::isfinite(s->h_ctl->res0) && std::isfinite(s->h_ctl->res1)) return IFX_OK;
  return fail(s, IFX_ERR_STATE, std::string(stage) + ": the residual is not finite after iteration " + std::to_string(s->h_ctl->iter) +
                                " — the solution has diverged (time step too large for the explicit convection, or NaN in the input state)");
}

int ifx::fetch_ctl(ifx_solver* s) {
  static_assert(sizeof(LoopCtl) % 4 == 0, "LoopCtl is copied in 4-byte words");
  return fetch_small(s, s->h_ctl, s->ctl, sizeof(LoopCtl));
}

void ifx::fill_bc(const ifx_solver* s, double* two_u, double* two_v) {
  const ifx_bc& b = s->opt.bc;
  two_u[0] = b.u_bc_w * 2.0; two_u[1] = b.u_bc_e * 2.0; two_u[2] = b.u_bc_s * 2.0; two_u[3] = b.u_bc_n * 2.0;
  two_v[0] = b.v_bc_w * 2.0; two_v[1] = b.v_bc_e * 2.0; two_v[2] = b.v_bc_s * 2.0; two_v[3] = b.v_bc_n * 2.0;
}

// Half-width (relative to the sum of magnitudes) of the band inside which the fused summation order
// and the reference's could disagree about `sum > tol`: gamma_d with d = the deepest chain of
// additions either order applies to one term, times a safety factor of 2.
double ifx::rounding_band(const ifx_solver* s, size_t nblocks, int rows_per_cta) {
  const double N = (double)s->L.nx * s->L.ny;
  const double d_ref = 10.0 + std::ceil(N / 256.0 / 512.0) + 1.0 + 8.0 + 2.0;
  const double d_fused = 8.0 * rows_per_cta + 5.0 + AD_WARPS + std::ceil((double)nblocks / AD_THREADS) + 5.0 + AD_WARPS + 2.0;
  return 2.0 * (d_ref + d_fused) * 1.1102230246251565e-16;
}

// ------------------------------------------------------------------------------------------------
// slabs: halo context of a launch, generic halo exchange, waits
// ------------------------------------------------------------------------------------------------
static XchgSync* peer_sync(ifx_solver* s, int r) {
  const size_t off = ((size_t)s->nseg_fields * s->peer_field_elems[r] * sizeof(double) + 255) / 256 * 256;
  return reinterpret_cast<XchgSync*>(reinterpret_cast<char*>(s->peer_seg[r]) + off);
}
static double* peer_field(ifx_solver* s, int r, int field_index) {
  return reinterpret_cast<double*>(s->peer_seg[r]) + (size_t)field_index * s->peer_field_elems[r];
}

// field_index: 0,1 = u[0],u[1]; 2,3 = v[0],v[1]; 4,5,(6) = p[0],p[1],(p[2])
void ifx::make_halo_ctx(ifx_solver* s, int group, int nfields, const int* out_field_index, HaloCtx* hx) {
  std::memset(hx, 0, sizeof(*hx));
  hx->nranks = s->connected ? s->opt.nranks : 1;
  hx->rank = s->opt.rank;
  if (hx->nranks == 1) return;
  const int r = s->opt.rank;
  hx->has_lo = r > 0;
  hx->has_hi = r < s->opt.nranks - 1;
  hx->seq = ++s->seq[group];
  hx->mseq = ++s->mseq;
  hx->wait_lo = s->sync->flags[group][0];
  hx->wait_hi = s->sync->flags[group][1];
  hx->gcw_lo = &s->sync->gcflag[0];
  hx->gcw_hi = &s->sync->gcflag[1];
  if (hx->has_lo) {
    hx->signal_lo = peer_sync(s, r - 1)->flags[group][1];
    for (int f = 0; f < nfields; f++)
      hx->peer_row_lo[f] = peer_field(s, r - 1, out_field_index[f]) + (size_t)(s->peer_nyl[r - 1] - 1) * s->L.pitch;
  }
  if (hx->has_hi) {
    hx->signal_hi = peer_sync(s, r + 1)->flags[group][0];
    for (int f = 0; f < nfields; f++) hx->peer_row_hi[f] = peer_field(s, r + 1, out_field_index[f]);
  }
  for (int q = 0; q < s->opt.nranks; q++) {
    XchgSync* ps = peer_sync(s, q);
    hx->mail[q] = &ps->mail[0][0][0];
  }
}

int ifx::halo_wait(ifx_solver* s, int group, unsigned need, int tile_cols) {
  if (!s->connected || s->opt.nranks == 1) return IFX_OK;
  const int ntiles = (s->L.nx - 2 + tile_cols - 1) / tile_cols;
  s->launches++;
  IFX_CUDA(s, launch_halo_wait(s->opt.rank > 0 ? s->sync->flags[group][0] : nullptr,
                               s->opt.rank < s->opt.nranks - 1 ? s->sync->flags[group][1] : nullptr, ntiles, need, s->stream));
  return IFX_OK;
}

// deliver my first / last owned row of the listed fields to the neighbours and publish `seq` of the group.
// ctl != null: in-loop use, skipped on the device when the loop finished before iteration `iter`.
int ifx::halo_push(ifx_solver* s, int group, int nfields, const int* field_index, int tile_cols, unsigned seq,
                   const LoopCtl* ctl, int iter, bool gc_flags) {
  if (!s->connected || s->opt.nranks == 1) return IFX_OK;
  const int r = s->opt.rank;
  HaloPushArgs a{};
  a.L = s->L;
  a.nfields = nfields;
  a.has_lo = r > 0; a.has_hi = r < s->opt.nranks - 1;
  a.seq = seq;
  a.tile_cols = tile_cols;
  a.ntiles = (s->L.nx - 2 + tile_cols - 1) / tile_cols;
  a.ctl = ctl; a.iter = iter;
  if (a.ntiles > IFX_MAX_TILES) return fail(s, IFX_ERR_INVALID, "too many column tiles for the flag array");
  for (int f = 0; f < nfields; f++) {
    a.src[f] = peer_field(s, r, field_index[f]);
    if (a.has_lo) a.dst_lo[f] = peer_field(s, r - 1, field_index[f]) + (size_t)(s->peer_nyl[r - 1] - 1) * s->L.pitch;
    if (a.has_hi) a.dst_hi[f] = peer_field(s, r + 1, field_index[f]);
  }
  if (a.has_lo) a.signal_lo = peer_sync(s, r - 1)->flags[group][1];
  if (a.has_hi) a.signal_hi = peer_sync(s, r + 1)->flags[group][0];
  if (gc_flags) {
    if (a.has_lo) a.gc_signal_lo = &peer_sync(s, r - 1)->gcflag[1];
    if (a.has_hi) a.gc_signal_hi = &peer_sync(s, r + 1)->gcflag[0];
  }
  s->launches++;
  IFX_CUDA(s, launch_halo_push(a, s->stream));
  return IFX_OK;
}

// push + wait for the neighbours' rows.  With nfields == 0 this is a pairwise barrier: everything a neighbour
// enqueued before its call is complete when the wait kernel lets this stream continue.
int ifx::halo_exchange(ifx_solver* s, int group, int nfields, const int* field_index, int tile_cols, bool gc_flags) {
  if (!s->connected || s->opt.nranks == 1) return IFX_OK;
  const unsigned seq = ++s->seq[group];
  int rc = halo_push(s, group, nfields, field_index, tile_cols, seq, nullptr, 0, gc_flags);
  if (rc != IFX_OK) return rc;
  return halo_wait(s, group, seq, tile_cols);
}

// immersed bodies on a slab run: ghost-cell stencils may read the neighbours' memory, and the sweeps defer
// their flags to a push kernel behind the ghost-cell kernel.  set_bodies is collective, so every rank agrees.
bool ifx::bodies_on_slabs(const ifx_solver* s) { return s->connected && s->opt.nranks > 1 && s->nbodies > 0; }

// field f0 (and f1) of the neighbours, as the ghost-cell kernels' remote source
GcPeers ifx::gc_peers(ifx_solver* s, int f0, int f1) {
  GcPeers pr{};
  if (!s->connected || s->opt.nranks == 1) return pr;
  const int r = s->opt.rank;
  const int f[2] = {f0, f1};
  for (int k = 0; k < 2; k++) {
    if (f[k] < 0) continue;
    if (r > 0) pr.lo[k] = peer_field(s, r - 1, f[k]);
    if (r < s->opt.nranks - 1) pr.hi[k] = peer_field(s, r + 1, f[k]);
  }
  return pr;
}

// ------------------------------------------------------------------------------------------------
// lifetime
// ------------------------------------------------------------------------------------------------
extern "C" int ifx_create(const ifx_input* in, const double* xf, const double* yf,
                          const ifx_options* opt_in, ifx_solver** out) {
  if (!in || !xf || !yf || !out) return fail(nullptr, IFX_ERR_INVALID, "null argument");
  *out = nullptr;
  ifx_options opt;
  if (opt_in) opt = *opt_in; else ifx_default_options(&opt);
  if (opt.abi_version != IFX_ABI_VERSION) return fail(nullptr, IFX_ERR_INVALID, "ABI version mismatch");
  if (in->nx < 4 || in->ny < 4) return fail(nullptr, IFX_ERR_INVALID, "grid too small");
  if ((long long)in->nx * in->ny >= (1LL << 31)) return fail(nullptr, IFX_ERR_INVALID, "nx*ny must fit 32-bit ids");
  if (opt.j_begin == 0 && opt.j_end == 0) { opt.j_begin = 1; opt.j_end = in->ny - 1; }
  if (opt.j_begin < 1 || opt.j_end > in->ny - 1 || opt.j_begin >= opt.j_end)
    return fail(nullptr, IFX_ERR_INVALID, "bad slab rows");
  if (opt.sweeps_per_batch < 1) opt.sweeps_per_batch = 64;
  // PPE_Solver of the input file (main.cu:42; the reference documents "1. Point GS, 2. Line SOR" and runs point Jacobi):
  // the reference-compatible mode always reproduces its Jacobi; the full mode honours 2 .. 5
  if (opt.ppe_solver == 0)
    opt.ppe_solver = (opt.compat == IFX_COMPAT_FULL && in->PPE_solver >= 2 && in->PPE_solver <= 5) ? in->PPE_solver : 1;
  if (opt.ppe_omega == 0.0) opt.ppe_omega = (in->w_PPE != 0) ? (double)in->w_PPE : 1.0;
  if (opt.ppe_solver < 1 || opt.ppe_solver > 5)
    return fail(nullptr, IFX_ERR_INVALID, "ppe_solver: 1 point Jacobi, 2 zebra line SOR, 3 red-black SOR, 4 multigrid (point smoother), "
                                          "5 multigrid (line smoother)");
  if (opt.ppe_solver != 1 && opt.compat != IFX_COMPAT_FULL)
    return fail(nullptr, IFX_ERR_INVALID, "SOR / line SOR / multigrid need IFX_COMPAT_FULL (the reference mode reproduces the reference's Jacobi)");
  if (opt.ppe_solver != 1 && !(opt.ppe_omega > 0.0 && opt.ppe_omega < 2.0))
    return fail(nullptr, IFX_ERR_INVALID, "SOR needs 0 < ppe_omega < 2");
  if (opt.ppe_solver == 4 || opt.ppe_solver == 5) {
    int lx[IFX_MG_MAX_LEVELS], ly[IFX_MG_MAX_LEVELS];
    if (mg_plan(in->nx - 2, in->ny - 2, lx, ly) < 2)
      return fail(nullptr, IFX_ERR_INVALID, "multigrid needs at least 3 cells in both directions");
  }
  if ((opt.ppe_solver == 2 || opt.ppe_solver >= 4) && opt.nranks > 1)
    return fail(nullptr, IFX_ERR_INVALID, "line SOR and multigrid are single-GPU for now (slab runs: ppe_solver 1 or 3)");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    return fail(nullptr, IFX_ERR_CUDA, "no CUDA device: the B200 kernels are the only implementation (no CPU fallback)");
  }
  if (opt.device < 0 || opt.device >= ndev) return fail(nullptr, IFX_ERR_INVALID, "bad device ordinal");

  ifx_solver* s = new (std::nothrow) ifx_solver();
  if (!s) return fail(nullptr, IFX_ERR_NOMEM, "host allocation failed");
  s->in = *in;
  s->opt = opt;
  s->device = opt.device;
  if (const char* e = std::getenv("IFX_ROWS_PER_CTA")) s->rows_override = std::atoi(e);
  auto bail = [&](int code) { g_create_error = s->err; ifx_destroy(s); return code; };

  cudaError_t e = cudaSetDevice(s->device);
  if (e != cudaSuccess) { s->err = cudaGetErrorString(e); return bail(IFX_ERR_CUDA); }
  e = cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking);
  if (e != cudaSuccess) { s->err = cudaGetErrorString(e); return bail(IFX_ERR_CUDA); }
  s->own_stream = true;

  Layout& L = s->L;
  L.nx = in->nx; L.ny = in->ny;
  L.pitch = ((IFX_PADL + in->nx + 1) + 15) / 16 * 16;      // +1: a double2 may straddle the last column
  L.jb = opt.j_begin; L.je = opt.j_end;
  L.j0 = opt.j_begin - 1;
  L.nyl = (opt.j_end - opt.j_begin) + 2;
  s->field_elems = (size_t)L.pitch * L.nyl + 64;

  int rc = build_metrics(s, xf, yf);
  if (rc != IFX_OK) return bail(rc);

  auto alloc_field = [&](double** p) -> int {
    IFX_CUDA(s, cudaMalloc(p, sizeof(double) * s->field_elems));
    IFX_CUDA(s, cudaMemsetAsync(*p, 0, sizeof(double) * s->field_elems, s->stream));
    return IFX_OK;
  };
  if (opt.nranks < 1 || opt.nranks > IFX_MAX_RANKS || opt.rank < 0 || opt.rank >= opt.nranks) {
    s->err = "bad rank / nranks (1..8 slabs)";
    return bail(IFX_ERR_INVALID);
  }
  // exchange segment: the ping-pong fields + the synchronisation area, one allocation (see common.cuh).  A slab run
  // keeps a third pressure buffer: its stop decision lags one sweep behind (no cross-GPU wait per sweep), so the
  // converged iterate must survive two more sweeps (run_ppe_loop).
  s->np = opt.nranks > 1 ? 3 : 2;
  s->nseg_fields = 4 + s->np;
  s->sync_off = ((size_t)s->nseg_fields * s->field_elems * sizeof(double) + 255) / 256 * 256;
  s->seg_bytes = s->sync_off + sizeof(XchgSync);
  if (cudaMalloc(&s->seg, s->seg_bytes) != cudaSuccess) { s->err = "cudaMalloc exchange segment"; return bail(IFX_ERR_CUDA); }
  cudaMemsetAsync(s->seg, 0, s->seg_bytes, s->stream);
  s->u[0] = s->seg; s->u[1] = s->seg + s->field_elems;
  s->v[0] = s->seg + 2 * s->field_elems; s->v[1] = s->seg + 3 * s->field_elems;
  for (int b = 0; b < s->np; b++) s->p[b] = s->seg + (size_t)(4 + b) * s->field_elems;
  s->sync = reinterpret_cast<XchgSync*>(reinterpret_cast<char*>(s->seg) + s->sync_off);
  s->peer_seg[opt.rank] = s->seg;
  s->peer_field_elems[opt.rank] = s->field_elems;
  s->peer_nyl[opt.rank] = L.nyl;
  double** fields[] = {&s->sx, &s->sy};
  for (double** f : fields)
    if ((rc = alloc_field(f)) != IFX_OK) return bail(rc);
  if (opt.compat == IFX_COMPAT_FULL || in->nx > in->ny) {
    if ((rc = alloc_field(&s->uf)) != IFX_OK) return bail(rc);
    if ((rc = alloc_field(&s->vf)) != IFX_OK) return bail(rc);
  }
  if (opt.compat == IFX_COMPAT_FULL) {
    if ((rc = alloc_field(&s->rhs)) != IFX_OK) return bail(rc);
    if (cudaMalloc(&s->d_ub, sizeof(double) * 64) != cudaSuccess || cudaMalloc(&s->d_vb, sizeof(double) * 64) != cudaSuccess) {
      s->err = "cudaMalloc body velocities";
      return bail(IFX_ERR_CUDA);
    }
    cudaMemsetAsync(s->d_ub, 0, sizeof(double) * 64, s->stream);
    cudaMemsetAsync(s->d_vb, 0, sizeof(double) * 64, s->stream);
  }
  if (cudaMalloc(&s->celltype, s->field_elems) != cudaSuccess) { s->err = "cudaMalloc celltype"; return bail(IFX_ERR_CUDA); }
  launch_fill_u8(s->celltype, s->field_elems, IFX_FLUID, s->stream);
  s->launches++;
  if (opt.compat == IFX_COMPAT_FULL) {
    if (cudaMalloc(&s->facemask, s->field_elems) != cudaSuccess) { s->err = "cudaMalloc facemask"; return bail(IFX_ERR_CUDA); }
    cudaMemsetAsync(s->facemask, 0, s->field_elems, s->stream);
  }
  if (cudaMalloc(&s->ctl, sizeof(LoopCtl)) != cudaSuccess || cudaMallocHost(&s->h_ctl, sizeof(LoopCtl)) != cudaSuccess ||
      cudaMallocHost(&s->h_counters, sizeof(int) * 4) != cudaSuccess ||
      cudaMalloc(&s->red_out, sizeof(double) * 4) != cudaSuccess) {
    s->err = "control block allocation failed";
    return bail(IFX_ERR_CUDA);
  }
  cudaMemsetAsync(s->ctl, 0, sizeof(LoopCtl), s->stream);
  for (auto& ev : s->ev) cudaEventCreate(&ev);
  e = cudaStreamSynchronize(s->stream);
  if (e != cudaSuccess) { s->err = cudaGetErrorString(e); return bail(IFX_ERR_CUDA); }
  *out = s;
  return IFX_OK;
}

extern "C" int ifx_destroy(ifx_solver* s) {
  if (!s) return IFX_OK;
  cudaSetDevice(s->device);
  if (s->stream) cudaStreamSynchronize(s->stream);
  for (int r = 0; r < IFX_MAX_RANKS; r++)
    if (s->peer_seg[r] && r != s->opt.rank) cudaIpcCloseMemHandle(s->peer_seg[r]);
  double* fields[] = {s->seg, s->sx, s->sy, s->rhs, s->uf, s->vf,
                      s->partials, s->res_a, s->res_b, s->red_partial, s->red_out, s->d_xm, s->d_ym, s->d_ub, s->d_vb,
                      s->d_bbox, s->gc.w_dir, s->gc.w_neu, s->gc.bi, s->gc.ip, s->gc_tmp_a, s->gc_tmp_b};
  for (double* f : fields) if (f) cudaFree(f);
  int* ifields[] = {s->d_body_off, s->gc.cell, s->gc.ref_id, s->gc.stencil, s->gc.stencil_ref, s->gc.body,
                    s->d_counters, s->d_rowcount, s->d_rowstart};
  for (int* f : ifields) if (f) cudaFree(f);
  for (double* t : s->tables) cudaFree(t);
  if (s->mg_graph) cudaGraphExecDestroy(s->mg_graph);
  for (double* q : s->line_f) if (q) cudaFree(q);
  for (int l = 1; l < s->mg_levels; l++) {
    double* a[] = {s->mg[l].GE, s->mg[l].GN, s->mg[l].e, s->mg[l].R, s->mg[l].inv_x, s->mg[l].cp_x, s->mg[l].inv_y, s->mg[l].cp_y,
                   s->mg[l].dp};
    for (double* q : a) if (q) cudaFree(q);
  }
  if (s->celltype) cudaFree(s->celltype);
  if (s->facemask) cudaFree(s->facemask);
  if (s->ctl) cudaFree(s->ctl);
  if (s->h_ctl) cudaFreeHost(s->h_ctl);
  if (s->h_counters) cudaFreeHost(s->h_counters);
  if (s->h_stage) cudaFreeHost(s->h_stage);
  for (auto& ev : s->ev) if (ev) cudaEventDestroy(ev);
  if (s->own_stream && s->stream) cudaStreamDestroy(s->stream);
  delete s;
  return IFX_OK;
}

extern "C" int ifx_set_stream(ifx_solver* s, void* stream) {
  if (!s) return IFX_ERR_INVALID;
  cudaSetDevice(s->device);
  if (s->stream) cudaStreamSynchronize(s->stream);
  if (s->own_stream && s->stream) cudaStreamDestroy(s->stream);
  s->stream = (cudaStream_t)stream;
  s->own_stream = false;
  return IFX_OK;
}

extern "C" int ifx_synchronize(ifx_solver* s) {
  if (!s) return IFX_ERR_INVALID;
  IFX_CUDA(s, cudaSetDevice(s->device));
  IFX_CUDA(s, cudaStreamSynchronize(s->stream));
  return IFX_OK;
}

extern "C" long long ifx_launch_count(const ifx_solver* s) { return s ? s->launches : 0; }

// initializeData (preSim.cu:201-217): vortex IC, iBlank.  In IFX_COMPAT_REFERENCE iBlank == 1
// (preSim.cu:133); with bodies set (ifx_set_bodies) the classification runs in ifx_iblank_update.
extern "C" int ifx_initialize(ifx_solver* s) {
  if (!s) return IFX_ERR_INVALID;
  IFX_CUDA(s, cudaSetDevice(s->device));
  const Layout& L = s->L;
  for (int b = 0; b < s->np; b++) {   // all ping-pong partners start from the same field
    s->launches++;
    IFX_CUDA(s, launch_init_vortex(L, s->M.xc, s->M.yc, s->u[b & 1], s->v[b & 1], s->p[b], s->stream));
  }
  s->cur_uv = 0; s->cur_p = 0;
  s->launches++;
  IFX_CUDA(s, launch_fill_u8(s->celltype, s->field_elems, IFX_FLUID, s->stream));
  if (s->uf) {   // initializeKernel zeroes the face arrays (preSim.cu:79-96)
    IFX_CUDA(s, cudaMemsetAsync(s->uf, 0, sizeof(double) * s->field_elems, s->stream));
    IFX_CUDA(s, cudaMemsetAsync(s->vf, 0, sizeof(double) * s->field_elems, s->stream));
  }
  s->faces_valid = false;
  s->state_bc_fresh = false;
  s->bodies_dirty = s->nbodies > 0;      // the cell types were just reset to all-fluid
  s->facemask_valid = false;
  s->has_gc = false;
  s->gc.count = 0;
  s->mg_valid = false;               // multigrid hierarchy and line eliminations were built for the old cell types
  s->line_factor_valid = false;
  s->initialized = true;
  IFX_CUDA(s, cudaStreamSynchronize(s->stream));
  return IFX_OK;
}

// ------------------------------------------------------------------------------------------------
// state access
// ------------------------------------------------------------------------------------------------
struct FieldView { double* dev; int i0, width, jl0, rows; bool is_u8; int raw; const double* table; };

static int field_view(ifx_solver* s, ifx_field f, FieldView* v) {
  const Layout& L = s->L;
  *v = FieldView{nullptr, 0, L.nx, 0, L.nyl, false, 0, nullptr};
  switch (f) {
    case IFX_FIELD_U: v->dev = s->u[s->cur_uv]; break;
    case IFX_FIELD_V: v->dev = s->v[s->cur_uv]; break;
    case IFX_FIELD_P: v->dev = s->p[s->cur_p]; break;
    case IFX_FIELD_SX: v->dev = s->sx; break;
    case IFX_FIELD_SY: v->dev = s->sy; break;
    case IFX_FIELD_PPE_RHS: v->dev = s->rhs; break;
    case IFX_FIELD_IBLANK: v->is_u8 = true; v->raw = 0; break;
    case IFX_FIELD_CELLTYPE: v->is_u8 = true; v->raw = 1; break;
    // UF(i,j) = face east of cell (i,j): i = 0..nx-2 on interior rows  (Data.u.velf, ADSolver.cu:167-177)
    case IFX_FIELD_UF: v->dev = s->uf; v->i0 = 0; v->width = L.nx - 1; v->jl0 = 1; v->rows = L.nyl - 2; break;
    // VF(i,j) = face north of cell (i,j): i = 1..nx-2, j = 0..ny-2     (Data.v.velf, ADSolver.cu:179-186)
    case IFX_FIELD_VF: v->dev = s->vf; v->i0 = 1; v->width = L.nx - 2; v->jl0 = 0; v->rows = L.nyl - 1; break;
    case IFX_FIELD_XC: v->table = s->h_xc.data(); v->width = L.nx; v->rows = 1; break;
    case IFX_FIELD_YC: v->table = s->h_yc.data(); v->width = L.ny; v->rows = 1; break;
    default: return fail(s, IFX_ERR_INVALID, "unknown field");
  }
  if (!v->table && !v->is_u8 && !v->dev) return fail(s, IFX_ERR_STATE, "field not allocated in this mode");
  return IFX_OK;
}

extern "C" size_t ifx_field_size(const ifx_solver* s, ifx_field f) {
  if (!s) return 0;
  FieldView v;
  if (field_view(const_cast<ifx_solver*>(s), f, &v) != IFX_OK) return 0;
  return (size_t)v.width * v.rows;
}

static int transfer_block_rows(int width) {
  const long long rows = (16LL << 20) / ((long long)width * (long long)sizeof(double));
  return (int)std::max(1LL, rows);
}

static int set_field_impl(ifx_solver* s, ifx_field f, const double* host, size_t n, bool sync) {
  if (!s || !host) return IFX_ERR_INVALID;
  IFX_CUDA(s, cudaSetDevice(s->device));
  FieldView v;
  int rc = field_view(s, f, &v);
  if (rc != IFX_OK) return rc;
  if (v.table || v.is_u8) return fail(s, IFX_ERR_INVALID, "field is read-only");
  if (n != (size_t)v.width * v.rows) return fail(s, IFX_ERR_INVALID, "size mismatch");
  const Layout& L = s->L;
  // async transfers go in row blocks of ~16 MB: a copy engine serves its queue in order, and the few-byte control
  // copies of a step running on ANOTHER handle (stop flags, ghost-cell counts, marker uploads) must be able to slip in
  // between the blocks instead of waiting behind a whole field (measured: one whole-field copy per direction stalls
  // the concurrent step by the full transfer time)
  const int blk = sync ? v.rows : transfer_block_rows(v.width);
  for (int r0 = 0; r0 < v.rows; r0 += blk) {
    const int nr = std::min(blk, v.rows - r0);
    IFX_CUDA(s, cudaMemcpy2DAsync(v.dev + lidx(L, v.i0, v.jl0 + r0), sizeof(double) * L.pitch, host + (size_t)r0 * v.width,
                                  sizeof(double) * v.width, sizeof(double) * v.width, nr, cudaMemcpyHostToDevice, s->stream));
  }
  if (f == IFX_FIELD_U || f == IFX_FIELD_V || f == IFX_FIELD_P) {
    // keep the ping-pong partner's ghost ring consistent (a freshly set state has no history)
    for (int k = 1; k < (f == IFX_FIELD_P ? s->np : 2); k++) {
      double* other = (f == IFX_FIELD_U) ? s->u[s->cur_uv ^ 1] : (f == IFX_FIELD_V) ? s->v[s->cur_uv ^ 1] : s->p[(s->cur_p + k) % s->np];
      s->launches++;
      IFX_CUDA(s, launch_copy_ring(L, v.dev, other, nullptr, nullptr, s->stream));
    }
    if (f != IFX_FIELD_P) { s->faces_valid = false; s->state_bc_fresh = false; }
  }
  if (f == IFX_FIELD_UF || f == IFX_FIELD_VF) s->faces_valid = true;
  if (sync) IFX_CUDA(s, cudaStreamSynchronize(s->stream));
  return IFX_OK;
}

extern "C" int ifx_set_field(ifx_solver* s, ifx_field f, const double* host, size_t n) { return set_field_impl(s, f, host, n, true); }
// enqueue only: the copy runs on the handle's stream, ahead of whatever is called next on this handle; `host` (pinned
// memory, or the copy is not asynchronous) must stay untouched until ifx_synchronize or the next blocking call
extern "C" int ifx_set_field_async(ifx_solver* s, ifx_field f, const double* host, size_t n) { return set_field_impl(s, f, host, n, false); }

static int get_field_impl(ifx_solver* s, ifx_field f, double* host, size_t n, bool sync) {
  if (!s || !host) return IFX_ERR_INVALID;
  IFX_CUDA(s, cudaSetDevice(s->device));
  FieldView v;
  int rc = field_view(s, f, &v);
  if (rc != IFX_OK) return rc;
  if (n != (size_t)v.width * v.rows) return fail(s, IFX_ERR_INVALID, "size mismatch");
  const Layout& L = s->L;
  if (v.table) { std::memcpy(host, v.table, sizeof(double) * n); return IFX_OK; }
  if (v.is_u8 && !sync) return fail(s, IFX_ERR_INVALID, "cell types are read synchronously");
  if (v.is_u8) {
    double* tmp = nullptr;
    IFX_CUDA(s, cudaMalloc(&tmp, sizeof(double) * n));
    s->launches++;
    cudaError_t e = launch_pack_u8(L, s->celltype, tmp, v.raw, s->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(host, tmp, sizeof(double) * n, cudaMemcpyDeviceToHost, s->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s->stream);
    cudaFree(tmp);
    IFX_CUDA(s, e);
    return IFX_OK;
  }
  const int blk = sync ? v.rows : transfer_block_rows(v.width);
  for (int r0 = 0; r0 < v.rows; r0 += blk) {
    const int nr = std::min(blk, v.rows - r0);
    IFX_CUDA(s, cudaMemcpy2DAsync(host + (size_t)r0 * v.width, sizeof(double) * v.width, v.dev + lidx(L, v.i0, v.jl0 + r0),
                                  sizeof(double) * L.pitch, sizeof(double) * v.width, nr, cudaMemcpyDeviceToHost, s->stream));
  }
  if (sync) IFX_CUDA(s, cudaStreamSynchronize(s->stream));
  return IFX_OK;
}

extern "C" int ifx_get_field(ifx_solver* s, ifx_field f, double* host, size_t n) { return get_field_impl(s, f, host, n, true); }
// enqueue only: `host` (pinned) holds the field once ifx_synchronize (or any blocking call on this handle) returned
extern "C" int ifx_get_field_async(ifx_solver* s, ifx_field f, double* host, size_t n) { return get_field_impl(s, f, host, n, false); }

// saveDataToFile (postSim.cu:10-39): D2H then the Tecplot writer.
extern "C" int ifx_save_field(ifx_solver* s, ifx_field f, const char* filename) {
  if (!s || !filename) return IFX_ERR_INVALID;
  if (s->opt.nranks != 1) return fail(s, IFX_ERR_INVALID, "ifx_save_field: gather the slabs first");
  const size_t n = ifx_field_size(s, f);
  if (n != (size_t)s->L.nx * s->L.ny) return fail(s, IFX_ERR_INVALID, "only cell-centred fields can be saved");
  std::vector<double> h(n);
  int rc = ifx_get_field(s, f, h.data(), n);
  if (rc != IFX_OK) return rc;
  rc = ifx_write_results_to_file(s->h_xc.data(), s->h_yc.data(), h.data(), s->L.nx, s->L.ny, filename);
  if (rc != IFX_OK) return fail(s, rc, std::string("cannot write ") + filename);
  return IFX_OK;
}

// ------------------------------------------------------------------------------------------------
// Reduction (preSim.cu:376-445) on host data
// ------------------------------------------------------------------------------------------------
extern "C" int ifx_reduce_sum(ifx_solver* s, const double* host, size_t n, double* out) {
  if (!s || !host || !out || n == 0 || n >= (1ull << 31)) return IFX_ERR_INVALID;
  IFX_CUDA(s, cudaSetDevice(s->device));
  double *d_in = nullptr, *d_part = nullptr;
  IFX_CUDA(s, cudaMalloc(&d_in, sizeof(double) * n));
  IFX_CUDA(s, cudaMalloc(&d_part, sizeof(double) * ((n + 255) / 256)));
  cudaError_t e = cudaMemcpyAsync(d_in, host, sizeof(double) * n, cudaMemcpyHostToDevice, s->stream);
  if (e == cudaSuccess) e = launch_reduce6(d_in, n, d_part, s->red_out, s->stream);
  s->launches += 2;
  if (e == cudaSuccess) e = cudaMemcpyAsync(out, s->red_out, sizeof(double), cudaMemcpyDeviceToHost, s->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(s->stream);
  cudaFree(d_in); cudaFree(d_part);
  IFX_CUDA(s, e);
  return IFX_OK;
}

// ------------------------------------------------------------------------------------------------
// predictor — ImmerseFlow::ADsolver(), ADSolver.cu:268-395
// ------------------------------------------------------------------------------------------------
static int launch_ad_jacobi(ifx_solver* s, AdJacobiArgs& a, dim3 grid, bool write_res) {
  s->launches++;
  IFX_CUDA(s, ifx::launch_ad_jacobi_v4(a, grid, s->stream, write_res));
  return IFX_OK;
}

// reference-order evaluation of the two residual arrays + stop decision (ADSolver.cu:360-366, :315)
int ifx::exact_decide(ifx_solver* s, const ReduceCfg& rc, bool two_arrays) {
  const size_t N = (size_t)s->L.nx * s->L.ny;
  int r = reduce_reference_order(s, s->res_a, N, s->red_out);
  if (r != IFX_OK) return r;
  if (two_arrays) {
    if ((r = reduce_reference_order(s, s->res_b, N, s->red_out + 1)) != IFX_OK) return r;
  } else {
    // Poisson: second sum = sum |r| of the same array, same summation order
    if ((r = reduce_reference_order(s, s->res_a, N, s->red_out + 1, true)) != IFX_OK) return r;
  }
  s->launches++;
  IFX_CUDA(s, launch_decide_exact(s->ctl, s->red_out, rc, s->stream));
  return IFX_OK;
}

int ifx::run_ad_loop(ifx_solver* s, ifx_step_stats* st, bool full) {
  const Layout& L = s->L;
  const bool exact = s->opt.reduce_mode == IFX_REDUCE_REFERENCE;
  // the reference's loop starts from uRes = vRes = 1.0 (ADSolver.cu:313-315): a tolerance of 2 or more means no iteration
  const int itermax = (1.0 + 1.0 > s->opt.ad_tol) ? s->in.AD_itermax : 0;
  const int ry = rows_per_cta_for(s, 2);
  const dim3 grid = tile_grid(s, ry, 2);
  const size_t nblocks = (size_t)grid.x * grid.y;
  // the source pass always uses the 256-column direct-load tiles
  const dim3 grid_src((L.nx - 2 + TILE_COLS - 1) / TILE_COLS, grid.y, 1);
  int rc = ensure_partials(s, nblocks);
  if (rc != IFX_OK) return rc;
  if (exact && (rc = ensure_exact_buffers(s)) != IFX_OK) return rc;

  IFX_CUDA(s, cudaEventRecord(s->ev[0], s->stream));
  const int base = s->cur_uv;
  const bool slabs = s->connected && s->opt.nranks > 1;
  const int tw_ad = v4_tile_cols(2);
  if (slabs) {   // halo rows of the starting field (also what the source pass reads)
    const int fi[2] = {base, 2 + base};
    if ((rc = halo_exchange(s, 0, 2, fi, tw_ad, full && bodies_on_slabs(s))) != IFX_OK) return rc;
  }

  // ---- velf + BC + ADSource (ADSolver.cu:298-311) in one pass
  AdSourceArgs sa{};
  sa.L = L; sa.M = s->M;
  sa.u = s->u[base]; sa.v = s->v[base];
  sa.uf = s->uf; sa.vf = s->vf;
  sa.sx = s->sx; sa.sy = s->sy;
  fill_bc(s, sa.two_bc_u, sa.two_bc_v);
  sa.rows_per_cta = ry;
  if (full) {
    IFX_CUDA(s, launch_ad_source(sa, grid_src, s->stream, SRC_FACES));
  } else if (L.nx <= L.ny) {
    IFX_CUDA(s, launch_ad_source(sa, grid_src, s->stream, SRC_REF_VF_ZERO));     // vf == 0 (App. A Q2)
  } else {
    return fail(s, IFX_ERR_INVALID, "IFX_COMPAT_REFERENCE with nx > ny: the reference overruns its vf allocation "
                                    "(preSim.cu:153, ADSolver.cu:179-186); use IFX_COMPAT_FULL");
  }
  s->launches++;

  IFX_CUDA(s, cudaMemsetAsync(s->ctl, 0, sizeof(LoopCtl), s->stream));
  IFX_CUDA(s, cudaEventRecord(s->ev[8], s->stream));        // from here to ev[9]: the Jacobi sweeps (+ ghost-cell kernels)
  if (itermax <= 0) {   // while-condition false on entry: zero iterations (ADSolver.cu:315)
    if (st) { st->ad_iters = 0; st->ad_ures = 1.0; st->ad_vres = 1.0; }
    IFX_CUDA(s, cudaStreamSynchronize(s->stream));
    return IFX_OK;
  }

  AdJacobiArgs ja{};
  ja.L = L; ja.M = s->M;
  ja.sx = s->sx; ja.sy = s->sy; ja.celltype = s->celltype;
  ja.res_u = s->res_a; ja.res_v = s->res_b;
  ja.partials = s->partials; ja.ctl = s->ctl;
  fill_bc(s, ja.two_bc_u, ja.two_bc_v);
  ja.rows_per_cta = ry;
  ja.rc.itermax = itermax; ja.rc.tol = s->opt.ad_tol; ja.rc.use_second = 1; ja.rc.test_abs = 0;
  ja.rc.decide = exact ? 0 : 1;
  ja.rc.certify = (exact || slabs) ? 0 : 1;     // slabs: the global sum is taken as is (+-1 at the rounding edge)
  ja.rc.band = rounding_band(s, nblocks, ry);
  // Slabs: the stop decision LAGS one sweep (every launch posts its partial sums and decides on its predecessor's, a
  // flush kernel decides on the last launch of a batch).  Sweep m+1 then runs although iteration m converged; it
  // writes the buffer of iterate m-1, so iterate m — the result — is intact.  No rank waits for another rank's
  // residual on its critical path any more (76 global lock steps per step in round 1, profiles/r1_scaling.md).
  // Full mode only: the extra sweep also rewrites the ghost ring of ITS input buffer — the result's — which the full
  // step refreshes anyway, whereas the reference-compatible step reads that ring as the reference left it (App. A Q5).
  const bool lag = slabs && full;
  bool lag_first = true;
  const int no_exchange = 0;
  unsigned seq_before = s->seq[0];
  // ghost cells on slabs: sweep (delivers its boundary rows, ghost cells with throw-away values) -> ghost-cell kernel
  // (reads the neighbours' previous iterate over NVLink, closes the ghost cells — also those in the rows the sweep
  // delivered — and publishes the neighbours' gcflag).  Two launches per iteration; the next sweep's boundary tiles wait
  // for both the tile flags and the gcflag (HaloCtx.defer)
  const bool defer = full && bodies_on_slabs(s);
  if (defer && ry < IFX_GC_REACH) return fail(s, IFX_ERR_INVALID, "IFX_ROWS_PER_CTA must be >= 4 with bodies on slabs");

  auto set_iter = [&](int m) {
    const int src = (base + m - 1) & 1;
    ja.uC = s->u[src]; ja.vC = s->v[src];
    ja.uT = s->u[src ^ 1]; ja.vT = s->v[src ^ 1];
    ja.rc.eval_iter = m;
    ja.rc.lag = lag ? 1 : 0; ja.rc.lag_first = lag_first ? 1 : 0;
    ja.rc.no_exchange = no_exchange;
    lag_first = false;
    const int fo[2] = {src ^ 1, 2 + (src ^ 1)};
    make_halo_ctx(s, 0, 2, fo, &ja.hx);
    ja.hx.defer = defer ? 1 : 0;
  };

  int m = 0, fallbacks = 0;
  int batch = std::min(itermax, std::max(1, s->last_ad_iters + 1));
  for (;;) {
    const int todo = std::min(batch, itermax - m);
    for (int b = 0; b < todo; b++) {
      set_iter(++m);
      ja.force = 0;
      if ((rc = launch_ad_jacobi(s, ja, grid, exact)) != IFX_OK) return rc;
      if (full && (s->has_gc || defer)) {   // ghost cells of iterate m from iterate m-1 (Jacobi-lagged, like every other cell)
        s->launches++;
        const int src = (base + m - 1) & 1, dst = src ^ 1;
        GcPush gp{};
        if (defer) {
          const int r = s->opt.rank;
          gp.active = 1;
          gp.has_lo = r > 0; gp.has_hi = r < s->opt.nranks - 1;
          gp.pitch = L.pitch;
          gp.row_lo = (L.jb - L.j0) * L.pitch; gp.row_hi = (L.je - 1 - L.j0) * L.pitch;
          const int fo[2] = {dst, 2 + dst};
          for (int f = 0; f < 2; f++) {
            if (gp.has_lo) gp.dst_lo[f] = peer_field(s, r - 1, fo[f]) + (size_t)(s->peer_nyl[r - 1] - 1) * L.pitch;
            if (gp.has_hi) gp.dst_hi[f] = peer_field(s, r + 1, fo[f]);
          }
          if (gp.has_lo) gp.signal_lo = &peer_sync(s, r - 1)->gcflag[1];
          if (gp.has_hi) gp.signal_hi = &peer_sync(s, r + 1)->gcflag[0];
          gp.seq = ja.hx.seq;
          gp.ticket = reinterpret_cast<unsigned*>(s->d_counters + 2);
        }
        IFX_CUDA(s, launch_gc_velocity(s->gc.count, s->gc.cell, s->gc.stencil, s->gc.w_dir, s->gc.body, s->d_ub, s->d_vb,
                                       ja.uC, ja.vC, gc_peers(s, src, 2 + src), ja.uT, ja.vT, 0, s->ctl, m, s->stream,
                                       defer ? &gp : nullptr));
      }
      if (exact && (rc = exact_decide(s, ja.rc, true)) != IFX_OK) return rc;
    }
    if (lag) {      // the batch's last sweep has no successor to decide on it
      s->launches++;
      IFX_CUDA(s, launch_lag_flush(s->ctl, ja.rc, ja.hx, s->stream));
      lag_first = true;
    }
    if ((rc = fetch_ctl(s)) != IFX_OK) return rc;
    if (s->h_ctl->done && s->h_ctl->ambiguous && !slabs) {
      // the fused sum of iteration h_ctl->iter is within rounding of the tolerance: re-evaluate that
      // iteration's residual in the reference's summation order (iterate and predecessor are intact)
      fallbacks++;
      if ((rc = ensure_exact_buffers(s)) != IFX_OK) return rc;
      ja.res_u = s->res_a; ja.res_v = s->res_b;
      m = s->h_ctl->iter;
      set_iter(m);
      ja.force = 1; ja.rc.decide = 0;
      if ((rc = launch_ad_jacobi(s, ja, grid, true)) != IFX_OK) return rc;
      if (full && s->has_gc) {   // the sweep leaves throw-away values at ghost cells: restore iterate m's closure
        const int src = (base + m - 1) & 1;
        s->launches++;
        IFX_CUDA(s, launch_gc_velocity(s->gc.count, s->gc.cell, s->gc.stencil, s->gc.w_dir, s->gc.body, s->d_ub, s->d_vb,
                                       ja.uC, ja.vC, gc_peers(s, src, 2 + src), ja.uT, ja.vT, 0, nullptr, 0, s->stream));
      }
      if ((rc = exact_decide(s, ja.rc, true)) != IFX_OK) return rc;
      ja.rc.decide = 1;
      if ((rc = fetch_ctl(s)) != IFX_OK) return rc;
    }
    if (s->h_ctl->done) break;
    if (m >= itermax) return fail(s, IFX_ERR_STATE, "predictor loop ran past AD_itermax without a decision");
    batch = 2;
  }
  IFX_CUDA(s, cudaEventRecord(s->ev[9], s->stream));
  const int K = s->h_ctl->iter;
  s->cur_uv = (base + K) & 1;
  s->last_ad_iters = K;
  std::memcpy(s->ad_hist, s->h_ctl->hist, sizeof(s->ad_hist));
  if (slabs && (rc = halo_wait(s, 0, seq_before + (unsigned)K, tw_ad)) != IFX_OK) return rc;   // rows of iterate K have landed
  if (K == 1 && !full) {   // final buffer's ghost ring was never written this step: give it BC(start field)
    s->launches++;
    IFX_CUDA(s, launch_copy_ring(L, s->u[base], s->u[s->cur_uv], s->v[base], s->v[s->cur_uv], s->stream));
  }
  IFX_CUDA(s, cudaEventRecord(s->ev[1], s->stream));
  IFX_CUDA(s, cudaEventSynchronize(s->ev[1]));
  if (st) {
    st->ad_iters = K;
    st->ad_ures = s->h_ctl->res0;
    st->ad_vres = s->h_ctl->res1;
    st->exact_fallbacks += fallbacks;
    cudaEventElapsedTime(&st->ms_ad, s->ev[0], s->ev[1]);
    cudaEventElapsedTime(&st->ms_ad_sweeps, s->ev[8], s->ev[9]);
  }
  return check_residual_finite(s, "predictor");
}

// ------------------------------------------------------------------------------------------------
// Poisson — ImmerseFlow::PPESolver(), PPESolver.cu:137-205
// ------------------------------------------------------------------------------------------------
// face masks of the owned rows from the current cell types (the halo rows' types are classified locally, so the
// first / last owned row of a slab sees its true neighbours)
int ifx::ensure_facemask(ifx_solver* s) {
  if (s->facemask_valid) return IFX_OK;
  if (!s->facemask) return fail(s, IFX_ERR_STATE, "face masks exist in IFX_COMPAT_FULL only");
  s->launches++;
  IFX_CUDA(s, launch_build_facemask(s->L, s->celltype, s->facemask, 1, s->L.nyl - 1, s->stream));
  s->facemask_valid = true;
  return IFX_OK;
}

int ifx::enqueue_ppe_sweep(ifx_solver* s, PpeSweepArgs& a, dim3 grid, bool laplace_ref, bool write_res) {
  s->launches++;
  a.wide = (!laplace_ref && ppe_wide_tiles(s)) ? 1 : 0;       // the geometry `grid` was sized for (tile_grid, mode 1)
  IFX_CUDA(s, ifx::launch_ppe_sweep_v4(a, grid, s->stream, laplace_ref, write_res));
  return IFX_OK;
}

int ifx::run_ppe_loop(ifx_solver* s, ifx_step_stats* st, bool laplace_ref) {
  const Layout& L = s->L;
  const bool exact = s->opt.reduce_mode == IFX_REDUCE_REFERENCE;
  // the reference's loop starts from res = 1.0 (PPESolver.cu:170-172): a tolerance of 1 or more means no sweep at all
  const int itermax = (1.0 > s->opt.ppe_tol) ? s->in.PPE_itermax : 0;
  const int pmode = laplace_ref ? 0 : 1;
  const int ry = rows_per_cta_for(s, pmode);
  const dim3 grid = tile_grid(s, ry, pmode);
  const size_t nblocks = (size_t)grid.x * grid.y;
  int rc = ensure_partials(s, nblocks);
  if (rc != IFX_OK) return rc;
  if (exact && (rc = ensure_exact_buffers(s)) != IFX_OK) return rc;

  IFX_CUDA(s, cudaEventRecord(s->ev[2], s->stream));
  const int base = s->cur_p;
  const bool slabs = s->connected && s->opt.nranks > 1;
  const int tw_ppe = tile_cols_for(s, pmode);
  if (laplace_ref) {
    // set_pressure_BC (PPESolver.cu:164); the ring is then carried through every sweep (:21)
    s->launches += 2;
    IFX_CUDA(s, launch_set_pressure_bc_ref(L, s->p[base], nullptr, s->stream));
    for (int k = 1; k < s->np; k++) {
      if (k > 1) s->launches++;
      IFX_CUDA(s, launch_copy_ring(L, s->p[base], s->p[(base + k) % s->np], nullptr, nullptr, s->stream));
    }
  }
  IFX_CUDA(s, cudaMemsetAsync(s->ctl, 0, sizeof(LoopCtl), s->stream));
  int K = 0, fallbacks = 0;
  if (slabs) {
    const int fi[1] = {4 + base};
    if ((rc = halo_exchange(s, 1, 1, fi, tw_ppe)) != IFX_OK) return rc;
  }
  const unsigned seq_before = s->seq[1];
  if (itermax > 0) {
    PpeSweepArgs pa{};
    pa.L = L; pa.M = s->M;
    if (!laplace_ref && (rc = ensure_facemask(s)) != IFX_OK) return rc;
    pa.rhs = s->rhs; pa.facemask = s->facemask;
    pa.res = s->res_a; pa.partials = s->partials; pa.ctl = s->ctl;
    pa.rows_per_cta = ry;
    pa.rc.itermax = itermax; pa.rc.tol = s->opt.ppe_tol; pa.rc.use_second = 0;
    pa.rc.test_abs = s->opt.ppe_abs_residual ? 1 : 0;
    pa.rc.decide = exact ? 0 : 1;
    pa.rc.certify = (exact || slabs) ? 0 : 1;
    pa.rc.band = rounding_band(s, nblocks, ry);
    // red-black SOR: iteration m = colour-0 half-sweep base -> partner (it also evaluates the
    // residual of iterate m-1 and takes the stop decision, like a Jacobi sweep) + colour-1 half-sweep partner -> base.
    // The iterate always ends up in the buffer it started in.
    const bool sor = !laplace_ref && s->opt.ppe_solver == 3;
    pa.sor = sor ? 1 : 0; pa.sor_omega = s->opt.ppe_omega;
    const int decide = pa.rc.decide;
    // Slabs, Jacobi: lagged stop decision (see run_ad_loop).  Sweep m evaluates the residual of iterate m-1; the
    // decision falls at the end of sweep m+1, which has written iterate m+1 by then — so the buffers rotate three
    // ways and iterate m-1, the result, is still intact.  Red-black SOR keeps the lock step (its iterate lives in one
    // buffer; a late decision would find it overwritten).
    const int np = s->np;
    const bool lag = slabs && !sor;
    bool lag_first = true;
    const int no_exchange = 0;      // (1 only for the forced re-creation sweeps below: nothing to decide, nothing to exchange)
    const int partner = (base + 1) % np;      // SOR: the other buffer of the pair
    auto set_sweep = [&](int m) {     // sweep m: iterate m-1 -> iterate m, evaluates residual(iterate m-1)
      const int src = sor ? base : (base + m - 1) % np;
      const int dst = sor ? partner : (base + m) % np;
      pa.pC = s->p[src]; pa.pT = s->p[dst];
      pa.rc.eval_iter = m - 1;
      pa.rc.decide = decide;
      pa.rc.lag = lag ? 1 : 0; pa.rc.lag_first = lag_first ? 1 : 0;
      pa.rc.no_exchange = no_exchange;
      lag_first = false;
      pa.sor_colour = 0;
      const int fo[1] = {4 + dst};
      make_halo_ctx(s, 1, 1, fo, &pa.hx);
    };
    auto set_black = [&]() {          // second half of a SOR iteration: no residual bookkeeping
      pa.pC = s->p[partner]; pa.pT = s->p[base];
      pa.rc.eval_iter = 0; pa.rc.decide = 0;
      pa.sor_colour = 1;
      const int fo[1] = {4 + base};
      make_halo_ctx(s, 1, 1, fo, &pa.hx);
    };
    if (sor) {
      int m = 0;
      for (;;) {
        const int todo = std::min(s->opt.sweeps_per_batch, itermax + 1 - m);
        for (int b = 0; b < todo; b++) {
          set_sweep(++m);
          pa.force = 0;
          if ((rc = enqueue_ppe_sweep(s, pa, grid, laplace_ref, exact)) != IFX_OK) return rc;
          if (exact && m > 1 && (rc = exact_decide(s, pa.rc, false)) != IFX_OK) return rc;
          set_black();
          if ((rc = enqueue_ppe_sweep(s, pa, grid, laplace_ref, false)) != IFX_OK) return rc;
        }
        if ((rc = fetch_ctl(s)) != IFX_OK) return rc;
        if (s->h_ctl->done && s->h_ctl->ambiguous && !slabs) {
          fallbacks++;
          if ((rc = ensure_exact_buffers(s)) != IFX_OK) return rc;
          pa.res = s->res_a;
          m = s->h_ctl->iter + 1;          // the sweep that evaluated the ambiguous residual
          set_sweep(m);
          pa.force = 1; pa.rc.decide = 0;
          if ((rc = enqueue_ppe_sweep(s, pa, grid, laplace_ref, true)) != IFX_OK) return rc;
          if ((rc = exact_decide(s, pa.rc, false)) != IFX_OK) return rc;
          pa.rc.decide = 1;
          if ((rc = fetch_ctl(s)) != IFX_OK) return rc;
          if (!s->h_ctl->done) {           // not converged after all: the iteration's second half was skipped, run it
            set_black();
            pa.force = 0;
            if ((rc = enqueue_ppe_sweep(s, pa, grid, laplace_ref, false)) != IFX_OK) return rc;
          }
        }
        if (s->h_ctl->done) break;
        if (m >= itermax + 1) return fail(s, IFX_ERR_STATE, "Poisson loop ran past PPE_itermax without a decision");
      }
      K = s->h_ctl->iter;
      s->cur_p = base;
      if (slabs && (rc = halo_wait(s, 1, seq_before + (unsigned)(2 * K), tw_ppe)) != IFX_OK) return rc;
    } else {
      // ---- point Jacobi.  A launch starts from the newest stored iterate `it` (buffer `cur`) and is either a single
      // sweep (it -> it+1, evaluates r(p_it)) or a PAIR (kernels_pair.cu: it -> it+2 in one pass over memory, evaluates
      // r(p_it) and r(p_it+1); iterate it+1 exists in shared memory only).  where[k] = buffer holding iterate k, or -1.
      std::vector<int> where((size_t)itermax + 4, -1);
      std::vector<unsigned> seq_at((size_t)itermax + 4, seq_before);
      where[0] = base;
      int it = 0, cur = base;
      const int nyi = L.je - L.jb;
      const bool pairs = s->opt.ppe_pairs && !laplace_ref && !exact && !slabs && nyi >= 4;
      PpeSweepArgs pp = pa;                            // pair launches: own tile geometry, 4 partial sums per CTA
      dim3 grid2(1, 1, 1);
      if (pairs) {
        const int tw2 = pair_tile_cols();
        const int ry2 = std::max(2, std::min(ry, 255));
        int ty2 = (nyi + ry2 - 1) / ry2;
        if (ty2 > 1 && nyi - (ty2 - 1) * ry2 == 1) ty2 -= 1;      // the last tile takes the odd row: never a 1-row tile
        grid2 = dim3((L.nx - 2 + tw2 - 1) / tw2, ty2, 1);
        if ((rc = ensure_partials(s, 2 * (size_t)grid2.x * grid2.y)) != IFX_OK) return rc;
        pa.partials = s->partials;
        pp = pa;
        pp.rows_per_cta = ry2;
        pp.rc.band = rounding_band(s, (size_t)grid2.x * grid2.y, 2 * (ry2 + 1));
      }
      // one launch; eval = index of the iterate whose residual it evaluates (single) / of the first of the two (pair)
      auto launch = [&](bool pair, int src, int dst, int eval, int force, int decide_now, bool wres, bool exchange) -> int {
        PpeSweepArgs& q = pair ? pp : pa;
        q.pC = s->p[src]; q.pT = s->p[dst];
        q.rc.eval_iter = eval;
        q.rc.decide = decide_now;
        q.rc.lag = (lag && exchange) ? 1 : 0; q.rc.lag_first = lag_first ? 1 : 0;
        q.rc.no_exchange = exchange ? no_exchange : 1;
        if (exchange) lag_first = false;
        q.force = force;
        q.sor_colour = 0;
        const int fo[1] = {4 + dst};
        make_halo_ctx(s, 1, 1, fo, &q.hx);
        if (pair) {
          s->launches++;
          IFX_CUDA(s, launch_ppe_pair(q, grid2, s->stream));
          return IFX_OK;
        }
        return enqueue_ppe_sweep(s, q, grid, laplace_ref, wres);
      };
      // re-create iterate k, which a pair only held in shared memory, from iterate k-1 (one forced single sweep)
      auto recreate = [&](int k) -> int {
        const int src = where[k - 1], dst = (src + 1) % np;
        if (src < 0) return fail(s, IFX_ERR_STATE, "Poisson loop: the predecessor of an unstored iterate is gone");
        int r2 = launch(false, src, dst, 0, 1, 0, false, false);
        if (r2 != IFX_OK) return r2;
        where[k] = dst; seq_at[k] = pa.hx.seq;
        for (size_t z = (size_t)k + 1; z < where.size(); z++) where[z] = -1;
        return IFX_OK;
      };
      const PpeSweepArgs* last = &pa;
      for (;;) {
        int launched = 0;
        while (launched < s->opt.sweeps_per_batch && it <= itermax) {
          const int dst = (cur + 1) % np;
          const bool pair = pairs && it + 1 <= itermax;
          if ((rc = launch(pair, cur, dst, it, 0, decide, exact, true)) != IFX_OK) return rc;
          last = pair ? &pp : &pa;
          if (exact && it >= 1 && (rc = exact_decide(s, pa.rc, false)) != IFX_OK) return rc;
          it += pair ? 2 : 1;
          where[it] = dst; seq_at[it] = last->hx.seq;
          cur = dst;
          launched++;
        }
        if (lag && launched > 0) {
          s->launches++;
          IFX_CUDA(s, launch_lag_flush(s->ctl, last->rc, last->hx, s->stream));
          lag_first = true;
        }
        if ((rc = fetch_ctl(s)) != IFX_OK) return rc;
        if (s->h_ctl->done && s->h_ctl->ambiguous && !slabs) {
          // the fused sum of iterate q is within rounding of the tolerance: evaluate it in the reference's summation
          // order — a forced single sweep from iterate q that writes the residual array (and iterate q+1)
          fallbacks++;
          if ((rc = ensure_exact_buffers(s)) != IFX_OK) return rc;
          pa.res = s->res_a;
          const int q = s->h_ctl->iter;
          if (where[q] < 0 && (rc = recreate(q)) != IFX_OK) return rc;
          const int src = where[q], dst = (src + 1) % np;
          if ((rc = launch(false, src, dst, q, 1, 0, true, false)) != IFX_OK) return rc;
          if ((rc = exact_decide(s, pa.rc, false)) != IFX_OK) return rc;
          if ((rc = fetch_ctl(s)) != IFX_OK) return rc;
          for (size_t z = (size_t)q + 1; z < where.size(); z++) where[z] = -1;
          where[q + 1] = dst; seq_at[q + 1] = pa.hx.seq;
          it = q + 1; cur = dst;
        }
        if (s->h_ctl->done) break;
        if (it > itermax) return fail(s, IFX_ERR_STATE, "Poisson loop ran past PPE_itermax without a decision");
      }
      K = s->h_ctl->iter;
      if (where[K] < 0 && (rc = recreate(K)) != IFX_OK) return rc;      // the rule fired on a pair's intermediate iterate
      s->cur_p = where[K];
      if (slabs && (rc = halo_wait(s, 1, seq_at[K], tw_ppe)) != IFX_OK) return rc;
    }
  }
  if (laplace_ref) {   // PPESolver.cu:195
    s->launches++;
    IFX_CUDA(s, launch_set_pressure_bc_ref(L, s->p[s->cur_p], nullptr, s->stream));
  }
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
