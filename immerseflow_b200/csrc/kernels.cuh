// kernels.cuh — kernel argument blocks, launch geometry and the fused residual reduction.
#pragma once
#include "common.cuh"
#include "facemask.cuh"

namespace ifx {

constexpr int AD_WARPS = 4;                 // warps per CTA, side by side in x
constexpr int AD_THREADS = AD_WARPS * 32;   // 128 threads -> 256 columns per CTA
constexpr int TILE_COLS = AD_WARPS * 64;

// How a sweep's residual feeds the stop rule.
struct ReduceCfg {
  int eval_iter;    // index of the iterate whose residual this launch evaluates (0: no decision)
  int itermax;
  int decide;       // 0: leave the decision to the reference-order reduction kernels
  int use_second;   // stop test on res0 + res1 (predictor: uRes + vRes, ADSolver.cu:315)
  int test_abs;     // PPE: test res1 (= sum |r|) instead of the signed res0
  int certify;      // flag decisions that fall inside the rounding band instead of taking them
  int lag;          // slabs: post this launch's partials, decide on the PREVIOUS launch's (no cross-GPU wait on the
                    // critical path); the loop's last launch is decided by a flush kernel (launch_lag_flush)
  int lag_first;    // lag: first launch of the loop or after a flush — nothing to decide yet
  int no_exchange;  // slabs: this launch takes part in no residual exchange (forced re-creation of an iterate: no decision)
  double tol;
  double band;      // relative half-width of the rounding band (times the abs-sum)
};

struct AdJacobiArgs {
  Layout L;
  Metrics M;
  double* uC; double* vC;             // input iterate (its ghost ring is written, see kernel)
  double* uT; double* vT;             // output iterate
  const double* sx; const double* sy;
  const uint8_t* celltype;
  double* res_u; double* res_v;       // reference-layout residual arrays (WRITE_RES only)
  double* partials;                   // 2 doubles per CTA
  LoopCtl* ctl;
  ReduceCfg rc;
  double two_bc_u[4], two_bc_v[4];    // 2*bc for W, E, S, N
  int rows_per_cta;
  int force;                          // run even if ctl->done (reference-order re-evaluation)
  HaloCtx hx;                         // slab neighbours (nranks == 1: unused)
};

struct AdSourceArgs {
  Layout L;
  Metrics M;
  const double* u; const double* v;
  const double* uf; const double* vf;
  double* sx; double* sy;
  double two_bc_u[4], two_bc_v[4];
  int rows_per_cta;
};

struct PpeSweepArgs {
  Layout L;
  Metrics M;
  double* pC;                         // input iterate
  double* pT;                         // output iterate
  const double* rhs;
  const uint8_t* facemask;            // general operator: face mask per cell (k_build_facemask); Laplace: unused
  double* res;                        // reference-layout residual array (WRITE_RES only)
  double* partials;
  LoopCtl* ctl;
  ReduceCfg rc;
  int rows_per_cta;
  int force;
  int wide;                           // general operator: four columns per thread / 512-column tiles (capi.cu: ppe_wide_tiles)
  int sor, sor_colour;                // red-black SOR half-sweep of the given colour instead of a Jacobi sweep
  double sor_omega;
  HaloCtx hx;
};

// ---------------------------------------------------------------------------------------------
// Fused residual reduction: thread partials -> warp shuffle -> shared memory -> one pair of
// doubles per CTA in global memory -> the LAST CTA to finish (atomic ticket) adds the CTA
// partials in a fixed order and takes the stop decision on the device.  Deterministic: the
// order depends only on the launch geometry, never on scheduling.
// ---------------------------------------------------------------------------------------------
// global residual over the slabs: every rank posts its partial into every rank's mailbox (NVLink P2P stores),
// then adds all of them in rank order — same operands, same order, same stop decision everywhere.  One thread of
// one CTA per launch runs this: kept out of line so that it does not set the register count of the sweep kernels.
// Two halves: POST (my partial of launch `tag` into everybody's mailbox) and COLLECT (add the partials of launch
// `tag`, waiting for stragglers).  A lock-step launch does both for its own tag; a LAGGED launch posts its own and
// collects the previous launch's, which arrived a whole sweep ago — the ranks then drift by up to one launch instead
// of meeting 76 times per step (profiles/r1_scaling.md).
// A mailbox entry carries up to two residual pairs (a launch of the pair kernel, kernels_pair.cu, evaluates two
// iterates) and says so itself: {a0, b0, a1, b1, number of pairs, iterate index of the first pair}.
static __device__ __noinline__ void slab_post_partial(const HaloCtx& hx, unsigned tag, const double* v, int npairs,
                                                      int eval_first) {
  // Fire and forget: every 8-byte packet carries 32 bits of data and the launch tag, so the reader needs neither a flag
  // nor a fence — and the one thread that runs at the tail of every launch does not sit out an NVLink round trip
  // (measured at 8 GPUs: the fenced data-then-flag version cost 9 us per launch, 3.5 % of the step).
  const unsigned slot = tag & (IFX_MAIL_SLOTS - 1);
  const unsigned me = (unsigned)hx.rank;
  const double w[IFX_MAIL_VALS] = {v[0], v[1], v[2], v[3], (double)npairs, (double)eval_first};
  for (int r = 0; r < hx.nranks; ++r) {
    unsigned long long* m = hx.mail[r] + ((size_t)slot * IFX_MAX_RANKS + me) * (2 * IFX_MAIL_VALS);
#pragma unroll
    for (int q = 0; q < IFX_MAIL_VALS; ++q) {
      const unsigned long long bits = (unsigned long long)__double_as_longlong(w[q]);
      st_relaxed_sys_u64(m + 2 * q, ((unsigned long long)tag << 32) | (bits & 0xffffffffull));
      st_relaxed_sys_u64(m + 2 * q + 1, ((unsigned long long)tag << 32) | (bits >> 32));
    }
  }
}
static __device__ __noinline__ void slab_collect_partials(const HaloCtx& hx, unsigned tag, double* v, int& npairs,
                                                          int& eval_first) {
  const unsigned slot = tag & (IFX_MAIL_SLOTS - 1);
  const unsigned me = (unsigned)hx.rank;
  double g[4] = {0.0, 0.0, 0.0, 0.0};
  for (int r = 0; r < hx.nranks; ++r) {      // rank order: same operands, same order, same sums on every rank
    const unsigned long long* m = hx.mail[me] + ((size_t)slot * IFX_MAX_RANKS + r) * (2 * IFX_MAIL_VALS);
    // all twelve packets of a rank are loaded back to back (independent loads: one L2 round trip, not twelve) and
    // re-read together until every one carries the tag — with a lagged decision they arrived a sweep ago
    unsigned long long pk[2 * IFX_MAIL_VALS];
    bool all;
    do {
#pragma unroll
      for (int q = 0; q < 2 * IFX_MAIL_VALS; ++q) pk[q] = ld_relaxed_sys_u64(m + q);
      all = true;
#pragma unroll
      for (int q = 0; q < 2 * IFX_MAIL_VALS; ++q) all = all && ((pk[q] >> 32) == tag);
      if (!all) __nanosleep(64);
    } while (!all);
    double w[IFX_MAIL_VALS];
#pragma unroll
    for (int q = 0; q < IFX_MAIL_VALS; ++q)
      w[q] = __longlong_as_double((long long)((pk[2 * q + 1] << 32) | (pk[2 * q] & 0xffffffffull)));
    g[0] += w[0]; g[1] += w[1]; g[2] += w[2]; g[3] += w[3];
    if (r == 0) { npairs = (int)w[4]; eval_first = (int)w[5]; }
  }
  v[0] = g[0]; v[1] = g[1]; v[2] = g[2]; v[3] = g[3];
}

// the stop decision on a pair of global residual sums (one thread)
__device__ __forceinline__ void decide_on_residual(LoopCtl* ctl, const ReduceCfg& rc, int eval_iter, double a, double b) {
  ctl->res0 = a;
  ctl->res1 = b;
  if (eval_iter < 1) return;
  if (eval_iter <= 64) { ctl->hist[2 * (eval_iter - 1)] = a; ctl->hist[2 * (eval_iter - 1) + 1] = b; }
  if (!rc.decide) return;
  ctl->iter = eval_iter;
  const double S = rc.use_second ? a + b : (rc.test_abs ? b : a);
  const double A = rc.use_second ? S : b;           // sum of magnitudes
  const bool at_max = eval_iter >= rc.itermax;
  if (rc.certify && !at_max && fabs(S - rc.tol) <= rc.band * A) {
    ctl->ambiguous = 1;
    ctl->done = 1;
  } else if (!(S > rc.tol) || at_max) {
    ctl->hit_max = (S > rc.tol) ? 1 : 0;
    ctl->done = 1;
  }
}
// ... on the one or two iterates a launch evaluated, in order; the second only if the first did not end the loop
__device__ __forceinline__ void decide_on_residuals(LoopCtl* ctl, const ReduceCfg& rc, int eval_first, int npairs, const double* v) {
  decide_on_residual(ctl, rc, eval_first, v[0], v[1]);
  if (npairs > 1 && !ctl->done) decide_on_residual(ctl, rc, eval_first + 1, v[2], v[3]);
}

// what the one thread that holds a launch's global-memory sums does with them: single GPU -> decide; slabs -> post
// them, then decide on this launch's sums (lock step: wait for every rank) or on the PREVIOUS launch's (lagged)
__device__ __forceinline__ void finish_residuals(LoopCtl* ctl, const ReduceCfg& rc, const HaloCtx* hx, double* v, int npairs) {
  int eval_first = rc.eval_iter;
  bool decide = true;
  if (hx && hx->nranks > 1 && !rc.no_exchange) {
    slab_post_partial(*hx, hx->mseq, v, npairs, eval_first);
    if (rc.lag) {                        // the previous launch's sums (every rank posted them a sweep ago)
      decide = !rc.lag_first;
      if (decide) slab_collect_partials(*hx, hx->mseq - 1, v, npairs, eval_first);
    } else {
      slab_collect_partials(*hx, hx->mseq, v, npairs, eval_first);
    }
  }
  if (decide) decide_on_residuals(ctl, rc, eval_first, npairs, v);
}

// NP = 1: (r[0], r[1]) is the residual pair of rc.eval_iter; NP = 2: (r[2], r[3]) in addition, of rc.eval_iter + 1
template <int THREADS, int NP>
__device__ __forceinline__ void block_reduce_and_decide_n(double (&r)[2 * NP], double* partials, LoopCtl* ctl,
                                                          const ReduceCfg& rc, unsigned bid, unsigned nblocks,
                                                          const HaloCtx* hx) {
  constexpr int NW = THREADS / 32;
  constexpr int NV = 2 * NP;
  __shared__ double sh[NV][NW];
  __shared__ int s_last;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int q = 0; q < NV; ++q) r[q] = warp_sum(r[q]);
  if (lane == 0) {
#pragma unroll
    for (int q = 0; q < NV; ++q) sh[q][warp] = r[q];
  }
  __syncthreads();
  if (threadIdx.x == 0) {
#pragma unroll
    for (int q = 0; q < NV; ++q) {
      double a = sh[q][0];
#pragma unroll
      for (int w = 1; w < NW; ++w) a += sh[q][w];
      partials[NV * bid + q] = a;
    }
    __threadfence();
    const unsigned t = atomicAdd(&ctl->ticket, 1u);
    s_last = (t == nblocks - 1);
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  double acc[NV];
#pragma unroll
  for (int q = 0; q < NV; ++q) acc[q] = 0.0;
  for (unsigned k = threadIdx.x; k < nblocks; k += THREADS) {
#pragma unroll
    for (int q = 0; q < NV; ++q) acc[q] += __ldcg(partials + NV * k + q);
  }
#pragma unroll
  for (int q = 0; q < NV; ++q) acc[q] = warp_sum(acc[q]);
  __syncthreads();
  if (lane == 0) {
#pragma unroll
    for (int q = 0; q < NV; ++q) sh[q][warp] = acc[q];
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double v[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
    for (int q = 0; q < NV; ++q) {
      double a = sh[q][0];
#pragma unroll
      for (int w = 1; w < NW; ++w) a += sh[q][w];
      v[q] = a;
    }
    ctl->ticket = 0;
    finish_residuals(ctl, rc, hx, v, NP);
    __threadfence();
  }
}

template <int THREADS>
__device__ __forceinline__ void block_reduce_and_decide(double r0, double r1, double* partials,
                                                        LoopCtl* ctl, const ReduceCfg& rc,
                                                        unsigned bid, unsigned nblocks, const HaloCtx* hx = nullptr) {
  double r[2] = {r0, r1};
  block_reduce_and_decide_n<THREADS, 1>(r, partials, ctl, rc, bid, nblocks, hx);
}
template <int THREADS>
__device__ __forceinline__ void block_reduce_and_decide_pair(double a0, double a1, double b0, double b1, double* partials,
                                                             LoopCtl* ctl, const ReduceCfg& rc, unsigned bid,
                                                             unsigned nblocks, const HaloCtx* hx = nullptr) {
  double r[4] = {a0, a1, b0, b1};
  block_reduce_and_decide_n<THREADS, 2>(r, partials, ctl, rc, bid, nblocks, hx);
}

// ---------------------------------------------------------------------------------------------
// Host-side launchers.  Every __global__ is launched from the translation unit that defines it.
// ---------------------------------------------------------------------------------------------
// kernels_ad.cu
enum AdSourceVariant { SRC_REF_VF_ZERO = 0, SRC_REF_VF_ARRAY = 1, SRC_FACES = 2 };
cudaError_t launch_ad_source(const AdSourceArgs& a, dim3 grid, cudaStream_t st, AdSourceVariant v);
cudaError_t launch_copy_ring(const Layout& L, const double* s0, double* d0, const double* s1, double* d1,
                             cudaStream_t st);
// kernels_ppe.cu
cudaError_t launch_set_pressure_bc_ref(const Layout& L, double* p0, double* p1, cudaStream_t st);
// kernels_v4.cu — the sweep kernels: bulk-copy row pipeline, lean interior path, in-line shared-reciprocal division
cudaError_t launch_ad_jacobi_v4(const AdJacobiArgs& a, dim3 grid, cudaStream_t st, bool write_res);
cudaError_t launch_ppe_sweep_v4(const PpeSweepArgs& a, dim3 grid, cudaStream_t st, bool laplace_ref, bool write_res);
// kernels_pair.cu — two Jacobi sweeps of the general Poisson operator per pass over memory (opt-in, single GPU: measured
// bit-identical but not faster, see the file's header); a.rc.eval_iter = index of the INPUT iterate; a.partials holds
// 4 doubles per CTA
cudaError_t launch_ppe_pair(const PpeSweepArgs& a, dim3 grid, cudaStream_t st);
int pair_tile_cols();
int v4_tile_cols(int mode /*0: Laplace, 1: general Poisson, 2: predictor*/, bool wide = false /*mode 1: PpeSweepArgs::wide*/);
// kernels_full.cu — PPE source term, projection, BC refresh (IFX_COMPAT_FULL)
cudaError_t launch_apply_ring(const Layout& L, double* q, const double* two_bc, int neumann, cudaStream_t st);
cudaError_t launch_faces_init(const Layout& L, const Metrics& M, const uint8_t* ct, const double* ub, const double* vb,
                              const double* u, const double* v, double* uf, double* vf, cudaStream_t st);
cudaError_t launch_ppe_rhs(const Layout& L, const Metrics& M, const uint8_t* ct, const double* ub, const double* vb,
                           const double* u, const double* v, double* rhs, cudaStream_t st);
cudaError_t launch_correct(const Layout& L, const Metrics& M, const uint8_t* ct, const double* ub, const double* vb,
                           const double* us, const double* vs, const double* p, double* un, double* vn, double* uf,
                           double* vf, cudaStream_t st);
// kernels_ib.cu — classification, ghost-cell list and stencils, ghost-cell values
struct BodySet {                   // marker polygons of all bodies (device pointers)
  int nbodies;
  const int* off;                  // nbodies+1 offsets into xm / ym
  const double* xm; const double* ym;
  const double* bbox;              // [xmin, xmax, ymin, ymax] per body
};
struct SlabGeom {                  // what the ghost-cell stencils need to know about the neighbour slabs
  int has_lo, has_hi;
  int nyl_lo, nyl_hi;              // rows the neighbours store (their owned rows + 2)
};
struct GcPeers {                   // the source field(s) of a ghost-cell evaluation in the neighbours' memory
  const double* lo[2];
  const double* hi[2];
};
cudaError_t launch_classify(const Layout& L, const double* xc, const double* yc, const BodySet& B, uint8_t* celltype,
                            int* rowcount, cudaStream_t st);
cudaError_t launch_gc_count(const Layout& L, const int* rowcount, int* rowstart, int* total, cudaStream_t st);
cudaError_t launch_gc_build(const Layout& L, const double* xc, const double* yc, const BodySet& B, const SlabGeom& sg,
                            const uint8_t* celltype, const int* rowstart, int ngc, int* cell, int* ref_id,
                            int* body, int* stencil, int* stencil_ref, double* wd, double* wn, double* bi, double* ip,
                            int* err, cudaStream_t st);
// in-loop ghost-cell kernel of a slab run (HaloCtx.defer): besides closing its ghost cells it overwrites those that lie
// in the slab's first / last owned row in the neighbours' halo rows (the sweep delivered throw-away values there) and,
// when the whole grid is through, publishes the neighbours' gcflag
struct GcPush {
  int active;
  int has_lo, has_hi;
  int row_lo, row_hi;              // padded offset of the start of my first / last owned row (lidx(L, 0, jl) - IFX_PADL)
  int pitch;
  double* dst_lo[2];               // lower neighbour's top halo row start (u, v); upper neighbour's bottom halo row start
  double* dst_hi[2];
  unsigned* signal_lo; unsigned* signal_hi;
  unsigned seq;
  unsigned* ticket;
};
cudaError_t launch_gc_velocity(int ngc, const int* cell, const int* stencil, const double* wd, const int* body,
                               const double* ub, const double* vb, const double* usrc, const double* vsrc,
                               const GcPeers& pr, double* udst, double* vdst, int gather, const LoopCtl* ctl, int iter,
                               cudaStream_t st, const GcPush* push = nullptr);
cudaError_t launch_gc_pressure(int ngc, const int* cell, const int* stencil, const double* wn, const double* psrc,
                               const GcPeers& pr, double* pdst, int gather, cudaStream_t st);
cudaError_t launch_gc_scatter(int ngc, const int* cell, const double* a, double* qa, const double* b, double* qb,
                              cudaStream_t st);
// kernels_halo.cu — halo delivery outside the sweep kernels (between stages) and flag waits
struct HaloPushArgs {
  Layout L;
  int nfields;
  const double* src[4];            // local fields
  double* dst_lo[4];               // lower neighbour's top-halo row start per field (or null)
  double* dst_hi[4];
  int has_lo, has_hi;
  unsigned seq;
  unsigned* signal_lo; unsigned* signal_hi;
  unsigned* gc_signal_lo; unsigned* gc_signal_hi;   // also publish the neighbours' gcflag (start of a predictor loop) or null
  int tile_cols, ntiles;
  const LoopCtl* ctl;              // in-loop use: skip when the loop finished before iteration `iter` (else null)
  int iter;
};
cudaError_t launch_halo_push(const HaloPushArgs& a, cudaStream_t st);
cudaError_t launch_halo_wait(const unsigned* wait_lo, const unsigned* wait_hi, int ntiles, unsigned need, cudaStream_t st);
// lagged stop decision: decide on the LAST launch of a batch (rc, hx: that launch's), unless the loop is already done
cudaError_t launch_lag_flush(LoopCtl* ctl, const ReduceCfg& rc, const HaloCtx& hx, cudaStream_t st);
// kernels_reduce.cu — the reference's summation order (preSim.cu:12-50, 376-441)
cudaError_t launch_reduce6(const double* in, size_t n, double* partial, double* out, cudaStream_t st,
                           bool abs_values = false);
cudaError_t launch_decide_exact(LoopCtl* ctl, const double* sums, const ReduceCfg& rc, cudaStream_t st);
// kernels_misc.cu / kernels_ic.cu
cudaError_t launch_init_vortex(const Layout& L, const double* xc, const double* yc, double* u, double* v, double* p,
                               cudaStream_t st);
cudaError_t launch_fill_u8(uint8_t* p, size_t n, uint8_t v, cudaStream_t st);
// (face masks: facemask.cuh)
cudaError_t launch_copy_words(void* dst, const void* src, size_t bytes, cudaStream_t st);   // bytes: multiple of 4
cudaError_t launch_pack_u8(const Layout& L, const uint8_t* padded, double* dense, int raw, cudaStream_t st);

}  // namespace ifx
