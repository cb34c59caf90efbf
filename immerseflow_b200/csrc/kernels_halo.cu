// kernels_halo.cu — halo rows between slabs, outside the sweep kernels (which deliver their own boundary rows,
// see kernels_v4.cu): after the projection, at the start of a solve, after a state upload.
// Rows are contiguous (id = i + j*nx), so one halo message is one row segment per field, stored straight into
// the neighbour's mapped memory over NVLink; each tile-wide block then publishes the group's sequence number
// with a system-scope release store, exactly like the sweep kernels do, so consumers cannot tell the difference.
#include "kernels.cuh"

namespace ifx {

static __global__ void k_halo_push(HaloPushArgs a) {
  if (a.ctl && a.ctl->done && a.ctl->iter < a.iter) return;
  const Layout& L = a.L;
  const int b = blockIdx.x;
  // tile b covers interior columns [1 + b*TW, 1 + (b+1)*TW); the first / last tile also carry the ghost columns
  int c0 = 1 + b * a.tile_cols, c1 = min(c0 + a.tile_cols, L.nx - 1);
  if (b == 0) c0 = 0;
  if (b == a.ntiles - 1) c1 = L.nx;
  const int jl_lo = L.jb - L.j0, jl_hi = L.je - 1 - L.j0;        // my first / last owned row (local index)
  for (int f = 0; f < a.nfields; ++f) {
    const double* src = a.src[f];
    for (int i = c0 + threadIdx.x; i < c1; i += blockDim.x) {
      if (a.has_lo) a.dst_lo[f][IFX_PADL + i] = src[lidx(L, i, jl_lo)];
      if (a.has_hi) a.dst_hi[f][IFX_PADL + i] = src[lidx(L, i, jl_hi)];
    }
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    if (a.has_lo) st_release_sys(a.signal_lo + b, a.seq);
    if (a.has_hi) st_release_sys(a.signal_hi + b, a.seq);
    if (b == 0) {      // start of a predictor loop with bodies: nothing of the neighbours' is being read by a ghost-cell kernel
      if (a.has_lo && a.gc_signal_lo) st_release_sys(a.gc_signal_lo, a.seq);
      if (a.has_hi && a.gc_signal_hi) st_release_sys(a.gc_signal_hi, a.seq);
    }
  }
}

static __global__ void k_halo_wait(const unsigned* wait_lo, const unsigned* wait_hi, int ntiles, unsigned need) {
  for (int b = threadIdx.x; b < ntiles; b += blockDim.x) {
    if (wait_lo) wait_seq_ge(wait_lo + b, need);
    if (wait_hi) wait_seq_ge(wait_hi + b, need);
  }
  __threadfence_system();
}

// Lagged stop decision (ReduceCfg.lag): every sweep decides on its predecessor's residual; the last sweep of a batch
// has no successor, so this one-thread kernel collects its partials and decides.  Every rank runs it at the same
// point of the same launch sequence and adds the same operands in the same order: same decision everywhere.
static __global__ void k_lag_flush(LoopCtl* ctl, ReduceCfg rc, HaloCtx hx) {
  if (threadIdx.x != 0 || ctl->done || rc.no_exchange) return;
  double v[4];
  int npairs = 1, eval_first = rc.eval_iter;
  slab_collect_partials(hx, hx.mseq, v, npairs, eval_first);
  decide_on_residuals(ctl, rc, eval_first, npairs, v);
  __threadfence();
}
cudaError_t launch_lag_flush(LoopCtl* ctl, const ReduceCfg& rc, const HaloCtx& hx, cudaStream_t st) {
  k_lag_flush<<<1, 32, 0, st>>>(ctl, rc, hx);
  return cudaGetLastError();
}

cudaError_t launch_halo_push(const HaloPushArgs& a, cudaStream_t st) {
  k_halo_push<<<a.ntiles, 256, 0, st>>>(a);
  return cudaGetLastError();
}

cudaError_t launch_halo_wait(const unsigned* wait_lo, const unsigned* wait_hi, int ntiles, unsigned need, cudaStream_t st) {
  k_halo_wait<<<1, 256, 0, st>>>(wait_lo, wait_hi, ntiles, need);
  return cudaGetLastError();
}

}  // namespace ifx
