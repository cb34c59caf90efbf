// kernels_reduce.cu — sum reduction in the REFERENCE's summation order.
//
// The reference sums residual arrays with Harris' reduce6<256> in two passes
// (preSim.cu:12-50 driven by ImmerseFlow::Reduction, :376-441): level 1 runs B = ceil(N/256)
// blocks, block b folding elements [512b, 512b+512) as (g[i] + g[i+256]) followed by the tree
// +128, +64, +32, +16, +8, +4, +2, +1; level 2 is one block over the B partials, each thread
// first accumulating (P[i] + P[i+256]) for i = t, t+512, ... sequentially.  No atomics, so the
// order — and therefore every bit of the result — is fixed.  These kernels reproduce exactly that
// pairing (with __syncwarp-correct shuffles instead of the reference's implicit warp-synchronous
// volatile tail) so residuals and iteration counts can be compared bit for bit.
#include "kernels.cuh"

namespace ifx {

// The reference tree for 256 values held one per thread: s[t] += s[t+128]; += s[t+64]; then the warp
// tail += s[t+32], 16, 8, 4, 2, 1.  Returns the total in thread 0.
__device__ __forceinline__ double tree256(double s, double* sh) {
  const unsigned t = threadIdx.x;
  sh[t] = s;
  __syncthreads();
  if (t < 128) sh[t] = s = s + sh[t + 128];
  __syncthreads();
  if (t < 64) sh[t] = s = s + sh[t + 64];
  __syncthreads();
  if (t < 32) {
    s = s + sh[t + 32];
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) s = s + __shfl_down_sync(0xffffffffu, s, off);
  }
  return s;
}

template <bool ABS>
static __global__ void __launch_bounds__(256) k_reduce6_level1(const double* __restrict__ in,
                                                        double* __restrict__ partial, unsigned n) {
  __shared__ double sh[256];
  const unsigned t = threadIdx.x;
  // gridSize = 512*B >= 2N, so the reference's while loop (preSim.cu:30-40) runs at most once
  const unsigned i = blockIdx.x * 512u + t;
  double s = 0.0;
  if (i < n) {
    double v;
    if (ABS) v = (i + 256u < n) ? fabs(in[i]) + fabs(in[i + 256u]) : fabs(in[i]);   // sum of |r| in the same order
    else v = (i + 256u < n) ? in[i] + in[i + 256u] : in[i];
    s = s + v;                                       // sdata[tid] (= 0) += ...
  }
  s = tree256(s, sh);
  if (t == 0) partial[blockIdx.x] = s;
}

static __global__ void __launch_bounds__(256) k_reduce6_level2(const double* __restrict__ partial, unsigned nb,
                                                        double* __restrict__ out) {
  __shared__ double sh[256];
  const unsigned t = threadIdx.x;
  double s = 0.0;
  for (unsigned i = t; i < nb; i += 512u) {          // one block: gridSize = 512
    double v = (i + 256u < nb) ? partial[i] + partial[i + 256u] : partial[i];
    s = s + v;
  }
  s = tree256(s, sh);
  if (t == 0) *out = s;
}

// Stop decision from reference-order sums (sums[0], sums[1]); mirrors the tail of
// block_reduce_and_decide without the certification band (the sums ARE the reference's).
static __global__ void k_decide_exact(LoopCtl* ctl, const double* sums, ReduceCfg rc) {
  if (ctl->done && !ctl->ambiguous) return;
  const double a = sums[0], b = sums[1];
  ctl->res0 = a; ctl->res1 = b;
  if (rc.eval_iter >= 1) {
    if (rc.eval_iter <= 64) { ctl->hist[2 * (rc.eval_iter - 1)] = a; ctl->hist[2 * (rc.eval_iter - 1) + 1] = b; }
    ctl->iter = rc.eval_iter;
    const double S = rc.use_second ? a + b : (rc.test_abs ? b : a);
    const bool at_max = rc.eval_iter >= rc.itermax;
    ctl->ambiguous = 0;
    if (!(S > rc.tol) || at_max) { ctl->hit_max = (S > rc.tol) ? 1 : 0; ctl->done = 1; }
    else ctl->done = 0;
  }
}

// ImmerseFlow::Reduction (preSim.cu:376-445): B = ceil(n/256) blocks (preSim.cu:193), then one block.
cudaError_t launch_reduce6(const double* in, size_t n, double* partial, double* out, cudaStream_t st, bool abs_values) {
  const unsigned B = (unsigned)((n + 255) / 256);
  if (abs_values) k_reduce6_level1<true><<<B, 256, 0, st>>>(in, partial, (unsigned)n);
  else k_reduce6_level1<false><<<B, 256, 0, st>>>(in, partial, (unsigned)n);
  k_reduce6_level2<<<1, 256, 0, st>>>(partial, B, out);
  return cudaGetLastError();
}
cudaError_t launch_decide_exact(LoopCtl* ctl, const double* sums, const ReduceCfg& rc, cudaStream_t st) {
  k_decide_exact<<<1, 1, 0, st>>>(ctl, sums, rc);
  return cudaGetLastError();
}

}  // namespace ifx
