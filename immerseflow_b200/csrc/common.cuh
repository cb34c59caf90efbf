// common.cuh — shared device/host definitions for the B200-native ImmerseFlow++ path.
//
// HBM layout.  Every cell-centred field of a slab is one padded 2-D array:
//     element (i, jl)  ->  base[ jl*pitch + IFX_PADL + i ],   i = 0..nx-1 (ghost-inclusive),
//     jl = local row, global row j = j0 + jl.
// IFX_PADL = 15 puts the first interior column (i = 1) on a 128-byte boundary and pitch is a
// multiple of 16 doubles, so a warp reading 64 consecutive interior columns as double2 touches
// exactly four full 128-B lines.  The reference's unpadded id = i + j*nx layout
// (globalVariables.cuh) only exists at the C-ABI boundary (ifx_set_field / ifx_get_field).
//
// Arithmetic: this directory is compiled with -fmad=false.  Every fused multiply-add is written
// out as fma() and mirrors the contraction nvcc emits for the reference kernels (read from the
// reference's sm_100 PTX/SASS), so results are bit-identical to the reference CUDA build.
#pragma once
// IFX_HOST_SHIM: tests/cuda_host_shim.h compiles the thread-per-cell kernels (kernels_mg.cu, kernels_diag.cu) with
// g++ and runs their grids serially on the CPU, so their arithmetic and indexing are checked against the oracle
// in the CPU test suite as well as on the GPU.  Test infrastructure only; the product is never built this way.
#ifndef IFX_HOST_SHIM
#include <cuda_runtime.h>
#endif
#include <stdint.h>

#define IFX_PADL 15
#define IFX_WARP 32

// cell types (uint8).  The reference's iBlank is a double array, 1.0 fluid / 0.0 solid
// (globalVariables.cuh:50-52); ghost cells are solid cells adjacent to fluid.
#define IFX_SOLID 0
#define IFX_FLUID 1
#define IFX_GHOST 2

namespace ifx {

struct Layout {
  int nx, ny;      // global ghost-inclusive sizes
  int pitch;       // doubles per stored row
  int nyl;         // stored local rows (owned interior rows + one halo/ghost row each side)
  int j0;          // global row index of local row 0
  int jb, je;      // owned interior rows, global, [jb, je)
};

__host__ __device__ __forceinline__ size_t lidx(const Layout& L, int i, int jl) {
  return (size_t)jl * L.pitch + IFX_PADL + i;
}

// 1-D metric / coefficient tables (device pointers; x tables have nx entries, y tables ny entries,
// indexed by GLOBAL i / j).  All values are computed on the device by k_build_metrics with the
// reference's per-cell expressions, which only depend on i (x tables) or j (y tables).
struct Metrics {
  const double* dx;     // dx[i] = xf[i]-xf[i-1]; dx[0]=dx[1], dx[nx-1]=dx[nx-2]   (preSim.cu:329-355)
  const double* dy;
  const double* xc;     // cell centres incl. ghost centres                       (preSim.cu:294-307)
  const double* yc;
  const double* rcpx;   // rcpx[i] = 1/(dx[i]+dx[i+1]), i = 0..nx-2   (ADSolver.cu:66,173)
  const double* rcpy;
  const double* kxh;    // (dt/dx[i])*0.5                               (ADSolver.cu:66)
  const double* kyh;
  // predictor (ADSolver.cu:34-39), k = dt/Re
  const double* ad_cE;  // k * 2/(dx_i (dx_i+dx_ip1))
  const double* ad_cW;  // k * 2/(dx_i (dx_i+dx_im1))
  const double* ad_px;  // fma(k, ax_p+ax_m, 1.0)
  const double* ad_cN;
  const double* ad_cS;
  const double* ad_sy;  // ay_p+ay_m      -> cP = fma(k, ad_sy[j], ad_px[i])
  // Poisson (PPESolver.cu:93-99)
  const double* pp_cE;  // 2/(dx_i (dx_i+dx_ip1))
  const double* pp_cW;
  const double* pp_sx;  // ax_p+ax_m
  const double* pp_cN;
  const double* pp_cS;
  const double* pp_sy;  // ay_p+ay_m      -> cP = -(pp_sx[i] + pp_sy[j])
  double k;             // dt/Re
  double dt;
};

// Device-resident loop control: lets the host enqueue sweeps blindly while the stop decision is
// taken on the device by the last CTA of each sweep (no host round trip per iteration, unlike the
// reference's blocking 8-byte D2H per Reduction(), preSim.cu:441).
struct LoopCtl {
  int iter;            // iterations whose residual has been evaluated
  int done;            // 1: later launches exit immediately
  int ambiguous;       // fused sum landed inside the rounding band of the tolerance
  int hit_max;
  unsigned ticket;     // CTA completion counter
  int pad_;
  double res0, res1;   // last evaluated residual sums (u,v) or (signed, abs) for PPE
  double hist[2 * 64]; // first 64 iterations' residuals
};

// ---------------------------------------------------------------------------------------------------------
// Slab decomposition (one process per GPU).  Every rank keeps its exchangeable fields and a small
// synchronisation area in ONE allocation that its neighbours map through CUDA IPC; halo rows are
// delivered by plain stores to the mapped peer pointer from inside the sweep kernels (NVLink P2P), followed
// by a release-store of a sequence number the consumer acquires before it reads the row.
// ---------------------------------------------------------------------------------------------------------
#define IFX_MAX_RANKS 8
#define IFX_MAX_TILES 1024
#define IFX_SYNC_GROUPS 3          // 0: predictor sweeps (u, v), 1: Poisson sweeps (p), 2: everything else
#define IFX_MAIL_SLOTS 8
#define IFX_MAIL_VALS 6            // doubles per mailbox entry: two residual pairs, their count, the first iterate index
#define IFX_SEG_FIELDS_MAX 8       // fields an exchange segment may hold: u[2], v[2], p[3] (+1 spare)
#define IFX_GC_REACH 4             // rows beyond its slab a ghost-cell stencil may read from the neighbour's memory

struct XchgSync {
  unsigned flags[IFX_SYNC_GROUPS][2][IFX_MAX_TILES];      // [group][0: written by lower nbr, 1: by upper][tile]
  unsigned gcflag[2];                                     // predictor with bodies: the neighbour's ghost-cell kernel of
                                                          // iteration seq is through ([0]: lower neighbour, [1]: upper)
  // per-rank residual partials of one launch, as self-validating 8-byte packets {32 data bits, 32-bit launch tag}: two
  // packets per double, no fence and no separate flag (the layout of NCCL's LL protocol)
  unsigned long long mail[IFX_MAIL_SLOTS][IFX_MAX_RANKS][2 * IFX_MAIL_VALS];
};

struct HaloCtx {
  int nranks, rank;
  int has_lo, has_hi;              // a neighbour slab below / above
  unsigned seq;                    // sequence number of this launch within its sync group
  int defer;                       // 1: predictor with immersed bodies on slabs.  The ghost cells of an iterate are closed by
                                   // a kernel of their own (kernels_ib.cu) that reads the neighbours' previous iterate over
                                   // NVLink, overwrites the ghost cells in the rows this launch delivered and then publishes
                                   // gcflag: halo rows are complete, and rows near a slab boundary may be overwritten, only
                                   // after the neighbours' gcflag of the previous iteration
  const unsigned* gcw_lo;          // my gcflag[0] / gcflag[1] (written by the lower / upper neighbour)
  const unsigned* gcw_hi;
  unsigned mseq;                   // global sweep counter (mailbox slot / tag)
  const unsigned* wait_lo;         // my flags written by the lower neighbour, per tile
  const unsigned* wait_hi;
  unsigned* signal_lo;             // the lower neighbour's "written by upper" flags
  unsigned* signal_hi;
  double* peer_row_lo[2];          // lower neighbour's top halo row of the OUTPUT buffer (per field), row start
  double* peer_row_hi[2];          // upper neighbour's bottom halo row
  unsigned long long* mail[IFX_MAX_RANKS];   // every rank's mail[][][] (mine included)
};

#ifndef IFX_HOST_SHIM
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_relaxed_sys(const unsigned* p) {
  unsigned v;
  asm volatile("ld.relaxed.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_relaxed_sys(unsigned* p, unsigned v) {
  asm volatile("st.relaxed.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_relaxed_sys_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_relaxed_sys_u64(unsigned long long* p, unsigned long long v) {
  asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ void wait_seq_ge(const unsigned* p, unsigned need) {
  // sequence numbers wrap after 2^32 launches; compare as a signed distance
  while ((int)(ld_acquire_sys(p) - need) < 0) { __nanosleep(64); }
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) v += __shfl_down_sync(0xffffffffu, v, off);
  return v;
}
#endif  // IFX_HOST_SHIM

}  // namespace ifx
