// kernels_ppe.cu — pressure-Poisson point-Jacobi sweep with the residual of the INPUT iterate
// evaluated in the same pass.
//
// Reference per sweep (PPESolver.cu:172-188): jacobiIteration (56 B/cell: p + 5 coefficient
// arrays + p_new), Compute_Residual (56 B/cell), reduce6 x2 (8 B/cell), a blocking D2H.
// Here: ONE launch per sweep, 16 B/cell for the Laplace variant the reference ships (24 B/cell
// with a source term): read p once (rows rolled through registers), write p_new.  Sweep m+1
// evaluates the residual of iterate m — the one the reference tests after sweep m — so the
// stop decision is identical; when it fires, iterate m is still intact in the input buffer.
#include "kernels.cuh"
#include "stencil_math.cuh"

namespace ifx {

template <bool LAPLACE_REF, bool WRITE_RES, bool HAS_GC>
static __global__ void __launch_bounds__(AD_THREADS)
k_ppe_sweep(PpeSweepArgs a) {
  if (a.ctl->done && !a.force) return;
  const Layout L = a.L;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int i = 1 + (blockIdx.x * AD_WARPS + warp) * 64 + lane * 2;
  const int jfirst = L.jb + blockIdx.y * a.rows_per_cta;
  const int jlast = min(jfirst + a.rows_per_cta, L.je);
  const int nxm2 = L.nx - 2, nym2 = L.ny - 2;
  const bool act0 = i <= nxm2, act1 = i + 1 <= nxm2;
  // Out-of-range lanes are clamped to a valid address; the lane sitting on the E ghost column must load it for
  // real: its pC.x is the east neighbour (stored Dirichlet ring) of the last interior column next door.
  const int ic = (act0 || (LAPLACE_REF && i == L.nx - 1)) ? i : 1;

  const double cE0 = a.M.pp_cE[ic], cW0 = a.M.pp_cW[ic], sx0 = a.M.pp_sx[ic];
  const double cE1 = a.M.pp_cE[ic + 1], cW1 = a.M.pp_cW[ic + 1], sx1 = a.M.pp_sx[ic + 1];

  // LAPLACE_REF reads the stored ghost ring (Dirichlet scaffolding, PPESolver.cu:54-71);
  // the general variant applies homogeneous Neumann through virtual ghosts.
  const bool ring_w = (ic == 1), ring_e0 = (ic == nxm2), ring_e1 = (ic + 1 == nxm2);
  const bool need_hw = (lane == 0) && (LAPLACE_REF || !ring_w);
  const bool need_he = (lane == 31) && (LAPLACE_REF ? (ic + 2 <= L.nx - 1) : (ic + 2 <= nxm2));

  auto ld2 = [&](const double* p, int jl) -> double2 {
    return *reinterpret_cast<const double2*>(p + lidx(L, ic, jl));
  };

  int jl = jfirst - L.j0;
  double2 pS = make_double2(0, 0), pC, pN = make_double2(0, 0);
  double hw = 0, he = 0, hwN = 0, heN = 0;
  if (LAPLACE_REF || jfirst > 1) pS = ld2(a.pC, jl - 1);
  pC = ld2(a.pC, jl);
  if (need_hw) hw = a.pC[lidx(L, ic - 1, jl)];
  if (need_he) he = a.pC[lidx(L, ic + 2, jl)];

  double rsum = 0.0, rabs = 0.0;

  for (int j = jfirst; j < jlast; ++j, ++jl) {
    const bool top = (j == nym2), bot = (j == 1);
    if (LAPLACE_REF || !top) {
      pN = ld2(a.pC, jl + 1);
      if (j + 1 < jlast) {
        if (need_hw) hwN = a.pC[lidx(L, ic - 1, jl + 1)];
        if (need_he) heN = a.pC[lidx(L, ic + 2, jl + 1)];
      }
    }
    const double cN = a.M.pp_cN[j], cS = a.M.pp_cS[j], sy = a.M.pp_sy[j];
    double pW0 = __shfl_up_sync(0xffffffffu, pC.y, 1);
    double pE1 = __shfl_down_sync(0xffffffffu, pC.x, 1);
    if (lane == 0) pW0 = hw;
    if (lane == 31) pE1 = he;
    double pE0 = pC.y, pW1 = pC.x;
    double2 pSs = pS, pNn = pN;
    if (!LAPLACE_REF) {
      if (ring_w) pW0 = pC.x;
      if (ring_e0) pE0 = pC.x;
      if (ring_e1) pE1 = pC.y;
      if (bot) pSs = pC;
      if (top) pNn = pC;
    }
    const double cP0 = -(sx0 + sy), cP1 = -(sx1 + sy);   // PPESolver.cu:93-94
    const size_t o = lidx(L, ic, jl);

    double2 pn, r;
    if (LAPLACE_REF) {
      const double t0 = ppe_offdiag(pW0, cW0, pE0, cE0, pNn.x, cN, pSs.x, cS);
      const double t1 = ppe_offdiag(pW1, cW1, pE1, cE1, pNn.y, cN, pSs.y, cS);
      pn.x = (-t0) / cP0; pn.y = (-t1) / cP1;                          // PPESolver.cu:24-27
      const double q0 = ppe_apply(pC.x, cP0, pW0, cW0, pE0, cE0, pNn.x, cN, pSs.x, cS);   // :42-46
      const double q1 = ppe_apply(pC.y, cP1, pW1, cW1, pE1, cE1, pNn.y, cN, pSs.y, cS);
      r.x = act0 ? q0 : 0.0; r.y = act1 ? q1 : 0.0;
      if (act1) *reinterpret_cast<double2*>(a.pT + o) = pn;
      else if (act0) a.pT[o] = pn.x;
    } else {
      const double2 f = *reinterpret_cast<const double2*>(a.rhs + o);
      const uchar2 ct = *reinterpret_cast<const uchar2*>(a.celltype + o);
      const double t0 = ppe_offdiag(pW0, cW0, pE0, cE0, pNn.x, cN, pSs.x, cS);
      const double t1 = ppe_offdiag(pW1, cW1, pE1, cE1, pNn.y, cN, pSs.y, cS);
      const double q0 = ppe_apply(pC.x, cP0, pW0, cW0, pE0, cE0, pNn.x, cN, pSs.x, cS);
      const double q1 = ppe_apply(pC.y, cP1, pW1, cW1, pE1, cE1, pNn.y, cN, pSs.y, cS);
      const bool fl0 = ct.x == IFX_FLUID, fl1 = ct.y == IFX_FLUID;
      pn.x = fl0 ? (f.x - t0) / cP0 : pC.x;      // solid cells keep their value in both buffers
      pn.y = fl1 ? (f.y - t1) / cP1 : pC.y;
      r.x = (act0 && fl0) ? f.x - q0 : 0.0;
      r.y = (act1 && fl1) ? f.y - q1 : 0.0;
      if (HAS_GC) {
        if (act0 && ct.x != IFX_GHOST) a.pT[o] = pn.x;
        if (act1 && ct.y != IFX_GHOST) a.pT[o + 1] = pn.y;
      } else if (act1) *reinterpret_cast<double2*>(a.pT + o) = pn;
      else if (act0) a.pT[o] = pn.x;
    }
    rsum += r.x; rsum += r.y;
    rabs += fabs(r.x); rabs += fabs(r.y);
    if (WRITE_RES) {
      const size_t ro = (size_t)j * L.nx + ic;
      if (act0) a.res[ro] = r.x;
      if (act1) a.res[ro + 1] = r.y;
    }
    pS = pC; pC = pN; hw = hwN; he = heN;
  }
  block_reduce_and_decide<AD_THREADS>(rsum, rabs, a.partials, a.ctl, a.rc,
                                      blockIdx.y * gridDim.x + blockIdx.x, gridDim.x * gridDim.y);
}

cudaError_t launch_ppe_sweep(const PpeSweepArgs& a, dim3 grid, cudaStream_t st, bool laplace_ref, bool write_res,
                             bool has_gc) {
  if (laplace_ref) {
    if (write_res) k_ppe_sweep<true, true, false><<<grid, AD_THREADS, 0, st>>>(a);
    else k_ppe_sweep<true, false, false><<<grid, AD_THREADS, 0, st>>>(a);
  } else if (has_gc) {
    if (write_res) k_ppe_sweep<false, true, true><<<grid, AD_THREADS, 0, st>>>(a);
    else k_ppe_sweep<false, false, true><<<grid, AD_THREADS, 0, st>>>(a);
  } else {
    if (write_res) k_ppe_sweep<false, true, false><<<grid, AD_THREADS, 0, st>>>(a);
    else k_ppe_sweep<false, false, false><<<grid, AD_THREADS, 0, st>>>(a);
  }
  return cudaGetLastError();
}

// set_pressure_BC (PPESolver.cu:54-71): p = 100 where i == 0 or j == 0, applied to both ping-pong
// buffers (the reference copies boundary cells through every sweep, PPESolver.cu:21).
static __global__ void k_set_pressure_bc_ref(Layout L, double* p0, double* p1) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < L.nyl) { p0[lidx(L, 0, t)] = 100.0; if (p1) p1[lidx(L, 0, t)] = 100.0; }
  if (t < L.nx && L.j0 == 0) { p0[lidx(L, t, 0)] = 100.0; if (p1) p1[lidx(L, t, 0)] = 100.0; }
}

cudaError_t launch_set_pressure_bc_ref(const Layout& L, double* p0, double* p1, cudaStream_t st) {
  const int n1 = L.nyl > L.nx ? L.nyl : L.nx;
  k_set_pressure_bc_ref<<<(n1 + 127) / 128, 128, 0, st>>>(L, p0, p1);
  return cudaGetLastError();
}

}  // namespace ifx
