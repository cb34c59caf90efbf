// kernels_ad.cu — momentum predictor (advection source + implicit-diffusion Jacobi sweeps).
//
// Replaces the reference's per-iteration kernel train (ADSolver.cu:317-366):
//   set_velocity_BC, Compute_velf, ADusolver_kernel, ADvsolver_kernel, set_velocity_BC,
//   Compute_uResidual_AD, Compute_vResidual_AD, 2 x (reduce6, reduce6)      -> 11 launches, 248 B/cell
// by ONE launch per iteration (k_ad_jacobi) moving 48 B/cell (+1 B cell type):
//   read u,v (5-point, rows rolled through registers), sx, sy; write u', v'; the ghost ring of the
//   input buffer is produced on the fly ("virtual ghosts") and stored exactly as the reference's
//   BC kernel would have left it; |u'-u|, |v'-v| are summed in the same pass.
// Coefficients are separable and time-invariant: 1-D tables instead of the
// reference's five N-sized arrays rebuilt every step (ADSolver.cu:275-291).
#include "kernels.cuh"
#include "stencil_math.cuh"

namespace ifx {

// ---------------------------------------------------------------------------------------------
// k_ad_jacobi: one point-Jacobi iteration of (I - dt/Re Lap) q = s for q = u and q = v.
// Tile = (64*AD_WARPS) columns x rows_per_cta rows; lane -> 2 adjacent columns (double2, 16-B
// aligned thanks to IFX_PADL), rows marched with south/centre/north kept in registers; W/E
// neighbours by warp shuffle, warp-edge columns by one scalar load per edge lane.
// ---------------------------------------------------------------------------------------------
template <bool WRITE_RES, bool HAS_GC>
static __global__ void __launch_bounds__(AD_THREADS)
k_ad_jacobi(AdJacobiArgs a) {
  if (a.ctl->done && !a.force) return;
  const Layout L = a.L;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int i = 1 + (blockIdx.x * AD_WARPS + warp) * 64 + lane * 2;   // first of my two columns
  const int jfirst = L.jb + blockIdx.y * a.rows_per_cta;
  const int jlast = min(jfirst + a.rows_per_cta, L.je);                // exclusive
  const int nxm2 = L.nx - 2, nym2 = L.ny - 2;
  const bool act0 = i <= nxm2, act1 = i + 1 <= nxm2;
  // clamp so that out-of-range lanes read valid memory (their results are discarded)
  const int ic = act0 ? i : 1;

  double resu = 0.0, resv = 0.0;

  // column coefficients (ADSolver.cu:36-37 and the x half of :34)
  const double cE0 = a.M.ad_cE[ic], cW0 = a.M.ad_cW[ic], px0 = a.M.ad_px[ic];
  const double cE1 = a.M.ad_cE[ic + 1], cW1 = a.M.ad_cW[ic + 1], px1 = a.M.ad_px[ic + 1];
  const double k = a.M.k;

  const bool need_hw = (lane == 0) && (ic > 1);            // west halo is a stored cell
  const bool need_he = (lane == 31) && (ic + 2 <= nxm2);   // east halo is a stored interior cell

  auto ld2 = [&](const double* p, int jl) -> double2 {
    return *reinterpret_cast<const double2*>(p + lidx(L, ic, jl));
  };

  // prime the rolling registers: south = row jfirst-1, centre = row jfirst
  int jl = jfirst - L.j0;
  double2 uS, vS, uC, vC, uN, vN;
  double uhw = 0, uhe = 0, vhw = 0, vhe = 0, uhwN = 0, uheN = 0, vhwN = 0, vheN = 0;
  if (jfirst > 1) {   // row below is a stored row (interior or slab halo), not the S ghost
    uS = ld2(a.uC, jl - 1); vS = ld2(a.vC, jl - 1);
  } else { uS = make_double2(0, 0); vS = uS; }
  uC = ld2(a.uC, jl); vC = ld2(a.vC, jl);
  if (need_hw) { uhw = a.uC[lidx(L, ic - 1, jl)]; vhw = a.vC[lidx(L, ic - 1, jl)]; }
  if (need_he) { uhe = a.uC[lidx(L, ic + 2, jl)]; vhe = a.vC[lidx(L, ic + 2, jl)]; }

  for (int j = jfirst; j < jlast; ++j, ++jl) {
    const bool top = (j == nym2), bot = (j == 1);
    // next row (north) — a stored row unless j is the last interior row of the global grid
    if (!top) {
      uN = ld2(a.uC, jl + 1); vN = ld2(a.vC, jl + 1);
      if (j + 1 < jlast) {
        if (need_hw) { uhwN = a.uC[lidx(L, ic - 1, jl + 1)]; vhwN = a.vC[lidx(L, ic - 1, jl + 1)]; }
        if (need_he) { uheN = a.uC[lidx(L, ic + 2, jl + 1)]; vheN = a.vC[lidx(L, ic + 2, jl + 1)]; }
      }
    }
    const double2 s_x = *reinterpret_cast<const double2*>(a.sx + lidx(L, ic, jl));
    const double2 s_y = *reinterpret_cast<const double2*>(a.sy + lidx(L, ic, jl));
    const uchar2 ct = *reinterpret_cast<const uchar2*>(a.celltype + lidx(L, ic, jl));
    const double cN = a.M.ad_cN[j], cS = a.M.ad_cS[j], sy = a.M.ad_sy[j];

    // W/E neighbours: shuffle inside the warp, halo loads at the warp edges
    double uW0 = __shfl_up_sync(0xffffffffu, uC.y, 1), vW0 = __shfl_up_sync(0xffffffffu, vC.y, 1);
    double uE1 = __shfl_down_sync(0xffffffffu, uC.x, 1), vE1 = __shfl_down_sync(0xffffffffu, vC.x, 1);
    if (lane == 0) { uW0 = uhw; vW0 = vhw; }
    if (lane == 31) { uE1 = uhe; vE1 = vhe; }
    // virtual ghosts (set_velocity_BC, ADSolver.cu:199-217): ghost = -interior + 2*bc
    const double ugw = a.two_bc_u[0] - uC.x, vgw = a.two_bc_v[0] - vC.x;   // used when ic == 1
    if (ic == 1) { uW0 = ugw; vW0 = vgw; }
    double uE0 = uC.y, vE0 = vC.y, uW1 = uC.x, vW1 = vC.x;
    const bool e0 = (ic == nxm2), e1 = (ic + 1 == nxm2);
    if (e0) { uE0 = a.two_bc_u[1] - uC.x; vE0 = a.two_bc_v[1] - vC.x; }
    if (e1) { uE1 = a.two_bc_u[1] - uC.y; vE1 = a.two_bc_v[1] - vC.y; }
    double2 uSs = uS, vSs = vS, uNn = uN, vNn = vN;
    if (bot) {
      uSs = make_double2(a.two_bc_u[2] - uC.x, a.two_bc_u[2] - uC.y);
      vSs = make_double2(a.two_bc_v[2] - vC.x, a.two_bc_v[2] - vC.y);
    }
    if (top) {
      uNn = make_double2(a.two_bc_u[3] - uC.x, a.two_bc_u[3] - uC.y);
      vNn = make_double2(a.two_bc_v[3] - vC.x, a.two_bc_v[3] - vC.y);
    }

    const double cP0 = fma(k, sy, px0), cP1 = fma(k, sy, px1);
    const double ib0 = (ct.x == IFX_FLUID) ? 1.0 : 0.0, ib1 = (ct.y == IFX_FLUID) ? 1.0 : 0.0;
    double2 un, vn;
    un.x = jac_cell(s_x.x, cE0, uE0, cW0, uW0, cN, uNn.x, cS, uSs.x, ib0, cP0);
    un.y = jac_cell(s_x.y, cE1, uE1, cW1, uW1, cN, uNn.y, cS, uSs.y, ib1, cP1);
    vn.x = jac_cell(s_y.x, cE0, vE0, cW0, vW0, cN, vNn.x, cS, vSs.x, ib0, cP0);
    vn.y = jac_cell(s_y.y, cE1, vE1, cW1, vW1, cN, vNn.y, cS, vSs.y, ib1, cP1);

    const size_t o = lidx(L, ic, jl);
    if (HAS_GC) {
      // ghost cells are owned by the ghost-cell CTAs of this launch; leave them untouched here
      if (act0 && ct.x != IFX_GHOST) a.uT[o] = un.x, a.vT[o] = vn.x;
      if (act1 && ct.y != IFX_GHOST) a.uT[o + 1] = un.y, a.vT[o + 1] = vn.y;
    } else if (act1) {
      *reinterpret_cast<double2*>(a.uT + o) = un;
      *reinterpret_cast<double2*>(a.vT + o) = vn;
    } else if (act0) {
      a.uT[o] = un.x; a.vT[o] = vn.x;
    }

    // Compute_{u,v}Residual_AD (ADSolver.cu:131-137): |new - old| where iBlank == 1
    double ru0 = 0, ru1 = 0, rv0 = 0, rv1 = 0;
    if (act0 && ct.x == IFX_FLUID) { ru0 = fabs(uC.x - un.x); rv0 = fabs(vC.x - vn.x); }
    if (act1 && ct.y == IFX_FLUID) { ru1 = fabs(uC.y - un.y); rv1 = fabs(vC.y - vn.y); }
    resu += ru0; resu += ru1; resv += rv0; resv += rv1;
    if (WRITE_RES) {   // reference layout, for the reference-order reduction
      const size_t r = (size_t)j * L.nx + ic;
      if (act0) { a.res_u[r] = ru0; a.res_v[r] = rv0; }
      if (act1) { a.res_u[r + 1] = ru1; a.res_v[r + 1] = rv1; }
    }

    // ghost ring of the INPUT buffer, exactly what set_velocity_BC leaves there (nobody reads it in
    // this launch: all ghost reads above are virtual).  Corners: 2bc - (2bc - diagonal neighbour).
    if (act0) {
      if (ic == 1) { a.uC[lidx(L, 0, jl)] = ugw; a.vC[lidx(L, 0, jl)] = vgw; }
      if (e0) { a.uC[lidx(L, L.nx - 1, jl)] = uE0; a.vC[lidx(L, L.nx - 1, jl)] = vE0; }
      if (bot) {
        a.uC[lidx(L, ic, jl - 1)] = uSs.x; a.vC[lidx(L, ic, jl - 1)] = vSs.x;
        if (ic == 1) {
          a.uC[lidx(L, 0, jl - 1)] = a.two_bc_u[2] - ugw; a.vC[lidx(L, 0, jl - 1)] = a.two_bc_v[2] - vgw;
        }
        if (e0) {
          a.uC[lidx(L, L.nx - 1, jl - 1)] = a.two_bc_u[1] - uSs.x;
          a.vC[lidx(L, L.nx - 1, jl - 1)] = a.two_bc_v[1] - vSs.x;
        }
      }
      if (top) {
        a.uC[lidx(L, ic, jl + 1)] = uNn.x; a.vC[lidx(L, ic, jl + 1)] = vNn.x;
        if (ic == 1) {
          a.uC[lidx(L, 0, jl + 1)] = a.two_bc_u[3] - ugw; a.vC[lidx(L, 0, jl + 1)] = a.two_bc_v[3] - vgw;
        }
        if (e0) {
          a.uC[lidx(L, L.nx - 1, jl + 1)] = a.two_bc_u[3] - uE0;
          a.vC[lidx(L, L.nx - 1, jl + 1)] = a.two_bc_v[3] - vE0;
        }
      }
    }
    if (act1) {
      if (e1) { a.uC[lidx(L, L.nx - 1, jl)] = uE1; a.vC[lidx(L, L.nx - 1, jl)] = vE1; }
      if (bot) {
        a.uC[lidx(L, ic + 1, jl - 1)] = uSs.y; a.vC[lidx(L, ic + 1, jl - 1)] = vSs.y;
        if (e1) {
          a.uC[lidx(L, L.nx - 1, jl - 1)] = a.two_bc_u[1] - uSs.y;
          a.vC[lidx(L, L.nx - 1, jl - 1)] = a.two_bc_v[1] - vSs.y;
        }
      }
      if (top) {
        a.uC[lidx(L, ic + 1, jl + 1)] = uNn.y; a.vC[lidx(L, ic + 1, jl + 1)] = vNn.y;
        if (e1) {
          a.uC[lidx(L, L.nx - 1, jl + 1)] = a.two_bc_u[3] - uE1;
          a.vC[lidx(L, L.nx - 1, jl + 1)] = a.two_bc_v[3] - vE1;
        }
      }
    }

    // roll
    uS = uC; vS = vC; uC = uN; vC = vN;
    uhw = uhwN; uhe = uheN; vhw = vhwN; vhe = vheN;
  }

  block_reduce_and_decide<AD_THREADS>(resu, resv, a.partials, a.ctl, a.rc,
                                      blockIdx.y * gridDim.x + blockIdx.x, gridDim.x * gridDim.y);
}

cudaError_t launch_ad_jacobi(const AdJacobiArgs& a, dim3 grid, cudaStream_t st, bool write_res, bool has_gc) {
  if (write_res) {
    if (has_gc) k_ad_jacobi<true, true><<<grid, AD_THREADS, 0, st>>>(a);
    else k_ad_jacobi<true, false><<<grid, AD_THREADS, 0, st>>>(a);
  } else {
    if (has_gc) k_ad_jacobi<false, true><<<grid, AD_THREADS, 0, st>>>(a);
    else k_ad_jacobi<false, false><<<grid, AD_THREADS, 0, st>>>(a);
  }
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// k_ad_source: explicit convection right-hand side (ADSource, ADSolver.cu:46-79) fused with the
// face-velocity interpolation (Compute_velf, :162-189) and the BC refresh (:304).
//
// UF_INLINE (reference compat): face velocities are recomputed from u as the reference's velf
// kernel sees it — i.e. from the buffer BEFORE the BC refresh, so boundary faces use the stale
// ghost values stored in memory — while the cell-centre interpolation of the
// same expression uses the refreshed ("virtual") ghosts.  VF_ZERO: vf == +0.0 everywhere, which
// is what the reference's bug leaves for nx <= ny (App. A Q2); the arithmetic is carried out
// with the zeros so the bits (including signed zeros) match.
// !UF_INLINE: uf/vf come from stored (projected) face arrays — IFX_COMPAT_FULL.
// ---------------------------------------------------------------------------------------------
template <bool UF_INLINE, bool VF_ZERO>
static __global__ void __launch_bounds__(AD_THREADS)
k_ad_source(AdSourceArgs a) {
  const Layout L = a.L;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // a warp covers 64 columns as two coalesced 32-column halves; rows are the OUTER loop so the rows a warp just
  // read as "north" are still in L1 when they become "centre" and "south" (the first cut walked each column pair
  // down the whole tile with 16-byte lane stride: ncu showed 3x the algorithmic DRAM traffic)
  const int i0 = 1 + (blockIdx.x * AD_WARPS + warp) * 64 + lane;
  const int jfirst = L.jb + blockIdx.y * a.rows_per_cta;
  const int jlast = min(jfirst + a.rows_per_cta, L.je);
  const int nxm2 = L.nx - 2, nym2 = L.ny - 2;

  for (int j = jfirst; j < jlast; ++j) {
    const int jl = j - L.j0;
    const double dy_j = a.M.dy[j], dy_jp1 = a.M.dy[j + 1], dy_jm1 = a.M.dy[j - 1];
    const double rcp_n = a.M.rcpy[j], rcp_s = a.M.rcpy[j - 1];
    const double ky = a.M.kyh[j];
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int i = i0 + 32 * e;
      if (i > nxm2) continue;
      const double dx_i = a.M.dx[i], dx_ip1 = a.M.dx[i + 1], dx_im1 = a.M.dx[i - 1];
      const double rcp_e = a.M.rcpx[i], rcp_w = a.M.rcpx[i - 1];
      const double kx = a.M.kxh[i];
      const size_t o = lidx(L, i, jl);
      // stored values (ghost cells as left by the previous step)
      const double uc = a.u[o], vc = a.v[o];
      const double uE_m = a.u[o + 1], uW_m = a.u[o - 1], uN_m = a.u[o + L.pitch], uS_m = a.u[o - L.pitch];
      const double vE_m = a.v[o + 1], vW_m = a.v[o - 1], vN_m = a.v[o + L.pitch], vS_m = a.v[o - L.pitch];
      // refreshed ghosts (set_velocity_BC at ADSolver.cu:304 runs between velf and ADSource)
      const double uE = (i == nxm2) ? a.two_bc_u[1] - uc : uE_m, vE = (i == nxm2) ? a.two_bc_v[1] - vc : vE_m;
      const double uW = (i == 1) ? a.two_bc_u[0] - uc : uW_m, vW = (i == 1) ? a.two_bc_v[0] - vc : vW_m;
      const double uN = (j == nym2) ? a.two_bc_u[3] - uc : uN_m, vN = (j == nym2) ? a.two_bc_v[3] - vc : vN_m;
      const double uS = (j == 1) ? a.two_bc_u[2] - uc : uS_m, vS = (j == 1) ? a.two_bc_v[2] - vc : vS_m;

      double uf_e, uf_w, vf_n, vf_s;
      if (UF_INLINE) {
        // Compute_velf (ADSolver.cu:173-174): uf = rcp(dxL+dxR) * fma(uR, dxL, uL*dxR), stale ghosts
        uf_e = rcp_e * fma(uE_m, dx_i, uc * dx_ip1);
        uf_w = rcp_w * fma(uc, dx_im1, uW_m * dx_i);
      } else {
        uf_e = a.uf[o]; uf_w = a.uf[o - 1];
      }
      if (VF_ZERO) { vf_n = 0.0; vf_s = 0.0; }
      else { vf_n = a.vf[o]; vf_s = a.vf[o - L.pitch]; }

      const double ue = fma(dx_i, uE, dx_ip1 * uc) * rcp_e;
      const double uw = fma(dx_im1, uc, dx_i * uW) * rcp_w;
      const double un = fma(dy_j, uN, dy_jp1 * uc) * rcp_n;
      const double us = fma(dy_jm1, uc, dy_j * uS) * rcp_s;
      const double dfx = fma(uw, -uf_w, ue * uf_e);
      const double dfy = fma(us, -vf_s, un * vf_n);
      a.sx[o] = fma(-ky, dfy, fma(-kx, dfx, uc));

      const double ve = fma(dx_ip1, vc, dx_i * vE) * rcp_e;
      const double vw = fma(dx_im1, vc, dx_i * vW) * rcp_w;
      const double vn = fma(dy_j, vN, dy_jp1 * vc) * rcp_n;
      const double vs = fma(dy_jm1, vc, dy_j * vS) * rcp_s;
      const double gfx = fma(ve, uf_e, -(vw * uf_w));
      const double gfy = fma(vn, vf_n, -(vs * vf_s));
      a.sy[o] = fma(-ky, gfy, fma(-kx, gfx, vc));
    }
  }
}

cudaError_t launch_ad_source(const AdSourceArgs& a, dim3 grid, cudaStream_t st, AdSourceVariant v) {
  switch (v) {
    case SRC_REF_VF_ZERO: k_ad_source<true, true><<<grid, AD_THREADS, 0, st>>>(a); break;
    case SRC_REF_VF_ARRAY: k_ad_source<true, false><<<grid, AD_THREADS, 0, st>>>(a); break;
    case SRC_FACES: k_ad_source<false, false><<<grid, AD_THREADS, 0, st>>>(a); break;
  }
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// k_copy_ring: copy the ghost ring (and nothing else) from one buffer pair to another.
// Used for the K == 1 corner case of the predictor (see oracle orc_ADsolver) and to seed
// ping-pong partners.
// ---------------------------------------------------------------------------------------------
static __global__ void k_copy_ring(Layout L, const double* __restrict__ s0, double* __restrict__ d0,
                            const double* __restrict__ s1, double* __restrict__ d1) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int nrows = L.nyl;
  // columns 0 and nx-1 of every stored row
  if (t < nrows) {
    const size_t a = lidx(L, 0, t), b = lidx(L, L.nx - 1, t);
    d0[a] = s0[a]; d0[b] = s0[b];
    if (s1) { d1[a] = s1[a]; d1[b] = s1[b]; }
  }
  // rows 0 and ny-1 (only where this slab holds them)
  if (t < L.nx) {
    if (L.j0 == 0) { const size_t a = lidx(L, t, 0); d0[a] = s0[a]; if (s1) d1[a] = s1[a]; }
    if (L.j0 + L.nyl == L.ny) {
      const size_t a = lidx(L, t, L.nyl - 1); d0[a] = s0[a]; if (s1) d1[a] = s1[a];
    }
  }
}

cudaError_t launch_copy_ring(const Layout& L, const double* s0, double* d0, const double* s1, double* d1,
                             cudaStream_t st) {
  const int n1 = L.nyl > L.nx ? L.nyl : L.nx;
  k_copy_ring<<<(n1 + 127) / 128, 128, 0, st>>>(L, s0, d0, s1, d1);
  return cudaGetLastError();
}

}  // namespace ifx
