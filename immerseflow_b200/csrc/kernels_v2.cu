// kernels_v2.cu — the sweep kernels re-built around an async bulk-copy (TMA 1-D) pipeline.
//
// Same arithmetic, same launch semantics, same reductions as kernels_ad.cu / kernels_ppe.cu (v1) — the
// cell math comes from stencil_math.cuh, so v1 and v2 agree bit for bit — but the loads are issued by
// one elected thread of a producer warp as cp.async.bulk row segments landing in a ring of shared-memory
// stages (mbarrier complete_tx), while CW consumer warps roll the rows through registers and write the
// results straight to HBM as 16-byte vectors.  In-flight bytes per CTA = STAGES x row-segment bytes.
//
// Tile = 64*CW interior columns x rows_per_cta rows.  Stage k of a CTA carries row (jfirst-1+k) of the
// stencil field(s) with a 2-column halo on each side (keeps every segment 16-byte aligned: the first
// interior column sits on a 128-byte boundary, IFX_PADL) and, for rows the CTA owns, the matching row
// of the point-wise fields (rhs / sx, sy, cell type).
#include "kernels.cuh"
#include "pipeline.cuh"
#include "stencil_math.cuh"

namespace ifx {

template <int CW>
struct TileGeom {
  static constexpr int TW = 64 * CW;        // interior columns per CTA
  static constexpr int SEG = TW + 4;        // stencil-field row segment incl. halo
  static constexpr int THREADS = 32 * (CW + 1);
};

// bytes of a row segment starting at padded-row offset `off` (doubles), clamped so it never leaves the row
__device__ __forceinline__ uint32_t seg_bytes(int want, int off, int pitch) {
  const int n = min(want, pitch - off);
  return (uint32_t)(n > 0 ? n : 0) * 8u;
}

// =================================================================================================
// Poisson sweep (see kernels_ppe.cu for the semantics)
// =================================================================================================
template <bool LAPLACE_REF, bool WRITE_RES, bool HAS_GC, int CW, int STAGES>
static __global__ void __launch_bounds__(32 * (CW + 1))
k_ppe_sweep_v2(PpeSweepArgs a) {
  using G = TileGeom<CW>;
  if (a.ctl->done && !a.force) return;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  // stage layout: [p seg: SEG doubles][rhs: TW doubles][ct: TW bytes], 128-byte aligned stages
  constexpr int STAGE_BYTES = ((G::SEG * 8 + (LAPLACE_REF ? 0 : G::TW * 8 + G::TW)) + 127) / 128 * 128;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + (size_t)STAGES * STAGE_BYTES);
  const uint32_t bar_full = smem_u32(bars), bar_empty = smem_u32(bars + STAGES);

  const Layout L = a.L;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int i0 = 1 + blockIdx.x * G::TW;                                  // first interior column of the tile
  const int jfirst = L.jb + blockIdx.y * a.rows_per_cta;
  const int jlast = min(jfirst + a.rows_per_cta, L.je);
  const int nst = (jlast - jfirst) + 2;                                  // rows jfirst-1 .. jlast
  const int nxm2 = L.nx - 2, nym2 = L.ny - 2;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, CW); }
    mbar_fence_init();
  }
  __syncthreads();

  double rsum = 0.0, rabs = 0.0;

  if (warp == CW) {
    // ------------------------------ producer ------------------------------
    if (lane == 0) {
      const int off_seg = IFX_PADL + i0 - 2, off_pt = IFX_PADL + i0;
      const uint32_t b_seg = seg_bytes(G::SEG, off_seg, L.pitch);
      const uint32_t b_pt = seg_bytes(G::TW, off_pt, L.pitch);
      for (int k = 0; k < nst; ++k) {
        const int s = k % STAGES;
        if (k >= STAGES) mbar_wait(bar_empty + 8 * s, ((k / STAGES) - 1) & 1);
        const size_t row = (size_t)(jfirst - 1 - L.j0 + k) * L.pitch;
        const uint32_t dst = smem_u32(smem_raw + (size_t)s * STAGE_BYTES);
        const bool owned = (k >= 1 && k <= nst - 2);
        uint32_t tx = b_seg;
        if (!LAPLACE_REF && owned) tx += b_pt + b_pt / 8;
        mbar_arrive_expect_tx(bar_full + 8 * s, tx);
        bulk_g2s(dst, a.pC + row + off_seg, b_seg, bar_full + 8 * s);
        if (!LAPLACE_REF && owned) {
          bulk_g2s(dst + G::SEG * 8, a.rhs + row + off_pt, b_pt, bar_full + 8 * s);
          bulk_g2s(dst + G::SEG * 8 + G::TW * 8, a.celltype + row + off_pt, b_pt / 8, bar_full + 8 * s);
        }
      }
    }
    __syncwarp();
  } else {
    // ------------------------------ consumers ------------------------------
    const int i = i0 + warp * 64 + lane * 2;                             // my two columns
    const int c = 2 + warp * 64 + lane * 2;                              // their index inside the segment
    const bool act0 = i <= nxm2, act1 = i + 1 <= nxm2;
    const int ic = act0 ? i : 1;                                         // for table lookups only
    const double cE0 = a.M.pp_cE[ic], cW0 = a.M.pp_cW[ic], sx0 = a.M.pp_sx[ic];
    const double cE1 = a.M.pp_cE[ic + 1], cW1 = a.M.pp_cW[ic + 1], sx1 = a.M.pp_sx[ic + 1];
    const bool ring_w = (i == 1), ring_e0 = (i == nxm2), ring_e1 = (i + 1 == nxm2);

    double2 pS = make_double2(0, 0), pC = pS, pN = pS;
    double hwC = 0, heC = 0, hwN = 0, heN = 0;
    double2 fC = pS, fN = pS;
    uchar2 ctC = make_uchar2(IFX_FLUID, IFX_FLUID), ctN = ctC;

    for (int k = 0; k < nst; ++k) {
      const int s = k % STAGES;
      mbar_wait(bar_full + 8 * s, (k / STAGES) & 1);
      const unsigned char* st = smem_raw + (size_t)s * STAGE_BYTES;
      const double* prow = reinterpret_cast<const double*>(st);
      pN = *reinterpret_cast<const double2*>(prow + c);
      if (lane == 0) hwN = prow[c - 1];
      if (lane == 31) heN = prow[c + 2];
      if (!LAPLACE_REF) {
        fN = *reinterpret_cast<const double2*>(prow + G::SEG + (c - 2));
        ctN = *reinterpret_cast<const uchar2*>(st + G::SEG * 8 + G::TW * 8 + (c - 2));
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_empty + 8 * s);

      if (k >= 2) {
        const int j = jfirst + k - 2, jl = j - L.j0;
        const bool top = (j == nym2), bot = (j == 1);
        const double cN = a.M.pp_cN[j], cS = a.M.pp_cS[j], sy = a.M.pp_sy[j];
        double pW0 = __shfl_up_sync(0xffffffffu, pC.y, 1);
        double pE1 = __shfl_down_sync(0xffffffffu, pC.x, 1);
        if (lane == 0) pW0 = hwC;
        if (lane == 31) pE1 = heC;
        double pE0 = pC.y, pW1 = pC.x;
        double2 pSs = pS, pNn = pN;
        if (!LAPLACE_REF) {      // homogeneous Neumann through virtual ghosts
          if (ring_w) pW0 = pC.x;
          if (ring_e0) pE0 = pC.x;
          if (ring_e1) pE1 = pC.y;
          if (bot) pSs = pC;
          if (top) pNn = pC;
        }
        const double cP0 = -(sx0 + sy), cP1 = -(sx1 + sy);               // PPESolver.cu:93-94
        const size_t o = lidx(L, i, jl);
        const double t0 = ppe_offdiag(pW0, cW0, pE0, cE0, pNn.x, cN, pSs.x, cS);
        const double t1 = ppe_offdiag(pW1, cW1, pE1, cE1, pNn.y, cN, pSs.y, cS);
        const double q0 = ppe_apply(pC.x, cP0, pW0, cW0, pE0, cE0, pNn.x, cN, pSs.x, cS);
        const double q1 = ppe_apply(pC.y, cP1, pW1, cW1, pE1, cE1, pNn.y, cN, pSs.y, cS);
        double2 pn, r;
        if (LAPLACE_REF) {
          pn.x = (-t0) / cP0; pn.y = (-t1) / cP1;                        // PPESolver.cu:24-27
          r.x = act0 ? q0 : 0.0; r.y = act1 ? q1 : 0.0;
          if (act1) *reinterpret_cast<double2*>(a.pT + o) = pn;
          else if (act0) a.pT[o] = pn.x;
        } else {
          const bool fl0 = ctC.x == IFX_FLUID, fl1 = ctC.y == IFX_FLUID;
          pn.x = fl0 ? (fC.x - t0) / cP0 : pC.x;
          pn.y = fl1 ? (fC.y - t1) / cP1 : pC.y;
          r.x = (act0 && fl0) ? fC.x - q0 : 0.0;
          r.y = (act1 && fl1) ? fC.y - q1 : 0.0;
          if (HAS_GC) {
            if (act0 && ctC.x != IFX_GHOST) a.pT[o] = pn.x;
            if (act1 && ctC.y != IFX_GHOST) a.pT[o + 1] = pn.y;
          } else if (act1) *reinterpret_cast<double2*>(a.pT + o) = pn;
          else if (act0) a.pT[o] = pn.x;
        }
        rsum += r.x; rsum += r.y;
        rabs += fabs(r.x); rabs += fabs(r.y);
        if (WRITE_RES) {
          const size_t ro = (size_t)j * L.nx + i;
          if (act0) a.res[ro] = r.x;
          if (act1) a.res[ro + 1] = r.y;
        }
      }
      pS = pC; pC = pN; hwC = hwN; heC = heN; fC = fN; ctC = ctN;
    }
  }
  block_reduce_and_decide<G::THREADS>(rsum, rabs, a.partials, a.ctl, a.rc,
                                      blockIdx.y * gridDim.x + blockIdx.x, gridDim.x * gridDim.y);
}

// =================================================================================================
// Predictor Jacobi iteration (see kernels_ad.cu for the semantics)
// =================================================================================================
template <bool WRITE_RES, bool HAS_GC, int CW, int STAGES>
static __global__ void __launch_bounds__(32 * (CW + 1))
k_ad_jacobi_v2(AdJacobiArgs a) {
  using G = TileGeom<CW>;
  if (a.ctl->done && !a.force) return;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  // stage layout: [u seg][v seg][sx][sy][ct]
  constexpr int OFF_V = G::SEG * 8, OFF_SX = 2 * G::SEG * 8, OFF_SY = OFF_SX + G::TW * 8, OFF_CT = OFF_SY + G::TW * 8;
  constexpr int STAGE_BYTES = (OFF_CT + G::TW + 127) / 128 * 128;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + (size_t)STAGES * STAGE_BYTES);
  const uint32_t bar_full = smem_u32(bars), bar_empty = smem_u32(bars + STAGES);

  const Layout L = a.L;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int i0 = 1 + blockIdx.x * G::TW;
  const int jfirst = L.jb + blockIdx.y * a.rows_per_cta;
  const int jlast = min(jfirst + a.rows_per_cta, L.je);
  const int nst = (jlast - jfirst) + 2;
  const int nxm2 = L.nx - 2, nym2 = L.ny - 2;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, CW); }
    mbar_fence_init();
  }
  __syncthreads();

  double resu = 0.0, resv = 0.0;

  if (warp == CW) {
    if (lane == 0) {
      const int off_seg = IFX_PADL + i0 - 2, off_pt = IFX_PADL + i0;
      const uint32_t b_seg = seg_bytes(G::SEG, off_seg, L.pitch);
      const uint32_t b_pt = seg_bytes(G::TW, off_pt, L.pitch);
      for (int k = 0; k < nst; ++k) {
        const int s = k % STAGES;
        if (k >= STAGES) mbar_wait(bar_empty + 8 * s, ((k / STAGES) - 1) & 1);
        const size_t row = (size_t)(jfirst - 1 - L.j0 + k) * L.pitch;
        const uint32_t dst = smem_u32(smem_raw + (size_t)s * STAGE_BYTES);
        const uint32_t bf = bar_full + 8 * s;
        const bool owned = (k >= 1 && k <= nst - 2);
        mbar_arrive_expect_tx(bf, 2 * b_seg + (owned ? 2 * b_pt + b_pt / 8 : 0));
        bulk_g2s(dst, a.uC + row + off_seg, b_seg, bf);
        bulk_g2s(dst + OFF_V, a.vC + row + off_seg, b_seg, bf);
        if (owned) {
          bulk_g2s(dst + OFF_SX, a.sx + row + off_pt, b_pt, bf);
          bulk_g2s(dst + OFF_SY, a.sy + row + off_pt, b_pt, bf);
          bulk_g2s(dst + OFF_CT, a.celltype + row + off_pt, b_pt / 8, bf);
        }
      }
    }
    __syncwarp();
  } else {
    const int i = i0 + warp * 64 + lane * 2;
    const int c = 2 + warp * 64 + lane * 2;
    const bool act0 = i <= nxm2, act1 = i + 1 <= nxm2;
    const int ic = act0 ? i : 1;
    const double cE0 = a.M.ad_cE[ic], cW0 = a.M.ad_cW[ic], px0 = a.M.ad_px[ic];
    const double cE1 = a.M.ad_cE[ic + 1], cW1 = a.M.ad_cW[ic + 1], px1 = a.M.ad_px[ic + 1];
    const double kk = a.M.k;
    const bool w0 = (i == 1), e0 = (i == nxm2), e1 = (i + 1 == nxm2);

    const double2 z = make_double2(0, 0);
    double2 uS = z, vS = z, uC = z, vC = z, uN = z, vN = z, sxC = z, syC = z, sxN = z, syN = z;
    double uhwC = 0, uheC = 0, vhwC = 0, vheC = 0, uhwN = 0, uheN = 0, vhwN = 0, vheN = 0;
    uchar2 ctC = make_uchar2(IFX_FLUID, IFX_FLUID), ctN = ctC;

    for (int k = 0; k < nst; ++k) {
      const int s = k % STAGES;
      mbar_wait(bar_full + 8 * s, (k / STAGES) & 1);
      const unsigned char* st = smem_raw + (size_t)s * STAGE_BYTES;
      const double* urow = reinterpret_cast<const double*>(st);
      const double* vrow = reinterpret_cast<const double*>(st + OFF_V);
      uN = *reinterpret_cast<const double2*>(urow + c);
      vN = *reinterpret_cast<const double2*>(vrow + c);
      if (lane == 0) { uhwN = urow[c - 1]; vhwN = vrow[c - 1]; }
      if (lane == 31) { uheN = urow[c + 2]; vheN = vrow[c + 2]; }
      sxN = *reinterpret_cast<const double2*>(st + OFF_SX + (c - 2) * 8);
      syN = *reinterpret_cast<const double2*>(st + OFF_SY + (c - 2) * 8);
      ctN = *reinterpret_cast<const uchar2*>(st + OFF_CT + (c - 2));
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_empty + 8 * s);

      if (k >= 2) {
        const int j = jfirst + k - 2, jl = j - L.j0;
        const bool top = (j == nym2), bot = (j == 1);
        const double cN = a.M.ad_cN[j], cS = a.M.ad_cS[j], sy = a.M.ad_sy[j];

        double uW0 = __shfl_up_sync(0xffffffffu, uC.y, 1), vW0 = __shfl_up_sync(0xffffffffu, vC.y, 1);
        double uE1 = __shfl_down_sync(0xffffffffu, uC.x, 1), vE1 = __shfl_down_sync(0xffffffffu, vC.x, 1);
        if (lane == 0) { uW0 = uhwC; vW0 = vhwC; }
        if (lane == 31) { uE1 = uheC; vE1 = vheC; }
        // virtual ghosts (set_velocity_BC, ADSolver.cu:199-217): ghost = -interior + 2*bc
        const double ugw = a.two_bc_u[0] - uC.x, vgw = a.two_bc_v[0] - vC.x;
        if (w0) { uW0 = ugw; vW0 = vgw; }
        double uE0 = uC.y, vE0 = vC.y, uW1 = uC.x, vW1 = vC.x;
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
        const double cP0 = fma(kk, sy, px0), cP1 = fma(kk, sy, px1);
        const double ib0 = (ctC.x == IFX_FLUID) ? 1.0 : 0.0, ib1 = (ctC.y == IFX_FLUID) ? 1.0 : 0.0;
        double2 un, vn;
        un.x = jac_cell(sxC.x, cE0, uE0, cW0, uW0, cN, uNn.x, cS, uSs.x, ib0, cP0);
        un.y = jac_cell(sxC.y, cE1, uE1, cW1, uW1, cN, uNn.y, cS, uSs.y, ib1, cP1);
        vn.x = jac_cell(syC.x, cE0, vE0, cW0, vW0, cN, vNn.x, cS, vSs.x, ib0, cP0);
        vn.y = jac_cell(syC.y, cE1, vE1, cW1, vW1, cN, vNn.y, cS, vSs.y, ib1, cP1);

        const size_t o = lidx(L, i, jl);
        if (HAS_GC) {
          if (act0 && ctC.x != IFX_GHOST) a.uT[o] = un.x, a.vT[o] = vn.x;
          if (act1 && ctC.y != IFX_GHOST) a.uT[o + 1] = un.y, a.vT[o + 1] = vn.y;
        } else if (act1) {
          *reinterpret_cast<double2*>(a.uT + o) = un;
          *reinterpret_cast<double2*>(a.vT + o) = vn;
        } else if (act0) {
          a.uT[o] = un.x; a.vT[o] = vn.x;
        }

        double ru0 = 0, ru1 = 0, rv0 = 0, rv1 = 0;                       // ADSolver.cu:131-137
        if (act0 && ctC.x == IFX_FLUID) { ru0 = fabs(uC.x - un.x); rv0 = fabs(vC.x - vn.x); }
        if (act1 && ctC.y == IFX_FLUID) { ru1 = fabs(uC.y - un.y); rv1 = fabs(vC.y - vn.y); }
        resu += ru0; resu += ru1; resv += rv0; resv += rv1;
        if (WRITE_RES) {
          const size_t r = (size_t)j * L.nx + i;
          if (act0) { a.res_u[r] = ru0; a.res_v[r] = rv0; }
          if (act1) { a.res_u[r + 1] = ru1; a.res_v[r + 1] = rv1; }
        }

        // ghost ring of the INPUT buffer (see kernels_ad.cu); only tiles touching the boundary get here
        if ((w0 | e0 | e1 | bot | top) && act0) {
          if (w0) { a.uC[lidx(L, 0, jl)] = ugw; a.vC[lidx(L, 0, jl)] = vgw; }
          if (e0) { a.uC[lidx(L, L.nx - 1, jl)] = uE0; a.vC[lidx(L, L.nx - 1, jl)] = vE0; }
          if (act1 && e1) { a.uC[lidx(L, L.nx - 1, jl)] = uE1; a.vC[lidx(L, L.nx - 1, jl)] = vE1; }
          if (bot) {
            a.uC[lidx(L, i, jl - 1)] = uSs.x; a.vC[lidx(L, i, jl - 1)] = vSs.x;
            if (act1) { a.uC[lidx(L, i + 1, jl - 1)] = uSs.y; a.vC[lidx(L, i + 1, jl - 1)] = vSs.y; }
            if (w0) { a.uC[lidx(L, 0, jl - 1)] = a.two_bc_u[2] - ugw; a.vC[lidx(L, 0, jl - 1)] = a.two_bc_v[2] - vgw; }
            if (e0) {
              a.uC[lidx(L, L.nx - 1, jl - 1)] = a.two_bc_u[1] - uSs.x;
              a.vC[lidx(L, L.nx - 1, jl - 1)] = a.two_bc_v[1] - vSs.x;
            }
            if (act1 && e1) {
              a.uC[lidx(L, L.nx - 1, jl - 1)] = a.two_bc_u[1] - uSs.y;
              a.vC[lidx(L, L.nx - 1, jl - 1)] = a.two_bc_v[1] - vSs.y;
            }
          }
          if (top) {
            a.uC[lidx(L, i, jl + 1)] = uNn.x; a.vC[lidx(L, i, jl + 1)] = vNn.x;
            if (act1) { a.uC[lidx(L, i + 1, jl + 1)] = uNn.y; a.vC[lidx(L, i + 1, jl + 1)] = vNn.y; }
            if (w0) { a.uC[lidx(L, 0, jl + 1)] = a.two_bc_u[3] - ugw; a.vC[lidx(L, 0, jl + 1)] = a.two_bc_v[3] - vgw; }
            if (e0) {
              a.uC[lidx(L, L.nx - 1, jl + 1)] = a.two_bc_u[3] - uE0;
              a.vC[lidx(L, L.nx - 1, jl + 1)] = a.two_bc_v[3] - vE0;
            }
            if (act1 && e1) {
              a.uC[lidx(L, L.nx - 1, jl + 1)] = a.two_bc_u[3] - uE1;
              a.vC[lidx(L, L.nx - 1, jl + 1)] = a.two_bc_v[3] - vE1;
            }
          }
        }
      }
      uS = uC; vS = vC; uC = uN; vC = vN;
      uhwC = uhwN; uheC = uheN; vhwC = vhwN; vheC = vheN;
      sxC = sxN; syC = syN; ctC = ctN;
    }
  }
  block_reduce_and_decide<G::THREADS>(resu, resv, a.partials, a.ctl, a.rc,
                                      blockIdx.y * gridDim.x + blockIdx.x, gridDim.x * gridDim.y);
}

// =================================================================================================
// launchers
// =================================================================================================
template <typename K>
static cudaError_t set_smem(K kernel, size_t bytes) {
  return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
}

template <int CW, int STAGES>
static cudaError_t ppe_v2_dispatch(const PpeSweepArgs& a, dim3 grid, cudaStream_t st, bool laplace_ref, bool write_res,
                                   bool has_gc) {
  using G = TileGeom<CW>;
  const size_t stage_l = ((size_t)G::SEG * 8 + 127) / 128 * 128;
  const size_t stage_g = ((size_t)G::SEG * 8 + G::TW * 8 + G::TW + 127) / 128 * 128;
  const size_t sm_l = STAGES * stage_l + 2 * STAGES * 8, sm_g = STAGES * stage_g + 2 * STAGES * 8;
#define IFX_GO(KERNEL, SM)                                          \
  do {                                                              \
    static bool once = false;                                       \
    if (!once) { cudaError_t e = set_smem(KERNEL, SM); if (e != cudaSuccess) return e; once = true; } \
    KERNEL<<<grid, G::THREADS, SM, st>>>(a);                        \
  } while (0)
  if (laplace_ref) {
    if (write_res) IFX_GO((k_ppe_sweep_v2<true, true, false, CW, STAGES>), sm_l);
    else IFX_GO((k_ppe_sweep_v2<true, false, false, CW, STAGES>), sm_l);
  } else if (has_gc) {
    if (write_res) IFX_GO((k_ppe_sweep_v2<false, true, true, CW, STAGES>), sm_g);
    else IFX_GO((k_ppe_sweep_v2<false, false, true, CW, STAGES>), sm_g);
  } else {
    if (write_res) IFX_GO((k_ppe_sweep_v2<false, true, false, CW, STAGES>), sm_g);
    else IFX_GO((k_ppe_sweep_v2<false, false, false, CW, STAGES>), sm_g);
  }
  return cudaGetLastError();
}

template <int CW, int STAGES>
static cudaError_t ad_v2_dispatch(const AdJacobiArgs& a, dim3 grid, cudaStream_t st, bool write_res, bool has_gc) {
  using G = TileGeom<CW>;
  const size_t stage = ((size_t)2 * G::SEG * 8 + 2 * G::TW * 8 + G::TW + 127) / 128 * 128;
  const size_t sm = STAGES * stage + 2 * STAGES * 8;
  if (write_res) {
    if (has_gc) IFX_GO((k_ad_jacobi_v2<true, true, CW, STAGES>), sm);
    else IFX_GO((k_ad_jacobi_v2<true, false, CW, STAGES>), sm);
  } else {
    if (has_gc) IFX_GO((k_ad_jacobi_v2<false, true, CW, STAGES>), sm);
    else IFX_GO((k_ad_jacobi_v2<false, false, CW, STAGES>), sm);
  }
#undef IFX_GO
  return cudaGetLastError();
}

cudaError_t launch_ppe_sweep_v2(const PpeSweepArgs& a, dim3 grid, cudaStream_t st, bool laplace_ref, bool write_res,
                                bool has_gc, int tune) {
  switch (tune) {
    case 1: return ppe_v2_dispatch<4, 16>(a, grid, st, laplace_ref, write_res, has_gc);
    case 2: return ppe_v2_dispatch<4, 24>(a, grid, st, laplace_ref, write_res, has_gc);
    default: return ppe_v2_dispatch<4, 8>(a, grid, st, laplace_ref, write_res, has_gc);
  }
}

cudaError_t launch_ad_jacobi_v2(const AdJacobiArgs& a, dim3 grid, cudaStream_t st, bool write_res, bool has_gc, int tune) {
  switch (tune) {
    case 1: return ad_v2_dispatch<4, 8>(a, grid, st, write_res, has_gc);
    case 2: return ad_v2_dispatch<4, 12>(a, grid, st, write_res, has_gc);
    default: return ad_v2_dispatch<4, 4>(a, grid, st, write_res, has_gc);
  }
}

}  // namespace ifx
