// kernels_v3.cu — sweep kernels, third cut: bulk-copy (TMA 1-D) row pipeline, stencil rows read straight
// from the shared-memory ring (no register rolling, no shuffles), 2*NC2 columns per thread.
//
// Why (ncu, r1_v2_*): v2 already moved exactly the algorithmic bytes, but the Poisson sweep was ISSUE-bound
// (76 % issue-slot utilisation, 160 thread-instructions per cell, most of them per-row overhead amortised
// over only two cells) and the predictor was latency-bound at 2 CTAs/SM (139 registers: u and v rolled
// through registers by the same thread).  v3 therefore
//   * gives every thread 2*NC2 adjacent columns of a row, so per-row overhead (barrier wait, table loads,
//     index math) is paid once per 4-8 cells;
//   * keeps rows S, C, N of the stencil in three live stages of the ring and reads them with LDS.128 —
//     no loop-carried register state, no MOV chains, no warp shuffles;
//   * splits the predictor by FIELD across warps: warps [0,CW) relax u, warps [CW,2CW) relax v, both
//     fed by the same stage (u, v, sx, sy, cell type), which doubles the resident warps per byte of
//     shared memory and halves the registers per thread.
// Arithmetic is stencil_math.cuh's, bit-identical to v1/v2 and to the reference build.
#include "kernels.cuh"
#include "pipeline.cuh"
#include "stencil_math.cuh"

namespace ifx {

enum SweepMode { MODE_PPE_LAPLACE = 0, MODE_PPE_GENERAL = 1, MODE_AD = 2 };

// bytes of a row segment starting at padded-row offset `off` (doubles), clamped so it never leaves the row
__device__ __forceinline__ uint32_t seg_bytes_v3(int want, int off, int pitch) {
  const int n = min(want, pitch - off);
  return (uint32_t)(n > 0 ? n : 0) * 8u;
}

template <int MODE, int NC2, int CW>
struct V3Geom {
  static constexpr int NCOL = 2 * NC2;                    // columns per thread
  static constexpr int TW = 32 * NCOL * CW;               // interior columns per CTA
  static constexpr int SEG = TW + 4;                      // stencil row segment incl. 2-column halos
  static constexpr int NFIELD = (MODE == MODE_AD) ? 2 : 1;
  static constexpr int CWARPS = CW * NFIELD;              // consumer warps
  static constexpr int THREADS = 32 * (CWARPS + 1);
  // stage layout (bytes): [field0 seg][field1 seg]?[pt0: TW doubles]?[pt1]?[ct: TW bytes]?
  static constexpr int OFF_F1 = SEG * 8;
  static constexpr int OFF_PT0 = NFIELD * SEG * 8;
  static constexpr int NPT = (MODE == MODE_PPE_LAPLACE) ? 0 : NFIELD;
  static constexpr int OFF_CT = OFF_PT0 + NPT * TW * 8;
  static constexpr int STAGE_BYTES = ((OFF_CT + (MODE == MODE_PPE_LAPLACE ? 0 : TW)) + 127) / 128 * 128;
};

struct SweepArgsV3 {
  Layout L;
  Metrics M;
  double* fC[2];                 // input iterate(s): p | u, v   (ghost ring of u, v is written, see kernels_ad.cu)
  double* fT[2];                 // output iterate(s)
  const double* pt[2];           // point-wise source: rhs | sx, sy
  const uint8_t* celltype;
  double* res[2];                // reference-layout residual arrays (WRITE_RES)
  double* partials;
  LoopCtl* ctl;
  ReduceCfg rc;
  double two_bc[2][4];           // 2*bc per field for W, E, S, N (MODE_AD)
  int rows_per_cta;
  int force;
};

template <int MODE, bool WRITE_RES, bool HAS_GC, int NC2, int CW, int STAGES>
static __global__ void __launch_bounds__(32 * (CW * (MODE == MODE_AD ? 2 : 1) + 1))
k_sweep_v3(SweepArgsV3 a) {
  using G = V3Geom<MODE, NC2, CW>;
  static_assert((STAGES & (STAGES - 1)) == 0 && STAGES >= 4, "STAGES must be a power of two >= 4");
  if (a.ctl->done && !a.force) return;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + (size_t)STAGES * G::STAGE_BYTES);
  const uint32_t bar_full = smem_u32(bars), bar_empty = smem_u32(bars + STAGES);

  const Layout L = a.L;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int i0 = 1 + blockIdx.x * G::TW;
  const int jfirst = L.jb + blockIdx.y * a.rows_per_cta;
  const int jlast = min(jfirst + a.rows_per_cta, L.je);
  const int nst = (jlast - jfirst) + 2;                    // rows jfirst-1 .. jlast
  const int nxm2 = L.nx - 2, nym2 = L.ny - 2;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, G::CWARPS); }
    mbar_fence_init();
  }
  __syncthreads();

  double r0 = 0.0, r1 = 0.0;

  if (warp == G::CWARPS) {
    // ------------------------------------ producer ------------------------------------
    if (lane == 0) {
      const int off_seg = IFX_PADL + i0 - 2, off_pt = IFX_PADL + i0;
      const uint32_t b_seg = seg_bytes_v3(G::SEG, off_seg, L.pitch);
      const uint32_t b_pt = seg_bytes_v3(G::TW, off_pt, L.pitch);
      size_t row = (size_t)(jfirst - 1 - L.j0) * L.pitch;
      for (int k = 0; k < nst; ++k, row += L.pitch) {
        const int s = k & (STAGES - 1);
        if (k >= STAGES) mbar_wait(bar_empty + 8 * s, ((k / STAGES) - 1) & 1);
        const uint32_t dst = smem_u32(smem_raw) + (uint32_t)s * G::STAGE_BYTES;
        const uint32_t bf = bar_full + 8 * s;
        const bool owned = (k >= 1 && k <= nst - 2);
        uint32_t tx = G::NFIELD * b_seg;
        if (MODE != MODE_PPE_LAPLACE && owned) tx += G::NPT * b_pt + b_pt / 8;
        mbar_arrive_expect_tx(bf, tx);
        bulk_g2s(dst, a.fC[0] + row + off_seg, b_seg, bf);
        if (G::NFIELD == 2) bulk_g2s(dst + G::OFF_F1, a.fC[1] + row + off_seg, b_seg, bf);
        if (MODE != MODE_PPE_LAPLACE && owned) {
          bulk_g2s(dst + G::OFF_PT0, a.pt[0] + row + off_pt, b_pt, bf);
          if (G::NPT == 2) bulk_g2s(dst + G::OFF_PT0 + G::TW * 8, a.pt[1] + row + off_pt, b_pt, bf);
          bulk_g2s(dst + G::OFF_CT, a.celltype + row + off_pt, b_pt / 8, bf);
        }
      }
    }
    __syncwarp();
  } else {
    // ------------------------------------ consumers ------------------------------------
    const int fld = (MODE == MODE_AD) ? (warp / CW) : 0;                // which field this warp relaxes
    const int wq = (MODE == MODE_AD) ? (warp % CW) : warp;
    const int cl = (wq * 32 + lane) * G::NCOL;                          // first of my columns inside the tile
    const int i = i0 + cl;                                              // ... in the grid
    const int c = 2 + cl;                                               // ... inside a row segment
    double* __restrict__ outT = a.fT[fld];
    double* __restrict__ ringC = a.fC[fld];
    const uint32_t off_f = (uint32_t)fld * G::OFF_F1 + (uint32_t)c * 8;
    const uint32_t off_p = G::OFF_PT0 + (uint32_t)fld * G::TW * 8 + (uint32_t)cl * 8;
    const uint32_t off_c = G::OFF_CT + (uint32_t)cl;

    // column coefficients
    double cE[G::NCOL], cW[G::NCOL], cX[G::NCOL];
#pragma unroll
    for (int q = 0; q < G::NCOL; ++q) {
      const int iq = (i + q <= nxm2) ? i + q : 1;
      if (MODE == MODE_AD) { cE[q] = a.M.ad_cE[iq]; cW[q] = a.M.ad_cW[iq]; cX[q] = a.M.ad_px[iq]; }
      else { cE[q] = a.M.pp_cE[iq]; cW[q] = a.M.pp_cW[iq]; cX[q] = a.M.pp_sx[iq]; }
    }
    const double kk = a.M.k;
    const double bcW = a.two_bc[fld][0], bcE = a.two_bc[fld][1], bcS = a.two_bc[fld][2], bcN = a.two_bc[fld][3];

    for (int k = 0; k < nst; ++k) {
      const int s = k & (STAGES - 1);
      mbar_wait(bar_full + 8 * s, (k / STAGES) & 1);
      if (k < 2) continue;
      const unsigned char* stN = smem_raw + (size_t)s * G::STAGE_BYTES;
      const unsigned char* stC = smem_raw + (size_t)((k - 1) & (STAGES - 1)) * G::STAGE_BYTES;
      const unsigned char* stS = smem_raw + (size_t)((k - 2) & (STAGES - 1)) * G::STAGE_BYTES;
      const int j = jfirst + k - 2, jl = j - L.j0;
      const bool top = (j == nym2), bot = (j == 1);

      double qC[G::NCOL + 2], qN[G::NCOL], qS[G::NCOL], src[G::NCOL];
      unsigned char ct[G::NCOL];
#pragma unroll
      for (int h = 0; h < NC2; ++h) {
        const double2 vc = *reinterpret_cast<const double2*>(stC + off_f + 16 * h);
        const double2 vn = *reinterpret_cast<const double2*>(stN + off_f + 16 * h);
        const double2 vs = *reinterpret_cast<const double2*>(stS + off_f + 16 * h);
        qC[1 + 2 * h] = vc.x; qC[2 + 2 * h] = vc.y;
        qN[2 * h] = vn.x; qN[2 * h + 1] = vn.y;
        qS[2 * h] = vs.x; qS[2 * h + 1] = vs.y;
        if (MODE != MODE_PPE_LAPLACE) {
          const double2 sv = *reinterpret_cast<const double2*>(stC + off_p + 16 * h);
          src[2 * h] = sv.x; src[2 * h + 1] = sv.y;
          const uchar2 cv = *reinterpret_cast<const uchar2*>(stC + off_c + 2 * h);
          ct[2 * h] = cv.x; ct[2 * h + 1] = cv.y;
        }
      }
      qC[0] = *reinterpret_cast<const double*>(stC + off_f - 8);              // west neighbour of my first column
      qC[G::NCOL + 1] = *reinterpret_cast<const double*>(stC + off_f + 8 * G::NCOL);   // east neighbour of my last
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_empty + 8 * ((k - 2) & (STAGES - 1)));     // row S is no longer needed

      const double cN = (MODE == MODE_AD) ? a.M.ad_cN[j] : a.M.pp_cN[j];
      const double cS = (MODE == MODE_AD) ? a.M.ad_cS[j] : a.M.pp_cS[j];
      const double sy = (MODE == MODE_AD) ? a.M.ad_sy[j] : a.M.pp_sy[j];
      const size_t o = lidx(L, i, jl);
      double out[G::NCOL];
      bool wr[G::NCOL];
#pragma unroll
      for (int q = 0; q < G::NCOL; ++q) {
        const int iq = i + q;
        const bool act = iq <= nxm2;
        const double pc = qC[q + 1];
        double pw = qC[q], pe = qC[q + 2], pn = qN[q], ps = qS[q];
        if (MODE == MODE_AD) {
          // virtual ghosts (set_velocity_BC, ADSolver.cu:199-217): ghost = -interior + 2*bc
          if (iq == 1) pw = bcW - pc;
          if (iq == nxm2) pe = bcE - pc;
          if (bot) ps = bcS - pc;
          if (top) pn = bcN - pc;
          const double cP = fma(kk, sy, cX[q]);                                   // ADSolver.cu:34
          const bool fluid = ct[q] == IFX_FLUID;
          const double nv = jac_cell(src[q], cE[q], pe, cW[q], pw, cN, pn, cS, ps, fluid ? 1.0 : 0.0, cP);
          out[q] = nv;
          wr[q] = act && !(HAS_GC && ct[q] == IFX_GHOST);
          const double rr = (act && fluid) ? fabs(pc - nv) : 0.0;                 // ADSolver.cu:131-137
          if (fld == 0) r0 += rr; else r1 += rr;
          if (WRITE_RES && act) a.res[fld][(size_t)j * L.nx + iq] = rr;
          // ghost ring of the INPUT buffer, as set_velocity_BC leaves it (nobody reads it in this launch)
          if (act && (iq == 1 || iq == nxm2 || bot || top)) {
            if (iq == 1) ringC[lidx(L, 0, jl)] = pw;
            if (iq == nxm2) ringC[lidx(L, L.nx - 1, jl)] = pe;
            if (bot) {
              ringC[lidx(L, iq, jl - 1)] = ps;
              if (iq == 1) ringC[lidx(L, 0, jl - 1)] = bcS - pw;                  // corners: 2bc - (2bc - diagonal)
              if (iq == nxm2) ringC[lidx(L, L.nx - 1, jl - 1)] = bcE - ps;
            }
            if (top) {
              ringC[lidx(L, iq, jl + 1)] = pn;
              if (iq == 1) ringC[lidx(L, 0, jl + 1)] = bcN - pw;
              if (iq == nxm2) ringC[lidx(L, L.nx - 1, jl + 1)] = bcN - pe;
            }
          }
        } else {
          if (MODE == MODE_PPE_GENERAL) {                                         // homogeneous Neumann, virtual
            if (iq == 1) pw = pc;
            if (iq == nxm2) pe = pc;
            if (bot) ps = pc;
            if (top) pn = pc;
          }
          const double cP = -(cX[q] + sy);                                        // PPESolver.cu:93-94
          const double t = ppe_offdiag(pw, cW[q], pe, cE[q], pn, cN, ps, cS);
          const double qq = ppe_apply(pc, cP, pw, cW[q], pe, cE[q], pn, cN, ps, cS);
          double rr;
          if (MODE == MODE_PPE_LAPLACE) {
            out[q] = (-t) / cP;                                                   // PPESolver.cu:24-27
            rr = act ? qq : 0.0;                                                  // :42-46
            wr[q] = act;
          } else {
            const bool fluid = ct[q] == IFX_FLUID;
            out[q] = fluid ? (src[q] - t) / cP : pc;
            rr = (act && fluid) ? src[q] - qq : 0.0;
            wr[q] = act && !(HAS_GC && ct[q] == IFX_GHOST);
          }
          r0 += rr; r1 += fabs(rr);
          if (WRITE_RES && act) a.res[0][(size_t)j * L.nx + iq] = rr;
        }
      }
#pragma unroll
      for (int h = 0; h < NC2; ++h) {
        if (wr[2 * h] && wr[2 * h + 1]) *reinterpret_cast<double2*>(outT + o + 2 * h) = make_double2(out[2 * h], out[2 * h + 1]);
        else {
          if (wr[2 * h]) outT[o + 2 * h] = out[2 * h];
          if (wr[2 * h + 1]) outT[o + 2 * h + 1] = out[2 * h + 1];
        }
      }
    }
    // release the two stages still held (rows jlast-1, jlast) — nobody waits for them, nothing to do
  }
  block_reduce_and_decide<G::THREADS>(r0, r1, a.partials, a.ctl, a.rc,
                                      blockIdx.y * gridDim.x + blockIdx.x, gridDim.x * gridDim.y);
}

// =================================================================================================
// launchers
// =================================================================================================
template <int MODE, int NC2, int CW, int STAGES>
static cudaError_t v3_dispatch(const SweepArgsV3& a, dim3 grid, cudaStream_t st, bool write_res, bool has_gc) {
  using G = V3Geom<MODE, NC2, CW>;
  const size_t sm = (size_t)STAGES * G::STAGE_BYTES + 2 * STAGES * 8;
#define IFX_GO3(WR, GC)                                                                                  \
  do {                                                                                                   \
    auto kern = k_sweep_v3<MODE, WR, GC, NC2, CW, STAGES>;                                               \
    static bool once = false;                                                                            \
    if (!once) {                                                                                         \
      cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);  \
      if (e != cudaSuccess) return e;                                                                    \
      once = true;                                                                                       \
    }                                                                                                    \
    kern<<<grid, G::THREADS, sm, st>>>(a);                                                               \
  } while (0)
  if (MODE == MODE_PPE_LAPLACE) {
    if (write_res) IFX_GO3(true, false); else IFX_GO3(false, false);
  } else if (has_gc) {
    if (write_res) IFX_GO3(true, true); else IFX_GO3(false, true);
  } else {
    if (write_res) IFX_GO3(true, false); else IFX_GO3(false, false);
  }
#undef IFX_GO3
  return cudaGetLastError();
}

int v3_tile_cols(int mode, int tune) {
  // must mirror the dispatch tables below
  if (mode == MODE_AD) {
    switch (tune) { case 1: return V3Geom<MODE_AD, 2, 4>::TW; case 2: return V3Geom<MODE_AD, 4, 2>::TW;
                    case 3: return V3Geom<MODE_AD, 1, 4>::TW; default: return V3Geom<MODE_AD, 2, 2>::TW; }
  }
  switch (tune) { case 1: return V3Geom<MODE_PPE_LAPLACE, 2, 4>::TW; case 2: return V3Geom<MODE_PPE_LAPLACE, 4, 4>::TW;
                  case 3: return V3Geom<MODE_PPE_LAPLACE, 2, 2>::TW; default: return V3Geom<MODE_PPE_LAPLACE, 4, 2>::TW; }
}

static SweepArgsV3 from_ppe(const PpeSweepArgs& p) {
  SweepArgsV3 a{};
  a.L = p.L; a.M = p.M;
  a.fC[0] = p.pC; a.fT[0] = p.pT; a.pt[0] = p.rhs; a.celltype = p.celltype; a.res[0] = p.res;
  a.partials = p.partials; a.ctl = p.ctl; a.rc = p.rc; a.rows_per_cta = p.rows_per_cta; a.force = p.force;
  return a;
}

cudaError_t launch_ppe_sweep_v3(const PpeSweepArgs& p, dim3 grid, cudaStream_t st, bool laplace_ref, bool write_res,
                                bool has_gc, int tune) {
  const SweepArgsV3 a = from_ppe(p);
#define IFX_PPE(NC2, CW, ST)                                                                              \
  return laplace_ref ? v3_dispatch<MODE_PPE_LAPLACE, NC2, CW, ST>(a, grid, st, write_res, false)          \
                     : v3_dispatch<MODE_PPE_GENERAL, NC2, CW, ST>(a, grid, st, write_res, has_gc)
  switch (tune) {
    case 1: IFX_PPE(2, 4, 8);
    case 2: IFX_PPE(4, 4, 8);
    case 3: IFX_PPE(2, 2, 8);
    default: IFX_PPE(4, 2, 8);
  }
#undef IFX_PPE
}

cudaError_t launch_ad_jacobi_v3(const AdJacobiArgs& p, dim3 grid, cudaStream_t st, bool write_res, bool has_gc, int tune) {
  SweepArgsV3 a{};
  a.L = p.L; a.M = p.M;
  a.fC[0] = p.uC; a.fC[1] = p.vC; a.fT[0] = p.uT; a.fT[1] = p.vT;
  a.pt[0] = p.sx; a.pt[1] = p.sy; a.celltype = p.celltype; a.res[0] = p.res_u; a.res[1] = p.res_v;
  a.partials = p.partials; a.ctl = p.ctl; a.rc = p.rc; a.rows_per_cta = p.rows_per_cta; a.force = p.force;
  for (int q = 0; q < 4; ++q) { a.two_bc[0][q] = p.two_bc_u[q]; a.two_bc[1][q] = p.two_bc_v[q]; }
  switch (tune) {
    case 1: return v3_dispatch<MODE_AD, 2, 4, 8>(a, grid, st, write_res, has_gc);
    case 2: return v3_dispatch<MODE_AD, 4, 2, 8>(a, grid, st, write_res, has_gc);
    case 3: return v3_dispatch<MODE_AD, 1, 4, 8>(a, grid, st, write_res, has_gc);
    default: return v3_dispatch<MODE_AD, 2, 2, 8>(a, grid, st, write_res, has_gc);
  }
}

}  // namespace ifx
