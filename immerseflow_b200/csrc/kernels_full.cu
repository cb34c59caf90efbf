// kernels_full.cu — the stages of the fractional step the reference leaves empty: PPE source term
// (calculateSourcePPE is a stub, PPESolver.cu:106-135), projection (AD_PPE_Correction.cu:1-12 is empty), and the
// boundary-condition refresh for arbitrary BC values (the reference hard-codes them, ADSolver.cu:200-216).
// Semantics: oracle/ifx_oracle_full.c (PARITY UNPINNED), arithmetic identical to it operation for operation.
//
// HBM traffic per cell (algorithmic): source term 16 B read (u*, v*) + 8 B write (+1 B cell type);
// projection 24 B read (u*, v*, p) + 32 B write (u, v, uf, vf) (+1 B): face velocities are produced in the same
// pass that corrects the cell velocities — they are next step's convecting velocities.
#include "kernels.cuh"
#include "fp64_div.cuh"

namespace ifx {

struct FaceCtx {
  Layout L;
  Metrics M;
  const uint8_t* ct;
  const double* ub;      // body velocities (device, 64 entries)
  const double* vb;
};

// velocity on the face east of cell (i, jl) — Compute_velf's interpolation (ADSolver.cu:173-174) with the
// closed-face rule of the oracle
__device__ __forceinline__ double face_u(const FaceCtx& c, const double* __restrict__ u, int i, int jl, bool* open) {
  const size_t o = lidx(c.L, i, jl);
  const uint8_t cw = c.ct[o], ce = c.ct[o + 1];
  if (ce != IFX_FLUID) { *open = false; return c.ub[ce >> 2]; }
  if (cw != IFX_FLUID) { *open = false; return c.ub[cw >> 2]; }
  *open = true;
  return c.M.rcpx[i] * fma(u[o + 1], c.M.dx[i], u[o] * c.M.dx[i + 1]);
}
// velocity on the face north of cell (i, jl) (global row j)
__device__ __forceinline__ double face_v(const FaceCtx& c, const double* __restrict__ v, int i, int jl, int j, bool* open) {
  const size_t o = lidx(c.L, i, jl);
  const uint8_t cs = c.ct[o], cn = c.ct[o + c.L.pitch];
  if (cn != IFX_FLUID) { *open = false; return c.vb[cn >> 2]; }
  if (cs != IFX_FLUID) { *open = false; return c.vb[cs >> 2]; }
  *open = true;
  return c.M.rcpy[j] * fma(c.M.dy[j], v[o + c.L.pitch], v[o] * c.M.dy[j + 1]);
}

// ---------------------------------------------------------------------------------------------
// ghost ring of a field, in place.  neumann = 0: ghost = 2bc - interior (set_velocity_BC, ADSolver.cu:199-217,
// with BC values as parameters); neumann = 1: ghost = interior.  Corners: 2bc - (2bc - diagonal) / diagonal.
// ---------------------------------------------------------------------------------------------
static __global__ void k_apply_ring(Layout L, double* __restrict__ q, double bW, double bE, double bS, double bN, int neumann) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  auto ghost = [&](double two_bc, double interior) { return neumann ? interior : two_bc - interior; };
  if (t >= 1 && t < L.nyl - 1) {            // columns 0 and nx-1 of every stored row except the grid's ghost rows
    const int j = L.j0 + t;
    if (j >= 1 && j <= L.ny - 2) {
      q[lidx(L, 0, t)] = ghost(bW, q[lidx(L, 1, t)]);
      q[lidx(L, L.nx - 1, t)] = ghost(bE, q[lidx(L, L.nx - 2, t)]);
    }
  }
  if (t < L.nx) {
    const int ti = t == 0 ? 1 : (t == L.nx - 1 ? L.nx - 2 : t);        // corners take the diagonal neighbour
    if (L.j0 == 0) {
      double v = ghost(bS, q[lidx(L, ti, 1)]);
      if (t == 0) v = ghost(bS, ghost(bW, q[lidx(L, 1, 1)]));
      if (t == L.nx - 1) v = ghost(bE, ghost(bS, q[lidx(L, L.nx - 2, 1)]));
      q[lidx(L, t, 0)] = v;
    }
    if (L.j0 + L.nyl == L.ny) {
      const int jt = L.nyl - 1;
      double v = ghost(bN, q[lidx(L, ti, jt - 1)]);
      if (t == 0) v = ghost(bN, ghost(bW, q[lidx(L, 1, jt - 1)]));
      if (t == L.nx - 1) v = ghost(bN, ghost(bE, q[lidx(L, L.nx - 2, jt - 1)]));
      q[lidx(L, t, jt)] = v;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// face velocities from cell velocities (first step / after the state was set from outside)
// ---------------------------------------------------------------------------------------------
static __global__ void k_faces_init(FaceCtx c, const double* __restrict__ u, const double* __restrict__ v,
                                    double* __restrict__ uf, double* __restrict__ vf) {
  const Layout& L = c.L;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int jl = blockIdx.y;                     // 0 .. nyl-2
  if (i >= L.nx - 1) return;
  const int j = L.j0 + jl;
  bool open;
  if (jl >= 1) uf[lidx(L, i, jl)] = face_u(c, u, i, jl, &open);                 // rows with an east face: owned rows
  if (i >= 1) vf[lidx(L, i, jl)] = face_v(c, v, i, jl, j, &open);               // north faces of rows j0 .. j0+nyl-2
}

// =============================================================================================
// Strip kernels (source term of the Poisson equation, projection): a block owns 256 columns x HRY rows, a thread
// two adjacent columns (16-byte loads and stores) and walks down its rows with the previous row's values, cell
// types and north-face velocity still in registers, so every field is read once.  The divisions by dx_i, dy_j
// and dt share one refined reciprocal per column / row / launch (fp64_div.cuh) — the quotients are the IEEE
// quotients of the oracle's `/`, the instruction count is a third.
// =============================================================================================
constexpr int HT = 128;          // threads per block
constexpr int HRY = 32;          // rows per block

// velocity on the face between a "low" and a "high" cell (west/east or south/north): Compute_velf's interpolation
// (ADSolver.cu:173-174) with the closed-face rule of the oracle
__device__ __forceinline__ double face_val(unsigned c_lo, unsigned c_hi, double q_lo, double q_hi, double d_lo, double d_hi,
                                           double rc, const double* __restrict__ body, bool& open) {
  if (c_hi != IFX_FLUID) { open = false; return body[c_hi >> 2]; }
  if (c_lo != IFX_FLUID) { open = false; return body[c_lo >> 2]; }
  open = true;
  return rc * fma(q_hi, d_lo, q_lo * d_hi);
}

struct StripCols {               // what a thread knows about its two columns
  int i, ie;                     // first column, clamped index of the column right of the pair
  bool two;                      // second column is an interior column too
  double dxm, dx0, dx1, dx2;     // dx[i-1 .. i+2]
  double rxm, rx0, rx1;          // rcpx[i-1 .. i+1]
  double ydx0, ydx1;             // refined reciprocals of dx0, dx1
};

__device__ __forceinline__ StripCols strip_cols(const Layout& L, const Metrics& M, int i) {
  StripCols s;
  s.i = i; s.ie = min(i + 2, L.nx - 1);
  s.two = (i + 1 <= L.nx - 2);
  s.dxm = M.dx[i - 1]; s.dx0 = M.dx[i]; s.dx1 = M.dx[i + 1]; s.dx2 = M.dx[s.ie];
  s.rxm = M.rcpx[i - 1]; s.rx0 = M.rcpx[i]; s.rx1 = M.rcpx[min(i + 1, L.nx - 2)];
  s.ydx0 = rcp_refined(s.dx0); s.ydx1 = rcp_refined(s.dx1);
  return s;
}

struct StripRows {               // per-block row tables: dy[j0-1 ..], rcpy[j0-1 ..], refined reciprocal of dy[j0 ..]
  double dy[HRY + 2], rcpy[HRY + 1], ydy[HRY];
};

__device__ __forceinline__ void strip_rows(StripRows& t, const Metrics& M, int j0, int nrow) {
  for (int r = threadIdx.x; r < nrow + 2; r += HT) {
    t.dy[r] = M.dy[j0 - 1 + r];
    if (r < nrow + 1) t.rcpy[r] = M.rcpy[j0 - 1 + r];
    if (r < nrow) t.ydy[r] = rcp_refined(M.dy[j0 + r]);
  }
  __syncthreads();
}

__device__ __forceinline__ uchar2 ld_ct2(const uint8_t* __restrict__ ct, size_t o) {
  return *reinterpret_cast<const uchar2*>(ct + o);
}
__device__ __forceinline__ double2 ld_d2(const double* __restrict__ q, size_t o) {
  return *reinterpret_cast<const double2*>(q + o);
}

// ---------------------------------------------------------------------------------------------
// a15: rhs = ((uf_e - uf_w)/dx_i + (vf_n - vf_s)/dy_j)/dt on fluid cells, 0 elsewhere
// ---------------------------------------------------------------------------------------------
static __global__ void __launch_bounds__(HT)
k_ppe_rhs(FaceCtx c, const double* __restrict__ u, const double* __restrict__ v, double* __restrict__ rhs) {
  const Layout& L = c.L;
  __shared__ StripRows t;
  const int j0 = L.jb + blockIdx.y * HRY;
  const int nrow = min(HRY, L.je - j0);
  strip_rows(t, c.M, j0, nrow);
  const int i = 1 + 2 * (blockIdx.x * HT + threadIdx.x);
  if (i > L.nx - 2) return;
  const StripCols sc = strip_cols(L, c.M, i);
  const double dt = c.M.dt, ydt = rcp_refined(dt);

  size_t o = lidx(L, i, j0 - L.j0);
  double2 vC = ld_d2(v, o);
  uchar2 ctC = ld_ct2(c.ct, o);
  double vfs[2];
  {   // south faces of the first row
    const double2 vS = ld_d2(v, o - L.pitch);
    const uchar2 ctS = ld_ct2(c.ct, o - L.pitch);
    bool open;
    vfs[0] = face_val(ctS.x, ctC.x, vS.x, vC.x, t.dy[0], t.dy[1], t.rcpy[0], c.vb, open);
    vfs[1] = face_val(ctS.y, ctC.y, vS.y, vC.y, t.dy[0], t.dy[1], t.rcpy[0], c.vb, open);
  }
  for (int r = 0; r < nrow; ++r, o += L.pitch) {
    const double2 vN = ld_d2(v, o + L.pitch);
    const uchar2 ctN = ld_ct2(c.ct, o + L.pitch);
    const double2 uc = ld_d2(u, o);
    const double uW = u[o - 1], uE = u[o + (sc.ie - i)];
    const unsigned ctW = c.ct[o - 1], ctE = c.ct[o + (sc.ie - i)];
    const double dyj = t.dy[r + 1], dyn = t.dy[r + 2], ryj = t.rcpy[r + 1], ydy = t.ydy[r];
    bool open;
    const double fw = face_val(ctW, ctC.x, uW, uc.x, sc.dxm, sc.dx0, sc.rxm, c.ub, open);
    const double fm = face_val(ctC.x, ctC.y, uc.x, uc.y, sc.dx0, sc.dx1, sc.rx0, c.ub, open);
    const double fe = face_val(ctC.y, ctE, uc.y, uE, sc.dx1, sc.dx2, sc.rx1, c.ub, open);
    const double vn0 = face_val(ctC.x, ctN.x, vC.x, vN.x, dyj, dyn, ryj, c.vb, open);
    const double vn1 = face_val(ctC.y, ctN.y, vC.y, vN.y, dyj, dyn, ryj, c.vb, open);
    const double nx0 = fm - fw, nx1 = fe - fm, ny0 = vn0 - vfs[0], ny1 = vn1 - vfs[1];
    bool ok = true;
    const double s0 = div_checked(nx0, sc.dx0, sc.ydx0, ok) + div_checked(ny0, dyj, ydy, ok);
    const double s1 = div_checked(nx1, sc.dx1, sc.ydx1, ok) + div_checked(ny1, dyj, ydy, ok);
    double r0 = div_checked(s0, dt, ydt, ok), r1 = div_checked(s1, dt, ydt, ok);
    if (!ok) {      // an operand outside the fast path's range: plain IEEE divisions, same results by definition
      r0 = (nx0 / sc.dx0 + ny0 / dyj) / dt;
      r1 = (nx1 / sc.dx1 + ny1 / dyj) / dt;
    }
    if (ctC.x != IFX_FLUID) r0 = 0.0;
    if (ctC.y != IFX_FLUID) r1 = 0.0;
    if (sc.two) *reinterpret_cast<double2*>(rhs + o) = make_double2(r0, r1);
    else rhs[o] = r0;
    vfs[0] = vn0; vfs[1] = vn1;
    vC = vN; ctC = ctN;
  }
}

// ---------------------------------------------------------------------------------------------
// a18: projection (oracle: orc_correct).  Reads u*, v*, p; writes u, v into the partner buffers and the face
// velocities uf (east faces; column 1 also its west face), vf (north faces; the slab's first row also its south).
// ---------------------------------------------------------------------------------------------
static __global__ void __launch_bounds__(HT)
k_correct(FaceCtx c, const double* __restrict__ us, const double* __restrict__ vs, const double* __restrict__ p,
          double* __restrict__ un, double* __restrict__ vn, double* __restrict__ uf, double* __restrict__ vf) {
  const Layout& L = c.L;
  __shared__ StripRows t;
  const int j0 = L.jb + blockIdx.y * HRY;
  const int nrow = min(HRY, L.je - j0);
  strip_rows(t, c.M, j0, nrow);
  const int i = 1 + 2 * (blockIdx.x * HT + threadIdx.x);
  if (i > L.nx - 2) return;
  const StripCols sc = strip_cols(L, c.M, i);
  const int nxm2 = L.nx - 2, nym2 = L.ny - 2;
  const double dt = c.M.dt;
  const int de = sc.ie - i;                       // offset of the column right of my pair (clamped to the ring)

  size_t o = lidx(L, i, j0 - L.j0);
  double2 pS = ld_d2(p, o - L.pitch), pC = ld_d2(p, o), vC = ld_d2(vs, o);
  uchar2 ctS = ld_ct2(c.ct, o - L.pitch), ctC = ld_ct2(c.ct, o);
  if (j0 == L.jb) {   // south faces of the slab's first row (oracle: every face has one owner; here the row above it)
    const double2 vS = ld_d2(vs, o - L.pitch);
    bool open;
    double f0 = face_val(ctS.x, ctC.x, vS.x, vC.x, t.dy[0], t.dy[1], t.rcpy[0], c.vb, open);
    if (open && j0 - 1 >= 1) f0 = f0 - dt * ((pC.x - pS.x) * (2.0 * t.rcpy[0]));
    double f1 = face_val(ctS.y, ctC.y, vS.y, vC.y, t.dy[0], t.dy[1], t.rcpy[0], c.vb, open);
    if (open && j0 - 1 >= 1) f1 = f1 - dt * ((pC.y - pS.y) * (2.0 * t.rcpy[0]));
    if (sc.two) *reinterpret_cast<double2*>(vf + o - L.pitch) = make_double2(f0, f1);
    else vf[o - L.pitch] = f0;
  }
  for (int r = 0; r < nrow; ++r, o += L.pitch) {
    const int j = j0 + r;
    const double2 pN = ld_d2(p, o + L.pitch), vN = ld_d2(vs, o + L.pitch), uc = ld_d2(us, o);
    const uchar2 ctN = ld_ct2(c.ct, o + L.pitch);
    const double uW = us[o - 1], uE = us[o + de], pW = p[o - 1], pE = p[o + de];
    const unsigned ctW = c.ct[o - 1], ctE = c.ct[o + de];
    const double dym = t.dy[r], dyj = t.dy[r + 1], dyn = t.dy[r + 2], rym = t.rcpy[r], ryj = t.rcpy[r + 1], ydy = t.ydy[r];
    const bool bot = (j == 1), top = (j == nym2);

    double uo0 = uc.x, uo1 = uc.y, vo0 = vC.x, vo1 = vC.y;
    {
      // cell (i, j)
      const double pc = pC.x;
      const double pWv = (i == 1 || ctW != IFX_FLUID) ? pc : pW;
      const double pEv = (i == nxm2 || ctC.y != IFX_FLUID) ? pc : pC.y;
      const double pSv = (bot || ctS.x != IFX_FLUID) ? pc : pS.x;
      const double pNv = (top || ctN.x != IFX_FLUID) ? pc : pN.x;
      const double pe = sc.rx0 * fma(pEv, sc.dx0, pc * sc.dx1);
      const double pw = sc.rxm * fma(pc, sc.dxm, pWv * sc.dx0);
      const double pn = ryj * fma(pNv, dyj, pc * dyn);
      const double ps = rym * fma(pc, dym, pSv * dyj);
      const double gx = pe - pw, gy = pn - ps;
      bool ok = true;
      double qx = div_checked(gx, sc.dx0, sc.ydx0, ok), qy = div_checked(gy, dyj, ydy, ok);
      if (!ok) { qx = gx / sc.dx0; qy = gy / dyj; }
      if (ctC.x == IFX_FLUID) { uo0 = uc.x - dt * qx; vo0 = vC.x - dt * qy; }
    }
    {
      // cell (i+1, j)
      const double pc = pC.y;
      const double pWv = (ctC.x != IFX_FLUID) ? pc : pC.x;                          // i+1 >= 2: never the first column
      const double pEv = (i + 1 == nxm2 || ctE != IFX_FLUID) ? pc : pE;
      const double pSv = (bot || ctS.y != IFX_FLUID) ? pc : pS.y;
      const double pNv = (top || ctN.y != IFX_FLUID) ? pc : pN.y;
      const double pe = sc.rx1 * fma(pEv, sc.dx1, pc * sc.dx2);
      const double pw = sc.rx0 * fma(pc, sc.dx0, pWv * sc.dx1);
      const double pn = ryj * fma(pNv, dyj, pc * dyn);
      const double ps = rym * fma(pc, dym, pSv * dyj);
      const double gx = pe - pw, gy = pn - ps;
      bool ok = true;
      double qx = div_checked(gx, sc.dx1, sc.ydx1, ok), qy = div_checked(gy, dyj, ydy, ok);
      if (!ok) { qx = gx / sc.dx1; qy = gy / dyj; }
      if (ctC.y == IFX_FLUID) { uo1 = uc.y - dt * qx; vo1 = vC.y - dt * qy; }
    }
    // faces owned by my cells: east and north; column 1 also owns its west face (grid boundary: never corrected)
    bool open;
    double fu0 = face_val(ctC.x, ctC.y, uc.x, uc.y, sc.dx0, sc.dx1, sc.rx0, c.ub, open);
    if (open && i <= L.nx - 3) fu0 = fu0 - dt * ((pC.y - pC.x) * (2.0 * sc.rx0));
    double fu1 = face_val(ctC.y, ctE, uc.y, uE, sc.dx1, sc.dx2, sc.rx1, c.ub, open);
    if (open && i + 1 <= L.nx - 3) fu1 = fu1 - dt * ((pE - pC.y) * (2.0 * sc.rx1));
    double fv0 = face_val(ctC.x, ctN.x, vC.x, vN.x, dyj, dyn, ryj, c.vb, open);
    if (open && j <= L.ny - 3) fv0 = fv0 - dt * ((pN.x - pC.x) * (2.0 * ryj));
    double fv1 = face_val(ctC.y, ctN.y, vC.y, vN.y, dyj, dyn, ryj, c.vb, open);
    if (open && j <= L.ny - 3) fv1 = fv1 - dt * ((pN.y - pC.y) * (2.0 * ryj));
    if (sc.two) {
      *reinterpret_cast<double2*>(un + o) = make_double2(uo0, uo1);
      *reinterpret_cast<double2*>(vn + o) = make_double2(vo0, vo1);
      *reinterpret_cast<double2*>(uf + o) = make_double2(fu0, fu1);
      *reinterpret_cast<double2*>(vf + o) = make_double2(fv0, fv1);
    } else {
      un[o] = uo0; vn[o] = vo0; uf[o] = fu0; vf[o] = fv0;
    }
    if (i == 1) uf[o - 1] = face_val(ctW, ctC.x, uW, uc.x, sc.dxm, sc.dx0, sc.rxm, c.ub, open);
    pS = pC; pC = pN; vC = vN; ctS = ctC; ctC = ctN;
  }
}

// ---------------------------------------------------------------------------------------------
// launchers
// ---------------------------------------------------------------------------------------------
cudaError_t launch_apply_ring(const Layout& L, double* q, const double* two_bc /*W,E,S,N or null*/, int neumann,
                              cudaStream_t st) {
  const int n1 = L.nyl > L.nx ? L.nyl : L.nx;
  const double z[4] = {0, 0, 0, 0};
  const double* b = two_bc ? two_bc : z;
  k_apply_ring<<<(n1 + 127) / 128, 128, 0, st>>>(L, q, b[0], b[1], b[2], b[3], neumann);
  return cudaGetLastError();
}

static FaceCtx make_ctx(const Layout& L, const Metrics& M, const uint8_t* ct, const double* ub, const double* vb) {
  FaceCtx c; c.L = L; c.M = M; c.ct = ct; c.ub = ub; c.vb = vb; return c;
}

cudaError_t launch_faces_init(const Layout& L, const Metrics& M, const uint8_t* ct, const double* ub, const double* vb,
                              const double* u, const double* v, double* uf, double* vf, cudaStream_t st) {
  dim3 g((L.nx - 1 + 127) / 128, L.nyl - 1);
  k_faces_init<<<g, 128, 0, st>>>(make_ctx(L, M, ct, ub, vb), u, v, uf, vf);
  return cudaGetLastError();
}

cudaError_t launch_ppe_rhs(const Layout& L, const Metrics& M, const uint8_t* ct, const double* ub, const double* vb,
                           const double* u, const double* v, double* rhs, cudaStream_t st) {
  dim3 g((L.nx - 2 + 2 * HT - 1) / (2 * HT), (L.je - L.jb + HRY - 1) / HRY);
  k_ppe_rhs<<<g, HT, 0, st>>>(make_ctx(L, M, ct, ub, vb), u, v, rhs);
  return cudaGetLastError();
}

cudaError_t launch_correct(const Layout& L, const Metrics& M, const uint8_t* ct, const double* ub, const double* vb,
                           const double* us, const double* vs, const double* p, double* un, double* vn, double* uf,
                           double* vf, cudaStream_t st) {
  dim3 g((L.nx - 2 + 2 * HT - 1) / (2 * HT), (L.je - L.jb + HRY - 1) / HRY);
  k_correct<<<g, HT, 0, st>>>(make_ctx(L, M, ct, ub, vb), us, vs, p, un, vn, uf, vf);
  return cudaGetLastError();
}

}  // namespace ifx
