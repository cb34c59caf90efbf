// kernels_full.cu — the stages of the fractional step the reference leaves empty: PPE source term
// (calculateSourcePPE is a stub, PPESolver.cu:106-135), projection (AD_PPE_Correction.cu:1-12 is empty), and the
// boundary-condition refresh for arbitrary BC values (the reference hard-codes them, ADSolver.cu:200-216).
// Semantics: oracle/ifx_oracle_full.c (PARITY UNPINNED), arithmetic identical to it operation for operation.
//
// HBM traffic per cell (algorithmic): source term 16 B read (u*, v*) + 8 B write (+1 B cell type);
// projection 24 B read (u*, v*, p) + 32 B write (u, v, uf, vf) (+1 B): face velocities are produced in the same
// pass that corrects the cell velocities — they are next step's convecting velocities.
#include "kernels.cuh"

namespace ifx {

constexpr int ROWS_PER_BLOCK = 8;

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

// ---------------------------------------------------------------------------------------------
// a15: rhs = ((uf_e - uf_w)/dx_i + (vf_n - vf_s)/dy_j)/dt on fluid cells, 0 elsewhere
// ---------------------------------------------------------------------------------------------
static __global__ void k_ppe_rhs(FaceCtx c, const double* __restrict__ u, const double* __restrict__ v,
                                 double* __restrict__ rhs) {
  const Layout& L = c.L;
  const int i = 1 + blockIdx.x * blockDim.x + threadIdx.x;
  if (i > L.nx - 2) return;
  const int j0 = L.jb + blockIdx.y * ROWS_PER_BLOCK;
#pragma unroll 2
  for (int j = j0; j < min(j0 + ROWS_PER_BLOCK, L.je); ++j) {     // consecutive rows: the row read as "north" is reused
    const int jl = j - L.j0;
    const size_t o = lidx(L, i, jl);
    double r = 0.0;
    if (c.ct[o] == IFX_FLUID) {
      bool open;
      const double ufe = face_u(c, u, i, jl, &open), ufw = face_u(c, u, i - 1, jl, &open);
      const double vfn = face_v(c, v, i, jl, j, &open), vfs = face_v(c, v, i, jl - 1, j - 1, &open);
      r = ((ufe - ufw) / c.M.dx[i] + (vfn - vfs) / c.M.dy[j]) / c.M.dt;
    }
    rhs[o] = r;
  }
}

// ---------------------------------------------------------------------------------------------
// a18: projection (oracle: orc_correct).  Reads u*, v*, p; writes u, v into the partner buffers and uf, vf.
// ---------------------------------------------------------------------------------------------
static __global__ void k_correct(FaceCtx c, const double* __restrict__ us, const double* __restrict__ vs,
                                 const double* __restrict__ p, double* __restrict__ un, double* __restrict__ vn,
                                 double* __restrict__ uf, double* __restrict__ vf) {
  const Layout& L = c.L;
  const int i = 1 + blockIdx.x * blockDim.x + threadIdx.x;
  if (i > L.nx - 2) return;
  const int jstart = L.jb + blockIdx.y * ROWS_PER_BLOCK;
  const int nxm2 = L.nx - 2, nym2 = L.ny - 2;
  const double dt = c.M.dt;
#pragma unroll 2
  for (int j = jstart; j < min(jstart + ROWS_PER_BLOCK, L.je); ++j) {
    const int jl = j - L.j0;
    const size_t o = lidx(L, i, jl);
    const double pc = p[o];
    double u_out = us[o], v_out = vs[o];
    if (c.ct[o] == IFX_FLUID) {
      const double dx_i = c.M.dx[i], dx_ip1 = c.M.dx[i + 1], dx_im1 = c.M.dx[i - 1];
      const double dy_j = c.M.dy[j], dy_jp1 = c.M.dy[j + 1], dy_jm1 = c.M.dy[j - 1];
      const double pW = (i == 1 || c.ct[o - 1] != IFX_FLUID) ? pc : p[o - 1];
      const double pE = (i == nxm2 || c.ct[o + 1] != IFX_FLUID) ? pc : p[o + 1];
      const double pS = (j == 1 || c.ct[o - L.pitch] != IFX_FLUID) ? pc : p[o - L.pitch];
      const double pN = (j == nym2 || c.ct[o + L.pitch] != IFX_FLUID) ? pc : p[o + L.pitch];
      const double pe = c.M.rcpx[i] * fma(pE, dx_i, pc * dx_ip1);
      const double pw = c.M.rcpx[i - 1] * fma(pc, dx_im1, pW * dx_i);
      const double pn = c.M.rcpy[j] * fma(pN, dy_j, pc * dy_jp1);
      const double ps = c.M.rcpy[j - 1] * fma(pc, dy_jm1, pS * dy_j);
      u_out = us[o] - dt * ((pe - pw) / dx_i);
      v_out = vs[o] - dt * ((pn - ps) / dy_j);
    }
    un[o] = u_out;
    vn[o] = v_out;

    // faces owned by this cell: east and north; the first column / first owned row also own their west / south face
    bool open;
    {
      double val = face_u(c, us, i, jl, &open);
      if (open && i <= L.nx - 3) val = val - dt * ((p[o + 1] - pc) * (2.0 * c.M.rcpx[i]));
      uf[o] = val;
      if (i == 1) uf[o - 1] = face_u(c, us, 0, jl, &open);          // grid-boundary face: never corrected
    }
    {
      double val = face_v(c, vs, i, jl, j, &open);
      if (open && j <= L.ny - 3) val = val - dt * ((p[o + L.pitch] - pc) * (2.0 * c.M.rcpy[j]));
      vf[o] = val;
      if (j == L.jb) {
        double vs_ = face_v(c, vs, i, jl - 1, j - 1, &open);
        if (open && j - 1 >= 1) vs_ = vs_ - dt * ((pc - p[o - L.pitch]) * (2.0 * c.M.rcpy[j - 1]));
        vf[o - L.pitch] = vs_;
      }
    }
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
  dim3 g((L.nx - 2 + 255) / 256, (L.je - L.jb + ROWS_PER_BLOCK - 1) / ROWS_PER_BLOCK);
  k_ppe_rhs<<<g, 256, 0, st>>>(make_ctx(L, M, ct, ub, vb), u, v, rhs);
  return cudaGetLastError();
}

cudaError_t launch_correct(const Layout& L, const Metrics& M, const uint8_t* ct, const double* ub, const double* vb,
                           const double* us, const double* vs, const double* p, double* un, double* vn, double* uf,
                           double* vf, cudaStream_t st) {
  dim3 g((L.nx - 2 + 255) / 256, (L.je - L.jb + ROWS_PER_BLOCK - 1) / ROWS_PER_BLOCK);
  k_correct<<<g, 256, 0, st>>>(make_ctx(L, M, ct, ub, vb), us, vs, p, un, vn, uf, vf);
  return cudaGetLastError();
}

}  // namespace ifx
