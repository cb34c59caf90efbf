// kernels_ib.cu — sharp-interface immersed boundary: cell classification (iBlank), ghost-cell list, body
// intercepts, image points and interpolation stencils.
//
// The reference has only a stub here (iBlankComputeKernel assigns 1.0 to every cell, preSim.cu:123-133; no
// ghost-cell code exists), so the semantics are those of oracle/ifx_oracle_full.c — PARITY UNPINNED — and
// the arithmetic is plain IEEE +,-,*,/ in the oracle's order (this TU is compiled with -fmad=false), which
// makes the integer maps AND the weights bit-identical between the CPU oracle and this path.
//
// Cell type byte: low two bits 0 solid / 1 fluid / 2 ghost cell, upper six bits = owning body (<= 63 bodies).
#include "kernels.cuh"

namespace ifx {

__device__ __forceinline__ bool point_in_polygon(double x, double y, const double* __restrict__ xm,
                                                 const double* __restrict__ ym, int n) {
  bool inside = false;
  for (int k = 0; k < n; k++) {
    const int k2 = (k + 1 == n) ? 0 : k + 1;
    const double xa = xm[k], ya = ym[k], xb = xm[k2], yb = ym[k2];
    if ((ya > y) != (yb > y)) {
      const double xi = xa + (y - ya) * (xb - xa) / (yb - ya);
      if (x < xi) inside = !inside;
    }
  }
  return inside;
}

// a16: SOLID when the cell centre is inside a body polygon; the ghost ring of the grid stays FLUID.
// bbox = per-body [xmin, xmax, ymin, ymax]: a centre outside the box cannot be inside the polygon, so skipping the
// crossing test there does not change the result.
static __global__ void k_classify(Layout L, const double* __restrict__ xc, const double* __restrict__ yc, int nbodies,
                                  const int* __restrict__ off, const double* __restrict__ xm,
                                  const double* __restrict__ ym, const double* __restrict__ bbox,
                                  uint8_t* __restrict__ celltype) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int jl = blockIdx.y;
  if (i >= L.nx || jl >= L.nyl) return;
  const int j = L.j0 + jl;
  uint8_t t = IFX_FLUID;
  if (i > 0 && i < L.nx - 1 && j > 0 && j < L.ny - 1) {
    const double x = xc[i], y = yc[j];
    for (int b = 0; b < nbodies; b++) {
      if (x < bbox[4 * b] || x > bbox[4 * b + 1] || y < bbox[4 * b + 2] || y > bbox[4 * b + 3]) continue;
      if (point_in_polygon(x, y, xm + off[b], ym + off[b], off[b + 1] - off[b])) { t = (uint8_t)(b << 2); break; }
    }
  }
  celltype[lidx(L, i, jl)] = t;
}

// ghost cell = non-fluid cell with a fluid 4-neighbour.  In place: only 0 -> 2 transitions in the type bits,
// and the test looks for == FLUID, so concurrent marking cannot change anybody's answer.
// (slab note: rows jb-1 and je are halo rows classified above with the same rule, so neighbours are complete.)
static __global__ void k_mark_ghost(Layout L, uint8_t* __restrict__ celltype) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int jl = blockIdx.y;
  if (i >= L.nx || jl >= L.nyl) return;
  const int j = L.j0 + jl;
  if (!(i > 0 && i < L.nx - 1 && j > 0 && j < L.ny - 1)) return;
  if (jl == 0 || jl == L.nyl - 1) return;       // halo rows of a slab: owned by the neighbour rank
  const size_t o = lidx(L, i, jl);
  const uint8_t c = celltype[o];
  if (c == IFX_FLUID || (c & 3) == IFX_GHOST) return;
  if (celltype[o - 1] == IFX_FLUID || celltype[o + 1] == IFX_FLUID || celltype[o - L.pitch] == IFX_FLUID ||
      celltype[o + L.pitch] == IFX_FLUID)
    celltype[o] = c | IFX_GHOST;
}

// ordered compaction, pass 1: ghost cells per owned row
static __global__ void k_gc_count_rows(Layout L, const uint8_t* __restrict__ celltype, int* __restrict__ rowcount) {
  const int jl = 1 + blockIdx.x;                 // owned local rows 1 .. nyl-2
  int n = 0;
  for (int i = 1 + threadIdx.x; i < L.nx - 1; i += blockDim.x) n += ((celltype[lidx(L, i, jl)] & 3) == IFX_GHOST);
  __shared__ int sh[32];
  for (int off = 16; off > 0; off >>= 1) n += __shfl_down_sync(0xffffffffu, n, off);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = n;
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
    for (int w = 0; w < (int)(blockDim.x >> 5); w++) t += sh[w];
    rowcount[blockIdx.x] = t;
  }
}

// pass 2: exclusive scan of the row counts (one block; nrows <= a few 10^4)
static __global__ void k_gc_scan_rows(int nrows, const int* __restrict__ rowcount, int* __restrict__ rowstart,
                                      int* __restrict__ total) {
  __shared__ int carry;
  __shared__ int sh[1024];
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int base = 0; base < nrows; base += blockDim.x) {
    const int r = base + threadIdx.x;
    const int v = (r < nrows) ? rowcount[r] : 0;
    sh[threadIdx.x] = v;
    __syncthreads();
    for (int off = 1; off < (int)blockDim.x; off <<= 1) {      // Hillis-Steele inclusive scan
      const int t = (threadIdx.x >= (unsigned)off) ? sh[threadIdx.x - off] : 0;
      __syncthreads();
      sh[threadIdx.x] += t;
      __syncthreads();
    }
    if (r < nrows) rowstart[r] = carry + sh[threadIdx.x] - v;
    __syncthreads();
    if (threadIdx.x == blockDim.x - 1) carry += sh[threadIdx.x];
    __syncthreads();
  }
  if (threadIdx.x == 0) { rowstart[nrows] = carry; *total = carry; }
}

// pass 3: write the ghost cells of each row in increasing i (=> the whole list in increasing reference id)
static __global__ void k_gc_fill(Layout L, const uint8_t* __restrict__ celltype, const int* __restrict__ rowstart,
                                 int capacity, int* __restrict__ cell, int* __restrict__ ref_id, int* __restrict__ body) {
  const int jl = 1 + blockIdx.x;
  const int j = L.j0 + jl;
  __shared__ int warp_tot[32];
  __shared__ int base;
  if (threadIdx.x == 0) base = rowstart[blockIdx.x];
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int i0 = 1; i0 < L.nx - 1; i0 += blockDim.x) {
    const int i = i0 + threadIdx.x;
    const uint8_t c = (i < L.nx - 1) ? celltype[lidx(L, i, jl)] : (uint8_t)IFX_FLUID;
    const bool g = (c & 3) == IFX_GHOST;
    const unsigned m = __ballot_sync(0xffffffffu, g);
    if (lane == 0) warp_tot[warp] = __popc(m);
    __syncthreads();
    int before = 0;
    for (int w = 0; w < warp; w++) before += warp_tot[w];
    int chunk = 0;
    for (int w = 0; w < nw; w++) chunk += warp_tot[w];
    if (g) {
      const int k = base + before + __popc(m & ((1u << lane) - 1u));
      if (k < capacity) { cell[k] = (int)lidx(L, i, jl); ref_id[k] = i + j * L.nx; body[k] = c >> 2; }
    }
    __syncthreads();
    if (threadIdx.x == 0) base += chunk;
    __syncthreads();
  }
}

__device__ __forceinline__ int lower_index(const double* __restrict__ c, int n, double x) {
  int lo = 0, hi = n - 2;
  if (x < c[0]) return 0;
  while (lo < hi) {
    const int mid = (lo + hi + 1) / 2;
    if (c[mid] <= x) lo = mid; else hi = mid - 1;
  }
  return lo;
}

// a17: one thread per ghost cell — body intercept, image point, enclosing box, weights (oracle: orc_ghost_cells)
static __global__ void k_gc_geometry(Layout L, const double* __restrict__ xc, const double* __restrict__ yc,
                                     const int* __restrict__ off, const double* __restrict__ xm,
                                     const double* __restrict__ ym, const uint8_t* __restrict__ celltype, int ngc,
                                     const int* __restrict__ ref_id, const int* __restrict__ body,
                                     int* __restrict__ stencil, int* __restrict__ stencil_ref, double* __restrict__ wd,
                                     double* __restrict__ wn, double* __restrict__ bi, double* __restrict__ ip) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= ngc) return;
  const int id = ref_id[g];
  const int i = id % L.nx, j = id / L.nx;
  const int b = body[g];
  const double xg = xc[i], yg = yc[j];
  // closest point on the body's boundary (first minimum in edge order)
  const double* bxm = xm + off[b];
  const double* bym = ym + off[b];
  const int n = off[b + 1] - off[b];
  double best = INFINITY, bx = xg, by = yg;
  for (int k = 0; k < n; k++) {
    const int k2 = (k + 1 == n) ? 0 : k + 1;
    const double ax = bxm[k], ay = bym[k], ex = bxm[k2] - ax, ey = bym[k2] - ay;
    const double len2 = ex * ex + ey * ey;
    double t = 0.0;
    if (len2 > 0.0) {
      t = ((xg - ax) * ex + (yg - ay) * ey) / len2;
      if (t < 0.0) t = 0.0;
      if (t > 1.0) t = 1.0;
    }
    const double px = ax + t * ex, py = ay + t * ey;
    const double d2 = (xg - px) * (xg - px) + (yg - py) * (yg - py);
    if (d2 < best) { best = d2; bx = px; by = py; }
  }
  const double xi = xg + 2.0 * (bx - xg), yi = yg + 2.0 * (by - yg);
  const int i0 = lower_index(xc, L.nx, xi), j0 = lower_index(yc, L.ny, yi);
  const double a = (xi - xc[i0]) / (xc[i0 + 1] - xc[i0]);
  const double bb = (yi - yc[j0]) / (yc[j0 + 1] - yc[j0]);
  const int ni[4] = {i0, i0 + 1, i0, i0 + 1}, nj[4] = {j0, j0, j0 + 1, j0 + 1};
  double w[4] = {(1.0 - a) * (1.0 - bb), a * (1.0 - bb), (1.0 - a) * bb, a * bb};
  double W = 0.0, ws = 0.0;
  int nid[4];
  for (int m = 0; m < 4; m++) {
    nid[m] = ni[m] + nj[m] * L.nx;
    // the stencil may reach one row into the neighbour slab: rows outside [j0, j0+nyl) are clamped
    int jl = nj[m] - L.j0;
    jl = jl < 0 ? 0 : (jl > L.nyl - 1 ? L.nyl - 1 : jl);
    const uint8_t c = celltype[lidx(L, ni[m], jl)];
    if (nid[m] != id && (c & 3) == IFX_SOLID) w[m] = 0.0;
    W = W + w[m];
  }
  int kept = 0;
  for (int m = 0; m < 4; m++) {
    w[m] = (W > 0.0) ? w[m] / W : 0.0;
    if (nid[m] == id) { ws = w[m]; w[m] = 0.0; }
    else if (w[m] != 0.0) kept++;
  }
  bi[2 * g] = bx; bi[2 * g + 1] = by;
  ip[2 * g] = xi; ip[2 * g + 1] = yi;
  for (int m = 0; m < 4; m++) {
    stencil_ref[4 * g + m] = nid[m];
    int jl = nj[m] - L.j0;
    jl = jl < 0 ? 0 : (jl > L.nyl - 1 ? L.nyl - 1 : jl);
    stencil[4 * g + m] = (int)lidx(L, ni[m], jl);
  }
  if (kept == 0 || (1.0 - ws) < 1e-12) {
    for (int m = 0; m < 4; m++) { wd[5 * g + m] = 0.0; wn[4 * g + m] = 0.0; }
    wd[5 * g + 4] = 1.0;
  } else {
    for (int m = 0; m < 4; m++) {
      wd[5 * g + m] = -w[m] / (1.0 + ws);
      wn[4 * g + m] = w[m] / (1.0 - ws);
    }
    wd[5 * g + 4] = 2.0 / (1.0 + ws);
  }
}

// ---------------------------------------------------------------------------------------------
// ghost-cell values.  Dirichlet (u, v): phi = cd*phi_BI + sum wd[m]*phi_m;  Neumann (p): phi = sum wn[m]*phi_m.
// dst may be another buffer (in-loop, Jacobi-lagged) or a gather array (in-place refresh = eval + scatter).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ double gc_dirichlet(const double* __restrict__ q, const int* __restrict__ st,
                                               const double* __restrict__ wd, double phi_bi) {
  double t = wd[4] * phi_bi;
  t = fma(wd[0], q[st[0]], t);
  t = fma(wd[1], q[st[1]], t);
  t = fma(wd[2], q[st[2]], t);
  t = fma(wd[3], q[st[3]], t);
  return t;
}

static __global__ void k_gc_velocity(int ngc, const int* __restrict__ cell, const int* __restrict__ stencil,
                                     const double* __restrict__ wd, const int* __restrict__ body,
                                     const double* __restrict__ ub, const double* __restrict__ vb,
                                     const double* __restrict__ usrc, const double* __restrict__ vsrc,
                                     double* __restrict__ udst, double* __restrict__ vdst, int gather,
                                     const LoopCtl* ctl, int iter) {
  // in-loop use: iteration `iter` ran iff the loop was not already finished by an earlier iteration
  if (ctl && ctl->done && ctl->iter < iter) return;
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= ngc) return;
  const double pu = gc_dirichlet(usrc, stencil + 4 * g, wd + 5 * g, ub[body[g]]);
  const double pv = gc_dirichlet(vsrc, stencil + 4 * g, wd + 5 * g, vb[body[g]]);
  const int o = gather ? g : cell[g];
  udst[o] = pu; vdst[o] = pv;
}

static __global__ void k_gc_pressure(int ngc, const int* __restrict__ cell, const int* __restrict__ stencil,
                                     const double* __restrict__ wn, const double* __restrict__ psrc,
                                     double* __restrict__ pdst, int gather) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= ngc) return;
  const int* st = stencil + 4 * g;
  const double* w = wn + 4 * g;
  double t = w[0] * psrc[st[0]];
  t = fma(w[1], psrc[st[1]], t);
  t = fma(w[2], psrc[st[2]], t);
  t = fma(w[3], psrc[st[3]], t);
  pdst[gather ? g : cell[g]] = t;
}

static __global__ void k_gc_scatter(int ngc, const int* __restrict__ cell, const double* __restrict__ a,
                                    double* __restrict__ qa, const double* __restrict__ b, double* __restrict__ qb) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= ngc) return;
  qa[cell[g]] = a[g];
  if (b) qb[cell[g]] = b[g];
}

// ---------------------------------------------------------------------------------------------
// launchers
// ---------------------------------------------------------------------------------------------
cudaError_t launch_classify(const Layout& L, const double* xc, const double* yc, int nbodies, const int* off,
                            const double* xm, const double* ym, const double* bbox, uint8_t* celltype, cudaStream_t st) {
  dim3 g((L.nx + 127) / 128, L.nyl);
  k_classify<<<g, 128, 0, st>>>(L, xc, yc, nbodies, off, xm, ym, bbox, celltype);
  k_mark_ghost<<<g, 128, 0, st>>>(L, celltype);
  return cudaGetLastError();
}

cudaError_t launch_gc_count(const Layout& L, const uint8_t* celltype, int* rowcount, int* rowstart, int* total,
                            cudaStream_t st) {
  const int nrows = L.nyl - 2;
  k_gc_count_rows<<<nrows, 256, 0, st>>>(L, celltype, rowcount);
  k_gc_scan_rows<<<1, 1024, 0, st>>>(nrows, rowcount, rowstart, total);
  return cudaGetLastError();
}

cudaError_t launch_gc_build(const Layout& L, const double* xc, const double* yc, const int* off, const double* xm,
                            const double* ym, const uint8_t* celltype, const int* rowstart, int ngc, int* cell, int* ref_id,
                            int* body, int* stencil, int* stencil_ref, double* wd, double* wn, double* bi, double* ip,
                            cudaStream_t st) {
  if (ngc <= 0) return cudaSuccess;
  k_gc_fill<<<L.nyl - 2, 256, 0, st>>>(L, celltype, rowstart, ngc, cell, ref_id, body);
  k_gc_geometry<<<(ngc + 127) / 128, 128, 0, st>>>(L, xc, yc, off, xm, ym, celltype, ngc, ref_id, body, stencil, stencil_ref,
                                                   wd, wn, bi, ip);
  return cudaGetLastError();
}

cudaError_t launch_gc_velocity(int ngc, const int* cell, const int* stencil, const double* wd, const int* body,
                               const double* ub, const double* vb, const double* usrc, const double* vsrc, double* udst,
                               double* vdst, int gather, const LoopCtl* ctl, int iter, cudaStream_t st) {
  if (ngc <= 0) return cudaSuccess;
  k_gc_velocity<<<(ngc + 127) / 128, 128, 0, st>>>(ngc, cell, stencil, wd, body, ub, vb, usrc, vsrc, udst, vdst, gather, ctl,
                                                   iter);
  return cudaGetLastError();
}

cudaError_t launch_gc_pressure(int ngc, const int* cell, const int* stencil, const double* wn, const double* psrc,
                               double* pdst, int gather, cudaStream_t st) {
  if (ngc <= 0) return cudaSuccess;
  k_gc_pressure<<<(ngc + 127) / 128, 128, 0, st>>>(ngc, cell, stencil, wn, psrc, pdst, gather);
  return cudaGetLastError();
}

cudaError_t launch_gc_scatter(int ngc, const int* cell, const double* a, double* qa, const double* b, double* qb,
                              cudaStream_t st) {
  if (ngc <= 0) return cudaSuccess;
  k_gc_scatter<<<(ngc + 127) / 128, 128, 0, st>>>(ngc, cell, a, qa, b, qb);
  return cudaGetLastError();
}

}  // namespace ifx
