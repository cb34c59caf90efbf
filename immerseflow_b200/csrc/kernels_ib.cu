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
// crossing test there does not change the result.  A pure function of (i, j) and the geometry: every rank of a
// slab run can evaluate it for rows it does not store and get the owner's answer.
__device__ __forceinline__ uint8_t classify_cell(const Layout& L, const double* __restrict__ xc,
                                                 const double* __restrict__ yc, const BodySet& B, int i, int j) {
  if (!(i > 0 && i < L.nx - 1 && j > 0 && j < L.ny - 1)) return IFX_FLUID;
  const double x = xc[i], y = yc[j];
  for (int b = 0; b < B.nbodies; b++) {
    if (x < B.bbox[4 * b] || x > B.bbox[4 * b + 1] || y < B.bbox[4 * b + 2] || y > B.bbox[4 * b + 3]) continue;
    if (point_in_polygon(x, y, B.xm + B.off[b], B.ym + B.off[b], B.off[b + 1] - B.off[b])) return (uint8_t)(b << 2);
  }
  return IFX_FLUID;
}

// full cell type (ghost bit included) of a cell outside the rows this rank stores
__device__ uint8_t celltype_of(const Layout& L, const double* __restrict__ xc, const double* __restrict__ yc,
                               const BodySet& B, int i, int j) {
  const uint8_t c = classify_cell(L, xc, yc, B, i, j);
  if (c == IFX_FLUID) return c;
  if (classify_cell(L, xc, yc, B, i - 1, j) == IFX_FLUID || classify_cell(L, xc, yc, B, i + 1, j) == IFX_FLUID ||
      classify_cell(L, xc, yc, B, i, j - 1) == IFX_FLUID || classify_cell(L, xc, yc, B, i, j + 1) == IFX_FLUID)
    return c | IFX_GHOST;
  return c;
}

// ---------------------------------------------------------------------------------------------------------
// Classification, ghost marking and the per-row ghost-cell count in ONE pass that only writes: one block per
// stored row.  The crossing test of point_in_polygon compares x with xi = xa + (y-ya)*(xb-xa)/(yb-ya), which
// does not depend on x: the block evaluates the straddling edges of rows j-1, j, j+1 once (a few entries per
// body), and every cell of the row then answers "first body containing me" and "is a 4-neighbour fluid" by
// counting x < xi over those short lists — the same comparisons on the same doubles as classify_cell, hence the
// same bytes as the cell-by-cell definition (the fallback below, used when a list overflows).
// ---------------------------------------------------------------------------------------------------------
constexpr int XCAP = 768;            // crossings kept per row

struct RowCrossings {
  double xi[3][XCAP];
  unsigned char body[3][XCAP];
  int n[3];
  double lo[3], hi[3];               // min / max xi of a row: outside, every parity is even
};

__device__ __forceinline__ uint8_t classify_from_row(const Layout& L, const double* __restrict__ xc, const BodySet& B,
                                                     const RowCrossings& rc, int r, int jr, int i) {
  if (!(i > 0 && i < L.nx - 1 && jr > 0 && jr < L.ny - 1)) return IFX_FLUID;
  const double x = xc[i];
  if (rc.n[r] == 0 || !(x < rc.hi[r])) return IFX_FLUID;
  unsigned long long odd = 0ull;
  for (int e = 0; e < rc.n[r]; e++)
    if (x < rc.xi[r][e]) odd ^= 1ull << rc.body[r][e];
  while (odd) {
    const int b = __ffsll((long long)odd) - 1;
    if (!(x < B.bbox[4 * b] || x > B.bbox[4 * b + 1])) return (uint8_t)(b << 2);
    odd &= odd - 1;
  }
  return IFX_FLUID;
}

static __global__ void __launch_bounds__(256)
k_classify_rows(Layout L, const double* __restrict__ xc, const double* __restrict__ yc, BodySet B,
                uint8_t* __restrict__ celltype, int* __restrict__ rowcount) {
  __shared__ RowCrossings rc;
  __shared__ int overflow;
  __shared__ int warp_cnt[8];
  const int jl = blockIdx.x;
  const int j = L.j0 + jl;
  if (threadIdx.x < 3) { rc.n[threadIdx.x] = 0; }
  if (threadIdx.x == 0) overflow = 0;
  __syncthreads();
  for (int r = 0; r < 3; r++) {
    const int jr = j - 1 + r;
    if (jr < 1 || jr > L.ny - 2) continue;
    const double y = yc[jr];
    for (int b = 0; b < B.nbodies; b++) {
      if (y < B.bbox[4 * b + 2] || y > B.bbox[4 * b + 3]) continue;         // classify_cell skips the body
      const double* xm = B.xm + B.off[b];
      const double* ym = B.ym + B.off[b];
      const int n = B.off[b + 1] - B.off[b];
      for (int k = threadIdx.x; k < n; k += blockDim.x) {
        const int k2 = (k + 1 == n) ? 0 : k + 1;
        const double xa = xm[k], ya = ym[k], xb = xm[k2], yb = ym[k2];
        if ((ya > y) != (yb > y)) {
          const double xi = xa + (y - ya) * (xb - xa) / (yb - ya);
          const int e = atomicAdd(&rc.n[r], 1);
          if (e < XCAP) { rc.xi[r][e] = xi; rc.body[r][e] = (unsigned char)b; } else overflow = 1;
        }
      }
    }
  }
  __syncthreads();
  if (threadIdx.x < 3) {
    const int r = threadIdx.x;
    double lo = INFINITY, hi = -INFINITY;
    const int n = min(rc.n[r], XCAP);
    for (int e = 0; e < n; e++) { lo = fmin(lo, rc.xi[r][e]); hi = fmax(hi, rc.xi[r][e]); }
    rc.lo[r] = lo; rc.hi[r] = hi;
  }
  __syncthreads();
  const bool slow = overflow != 0;
  const bool interior_row = j > 0 && j < L.ny - 1;
  int ghosts = 0;
  uint8_t* row = celltype + (size_t)jl * L.pitch;
  for (int c = threadIdx.x; c * 16 < L.pitch; c += blockDim.x) {           // 16 bytes per thread: cells 16c-15 .. 16c
    unsigned w[4] = {0x01010101u, 0x01010101u, 0x01010101u, 0x01010101u};  // padding and the ring are FLUID
    const int ibase = 16 * c - IFX_PADL;
    const bool maybe = interior_row && (slow || (rc.n[1] > 0 && xc[max(ibase, 0)] < rc.hi[1] &&
                                                 !(xc[min(ibase + 15, L.nx - 1)] < rc.lo[1])));
    // Most chunks hold no change point — no crossing of rows j-1, j, j+1 and no bounding-box edge between the cell
    // left of the chunk and the cell right of it — so all 16 cells (and their west/east neighbours) classify like
    // the first one: three look-ups instead of up to eighty.
    bool uniform = maybe && !slow && ibase >= 1 && ibase + 15 <= L.nx - 2;
    if (uniform) {
      const double xa = xc[ibase - 1], xb = xc[ibase + 16];
      for (int r = 0; r < 3; r++)
        for (int e = 0; e < rc.n[r]; e++) uniform = uniform && !(rc.xi[r][e] >= xa && rc.xi[r][e] <= xb);
      for (int b = 0; b < B.nbodies; b++)
        uniform = uniform && !(B.bbox[4 * b] >= xa && B.bbox[4 * b] <= xb) && !(B.bbox[4 * b + 1] >= xa && B.bbox[4 * b + 1] <= xb);
    }
    if (uniform) {
      uint8_t t = classify_from_row(L, xc, B, rc, 1, j, ibase);
      if (t != IFX_FLUID && (classify_from_row(L, xc, B, rc, 0, j - 1, ibase) == IFX_FLUID ||
                             classify_from_row(L, xc, B, rc, 2, j + 1, ibase) == IFX_FLUID))
        t |= IFX_GHOST;
      ghosts += ((t & 3) == IFX_GHOST) ? 16 : 0;
      w[0] = w[1] = w[2] = w[3] = 0x01010101u * (unsigned)t;
    } else if (maybe) {
#pragma unroll 1
      for (int k = 0; k < 16; k++) {
        const int i = ibase + k;
        if (i < 1 || i > L.nx - 2) continue;
        uint8_t t;
        if (slow) {
          t = celltype_of(L, xc, yc, B, i, j);                 // the cell-by-cell definition
        } else {
          t = classify_from_row(L, xc, B, rc, 1, j, i);
          if (t != IFX_FLUID &&
              (classify_from_row(L, xc, B, rc, 1, j, i - 1) == IFX_FLUID || classify_from_row(L, xc, B, rc, 1, j, i + 1) == IFX_FLUID ||
               classify_from_row(L, xc, B, rc, 0, j - 1, i) == IFX_FLUID || classify_from_row(L, xc, B, rc, 2, j + 1, i) == IFX_FLUID))
            t |= IFX_GHOST;
        }
        ghosts += ((t & 3) == IFX_GHOST);
        w[k >> 2] = (w[k >> 2] & ~(0xffu << (8 * (k & 3)))) | ((unsigned)t << (8 * (k & 3)));
      }
    }
    *reinterpret_cast<uint4*>(row + 16 * c) = make_uint4(w[0], w[1], w[2], w[3]);
  }
  // ordered compaction, pass 1: ghost cells per owned row
  for (int off = 16; off > 0; off >>= 1) ghosts += __shfl_down_sync(0xffffffffu, ghosts, off);
  if ((threadIdx.x & 31) == 0) warp_cnt[threadIdx.x >> 5] = ghosts;
  __syncthreads();
  if (threadIdx.x == 0 && jl >= 1 && jl <= L.nyl - 2) {
    int t = 0;
    for (int wv = 0; wv < 8; wv++) t += warp_cnt[wv];
    rowcount[jl - 1] = t;
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

// pass 3: write the ghost cells of each row in increasing i (=> the whole list in increasing reference id).
// 16 cells (one uint4) per thread; rows without ghost cells leave at once.
static __global__ void __launch_bounds__(256)
k_gc_fill(Layout L, const uint8_t* __restrict__ celltype, const int* __restrict__ rowstart,
          int capacity, int* __restrict__ cell, int* __restrict__ ref_id, int* __restrict__ body) {
  const int jl = 1 + blockIdx.x;
  const int j = L.j0 + jl;
  if (rowstart[blockIdx.x + 1] == rowstart[blockIdx.x]) return;
  __shared__ int warp_tot[8];
  __shared__ int base;
  if (threadIdx.x == 0) base = rowstart[blockIdx.x];
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint8_t* row = celltype + (size_t)jl * L.pitch;
  for (int c0 = 0; c0 * 16 < L.pitch; c0 += 256) {
    const int c = c0 + threadIdx.x;
    uint4 v = make_uint4(0x01010101u, 0x01010101u, 0x01010101u, 0x01010101u);
    if (c * 16 < L.pitch) v = *reinterpret_cast<const uint4*>(row + 16 * c);
    const unsigned w[4] = {v.x, v.y, v.z, v.w};
    unsigned gmask = 0;                                   // bit k: cell 16c-15+k is a ghost cell
#pragma unroll
    for (int k = 0; k < 16; k++) {
      const unsigned t = (w[k >> 2] >> (8 * (k & 3))) & 0xffu;
      const int i = 16 * c - IFX_PADL + k;
      if ((t & 3u) == IFX_GHOST && i >= 1 && i <= L.nx - 2) gmask |= 1u << k;
    }
    const int mine = __popc(gmask);
    int incl = mine;                                      // inclusive scan over the warp
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, incl, off);
      if (lane >= off) incl += t;
    }
    if (lane == 31) warp_tot[warp] = incl;
    __syncthreads();
    int before = 0, chunk = 0;
    for (int wv = 0; wv < 8; wv++) { if (wv < warp) before += warp_tot[wv]; chunk += warp_tot[wv]; }
    int k = base + before + incl - mine;
    while (gmask) {
      const int b = __ffs(gmask) - 1;
      gmask &= gmask - 1;
      const int i = 16 * c - IFX_PADL + b;
      const unsigned t = (w[b >> 2] >> (8 * (b & 3))) & 0xffu;
      if (k < capacity) { cell[k] = (int)lidx(L, i, jl); ref_id[k] = i + j * L.nx; body[k] = (int)(t >> 2); }
      k++;
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
//
// Slabs: a stencil node in a row this rank does not store is read from the owner's memory when the ghost-cell
// values are evaluated (NVLink P2P load).  Its stencil entry is then negative: -(1 + index in the owner's field),
// bit 30 of the index selecting the upper neighbour; its cell type is recomputed from the geometry.  A node more
// than IFX_GC_REACH rows beyond the slab raises *err (the sweep kernels only fence that many rows, kernels_v4.cu).
static __global__ void k_gc_geometry(Layout L, const double* __restrict__ xc, const double* __restrict__ yc, BodySet B,
                                     SlabGeom sg, const uint8_t* __restrict__ celltype, int ngc,
                                     const int* __restrict__ ref_id, const int* __restrict__ body,
                                     int* __restrict__ stencil, int* __restrict__ stencil_ref, double* __restrict__ wd,
                                     double* __restrict__ wn, double* __restrict__ bi, double* __restrict__ ip,
                                     int* __restrict__ err) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= ngc) return;
  const int* off = B.off;
  const double* xm = B.xm;
  const double* ym = B.ym;
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
  int nid[4], code[4];
  for (int m = 0; m < 4; m++) {
    nid[m] = ni[m] + nj[m] * L.nx;
    const int jl = nj[m] - L.j0;
    uint8_t c;
    if (jl >= 0 && jl < L.nyl) {
      code[m] = (int)lidx(L, ni[m], jl);
      c = celltype[code[m]];
    } else {
      c = celltype_of(L, xc, yc, B, ni[m], nj[m]);
      code[m] = 0;
      if (jl < 0) {                                   // owned by the lower neighbour: its stored rows start at jb_lo - 1
        const int jp = nj[m] - (L.jb - sg.nyl_lo + 1);
        if (!sg.has_lo || jp < 1 || nj[m] < L.jb - IFX_GC_REACH) *err = 1;
        else code[m] = -(1 + (int)((size_t)jp * L.pitch + IFX_PADL + ni[m]));
      } else {                                        // upper neighbour: its stored rows start at je - 1
        const int jp = nj[m] - (L.je - 1);
        if (!sg.has_hi || jp > sg.nyl_hi - 2 || nj[m] > L.je - 1 + IFX_GC_REACH) *err = 1;
        else code[m] = -(1 + (int)(0x40000000u | (unsigned)((size_t)jp * L.pitch + IFX_PADL + ni[m])));
      }
    }
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
    stencil[4 * g + m] = code[m];
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
__device__ __forceinline__ double gc_node(const double* __restrict__ q, const double* qlo, const double* qhi, int code) {
  if (code >= 0) return q[code];
  const unsigned c = (unsigned)(-(code + 1));
  const double* p = (c & 0x40000000u) ? qhi + (c & 0x3fffffffu) : qlo + c;
  double v;           // the owner's memory over NVLink: never through a non-coherent cache
  asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
  return v;
}

__device__ __forceinline__ double gc_dirichlet(const double* __restrict__ q, const double* qlo, const double* qhi,
                                               const int* __restrict__ st, const double* __restrict__ wd, double phi_bi) {
  double t = wd[4] * phi_bi;
  t = fma(wd[0], gc_node(q, qlo, qhi, st[0]), t);
  t = fma(wd[1], gc_node(q, qlo, qhi, st[1]), t);
  t = fma(wd[2], gc_node(q, qlo, qhi, st[2]), t);
  t = fma(wd[3], gc_node(q, qlo, qhi, st[3]), t);
  return t;
}

static __global__ void k_gc_velocity(int ngc, const int* __restrict__ cell, const int* __restrict__ stencil,
                                     const double* __restrict__ wd, const int* __restrict__ body,
                                     const double* __restrict__ ub, const double* __restrict__ vb,
                                     const double* __restrict__ usrc, const double* __restrict__ vsrc, GcPeers pr,
                                     double* __restrict__ udst, double* __restrict__ vdst, int gather,
                                     const LoopCtl* ctl, int iter, GcPush ps) {
  // in-loop use: iteration `iter` ran iff the loop was not already finished by an earlier iteration
  if (ctl && ctl->done && ctl->iter < iter) return;
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g < ngc) {
    const double pu = gc_dirichlet(usrc, pr.lo[0], pr.hi[0], stencil + 4 * g, wd + 5 * g, ub[body[g]]);
    const double pv = gc_dirichlet(vsrc, pr.lo[1], pr.hi[1], stencil + 4 * g, wd + 5 * g, vb[body[g]]);
    const int o = gather ? g : cell[g];
    udst[o] = pu; vdst[o] = pv;
    if (ps.active) {      // a ghost cell of my first / last owned row also lives in the neighbour's halo row
      if (ps.has_lo && o >= ps.row_lo && o < ps.row_lo + ps.pitch) { ps.dst_lo[0][o - ps.row_lo] = pu; ps.dst_lo[1][o - ps.row_lo] = pv; }
      if (ps.has_hi && o >= ps.row_hi && o < ps.row_hi + ps.pitch) { ps.dst_hi[0][o - ps.row_hi] = pu; ps.dst_hi[1][o - ps.row_hi] = pv; }
    }
  }
  if (!ps.active) return;
  // the whole grid is through (ticket): my remote reads of the neighbours' previous iterate are over and the halo rows
  // I feed are complete — tell both neighbours
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned t = atomicAdd(ps.ticket, 1u);
    if (t == gridDim.x - 1) {
      *ps.ticket = 0u;
      __threadfence_system();
      if (ps.has_lo) st_release_sys(ps.signal_lo, ps.seq);
      if (ps.has_hi) st_release_sys(ps.signal_hi, ps.seq);
    }
  }
}

static __global__ void k_gc_pressure(int ngc, const int* __restrict__ cell, const int* __restrict__ stencil,
                                     const double* __restrict__ wn, const double* __restrict__ psrc, GcPeers pr,
                                     double* __restrict__ pdst, int gather) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= ngc) return;
  const int* st = stencil + 4 * g;
  const double* w = wn + 4 * g;
  double t = w[0] * gc_node(psrc, pr.lo[0], pr.hi[0], st[0]);
  t = fma(w[1], gc_node(psrc, pr.lo[0], pr.hi[0], st[1]), t);
  t = fma(w[2], gc_node(psrc, pr.lo[0], pr.hi[0], st[2]), t);
  t = fma(w[3], gc_node(psrc, pr.lo[0], pr.hi[0], st[3]), t);
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
cudaError_t launch_classify(const Layout& L, const double* xc, const double* yc, const BodySet& B, uint8_t* celltype,
                            int* rowcount, cudaStream_t st) {
  k_classify_rows<<<L.nyl, 256, 0, st>>>(L, xc, yc, B, celltype, rowcount);
  return cudaGetLastError();
}

cudaError_t launch_gc_count(const Layout& L, const int* rowcount, int* rowstart, int* total, cudaStream_t st) {
  k_gc_scan_rows<<<1, 1024, 0, st>>>(L.nyl - 2, rowcount, rowstart, total);
  return cudaGetLastError();
}

cudaError_t launch_gc_build(const Layout& L, const double* xc, const double* yc, const BodySet& B, const SlabGeom& sg,
                            const uint8_t* celltype, const int* rowstart, int ngc, int* cell, int* ref_id,
                            int* body, int* stencil, int* stencil_ref, double* wd, double* wn, double* bi, double* ip,
                            int* err, cudaStream_t st) {
  if (ngc <= 0) return cudaSuccess;
  k_gc_fill<<<L.nyl - 2, 256, 0, st>>>(L, celltype, rowstart, ngc, cell, ref_id, body);
  k_gc_geometry<<<(ngc + 127) / 128, 128, 0, st>>>(L, xc, yc, B, sg, celltype, ngc, ref_id, body, stencil, stencil_ref,
                                                   wd, wn, bi, ip, err);
  return cudaGetLastError();
}

cudaError_t launch_gc_velocity(int ngc, const int* cell, const int* stencil, const double* wd, const int* body,
                               const double* ub, const double* vb, const double* usrc, const double* vsrc,
                               const GcPeers& pr, double* udst, double* vdst, int gather, const LoopCtl* ctl, int iter,
                               cudaStream_t st, const GcPush* push) {
  GcPush ps{};
  if (push) ps = *push;
  if (ngc <= 0 && !ps.active) return cudaSuccess;        // (a slab without ghost cells still has its flag to publish)
  const int nblk = ngc > 0 ? (ngc + 127) / 128 : 1;
  k_gc_velocity<<<nblk, 128, 0, st>>>(ngc, cell, stencil, wd, body, ub, vb, usrc, vsrc, pr, udst, vdst, gather, ctl, iter, ps);
  return cudaGetLastError();
}

cudaError_t launch_gc_pressure(int ngc, const int* cell, const int* stencil, const double* wn, const double* psrc,
                               const GcPeers& pr, double* pdst, int gather, cudaStream_t st) {
  if (ngc <= 0) return cudaSuccess;
  k_gc_pressure<<<(ngc + 127) / 128, 128, 0, st>>>(ngc, cell, stencil, wn, psrc, pr, pdst, gather);
  return cudaGetLastError();
}

cudaError_t launch_gc_scatter(int ngc, const int* cell, const double* a, double* qa, const double* b, double* qb,
                              cudaStream_t st) {
  if (ngc <= 0) return cudaSuccess;
  k_gc_scatter<<<(ngc + 127) / 128, 128, 0, st>>>(ngc, cell, a, qa, b, qb);
  return cudaGetLastError();
}

}  // namespace ifx
