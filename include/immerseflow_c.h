/*
 * immerseflow_c.h — C-ABI of the B200-native ImmerseFlow++ fractional-step path.
 *
 * The reference has no library boundary: its seam is (i) the files in ../inputs and ../results
 * and (ii) the `ImmerseFlow` member functions that `main` calls (reference src/main.cu:75-101,
 * src/header/globalVariables.cuh:70-88).  Each entry point below replaces one of those member
 * functions / free functions; the reference interface it replaces is cited as file:line
 * (paths relative to the reference repo root).  Plain pointers and sizes only — no C++ or torch
 * types cross this boundary.  All functions return IFX_OK (0) or a negative ifx_status; the
 * message is available from ifx_last_error().  Nothing here ever calls exit() (the reference's
 * CHECK_CUDA_ERROR does, src/header/globalVariables.cuh:91-107; the CLI driver maps a non-zero
 * status to that behaviour).
 *
 * Host arrays passed in or out use the REFERENCE layout: row-major, id = i + j*nx, i fastest,
 * ghost-inclusive nx = nx_cells+2, ny = ny_cells+2 (src/main.cu:55-58), fp64.
 * Device memory is owned by the handle for its whole life (no per-step cudaMalloc).
 */
#ifndef IMMERSEFLOW_C_H
#define IMMERSEFLOW_C_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define IFX_ABI_VERSION 2

typedef enum {
  IFX_OK = 0,
  IFX_ERR_INVALID = -1,   /* bad argument / unsupported combination            */
  IFX_ERR_IO = -2,        /* file could not be opened / parsed                 */
  IFX_ERR_CUDA = -3,      /* a CUDA runtime call or kernel failed              */
  IFX_ERR_NOMEM = -4,
  IFX_ERR_STATE = -5      /* call sequence error (e.g. step before initialize) */
} ifx_status;

/* Mirror of `struct CFDInput` (src/header/globalVariables.cuh:15-33) with the derived sizes
 * exactly as readInputFile leaves them (src/main.cu:55-58): nx, ny are GHOST-INCLUSIVE. */
typedef struct {
  int Restart, Restart_Time;
  int nx, ny;           /* nx_cells + 2, ny_cells + 2 */
  int nxf, nyf;         /* nx_cells + 1, ny_cells + 1 */
  double Lx, Ly;
  int w_AD, w_PPE, AD_itermax, PPE_itermax, AD_solver, PPE_solver;
  double ErrorMax, tmax, dt, Re, mu;
  int Write_Interval;
} ifx_input;

/* Mirror of `struct BC` (src/header/globalVariables.cuh:35-39).  The reference never fills it and
 * hard-codes u = 1, v = 0 on all four sides (src/include/ADSolver.cu:200-216); those are the
 * defaults from ifx_default_options(). */
typedef struct {
  double u_bc_w, u_bc_e, u_bc_n, u_bc_s;
  double v_bc_w, v_bc_e, v_bc_n, v_bc_s;
  double p_bc_w, p_bc_e, p_bc_n, p_bc_s;
} ifx_bc;

typedef enum {
  /* Reproduce the reference AS WRITTEN: predictor only in ifx_step, vf never
   * computed when nx <= ny, velf taken before the BC refresh, ghost values lagging two iterates,
   * Laplace PPE with p = 100 on W/S in ifx_ppe_solve, iBlank == 1. */
  IFX_COMPAT_REFERENCE = 0,
  /* The complete fractional step north_star describes: predictor -> PPE (source term, Neumann
   * BC, ghost-cell IB) -> projection.  Semantics defined by oracle/ifx_oracle_full.c (the
   * reference has no code for these stages). */
  IFX_COMPAT_FULL = 1
} ifx_compat;

typedef enum {
  IFX_REDUCE_FUSED = 0,      /* residual sums fused into the sweep kernels; the stop decision is
                                certified against the reference summation order and re-evaluated
                                in that order only when within the rounding band of the tolerance */
  IFX_REDUCE_REFERENCE = 1   /* every residual summed in the reference's reduce6<256> order
                                (src/include/preSim.cu:12-50,376-441): bit-identical residuals */
} ifx_reduce_mode;

typedef struct {
  int abi_version;          /* IFX_ABI_VERSION */
  int device;               /* CUDA device ordinal (reference: always 0, preSim.cu:189) */
  ifx_compat compat;
  ifx_reduce_mode reduce_mode;
  ifx_bc bc;
  double ad_tol;            /* reference hard-codes 1e-6 = pow(10,-6) (ADSolver.cu:315) */
  double ppe_tol;           /* reference hard-codes 1e-6 (PPESolver.cu:172)             */
  int ppe_abs_residual;     /* 0: signed sum as the reference (PPESolver.cu:42-46,172); 1: sum |r| */
  /* slab decomposition along j (rows): this handle owns interior rows [j_begin, j_end) of the
   * global grid, 1 <= j_begin < j_end <= ny-1.  Single GPU: j_begin = 1, j_end = ny-1. */
  int rank, nranks;
  int j_begin, j_end;
  int sweeps_per_batch;     /* Poisson sweeps enqueued between host looks at the stop flag */
  int use_graphs;           /* 1 (default): PPE_Solver 4 replays the coarse part of its V-cycle (~130 small launches) from a CUDA
                               graph captured once per handle; results identical (measured on B200: cavity 1024^2 6.9 -> 3.8 ms
                               per solve, 4096^2 19.0 -> 16.0 ms; the capture costs one slower first solve).  0: plain launches.
                               (The sweep loops themselves are plain launches in batches; a graph `while` node for
                               them is DESIGN.md's next step for launch-bound small grids.) */
  /* Poisson iteration (IFX_COMPAT_FULL).  The reference documents PPE_Solver "1. Point GS, 2. Line
   * SOR" and w-PPE in inputs.txt, parses them (main.cu:42) and always runs point Jacobi.  0 = take
   * ifx_input.PPE_solver (full mode; the reference-compatible mode always runs the reference's Jacobi);
   *   1 = point Jacobi (the reference's sweep);
   *   2 = zebra line SOR with factor ppe_omega: x-lines even / odd, y-lines even / odd, each line solved exactly;
   *   3 = red-black SOR with factor ppe_omega;
   *   4 = geometric multigrid, V(2,2) cycles smoothed by red-black SOR — for near-isotropic cells (uniform grids);
   *   5 = geometric multigrid smoothed by the line relaxation of 2 — for the stretched grids the reference ships.
   * For 4 and 5 an "iteration" of PPE_itermax / ifx_step_stats.ppe_sweeps is then one V-cycle.
   * 2, 4 and 5 are single-GPU for now. */
  int ppe_solver;
  double ppe_omega;         /* 0 = take ifx_input.w_PPE (an int in the reference's struct, globalVariables.cuh:26) */
  /* 1: the few-byte control traffic of a step (stop flags, ghost-cell counts, marker uploads) moves through kernels
   * that read / write page-locked host memory directly instead of through cudaMemcpyAsync.  Results are identical.
   * For processes that run several handles at once with bulk ifx_set/get_field_async transfers in flight: a copy
   * engine serves whole-field transfers ahead of a step's control copies and stalls it (measured on B200: a
   * concurrent download delays a 31 ms step by 16 ms; bench.py's end-to-end pipeline sets this). */
  int zero_copy_control;
  /* 1 (ppe_solver 1, IFX_COMPAT_FULL, single GPU): point-Jacobi sweeps run two per pass over memory (kernels_pair.cu:
   * the intermediate iterate lives in shared memory only, both residuals are evaluated, results and iteration counts
   * are bit-identical).  Halves the DRAM traffic of the Poisson solve; measured on B200 it is NOT faster than single
   * sweeps (issue-bound: profiles/r2_pair_kernel.md), hence off by default. */
  int ppe_pairs;
  int reserved[3];
} ifx_options;

/* Per-call statistics (replaces the reference's printf of "iter = %d %f %f", ADSolver.cu:369). */
typedef struct {
  int ad_iters;
  double ad_ures, ad_vres;  /* residual sums at exit */
  int ppe_sweeps;
  double ppe_residual;
  int exact_fallbacks;      /* stop decisions that had to be re-evaluated in reference order */
  float ms_ad, ms_ppe, ms_correct, ms_ib, ms_total;   /* CUDA-event stage timings */
  float ms_ad_sweeps;       /* of ms_ad, the Jacobi sweeps alone (ad_iters launches + their ghost-cell kernels): what a
                               per-launch bandwidth of the predictor sweep is computed from */
} ifx_step_stats;

typedef enum {
  IFX_FIELD_U = 0,      /* Data.u.velc */
  IFX_FIELD_V = 1,      /* Data.v.velc */
  IFX_FIELD_P = 2,      /* Data.p      */
  IFX_FIELD_IBLANK = 3, /* ibm.iBlank as doubles, 1.0 fluid / 0.0 solid (globalVariables.cuh:50-52) */
  IFX_FIELD_UF = 4,     /* Data.u.velf, (nx-1)*(ny-2) values, id = i + (j-1)*(nx-1)            */
  IFX_FIELD_VF = 5,     /* Data.v.velf, logical extent (nx-2)*(ny-1), id = (i-1) + j*(nx-2)    */
  IFX_FIELD_SX = 6,     /* predictor right-hand sides (ADSolver.cu:55-74) */
  IFX_FIELD_SY = 7,
  IFX_FIELD_PPE_RHS = 8,
  IFX_FIELD_XC = 9,     /* nx values */
  IFX_FIELD_YC = 10,    /* ny values */
  IFX_FIELD_CELLTYPE = 11 /* as doubles: 1 fluid, 0 solid, 2 ghost cell */
} ifx_field;

typedef struct ifx_solver ifx_solver; /* opaque; replaces `struct ImmerseFlow` (globalVariables.cuh:70-88) */

/* ---- library / diagnostics -------------------------------------------------------------- */
int ifx_abi_version(void);
const char* ifx_last_error(const ifx_solver* s);      /* s may be NULL: last error of failed create */
int ifx_device_count(void);
void ifx_default_options(ifx_options* opt);

/* ---- file contract (host only, no CUDA) -------------------------------------------------- */
/* replaces readInputFile(), src/main.cu:10-59: keyword line then value line, separators = / _ */
int ifx_read_input_file(const char* path, ifx_input* in);
/* replaces the grid loops of readGridData(), src/include/preSim.cu:268-291: n "index value" pairs */
int ifx_read_grid_file(const char* path, int n, double* faces);
/* replaces write_results_to_file(), src/include/postSim.cu:41-66 (Tecplot ASCII POINT, "%f,%f,%f") */
int ifx_write_results_to_file(const double* x, const double* y, const double* data,
                              int ni, int nj, const char* filename);

/* ---- lifetime ---------------------------------------------------------------------------- */
/* replaces CUDAQuery() + allocation() + readGridData() metrics (preSim.cu:147-162,187-199,294-363).
 * xf: in->nxf face coordinates, yf: in->nyf. */
int ifx_create(const ifx_input* in, const double* xf, const double* yf,
               const ifx_options* opt, ifx_solver** out);
/* replaces freeAllocation() (preSim.cu:164-179), without its hidden rewrite of uc.dat */
int ifx_destroy(ifx_solver* s);

/* replaces initializeData(), preSim.cu:201-217: vortex IC (initializeKernel) + iBlank */
int ifx_initialize(ifx_solver* s);

/* ---- state access (host buffers, reference layout; slab handles move only their rows) ------ */
size_t ifx_field_size(const ifx_solver* s, ifx_field f);      /* number of doubles */
int ifx_set_field(ifx_solver* s, ifx_field f, const double* host, size_t n);
int ifx_get_field(ifx_solver* s, ifx_field f, double* host, size_t n);
/* The same, enqueued on the handle's stream without waiting: the transfer overlaps whatever OTHER handles are running
 * (several handles of one process = several simulations in flight: one uploading its next state, one stepping, one
 * downloading its result — bench.py's end-to-end pipeline).  `host` must be page-locked and stay untouched until
 * ifx_synchronize() or the next blocking call on this handle returns. */
int ifx_set_field_async(ifx_solver* s, ifx_field f, const double* host, size_t n);
int ifx_get_field_async(ifx_solver* s, ifx_field f, double* host, size_t n);
/* replaces saveDataToFile(), postSim.cu:10-39: D2H + Tecplot ASCII */
int ifx_save_field(ifx_solver* s, ifx_field f, const char* filename);
/* Restart files: what `Restart` / `Restart_Time` / `Write_Interval` of inputs.txt ask for (parsed at
 * main.cu:27-30,47-50, never used by the reference; predecessor: test/UTIL_PRE_SIM.f90:120-150).  Raw fp64 of
 * u, v, p (+ uf, vf in IFX_COMPAT_FULL); a run continued from a file is bit-identical to an uninterrupted one.
 * Slab handles write / read their own rows (one file per rank). */
int ifx_checkpoint_write(ifx_solver* s, const char* path, long long step, double time);
int ifx_checkpoint_read(ifx_solver* s, const char* path, long long* step, double* time);

/* ---- the hot path -------------------------------------------------------------------------
 * Every solve returns IFX_ERR_STATE when the residual it ends on is not finite: NaN > tol is false, so the reference's
 * stop rules (ADSolver.cu:315, PPESolver.cu:172) would end "converged" on a diverged state. */
/* replaces ImmerseFlow::ADsolver(), src/include/ADSolver.cu:268-395 (one predictor step; no file I/O) */
int ifx_ad_solve(ifx_solver* s, ifx_step_stats* stats);
/* replaces ImmerseFlow::PPESolver(), src/include/PPESolver.cu:137-205 */
int ifx_ppe_solve(ifx_solver* s, ifx_step_stats* stats);
/* the projection the reference leaves empty, src/include/AD_PPE_Correction.cu:1-12 */
int ifx_correct(ifx_solver* s, ifx_step_stats* stats);
/* one time step of main()'s loop, src/main.cu:93-96.  IFX_COMPAT_REFERENCE: == ifx_ad_solve.
 * IFX_COMPAT_FULL: IB update (if bodies moved) -> predictor -> PPE -> correction. */
int ifx_step(ifx_solver* s, ifx_step_stats* stats);
/* per-iteration (uRes, vRes) of the last predictor solve, what the reference prints as "iter = %d %f %f"
 * (src/include/ADSolver.cu:369); at most 64 iterations are kept.  Returns the number of pairs written. */
int ifx_get_residual_history(ifx_solver* s, double* pairs, int capacity);
/* replaces ImmerseFlow::Reduction(), src/include/preSim.cu:376-445: sum of n doubles (host input),
 * bit-identical summation order */
int ifx_reduce_sum(ifx_solver* s, const double* host, size_t n, double* out);

/* ---- immersed boundary --------------------------------------------------------------------- */
/* Bodies are closed polygons of surface markers (counter-clockwise), the 2-D analogue of the
 * predecessor's marker meshes; body b uses markers [offsets[b], offsets[b+1]).  ubody/vbody: rigid
 * velocity of each body (Dirichlet value at the body intercept), may be NULL (= 0).  Bodies must lie strictly inside
 * the first / last interior cell centres (a body on the grid boundary has ghost cells whose image points leave the
 * grid); otherwise IFX_ERR_INVALID, and the solver is left without bodies. */
int ifx_set_bodies(ifx_solver* s, int nbodies, const int* offsets, const double* xm, const double* ym,
                   const double* ubody, const double* vbody);
/* replaces iBlankComputeKernel (preSim.cu:110-136) + everything the reference lacks: cell
 * classification, ghost-cell list, body intercepts, image points, interpolation stencils */
int ifx_iblank_update(ifx_solver* s, ifx_step_stats* stats);
int ifx_ghost_cell_count(const ifx_solver* s);
/* ghost-cell maps for parity tests: cell ids (reference id = i + j*nx), 4 stencil cell ids and 4
 * weights per ghost cell, body-intercept and image-point coordinates.  Any pointer may be NULL. */
int ifx_get_ghost_cells(ifx_solver* s, int* cell_id, int* stencil_id, double* weights,
                        double* bi_xy, double* ip_xy, int capacity);

/* ---- diagnostics (IFX_COMPAT_FULL, single GPU; SURVEY 8(f)-4 — the reference has none, its predecessor did:
 *      test/UTIL_PRE_SIM.f90:172-201) ------------------------------------------------------------------------- */
/* u, v, p interpolated bilinearly at n points (x[k], y[k]); ghost cells take part with their boundary values,
 * cells inside a body are dropped from the interpolation */
int ifx_probe(ifx_solver* s, int n, const double* x, const double* y, double* u, double* v, double* p);
/* surface force on every body, 4 doubles per body: pressure force x, y, viscous force x, y (per unit span; the drag
 * coefficient is 2 (Fpx + Fvx) / (U^2 D)).  Two probes per marker segment, 1.5 and 3 cell diagonals off the surface:
 * wall pressure by linear extrapolation, wall shear from a second-order one-sided difference. */
int ifx_body_forces(ifx_solver* s, double* forces, int capacity_bodies);

/* ---- multi-GPU slabs (one process per GPU; halos move over NVLink peer mappings) -------------- */
#define IFX_IPC_HANDLE_BYTES 128
/* export this rank's exchange segment; the launcher all-gathers the handles (torch.distributed) */
int ifx_ipc_export(ifx_solver* s, unsigned char handle[IFX_IPC_HANDLE_BYTES]);
/* map the neighbours' segments (NULL = domain boundary on that side) and all ranks' mailboxes */
int ifx_ipc_connect(ifx_solver* s, const unsigned char* all_handles, int nranks);

/* ---- plumbing -------------------------------------------------------------------------------- */
int ifx_set_stream(ifx_solver* s, void* cuda_stream);   /* run on the caller's stream (e.g. torch's) */
int ifx_synchronize(ifx_solver* s);
/* number of kernel launches issued by this handle since creation (bench.py's gpu_launches) */
long long ifx_launch_count(const ifx_solver* s);

#ifdef __cplusplus
}
#endif
#endif /* IMMERSEFLOW_C_H */
