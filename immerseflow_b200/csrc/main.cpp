// main.cpp — drop-in driver for the reference's `main` (src/main.cu:75-101), host C++ over the C-ABI.
//
// With no arguments it does what the reference binary does when run from its src/ directory: read
// ../inputs/inputs.txt and ../inputs/{x,y}grid.dat2 (the files the reference's code actually opens,
// preSim.cu:268,281), build the vortex initial condition, write ../results/final_results.dat (iBlank,
// preSim.cu:217), run `tmax` predictor steps and leave ../results/uc.dat and ../results/vc.dat
// (ADSolver.cu:378-379) in the reference's Tecplot ASCII format.  Options select the stretched grids,
// the complete fractional step (--mode full: predictor, Poisson with source term, projection, optional
// immersed bodies; also writes ../results/p.dat, PPESolver.cu:197) and where files live.
//
// Error behaviour follows the reference: a message on stderr and exit(1) (globalVariables.cuh:91-107,
// main.cu:12-15), here driven by the library's status codes instead of exit() calls inside the library.
#include "../../include/immerseflow_c.h"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <sstream>
#include <string>
#include <vector>

static void die(const std::string& msg) {
  std::cerr << msg << std::endl;
  std::exit(1);
}

struct Args {
  std::string input = "../inputs/inputs.txt";
  std::string xgrid = "../inputs/xgrid.dat2";
  std::string ygrid = "../inputs/ygrid.dat2";
  std::string results = "../results";
  std::string bodies;
  std::string mode = "reference";
  long steps = -1;
  bool write_every_step = false;
  bool reference_log = false;
  bool exact = false;
  bool checkpoints = false;      // write restart files every Write_Interval steps and after the last one
  int ppe_solver = 0;            // 0: PPE_Solver of inputs.txt; 1 point Jacobi; 2 line SOR; 3 red-black SOR; 4, 5 multigrid (full mode)
  double ppe_omega = 0.0;        // 0: w-PPE of inputs.txt (an integer there)
  std::string restart;           // explicit restart file (default when Restart != 0: <results>/restart.<Restart_Time>.ifx)
  int device = 0;
  // what the reference hard-codes (BC values ADSolver.cu:200-216, vortex IC preSim.cu:63-73) as options
  bool have_bc_u = false, have_bc_v = false;
  double bc_u[4] = {1, 1, 1, 1}, bc_v[4] = {0, 0, 0, 0};      // W, E, S, N
  std::string ic = "vortex";     // vortex | zero | uniform:U,V
  double ad_tol = 0.0, ppe_tol = 0.0;   // 0: the reference's hard-coded 1e-6 (ADSolver.cu:315, PPESolver.cu:172)
  std::string forces;            // per step and body: pressure and viscous force (ifx_body_forces)
  std::string probes, probe_out; // probe points in, per-step u v p out (ifx_probe)
};

// one immersed body of the bodies file: marker polygon at t = 0 + rigid translation
//   x(t) = x0 + ub t + ax sin(2 pi f t),  y(t) = y0 + vb t + ay sin(2 pi f t)
struct Body {
  std::vector<double> x0, y0;
  double ub = 0, vb = 0, ax = 0, ay = 0, freq = 0;
  bool moves() const { return ub != 0 || vb != 0 || ((ax != 0 || ay != 0) && freq != 0); }
};

static bool four(const std::string& v, double* out) {
  return std::sscanf(v.c_str(), "%lf,%lf,%lf,%lf", out, out + 1, out + 2, out + 3) == 4;
}

static void usage() {
  std::cout <<
      "immerseflow [--input FILE] [--xgrid FILE] [--ygrid FILE] [--stretched] [--results DIR]\n"
      "            [--mode reference|full] [--bodies FILE] [--steps N] [--write-every-step]\n"
      "            [--reference-log] [--exact-reduction] [--checkpoints] [--restart FILE] [--device K]\n"
      "            [--ppe-solver 1..5] [--ppe-omega W] [--bc-u W,E,S,N] [--bc-v W,E,S,N] [--ic vortex|zero|uniform:U,V]\n"
      "            [--forces FILE] [--probes FILE --probe-out FILE] [--ad-tol T] [--ppe-tol T]\n"
      "  defaults reproduce the reference binary run from src/: ../inputs/inputs.txt, ../inputs/{x,y}grid.dat2,\n"
      "  tmax predictor steps, ../results/{final_results,uc,vc}.dat.  --stretched picks ../inputs/{x,y}grid.dat.\n"
      "  inputs.txt `Write Interval` N: results are (re)written every N steps as well as after the last one; with\n"
      "  --checkpoints a restart file <results>/restart.<step, 7 digits>.ifx goes with them.  `Restart 1 T` in inputs.txt\n"
      "  (or --restart FILE) continues from <results>/restart.<T>.ifx: steps T+1 .. tmax, bit-identical to an unbroken run.\n"
      "  --ppe-solver 3 (or PPE_Solver 3 in inputs.txt; full mode): red-black SOR with factor --ppe-omega / w-PPE.\n"
      "  --ppe-solver 2 (PPE_Solver 2, the input file's \"Line SOR\"): zebra line relaxation, factor --ppe-omega / w-PPE.\n"
      "  --ppe-solver 4 / 5: geometric multigrid V(2,2) cycles smoothed by red-black SOR (uniform grids) / by line\n"
      "  relaxation (stretched grids); PPE_itermax then counts cycles.\n"
      "  --bodies FILE (full mode): `nbodies`, then per body a line `nmarkers ub vb [ax ay f]` and nmarkers lines `x y`\n"
      "  (counter-clockwise).  A body translates as x0 + ub t + ax sin(2 pi f t) (same in y); moving bodies are\n"
      "  re-classified every step.\n"
      "  --bc-u / --bc-v: wall values of u and v on the west, east, south, north side (reference: 1,1,1,1 / 0,0,0,0;\n"
      "  lid-driven cavity: --bc-u 0,0,0,1).  --ic: initial condition (reference: its Gaussian vortex).\n"
      "  --ad-tol / --ppe-tol: stop tolerances of the predictor and Poisson iterations.  The reference hard-codes 1e-6 for\n"
      "  both and ignores ErrorMax; its residuals are un-normalised sums over all cells, so on large grids 1e-6 is below\n"
      "  the rounding floor and the loops run to their iteration caps — give a tolerance that scales with the grid.\n"
      "  --forces FILE (full mode): one line per step and body `step time body Fpx Fpy Fvx Fvy`.\n"
      "  --probes FILE: lines `x y`; --probe-out FILE gets one line per step and point `step time k u v p`.\n";
}

static Args parse(int argc, char** argv) {
  Args a;
  for (int k = 1; k < argc; k++) {
    std::string o = argv[k];
    auto val = [&]() -> std::string { if (k + 1 >= argc) die("missing value for " + o); return argv[++k]; };
    if (o == "--input") a.input = val();
    else if (o == "--xgrid") a.xgrid = val();
    else if (o == "--ygrid") a.ygrid = val();
    else if (o == "--stretched") { a.xgrid = "../inputs/xgrid.dat"; a.ygrid = "../inputs/ygrid.dat"; }
    else if (o == "--results") a.results = val();
    else if (o == "--bodies") a.bodies = val();
    else if (o == "--mode") a.mode = val();
    else if (o == "--steps") a.steps = std::atol(val().c_str());
    else if (o == "--write-every-step") a.write_every_step = true;
    else if (o == "--reference-log") a.reference_log = true;
    else if (o == "--exact-reduction") a.exact = true;
    else if (o == "--checkpoints") a.checkpoints = true;
    else if (o == "--ppe-solver") a.ppe_solver = std::atoi(val().c_str());
    else if (o == "--ppe-omega") a.ppe_omega = std::atof(val().c_str());
    else if (o == "--restart") a.restart = val();
    else if (o == "--device") a.device = std::atoi(val().c_str());
    else if (o == "--bc-u") { if (!four(val(), a.bc_u)) die("--bc-u needs W,E,S,N"); a.have_bc_u = true; }
    else if (o == "--bc-v") { if (!four(val(), a.bc_v)) die("--bc-v needs W,E,S,N"); a.have_bc_v = true; }
    else if (o == "--ic") a.ic = val();
    else if (o == "--ad-tol") a.ad_tol = std::atof(val().c_str());
    else if (o == "--ppe-tol") a.ppe_tol = std::atof(val().c_str());
    else if (o == "--forces") a.forces = val();
    else if (o == "--probes") a.probes = val();
    else if (o == "--probe-out") a.probe_out = val();
    else if (o == "-h" || o == "--help") { usage(); std::exit(0); }
    else die("unknown option " + o);
  }
  if (a.mode != "reference" && a.mode != "full") die("--mode must be reference or full");
  if (a.ic != "vortex" && a.ic != "zero" && a.ic.rfind("uniform:", 0) != 0) die("--ic must be vortex, zero or uniform:U,V");
  if ((!a.forces.empty() || !a.probes.empty()) && a.mode != "full") die("--forces / --probes need --mode full");
  if (a.probes.empty() != a.probe_out.empty()) die("--probes and --probe-out go together");
  return a;
}

static void check(ifx_solver* s, int rc, const char* what) {
  if (rc != IFX_OK) die(std::string("ImmerseFlow error in ") + what + ": " + ifx_last_error(s));
}

int main(int argc, char** argv) {
  const Args a = parse(argc, argv);
  std::cout << "ImmerseFlow++ fractional-step path, B200-native build (mode: " << a.mode << ")" << std::endl;

  ifx_input in;
  if (ifx_read_input_file(a.input.c_str(), &in) != IFX_OK) die("Unable to open file: " + a.input);       // main.cu:12-15
  std::vector<double> xf(in.nxf), yf(in.nyf);
  if (ifx_read_grid_file(a.xgrid.c_str(), in.nxf, xf.data()) != IFX_OK) die("Error opening xgrid.dat");    // preSim.cu:270
  if (ifx_read_grid_file(a.ygrid.c_str(), in.nyf, yf.data()) != IFX_OK) die("Error opening ygrid.dat");    // preSim.cu:283

  ifx_options opt;
  ifx_default_options(&opt);
  opt.device = a.device;
  opt.compat = (a.mode == "full") ? IFX_COMPAT_FULL : IFX_COMPAT_REFERENCE;
  opt.reduce_mode = a.exact ? IFX_REDUCE_REFERENCE : IFX_REDUCE_FUSED;
  opt.ppe_abs_residual = (a.mode == "full") ? 1 : 0;
  opt.ppe_solver = (a.mode == "full") ? a.ppe_solver : 1;       // PPE_Solver / w-PPE (main.cu:42): full mode only
  opt.ppe_omega = a.ppe_omega;
  if (a.ad_tol > 0.0) opt.ad_tol = a.ad_tol;
  if (a.ppe_tol > 0.0) opt.ppe_tol = a.ppe_tol;
  opt.bc.u_bc_w = a.bc_u[0]; opt.bc.u_bc_e = a.bc_u[1]; opt.bc.u_bc_s = a.bc_u[2]; opt.bc.u_bc_n = a.bc_u[3];
  opt.bc.v_bc_w = a.bc_v[0]; opt.bc.v_bc_e = a.bc_v[1]; opt.bc.v_bc_s = a.bc_v[2]; opt.bc.v_bc_n = a.bc_v[3];
  ifx_solver* s = nullptr;
  if (ifx_create(&in, xf.data(), yf.data(), &opt, &s) != IFX_OK) die(std::string("ImmerseFlow error in ifx_create: ") + ifx_last_error(nullptr));
  std::printf("grid %d x %d cells (%d x %d with ghost cells), dt = %g, Re = %g, AD_itermax = %d, PPE_itermax = %d\n",
              in.nx - 2, in.ny - 2, in.nx, in.ny, in.dt, in.Re, in.AD_itermax, in.PPE_itermax);

  std::vector<Body> bodies;
  bool any_moves = false;
  if (!a.bodies.empty()) {
    if (a.mode != "full") die("--bodies needs --mode full (the reference has no immersed-boundary code to be compatible with)");
    std::ifstream f(a.bodies);
    if (!f) die("Unable to open file: " + a.bodies);
    int nb = 0;
    std::string line;
    while (std::getline(f, line)) { std::istringstream is(line); if (is >> nb) break; }
    if (nb < 0 || nb > 63) die("malformed bodies file: body count");
    for (int b = 0; b < nb; b++) {
      Body body;
      int n = 0;
      bool got = false;
      while (std::getline(f, line)) {        // header line: nmarkers ub vb [ax ay f]
        std::istringstream is(line);
        if (is >> n >> body.ub >> body.vb) { is >> body.ax >> body.ay >> body.freq; got = true; break; }
      }
      if (!got || n < 3) die("malformed bodies file: body header");
      for (int k = 0; k < n; k++) { double x, y; if (!(f >> x >> y)) die("malformed bodies file: markers"); body.x0.push_back(x); body.y0.push_back(y); }
      std::getline(f, line);                 // rest of the last marker line
      any_moves = any_moves || body.moves();
      bodies.push_back(body);
    }
  }
  // marker positions and rigid velocities at time t, handed to the library (re-classified lazily by the next step)
  auto place_bodies = [&](double t) {
    if (bodies.empty()) return;
    const double two_pi = 6.283185307179586476925286766559;
    std::vector<int> off(1, 0);
    std::vector<double> xm, ym, ub, vb;
    for (const Body& b : bodies) {
      const double sn = std::sin(two_pi * b.freq * t), cs = std::cos(two_pi * b.freq * t);
      const double dxb = b.ub * t + b.ax * sn, dyb = b.vb * t + b.ay * sn;
      ub.push_back(b.ub + b.ax * two_pi * b.freq * cs);
      vb.push_back(b.vb + b.ay * two_pi * b.freq * cs);
      for (size_t k = 0; k < b.x0.size(); k++) { xm.push_back(b.x0[k] + dxb); ym.push_back(b.y0[k] + dyb); }
      off.push_back((int)xm.size());
    }
    check(s, ifx_set_bodies(s, (int)bodies.size(), off.data(), xm.data(), ym.data(), ub.data(), vb.data()), "ifx_set_bodies");
  };
  place_bodies(0.0);

  check(s, ifx_initialize(s), "ifx_initialize");
  if (a.ic != "vortex") {                      // the reference only has its vortex (initializeKernel, preSim.cu:63-73)
    double u0 = 0.0, v0 = 0.0;
    if (a.ic != "zero" && std::sscanf(a.ic.c_str(), "uniform:%lf,%lf", &u0, &v0) != 2) die("--ic uniform:U,V");
    const size_t n = ifx_field_size(s, IFX_FIELD_U);
    std::vector<double> q(n, u0);
    check(s, ifx_set_field(s, IFX_FIELD_U, q.data(), n), "ifx_set_field(u)");
    q.assign(n, v0);
    check(s, ifx_set_field(s, IFX_FIELD_V, q.data(), n), "ifx_set_field(v)");
    q.assign(n, 0.0);
    check(s, ifx_set_field(s, IFX_FIELD_P, q.data(), n), "ifx_set_field(p)");
  }
  if (a.mode == "full") check(s, ifx_iblank_update(s, nullptr), "ifx_iblank_update");
  check(s, ifx_save_field(s, IFX_FIELD_IBLANK, (a.results + "/final_results.dat").c_str()), "final_results.dat");   // preSim.cu:217

  const long nsteps = a.steps >= 0 ? a.steps : (long)in.tmax;        // tmax is a step COUNT in the reference (main.cu:93)
  auto ckpt_name = [&](long step) {
    char b[32];
    std::snprintf(b, sizeof(b), "/restart.%07ld.ifx", step);
    return a.results + b;
  };
  // Restart / Restart_Time (main.cu:27-30; read and ignored by the reference): continue from a restart file
  long first = 0;
  if (in.Restart != 0 || !a.restart.empty()) {
    const std::string f = a.restart.empty() ? ckpt_name(in.Restart_Time) : a.restart;
    long long st = 0; double t = 0.0;
    check(s, ifx_checkpoint_read(s, f.c_str(), &st, &t), "ifx_checkpoint_read");
    first = (long)st;
    std::printf("restarted from %s: step %ld, t = %g\n", f.c_str(), first, t);
  }
  // diagnostics
  std::FILE* f_forces = nullptr;
  std::FILE* f_probes = nullptr;
  std::vector<double> prx, pry, pru, prv, prp, forces(4 * 64);
  if (!a.forces.empty() && !(f_forces = std::fopen(a.forces.c_str(), first ? "a" : "w"))) die("Unable to open file: " + a.forces);
  if (!a.probes.empty()) {
    std::ifstream f(a.probes);
    if (!f) die("Unable to open file: " + a.probes);
    double x, y;
    while (f >> x >> y) { prx.push_back(x); pry.push_back(y); }
    pru.resize(prx.size()); prv.resize(prx.size()); prp.resize(prx.size());
    if (!(f_probes = std::fopen(a.probe_out.c_str(), first ? "a" : "w"))) die("Unable to open file: " + a.probe_out);
  }
  if (any_moves && first > 0) place_bodies(first * in.dt);
  std::vector<double> hist(2 * 64);
  for (long step = first; step < nsteps; step++) {
    ifx_step_stats st;
    if (any_moves) place_bodies((step + 1) * in.dt);      // where the bodies are at the end of this step
    check(s, ifx_step(s, &st), "ifx_step");
    if (f_forces) {
      check(s, ifx_body_forces(s, forces.data(), 64), "ifx_body_forces");
      for (size_t b = 0; b < bodies.size(); b++)
        std::fprintf(f_forces, "%ld %.9g %zu %.12e %.12e %.12e %.12e\n", step + 1, (step + 1) * in.dt, b, forces[4 * b],
                     forces[4 * b + 1], forces[4 * b + 2], forces[4 * b + 3]);
    }
    if (f_probes) {
      check(s, ifx_probe(s, (int)prx.size(), prx.data(), pry.data(), pru.data(), prv.data(), prp.data()), "ifx_probe");
      for (size_t k = 0; k < prx.size(); k++)
        std::fprintf(f_probes, "%ld %.9g %zu %.12e %.12e %.12e\n", step + 1, (step + 1) * in.dt, k, pru[k], prv[k], prp[k]);
    }
    if (a.reference_log) {                                           // the reference's own lines, ADSolver.cu:274,313,369
      std::printf("dt=%f\n________AD slover________\n", in.dt);
      const int n = ifx_get_residual_history(s, hist.data(), 64);
      for (int k = 0; k < n; k++) std::printf("iter = %d %f %f\n", k + 1, hist[2 * k], hist[2 * k + 1]);
    } else if (a.mode == "full") {
      std::printf("step %ld: predictor %d iterations (%.3e, %.3e), Poisson %d sweeps (residual %.3e), %.3f ms\n", step + 1,
                  st.ad_iters, st.ad_ures, st.ad_vres, st.ppe_sweeps, st.ppe_residual, st.ms_total);
    } else {
      std::printf("step %ld: predictor %d iterations (%.3e, %.3e), %.3f ms\n", step + 1, st.ad_iters, st.ad_ures, st.ad_vres, st.ms_total);
    }
    const bool interval = in.Write_Interval > 0 && (step + 1) % in.Write_Interval == 0;     // main.cu:47-50
    if (a.checkpoints && (interval || step == nsteps - 1))
      check(s, ifx_checkpoint_write(s, ckpt_name(step + 1).c_str(), step + 1, (step + 1) * in.dt), "ifx_checkpoint_write");
    if (a.write_every_step || interval || step == nsteps - 1) {
      check(s, ifx_save_field(s, IFX_FIELD_U, (a.results + "/uc.dat").c_str()), "uc.dat");       // ADSolver.cu:378
      check(s, ifx_save_field(s, IFX_FIELD_V, (a.results + "/vc.dat").c_str()), "vc.dat");       // ADSolver.cu:379
      if (a.mode == "full") check(s, ifx_save_field(s, IFX_FIELD_P, (a.results + "/p.dat").c_str()), "p.dat");   // PPESolver.cu:197
    }
  }
  if (f_forces) std::fclose(f_forces);
  if (f_probes) std::fclose(f_probes);
  std::printf("%lld kernel launches\n", ifx_launch_count(s));
  ifx_destroy(s);
  return 0;
}
