// io.cpp — the file contract of the reference, host only (no CUDA).
//   ifx_read_input_file       <- readInputFile(), reference src/main.cu:10-59
//   ifx_read_grid_file        <- grid loops of ImmerseFlow::readGridData(), src/include/preSim.cu:268-291
//   ifx_write_results_to_file <- write_results_to_file(), src/include/postSim.cu:41-66
#include "../../include/immerseflow_c.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <sstream>
#include <string>
#include <thread>
#include <vector>

extern "C" int ifx_abi_version(void) { return IFX_ABI_VERSION; }

// Same grammar as the reference: a keyword line (matched by substring, in this order: "Restart",
// "nx", "Lx", "w-AD", "ErrorMax", "Write Interval") followed by one value line; empty lines and
// lines starting with '=' or '_' are skipped (main.cu:17-52).  Unlike the reference, fields the
// file does not mention are zero instead of uninitialised, and a missing file is a status code
// instead of exit(1).
extern "C" int ifx_read_input_file(const char* path, ifx_input* in) {
  if (!path || !in) return IFX_ERR_INVALID;
  std::memset(in, 0, sizeof(*in));
  std::ifstream f(path);
  if (!f.is_open()) return IFX_ERR_IO;
  std::string line;
  while (std::getline(f, line)) {
    if (line.empty() || line[0] == '=' || line[0] == '_') continue;
    std::istringstream iss(line);
    auto value_line = [&]() {
      std::getline(f, line);
      iss.str(line);
      iss.clear();
    };
    if (line.find("Restart") != std::string::npos) {
      value_line();
      iss >> in->Restart >> in->Restart_Time;
    } else if (line.find("nx") != std::string::npos) {
      value_line();
      iss >> in->nx >> in->ny;
    } else if (line.find("Lx") != std::string::npos) {
      value_line();
      iss >> in->Lx >> in->Ly;
    } else if (line.find("w-AD") != std::string::npos) {
      value_line();
      iss >> in->w_AD >> in->w_PPE >> in->AD_itermax >> in->PPE_itermax >> in->AD_solver >> in->PPE_solver;
    } else if (line.find("ErrorMax") != std::string::npos) {
      value_line();
      iss >> in->ErrorMax >> in->tmax >> in->dt >> in->Re >> in->mu;
    } else if (line.find("Write Interval") != std::string::npos) {
      value_line();
      iss >> in->Write_Interval;
    }
  }
  // main.cu:55-58: face counts, then two ghost layers
  in->nxf = in->nx + 1;
  in->nyf = in->ny + 1;
  in->nx += 2;
  in->ny += 2;
  if (in->nx < 3 || in->ny < 3) return IFX_ERR_IO;
  return IFX_OK;
}

// n whitespace-separated "index value" pairs (Fortran list-directed or %.7E both parse), preSim.cu:275-277.
extern "C" int ifx_read_grid_file(const char* path, int n, double* faces) {
  if (!path || !faces || n <= 0) return IFX_ERR_INVALID;
  std::ifstream f(path);
  if (!f) return IFX_ERR_IO;
  int id;
  for (int i = 0; i < n; i++) {
    if (!(f >> id >> faces[i])) return IFX_ERR_IO;
  }
  return IFX_OK;
}

// ---------------------------------------------------------------------------------------------------------
// "%f" without printf.  printf("%f") prints the exact binary value rounded half-to-even at the sixth decimal;
// fmt_f6 does the same with one FMA: hi = fl(v*1e6) and lo = fma(v, 1e6, -hi) give v*1e6 exactly as hi + lo, so
// the rounding decision (including exact ties such as 1/128 = 0.0078125 -> 0.007812) is taken on the exact
// product.  Valid for |v| < 2^52 / 1e6 and finite v; everything else goes through snprintf.  Returns the length.
// ---------------------------------------------------------------------------------------------------------
static inline int fmt_f6(char* out, double v) {
  const double a = std::fabs(v);
  if (!(a < 4.0e9)) return std::snprintf(out, 400, "%f", v);       // also NaN / inf
  const double hi = a * 1e6;
  const double lo = std::fma(a, 1e6, -hi);
  double r = std::nearbyint(hi);                                   // ties-to-even on hi
  const double d = hi - r;                                         // exact: |d| <= 0.5
  if (d == 0.5) { if (lo < 0.0) r = std::floor(hi); else if (lo > 0.0) r = std::floor(hi) + 1.0; }
  else if (d == -0.5) { if (lo > 0.0) r = std::floor(hi) + 1.0; else if (lo < 0.0) r = std::floor(hi); }
  unsigned long long n = (unsigned long long)r;
  unsigned long long ip = n / 1000000ull;
  unsigned fp = (unsigned)(n % 1000000ull);
  char tmp[24];
  int k = 0;
  do { tmp[k++] = (char)('0' + ip % 10); ip /= 10; } while (ip);
  char* o = out;
  if (std::signbit(v)) *o++ = '-';
  while (k) *o++ = tmp[--k];
  *o++ = '.';
  for (int q = 5; q >= 0; --q) { o[q] = (char)('0' + fp % 10); fp /= 10; }
  o += 6;
  return (int)(o - out);
}

// Tecplot ASCII POINT: three header lines, then "x,y,value" with i fastest, ghost cells included,
// "%f" (six decimals) — byte-compatible with postSim.cu:54-63 so results/plot.py still loads it.
// at 16384^2 one field is ~8 GB of text, so the rows are formatted by all host threads into
// per-thread buffers (x and y columns pre-formatted once) and written in order; the bytes are those of the
// reference's fprintf loop.
extern "C" int ifx_write_results_to_file(const double* x, const double* y, const double* data,
                                         int ni, int nj, const char* filename) {
  if (!x || !y || !data || !filename || ni <= 0 || nj <= 0) return IFX_ERR_INVALID;
  FILE* fp = std::fopen(filename, "w");
  if (fp == NULL) return IFX_ERR_IO;
  std::fprintf(fp, "TITLE = \"Post Processing Tecplot\"\n");
  std::fprintf(fp, "VARIABLES = \"X\",\"Y\",\"T\"\n");
  std::fprintf(fp, "ZONE T=\"BIG ZONE\", I=%d, J=%d, DATAPACKING=POINT\n", ni, nj);
  // "x," strings of a row, formatted once
  std::vector<char> xs;
  std::vector<unsigned> xoff((size_t)ni + 1);
  xs.reserve((size_t)ni * 12);
  for (int i = 0; i < ni; i++) {
    char b[408];
    int n = fmt_f6(b, x[i]);
    b[n++] = ',';
    xoff[i] = (unsigned)xs.size();
    xs.insert(xs.end(), b, b + n);
  }
  xoff[ni] = (unsigned)xs.size();
  unsigned nthreads = std::thread::hardware_concurrency();
  if (nthreads == 0) nthreads = 1;
  if (const char* e = std::getenv("IFX_IO_THREADS")) nthreads = (unsigned)std::max(1, std::atoi(e));
  nthreads = std::min<unsigned>(nthreads, 64u);
  // rows per batch: ~4 MB of text per thread and batch
  const int rows_per_chunk = std::max(1, (int)((4u << 20) / ((size_t)ni * 32 + 1)));
  const int batch_rows = rows_per_chunk * (int)nthreads;
  std::vector<std::vector<char>> bufs(nthreads);
  bool ok = true;
  for (int jb = 0; jb < nj && ok; jb += batch_rows) {
    const int jend = std::min(nj, jb + batch_rows);
    auto work = [&](unsigned t) {
      std::vector<char>& out = bufs[t];
      out.clear();
      const int j0 = jb + (int)t * rows_per_chunk, j1 = std::min(jend, j0 + rows_per_chunk);
      for (int j = j0; j < j1; j++) {
        char yb[408];
        int yn = fmt_f6(yb, y[j]);
        yb[yn++] = ',';
        const double* row = data + (size_t)j * ni;
        size_t pos = out.size();
        out.resize(pos + xs.size() + (size_t)ni * (yn + 1) + (size_t)ni * 24);
        for (int i = 0; i < ni; i++) {
          if (out.size() - pos < 1300) out.resize(out.size() + (1u << 16));
          char* o = out.data() + pos;
          const unsigned xl = xoff[i + 1] - xoff[i];
          std::memcpy(o, xs.data() + xoff[i], xl); o += xl;
          std::memcpy(o, yb, yn); o += yn;
          o += fmt_f6(o, row[i]);
          *o++ = '\n';
          pos = (size_t)(o - out.data());
        }
        out.resize(pos);
      }
    };
    std::vector<std::thread> th;
    const unsigned active = (unsigned)((jend - jb + rows_per_chunk - 1) / rows_per_chunk);
    for (unsigned t = 1; t < active; t++) th.emplace_back(work, t);
    work(0);
    for (auto& q : th) q.join();
    for (unsigned t = 0; t < active && ok; t++)
      ok = bufs[t].empty() || std::fwrite(bufs[t].data(), 1, bufs[t].size(), fp) == bufs[t].size();
  }
  const bool closed = std::fclose(fp) == 0;
  return (ok && closed) ? IFX_OK : IFX_ERR_IO;
}
