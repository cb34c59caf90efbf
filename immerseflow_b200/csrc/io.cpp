// io.cpp — the file contract of the reference, host only (no CUDA).
//   ifx_read_input_file       <- readInputFile(), reference src/main.cu:10-59
//   ifx_read_grid_file        <- grid loops of ImmerseFlow::readGridData(), src/include/preSim.cu:268-291
//   ifx_write_results_to_file <- write_results_to_file(), src/include/postSim.cu:41-66
#include "../../include/immerseflow_c.h"

#include <cstdio>
#include <cstring>
#include <fstream>
#include <sstream>
#include <string>

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

// Tecplot ASCII POINT: three header lines, then "x,y,value" with i fastest, ghost cells included,
// "%f" (six decimals) — byte-compatible with postSim.cu:54-63 so results/plot.py still loads it.
extern "C" int ifx_write_results_to_file(const double* x, const double* y, const double* data,
                                         int ni, int nj, const char* filename) {
  if (!x || !y || !data || !filename) return IFX_ERR_INVALID;
  FILE* fp = std::fopen(filename, "w");
  if (fp == NULL) return IFX_ERR_IO;
  // a large stdio buffer: the reference's unbuffered-size fprintf loop is the slowest part of its step
  static thread_local char buf[1 << 20];
  std::setvbuf(fp, buf, _IOFBF, sizeof(buf));
  std::fprintf(fp, "TITLE = \"Post Processing Tecplot\"\n");
  std::fprintf(fp, "VARIABLES = \"X\",\"Y\",\"T\"\n");
  std::fprintf(fp, "ZONE T=\"BIG ZONE\", I=%d, J=%d, DATAPACKING=POINT\n", ni, nj);
  for (int j = 0; j < nj; j++)
    for (int i = 0; i < ni; i++)
      std::fprintf(fp, "%f,%f,%f\n", x[i], y[j], data[(size_t)i + (size_t)j * ni]);
  return std::fclose(fp) == 0 ? IFX_OK : IFX_ERR_IO;
}
