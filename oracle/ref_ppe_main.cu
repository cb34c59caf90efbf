// ref_ppe_main.cu — tiny driver for the reference's (repaired) ImmerseFlow::PPESolver(),
// which the reference's own main never calls (src/main.cu:98 is commented out).
// TEST INFRASTRUCTURE ONLY.  Mirrors src/main.cu:75-101 with PPESolver() in place of the
// predictor loop; readInputFile is the same keyword/next-line grammar (src/main.cu:10-59),
// restated here because main.cu cannot be linked twice.
#include "globalVariables.cuh"
#include <cuda_runtime.h>
#include <fstream>
#include <iostream>
#include <sstream>
#include <string>

static void read_input(const std::string& fn, ImmerseFlow& S) {
  std::ifstream f(fn);
  if (!f.is_open()) { std::cerr << "Unable to open file: " << fn << std::endl; exit(1); }
  std::string line;
  auto next = [&](std::istringstream& iss) { getline(f, line); iss.str(line); iss.clear(); };
  while (getline(f, line)) {
    if (line.empty() || line[0] == '=' || line[0] == '_') continue;
    std::istringstream iss(line);
    if (line.find("Restart") != std::string::npos) { next(iss); iss >> S.Input.Restart >> S.Input.Restart_Time; }
    else if (line.find("nx") != std::string::npos) { next(iss); iss >> S.Input.nx >> S.Input.ny; }
    else if (line.find("Lx") != std::string::npos) { next(iss); iss >> S.Input.Lx >> S.Input.Ly; }
    else if (line.find("w-AD") != std::string::npos) {
      next(iss);
      iss >> S.Input.w_AD >> S.Input.w_PPE >> S.Input.AD_itermax >> S.Input.PPE_itermax >> S.Input.AD_solver >> S.Input.PPE_solver;
    } else if (line.find("ErrorMax") != std::string::npos) {
      next(iss);
      iss >> S.Input.ErrorMax >> S.Input.tmax >> S.Input.dt >> S.Input.Re >> S.Input.mu;
    } else if (line.find("Write Interval") != std::string::npos) { next(iss); iss >> S.Input.Write_Interval; }
  }
  S.Input.nxf = S.Input.nx + 1; S.Input.nyf = S.Input.ny + 1;
  S.Input.nx += 2; S.Input.ny += 2;
}

int main() {
  ImmerseFlow Solver;
  read_input("../inputs/inputs.txt", Solver);
  Solver.CUDAQuery();
  Solver.allocation();
  Solver.readGridData();
  Solver.initializeData();
  Solver.PPESolver();
  cudaDeviceSynchronize();
  return 0;
}
