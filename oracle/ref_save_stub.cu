// ref_save_stub.cu — link-time replacement for the reference's postSim.o.
// TEST INFRASTRUCTURE ONLY.  Same signature as the reference's saveDataToFile
// (src/header/postSim.cuh:10-13); instead of two N-line ASCII files per step it dumps raw fp64
// so parity tests can compare all 53 mantissa bits, or does nothing when timing.
//   IFX_REF_SAVE = none  : no output (timing runs)
//                  last  : <filename>.f64 rewritten on every call (default)
//                  all   : <filename>.<call#>.f64 per call
//                  ascii : additionally write the reference's Tecplot ASCII via our writer
#include "postSim.cuh"
#include <cuda_runtime.h>
#include <cstdio>
#include <iostream>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

void write_results_to_file(const REALTYPE* x, const REALTYPE* y, const REALTYPE* d, int ni, int nj,
                           const char* filename) {
  FILE* fp = fopen(filename, "w");
  if (!fp) { printf("Error opening file: %s\n", filename); return; }
  fprintf(fp, "TITLE = \"Post Processing Tecplot\"\n");
  fprintf(fp, "VARIABLES = \"X\",\"Y\",\"T\"\n");
  fprintf(fp, "ZONE T=\"BIG ZONE\", I=%d, J=%d, DATAPACKING=POINT\n", ni, nj);
  for (int j = 0; j < nj; j++)
    for (int i = 0; i < ni; i++) fprintf(fp, "%f,%f,%f\n", x[i], y[j], d[i + j * ni]);
  fclose(fp);
}

void saveDataToFile(int nx, int ny, const REALTYPE* x, const REALTYPE* y,
                    const REALTYPE* dataToSave, const char* filename) {
  static std::map<std::string, int> calls;
  const char* mode = getenv("IFX_REF_SAVE");
  if (!mode) mode = "last";
  int k = calls[filename]++;
  if (!strcmp(mode, "none")) { cudaDeviceSynchronize(); return; }
  std::vector<double> h((size_t)nx * ny);
  CHECK_CUDA_ERROR(cudaMemcpy(h.data(), dataToSave, sizeof(double) * h.size(), cudaMemcpyDeviceToHost));
  char path[4096];
  if (!strcmp(mode, "all")) snprintf(path, sizeof path, "%s.%d.f64", filename, k);
  else snprintf(path, sizeof path, "%s.f64", filename);
  FILE* fp = fopen(path, "wb");
  if (fp) { fwrite(h.data(), sizeof(double), h.size(), fp); fclose(fp); }
  if (!strcmp(mode, "ascii")) {
    std::vector<double> hx(nx), hy(ny);
    cudaMemcpy(hx.data(), x, sizeof(double) * nx, cudaMemcpyDeviceToHost);
    cudaMemcpy(hy.data(), y, sizeof(double) * ny, cudaMemcpyDeviceToHost);
    write_results_to_file(hx.data(), hy.data(), h.data(), nx, ny, filename);
  }
}
