// facemask_shim.cpp — TEST INFRASTRUCTURE: C entry point over the host-shim build of kernels_facemask.cu (see
// cuda_host_shim.h).  Arrays are host memory in the product's padded layout.
#include "cuda_host_shim.h"
thread_local dim3 threadIdx, blockIdx, blockDim, gridDim;
#include "../../immerseflow_b200/csrc/kernels_facemask.cu"

using namespace ifx;

extern "C" {
// single slab holding rows j0 .. j0 + nyl - 1 of a grid of ny rows; masks of local rows jl_lo .. jl_hi-1
void shim_build_facemask(int nx, int ny, int pitch, int nyl, int j0, const uint8_t* ct, uint8_t* fm, int jl_lo, int jl_hi) {
  launch_build_facemask(Layout{nx, ny, pitch, nyl, j0, j0 + 1, j0 + nyl - 1}, ct, fm, jl_lo, jl_hi, nullptr);
}
}
