// cuda_host_shim.h — TEST INFRASTRUCTURE.  Lets the thread-per-cell CUDA sources of the product
// (immerseflow_b200/csrc/kernels_mg.cu, kernels_diag.cu: no shared memory, no warp intrinsics, no atomics) compile as
// plain C++ and run their launch grids serially on the CPU, so the CPU test suite checks their arithmetic, indexing
// and launch geometry against the oracle.  The product is never built this way (IFX_HOST_SHIM is only defined by
// tests/shim/build.sh), and nothing here is a fallback: it exists because the development container has no GPU.
#pragma once
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstring>

struct dim3 {
  unsigned x, y, z;
  dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {}
};
struct double2 { double x, y; };
struct uint4 { unsigned x, y, z, w; };
inline uint4 make_uint4(unsigned a, unsigned b, unsigned c, unsigned d) { return uint4{a, b, c, d}; }
typedef void* cudaStream_t;
typedef int cudaError_t;
enum { cudaSuccess = 0 };

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__

extern thread_local dim3 threadIdx, blockIdx, blockDim, gridDim;

inline int min(int a, int b) { return a < b ? a : b; }      // CUDA's global overloads
inline int max(int a, int b) { return a > b ? a : b; }

template <class F>
inline void shim_launch(dim3 grid, dim3 block, F body) {
  gridDim = grid; blockDim = block;
  for (unsigned bz = 0; bz < grid.z; bz++)
    for (unsigned by = 0; by < grid.y; by++)
      for (unsigned bx = 0; bx < grid.x; bx++) {
        blockIdx = dim3(bx, by, bz);
        for (unsigned tz = 0; tz < block.z; tz++)
          for (unsigned ty = 0; ty < block.y; ty++)
            for (unsigned tx = 0; tx < block.x; tx++) {
              threadIdx = dim3(tx, ty, tz);
              body();
            }
      }
}
