// facemask.cuh — the face mask of the general Poisson operator (closed-face rule): one byte per cell,
// derived from the cell types once per classification (kernels_facemask.cu), read by the Poisson sweeps instead of the
// types of the cell and its four neighbours.  Free of device-only constructs: shared with the host shim of the tests.
#pragma once
#include "common.cuh"

#define IFX_FM_W 1u          // the west face is open: the neighbour is a fluid cell inside the grid
#define IFX_FM_E 2u
#define IFX_FM_S 4u
#define IFX_FM_N 8u
#define IFX_FM_FLUID 16u     // the cell itself is fluid
#define IFX_FM_PLAIN 0x1fu   // plain interior cell

namespace ifx {
// rows jl_lo .. jl_hi-1 of the padded layout (the rows above and below must be classified)
cudaError_t launch_build_facemask(const Layout& L, const uint8_t* celltype, uint8_t* facemask, int jl_lo, int jl_hi,
                                  cudaStream_t st);
}  // namespace ifx
