// stencil_math.cuh — per-cell arithmetic of the reference kernels, with the FMA contraction nvcc 12.9
// emits for them (-arch=sm_100, default flags; read from the reference's PTX/SASS) written out
// explicitly.  Compiled with -fmad=false: nothing here is contracted further.  Shared by every kernel
// variant so that all of them are bit-identical to each other and to the reference build.
#pragma once
#ifndef IFX_HOST_SHIM
#include <cuda_runtime.h>
#endif

namespace ifx {

// ADusolver_kernel / ADvsolver_kernel (ADSolver.cu:91-94, :112-115):
//   t = fma(cS,qS, fma(cN,qN, fma(cW,qW, fma(cE,qE, s))));  qnew = (iBlank*t)/cP
__device__ __forceinline__ double jac_cell(double s, double cE, double qE, double cW, double qW,
                                           double cN, double qN, double cS, double qS,
                                           double ib, double cP) {
  double t = fma(cE, qE, s);
  t = fma(cW, qW, t);
  t = fma(cN, qN, t);
  t = fma(cS, qS, t);
  return (ib * t) / cP;
}

// jacobiIteration (PPESolver.cu:24-27): t = pW*cW; fma(pE,cE,t); fma(pN,cN,t); fma(pS,cS,t)
__device__ __forceinline__ double ppe_offdiag(double pW, double cW, double pE, double cE,
                                              double pN, double cN, double pS, double cS) {
  double t = pW * cW;
  t = fma(pE, cE, t);
  t = fma(pN, cN, t);
  t = fma(pS, cS, t);
  return t;
}

// Compute_Residual (PPESolver.cu:42-46): q = pE*cE; fma(p,cP,q); fma(pW,cW,q); fma(pN,cN,q); fma(pS,cS,q)
__device__ __forceinline__ double ppe_apply(double p, double cP, double pW, double cW, double pE, double cE,
                                            double pN, double cN, double pS, double cS) {
  double q = pE * cE;
  q = fma(p, cP, q);
  q = fma(pW, cW, q);
  q = fma(pN, cN, q);
  q = fma(pS, cS, q);
  return q;
}

}  // namespace ifx
