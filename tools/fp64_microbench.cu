// fp64_microbench.cu — measures what bounds the sweep kernels' arithmetic on B200 (sm_100a):
// DFMA issue throughput and dependent latency, IEEE fp64 division throughput and latency.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false tools/fp64_microbench.cu -o gpurun_out/fp64_microbench
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP>
__global__ void k_dfma(double* out, double a, double b, int iters) {
  double x[ILP];
#pragma unroll
  for (int q = 0; q < ILP; ++q) x[q] = threadIdx.x * 1e-3 + q;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int q = 0; q < ILP; ++q) x[q] = fma(x[q], a, b);
  }
  double s = 0;
#pragma unroll
  for (int q = 0; q < ILP; ++q) s += x[q];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int ILP>
__global__ void k_ddiv(double* out, double a, double b, int iters) {
  double x[ILP];
#pragma unroll
  for (int q = 0; q < ILP; ++q) x[q] = 1.0 + threadIdx.x * 1e-3 + q;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int q = 0; q < ILP; ++q) x[q] = (x[q] + b) / a;
  }
  double s = 0;
#pragma unroll
  for (int q = 0; q < ILP; ++q) s += x[q];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
float time_ms(F f) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  f();
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  f();
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  return ms;
}

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  int sms = p.multiProcessorCount;
  double* out; cudaMalloc(&out, sizeof(double) * sms * 32 * 1024);
  const int iters = 4096;
  printf("device %s, %d SMs\n", p.name, sms);
  int clk_khz; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
  printf("clock attr %d kHz\n", clk_khz);
  // throughput: many warps, ILP 8
  for (int warps : {1, 2, 4, 8, 16, 32}) {
    float ms = time_ms([&] { k_dfma<8><<<sms, warps * 32>>>(out, 1.0000001, 1e-9, iters); });
    double inst = (double)sms * warps * 8.0 * iters;           // warp-instructions
    printf("DFMA ILP8 warps/SM=%2d : %.3f ms  -> %.2f warp-instr/ns total, %.3f warp-instr/clk/SM @1.9GHz, %.1f TFLOPS\n", warps, ms,
           inst / (ms * 1e6), inst / (ms * 1e6) / sms / 1.9, inst * 64 / (ms * 1e-3) / 1e12);
  }
  {
    float ms = time_ms([&] { k_dfma<1><<<sms, 32>>>(out, 1.0000001, 1e-9, iters); });
    printf("DFMA dependent chain, 1 warp/SM: %.3f ms -> %.1f ns per DFMA (%.1f cycles @1.9GHz)\n", ms, ms * 1e6 / iters, ms * 1e6 / iters * 1.9);
  }
  for (int warps : {1, 4, 8, 16, 32}) {
    float ms = time_ms([&] { k_ddiv<4><<<sms, warps * 32>>>(out, 1.0000001, 1e-9, iters / 4); });
    double n = (double)sms * warps * 32 * 4.0 * (iters / 4);   // divisions
    printf("DDIV ILP4 warps/SM=%2d : %.3f ms -> %.2f Gdiv/s total, %.2f div/clk/SM @1.9GHz\n", warps, ms, n / (ms * 1e6), n / (ms * 1e6) / sms / 1.9);
  }
  {
    float ms = time_ms([&] { k_ddiv<1><<<sms, 32>>>(out, 1.0000001, 1e-9, iters / 4); });
    printf("DDIV dependent chain: %.1f ns per (add+div) (%.0f cycles @1.9GHz)\n", ms * 1e6 / (iters / 4), ms * 1e6 / (iters / 4) * 1.9);
  }
  return 0;
}
