// Measures the sustained DFMA issue rate of one B200 (the roofline of the fp64-bound QUADPACK kernel):
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dfma_peak dfma_peak.cu && ./dfma_peak
// CH independent DFMA chains per thread, register operands; optional IMAD filler per DFMA (MIX) shows how many
// non-fp64 instructions the scheduler can issue in the shadow of the fp64 pipe.
#include <cstdio>
#include <cuda_runtime.h>

template <int CH, int MIX>
__global__ void __launch_bounds__(256) k(double* out, int iters, double a, double b, int m) {
  double x[CH];
  int y[MIX > 0 ? MIX : 1];
#pragma unroll
  for (int c = 0; c < CH; ++c) x[c] = threadIdx.x * 1e-9 + c;
#pragma unroll
  for (int c = 0; c < (MIX > 0 ? MIX : 1); ++c) y[c] = threadIdx.x + c;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int c = 0; c < CH; ++c) {
      x[c] = fma(x[c], a, b);
#pragma unroll
      for (int q = 0; q < MIX; ++q) y[q] = y[q] * m + q;
    }
  }
  double s = 0;
#pragma unroll
  for (int c = 0; c < CH; ++c) s += x[c];
  int t = 0;
#pragma unroll
  for (int c = 0; c < (MIX > 0 ? MIX : 1); ++c) t += y[c];
  if (s == 1.2345 || t == 12345) out[0] = s + t;
}

template <int CH, int MIX>
void run(int blocks_per_sm, int nsm, double clock_ghz) {
  double* d;
  cudaMalloc(&d, 8);
  int iters = 20000;
  int blocks = blocks_per_sm * nsm;
  k<CH, MIX><<<blocks, 256>>>(d, 100, 1.0000001, 1e-9, 3);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  k<CH, MIX><<<blocks, 256>>>(d, iters, 1.0000001, 1e-9, 3);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double dfma = (double)blocks * 256 * iters * CH;
  double per_clk_sm = dfma / (ms * 1e-3) / (clock_ghz * 1e9) / nsm;
  printf("{\"chains\": %d, \"imad_per_dfma\": %d, \"blocks_per_sm\": %d, \"ms\": %.3f, \"tflops_fp64\": %.2f, \"dfma_lanes_per_clk_per_sm\": %.1f}\n",
         CH, MIX, blocks_per_sm, ms, 2 * dfma / (ms * 1e-3) / 1e12, per_clk_sm);
  cudaFree(d);
}

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  int nsm = p.multiProcessorCount;
  int khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  double ghz = khz / 1e6;
  printf("{\"device\": \"%s\", \"sms\": %d, \"clock_ghz\": %.3f}\n", p.name, nsm, ghz);
  run<1, 0>(8, nsm, ghz);
  run<2, 0>(8, nsm, ghz);
  run<4, 0>(8, nsm, ghz);
  run<8, 0>(8, nsm, ghz);
  run<8, 0>(4, nsm, ghz);
  run<8, 0>(2, nsm, ghz);
  run<8, 1>(8, nsm, ghz);
  run<8, 2>(8, nsm, ghz);
  run<4, 1>(3, nsm, ghz);
  return 0;
}
