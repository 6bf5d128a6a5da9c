// Micro-benchmark: sustained DFMA rate of one B200 (non-tensor FP64 pipe), the denominator of the FP64 fraction
// reported for the implicit elastic SpMV.  24 independent accumulators per thread (as in k_spmv_dot_tmac), operands
// either registers only or one uniform/constant operand.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dfma_peak dfma_peak.cu && ./dfma_peak
#include <cstdio>
#include <cuda_runtime.h>
__constant__ double cA[32];
template <int MODE>
__global__ void __launch_bounds__(128) k(double *out, int iters, double seed) {
  double acc[24], p[8];
#pragma unroll
  for (int i = 0; i < 24; ++i) acc[i] = seed * (i + threadIdx.x);
#pragma unroll
  for (int i = 0; i < 8; ++i) p[i] = seed + i * 1e-9;
#pragma unroll 1
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
#pragma unroll
      for (int i = 0; i < 24; ++i) {
        if (MODE == 0)
          acc[i] = fma(acc[(i + 5) % 24 == i ? i : i], p[(i + r) & 7], acc[i]) ;
        else
          acc[i] = fma(p[(i + r) & 7], cA[(r * 3 + i % 3)], acc[i]);
      }
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 24; ++i) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
  double h[32];
  for (int i = 0; i < 32; ++i) h[i] = 1.0 + 1e-9 * i;
  cudaMemcpyToSymbol(cA, h, sizeof(h));
  double *d;
  cudaMalloc(&d, 148 * 16 * 128 * sizeof(double));
  cudaEvent_t a, b;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  const int iters = 20000;
  for (int mode = 0; mode < 2; ++mode)
    for (int bps = 1; bps <= 8; bps *= 2) {
      const int grid = 148 * bps;
      for (int rep = 0; rep < 2; ++rep) {
        cudaEventRecord(a);
        if (mode == 0) k<0><<<grid, 128>>>(d, iters, 1e-3); else k<1><<<grid, 128>>>(d, iters, 1e-3);
        cudaEventRecord(b);
        cudaEventSynchronize(b);
      }
      float ms;
      cudaEventElapsedTime(&ms, a, b);
      const double flop = 2.0 * 96 * iters * 128.0 * grid;
      printf("mode %d (%s) blocks/SM %d (%d warps/SM): %.3f ms  %.2f TFLOP/s\n", mode, mode ? "reg x const" : "reg x reg",
             bps, bps * 4, ms, flop / ms * 1e-9);
    }
  return 0;
}
