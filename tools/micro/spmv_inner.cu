// Micro-benchmark of the inner loop of k_spmv_dot_tmac (implicit elastic SpMV): 9 neighbour rows x 3 dx x 3 comps x TN nodes,
// p from a shared-memory brick (LDS.128), row-block values from a __grid_constant__ parameter (LDCU + uniform operand).
// Variants isolate the cost of each ingredient:
//   0 full (LDS p + LDCU A)      1 p in registers (no LDS), LDCU A      2 LDS p, A hoisted to 9 registers      3 neither
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o spmv_inner spmv_inner.cu && ./spmv_inner
#include <cstdio>
#include <cuda_runtime.h>
constexpr int PITCH = 34, BROWS = 60, TN = 8;
struct Rows { double a[3 * 270]; };
template <int VAR, int RU>
__global__ void __launch_bounds__(128) k(const __grid_constant__ Rows R, double *out, int iters, const int *sel) {
  extern __shared__ __align__(16) double sb[];
  for (int q = threadIdx.x; q < 3 * BROWS * PITCH; q += 128) sb[q] = 1e-3 * (q % 97);
  __syncthreads();
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, ry = lane & 7, rz = lane >> 3;
  double acc[TN][3];
#pragma unroll
  for (int t = 0; t < TN; ++t) acc[t][0] = acc[t][1] = acc[t][2] = 0.0;
  double pvf[3][10];
#pragma unroll
  for (int d = 0; d < 3; ++d)
#pragma unroll
    for (int h = 0; h < 10; ++h) pvf[d][h] = sb[d * 10 + h + lane];
  double ah[9];
#pragma unroll
  for (int q = 0; q < 9; ++q) ah[q] = R.a[q];
  const int mat = sel[blockIdx.x & 1];
#pragma unroll 1
  for (int it = 0; it < iters; ++it) {
#pragma unroll RU
    for (int row = 0; row < 9; ++row) {
      const int dk = row / 3, dj = row - dk * 3;
      const int rbase = ((rz + dk) * 10 + (ry + dj)) * PITCH + 8 * w;
      double pv[3][10];
#pragma unroll
      for (int d = 0; d < 3; ++d) {
        if (VAR == 0 || VAR == 2) {
          const double2 *s2 = reinterpret_cast<const double2 *>(sb + d * (BROWS * PITCH) + rbase);
#pragma unroll
          for (int h = 0; h < 5; ++h) { const double2 v = s2[h]; pv[d][2 * h] = v.x; pv[d][2 * h + 1] = v.y; }
        } else {
#pragma unroll
          for (int h = 0; h < 10; ++h) pv[d][h] = pvf[d][h];
        }
      }
#pragma unroll
      for (int di = 0; di < 3; ++di) {
        const double *a = &R.a[(row * 3 + di) * 10];
#pragma unroll
        for (int fj = 0; fj < 3; ++fj)
#pragma unroll
          for (int t = 0; t < TN; ++t) {
            const double pval = pv[fj][t + di];
            if (VAR == 0 || VAR == 1) {
              acc[t][0] += a[fj] * pval; acc[t][1] += a[3 + fj] * pval; acc[t][2] += a[6 + fj] * pval;
            } else {
              acc[t][0] += ah[fj] * pval; acc[t][1] += ah[3 + fj] * pval; acc[t][2] += ah[6 + fj] * pval;
            }
          }
      }
    }
    if (mat == 12345) pvf[0][0] += 1.0;  // keeps the loop body from being hoisted
  }
  double s = 0;
#pragma unroll
  for (int t = 0; t < TN; ++t) s += acc[t][0] + acc[t][1] + acc[t][2];
  out[blockIdx.x * 128 + threadIdx.x] = s;
}
template <int VAR, int RU>
void run(const Rows &R, double *d, const int *sel, const char *name) {
  const int smem = 3 * BROWS * PITCH * 8;
  cudaFuncSetAttribute(k<VAR, RU>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  const int iters = 400;
  for (int bps = 1; bps <= 4; ++bps) {
    const int grid = 148 * bps;
    float ms = 0;
    for (int rep = 0; rep < 2; ++rep) {
      cudaEventRecord(a); k<VAR, RU><<<grid, 128, smem>>>(R, d, iters, sel); cudaEventRecord(b); cudaEventSynchronize(b);
      cudaEventElapsedTime(&ms, a, b);
    }
    const double flop = 2.0 * 243 * TN * 128.0 * grid * iters;
    printf("%-44s unroll %d blocks/SM %d: %8.3f ms %6.2f TFLOP/s (%s)\n", name, RU, bps, ms, flop / ms * 1e-9,
           cudaGetErrorString(cudaGetLastError()));
  }
}
int main() {
  Rows R; for (int i = 0; i < 810; ++i) R.a[i] = 1.0 + 1e-6 * i;
  double *d; cudaMalloc(&d, 148 * 4 * 128 * 8);
  int hs[2] = {0, 0}, *sel; cudaMalloc(&sel, 8); cudaMemcpy(sel, hs, 8, cudaMemcpyHostToDevice);
  run<0, 1>(R, d, sel, "full: LDS p + LDCU A");
  run<0, 3>(R, d, sel, "full: LDS p + LDCU A");
  run<0, 9>(R, d, sel, "full: LDS p + LDCU A");
  run<1, 3>(R, d, sel, "p in registers + LDCU A");
  run<2, 3>(R, d, sel, "LDS p + A in 9 registers");
  run<3, 3>(R, d, sel, "p in registers + A in 9 registers");
  return 0;
}
