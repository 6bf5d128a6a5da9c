// ell_generic.cu -- the reference's ELL API for matrices that are NOT the 3-D / 3-field system of the hot path
// (include/ell.hpp; 2-D grids and other field counts, as exercised by the reference's own test/test_ell_1.cpp).
//
// Semantics are the reference's, with the explicit column table: y[i] = sum_j vals[i*nnz+j] * x[cols[i*nnz+j]]
// (src/ell.cpp:35-44) and the Jacobi-PCG of src/ell.cpp:66-122.  These are test-size systems (test_ell_1: 9 rows), so
// the whole solve is ONE thread block: rows strided over the threads, dot products by a fixed-order block reduction,
// no host round trip inside the loop.  It runs on the GPU like everything else in the library -- there is no CPU path.
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>

namespace {

constexpr int GT = 256;

#define GCK(call)                                                                                        \
  do {                                                                                                   \
    cudaError_t e_ = (call);                                                                             \
    if (e_ != cudaSuccess) {                                                                             \
      fprintf(stderr, "micropp-b200: CUDA error '%s' at %s:%d (%s)\n", cudaGetErrorString(e_), __FILE__, \
              __LINE__, #call);                                                                          \
      abort();                                                                                           \
    }                                                                                                    \
  } while (0)

__device__ double block_dot(const double *a, const double *b, int n, double *sm) {
  double acc = 0.0;
  for (int i = threadIdx.x; i < n; i += GT) acc += a[i] * b[i];
  sm[threadIdx.x] = acc;
  __syncthreads();
  for (int s = GT / 2; s > 0; s >>= 1) {
    if ((int)threadIdx.x < s) sm[threadIdx.x] += sm[threadIdx.x + s];
    __syncthreads();
  }
  const double r = sm[0];
  __syncthreads();
  return r;
}

__device__ void block_mvp(int nrow, int nnz, const int *cols, const double *vals, const double *x, double *y) {
  for (int i = threadIdx.x; i < nrow; i += GT) {
    double t = 0.0;
    const size_t ix = (size_t)i * nnz;
    for (int j = 0; j < nnz; ++j) t += vals[ix + j] * x[cols[ix + j]];
    y[i] = t;
  }
  __syncthreads();
}

__global__ void __launch_bounds__(GT) k_ell_generic_mvp(int nrow, int nnz, const int *cols, const double *vals,
                                                        const double *x, double *y) {
  block_mvp(nrow, nnz, cols, vals, x, y);
}

// work: k, r, z, p, Ap (nrow each); out[0] = iterations, out[1] = r.z at exit (*err of the reference)
__global__ void __launch_bounds__(GT)
    k_ell_generic_cg(int nrow, int nnz, int nfield, int shift, const int *cols, const double *vals, const double *b,
                     double *x, double *work, int max_its, double min_err, double rel_err, double *out) {
  __shared__ double sm[GT];
  double *k = work, *r = work + nrow, *z = work + 2 * (size_t)nrow, *p = work + 3 * (size_t)nrow,
         *Ap = work + 4 * (size_t)nrow;
  // k = 1 / diagonal: row (node, d) keeps its diagonal in slot shift*nfield + d (src/ell.cpp:73-76)
  for (int i = threadIdx.x; i < nrow; i += GT) {
    k[i] = 1 / vals[(size_t)i * nnz + shift * nfield + (i % nfield)];
    x[i] = 0.0;
  }
  __syncthreads();
  block_mvp(nrow, nnz, cols, vals, x, r);
  for (int i = threadIdx.x; i < nrow; i += GT) {
    r[i] = b[i] - r[i];
    z[i] = k[i] * r[i];
    p[i] = z[i];
  }
  __syncthreads();
  double rz = block_dot(r, z, nrow, sm);
  double pnorm = sqrt(block_dot(z, z, nrow, sm));
  const double pnorm0 = pnorm;
  int its = 0;
  while (its < max_its) {
    if (pnorm < min_err || pnorm < pnorm0 * rel_err) break;  // loop-head test (src/ell.cpp:93-94)
    block_mvp(nrow, nnz, cols, vals, p, Ap);
    const double alpha = rz / block_dot(p, Ap, nrow, sm);
    for (int i = threadIdx.x; i < nrow; i += GT) {
      x[i] += alpha * p[i];
      r[i] -= alpha * Ap[i];
      z[i] = k[i] * r[i];
    }
    __syncthreads();
    pnorm = sqrt(block_dot(z, z, nrow, sm));
    const double rz_n = block_dot(r, z, nrow, sm);
    const double beta = rz_n / rz;
    for (int i = threadIdx.x; i < nrow; i += GT) p[i] = z[i] + beta * p[i];
    __syncthreads();
    rz = rz_n;
    ++its;
  }
  if (threadIdx.x == 0) {
    out[0] = (double)its;
    out[1] = rz;
  }
}

struct DevBuf {
  void *p = nullptr;
  DevBuf(size_t bytes, const void *host) {
    GCK(cudaMalloc(&p, bytes ? bytes : 8));
    if (host) GCK(cudaMemcpy(p, host, bytes, cudaMemcpyHostToDevice));
  }
  ~DevBuf() { cudaFree(p); }
};

void need_device() {
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    fprintf(stderr, "micropp-b200: no CUDA device available; this library has no CPU path\n");
    abort();
  }
  GCK(cudaSetDevice(0));
}

}  // namespace

extern "C" void mgpu_ell_generic_mvp(int nrow, int nnz, const int *cols, const double *vals, const double *x,
                                     double *y) {
  need_device();
  const size_t nv = (size_t)nrow * nnz;
  DevBuf dc(nv * sizeof(int), cols), dv(nv * sizeof(double), vals), dx(nrow * sizeof(double), x),
      dy(nrow * sizeof(double), nullptr);
  k_ell_generic_mvp<<<1, GT>>>(nrow, nnz, (const int *)dc.p, (const double *)dv.p, (const double *)dx.p, (double *)dy.p);
  GCK(cudaGetLastError());
  GCK(cudaMemcpy(y, dy.p, nrow * sizeof(double), cudaMemcpyDeviceToHost));
}

extern "C" int mgpu_ell_generic_cg(int nrow, int nnz, int nfield, int shift, const int *cols, const double *vals,
                                   const double *b, double *x, int max_its, double min_err, double rel_err,
                                   double *err) {
  need_device();
  const size_t nv = (size_t)nrow * nnz;
  DevBuf dc(nv * sizeof(int), cols), dv(nv * sizeof(double), vals), db(nrow * sizeof(double), b),
      dx(nrow * sizeof(double), nullptr), dw(5 * (size_t)nrow * sizeof(double), nullptr), dout(2 * sizeof(double), nullptr);
  k_ell_generic_cg<<<1, GT>>>(nrow, nnz, nfield, shift, (const int *)dc.p, (const double *)dv.p, (const double *)db.p,
                              (double *)dx.p, (double *)dw.p, max_its, min_err, rel_err, (double *)dout.p);
  GCK(cudaGetLastError());
  double out[2];
  GCK(cudaMemcpy(out, dout.p, sizeof(out), cudaMemcpyDeviceToHost));
  GCK(cudaMemcpy(x, dx.p, nrow * sizeof(double), cudaMemcpyDeviceToHost));
  if (err) *err = out[1];
  return (int)out[0];
}
