// mgpu_kernels.cu -- hand-written sm_100a kernels of the RVE-homogenization hot path and the
// thin C-ABI (include/mgpu.h) that launches them.  FP64 throughout; no tensor cores (no stage is
// a dense contraction); every kernel is either HBM-bound (SpMV / CG vector updates / gather
// assembly of elastic RVEs) or FP64-pipe-bound (forward-difference tangent + B^T C B).
//
// Design in one paragraph: a *wave* of W RVEs (one per macro Gauss point) is resident in HBM.
// Every kernel is launched over (node- or element-blocks) x (a compacted list of slots), so one
// launch advances all RVEs of the wave by one algorithmic step.  Per-slot scalars (CG alpha/beta,
// norms, iteration counters, convergence flags) never leave the device inside a solve: the last
// block of each reducing kernel (ticket counter) folds the per-block partial sums in a fixed order
// and applies the reference's scalar logic (src/ell.cpp:86-119, src/solve.cpp:43-78).  Reductions
// are therefore deterministic and there is not a single floating-point atomic in the file.
//
// Assembly is a *gather by node* instead of the reference's element scatter: each ELL value is
// written exactly once, coalesced, with the element contributions added in the reference's element
// visiting order -- no atomics, no colours, no read-modify-write traffic.
//
// Map of the file (DESIGN.md section 5 has bytes, bounds and measurements of every kernel):
//   k_set_bc, k_elem_rhs + k_asm_rhs        boundary displacements, residual (per-element + node gather)
//   k_asm_mat_elastic, k_elem_ctan + k_asm_mat_general, k_rows_build     Jacobian (assembled) / row-block table
//   k_cg_init, k_spmv_dot, k_cg_update + k_fold_update, k_cg_pupdate, k_cg_finish      DPCG, assembled operator
//   k_spmv_dot_tmac (+ _tma, _tile, _imp), k_fold_spmv, k_cg_update_imp, k_cg_pupdate_imp   DPCG, implicit operator
//                                           of all-elastic RVEs (no per-RVE matrix; TMA-tiled, FP64-pipe-bound)
//   k_axpy_u, k_ave_stress, k_vars_new, k_elem_fields, k_compact        Newton update, averages, history, lists
//   k_slab_*                                z-slab mode of one large RVE over NVLink peer memory
//   mgpu_*                                  the C ABI; mgpu_newton_step_graph = one Newton step as one CUDA graph
#include <cuda.h>  // CUtensorMap (type + enums only; the encoder is fetched through cudaGetDriverEntryPoint)
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <vector>

#include "fe_math.cuh"
#include "mgpu.h"

#define CK(call)                                                                                         \
  do {                                                                                                   \
    cudaError_t e_ = (call);                                                                             \
    if (e_ != cudaSuccess) {                                                                             \
      fprintf(stderr, "micropp-b200: CUDA error '%s' at %s:%d (%s)\n", cudaGetErrorString(e_), __FILE__, \
              __LINE__, #call);                                                                          \
      abort();                                                                                           \
    }                                                                                                    \
  } while (0)

namespace {

constexpr int NT = 128;      // threads per block of node/element kernels
constexpr int NPLANE = 243;  // 27 neighbours x 3 x 3
// row blocks of the implicit operator: [27 neighbours][10] doubles, the 3x3 block of a neighbour in the first 9 --
// 80-B groups are 16-B aligned, so a neighbour's block is five 128-bit loads (global or shared)
constexpr int RB_NBR = 10, RB_LEN = 27 * RB_NBR;
constexpr int NLIST = 6;
constexpr int NRED = 6;      // max values reduced per kernel

struct MeshConst {
  int nx, ny, nz, nxny, nn, nn_pad;
  int nix, niy, niz, nint, nint_pad;  // interior nodes (the only rows the ELL storage keeps)
  // z-slab of a larger RVE (single-RVE domain decomposition): local plane k is global plane k + koff; the first /
  // last local plane is a halo plane (owned by the neighbour rank) when halo_lo / halo_hi is set, else a true face.
  // Reductions then stop at the slab-local sum (T.red) and the scalar tails run after the cross-rank all-reduce.
  int slab, koff, nz_glob, halo_lo, halo_hi, ez_own_lo, ez_own_hi;
  int nex, ney, nez, nelem, nelem_pad;
  int nvar;
  int nr_max_its, cg_max_its;
  double dx, dy, dz, wg;
  double nr_max_tol, nr_rel_tol, cg_abs_tol, cg_rel_tol;
  double dsh[8][24];
  mpp_material mat[3];
};

struct SlotTables {  // device arrays, one entry per slot
  mgpu_slot_state *state;
  const double **vars_old;
  double **vars_new;
  double **u_n;
  double **u_k;
  double *eps;     // [W][6]
  double *stress;  // [W][6]
  double *partial; // [W][NRED][nblk_max]
  double *red;     // [W][8] slab-local sums handed to the all-reduce (slab mode)
  int nblk_max;
};

struct VecPool {
  double *u, *b, *du, *k, *r, *z, *p, *Ap;  // [W][3*nn_pad]
  double *mat;                              // [W][243*nint_pad], interior rows, 32-node tiles
  double *mat_shared;                       // [243*nint_pad] (A0)
  double *gen;                              // [3*nn][81] host matrix of the generic ELL API (reference layout)
  // implicit operator of an all-elastic RVE: the ELL row block of an interior node is a pure function of the
  // materials of its 8 elements, so only the DISTINCT row blocks are kept (rows[id][243]) plus one id per node
  const double *rows;                       // [nrows][243]
  const double *rkinv;                      // [nrows][3]  1 / diagonal (the Jacobi preconditioner, src/ell.cpp:73-76)
  const int *rowid;                         // [nint_pad]
  size_t vstride, mstride;
};

// operator selector of the DPCG kernels
enum { OP_SLOT = 0, OP_SHARED = 1, OP_GENERIC = 2, OP_IMPLICIT = 3 };

// ------------------------------------------------------------------------------------------------
// small device helpers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void node_ijk(const MeshConst &P, int n, int &i, int &j, int &k) {
  k = n / P.nxny;
  const int r = n - k * P.nxny;
  j = r / P.nx;
  i = r - j * P.nx;
}
// interior-node index m (x fastest) -> grid coordinates and global node id
__device__ __forceinline__ int interior_node(const MeshConst &P, int m, int &i, int &j, int &k) {
  const int pl = P.nix * P.niy;
  const int kk = m / pl, r = m - kk * pl, jj = r / P.nix;
  i = r - jj * P.nix + 1;
  j = jj + 1;
  k = kk + 1;
  return k * P.nxny + j * P.nx + i;
}
__device__ __forceinline__ int interior_index(const MeshConst &P, int i, int j, int k) {
  return ((k - 1) * P.niy + (j - 1)) * P.nix + (i - 1);
}
__device__ __forceinline__ bool on_boundary(const MeshConst &P, int i, int j, int k) {
  return i == 0 || i == P.nx - 1 || j == 0 || j == P.ny - 1 || k == 0 || k == P.nz - 1;
}

// ELL values of one RVE are stored for INTERIOR nodes only (boundary rows are identity rows, ell_set_bc_3D
// src/ell-common.cpp:238-297, and are never read) in tiles of 32 consecutive interior nodes: [tile][243 planes][32].  A warp that owns
// one tile streams a single contiguous 62 KB chunk (plane after plane, 256 B per load, immediate offsets from one
// base register) -- DRAM page locality does not depend on how the compiler schedules the 243 loads.
__host__ __device__ __forceinline__ size_t aidx(int plane, int node) {
  return ((size_t)(node >> 5) * NPLANE + plane) * 32 + (node & 31);
}

// A batched kernel runs over (blocks) x (entries of a slot list).  `dcount` (optional) is a device-side entry
// count: inside a captured CUDA graph the launch shape is fixed while the number of still-active slots shrinks,
// so surplus blocks leave at once.  `yoff` is the offset of this launch inside the list (chunked launches).
struct Lst {
  const int *list;
  const int *dcount;
  int yoff;
};
__device__ __forceinline__ int slot_of(const Lst &L) {
  const int y = (int)blockIdx.y + L.yoff;
  if (L.dcount && y >= *L.dcount) return -1;
  return L.list[y];
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  return v;
}

// Sum NV values over the block; result valid in thread 0.  Fixed tree => deterministic.
template <int NV>
__device__ __forceinline__ void block_sum(double (&v)[NV], double *sm /* [NV][NT/32] */) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
  for (int q = 0; q < NV; ++q) {
    const double s = warp_sum(v[q]);
    if (lane == 0) sm[q * (NT / 32) + w] = s;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
#pragma unroll
    for (int q = 0; q < NV; ++q) {
      double s = 0.0;
#pragma unroll
      for (int ww = 0; ww < NT / 32; ++ww) s += sm[q * (NT / 32) + ww];
      v[q] = s;
    }
  }
  __syncthreads();
}

// Grid-wide deterministic reduction with a ticket: every block deposits its partial sums; the block
// that draws the last ticket re-reduces all partials in a fixed order.  Returns true in every thread
// of that last block; totals valid in its thread 0.
template <int NV>
__device__ __forceinline__ bool grid_sum(double (&v)[NV], double *partial, int pstride, unsigned *ticket, double *sm,
                                         int *sflag) {
  block_sum<NV>(v, sm);
  const int nblk = gridDim.x;
  if (threadIdx.x == 0) {
#pragma unroll
    for (int q = 0; q < NV; ++q) partial[q * pstride + blockIdx.x] = v[q];
    __threadfence();
    const unsigned t = atomicAdd(ticket, 1u);
    *sflag = (t == (unsigned)(nblk - 1));
  }
  __syncthreads();
  if (!*sflag) return false;
  __threadfence();
#pragma unroll
  for (int q = 0; q < NV; ++q) {
    double acc = 0.0;
    for (int b = threadIdx.x; b < nblk; b += NT) acc += __ldcg(&partial[q * pstride + b]);
    v[q] = acc;
  }
  block_sum<NV>(v, sm);
  if (threadIdx.x == 0) *ticket = 0u;
  return true;
}

__device__ __forceinline__ const double *fetch_vars(const double *vbase, int nelem_pad, int e, int gp, int nv,
                                                    double *buf) {
  if (!vbase) return nullptr;
#pragma unroll
  for (int q = 0; q < 7; ++q) buf[q] = (q < nv) ? __ldg(&vbase[(size_t)(q * 8 + gp) * nelem_pad + e]) : 0.0;
  return buf;
}

__device__ __forceinline__ void gather_ue(const MeshConst &P, const double *__restrict__ u, int ex, int ey, int ez,
                                          double *ue) {
  const int n0 = ez * P.nxny + ey * P.nx + ex;
#pragma unroll
  for (int a = 0; a < 8; ++a) {
    const int nd = n0 + corner_x(a) + corner_y(a) * P.nx + corner_z(a) * P.nxny;  // src/common.cpp:30-41
#pragma unroll
    for (int d = 0; d < 3; ++d) ue[a * 3 + d] = u[(size_t)d * P.nn_pad + nd];
  }
}

// ------------------------------------------------------------------------------------------------
// scalar tails of the reducing kernels: the reference's per-solve scalar logic, one thread per slot.
// In slab mode they run from k_tail after the cross-rank all-reduce of T.red.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ bool true_boundary(const MeshConst &P, int i, int j, int k) {
  const int kg = k + P.koff;
  return i == 0 || i == P.nx - 1 || j == 0 || j == P.ny - 1 || kg == 0 || kg == P.nz_glob - 1;
}

// mode 0: first residual of a Newton solve (sets norm0, its=0); 1: after an update (its++); 2: plain
__device__ __forceinline__ void tail_rhs(const MeshConst &P, mgpu_slot_state *st, double nrm2, int mode) {
  const double norm = sqrt(nrm2);
  st->norm = norm;
  if (mode == 2) return;
  int its;
  if (mode == 0) {
    st->norm0 = norm;
    st->nr_its = its = 0;
    st->solver_its = 0;
    st->converged = 0;
  } else {
    its = ++st->nr_its;
  }
  // loop head of src/solve.cpp:43-47 -- no test once nr_max_its solves have been spent
  int active = 0;
  if (its < P.nr_max_its) {
    if (norm < P.nr_max_tol || norm < st->norm0 * P.nr_rel_tol)
      st->converged = 1;
    else
      active = 1;
  }
  st->nr_active = active;
}
__device__ __forceinline__ void tail_cg_init(const MeshConst &P, mgpu_slot_state *st, double rz, double zz) {
  const double pn = sqrt(zz);
  st->rz = rz;
  st->pnorm0 = pn;
  st->pnorm = pn;
  st->cg_its = 0;
  // loop head of src/ell.cpp:93-94
  st->cg_active = (0 < P.cg_max_its) && !(pn < P.cg_abs_tol || pn < pn * P.cg_rel_tol);
}
__device__ __forceinline__ void tail_spmv(mgpu_slot_state *st, double pAp) {
  st->pAp = pAp;
  st->alpha = st->rz / pAp;  // src/ell.cpp:100
}
__device__ __forceinline__ void tail_cg_update(const MeshConst &P, mgpu_slot_state *st, double zz, double rz_n) {
  const double pn = sqrt(zz);
  st->pnorm = pn;
  st->beta = rz_n / st->rz;
  st->rz = rz_n;
  const int its = ++st->cg_its;
  st->cg_active = (its < P.cg_max_its) && !(pn < P.cg_abs_tol || pn < st->pnorm0 * P.cg_rel_tol);
}

// ------------------------------------------------------------------------------------------------
// u <- u_n / u_k ; u_k <- u
// ------------------------------------------------------------------------------------------------
__global__ void k_load_u(MeshConst P, const Lst L, SlotTables T, double *u_pool, size_t vstride,
                         int which) {
  const int slot = slot_of(L);
  if (slot < 0) return;
  const double *src = which ? T.u_k[slot] : T.u_n[slot];
  double *dst = u_pool + (size_t)slot * vstride;
  const int len = 3 * P.nn_pad;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < len; i += gridDim.x * blockDim.x) dst[i] = src[i];
}

__global__ void k_store_u(MeshConst P, const Lst L, SlotTables T, const double *u_pool,
                          size_t vstride, int which) {
  const int slot = slot_of(L);
  if (slot < 0) return;
  double *dst = which ? T.u_k[slot] : T.u_n[slot];
  const double *src = u_pool + (size_t)slot * vstride;
  const int len = 3 * P.nn_pad;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < len; i += gridDim.x * blockDim.x) dst[i] = src[i];
}

__global__ void k_zero_u(MeshConst P, const Lst L, double *u_pool, size_t vstride) {
  const int slot = slot_of(L);
  if (slot < 0) return;
  double *dst = u_pool + (size_t)slot * vstride;
  const int len = 3 * P.nn_pad;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < len; i += gridDim.x * blockDim.x) dst[i] = 0.0;
}

// ------------------------------------------------------------------------------------------------
// set_displ_bc (src/micro3D.cpp:27-78)
// ------------------------------------------------------------------------------------------------
__global__ void k_set_bc(const __grid_constant__ MeshConst P, const Lst L, SlotTables T,
                         double *u_pool, size_t vstride) {
  const int slot = slot_of(L);
  if (slot < 0) return;
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= P.nn) return;
  int i, j, k;
  node_ijk(P, n, i, j, k);
  if (!true_boundary(P, i, j, k)) return;  // halo planes of a slab are interior nodes of the RVE
  double eps[6], c[3], u3[3];
#pragma unroll
  for (int q = 0; q < 6; ++q) eps[q] = T.eps[slot * 6 + q];
  bc_coords(i, j, k + P.koff, P.nx, P.ny, P.nz_glob, P.dx, P.dy, P.dz, c);
  bc_displacement(eps, c, u3);
  double *u = u_pool + (size_t)slot * vstride;
#pragma unroll
  for (int d = 0; d < 3; ++d) u[(size_t)d * P.nn_pad + n] = u3[d];
}

// ------------------------------------------------------------------------------------------------
// assembly_rhs (src/assembly.cpp:28-103) in two kernels.
// (1) k_elem_rhs: thread per element: be = sum_gp B^T sigma wg (get_elem_rhs, src/assembly.cpp:124-138),
//     every Gauss-point stress evaluated exactly once, stored as [24][nelem_pad] per slot of the chunk.
// (2) k_asm_rhs: thread per node: adds the contributions of the 8 surrounding elements in the
//     reference's element visiting order (ez outer, ex inner), zeroes boundary rows, negates, and
//     reduces ||b||^2 (deterministic ticket reduction) with the Newton loop-head logic in its tail.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(NT)
    k_elem_rhs(const __grid_constant__ MeshConst P, const Lst L, SlotTables T,
               const double *__restrict__ u_pool, size_t vstride, const int *__restrict__ elem_type,
               double *__restrict__ bebuf, size_t bstride, int mode) {
  const int slot = slot_of(L);
  if (slot < 0) return;
  if (mode == 1 && !T.state[slot].nr_active) return;
  const double *u = u_pool + (size_t)slot * vstride;
  const double *vars = T.vars_old[slot];
  double *be = bebuf + (size_t)blockIdx.y * bstride;
  const int e = blockIdx.x * NT + threadIdx.x;
  if (e >= P.nelem) return;
  const int ez = e / (P.nex * P.ney);
  const int r = e - ez * P.nex * P.ney;
  const int ey = r / P.nex, ex = r - ey * P.nex;
  double ue[24], acc[24];
  gather_ue(P, u, ex, ey, ez, ue);
#pragma unroll
  for (int q = 0; q < 24; ++q) acc[q] = 0.0;
  const int type = __ldg(&elem_type[e]);
  const mpp_material m = P.mat[type];
  const int nv = mat_nvar(m.type);
  const double wg = P.wg;
#pragma unroll 1
  for (int gp = 0; gp < 8; ++gp) {
    double eps[6], sig[6], vbuf[7];
    gp_strain(P.dsh[gp], ue, eps);
    const double *v = fetch_vars(vars, P.nelem_pad, e, gp, nv, vbuf);
    mat_stress(m, eps, v, sig);
#pragma unroll
    for (int a = 0; a < 8; ++a) {
      const double gx = P.dsh[gp][a * 3 + 0], gy = P.dsh[gp][a * 3 + 1], gz = P.dsh[gp][a * 3 + 2];
      // be[i] += B[gp][j][i] * sig[j] * wg, j ascending (src/assembly.cpp:135-136)
      acc[a * 3 + 0] += gx * sig[0] * wg;
      acc[a * 3 + 0] += gy * sig[3] * wg;
      acc[a * 3 + 0] += gz * sig[4] * wg;
      acc[a * 3 + 1] += gy * sig[1] * wg;
      acc[a * 3 + 1] += gx * sig[3] * wg;
      acc[a * 3 + 1] += gz * sig[5] * wg;
      acc[a * 3 + 2] += gz * sig[2] * wg;
      acc[a * 3 + 2] += gx * sig[4] * wg;
      acc[a * 3 + 2] += gy * sig[5] * wg;
    }
  }
#pragma unroll
  for (int q = 0; q < 24; ++q) be[(size_t)q * P.nelem_pad + e] = acc[q];
}

// mode 0: first residual of a Newton solve (sets norm0, its=0); 1: after an update (its++); 2: plain
__global__ void __launch_bounds__(NT)
    k_asm_rhs(const __grid_constant__ MeshConst P, const Lst L, SlotTables T, double *b_pool,
              size_t vstride, const double *__restrict__ bebuf, size_t bstride, int mode) {
  __shared__ double sm[NRED * (NT / 32)];
  __shared__ int sflag;
  const int slot = slot_of(L);
  if (slot < 0) return;
  mgpu_slot_state *st = &T.state[slot];
  if (mode == 1 && !st->nr_active) return;
  double *b = b_pool + (size_t)slot * vstride;
  const double *be = bebuf + (size_t)blockIdx.y * bstride;
  const int n = blockIdx.x * NT + threadIdx.x;
  double nrm[1] = {0.0};
  if (n < P.nn) {
    int i, j, k;
    node_ijk(P, n, i, j, k);
    double bx = 0.0, by = 0.0, bz = 0.0;
    if (!on_boundary(P, i, j, k)) {
      // the eight elements around the node in the reference's visiting order (ez outer, ex inner)
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const int az = (c >> 2) & 1, ay = (c >> 1) & 1, ax = c & 1;  // element = (i-1+ax, j-1+ay, k-1+az)
        const int e = ((k - 1 + az) * P.ney + (j - 1 + ay)) * P.nex + (i - 1 + ax);
        const int a = corner_of(1 - ax, 1 - ay, 1 - az);
        bx += be[(size_t)(a * 3 + 0) * P.nelem_pad + e];
        by += be[(size_t)(a * 3 + 1) * P.nelem_pad + e];
        bz += be[(size_t)(a * 3 + 2) * P.nelem_pad + e];
      }
      bx = -bx;
      by = -by;
      bz = -bz;
    }
    b[n] = bx;
    b[(size_t)P.nn_pad + n] = by;
    b[(size_t)2 * P.nn_pad + n] = bz;
    nrm[0] = bx * bx + by * by + bz * bz;
  }
  double *partial = T.partial + (size_t)slot * NRED * T.nblk_max;
  if (grid_sum<1>(nrm, partial, T.nblk_max, &st->ticket, sm, &sflag) && threadIdx.x == 0) {
    if (P.slab)
      T.red[slot * 8] = nrm[0];
    else
      tail_rhs(P, st, nrm[0], mode);
  }
}

// ------------------------------------------------------------------------------------------------
// assembly_mat, all-elastic RVE (src/assembly.cpp:106-178 + src/ell-common.cpp:166-297).
// Element matrices of elastic materials do not depend on u (src/material.cpp:84-94), so each node
// gathers its 27 3x3 blocks from a per-material 24x24 table held in shared memory.
// ------------------------------------------------------------------------------------------------
template <int DI, int DJ, int DK>
__device__ __forceinline__ void gather_block_elastic(const double *__restrict__ s_ke, const int (&et)[8],
                                                     double (&acc)[9]) {
#pragma unroll
  for (int q = 0; q < 9; ++q) acc[q] = 0.0;
  // reference element order of assembly_mat: ex outermost, ez innermost (src/assembly.cpp:112-114)
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    const int ax = (c >> 2) & 1, ay = (c >> 1) & 1, az = c & 1;  // element = (i-1+ax, j-1+ay, k-1+az)
    const int lx = 1 - ax, ly = 1 - ay, lz = 1 - az;             // this node inside that element
    const int mx = lx + DI, my = ly + DJ, mz = lz + DK;          // the neighbour inside that element
    if (mx >= 0 && mx <= 1 && my >= 0 && my <= 1 && mz >= 0 && mz <= 1) {
      const int a = corner_of(lx, ly, lz), jn = corner_of(mx, my, mz);
      const double *ke = s_ke + et[c] * 576;
#pragma unroll
      for (int fi = 0; fi < 3; ++fi)
#pragma unroll
        for (int fj = 0; fj < 3; ++fj) acc[fi * 3 + fj] += ke[(a * 3 + fi) * 24 + jn * 3 + fj];
    }
  }
}

// value of plane q goes to A[q * stride] (A already points at the node's first plane)
template <int NBR>
__device__ __forceinline__ void asm_elastic_slot(const double *__restrict__ s_ke, const int (&et)[8], double *A,
                                                 int stride) {
  constexpr int DI = NBR % 3 - 1, DJ = (NBR / 3) % 3 - 1, DK = NBR / 9 - 1;
  double acc[9];
  gather_block_elastic<DI, DJ, DK>(s_ke, et, acc);
#pragma unroll
  for (int q = 0; q < 9; ++q) A[(size_t)(NBR * 9 + q) * stride] = acc[q];
}

template <int NBR>
struct AsmElasticLoop {
  static __device__ __forceinline__ void run(const double *__restrict__ s_ke, const int (&et)[8], double *A,
                                             int stride) {
    asm_elastic_slot<NBR>(s_ke, et, A, stride);
    AsmElasticLoop<NBR + 1>::run(s_ke, et, A, stride);
  }
};
template <>
struct AsmElasticLoop<27> {
  static __device__ __forceinline__ void run(const double *__restrict__, const int (&)[8], double *, int) {}
};

__global__ void __launch_bounds__(NT)
    k_asm_mat_elastic(const __grid_constant__ MeshConst P, const Lst L, double *mat_pool,
                      size_t mstride, double *mat_shared, const int *__restrict__ elem_type,
                      const double *__restrict__ ke_tab) {
  __shared__ double s_ke[3 * 576];
  for (int q = threadIdx.x; q < 3 * 576; q += NT) s_ke[q] = ke_tab[q];
  __syncthreads();
  const int slot = slot_of(L);
  if (slot < 0) return;
  double *A = mat_shared ? mat_shared : mat_pool + (size_t)slot * mstride;
  const int m = blockIdx.x * NT + threadIdx.x;
  if (m >= P.nint) return;
  int i, j, k;
  interior_node(P, m, i, j, k);
  int et[8];
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    const int ex = i - 1 + ((c >> 2) & 1), ey = j - 1 + ((c >> 1) & 1), ez = k - 1 + (c & 1);
    et[c] = __ldg(&elem_type[(ez * P.ney + ey) * P.nex + ex]);
  }
  AsmElasticLoop<0>::run(s_ke, et, A + aidx(0, m), 32);
}

// Distinct row blocks of the implicit elastic operator: thread per row id; `codes[id]` = sum_c type_c * 3^c over the
// 8 elements around a node (c as in k_asm_mat_elastic), the row block is gathered by the SAME code as the explicit
// assembly, so the implicit operator is bit-identical to the assembled one.
__global__ void __launch_bounds__(NT)
    k_rows_build(const int *__restrict__ codes, int nrows, double *rows, double *rkinv,
                 const double *__restrict__ ke_tab) {
  __shared__ double s_ke[3 * 576];
  for (int q = threadIdx.x; q < 3 * 576; q += NT) s_ke[q] = ke_tab[q];
  __syncthreads();
  const int id = blockIdx.x * NT + threadIdx.x;
  if (id >= nrows) return;
  int code = codes[id];
  int et[8];
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    et[c] = code % 3;
    code /= 3;
  }
  double blk[NPLANE];
  AsmElasticLoop<0>::run(s_ke, et, blk, 1);
  double *out = rows + (size_t)id * RB_LEN;
  for (int nbr = 0; nbr < 27; ++nbr) {
    for (int q = 0; q < 9; ++q) out[nbr * RB_NBR + q] = blk[nbr * 9 + q];
    out[nbr * RB_NBR + 9] = 0.0;
  }
#pragma unroll
  for (int d = 0; d < 3; ++d) rkinv[id * 3 + d] = 1 / blk[13 * 9 + d * 4];
}

// ------------------------------------------------------------------------------------------------
// assembly_mat, general (damage / plastic / mixed), in two kernels.
//
// (1) k_elem_ctan: thread per element.  Gathers the 24 element dofs once and evaluates the reference's
//     forward-difference tangent (src/material.cpp:49-63: 7 stress evaluations) at the 8 Gauss points,
//     storing C_gp (36 doubles, row-major) into a scratch buffer laid out [(gp*36+q)][nelem_pad] per
//     slot of the current chunk, so stores and the later loads are coalesced over elements.  Every
//     tangent is computed exactly once (a node-gather that recomputes it would do so 8 times).
// (2) k_asm_mat_general: 8 threads per node: thread (node, c) owns element c of the 8 around the node
//     and computes the 3 rows of that element's 24x24 matrix that belong to the node,
//     K_rows = sum_gp (G_a^T C_gp wg) B_gp, reading C_gp from the scratch buffer.  The 8 row blocks
//     are added into a shared-memory image of the node's ELL row in the reference's element order
//     (ex outermost), and the image is written out coalesced: each ELL value is stored exactly once.
// Elastic elements of a mixed RVE take their rows from the precomputed per-material table.
// ------------------------------------------------------------------------------------------------
constexpr int GN = 16;      // nodes per block of the gather kernel (GN*8 == NT)
constexpr int CTAN_LEN = 8 * 36;  // doubles per element in the tangent scratch

__global__ void __launch_bounds__(NT)
    k_elem_ctan(const __grid_constant__ MeshConst P, const Lst L, SlotTables T,
                const double *__restrict__ u_pool, size_t vstride, const int *__restrict__ elem_type,
                double *__restrict__ cbuf, size_t cstride) {
  const int slot = slot_of(L);
  if (slot < 0) return;
  const double *u = u_pool + (size_t)slot * vstride;
  const double *vars = T.vars_old[slot];
  double *cb = cbuf + (size_t)blockIdx.y * cstride;
  const int e = blockIdx.x * NT + threadIdx.x;
  if (e >= P.nelem) return;
  const int type = __ldg(&elem_type[e]);
  const mpp_material m = P.mat[type];
  if (m.type == MPP_ELASTIC) return;  // rows come from the per-material table
  const int ez = e / (P.nex * P.ney);
  const int r = e - ez * P.nex * P.ney;
  const int ey = r / P.nex, ex = r - ey * P.nex;
  double ue[24];
  gather_ue(P, u, ex, ey, ez, ue);
  const int nv = mat_nvar(m.type);
#pragma unroll 1
  for (int gp = 0; gp < 8; ++gp) {
    double eps[6], C[36], vbuf[7];
    gp_strain(P.dsh[gp], ue, eps);
    const double *v = fetch_vars(vars, P.nelem_pad, e, gp, nv, vbuf);
    mat_ctan(m, eps, v, C);
#pragma unroll
    for (int q = 0; q < 36; ++q) cb[(size_t)(gp * 36 + q) * P.nelem_pad + e] = C[q];
  }
}

__global__ void __launch_bounds__(NT)
    k_asm_mat_general(const __grid_constant__ MeshConst P, const Lst L, double *mat_pool,
                      size_t mstride, double *mat_shared, const int *__restrict__ elem_type,
                      const double *__restrict__ ke_tab, const double *__restrict__ cbuf, size_t cstride) {
  extern __shared__ double s_acc[];  // [GN][243]
  const int slot = slot_of(L);
  if (slot < 0) return;
  const double *cb = cbuf + (size_t)blockIdx.y * cstride;
  double *A = mat_shared ? mat_shared : mat_pool + (size_t)slot * mstride;

  for (int q = threadIdx.x; q < GN * NPLANE; q += NT) s_acc[q] = 0.0;
  __syncthreads();

  // element-major thread order: the 16 lanes of a half-warp hold the SAME corner element of 16 x-consecutive nodes,
  // i.e. 16 consecutive elements => their tangent loads are 128-B segments (node-major order: 8 segments of 32 B)
  const int ln = threadIdx.x % GN, c = threadIdx.x / GN;
  const int m = blockIdx.x * GN + ln;  // interior-node index
  int i = 0, j = 0, k = 0;
  const bool work = m < P.nint;
  if (work) interior_node(P, m, i, j, k);

  double R[72];
  int a = 0;
  if (work) {
    const int ax = (c >> 2) & 1, ay = (c >> 1) & 1, az = c & 1;
    const int ex = i - 1 + ax, ey = j - 1 + ay, ez = k - 1 + az;
    a = corner_of(1 - ax, 1 - ay, 1 - az);
    const int e = (ez * P.ney + ey) * P.nex + ex;
    const int type = __ldg(&elem_type[e]);
    if (P.mat[type].type == MPP_ELASTIC) {
      const double *ke = ke_tab + type * 576 + a * 72;
#pragma unroll
      for (int q = 0; q < 72; ++q) R[q] = __ldg(&ke[q]);
    } else {
#pragma unroll
      for (int q = 0; q < 72; ++q) R[q] = 0.0;
      const double wg = P.wg;
#pragma unroll 1
      for (int gp = 0; gp < 8; ++gp) {
        double C[36];
        const double *cg = cb + (size_t)(gp * 36) * P.nelem_pad + e;
#pragma unroll
        for (int q = 0; q < 36; ++q) C[q] = cg[(size_t)q * P.nelem_pad];
        const double gx = P.dsh[gp][a * 3 + 0] * wg, gy = P.dsh[gp][a * 3 + 1] * wg, gz = P.dsh[gp][a * 3 + 2] * wg;
#pragma unroll
        for (int fi = 0; fi < 3; ++fi) {
          // row fi of G_a^T C : non-zero B rows for column 3a+fi (src/micro3D.cpp:100-119)
          double t[6];
#pragma unroll
          for (int q = 0; q < 6; ++q) {
            if (fi == 0)
              t[q] = gx * C[0 * 6 + q] + gy * C[3 * 6 + q] + gz * C[4 * 6 + q];
            else if (fi == 1)
              t[q] = gy * C[1 * 6 + q] + gx * C[3 * 6 + q] + gz * C[5 * 6 + q];
            else
              t[q] = gz * C[2 * 6 + q] + gx * C[4 * 6 + q] + gy * C[5 * 6 + q];
          }
#pragma unroll
          for (int jn = 0; jn < 8; ++jn) {
            const double hx = P.dsh[gp][jn * 3 + 0], hy = P.dsh[gp][jn * 3 + 1], hz = P.dsh[gp][jn * 3 + 2];
            R[fi * 24 + jn * 3 + 0] += t[0] * hx + t[3] * hy + t[4] * hz;
            R[fi * 24 + jn * 3 + 1] += t[1] * hy + t[3] * hx + t[5] * hz;
            R[fi * 24 + jn * 3 + 2] += t[2] * hz + t[4] * hx + t[5] * hy;
          }
        }
      }
    }
  }

  // add the eight row blocks in the reference's element order (c ascending == ex outermost)
  for (int round = 0; round < 8; ++round) {
    if (work && c == round) {
      double *row = s_acc + ln * NPLANE;
      const int cxa = corner_x(a), cya = corner_y(a), cza = corner_z(a);
#pragma unroll
      for (int jn = 0; jn < 8; ++jn) {
        const int pl = nbr_slot(corner_x(jn) - cxa, corner_y(jn) - cya, corner_z(jn) - cza) * 9;
#pragma unroll
        for (int fi = 0; fi < 3; ++fi)
#pragma unroll
          for (int fj = 0; fj < 3; ++fj) row[pl + fi * 3 + fj] += R[fi * 24 + jn * 3 + fj];
      }
    }
    __syncthreads();
  }
  for (int q = threadIdx.x; q < GN * NPLANE; q += NT) {
    const int pl = q / GN, l = q % GN;
    const int md = blockIdx.x * GN + l;
    if (md < P.nint) A[aidx(pl, md)] = s_acc[l * NPLANE + pl];
  }
}

// ------------------------------------------------------------------------------------------------
// DPCG (src/ell.cpp:66-122)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(NT)
    k_cg_init(const __grid_constant__ MeshConst P, const Lst L, SlotTables T, VecPool V,
              int use_shared) {
  __shared__ double sm[NRED * (NT / 32)];
  __shared__ int sflag;
  const int slot = slot_of(L);
  if (slot < 0) return;
  mgpu_slot_state *st = &T.state[slot];
  // use_shared: 0 per-slot matrix, 1 the shared linear matrix (A0), 2 the generic host matrix (reference layout)
  const double *A = use_shared == OP_SHARED ? V.mat_shared : V.mat + (size_t)slot * V.mstride;
  const size_t vo = (size_t)slot * V.vstride;
  const int n = blockIdx.x * NT + threadIdx.x;
  double red[2] = {0.0, 0.0};
  if (n < P.nn) {
    int i, j, k;
    node_ijk(P, n, i, j, k);
    const bool bnd = on_boundary(P, i, j, k);
    const int m = bnd ? 0 : interior_index(P, i, j, k);
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      const size_t ix = vo + (size_t)d * P.nn_pad + n;
      double diag = 1.0;  // boundary rows are identity rows (ell_set_bc_3D)
      if (use_shared == OP_GENERIC)
        diag = V.gen[((size_t)n * 3 + d) * 81 + 13 * 3 + d];
      else if (use_shared != OP_IMPLICIT && !bnd)
        diag = A[aidx(13 * 9 + d * 4, m)];
      // src/ell.cpp:73-76 (the implicit operator keeps 1/diag per distinct row block)
      const double kk = (use_shared == OP_IMPLICIT && !bnd) ? __ldg(&V.rkinv[__ldg(&V.rowid[m]) * 3 + d]) : 1 / diag;
      const double r = V.b[ix];    // r = b - A*0 (src/ell.cpp:78-82)
      const double z = kk * r;
      if (use_shared != OP_IMPLICIT) {  // the implicit operator re-derives k (and z = k r) from its row table
        V.k[ix] = kk;
        V.z[ix] = z;
      }
      V.du[ix] = 0.0;
      V.r[ix] = r;
      V.p[ix] = z;
      V.Ap[ix] = 0.0;  // boundary entries of Ap are never written by the SpMV (p is 0 there)
      red[0] += r * z;
      red[1] += z * z;
    }
  }
  double *partial = T.partial + (size_t)slot * NRED * T.nblk_max;
  if (grid_sum<2>(red, partial, T.nblk_max, &st->ticket, sm, &sflag) && threadIdx.x == 0) {
    if (P.slab) {
      T.red[slot * 8] = red[0];
      T.red[slot * 8 + 1] = red[1];
    } else {
      tail_cg_init(P, st, red[0], red[1]);
    }
  }
}

// Ap = A p fused with p.Ap.  Thread per INTERIOR node: 3 rows, 243 coalesced value loads (one 256-B warp load per
// plane of the node's tile, immediate offsets), 81 p loads that hit L1 (each p value is reused by 27 nodes x 3
// rows).  Boundary rows are identity rows with p = 0 (b and x0 vanish there), so they contribute nothing.
template <int NBR>
__device__ __forceinline__ void spmv_slot(const double *__restrict__ a, const double *__restrict__ p, size_t npad,
                                          int n, int nx, int nxny, double &y0, double &y1, double &y2) {
  constexpr int DI = NBR % 3 - 1, DJ = (NBR / 3) % 3 - 1, DK = NBR / 9 - 1;
  const int q = n + DI + DJ * nx + DK * nxny;
  const double px = p[q], py = p[npad + q], pz = p[2 * npad + q];
  constexpr int o = NBR * 9 * 32;  // planes of one tile are 32 doubles apart
  y0 += a[o + 0] * px;
  y0 += a[o + 32] * py;
  y0 += a[o + 64] * pz;
  y1 += a[o + 96] * px;
  y1 += a[o + 128] * py;
  y1 += a[o + 160] * pz;
  y2 += a[o + 192] * px;
  y2 += a[o + 224] * py;
  y2 += a[o + 256] * pz;
}
template <int NBR>
struct SpmvLoop {
  static __device__ __forceinline__ void run(const double *__restrict__ a, const double *__restrict__ p,
                                             size_t npad, int n, int nx, int nxny, double &y0, double &y1,
                                             double &y2) {
    spmv_slot<NBR>(a, p, npad, n, nx, nxny, y0, y1, y2);
    SpmvLoop<NBR + 1>::run(a, p, npad, n, nx, nxny, y0, y1, y2);
  }
};
template <>
struct SpmvLoop<27> {
  static __device__ __forceinline__ void run(const double *__restrict__, const double *__restrict__, size_t, int,
                                             int, int, double &, double &, double &) {}
};

__global__ void __launch_bounds__(NT)
    k_spmv_dot(const __grid_constant__ MeshConst P, const Lst L, SlotTables T, VecPool V, int use_shared, int force) {
  __shared__ double sm[NRED * (NT / 32)];
  __shared__ int sflag;
  const int slot = slot_of(L);
  if (slot < 0) return;
  mgpu_slot_state *st = &T.state[slot];
  if (!force && !st->cg_active) return;
  const double *A = use_shared ? V.mat_shared : V.mat + (size_t)slot * V.mstride;
  const size_t vo = (size_t)slot * V.vstride;
  const double *p = V.p + vo;
  double *Ap = V.Ap + vo;
  const int m = blockIdx.x * NT + threadIdx.x;
  double red[1] = {0.0};
  if (m < P.nint) {
    int i, j, k;
    const int n = interior_node(P, m, i, j, k);
    double y0 = 0.0, y1 = 0.0, y2 = 0.0;
    const size_t npad = P.nn_pad;
    SpmvLoop<0>::run(A + aidx(0, m), p, npad, n, P.nx, P.nxny, y0, y1, y2);
    Ap[n] = y0;
    Ap[npad + n] = y1;
    Ap[2 * npad + n] = y2;
    red[0] = p[n] * y0 + p[npad + n] * y1 + p[2 * npad + n] * y2;
  }
  double *partial = T.partial + (size_t)slot * NRED * T.nblk_max;
  if (grid_sum<1>(red, partial, T.nblk_max, &st->ticket, sm, &sflag) && threadIdx.x == 0) {
    if (P.slab)
      T.red[slot * 8] = red[0];
    else
      tail_spmv(st, red[0]);
  }
}

// ------------------------------------------------------------------------------------------------
// Implicit operator of an all-elastic RVE (every material elastic => the Jacobian does not depend on u,
// src/material.cpp:84-94, and is the same for every macro Gauss point and every Newton step).  No matrix is
// stored per RVE: a thread owns one interior node, fetches the node's row block from the small table of distinct
// row blocks (L1-resident: almost every node of a warp uses the same block, so the 243 loads are broadcasts) and
// applies it to MR right-hand sides (slots) at once.  Per RVE the kernel moves 48 B/node (p in, Ap out) instead
// of 1992 B/node; the FMA order per slot is exactly that of k_spmv_dot, so results are bit-identical to the
// assembled path.
// ------------------------------------------------------------------------------------------------
constexpr int MR = 8;  // right-hand sides per thread
constexpr int UPD_VPT = 2;  // nodes per thread of k_cg_update / k_cg_update_imp

// blockIdx.y = group of R consecutive entries of the slot list (n_list entries; inside a graph the device-side count)
template <int R>
__global__ void __launch_bounds__(NT, 3)
    k_spmv_dot_imp(const __grid_constant__ MeshConst P, const Lst L, int n_list, SlotTables T, VecPool V, int force) {
  __shared__ double sm[R * (NT / 32)];
  __shared__ int s_slot[R];
  __shared__ int s_last[R];
  if (threadIdx.x < R) {
    const int yy = (int)blockIdx.y * R + (int)threadIdx.x + L.yoff;
    const int cnt = L.dcount ? min(*L.dcount, n_list + L.yoff) : n_list + L.yoff;
    int slot = yy < cnt ? L.list[yy] : -1;
    if (slot >= 0 && !force && !T.state[slot].cg_active) slot = -1;
    s_slot[threadIdx.x] = slot;
  }
  __syncthreads();
  int any = -1;
#pragma unroll
  for (int r = R - 1; r >= 0; --r)
    if (s_slot[r] >= 0) any = s_slot[r];
  if (any < 0) return;
  unsigned off[R];
#pragma unroll
  for (int r = 0; r < R; ++r) off[r] = (unsigned)((size_t)(s_slot[r] >= 0 ? s_slot[r] : any) * V.vstride);

  const int m = blockIdx.x * NT + threadIdx.x;
  double red[R];
#pragma unroll
  for (int r = 0; r < R; ++r) red[r] = 0.0;
  if (m < P.nint) {
    int i, j, k;
    const int n = interior_node(P, m, i, j, k);
    const size_t npad = P.nn_pad;
    double y[R][3];
#pragma unroll
    for (int r = 0; r < R; ++r) y[r][0] = y[r][1] = y[r][2] = 0.0;
    const double *a = V.rows + (size_t)__ldg(&V.rowid[m]) * RB_LEN;
    // rolled over the 9 (dz, dy) neighbour rows, unrolled over dx: bounds the loads the scheduler can hoist
#pragma unroll 1
    for (int row = 0; row < 9; ++row) {
      const int dk = row / 3 - 1, dj = row - (dk + 1) * 3 - 1;
      const int q0 = n + dj * P.nx + dk * P.nxny;
      const double *ar = a + row * 3 * RB_NBR;
#pragma unroll
      for (int di = -1; di <= 1; ++di) {
        double av[10];
#pragma unroll
        for (int t = 0; t < 5; ++t) {
          const double2 v = __ldg(reinterpret_cast<const double2 *>(ar + (di + 1) * RB_NBR) + t);
          av[2 * t] = v.x;
          av[2 * t + 1] = v.y;
        }
        // component-outer order: 3R independent DFMAs between two updates of the same accumulator
#pragma unroll
        for (int fj = 0; fj < 3; ++fj) {
#pragma unroll
          for (int r = 0; r < R; ++r) {
            const double pval = V.p[(size_t)off[r] + (size_t)fj * npad + (q0 + di)];
            y[r][0] += av[fj] * pval;
            y[r][1] += av[3 + fj] * pval;
            y[r][2] += av[6 + fj] * pval;
          }
        }
      }
    }
#pragma unroll
    for (int r = 0; r < R; ++r) {
      if (s_slot[r] >= 0) {
        const double *pp = V.p + ((size_t)off[r] + n);
        double *Ap = V.Ap + ((size_t)off[r] + n);
        Ap[0] = y[r][0];
        Ap[npad] = y[r][1];
        Ap[2 * npad] = y[r][2];
        red[r] = pp[0] * y[r][0] + pp[npad] * y[r][1] + pp[2 * npad] * y[r][2];
      }
    }
  }
  // per-slot deterministic ticket reductions (the same partial layout and summation order as grid_sum<1>)
  block_sum<R>(red, sm);
  if (threadIdx.x == 0) {
#pragma unroll
    for (int r = 0; r < R; ++r)
      if (s_slot[r] >= 0) T.partial[(size_t)s_slot[r] * NRED * T.nblk_max + blockIdx.x] = red[r];
    __threadfence();
#pragma unroll
    for (int r = 0; r < R; ++r) {
      int last = 0;
      if (s_slot[r] >= 0) last = atomicAdd(&T.state[s_slot[r]].ticket, 1u) == gridDim.x - 1;
      s_last[r] = last;
    }
  }
  __syncthreads();
#pragma unroll 1
  for (int r = 0; r < R; ++r) {
    if (!s_last[r]) continue;
    __threadfence();
    const int slot = s_slot[r];
    const double *partial = T.partial + (size_t)slot * NRED * T.nblk_max;
    double acc[1] = {0.0};
    for (int b = threadIdx.x; b < (int)gridDim.x; b += NT) acc[0] += __ldcg(&partial[b]);
    block_sum<1>(acc, sm);
    if (threadIdx.x == 0) {
      T.state[slot].ticket = 0u;
      if (P.slab)
        T.red[slot * 8] = acc[0];
      else
        tail_spmv(&T.state[slot], acc[0]);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Tiled version of the implicit SpMV (the default): k_spmv_dot_imp is bound by L1 wavefronts (81 p loads per
// node, ncu: l1tex 72 %, FP64 pipe 25 %), so this kernel stages p through shared memory and blocks 8 x-adjacent
// nodes per thread: a thread needs 10 x 9 x 3 p values for its 8 nodes (34 per node instead of 81), fetched as
// conflict-free 128-bit shared loads, which leaves the FP64 pipe (243 DFMA per node) as the bound.
//   block  = cb warps; warp w owns x-chunk w of the tile (8 nodes), lane l owns the (y, z) row (l & 7, l >> 3):
//            a tile is 8cb x 8 x 4 interior nodes, its p brick (8cb+2) x 10 x 6 x 3 doubles;
//   smem   : brick[d][bz][by][pitch = 8cb+2]; lanes of a quarter-warp differ in by => their 16-B accesses are
//            pitch*8 B apart, and (8cb+2)/2 is odd, so they cover all 32 banks;
//   rows   : almost every thread's 8 nodes share ONE row block (chunk_id >= 0): 243 L1-broadcast loads per 8 nodes;
//            chunks that straddle a material interface (chunk_id < 0) fetch the row block of each node.
// The FMA order per node is that of k_spmv_dot, so Ap is bit-identical; p.Ap is summed in a different (fixed) order.
// ------------------------------------------------------------------------------------------------
struct TileInfo {
  int cb, tiles_x, tiles_y, tiles_z, nchunk, pitch;
  const int *chunk_id;  // [niz][niy][nchunk]: row-block id shared by the chunk's nodes, or -1
  // TMA kernel: every chunk is computed with ONE pure-material row block (id 0..2, shared memory) and the nodes of
  // the chunk whose own block differs (material interfaces) are recomputed from the per-tile fix-up list
  const int *chunk_pure;  // [niz][niy][nchunk]: pure id | (8-bit mask of the nodes to fix up) << 8
  const int *fix_ptr;     // [ntiles + 1]
  const int2 *fix;        // x: lx | ry << 8 | rz << 12 (position inside the tile), y: row-block id
};
constexpr int TILE_Y = 8, TILE_Z = 4, BRICK_ROWS = (TILE_Y + 2) * (TILE_Z + 2);

__device__ __forceinline__ void cp_async8(double *smem_dst, const double *gsrc) {
  const unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(sa), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;\n" ::: "memory");
}

// UNIFORM: every node of the thread uses the row block `a_uni` (shared-memory copy of a pure-material block);
// otherwise node t uses rows[rid[t]] (global, L1).
template <bool UNIFORM>
__device__ __forceinline__ void tile_rows_apply(const double *__restrict__ a_uni, const double *__restrict__ rows,
                                                const int (&rid)[8], const double *__restrict__ brick, int pitch,
                                                int bx0, int ry, int rz, double (&acc)[8][3]) {
#pragma unroll(UNIFORM ? 3 : 1)
  for (int row = 0; row < 9; ++row) {
    const int dk = row / 3, dj = row - dk * 3;  // 0..2 (offset + 1)
    const int rbase = ((rz + dk) * (TILE_Y + 2) + (ry + dj)) * pitch + bx0;
    double pv[3][10];
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      const double2 *s2 = reinterpret_cast<const double2 *>(brick + d * (BRICK_ROWS * pitch) + rbase);
#pragma unroll
      for (int h = 0; h < 5; ++h) {
        const double2 v = s2[h];
        pv[d][2 * h] = v.x;
        pv[d][2 * h + 1] = v.y;
      }
    }
#pragma unroll
    for (int di = 0; di < 3; ++di) {
      const int ao = (row * 3 + di) * RB_NBR;
      if (UNIFORM) {
        double av[10];
#pragma unroll
        for (int q = 0; q < 5; ++q) {
          const double2 v = reinterpret_cast<const double2 *>(a_uni + ao)[q];
          av[2 * q] = v.x;
          av[2 * q + 1] = v.y;
        }
        // component-outer order: 24 independent DFMAs between two updates of the same accumulator (each accumulator
        // still receives its px, py, pz terms in this order)
#pragma unroll
        for (int fj = 0; fj < 3; ++fj) {
#pragma unroll
          for (int t = 0; t < 8; ++t) {
            const double pval = pv[fj][t + di];
            acc[t][0] += av[fj] * pval;
            acc[t][1] += av[3 + fj] * pval;
            acc[t][2] += av[6 + fj] * pval;
          }
        }
      } else {
#pragma unroll
        for (int t = 0; t < 8; ++t) {
          const double2 *a2 = reinterpret_cast<const double2 *>(rows + (size_t)rid[t] * RB_LEN + ao);
          double av[10];
#pragma unroll
          for (int q = 0; q < 5; ++q) {
            const double2 v = __ldg(a2 + q);
            av[2 * q] = v.x;
            av[2 * q + 1] = v.y;
          }
          const double px = pv[0][t + di], py = pv[1][t + di], pz = pv[2][t + di];
          acc[t][0] += av[0] * px;
          acc[t][0] += av[1] * py;
          acc[t][0] += av[2] * pz;
          acc[t][1] += av[3] * px;
          acc[t][1] += av[4] * py;
          acc[t][1] += av[5] * pz;
          acc[t][2] += av[6] * px;
          acc[t][2] += av[7] * py;
          acc[t][2] += av[8] * pz;
        }
      }
    }
  }
}

// every node of the thread uses the row block `a_uni` (shared memory)
__device__ __forceinline__ void tile_rows_apply_uniform(const double *__restrict__ a_uni,
                                                        const double *__restrict__ brick, int pitch, int bx0, int ry,
                                                        int rz, double (&acc)[8][3]) {
  const int rid[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  tile_rows_apply<true>(a_uni, nullptr, rid, brick, pitch, bx0, ry, rz, acc);
}

// about 384 threads x 168 registers per SM: 3 blocks of 4 warps, 2 of 6, ...
template <int CB>
__global__ void __launch_bounds__(32 * CB, (384 / (32 * CB)) > 0 ? 384 / (32 * CB) : 1)
    k_spmv_dot_tile(const __grid_constant__ MeshConst P, const Lst L, SlotTables T, VecPool V, TileInfo ti,
                    int force) {
  extern __shared__ __align__(16) double s_brick[];  // [3][BRICK_ROWS][pitch]
  __shared__ __align__(16) double s_rows[3 * RB_LEN];  // row blocks 0..2 = nodes surrounded by one material
  __shared__ double s_red[8];
  __shared__ int sflag;
  const int slot = slot_of(L);
  if (slot < 0) return;
  mgpu_slot_state *st = &T.state[slot];
  if (!force && !st->cg_active) return;
  for (int q = threadIdx.x; q < 3 * RB_LEN; q += 32 * CB) s_rows[q] = __ldg(&V.rows[q]);
  const size_t vo = (size_t)slot * V.vstride;
  const double *p = V.p + vo;
  double *Ap = V.Ap + vo;
  const size_t npad = P.nn_pad;
  constexpr int pitch = 8 * CB + 2, nw = CB;
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int b = blockIdx.x;
  const int tx = b % ti.tiles_x;
  b /= ti.tiles_x;
  const int ty = b % ti.tiles_y, tz = b / ti.tiles_y;
  const int X0 = tx * 8 * CB, Y0 = ty * TILE_Y, Z0 = tz * TILE_Z;  // grid coordinates of the brick origin

  // ---- p brick -> shared memory (rows of the brick are contiguous in global memory) ----
  for (int r = w; r < 3 * BRICK_ROWS; r += nw) {
    const int d = r / BRICK_ROWS, rr = r - d * BRICK_ROWS, bz = rr / (TILE_Y + 2), by = rr - bz * (TILE_Y + 2);
    const int gy = Y0 + by, gz = Z0 + bz;
    const bool row_in = gy < P.ny && gz < P.nz;
    const double *src = p + (size_t)d * npad + (size_t)gz * P.nxny + gy * P.nx + X0;
    double *dst = s_brick + r * pitch;
    for (int bx = lane; bx < pitch; bx += 32) {
      if (row_in && X0 + bx < P.nx)
        cp_async8(dst + bx, src + bx);
      else
        dst[bx] = 0.0;
    }
  }
  cp_async_wait_all();
  __syncthreads();

  // ---- 8 nodes per thread ----
  const int ry = lane & 7, rz = lane >> 3;
  const int c = tx * CB + w, jj = Y0 + ry, kk = Z0 + rz;  // chunk, interior y, interior z (0-based)
  double red = 0.0;
  const bool work = c < ti.nchunk && jj < P.niy && kk < P.niz;
  int cid = 0;
  if (work) cid = __ldg(&ti.chunk_id[(kk * P.niy + jj) * ti.nchunk + c]);
  // one path per warp: the per-node row path only when some lane's chunk straddles a material interface
  const bool uniform = __all_sync(0xffffffffu, !work || (cid >= 0 && cid < 3));
  if (work) {
    const int ii0 = c * 8;
    const int nvalid = min(8, P.nix - ii0);
    const int m0 = (kk * P.niy + jj) * P.nix + ii0;
    double acc[8][3];
#pragma unroll
    for (int t = 0; t < 8; ++t) acc[t][0] = acc[t][1] = acc[t][2] = 0.0;
    int rid[8];
    if (uniform) {
      rid[0] = cid;
      tile_rows_apply<true>(s_rows + cid * RB_LEN, V.rows, rid, s_brick, pitch, 8 * w, ry, rz, acc);
    } else {
#pragma unroll
      for (int t = 0; t < 8; ++t) rid[t] = cid >= 0 ? cid : __ldg(&V.rowid[m0 + min(t, nvalid - 1)]);
      tile_rows_apply<false>(s_rows, V.rows, rid, s_brick, pitch, 8 * w, ry, rz, acc);
    }
    const int n0 = (kk + 1) * P.nxny + (jj + 1) * P.nx + ii0 + 1;
    const int cbase = ((rz + 1) * (TILE_Y + 2) + (ry + 1)) * pitch + 8 * w + 1;
#pragma unroll
    for (int t = 0; t < 8; ++t) {
      if (t < nvalid) {
        Ap[n0 + t] = acc[t][0];
        Ap[npad + n0 + t] = acc[t][1];
        Ap[2 * npad + n0 + t] = acc[t][2];
        red += s_brick[cbase + t] * acc[t][0] + s_brick[BRICK_ROWS * pitch + cbase + t] * acc[t][1] +
               s_brick[2 * BRICK_ROWS * pitch + cbase + t] * acc[t][2];
      }
    }
  }

  // ---- p.Ap: deterministic ticket reduction over the tiles of this slot ----
  red = warp_sum(red);
  if (lane == 0) s_red[w] = red;
  __syncthreads();
  double *partial = T.partial + (size_t)slot * NRED * T.nblk_max;
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int ww = 0; ww < nw; ++ww) s += s_red[ww];
    partial[blockIdx.x] = s;
    __threadfence();
    sflag = atomicAdd(&st->ticket, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (!sflag) return;
  __threadfence();
  double acc2 = 0.0;
  for (int q = threadIdx.x; q < (int)gridDim.x; q += 32 * CB) acc2 += __ldcg(&partial[q]);
  acc2 = warp_sum(acc2);
  __syncthreads();
  if (lane == 0) s_red[w] = acc2;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int ww = 0; ww < nw; ++ww) s += s_red[ww];
    st->ticket = 0u;
    if (P.slab)
      T.red[slot * 8] = s;
    else
      tail_spmv(st, s);
  }
}

// ------------------------------------------------------------------------------------------------
// TMA version of the tiled implicit SpMV (default whenever nx is even: TMA needs 16-B global strides).  The p brick of
// a (tile, slot) item -- (8cb+2) x 10 x 6 nodes x 3 components -- is ONE cp.async.bulk.tensor.5d from the pool
// viewed as a rank-5 tensor (x, y, z, component, slot); out-of-grid parts are zero-filled by the TMA unit.  A block
// walks over its items (ntl consecutive tiles x rs slots) with a two-stage mbarrier pipeline: thread 0 arms the
// barrier and issues the load of item i+1 before the block computes item i, so the brick traffic (L2 -> smem)
// overlaps the FP64 work and costs no issue slots of the compute warps.  Compute, row tables and the per-slot
// deterministic reduction are those of k_spmv_dot_tile.
// ------------------------------------------------------------------------------------------------
constexpr int FOLD_PLANE = 2;  // planes 0/1 of the partial-sum buffer belong to the consumer's own grid_sum<2>

// sum of the n per-(tile, warp) partials of a slot, by ONE warp, in a fixed order; result in every lane
__device__ __forceinline__ double fold_partials(const double *partial, int n) {
  double acc = 0.0;
  for (int q = threadIdx.x & 31; q < n; q += 32) acc += __ldcg(&partial[q]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  return acc;
}

// z.z and r.z of k_cg_update / k_cg_update_imp: one warp per slot folds the per-block partials (planes 0 and 1) in a
// fixed order, then the scalar tail of the iteration and the loop-head test of the next one (src/ell.cpp:108-119)
__global__ void k_fold_update(const __grid_constant__ MeshConst P, const Lst L, SlotTables T, int nblk) {
  const int slot = slot_of(L);
  if (slot < 0) return;
  mgpu_slot_state *st = &T.state[slot];
  if (!st->cg_active) return;
  const double *partial = T.partial + (size_t)slot * NRED * T.nblk_max;
  const double zz = fold_partials(partial, nblk);
  const double rz = fold_partials(partial + T.nblk_max, nblk);
  if (threadIdx.x == 0) {
    if (P.slab) {
      T.red[slot * 8] = zz;
      T.red[slot * 8 + 1] = rz;
    } else {
      tail_cg_update(P, st, zz, rz);
    }
  }
}

// slab mode / forced applications: fold p.Ap right after the SpMV (one warp per slot)
__global__ void k_fold_spmv(const __grid_constant__ MeshConst P, const Lst L, SlotTables T, int nfold, int force) {
  const int slot = slot_of(L);
  if (slot < 0) return;
  mgpu_slot_state *st = &T.state[slot];
  if (!force && !st->cg_active) return;
  const double s = fold_partials(T.partial + ((size_t)slot * NRED + FOLD_PLANE) * T.nblk_max, nfold);
  if (threadIdx.x == 0) {
    if (P.slab)
      T.red[slot * 8] = s;
    else
      tail_spmv(st, s);
  }
}

constexpr int TMA_MAX_RS = 8;

__device__ __forceinline__ void mbar_init(uint64_t *bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(
                   (unsigned)__cvta_generic_to_shared(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, unsigned parity) {
  const unsigned a = (unsigned)__cvta_generic_to_shared(bar);
  unsigned ok;
  do {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(a), "r"(parity)
        : "memory");
  } while (!ok);
}
__device__ __forceinline__ void tma_load_5d(void *smem_dst, const CUtensorMap *tmap, uint64_t *bar, int c0, int c1,
                                            int c2, int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], "
      "[%2];\n" ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)),
      "l"(tmap), "r"((unsigned)__cvta_generic_to_shared(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}

template <int CB>
__global__ void __launch_bounds__(32 * CB, CB <= 4 ? 2 : 1)
    k_spmv_dot_tma(const __grid_constant__ MeshConst P, const Lst L, int n_list, SlotTables T, VecPool V, TileInfo ti,
                   const __grid_constant__ CUtensorMap tmap, int ntl, int rs, int force) {
  extern __shared__ unsigned char s_raw[];
  constexpr int pitch = 8 * CB + 2;
  constexpr int BRICK = 3 * BRICK_ROWS * pitch;                       // doubles
  constexpr int STAGE_BYTES = (BRICK * 8 + 127) / 128 * 128;
  __shared__ uint64_t s_full[2];
  __shared__ int s_slot[TMA_MAX_RS];
  __shared__ __align__(16) double s_rows[3 * RB_LEN];  // row blocks 0..2 = nodes surrounded by one material
  __shared__ int s_done[2];
  // 128-B aligned stages; plain array arithmetic keeps the pointer in the shared address space (LDS, not generic LD)
  unsigned char *s_base = s_raw + ((128u - ((unsigned)__cvta_generic_to_shared(s_raw) & 127u)) & 127u);
  for (int q = threadIdx.x; q < 3 * RB_LEN; q += 32 * CB) s_rows[q] = __ldg(&V.rows[q]);
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ntiles = ti.tiles_x * ti.tiles_y * ti.tiles_z;
  const int tile0 = blockIdx.x * ntl;
  const int nitems = min(ntl, ntiles - tile0) * rs;

  if ((int)threadIdx.x < rs) {
    const int yy = (int)blockIdx.y * rs + (int)threadIdx.x + L.yoff;
    const int cnt = L.dcount ? min(*L.dcount, n_list + L.yoff) : n_list + L.yoff;
    int slot = yy < cnt ? L.list[yy] : -1;
    if (slot >= 0 && !force && !T.state[slot].cg_active) slot = -1;
    s_slot[threadIdx.x] = slot;
  }
  if (threadIdx.x == 0) {
    s_done[0] = s_done[1] = 0;
    mbar_init(&s_full[0], 1);
    mbar_init(&s_full[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
  }
  __syncthreads();

  auto next_active = [&](int i) {
    while (i < nitems && s_slot[i % rs] < 0) ++i;
    return i;
  };
  auto issue = [&](int item, int stage) {  // thread 0
    const int tile = tile0 + item / rs, slot = s_slot[item % rs];
    const int tx = tile % ti.tiles_x, tyz = tile / ti.tiles_x, ty = tyz % ti.tiles_y, tz = tyz / ti.tiles_y;
    mbar_expect_tx(&s_full[stage], BRICK * 8);
    tma_load_5d(s_base + stage * STAGE_BYTES, &tmap, &s_full[stage], tx * 8 * CB, ty * TILE_Y, tz * TILE_Z, 0, slot);
  };

  const int ry = lane & 7, rz = lane >> 3;
  const size_t npad = P.nn_pad;
  // Warps run through the items without block-wide barriers: a warp waits for the brick of its item (full barrier),
  // computes, deposits its partial p.Ap, and signs the stage off; the LAST warp to sign off re-arms the stage with
  // the load of the item after next.  Items 0 and 1 are issued up front.
  int cur = next_active(0), k = 0;
  if (threadIdx.x == 0 && cur < nitems) {
    issue(cur, 0);
    const int second = next_active(cur + 1);
    if (second < nitems) issue(second, 1);
  }
  // per-tile thread state (independent of the slot: with ntl == 1 it is computed once per block)
  int cur_tile = -1, pure = 0, keep = 0, n0 = 0, f0 = 0, f1 = 0, nfix0 = 0;
  bool work = false;
  while (cur < nitems) {
    const int nxt = next_active(cur + 1);
    const int stage = k & 1;
    const int tile = tile0 + cur / rs, slot = s_slot[cur % rs];
    if (tile != cur_tile) {
      cur_tile = tile;
      const int tx = tile % ti.tiles_x, tyz = tile / ti.tiles_x, ty = tyz % ti.tiles_y, tz = tyz / ti.tiles_y;
      const int c = tx * CB + w, jj = ty * TILE_Y + ry, kk = tz * TILE_Z + rz;
      work = c < ti.nchunk && jj < P.niy && kk < P.niz;
      const int info = work ? __ldg(&ti.chunk_pure[(kk * P.niy + jj) * ti.nchunk + c]) : 0;
      pure = info & 0xff;
      const int ii0 = c * 8;
      const int nvalid = min(8, P.nix - ii0);
      keep = work ? (((1 << nvalid) - 1) & ~(info >> 8)) : 0;  // nodes this thread stores itself
      n0 = (kk + 1) * P.nxny + (jj + 1) * P.nx + ii0 + 1;
      f0 = __ldg(&ti.fix_ptr[tile]);
      f1 = __ldg(&ti.fix_ptr[tile + 1]);
      nfix0 = (ty * TILE_Y + 1) * P.nx + (tz * TILE_Z + 1) * P.nxny + tx * 8 * CB + 1;  // node of tile position 0
    }
    const double *s_brick = reinterpret_cast<const double *>(s_base + stage * STAGE_BYTES);
    mbar_wait(&s_full[stage], (k >> 1) & 1);

    double red = 0.0;
    double *Ap = V.Ap + (size_t)slot * V.vstride;
    if (work) {
      double acc[8][3];
#pragma unroll
      for (int t = 0; t < 8; ++t) acc[t][0] = acc[t][1] = acc[t][2] = 0.0;
      tile_rows_apply_uniform(s_rows + pure * RB_LEN, s_brick, pitch, 8 * w, ry, rz, acc);
      const int cbase = ((rz + 1) * (TILE_Y + 2) + (ry + 1)) * pitch + 8 * w + 1;
#pragma unroll
      for (int t = 0; t < 8; ++t) {
        if ((keep >> t) & 1) {
          Ap[n0 + t] = acc[t][0];
          Ap[npad + n0 + t] = acc[t][1];
          Ap[2 * npad + n0 + t] = acc[t][2];
          red += s_brick[cbase + t] * acc[t][0] + s_brick[BRICK_ROWS * pitch + cbase + t] * acc[t][1] +
                 s_brick[2 * BRICK_ROWS * pitch + cbase + t] * acc[t][2];
        }
      }
    }
    // fix-up: the nodes of this tile that sit on a material interface, one per thread, shared evenly by the warps
    // (the brick holds all the p values they need; their row blocks come from the L1-resident table)
    for (int f = f0 + w * 32 + lane; f < f1; f += 32 * CB) {
      const int2 e = __ldg(&ti.fix[f]);
      const int lx = e.x & 0xff, fy = (e.x >> 8) & 0xf, fz = (e.x >> 12) & 0xf;
      const double2 *a2 = reinterpret_cast<const double2 *>(V.rows + (size_t)e.y * RB_LEN);
      double y0 = 0.0, y1 = 0.0, y2 = 0.0;
#pragma unroll 3
      for (int row = 0; row < 9; ++row) {
        const int dk = row / 3, dj = row - dk * 3;
        const int rb = ((fz + dk) * (TILE_Y + 2) + (fy + dj)) * pitch + lx;
#pragma unroll
        for (int di = 0; di < 3; ++di) {
          double av[10];
#pragma unroll
          for (int q = 0; q < 5; ++q) {
            const double2 v = __ldg(a2 + (row * 3 + di) * (RB_NBR / 2) + q);
            av[2 * q] = v.x;
            av[2 * q + 1] = v.y;
          }
          const double px = s_brick[rb + di], py = s_brick[BRICK_ROWS * pitch + rb + di],
                       pz = s_brick[2 * BRICK_ROWS * pitch + rb + di];
          y0 += av[0] * px;
          y0 += av[1] * py;
          y0 += av[2] * pz;
          y1 += av[3] * px;
          y1 += av[4] * py;
          y1 += av[5] * pz;
          y2 += av[6] * px;
          y2 += av[7] * py;
          y2 += av[8] * pz;
        }
      }
      const int n = nfix0 + fz * P.nxny + fy * P.nx + lx;
      const int cb0 = ((fz + 1) * (TILE_Y + 2) + (fy + 1)) * pitch + lx + 1;
      Ap[n] = y0;
      Ap[npad + n] = y1;
      Ap[2 * npad + n] = y2;
      red += s_brick[cb0] * y0 + s_brick[BRICK_ROWS * pitch + cb0] * y1 + s_brick[2 * BRICK_ROWS * pitch + cb0] * y2;
    }
    __syncwarp();

    // stage sign-off; the last warp re-arms the stage with the item after next
    if (lane == 0) {
      const int done = atomicAdd(&s_done[stage], 1);
      if (done == CB - 1) {
        s_done[stage] = 0;
        const int after = nxt < nitems ? next_active(nxt + 1) : nitems;
        if (after < nitems) {
          asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
          issue(after, stage);
        }
      }
    }

    // p.Ap of this slot: one partial per (tile, warp) in plane FOLD_PLANE of the slot's partial-sum buffer.  No
    // ticket, no fence: the sum is folded in a fixed order by k_fold_spmv (one warp per slot), which the kernel
    // boundary orders after these stores.
    red = warp_sum(red);
    if (lane == 0) T.partial[((size_t)slot * NRED + FOLD_PLANE) * T.nblk_max + tile * CB + w] = red;
    cur = nxt;
    ++k;
  }
}

// ------------------------------------------------------------------------------------------------
// k_spmv_dot_tmac: the TMA-tiled implicit SpMV, second generation (default).  Differences to k_spmv_dot_tma, each
// taken from the ncu captures under profiles/r01c_* / r01d_*:
//  * the three pure-material row blocks are passed BY VALUE as a __grid_constant__ kernel parameter (6.5 KB in the
//    constant bank): the row-block value is an operand of the DFMA itself (uniform register filled by LDCU: no
//    LSU/shared-memory traffic, no vector registers).  The material is a compile-time constant of a three-way branch
//    (lanes of a warp whose chunks sit in different materials take their branches one after the other).  Halves the
//    shared-memory wavefronts, 128..168 registers instead of 221;
//  * TN = 7 or 8 nodes per thread and explicit tile descriptors with two lane shapes (8y x 4z and 4y x 8z rows per
//    warp, both with a 60-row brick): at 30^3 (28 interior nodes per edge) the 8-node / 8x4 tiling of k_spmv_dot_tma
//    executes 28672 node slots for 21952 nodes (77 %); 4 chunks of 7 nodes and a 4y x 8z strip for the last four
//    y rows execute 22400 (98 %);
//  * NSTAGE 1 (more resident blocks hide the TMA latency) or 2 (mbarrier pipeline), MINB resident blocks per SM, RU
//    unroll of the loop over the 9 neighbour rows: variants measured by tools/bench_imp_spmv.py;
//  * p.Ap of a thread is accumulated per component (3 independent chains instead of one of 72 additions).
// Arithmetic of Ap (FMA order per accumulator) is that of k_spmv_dot / k_spmv_dot_tma: Ap is bit-identical.
// ------------------------------------------------------------------------------------------------
struct PureRows {
  double a[3 * RB_LEN];
};

// x pitch of the p brick: even (rows stay 16-B aligned) with an odd number of 16-B units, so that the 128-bit loads
// of the 8 y-rows of a quarter-warp fall into 8 different bank groups
__host__ __device__ constexpr int tmac_pitch(int tn, int cb) {
  int p = (tn * cb + 2 + (tn & 1) + 1) & ~1;  // odd tn: the brick may start one node early (16-B aligned TMA origin)
  if (((p / 2) & 1) == 0) p += 2;
  return p;
}

struct TileInfo2 {
  int cb, tn, nchunk, pitch, ntiles;
  const int4 *tiles;      // [ntiles] x: first chunk, y / z: interior coordinates of the tile origin, w: lane shape
  const int *chunk_pure;  // [niz][niy][nchunk]: pure id | (TN-bit mask of the nodes to fix up) << 8
  const int *fix_ptr;     // [ntiles + 1]
  // fix-up tasks: one or two interface nodes that share a row block.  x: position of node A inside the tile
  // (lx | ry << 8 | rz << 12), y: position of node B or -1, z: row-block id
  const int4 *fix;
};

// S: the thread's window [TN*w, TN*w + TN + 2) starts S doubles after the 16-B aligned address the loads start from
template <int MAT, int RU, int PITCH, int TN, int S>
__device__ __forceinline__ void tile_rows_apply_const(const PureRows &R, const double *__restrict__ brick, int bx0,
                                                      int ry, int rz, int by_rows, double (&acc)[8][3]) {
#pragma unroll RU
  for (int row = 0; row < 9; ++row) {
    const int dk = row / 3, dj = row - dk * 3;  // 0..2 (offset + 1)
    const int rbase = ((rz + dk) * by_rows + (ry + dj)) * PITCH + bx0;
    double pv[3][10];
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      const double2 *s2 = reinterpret_cast<const double2 *>(brick + d * (BRICK_ROWS * PITCH) + rbase);
#pragma unroll
      for (int h = 0; h < 5; ++h) {
        const double2 v = s2[h];
        pv[d][2 * h] = v.x;
        pv[d][2 * h + 1] = v.y;
      }
    }
#pragma unroll
    for (int di = 0; di < 3; ++di) {
      const double *a = &R.a[MAT * RB_LEN + (row * 3 + di) * RB_NBR];
#pragma unroll
      for (int fj = 0; fj < 3; ++fj) {
#pragma unroll
        for (int t = 0; t < TN; ++t) {
          const double pval = pv[fj][t + di + S];
          acc[t][0] += a[fj] * pval;
          acc[t][1] += a[3 + fj] * pval;
          acc[t][2] += a[6 + fj] * pval;
        }
      }
    }
  }
}

// ODD: n0 is odd, i.e. the pairs (t, t + 1) with odd t are the 16-B aligned ones
template <int TN, int ODD>
__device__ __forceinline__ void store_ap_pairs(double *__restrict__ Ap, size_t npad, int n0, int keep,
                                               const double (&acc)[8][3]) {
#pragma unroll
  for (int t = 0; t < TN; ++t) {
    const bool pair_start = ((t + ODD) & 1) == 0 && t + 1 < TN;
    const bool pair_second = ((t + ODD) & 1) == 1 && t >= 1;
    if (pair_start) {
      const int both = (keep >> t) & 3;
      if (both == 3) {
#pragma unroll
        for (int d = 0; d < 3; ++d)
          *reinterpret_cast<double2 *>(Ap + d * npad + n0 + t) = make_double2(acc[t][d], acc[t + 1][d]);
      } else if (both & 1) {
#pragma unroll
        for (int d = 0; d < 3; ++d) Ap[d * npad + n0 + t] = acc[t][d];
      } else if (both & 2) {
#pragma unroll
        for (int d = 0; d < 3; ++d) Ap[d * npad + n0 + t + 1] = acc[t + 1][d];
      }
    } else if (!pair_second) {  // a single node at either end
      if ((keep >> t) & 1) {
#pragma unroll
        for (int d = 0; d < 3; ++d) Ap[d * npad + n0 + t] = acc[t][d];
      }
    }
  }
}

template <int CB, int TN, int NSTAGE, int MINB, int RU>
__global__ void __launch_bounds__(32 * CB, MINB)
    k_spmv_dot_tmac(const __grid_constant__ MeshConst P, const Lst L, int n_list, SlotTables T, VecPool V, TileInfo2 ti,
                    const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                    const __grid_constant__ PureRows R, int ntl, int rs, int force) {
  extern __shared__ unsigned char s_raw[];
  constexpr int pitch = tmac_pitch(TN, CB);
  constexpr int BRICK = 3 * BRICK_ROWS * pitch;  // doubles
  constexpr int STAGE_BYTES = (BRICK * 8 + 127) / 128 * 128;
  __shared__ uint64_t s_full[NSTAGE];
  __shared__ int s_slot[TMA_MAX_RS];
  __shared__ int s_done[NSTAGE];
  unsigned char *s_base = s_raw + ((128u - ((unsigned)__cvta_generic_to_shared(s_raw) & 127u)) & 127u);
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tile0 = blockIdx.x * ntl;
  const int nitems = min(ntl, ti.ntiles - tile0) * rs;

  if ((int)threadIdx.x < rs) {
    const int yy = (int)blockIdx.y * rs + (int)threadIdx.x + L.yoff;
    const int cnt = L.dcount ? min(*L.dcount, n_list + L.yoff) : n_list + L.yoff;
    int slot = yy < cnt ? L.list[yy] : -1;
    if (slot >= 0 && !force && !T.state[slot].cg_active) slot = -1;
    s_slot[threadIdx.x] = slot;
  }
  if (threadIdx.x == 0) {
#pragma unroll
    for (int q = 0; q < NSTAGE; ++q) {
      s_done[q] = 0;
      mbar_init(&s_full[q], 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
  }
  __syncthreads();

  auto next_active = [&](int i) {
    while (i < nitems && s_slot[i % rs] < 0) ++i;
    return i;
  };
  auto issue = [&](int item, int stage) {  // one thread
    const int4 td = __ldg(&ti.tiles[tile0 + item / rs]);
    const int slot = s_slot[item % rs];
    mbar_expect_tx(&s_full[stage], BRICK * 8);
    // the box starts at an even x (16-B aligned global address): with TN = 7 odd tile origins start one node early
    tma_load_5d(s_base + stage * STAGE_BYTES, td.w ? &tmap_b : &tmap_a, &s_full[stage], (td.x * TN) & ~1, td.y, td.z, 0,
                slot);
  };

  const size_t npad = P.nn_pad;
  int cur = next_active(0), k = 0;
  if (threadIdx.x == 0 && !(force & 4)) {
    int it = cur;
#pragma unroll
    for (int q = 0; q < NSTAGE; ++q) {
      if (it < nitems) issue(it, q);
      it = it < nitems ? next_active(it + 1) : nitems;
    }
  }
  // per-tile thread state (independent of the slot: with ntl == 1 it is computed once per block)
  int cur_tile = -1, pure = 0, keep = 0, n0 = 0, f0 = 0, f1 = 0, nfix0 = 0, ry = 0, rz = 0, by_rows = TILE_Y + 2;
  int xoff = 0;  // 1: the brick starts one node before the tile (see issue())
  const bool dbg_skip_compute = force & 2, dbg_skip_load = force & 4;  // measurement only (tools/bench_imp_spmv.py)
  bool work = false;
  while (cur < nitems) {
    const int nxt = next_active(cur + 1);
    const int stage = k % NSTAGE;
    const int tile = tile0 + cur / rs, slot = s_slot[cur % rs];
    if (tile != cur_tile) {
      cur_tile = tile;
      const int4 td = __ldg(&ti.tiles[tile]);
      // lane shape 0: 8 y-rows x 4 z-rows per warp (brick 10 x 6 rows); 1: 4 y-rows x 8 z-rows (brick 6 x 10 rows)
      ry = td.w ? (lane & 3) : (lane & 7);
      rz = td.w ? (lane >> 2) : (lane >> 3);
      by_rows = td.w ? TILE_Z + 2 : TILE_Y + 2;
      const int c = td.x + w, jj = td.y + ry, kk = td.z + rz;
      work = c < ti.nchunk && jj < P.niy && kk < P.niz;
      const int info = work ? __ldg(&ti.chunk_pure[(kk * P.niy + jj) * ti.nchunk + c]) : 0;
      pure = info & 0xff;
      const int ii0 = c * TN;
      const int nvalid = min(TN, P.nix - ii0);
      keep = work ? (((1 << nvalid) - 1) & ~(info >> 8)) : 0;  // nodes this thread stores itself
      n0 = (kk + 1) * P.nxny + (jj + 1) * P.nx + ii0 + 1;
      f0 = __ldg(&ti.fix_ptr[tile]);
      f1 = __ldg(&ti.fix_ptr[tile + 1]);
      nfix0 = (td.y + 1) * P.nx + (td.z + 1) * P.nxny + td.x * TN + 1;  // node of tile position 0
      xoff = (td.x * TN) & 1;
    }
    const double *s_brick = reinterpret_cast<const double *>(s_base + stage * STAGE_BYTES);
    if (!dbg_skip_load) mbar_wait(&s_full[stage], (k / NSTAGE) & 1);

    double red0 = 0.0, red1 = 0.0, red2 = 0.0;
    double *Ap = V.Ap + (size_t)slot * V.vstride;
    if (work && !dbg_skip_compute) {
      double acc[8][3];
#pragma unroll
      for (int t = 0; t < 8; ++t) acc[t][0] = acc[t][1] = acc[t][2] = 0.0;
      const int xs = TN * w + xoff, bx0 = xs & ~1;
      if ((TN & 1) && (xs & 1)) {  // odd window start: loads begin one double earlier (warp-uniform)
        if (pure == 0)
          tile_rows_apply_const<0, RU, pitch, TN, TN & 1>(R, s_brick, bx0, ry, rz, by_rows, acc);
        else if (pure == 1)
          tile_rows_apply_const<1, RU, pitch, TN, TN & 1>(R, s_brick, bx0, ry, rz, by_rows, acc);
        else
          tile_rows_apply_const<2, RU, pitch, TN, TN & 1>(R, s_brick, bx0, ry, rz, by_rows, acc);
      } else {
        if (pure == 0)
          tile_rows_apply_const<0, RU, pitch, TN, 0>(R, s_brick, bx0, ry, rz, by_rows, acc);
        else if (pure == 1)
          tile_rows_apply_const<1, RU, pitch, TN, 0>(R, s_brick, bx0, ry, rz, by_rows, acc);
        else
          tile_rows_apply_const<2, RU, pitch, TN, 0>(R, s_brick, bx0, ry, rz, by_rows, acc);
      }
      const int cbase = ((rz + 1) * by_rows + (ry + 1)) * pitch + xs + 1;
#pragma unroll
      for (int t = 0; t < TN; ++t) {
        if ((keep >> t) & 1) {
          red0 += s_brick[cbase + t] * acc[t][0];
          red1 += s_brick[BRICK_ROWS * pitch + cbase + t] * acc[t][1];
          red2 += s_brick[2 * BRICK_ROWS * pitch + cbase + t] * acc[t][2];
        }
      }
      // Ap stores.  Lanes of a warp write different grid rows, so every store instruction costs one sector per lane
      // whatever its width (ncu r01d: the L1 data pipe, not the FP64 pipe, limits k_spmv_dot_tma): nodes are paired
      // into 16-B stores wherever both nodes are kept and the pair is 16-B aligned (node index even)
      if (n0 & 1) {
        store_ap_pairs<TN, 1>(Ap, npad, n0, keep, acc);
      } else {
        store_ap_pairs<TN, 0>(Ap, npad, n0, keep, acc);
      }
    }
    // fix-up: the nodes of this tile that sit on a material interface (the brick holds all the p values they need;
    // their row blocks come from the L2-resident table).  One task per thread: ONE row block applied to one or two
    // nodes (6 independent accumulator chains instead of 3, half the row-block loads); tasks are sorted by row
    // block, so neighbouring lanes fetch the same sectors.
    for (int f = f0 + w * 32 + lane; f < (dbg_skip_compute ? f0 : f1); f += 32 * CB) {
      const int4 e = __ldg(&ti.fix[f]);
      const bool two = e.y >= 0;
      const int eb = two ? e.y : e.x;
      const int lxa = e.x & 0xff, fya = (e.x >> 8) & 0xf, fza = (e.x >> 12) & 0xf;
      const int lxb = eb & 0xff, fyb = (eb >> 8) & 0xf, fzb = (eb >> 12) & 0xf;
      const double2 *a2 = reinterpret_cast<const double2 *>(V.rows + (size_t)e.z * RB_LEN);
      double ya0 = 0.0, ya1 = 0.0, ya2 = 0.0, yb0 = 0.0, yb1 = 0.0, yb2 = 0.0;
#pragma unroll 3
      for (int row = 0; row < 9; ++row) {
        const int dk = row / 3, dj = row - dk * 3;
        const int ra = ((fza + dk) * by_rows + (fya + dj)) * pitch + lxa + xoff;
        const int rb = ((fzb + dk) * by_rows + (fyb + dj)) * pitch + lxb + xoff;
#pragma unroll
        for (int di = 0; di < 3; ++di) {
          double av[10];
#pragma unroll
          for (int q = 0; q < 5; ++q) {
            const double2 v = __ldg(a2 + (row * 3 + di) * (RB_NBR / 2) + q);
            av[2 * q] = v.x;
            av[2 * q + 1] = v.y;
          }
          const double pxa = s_brick[ra + di], pya = s_brick[BRICK_ROWS * pitch + ra + di],
                       pza = s_brick[2 * BRICK_ROWS * pitch + ra + di];
          const double pxb = s_brick[rb + di], pyb = s_brick[BRICK_ROWS * pitch + rb + di],
                       pzb = s_brick[2 * BRICK_ROWS * pitch + rb + di];
          ya0 += av[0] * pxa;
          yb0 += av[0] * pxb;
          ya0 += av[1] * pya;
          yb0 += av[1] * pyb;
          ya0 += av[2] * pza;
          yb0 += av[2] * pzb;
          ya1 += av[3] * pxa;
          yb1 += av[3] * pxb;
          ya1 += av[4] * pya;
          yb1 += av[4] * pyb;
          ya1 += av[5] * pza;
          yb1 += av[5] * pzb;
          ya2 += av[6] * pxa;
          yb2 += av[6] * pxb;
          ya2 += av[7] * pya;
          yb2 += av[7] * pyb;
          ya2 += av[8] * pza;
          yb2 += av[8] * pzb;
        }
      }
      {
        const int n = nfix0 + fza * P.nxny + fya * P.nx + lxa;
        const int cb0 = ((fza + 1) * by_rows + (fya + 1)) * pitch + lxa + xoff + 1;
        Ap[n] = ya0;
        Ap[npad + n] = ya1;
        Ap[2 * npad + n] = ya2;
        red0 += s_brick[cb0] * ya0;
        red1 += s_brick[BRICK_ROWS * pitch + cb0] * ya1;
        red2 += s_brick[2 * BRICK_ROWS * pitch + cb0] * ya2;
      }
      if (two) {
        const int n = nfix0 + fzb * P.nxny + fyb * P.nx + lxb;
        const int cb0 = ((fzb + 1) * by_rows + (fyb + 1)) * pitch + lxb + xoff + 1;
        Ap[n] = yb0;
        Ap[npad + n] = yb1;
        Ap[2 * npad + n] = yb2;
        red0 += s_brick[cb0] * yb0;
        red1 += s_brick[BRICK_ROWS * pitch + cb0] * yb1;
        red2 += s_brick[2 * BRICK_ROWS * pitch + cb0] * yb2;
      }
    }
    __syncwarp();

    // stage sign-off; the last warp re-arms the stage with the NSTAGE-th active item after this one
    if (lane == 0) {
      const int done = atomicAdd(&s_done[stage], 1);
      if (done == CB - 1) {
        s_done[stage] = 0;
        int after = nxt;
#pragma unroll
        for (int q = 1; q < NSTAGE; ++q) after = after < nitems ? next_active(after + 1) : nitems;
        if (after < nitems && !dbg_skip_load) {
          asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
          issue(after, stage);
        }
      }
    }
    // p.Ap of this slot: one partial per (tile, warp) in plane FOLD_PLANE of the slot's partial-sum buffer, folded in
    // a fixed order by k_fold_spmv (one warp per slot), which the kernel boundary orders after these stores
    const double red = warp_sum((red0 + red1) + red2);
    if (lane == 0) T.partial[((size_t)slot * NRED + FOLD_PLANE) * T.nblk_max + tile * CB + w] = red;
    cur = nxt;
    ++k;
  }
}

__device__ __forceinline__ double imp_kk(const MeshConst &P, const VecPool &V, int n, int d) {
  int i, j, k;
  node_ijk(P, n, i, j, k);
  if (on_boundary(P, i, j, k)) return 1.0;  // 1 / 1
  const int m = interior_index(P, i, j, k);
  return __ldg(&V.rkinv[__ldg(&V.rowid[m]) * 3 + d]);
}

// cg_update / cg_pupdate of the implicit operator: k = 1/diag comes from the row table and z = k r is recomputed
// instead of being stored (the same multiplication => the same bits).  x += alpha p (src/ell.cpp:102) is deferred to
// the p update of the same iteration (or k_cg_finish after the last one), where p is read anyway:
// 72 + 120 B/node per iteration instead of 144 + 72.
__global__ void __launch_bounds__(NT, 4)
    k_cg_update_imp(const __grid_constant__ MeshConst P, const Lst L, SlotTables T, VecPool V, int nfold) {
  __shared__ double sm[NRED * (NT / 32)];
    __shared__ double s_alpha;
  const int slot = slot_of(L);
  if (slot < 0) return;
  mgpu_slot_state *st = &T.state[slot];
  if (!st->cg_active) return;
  const size_t vo = (size_t)slot * V.vstride;
  // UPD_VPT nodes per thread (ncu r01h: one node per thread left the kernel latency-bound at 4.0 TB/s -- three
  // dependent loads (row id -> 1/diag -> product) and one block reduction per 256 nodes): all vector loads and the
  // row ids go out first, the fold below (and its barrier) then overlaps their latency
  const int nb = blockIdx.x * (NT * UPD_VPT) + threadIdx.x;
  double rr[UPD_VPT][3], ap[UPD_VPT][3];
  int rid[UPD_VPT];  // row-block id of the node (-1: boundary node, k = 1)
#pragma unroll
  for (int v = 0; v < UPD_VPT; ++v) {
    const int n = nb + v * NT;
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      const size_t ix = vo + (size_t)d * P.nn_pad + (n < P.nn ? n : 0);
      rr[v][d] = V.r[ix];
      ap[v][d] = V.Ap[ix];
    }
    rid[v] = -1;
    if (n < P.nn) {
      int i, j, k;
      node_ijk(P, n, i, j, k);
      if (!on_boundary(P, i, j, k)) rid[v] = __ldg(&V.rowid[interior_index(P, i, j, k)]);
    }
  }
  double alpha;
  if (nfold > 0) {
    // the TMA SpMV left per-(tile, warp) partials of p.Ap: every block folds them itself (same order => same bits)
    if (threadIdx.x < 32) {
      const double pAp = fold_partials(T.partial + ((size_t)slot * NRED + FOLD_PLANE) * T.nblk_max, nfold);
      if (threadIdx.x == 0) {
        const double a = st->rz / pAp;  // src/ell.cpp:100
        s_alpha = a;
        if (blockIdx.x == 0) {
          st->pAp = pAp;
          st->alpha = a;
        }
      }
    }
    __syncthreads();
    alpha = s_alpha;
  } else {
    alpha = st->alpha;
  }
  double red[2] = {0.0, 0.0};
#pragma unroll
  for (int v = 0; v < UPD_VPT; ++v) {
    const int n = nb + v * NT;
    if (n < P.nn) {
#pragma unroll
      for (int d = 0; d < 3; ++d) {
        const size_t ix = vo + (size_t)d * P.nn_pad + n;
        const double r = rr[v][d] - alpha * ap[v][d];
        V.r[ix] = r;
        const double kd = rid[v] < 0 ? 1.0 : __ldg(&V.rkinv[rid[v] * 3 + d]);  // as imp_kk: 1 / diagonal
        const double z = __dmul_rn(kd, r);  // rounded product, as when z is stored (k_cg_update)
        red[0] += z * z;
        red[1] += r * z;
      }
    }
  }
  // per-block partial sums only: no ticket, no fence -- k_fold_update (one warp per slot) folds them in a fixed order
  // and runs the scalar tail; the kernel boundary orders it after these stores
  block_sum<2>(red, sm);
  if (threadIdx.x == 0) {
    double *partial = T.partial + (size_t)slot * NRED * T.nblk_max;
    partial[blockIdx.x] = red[0];
    partial[T.nblk_max + blockIdx.x] = red[1];
  }
}

__global__ void __launch_bounds__(NT)
    k_cg_pupdate_imp(const __grid_constant__ MeshConst P, const Lst L, SlotTables T, VecPool V) {
  const int slot = slot_of(L);
  if (slot < 0) return;
  const mgpu_slot_state *st = &T.state[slot];
  if (!st->cg_active) return;
  const double beta = st->beta, alpha = st->alpha;
  const size_t vo = (size_t)slot * V.vstride;
  const int n = blockIdx.x * NT + threadIdx.x;
  if (n >= P.nn) return;
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    const size_t ix = vo + (size_t)d * P.nn_pad + n;
    const double z = __dmul_rn(imp_kk(P, V, n, d), V.r[ix]);  // never fused into the FMA below
    const double pp = V.p[ix];
    V.du[ix] = fma(alpha, pp, V.du[ix]);  // x += alpha p of this iteration (src/ell.cpp:102)
    V.p[ix] = z + beta * pp;
  }
}

// Arbitrary user matrix in the reference's own layout vals[row*81 + slot] (host-pointer ell_mvp / ell_solve_cgpd
// API): every row is read, neighbours outside the grid are skipped -- their stored value is 0 and the reference
// multiplies it by x[fj] (src/ell-common.cpp:102-130 + src/ell.cpp:39-41).  Test-size utility, not the hot path.
__global__ void __launch_bounds__(NT)
    k_spmv_generic(const __grid_constant__ MeshConst P, const Lst L, SlotTables T, VecPool V, int force) {
  __shared__ double sm[NRED * (NT / 32)];
  __shared__ int sflag;
  const int slot = slot_of(L);
  if (slot < 0) return;
  mgpu_slot_state *st = &T.state[slot];
  if (!force && !st->cg_active) return;
  const size_t vo = (size_t)slot * V.vstride;
  const double *p = V.p + vo;
  double *Ap = V.Ap + vo;
  const int n = blockIdx.x * NT + threadIdx.x;
  double red[1] = {0.0};
  if (n < P.nn) {
    int i, j, k;
    node_ijk(P, n, i, j, k);
    const size_t npad = P.nn_pad;
    double y[3] = {0.0, 0.0, 0.0};
    for (int nbr = 0; nbr < 27; ++nbr) {
      const int di = nbr % 3 - 1, dj = (nbr / 3) % 3 - 1, dk = nbr / 9 - 1;
      const int ii = i + di, jj = j + dj, kk = k + dk;
      if (ii < 0 || ii >= P.nx || jj < 0 || jj >= P.ny || kk < 0 || kk >= P.nz) continue;
      const int q = n + di + dj * P.nx + dk * P.nxny;
      const double px[3] = {p[q], p[npad + q], p[2 * npad + q]};
      for (int fi = 0; fi < 3; ++fi)
        for (int fj = 0; fj < 3; ++fj) y[fi] += V.gen[((size_t)n * 3 + fi) * 81 + nbr * 3 + fj] * px[fj];
    }
    Ap[n] = y[0];
    Ap[npad + n] = y[1];
    Ap[2 * npad + n] = y[2];
    red[0] = p[n] * y[0] + p[npad + n] * y[1] + p[2 * npad + n] * y[2];
  }
  double *partial = T.partial + (size_t)slot * NRED * T.nblk_max;
  if (grid_sum<1>(red, partial, T.nblk_max, &st->ticket, sm, &sflag) && threadIdx.x == 0) {
    tail_spmv(st, red[0]);
  }
}

// r -= alpha Ap ; z = k r ; z.z ; r.z   (src/ell.cpp:103-110), then the scalar tail of the iteration and the
// loop-head test of the next one (src/ell.cpp:93-94,108-119).  x += alpha p (src/ell.cpp:102) does not feed any
// of these: it is applied by k_cg_pupdate of the same iteration, which reads p anyway, or -- after the last
// iteration of a slot, whose p update is skipped -- by k_cg_finish.  Same FMA, same bits, 24 B/node less traffic.
__global__ void __launch_bounds__(NT, 4)
    k_cg_update(const __grid_constant__ MeshConst P, const Lst L, SlotTables T, VecPool V) {
  __shared__ double sm[NRED * (NT / 32)];
    const int slot = slot_of(L);
  if (slot < 0) return;
  mgpu_slot_state *st = &T.state[slot];
  if (!st->cg_active) return;
  const double alpha = st->alpha;
  const size_t vo = (size_t)slot * V.vstride;
  // UPD_VPT nodes per thread, same node order and summation order as k_cg_update_imp (bit-identical sums)
  const int nb = blockIdx.x * (NT * UPD_VPT) + threadIdx.x;
  double rr[UPD_VPT][3], ap[UPD_VPT][3], kk[UPD_VPT][3];
#pragma unroll
  for (int v = 0; v < UPD_VPT; ++v) {
    const int n = nb + v * NT;
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      const size_t ix = vo + (size_t)d * P.nn_pad + (n < P.nn ? n : 0);
      rr[v][d] = V.r[ix];
      ap[v][d] = V.Ap[ix];
      kk[v][d] = V.k[ix];
    }
  }
  double red[2] = {0.0, 0.0};
#pragma unroll
  for (int v = 0; v < UPD_VPT; ++v) {
    const int n = nb + v * NT;
    if (n < P.nn) {
#pragma unroll
      for (int d = 0; d < 3; ++d) {
        const size_t ix = vo + (size_t)d * P.nn_pad + n;
        const double r = rr[v][d] - alpha * ap[v][d];
        V.r[ix] = r;
        const double z = kk[v][d] * r;
        V.z[ix] = z;
        red[0] += z * z;
        red[1] += r * z;
      }
    }
  }
  // per-block partial sums only: no ticket, no fence -- k_fold_update (one warp per slot) folds them in a fixed order
  // and runs the scalar tail; the kernel boundary orders it after these stores
  block_sum<2>(red, sm);
  if (threadIdx.x == 0) {
    double *partial = T.partial + (size_t)slot * NRED * T.nblk_max;
    partial[blockIdx.x] = red[0];
    partial[T.nblk_max + blockIdx.x] = red[1];
  }
}

// p = z + beta p (src/ell.cpp:113); skipped once the slot has left the loop (p is dead then).
__global__ void __launch_bounds__(NT)
    k_cg_pupdate(const __grid_constant__ MeshConst P, const Lst L, SlotTables T, VecPool V) {
  const int slot = slot_of(L);
  if (slot < 0) return;
  const mgpu_slot_state *st = &T.state[slot];
  if (!st->cg_active) return;
  const double beta = st->beta, alpha = st->alpha;
  const size_t vo = (size_t)slot * V.vstride;
  const int n = blockIdx.x * NT + threadIdx.x;
  if (n >= P.nn) return;
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    const size_t ix = vo + (size_t)d * P.nn_pad + n;
    const double pp = V.p[ix];
    V.du[ix] = fma(alpha, pp, V.du[ix]);  // x += alpha p of this iteration (src/ell.cpp:102)
    V.p[ix] = V.z[ix] + beta * pp;
  }
}

// x += alpha p of the LAST iteration of every slot of the list that iterated at all: its p update was skipped because
// the slot had left the loop (cg_active == 0), so the deferred update of du is still pending.  Runs once per solve.
__global__ void __launch_bounds__(NT)
    k_cg_finish(const __grid_constant__ MeshConst P, const Lst L, SlotTables T, VecPool V) {
  const int slot = slot_of(L);
  if (slot < 0) return;
  const mgpu_slot_state *st = &T.state[slot];
  if (st->cg_its <= 0 || st->cg_active) return;
  const double alpha = st->alpha;
  const size_t vo = (size_t)slot * V.vstride;
  const int n = blockIdx.x * NT + threadIdx.x;
  if (n >= P.nn) return;
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    const size_t ix = vo + (size_t)d * P.nn_pad + n;
    V.du[ix] = fma(alpha, V.p[ix], V.du[ix]);
  }
}

// u += du (src/solve.cpp:73) and newton.solver_its += cg_its (src/solve.cpp:71)
__global__ void __launch_bounds__(NT)
    k_axpy_u(const __grid_constant__ MeshConst P, const Lst L, SlotTables T, VecPool V) {
  const int slot = slot_of(L);
  if (slot < 0) return;
  mgpu_slot_state *st = &T.state[slot];
  if (!st->nr_active) return;
  const size_t vo = (size_t)slot * V.vstride;
  const int n = blockIdx.x * NT + threadIdx.x;
  if (blockIdx.x == 0 && threadIdx.x == 0) st->solver_its += st->cg_its;
  if (n >= P.nn) return;
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    const size_t ix = vo + (size_t)d * P.nn_pad + n;
    V.u[ix] += V.du[ix];
  }
}

// ------------------------------------------------------------------------------------------------
// calc_ave_stress (src/average.cpp:58-82): thread per element
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(NT)
    k_ave_stress(const __grid_constant__ MeshConst P, const Lst L, SlotTables T,
                 const double *__restrict__ u_pool, size_t vstride, const int *__restrict__ elem_type) {
  __shared__ double sm[NRED * (NT / 32)];
  __shared__ int sflag;
  const int slot = slot_of(L);
  if (slot < 0) return;
  mgpu_slot_state *st = &T.state[slot];
  const double *u = u_pool + (size_t)slot * vstride;
  const double *vars = T.vars_old[slot];
  const int e = blockIdx.x * NT + threadIdx.x;
  double red[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
  // in a slab every element layer is summed by exactly one rank (the owner of its lower node plane)
  if (e < P.nelem && e / (P.nex * P.ney) >= P.ez_own_lo && e / (P.nex * P.ney) < P.ez_own_hi) {
    const int ez = e / (P.nex * P.ney);
    const int r = e - ez * P.nex * P.ney;
    const int ey = r / P.nex, ex = r - ey * P.nex;
    double ue[24];
    gather_ue(P, u, ex, ey, ez, ue);
    const int type = __ldg(&elem_type[e]);
    const mpp_material m = P.mat[type];
    const int nv = mat_nvar(m.type);
#pragma unroll 1
    for (int gp = 0; gp < 8; ++gp) {
      double eps[6], sig[6], vbuf[7];
      gp_strain(P.dsh[gp], ue, eps);
      const double *v = fetch_vars(vars, P.nelem_pad, e, gp, nv, vbuf);
      mat_stress(m, eps, v, sig);
#pragma unroll
      for (int q = 0; q < 6; ++q) red[q] = add_(red[q], mul_(sig[q], P.wg));  // src/average.cpp:70-72, uncontracted
    }
  }
  double *partial = T.partial + (size_t)slot * NRED * T.nblk_max;
  if (grid_sum<6>(red, partial, T.nblk_max, &st->ticket, sm, &sflag) && threadIdx.x == 0) {
#pragma unroll
    for (int q = 0; q < 6; ++q) {
      if (P.slab)
        T.red[slot * 8 + q] = red[q];
      else
        T.stress[slot * 6 + q] = red[q] / 1.0;  // vol_tot = 1 (src/micropp.cpp:58)
    }
  }
}

// ------------------------------------------------------------------------------------------------
// calc_vars_new (src/update.cpp:33-56): thread per element; write = 0 only raises the non-linear flag
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(NT)
    k_vars_new(const __grid_constant__ MeshConst P, const Lst L, SlotTables T,
               const double *__restrict__ u_pool, size_t vstride, const int *__restrict__ elem_type, int write) {
  const int slot = slot_of(L);
  if (slot < 0) return;
  mgpu_slot_state *st = &T.state[slot];
  const double *u = u_pool + (size_t)slot * vstride;
  const double *vars = T.vars_old[slot];
  double *vnew = write ? T.vars_new[slot] : nullptr;
  const int e = blockIdx.x * NT + threadIdx.x;
  bool nl = false;
  if (e < P.nelem) {
    const int type = __ldg(&elem_type[e]);
    const mpp_material m = P.mat[type];
    if (m.type != MPP_ELASTIC) {
      const int ez = e / (P.nex * P.ney);
      const int r = e - ez * P.nex * P.ney;
      const int ey = r / P.nex, ex = r - ey * P.nex;
      double ue[24];
      gather_ue(P, u, ex, ey, ez, ue);
      const int nv = mat_nvar(m.type);
#pragma unroll 1
      for (int gp = 0; gp < 8; ++gp) {
        double eps[6], vbuf[7], vn[7];
        gp_strain(P.dsh[gp], ue, eps);
        const double *v = fetch_vars(vars, P.nelem_pad, e, gp, nv, vbuf);
        // plastic evolute writes nothing without an old state (src/material.cpp:180-183)
        const bool wr = vnew && (m.type == MPP_DAMAGE || v != nullptr);
        nl |= mat_evolute(m, eps, v, wr ? vn : nullptr);
        if (wr)
          for (int q = 0; q < nv; ++q) vnew[(size_t)(q * 8 + gp) * P.nelem_pad + e] = vn[q];
      }
    }
  }
  if (__syncthreads_or(nl) && threadIdx.x == 0) atomicOr(&st->nl_flag, 1);
}

// ------------------------------------------------------------------------------------------------
// calc_fields (src/average.cpp:85-112): element averages of strain and stress for the VTU output; thread per
// element, out[e*6+v] (strain) and out[6*nelem + e*6+v] (stress) in the reference's element-major layout
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(NT)
    k_elem_fields(const __grid_constant__ MeshConst P, int slot, SlotTables T, const double *__restrict__ u_pool,
                  size_t vstride, const int *__restrict__ elem_type, double ivol, double *__restrict__ out) {
  const double *u = u_pool + (size_t)slot * vstride;
  const double *vars = T.vars_old[slot];
  const int e = blockIdx.x * NT + threadIdx.x;
  if (e >= P.nelem) return;
  const int ez = e / (P.nex * P.ney);
  const int r = e - ez * P.nex * P.ney;
  const int ey = r / P.nex, ex = r - ey * P.nex;
  double ue[24];
  gather_ue(P, u, ex, ey, ez, ue);
  const mpp_material m = P.mat[__ldg(&elem_type[e])];
  const int nv = mat_nvar(m.type);
  double ea[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0}, sa[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
#pragma unroll 1
  for (int gp = 0; gp < 8; ++gp) {
    double eps[6], sig[6], vbuf[7];
    gp_strain(P.dsh[gp], ue, eps);
    const double *v = fetch_vars(vars, P.nelem_pad, e, gp, nv, vbuf);
    mat_stress(m, eps, v, sig);
#pragma unroll
    for (int q = 0; q < 6; ++q) {
      ea[q] = add_(ea[q], mul_(eps[q], P.wg));
      sa[q] = add_(sa[q], mul_(sig[q], P.wg));
    }
  }
#pragma unroll
  for (int q = 0; q < 6; ++q) {
    out[(size_t)e * 6 + q] = mul_(ea[q], ivol);
    out[(size_t)(P.nelem + e) * 6 + q] = mul_(sa[q], ivol);
  }
}

// slab mode: scalar tails after the cross-rank all-reduce of T.red (kind: 0 rhs, 1 cg_init, 2 spmv, 3 cg_update,
// 4 average stress)
__global__ void k_tail(const __grid_constant__ MeshConst P, const Lst L, SlotTables T, int kind, int mode) {
  const int slot = slot_of(L);
  if (slot < 0 || threadIdx.x != 0) return;
  mgpu_slot_state *st = &T.state[slot];
  const double *red = T.red + slot * 8;
  switch (kind) {
    case 0:
      if (mode == 1 && !st->nr_active) return;
      tail_rhs(P, st, red[0], mode);
      break;
    case 1: tail_cg_init(P, st, red[0], red[1]); break;
    case 2:
      if (st->cg_active) tail_spmv(st, red[0]);
      break;
    case 3:
      if (st->cg_active) tail_cg_update(P, st, red[0], red[1]);
      break;
    default:
      for (int q = 0; q < 6; ++q) T.stress[slot * 6 + q] = red[q] / 1.0;
      break;
  }
}

// ------------------------------------------------------------------------------------------------
// Slab mode over PEER MEMORY (NVLink P2P): the cross-rank steps of a DPCG iteration without any NCCL call.
//  * every rank owns a small mailbox (SlabMail) that all ranks map (cudaIpc): the slab-local sums of a reducing kernel
//    are posted there with an epoch number (st.release.sys); the tail kernel of every rank waits for the epoch of
//    every mailbox (ld.acquire.sys), adds the partial sums in RANK ORDER (identical bits on all ranks) and runs the
//    reference's scalar logic -- an all-reduce of <= 6 doubles costs two tiny kernels instead of a collective;
//  * the halo planes of p are PULLED from the neighbours' vectors by k_slab_halo_pull once the neighbour has
//    published the epoch of its last p update.
// Write-after-read safety needs no extra flag: a rank overwrites p (next p update) only after the tail of the
// following reduction, which cannot complete before every neighbour has posted its partial sum, i.e. has finished
// the SpMV that followed its pull.  Two mailbox buffers (epoch parity) are enough for the same reason.
// All waits are bounded (about 2 s): a lost rank raises SlabMail::error instead of hanging the GPU.
// ------------------------------------------------------------------------------------------------
struct SlabMail {
  double red[2][8];
  unsigned long long red_epoch, p_epoch;  // published epochs (read by the peers)
  unsigned long long red_local, p_local;  // the owner's own counters: the epochs live on the device, so a whole chunk
                                          // of DPCG iterations is a replayable CUDA graph with constant arguments
  int error, pad;
};
constexpr int SLAB_MAX_RANKS = 16;
struct SlabPeers {
  SlabMail *mail[SLAB_MAX_RANKS];
};

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ bool spin_until(const unsigned long long *flag, unsigned long long epoch) {
  const long long t0 = clock64();
  while (ld_acquire_sys(flag) < epoch) {
    if (clock64() - t0 > 4000000000ll) return false;
    __nanosleep(64);
  }
  return true;
}

__global__ void k_slab_publish_p(SlabMail *mail) {
  const unsigned long long epoch = ++mail->p_local;
  __threadfence_system();
  st_release_sys(&mail->p_epoch, epoch);
}

// blockIdx.y = 0: low halo plane <- neighbour below, 1: high halo plane <- neighbour above
__global__ void __launch_bounds__(NT)
    k_slab_halo_pull(const __grid_constant__ MeshConst P, double *p_own, SlabMail *own, const double *p_lo,
                     long long lo_off, long long lo_npad, const SlabMail *mail_lo, const double *p_hi, long long hi_off,
                     long long hi_npad, const SlabMail *mail_hi) {
  const bool hi = blockIdx.y == 1;
  const double *src = hi ? p_hi : p_lo;
  if (!src) return;
  __shared__ int ok;
  if (threadIdx.x == 0) {
    const unsigned long long epoch = own->p_local;  // as many publishes as this rank has done itself
    ok = spin_until(hi ? &mail_hi->p_epoch : &mail_lo->p_epoch, epoch);
    if (!ok) own->error = 1;
  }
  __syncthreads();
  if (!ok) return;
  const long long off = hi ? hi_off : lo_off, npad = hi ? hi_npad : lo_npad;
  double *dst = p_own + (size_t)(hi ? P.nz - 1 : 0) * P.nxny;
  for (int i = blockIdx.x * NT + threadIdx.x; i < P.nxny; i += gridDim.x * NT) {
#pragma unroll
    for (int d = 0; d < 3; ++d) dst[(size_t)d * P.nn_pad + i] = __ldcv(src + (size_t)d * npad + off + i);
  }
}

__global__ void k_slab_post(const double *red, SlabMail *mail, int k) {
  if (threadIdx.x != 0) return;
  const unsigned long long epoch = ++mail->red_local;
  for (int q = 0; q < k; ++q) mail->red[epoch & 1][q] = red[q];
  __threadfence_system();
  st_release_sys(&mail->red_epoch, epoch);
}

// sum of the posted slab sums in rank order, then the scalar tail (k_tail) on this rank
__global__ void k_slab_gather_tail(const __grid_constant__ MeshConst P, const Lst L, SlotTables T,
                                   const __grid_constant__ SlabPeers peers, SlabMail *own, int nranks, int k,
                                   int kind, int mode) {
  const int slot = slot_of(L);
  if (slot < 0 || threadIdx.x != 0) return;
  const unsigned long long epoch = own->red_local;
  double tot[8] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
  for (int r = 0; r < nranks; ++r) {
    if (!spin_until(&peers.mail[r]->red_epoch, epoch)) own->error = 1;
    for (int q = 0; q < k; ++q) tot[q] += __ldcv(&peers.mail[r]->red[epoch & 1][q]);
  }
  double *red = T.red + slot * 8;
  for (int q = 0; q < k; ++q) red[q] = tot[q];
  mgpu_slot_state *st = &T.state[slot];
  switch (kind) {
    case 0:
      if (mode == 1 && !st->nr_active) return;
      tail_rhs(P, st, red[0], mode);
      break;
    case 1: tail_cg_init(P, st, red[0], red[1]); break;
    case 2:
      if (st->cg_active) tail_spmv(st, red[0]);
      break;
    case 3:
      if (st->cg_active) tail_cg_update(P, st, red[0], red[1]);
      break;
    default:
      for (int q = 0; q < 6; ++q) T.stress[slot * 6 + q] = red[q] / 1.0;
      break;
  }
}

__global__ void k_clear_nl(const int *__restrict__ list, int n, SlotTables T) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) T.state[list[i]].nl_flag = 0;
}

// ------------------------------------------------------------------------------------------------
// stable compaction of a slot list by an activity flag (single block)
// ------------------------------------------------------------------------------------------------
// n_in_dev (optional): device-side length of `in`; count2 (optional): second destination of the new length;
// inside a graph the kernel also drives the WHILE node (cond_set) and counts loop trips (trips).  in == out is
// allowed: every entry of a chunk is read before the barrier that precedes the chunk's writes, and writes never
// run ahead of reads.
__global__ void k_compact(const int *in, int n_in, const int *n_in_dev, int *out, int *count, int *count2,
                          SlotTables T, int mode, cudaGraphConditionalHandle cond, int cond_set, int *trips) {
  if (n_in_dev) n_in = *n_in_dev;
  __shared__ int s_warp[32];
  __shared__ int s_base;
  if (threadIdx.x == 0) s_base = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int start = 0; start < n_in; start += blockDim.x) {
    const int i = start + threadIdx.x;
    int slot = -1, keep = 0;
    if (i < n_in) {
      slot = in[i];
      const mgpu_slot_state *st = &T.state[slot];
      keep = mode ? st->cg_active : st->nr_active;
    }
    const unsigned bal = __ballot_sync(0xffffffffu, keep);
    if (lane == 0) s_warp[w] = __popc(bal);
    __syncthreads();
    int off = s_base;
    for (int ww = 0; ww < w; ++ww) off += s_warp[ww];
    if (keep) out[off + __popc(bal & ((1u << lane) - 1u))] = slot;
    __syncthreads();
    if (threadIdx.x == 0) {
      int tot = 0;
      for (int ww = 0; ww < nw; ++ww) tot += s_warp[ww];
      s_base += tot;
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    if (count) *count = s_base;
    if (count2) *count2 = s_base;
    if (trips) *trips += 1;
    if (cond_set) cudaGraphSetConditional(cond, s_base > 0 ? 1u : 0u);
  }
}

// ------------------------------------------------------------------------------------------------
// bit-exact regeneration of the reference's explicit ELL column table (src/ell-common.cpp:86-137):
// the product never stores it (columns are implicit); this exporter exists for parity tests.
// ------------------------------------------------------------------------------------------------
__global__ void k_ell_cols(int nx, int ny, int nz, int *cols) {
  const int nn = nx * ny * nz;
  const int row = blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= 3 * nn) return;
  const int n = row / 3;
  const int k = n / (nx * ny), r = n - k * nx * ny, j = r / nx, i = r - j * nx;
  for (int nbr = 0; nbr < 27; ++nbr) {
    const int di = nbr % 3 - 1, dj = (nbr / 3) % 3 - 1, dk = nbr / 9 - 1;
    const int ii = i + di, jj = j + dj, kk = k + dk;
    const bool in = ii >= 0 && ii < nx && jj >= 0 && jj < ny && kk >= 0 && kk < nz;
    const int m = in ? n + di + dj * nx + dk * nx * ny : 0;  // out-of-grid neighbours point at node 0
    for (int fj = 0; fj < 3; ++fj) cols[(size_t)row * 81 + nbr * 3 + fj] = m * 3 + fj;
  }
}

}  // namespace

// ================================================================================================
// host side of the thin layer
// ================================================================================================
struct mgpu_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  MeshConst mc;
  int ngp = 0, W = 0;
  bool all_elastic = true;
  bool implicit = false;  // all-elastic RVE served by the implicit operator (no per-slot matrices)
  int mat_slots = 0;      // slots of the explicit matrix pool
  int cg_op = OP_SLOT;    // operator of the DPCG solve in flight (set by mgpu_cg_init)
  int nrows = 0;
  TileInfo tile{};        // tiled implicit SpMV (k_spmv_dot_tile)
  int *d_chunk_id = nullptr, *d_chunk_pure = nullptr, *d_fix_ptr = nullptr;
  int2 *d_fix = nullptr;
  int nfix = 0;           // interior nodes whose row block is not a pure-material one (fix-up list of the TMA kernel)
  int tile_smem = 0;
  int imp_kernel = 1;     // 2 tiled + TMA (default when nx is even), 1 tiled + cp.async, 0 simple (MICROPP_IMP_KERNEL)
  CUtensorMap tmap_p;     // V.p as a rank-5 tensor (x, y, z, component, slot)
  int tma_smem = 0;
  PureRows pure_rows;     // host copy of row blocks 0..2 (kernel parameter of k_spmv_dot_tmac)
  int tma_variant = 0;    // 0: k_spmv_dot_tma (row blocks in shared memory); v >= 1: k_spmv_dot_tmac variant v
  // slab mode over peer memory (mgpu_slab_link)
  SlabMail *slab_mail = nullptr;
  SlabPeers slab_peers{};
  int slab_rank = -1, slab_size = 0;
  std::map<int, cudaGraphExec_t> slab_chunk_graphs;  // key: op * 1024 + iterations
  int slab_launches_per_chunk = 0;
  const double *slab_p_lo = nullptr, *slab_p_hi = nullptr;
  long long slab_lo_off = 0, slab_lo_npad = 0, slab_hi_off = 0, slab_hi_npad = 0;
  TileInfo2 tile2;        // tiling of k_spmv_dot_tmac (7 or 8 nodes per thread, two lane shapes)
  CUtensorMap tmap_a, tmap_b;  // V.p with the boxes of lane shape 0 (pitch x 10 x 6) and 1 (pitch x 6 x 10)
  int tile2_smem = 0;     // bytes of one brick
  int4 *d_tiles2 = nullptr;
  int *d_chunk_pure2 = nullptr, *d_fix_ptr2 = nullptr;
  int4 *d_fix2 = nullptr;
  int *d_elem_type = nullptr;
  double *d_ke = nullptr;
  double *d_be = nullptr;    // element residual scratch of assembly_rhs: [be_chunk][24][nelem_pad]
  int be_chunk = 0;
  double *d_ctan = nullptr;  // tangent scratch of the general Jacobian assembly: [ctan_chunk][288][nelem_pad]
  int ctan_chunk = 0;
  VecPool V{};
  SlotTables T{};
  int *d_list[NLIST] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  int *d_count = nullptr;
  int *d_cnt2 = nullptr;           // device-side list lengths used inside graphs: [0] Newton list, [1] CG list
  const int *dyn_count = nullptr;  // non-null while a graph is being captured: launches test it per block
  int *h_count = nullptr;  // pinned
  // persistent per-GP state
  double *d_ustore = nullptr;  // [ngp][2][3*nn_pad]
  std::vector<double *> u_n, u_k, vars_n, vars_k;
  std::vector<double *> var_chunks;
  std::vector<double *> var_free;
  size_t var_len = 0;  // doubles per vars buffer
  // slot tables (host mirrors)
  std::vector<int> slot_gp;
  std::vector<const double *> h_vars_old;
  std::vector<double *> h_vars_new, h_un, h_uk;
  // staging buffers for the host-pointer API
  std::vector<double *> stage_vars[2];
  // measurement
  bool prof = false;
  struct EvPair {
    cudaEvent_t a, b;
    int kind;
    int slots;
  };
  std::vector<EvPair> ev_live, ev_pool;
  double prof_acc[6] = {0, 0, 0, 0, 0, 0};
  cudaEvent_t t0 = nullptr, t1 = nullptr;
  unsigned long long launches = 0;
  struct StepGraph {
    cudaGraphExec_t exec;
    int fixed_launches, body_launches;
  };
  std::map<long long, StepGraph> step_graphs;  // key = bucket * 4 + operator
};

namespace {

inline Lst lst_of(const mgpu_ctx *c, int l, int off = 0) { return Lst{c->d_list[l], c->dyn_count, off}; }
inline dim3 int_grid(const mgpu_ctx *c, int n) { return dim3(std::max((c->mc.nint + NT - 1) / NT, 1), n, 1); }
inline dim3 node_grid(const mgpu_ctx *c, int n) { return dim3((c->mc.nn + NT - 1) / NT, n, 1); }
inline dim3 upd_grid(const mgpu_ctx *c, int n) { return dim3((c->mc.nn + NT * UPD_VPT - 1) / (NT * UPD_VPT), n, 1); }
inline dim3 elem_grid(const mgpu_ctx *c, int n) { return dim3((c->mc.nelem + NT - 1) / NT, n, 1); }

typedef void (*tile_kernel_t)(const MeshConst, const Lst, SlotTables, VecPool, TileInfo, int);
inline tile_kernel_t tile_kernel(int cb) {
  switch (cb) {
    case 1: return k_spmv_dot_tile<1>;
    case 2: return k_spmv_dot_tile<2>;
    case 3: return k_spmv_dot_tile<3>;
    case 4: return k_spmv_dot_tile<4>;
    case 5: return k_spmv_dot_tile<5>;
    case 6: return k_spmv_dot_tile<6>;
    case 7: return k_spmv_dot_tile<7>;
    default: return k_spmv_dot_tile<8>;
  }
}

typedef void (*tma_kernel_t)(const MeshConst, const Lst, int, SlotTables, VecPool, TileInfo, const CUtensorMap, int, int,
                             int);
inline tma_kernel_t tma_kernel(int cb) {
  switch (cb) {
    case 1: return k_spmv_dot_tma<1>;
    case 2: return k_spmv_dot_tma<2>;
    case 3: return k_spmv_dot_tma<3>;
    case 4: return k_spmv_dot_tma<4>;
    case 5: return k_spmv_dot_tma<5>;
    case 6: return k_spmv_dot_tma<6>;
    case 7: return k_spmv_dot_tma<7>;
    default: return k_spmv_dot_tma<8>;
  }
}

// variants of k_spmv_dot_tmac: {stages, resident blocks per SM, row unroll}; variant ids are 1-based
struct TmacVariant {
  int nstage, minb, ru;
};
constexpr int N_TMAC = 4;
static const TmacVariant kTmac[N_TMAC] = {{2, 2, 1}, {1, 3, 1}, {1, 4, 1}, {2, 2, 3}};
constexpr int TMAC_DEFAULT = 3;
typedef void (*tmac_kernel_t)(const MeshConst, const Lst, int, SlotTables, VecPool, TileInfo2, const CUtensorMap,
                              const CUtensorMap, const PureRows, int, int, int);
template <int TN, int NS, int MB, int RU>
inline tmac_kernel_t tmac_pick(int cb) {
  switch (cb) {
    case 1: return k_spmv_dot_tmac<1, TN, NS, MB, RU>;
    case 2: return k_spmv_dot_tmac<2, TN, NS, MB, RU>;
    case 3: return k_spmv_dot_tmac<3, TN, NS, MB, RU>;
    default: return k_spmv_dot_tmac<4, TN, NS, MB, RU>;
  }
}
template <int TN>
inline tmac_kernel_t tmac_kernel_tn(int variant, int cb) {
  switch (variant) {
    case 1: return tmac_pick<TN, 2, 2, 1>(cb);
    case 2: return tmac_pick<TN, 1, 3, 1>(cb);
    case 3: return tmac_pick<TN, 1, 4, 1>(cb);
    default: return tmac_pick<TN, 2, 2, 3>(cb);
  }
}
inline tmac_kernel_t tmac_kernel(int variant, int cb, int tn) {  // variant 1..N_TMAC
  return tn == 7 ? tmac_kernel_tn<7>(variant, cb) : tmac_kernel_tn<8>(variant, cb);
}

// Host -> device copy ORDERED WITH THE CONTEXT STREAM.  A plain cudaMemcpy from pageable memory may return before its
// DMA has landed, and the legacy default stream it runs on is not ordered with the non-blocking context stream the
// kernels use: a kernel launched right after it could read the old contents (seen as a rare wrong first operator
// application in the tests).  The copy is enqueued on the context stream and waited for (the source may be a temporary).
inline void h2d_sync(mgpu_ctx *c, void *dst, const void *src, size_t bytes) {
  CK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, c->stream));
  CK(cudaStreamSynchronize(c->stream));
}

struct ProfScope {
  mgpu_ctx *c;
  int kind, slots;
  bool on;
  mgpu_ctx::EvPair ev;
  ProfScope(mgpu_ctx *c_, int kind_, int slots_) : c(c_), kind(kind_), slots(slots_), on(c_->prof) {
    c->launches++;
    if (!on) return;
    if (c->ev_pool.empty()) {
      CK(cudaEventCreate(&ev.a));
      CK(cudaEventCreate(&ev.b));
    } else {
      ev = c->ev_pool.back();
      c->ev_pool.pop_back();
    }
    ev.kind = kind;
    ev.slots = slots;
    CK(cudaEventRecord(ev.a, c->stream));
  }
  ~ProfScope() {
    if (!on) return;
    CK(cudaEventRecord(ev.b, c->stream));
    c->ev_live.push_back(ev);
  }
};

void prof_drain(mgpu_ctx *c) {
  if (c->ev_live.empty()) return;
  CK(cudaStreamSynchronize(c->stream));
  for (auto &e : c->ev_live) {
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, e.a, e.b));
    switch (e.kind) {
      case 0:
        c->prof_acc[0] += ms;
        c->prof_acc[1] += 1;
        c->prof_acc[2] += e.slots;
        break;
      case 1: c->prof_acc[3] += ms; break;
      case 2: c->prof_acc[4] += ms; break;
      case 3: c->prof_acc[5] += ms; break;
      default: break;
    }
    c->ev_pool.push_back(e);
  }
  c->ev_live.clear();
}

void upload_slot_tables(mgpu_ctx *c) {
  const size_t W = c->W;
  CK(cudaMemcpyAsync((void *)c->T.vars_old, c->h_vars_old.data(), W * sizeof(double *), cudaMemcpyHostToDevice,
                     c->stream));
  CK(cudaMemcpyAsync((void *)c->T.vars_new, c->h_vars_new.data(), W * sizeof(double *), cudaMemcpyHostToDevice,
                     c->stream));
  CK(cudaMemcpyAsync((void *)c->T.u_n, c->h_un.data(), W * sizeof(double *), cudaMemcpyHostToDevice, c->stream));
  CK(cudaMemcpyAsync((void *)c->T.u_k, c->h_uk.data(), W * sizeof(double *), cudaMemcpyHostToDevice, c->stream));
}

double *vars_buffer_alloc(mgpu_ctx *c) {
  if (c->var_free.empty()) {
    // grow the pool by a chunk of buffers (never returned to the driver before mgpu_destroy)
    const size_t per = c->var_len;
    size_t nbuf = std::max<size_t>(2, std::min<size_t>(64, (size_t(1) << 30) / (per * sizeof(double) + 1)));
    double *chunk = nullptr;
    CK(cudaMalloc(&chunk, nbuf * per * sizeof(double)));
    c->var_chunks.push_back(chunk);
    for (size_t i = 0; i < nbuf; ++i) c->var_free.push_back(chunk + i * per);
  }
  double *p = c->var_free.back();
  c->var_free.pop_back();
  CK(cudaMemsetAsync(p, 0, c->var_len * sizeof(double), c->stream));
  return p;
}

// reference AoS [e][gp][7] (include/params.hpp:41) <-> internal [(v*8+gp)*nelem_pad + e]
void vars_ref_to_internal(const mgpu_ctx *c, const double *ref, std::vector<double> &out) {
  const int nelem = c->mc.nelem, npad = c->mc.nelem_pad, nv = c->mc.nvar;
  out.assign((size_t)nv * 8 * npad, 0.0);
  for (int e = 0; e < nelem; ++e)
    for (int gp = 0; gp < 8; ++gp)
      for (int v = 0; v < nv; ++v) out[(size_t)(v * 8 + gp) * npad + e] = ref[(size_t)e * 56 + gp * 7 + v];
}
void vars_internal_to_ref(const mgpu_ctx *c, const std::vector<double> &in, double *ref) {
  const int nelem = c->mc.nelem, npad = c->mc.nelem_pad, nv = c->mc.nvar;
  memset(ref, 0, sizeof(double) * (size_t)nelem * 56);
  for (int e = 0; e < nelem; ++e)
    for (int gp = 0; gp < 8; ++gp)
      for (int v = 0; v < nv; ++v) ref[(size_t)e * 56 + gp * 7 + v] = in[(size_t)(v * 8 + gp) * npad + e];
}
void aos_to_soa(const mgpu_ctx *c, const double *aos, std::vector<double> &soa) {
  const int nn = c->mc.nn, npad = c->mc.nn_pad;
  soa.assign((size_t)3 * npad, 0.0);
  for (int n = 0; n < nn; ++n)
    for (int d = 0; d < 3; ++d) soa[(size_t)d * npad + n] = aos[(size_t)n * 3 + d];
}
void soa_to_aos(const mgpu_ctx *c, const std::vector<double> &soa, double *aos) {
  const int nn = c->mc.nn, npad = c->mc.nn_pad;
  for (int n = 0; n < nn; ++n)
    for (int d = 0; d < 3; ++d) aos[(size_t)n * 3 + d] = soa[(size_t)d * npad + n];
}

double *vec_of(mgpu_ctx *c, int which) {
  switch (which) {
    case 0: return c->V.b;
    case 1: return c->V.du;
    case 2: return c->V.Ap;
    case 3: return c->V.p;
    case 4: return c->V.u;
    default: return c->V.r;
  }
}

// ---- pure host helpers of the implicit operator (no CUDA calls: also reachable from the CPU tests) ----
// Row-block id of every interior node: code = sum_c type_c 3^c over the 8 elements around the node (c as in
// k_asm_mat_elastic); ids 0..2 are reserved for nodes whose 8 elements are all of material 0 / 1 / 2.
static void implicit_row_ids(int nix, int niy, int niz, int nint_pad, const int *elem_type, std::vector<int> &codes,
                             std::vector<int> &rowid) {
  const int nex = nix + 1, ney = niy + 1;
  std::vector<int> code2id(6561, -1);
  codes.clear();
  rowid.assign(std::max(nint_pad, nix * niy * niz), 0);
  for (int t = 0; t < 3; ++t) {
    code2id[t * 3280] = t;
    codes.push_back(t * 3280);
  }
  const int nint = nix * niy * niz;
  for (int m = 0; m < nint; ++m) {
    const int pl = nix * niy;
    const int kk = m / pl, r = m - kk * pl, jj = r / nix, ii = r - jj * nix;
    const int i = ii + 1, j = jj + 1, k = kk + 1;
    int code = 0, w3 = 1;
    for (int cc = 0; cc < 8; ++cc) {
      const int ex = i - 1 + ((cc >> 2) & 1), ey = j - 1 + ((cc >> 1) & 1), ez = k - 1 + (cc & 1);
      code += w3 * elem_type[(ez * ney + ey) * nex + ex];
      w3 *= 3;
    }
    if (code2id[code] < 0) {
      code2id[code] = (int)codes.size();
      codes.push_back(code);
    }
    rowid[m] = code2id[code];
  }
}

// Tiling of k_spmv_dot_tmac: nodes per thread (7 or 8), warps per block, tile descriptors (two lane shapes), the pure
// row block + fix-up mask of every chunk, and the fix-up tasks of every tile.
struct TmacTiling {
  int tn = 8, cb = 1, nchunk = 0, pitch = 0;
  std::vector<int4> tiles, tasks;
  std::vector<int> chunk_pure, fix_ptr;
};
static TmacTiling tmac_tiling(int nix, int niy, int niz, const std::vector<int> &rowid) {
  TmacTiling t2;
  long best = -1;
  for (int tn = 8; tn >= 7; --tn)
    for (int cb = 4; cb >= 1; --cb) {
      const int nch = (nix + tn - 1) / tn;
      if (cb > nch) continue;
      const long exec = (long)((nch + cb - 1) / cb) * cb * tn;  // executed node slots per x row
      // fewer executed slots; blocks of 1 or 2 warps pay for their relatively larger halo and overheads
      const long score = exec * (cb >= 3 ? 100 : cb == 2 ? 115 : 140) + (4 - cb);
      if (best < 0 || score < best) {
        best = score;
        t2.tn = tn;
        t2.cb = cb;
      }
    }
  const int TN = t2.tn;
  t2.nchunk = (nix + TN - 1) / TN;
  t2.pitch = tmac_pitch(TN, t2.cb);
  // y is covered by 8-row tiles of lane shape 0; a remainder of 1..4 rows becomes a strip of shape-1 tiles
  const int yrem = niy % TILE_Y, y_a = (yrem >= 1 && yrem <= 4) ? niy - yrem : niy;
  std::vector<int4> &tiles = t2.tiles;
  const int tiles_x = (t2.nchunk + t2.cb - 1) / t2.cb;
  for (int z0 = 0; z0 < niz; z0 += TILE_Z)
    for (int y0 = 0; y0 < y_a; y0 += TILE_Y)
      for (int tx = 0; tx < tiles_x; ++tx) tiles.push_back(make_int4(tx * t2.cb, y0, z0, 0));
  if (y_a < niy)
    for (int z0 = 0; z0 < niz; z0 += TILE_Y)
      for (int tx = 0; tx < tiles_x; ++tx) tiles.push_back(make_int4(tx * t2.cb, y_a, z0, 1));
  const int ntiles = (int)tiles.size();
  std::vector<int> &chunk_pure = t2.chunk_pure, &fix_ptr = t2.fix_ptr;
  chunk_pure.assign((size_t)niz * niy * t2.nchunk, 0);
  fix_ptr.assign(ntiles + 1, 0);
  std::vector<int4> &tasks = t2.tasks;
  for (int tile = 0; tile < ntiles; ++tile) {
    const int4 td = tiles[tile];
    std::vector<int2> fix;
    fix_ptr[tile] = (int)tasks.size();
    const int ny_t = td.w ? TILE_Z : TILE_Y, nz_t = td.w ? TILE_Y : TILE_Z;
    for (int rz = 0; rz < nz_t; ++rz)
      for (int ry = 0; ry < ny_t; ++ry)
        for (int wq = 0; wq < t2.cb; ++wq) {
          const int kk = td.z + rz, jj = td.y + ry, cc = td.x + wq;
          if (kk >= niz || jj >= niy || cc >= t2.nchunk) continue;
          const int m0 = (kk * niy + jj) * nix + cc * TN, nv = std::min(TN, nix - cc * TN);
          int cnt[3] = {0, 0, 0};
          for (int t = 0; t < nv; ++t)
            if (rowid[m0 + t] < 3) cnt[rowid[m0 + t]]++;
          int pure = 0;
          for (int q = 1; q < 3; ++q)
            if (cnt[q] > cnt[pure]) pure = q;
          int mask = 0;
          for (int t = 0; t < nv; ++t)
            if (rowid[m0 + t] != pure) {
              mask |= 1 << t;
              int2 e;
              e.x = (wq * TN + t) | (ry << 8) | (rz << 12);
              e.y = rowid[m0 + t];
              fix.push_back(e);
            }
          chunk_pure[((size_t)kk * niy + jj) * t2.nchunk + cc] = pure | (mask << 8);
        }
    // lanes of a warp take consecutive entries: sorted by row-block id, neighbouring lanes fetch the same
    // row block (the loads of a fix-up round are bound by the distinct 32-B sectors a warp touches)
    std::stable_sort(fix.begin(), fix.end(), [](const int2 &a, const int2 &b) { return a.y < b.y; });
    // tasks: two nodes with the same row block share one task (one set of row-block loads)
    for (size_t q = 0; q < fix.size();) {
      if (q + 1 < fix.size() && fix[q + 1].y == fix[q].y) {
        tasks.push_back(make_int4(fix[q].x, fix[q + 1].x, fix[q].y, 0));
        q += 2;
      } else {
        tasks.push_back(make_int4(fix[q].x, -1, fix[q].y, 0));
        q += 1;
      }
    }
  }
  fix_ptr[ntiles] = (int)tasks.size();
  return t2;
}

}  // namespace

extern "C" {

// Host-only view of the implicit operator's tiling for an nx x ny x nz RVE (tests/test_tiling.py, no GPU needed).
// meta[6] = {nodes per thread, warps per block, chunks per x row, brick pitch, tiles, fix-up tasks}; the arrays may be
// null (size query): rowid[(nx-2)(ny-2)(nz-2)], tiles[4 * ntiles], chunk_pure[niz * niy * nchunk], fix_ptr[ntiles + 1],
// tasks[4 * ntasks].  Returns the number of distinct row blocks.
int mgpu_tmac_tiling_host(int nx, int ny, int nz, const int *elem_type, int *meta, int *rowid_out, int *tiles,
                          int *chunk_pure, int *fix_ptr, int *tasks) {
  const int nix = nx - 2, niy = ny - 2, niz = nz - 2;
  if (nix < 1 || niy < 1 || niz < 1) return 0;
  std::vector<int> codes, rowid;
  implicit_row_ids(nix, niy, niz, nix * niy * niz, elem_type, codes, rowid);
  const TmacTiling t = tmac_tiling(nix, niy, niz, rowid);
  if (meta) {
    meta[0] = t.tn;
    meta[1] = t.cb;
    meta[2] = t.nchunk;
    meta[3] = t.pitch;
    meta[4] = (int)t.tiles.size();
    meta[5] = (int)t.tasks.size();
  }
  if (rowid_out) memcpy(rowid_out, rowid.data(), sizeof(int) * (size_t)nix * niy * niz);
  if (tiles) memcpy(tiles, t.tiles.data(), sizeof(int4) * t.tiles.size());
  if (chunk_pure) memcpy(chunk_pure, t.chunk_pure.data(), sizeof(int) * t.chunk_pure.size());
  if (fix_ptr) memcpy(fix_ptr, t.fix_ptr.data(), sizeof(int) * t.fix_ptr.size());
  if (tasks && !t.tasks.empty()) memcpy(tasks, t.tasks.data(), sizeof(int4) * t.tasks.size());
  return (int)codes.size();
}

int mgpu_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
  return n;
}

mgpu_ctx *mgpu_create(const mgpu_config *cfg) {
  int ndev = 0;
  cudaError_t err = cudaGetDeviceCount(&ndev);
  if (err != cudaSuccess || ndev == 0) {
    // No CPU fallback by design: the hot path exists only as sm_100a kernels.
    fprintf(stderr, "micropp-b200: no CUDA device available (%s); this library has no CPU path\n",
            cudaGetErrorString(err));
    abort();
  }
  mgpu_ctx *c = new mgpu_ctx();
  c->device = cfg->device % ndev;
  CK(cudaSetDevice(c->device));
  CK(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
  CK(cudaEventCreate(&c->t0));
  CK(cudaEventCreate(&c->t1));

  MeshConst &P = c->mc;
  memset(&P, 0, sizeof(P));
  P.nx = cfg->nx;
  P.ny = cfg->ny;
  P.nz = cfg->nz;
  P.nxny = P.nx * P.ny;
  P.nn = P.nx * P.ny * P.nz;
  P.nn_pad = (P.nn + 31) / 32 * 32;
  P.nix = std::max(P.nx - 2, 0);
  P.niy = std::max(P.ny - 2, 0);
  P.niz = std::max(P.nz - 2, 0);
  P.nint = P.nix * P.niy * P.niz;
  P.nint_pad = std::max((P.nint + 31) / 32 * 32, 32);
  P.slab = cfg->slab;
  P.koff = cfg->slab ? cfg->koff : 0;
  P.nz_glob = cfg->slab ? cfg->nz_glob : P.nz;
  P.halo_lo = cfg->slab ? cfg->halo_lo : 0;
  P.halo_hi = cfg->slab ? cfg->halo_hi : 0;
  P.ez_own_lo = cfg->slab ? cfg->ez_own_lo : 0;
  P.ez_own_hi = cfg->slab ? cfg->ez_own_hi : std::max(P.nz - 1, 0);
  P.nex = P.nx - 1;
  P.ney = P.ny - 1;
  P.nez = P.nz - 1;
  P.nelem = P.nex * P.ney * P.nez;
  P.nelem_pad = (P.nelem + 31) / 32 * 32;
  P.nr_max_its = cfg->nr_max_its;
  P.cg_max_its = cfg->cg_max_its;
  P.dx = cfg->dx;
  P.dy = cfg->dy;
  P.dz = cfg->dz;
  P.wg = cfg->wg;
  P.nr_max_tol = cfg->nr_max_tol;
  P.nr_rel_tol = cfg->nr_rel_tol;
  P.cg_abs_tol = cfg->cg_abs_tol;
  P.cg_rel_tol = cfg->cg_rel_tol;
  memcpy(P.dsh, cfg->dsh, sizeof(P.dsh));
  int nvar = 0;
  c->all_elastic = true;
  for (int i = 0; i < 3; ++i) {
    mpp_material &m = P.mat[i];
    m.E = cfg->mat[i][0];
    m.nu = cfg->mat[i][1];
    m.Ka = cfg->mat[i][2];
    m.Sy = cfg->mat[i][3];
    m.k = cfg->mat[i][4];
    m.mu = cfg->mat[i][5];
    m.lambda = cfg->mat[i][6];
    m.Xt = cfg->mat[i][7];
    m.type = cfg->mat_type[i];
  }
  // only materials that actually occur in the micro-structure count
  bool used[3] = {false, false, false};
  for (int e = 0; e < P.nelem; ++e) {
    const int t = cfg->elem_type[e];
    if (t < 0 || t > 2) {
      fprintf(stderr, "micropp-b200: invalid element type %d\n", t);
      abort();
    }
    used[t] = true;
  }
  for (int i = 0; i < 3; ++i)
    if (used[i]) {
      nvar = std::max(nvar, mat_nvar(P.mat[i].type));
      if (P.mat[i].type != MPP_ELASTIC) c->all_elastic = false;
    }
  P.nvar = nvar;
  c->var_len = (size_t)std::max(nvar, 1) * 8 * P.nelem_pad;
  c->ngp = cfg->ngp;

  CK(cudaMalloc(&c->d_elem_type, sizeof(int) * std::max(P.nelem, 1)));
  h2d_sync(c, c->d_elem_type, cfg->elem_type, sizeof(int) * P.nelem);
  CK(cudaMalloc(&c->d_ke, sizeof(double) * 3 * 576));
  h2d_sync(c, c->d_ke, cfg->ke_elastic, sizeof(double) * 3 * 576);

  // persistent displacement state u_n,u_k of every FE Gauss point
  const size_t vlen = (size_t)3 * P.nn_pad;
  const int ngp = std::max(cfg->ngp, 0);
  if (ngp > 0) {
    CK(cudaMalloc(&c->d_ustore, sizeof(double) * vlen * 2 * ngp));
    CK(cudaMemset(c->d_ustore, 0, sizeof(double) * vlen * 2 * ngp));
  }
  c->u_n.resize(ngp);
  c->u_k.resize(ngp);
  c->vars_n.assign(ngp, nullptr);
  c->vars_k.assign(ngp, nullptr);
  for (int g = 0; g < ngp; ++g) {
    c->u_n[g] = c->d_ustore + (size_t)g * 2 * vlen;
    c->u_k[g] = c->u_n[g] + vlen;
  }

  // wave size from the HBM left after reserving room for internal variables of every GP
  const size_t mlen = (size_t)NPLANE * P.nint_pad;
  const int nblk_max = std::max((P.nn + NT - 1) / NT, (P.nelem + NT - 1) / NT);
  c->implicit = c->all_elastic && cfg->implicit_elastic && P.nint > 0;
  const size_t per_slot =
      sizeof(double) * ((c->implicit ? 0 : mlen) + 8 * vlen + (size_t)NRED * nblk_max + 12) + 256;
  size_t free_b = 0, total_b = 0;
  CK(cudaMemGetInfo(&free_b, &total_b));
  size_t reserve = sizeof(double) * mlen * (c->implicit ? 3 : 1) /* A0 (+ 2 explicit slots) */ + (size_t(1) << 30);
  const size_t ctan_per_slot = sizeof(double) * CTAN_LEN * (size_t)P.nelem_pad;
  const size_t be_per_slot = sizeof(double) * 24 * (size_t)P.nelem_pad;
  {
    long long ch = (long long)((total_b / 100) / std::max<size_t>(be_per_slot, 1));  // about 1% of the device
    c->be_chunk = (int)std::min<long long>(256, std::max<long long>(4, ch));
    reserve += be_per_slot * c->be_chunk;
  }
  if (!c->all_elastic) {
    // tangent scratch for a chunk of slots: about 4% of the device, 4..64 slots
    long long ch = (long long)((total_b / 25) / std::max<size_t>(ctan_per_slot, 1));
    c->ctan_chunk = (int)std::min<long long>(64, std::max<long long>(4, ch));
    reserve += ctan_per_slot * c->ctan_chunk;
  }
  if (!c->all_elastic) reserve += (size_t)ngp * 2 * c->var_len * sizeof(double);
  size_t avail = free_b > reserve ? free_b - reserve : 0;
  avail = (size_t)(avail * 0.92);
  long long W = (long long)(avail / per_slot);
  const int want = std::max(ngp, 6);  // at least the 6 unit-strain solves of calc_ctan_lin_fe_models
  if (W > want) W = want;
  if (cfg->wave_cap > 0 && W > cfg->wave_cap) W = cfg->wave_cap;
  if (const char *env = getenv("MICROPP_WAVE")) {
    const long long w = atoll(env);
    if (w > 0 && W > w) W = w;
  }
  if (W < 1) {
    fprintf(stderr, "micropp-b200: not enough device memory for one %dx%dx%d RVE\n", P.nx, P.ny, P.nz);
    abort();
  }
  c->W = (int)W;

  VecPool &V = c->V;
  V.vstride = vlen;
  V.mstride = mlen;
  double **vecs[8] = {&V.u, &V.b, &V.du, &V.k, &V.r, &V.z, &V.p, &V.Ap};
  for (auto pp : vecs) {
    CK(cudaMalloc(pp, sizeof(double) * vlen * W));
    CK(cudaMemset(*pp, 0, sizeof(double) * vlen * W));
  }
  // explicit matrices: one per slot, or just two (host-pointer API, debugging) next to the implicit operator
  c->mat_slots = c->implicit ? (int)std::min<long long>(W, 2) : (int)W;
  CK(cudaMalloc(&V.mat, sizeof(double) * mlen * c->mat_slots));
  c->be_chunk = std::min(c->be_chunk, (int)W);
  CK(cudaMalloc(&c->d_be, be_per_slot * c->be_chunk));
  if (!c->all_elastic) {
    c->ctan_chunk = std::min(c->ctan_chunk, (int)W);
    CK(cudaMalloc(&c->d_ctan, ctan_per_slot * c->ctan_chunk));
  }
  CK(cudaMemsetAsync(V.mat, 0, sizeof(double) * mlen * c->mat_slots, c->stream));  // tile padding stays defined
  V.mat_shared = nullptr;  // allocated on first use (use_A0)
  V.gen = nullptr;         // allocated on first use (generic host-matrix API)

  if (c->implicit) {
    // distinct row blocks: code = sum_c type_c 3^c over the 8 elements around an interior node
    std::vector<int> codes, rowid;
    implicit_row_ids(P.nix, P.niy, P.niz, P.nint_pad, cfg->elem_type, codes, rowid);
    c->nrows = (int)codes.size();
    int *d_codes = nullptr, *d_rowid = nullptr;
    double *d_rows = nullptr, *d_rkinv = nullptr;
    CK(cudaMalloc(&d_codes, sizeof(int) * c->nrows));
    CK(cudaMalloc(&d_rowid, sizeof(int) * P.nint_pad));
    CK(cudaMalloc(&d_rows, sizeof(double) * RB_LEN * c->nrows));
    CK(cudaMalloc(&d_rkinv, sizeof(double) * 3 * c->nrows));
    h2d_sync(c, d_codes, codes.data(), sizeof(int) * c->nrows);
    h2d_sync(c, d_rowid, rowid.data(), sizeof(int) * P.nint_pad);
    k_rows_build<<<(c->nrows + NT - 1) / NT, NT, 0, c->stream>>>(d_codes, c->nrows, d_rows, d_rkinv, c->d_ke);
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(c->stream));
    CK(cudaFree(d_codes));
    c->launches++;
    V.rows = d_rows;
    V.rkinv = d_rkinv;
    V.rowid = d_rowid;
    memset(&c->pure_rows, 0, sizeof(PureRows));
    CK(cudaMemcpy(c->pure_rows.a, d_rows, sizeof(double) * RB_LEN * std::min(c->nrows, 3), cudaMemcpyDeviceToHost));

    // tiling of k_spmv_dot_tile: chunks of 8 x-adjacent interior nodes, cb chunks (warps) per block
    TileInfo &ti = c->tile;
    ti.nchunk = (P.nix + 7) / 8;
    // blocks of 4 warps (2 resident blocks of 2 x 49 KB stages per SM) unless that idles too many warps
    int best_cb = 1;
    double best_score = 1e30;
    for (int cb = 1; cb <= std::min(ti.nchunk, 4); ++cb) {
      const int warps = (ti.nchunk + cb - 1) / cb * cb;  // executed warps per row block, idle ones included
      const double score = warps * (cb == 4 ? 1.0 : cb == 3 ? 1.15 : 1.5);
      if (score < best_score) {
        best_score = score;
        best_cb = cb;
      }
    }
    ti.cb = best_cb;
    ti.tiles_x = (ti.nchunk + ti.cb - 1) / ti.cb;
    ti.tiles_y = (P.niy + TILE_Y - 1) / TILE_Y;
    ti.tiles_z = (P.niz + TILE_Z - 1) / TILE_Z;
    ti.pitch = 8 * ti.cb + 2;
    c->tile_smem = (int)(sizeof(double) * 3 * BRICK_ROWS * ti.pitch);
    std::vector<int> chunk_id((size_t)P.niz * P.niy * ti.nchunk, -1);
    for (int kk = 0; kk < P.niz; ++kk)
      for (int jj = 0; jj < P.niy; ++jj)
        for (int cc = 0; cc < ti.nchunk; ++cc) {
          const int m0 = (kk * P.niy + jj) * P.nix + cc * 8, nv = std::min(8, P.nix - cc * 8);
          int id = rowid[m0];
          for (int t = 1; t < nv; ++t)
            if (rowid[m0 + t] != id) id = -1;
          chunk_id[((size_t)kk * P.niy + jj) * ti.nchunk + cc] = id;
        }
    {
      const int ntiles = ti.tiles_x * ti.tiles_y * ti.tiles_z;
      std::vector<int> chunk_pure(chunk_id.size(), 0), fix_ptr(ntiles + 1, 0);
      std::vector<int2> fix;
      for (int tz = 0; tz < ti.tiles_z; ++tz)
        for (int ty = 0; ty < ti.tiles_y; ++ty)
          for (int tx = 0; tx < ti.tiles_x; ++tx) {
            const int tile = (tz * ti.tiles_y + ty) * ti.tiles_x + tx;
            fix_ptr[tile] = (int)fix.size();
            for (int rz = 0; rz < TILE_Z; ++rz)
              for (int ry = 0; ry < TILE_Y; ++ry)
                for (int wq = 0; wq < ti.cb; ++wq) {
                  const int kk = tz * TILE_Z + rz, jj = ty * TILE_Y + ry, cc = tx * ti.cb + wq;
                  if (kk >= P.niz || jj >= P.niy || cc >= ti.nchunk) continue;
                  const int m0 = (kk * P.niy + jj) * P.nix + cc * 8, nv = std::min(8, P.nix - cc * 8);
                  int cnt[3] = {0, 0, 0};
                  for (int t = 0; t < nv; ++t)
                    if (rowid[m0 + t] < 3) cnt[rowid[m0 + t]]++;
                  int pure = 0;
                  for (int q = 1; q < 3; ++q)
                    if (cnt[q] > cnt[pure]) pure = q;
                  int mask = 0;
                  for (int t = 0; t < nv; ++t)
                    if (rowid[m0 + t] != pure) {
                      mask |= 1 << t;
                      int2 e;
                      e.x = (wq * 8 + t) | (ry << 8) | (rz << 12);
                      e.y = rowid[m0 + t];
                      fix.push_back(e);
                    }
                  chunk_pure[((size_t)kk * P.niy + jj) * ti.nchunk + cc] = pure | (mask << 8);
                }
          }
      fix_ptr[ntiles] = (int)fix.size();
      c->nfix = (int)fix.size();
      CK(cudaMalloc(&c->d_chunk_pure, sizeof(int) * chunk_pure.size()));
      h2d_sync(c, c->d_chunk_pure, chunk_pure.data(), sizeof(int) * chunk_pure.size());
      CK(cudaMalloc(&c->d_fix_ptr, sizeof(int) * fix_ptr.size()));
      h2d_sync(c, c->d_fix_ptr, fix_ptr.data(), sizeof(int) * fix_ptr.size());
      CK(cudaMalloc(&c->d_fix, sizeof(int2) * std::max<size_t>(fix.size(), 1)));
      if (!fix.empty()) h2d_sync(c, c->d_fix, fix.data(), sizeof(int2) * fix.size());
      ti.chunk_pure = c->d_chunk_pure;
      ti.fix_ptr = c->d_fix_ptr;
      ti.fix = c->d_fix;
    }
    CK(cudaMalloc(&c->d_chunk_id, sizeof(int) * chunk_id.size()));
    h2d_sync(c, c->d_chunk_id, chunk_id.data(), sizeof(int) * chunk_id.size());
    ti.chunk_id = c->d_chunk_id;
    CK(cudaFuncSetAttribute(tile_kernel(ti.cb), cudaFuncAttributeMaxDynamicSharedMemorySize, c->tile_smem));
    c->imp_kernel = 1;
    if (P.nx % 2 == 0) {
      // TMA descriptor of the p pool; the encoder comes from the driver through the runtime (no -lcuda)
      typedef CUresult (*encode_t)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                   const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
      void *fn = nullptr;
      cudaDriverEntryPointQueryResult qres;
      if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) == cudaSuccess && fn &&
          qres == cudaDriverEntryPointSuccess) {
        const cuuint64_t gdim[5] = {(cuuint64_t)P.nx, (cuuint64_t)P.ny, (cuuint64_t)P.nz, 3, (cuuint64_t)W};
        const cuuint64_t gstr[4] = {(cuuint64_t)P.nx * 8, (cuuint64_t)P.nxny * 8, (cuuint64_t)P.nn_pad * 8,
                                    (cuuint64_t)vlen * 8};
        const cuuint32_t box[5] = {(cuuint32_t)ti.pitch, TILE_Y + 2, TILE_Z + 2, 3, 1};
        const cuuint32_t estr[5] = {1, 1, 1, 1, 1};
        const CUresult r = ((encode_t)fn)(&c->tmap_p, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 5, (void *)V.p, gdim, gstr, box,
                                          estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                          CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r == CUDA_SUCCESS) {
          c->imp_kernel = 2;
          c->tma_smem = 2 * ((c->tile_smem + 127) / 128 * 128) + 128;
          CK(cudaFuncSetAttribute(tma_kernel(ti.cb), cudaFuncAttributeMaxDynamicSharedMemorySize, c->tma_smem));
          // ---- tiling of k_spmv_dot_tmac: TN nodes per thread, cb warps per block, tile descriptors ----
          {
            TileInfo2 &t2 = c->tile2;
            TmacTiling tt = tmac_tiling(P.nix, P.niy, P.niz, rowid);
            t2.tn = tt.tn;
            t2.cb = tt.cb;
            t2.nchunk = tt.nchunk;
            t2.pitch = tt.pitch;
            t2.ntiles = (int)tt.tiles.size();
            const int TN = t2.tn;
            c->tile2_smem = (int)(sizeof(double) * 3 * BRICK_ROWS * t2.pitch);
            const std::vector<int4> &tiles = tt.tiles, &tasks = tt.tasks;
            const std::vector<int> &chunk_pure = tt.chunk_pure, &fix_ptr = tt.fix_ptr;
            CK(cudaMalloc(&c->d_tiles2, sizeof(int4) * tiles.size()));
            h2d_sync(c, c->d_tiles2, tiles.data(), sizeof(int4) * tiles.size());
            CK(cudaMalloc(&c->d_chunk_pure2, sizeof(int) * chunk_pure.size()));
            h2d_sync(c, c->d_chunk_pure2, chunk_pure.data(), sizeof(int) * chunk_pure.size());
            CK(cudaMalloc(&c->d_fix_ptr2, sizeof(int) * fix_ptr.size()));
            h2d_sync(c, c->d_fix_ptr2, fix_ptr.data(), sizeof(int) * fix_ptr.size());
            CK(cudaMalloc(&c->d_fix2, sizeof(int4) * std::max<size_t>(tasks.size(), 1)));
            if (!tasks.empty())
              h2d_sync(c, c->d_fix2, tasks.data(), sizeof(int4) * tasks.size());
            t2.tiles = c->d_tiles2;
            t2.chunk_pure = c->d_chunk_pure2;
            t2.fix_ptr = c->d_fix_ptr2;
            t2.fix = c->d_fix2;
            const cuuint32_t box_a[5] = {(cuuint32_t)t2.pitch, TILE_Y + 2, TILE_Z + 2, 3, 1};
            const cuuint32_t box_b[5] = {(cuuint32_t)t2.pitch, TILE_Z + 2, TILE_Y + 2, 3, 1};
            const CUresult ra = ((encode_t)fn)(&c->tmap_a, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 5, (void *)V.p, gdim, gstr,
                                               box_a, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                               CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            const CUresult rb = ((encode_t)fn)(&c->tmap_b, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 5, (void *)V.p, gdim, gstr,
                                               box_b, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                               CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (ra == CUDA_SUCCESS && rb == CUDA_SUCCESS && t2.ntiles * t2.cb <= nblk_max) {
              for (int v = 1; v <= N_TMAC; ++v)
                CK(cudaFuncSetAttribute(tmac_kernel(v, t2.cb, TN), cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        kTmac[v - 1].nstage * ((c->tile2_smem + 127) / 128 * 128) + 128));
              c->tma_variant = TMAC_DEFAULT;
              if (const char *env = getenv("MICROPP_TMA_VARIANT"))
                c->tma_variant = std::min(std::max(atoi(env), 0), N_TMAC);
            } else {
              t2.ntiles = 0;  // k_spmv_dot_tmac unavailable: k_spmv_dot_tma serves kernel 2
            }
          }
        } else {
          fprintf(stderr, "micropp-b200: cuTensorMapEncodeTiled failed (%d); using the cp.async tiled kernel\n", (int)r);
        }
      }
    }
    if (const char *env = getenv("MICROPP_IMP_KERNEL")) c->imp_kernel = std::min(c->imp_kernel, std::max(atoi(env), 0));
    if (ti.tiles_x * ti.tiles_y * ti.tiles_z > nblk_max) c->imp_kernel = 0;  // partial-sum buffer too small
  }

  SlotTables &T = c->T;
  T.nblk_max = nblk_max;
  CK(cudaMalloc(&T.state, sizeof(mgpu_slot_state) * W));
  CK(cudaMemset(T.state, 0, sizeof(mgpu_slot_state) * W));
  CK(cudaMalloc((void **)&T.vars_old, sizeof(double *) * W));
  CK(cudaMalloc((void **)&T.vars_new, sizeof(double *) * W));
  CK(cudaMalloc((void **)&T.u_n, sizeof(double *) * W));
  CK(cudaMalloc((void **)&T.u_k, sizeof(double *) * W));
  CK(cudaMalloc(&T.eps, sizeof(double) * 6 * W));
  CK(cudaMalloc(&T.stress, sizeof(double) * 6 * W));
  CK(cudaMemset(T.eps, 0, sizeof(double) * 6 * W));
  CK(cudaMemset(T.stress, 0, sizeof(double) * 6 * W));
  CK(cudaMalloc(&T.partial, sizeof(double) * (size_t)NRED * nblk_max * W));
  CK(cudaMalloc(&T.red, sizeof(double) * 8 * W));
  CK(cudaMemset(T.red, 0, sizeof(double) * 8 * W));
  for (int l = 0; l < NLIST; ++l) CK(cudaMalloc(&c->d_list[l], sizeof(int) * W));
  CK(cudaMalloc(&c->d_count, sizeof(int)));
  CK(cudaMallocHost(&c->h_count, 4 * sizeof(int)));
  CK(cudaMalloc(&c->d_cnt2, 4 * sizeof(int)));
  CK(cudaMemset(c->d_cnt2, 0, 4 * sizeof(int)));

  c->slot_gp.assign(W, -1);
  c->h_vars_old.assign(W, nullptr);
  c->h_vars_new.assign(W, nullptr);
  c->h_un.assign(W, nullptr);
  c->h_uk.assign(W, nullptr);
  upload_slot_tables(c);

  CK(cudaFuncSetAttribute(k_asm_mat_general, cudaFuncAttributeMaxDynamicSharedMemorySize,
                          (int)(GN * NPLANE * sizeof(double))));
  CK(cudaStreamSynchronize(c->stream));
  CK(cudaDeviceSynchronize());  // the cudaMemset calls above ran on the legacy stream, which the context stream ignores
  return c;
}

void mgpu_destroy(mgpu_ctx *c) {
  if (!c) return;
  CK(cudaSetDevice(c->device));
  cudaStreamSynchronize(c->stream);
  prof_drain(c);
  for (auto &e : c->ev_pool) {
    cudaEventDestroy(e.a);
    cudaEventDestroy(e.b);
  }
  double *vecs[9] = {c->V.u, c->V.b, c->V.du, c->V.k, c->V.r, c->V.z, c->V.p, c->V.Ap, c->V.mat};
  for (auto p : vecs) cudaFree(p);
  if (c->V.mat_shared) cudaFree(c->V.mat_shared);
  if (c->V.gen) cudaFree(c->V.gen);
  if (c->V.rows) cudaFree((void *)c->V.rows);
  if (c->V.rkinv) cudaFree((void *)c->V.rkinv);
  if (c->V.rowid) cudaFree((void *)c->V.rowid);
  if (c->d_chunk_id) cudaFree(c->d_chunk_id);
  if (c->d_chunk_pure) cudaFree(c->d_chunk_pure);
  if (c->d_fix_ptr) cudaFree(c->d_fix_ptr);
  for (auto &kv : c->slab_chunk_graphs) cudaGraphExecDestroy(kv.second);
  if (c->slab_mail) cudaFree(c->slab_mail);
  if (c->d_tiles2) cudaFree(c->d_tiles2);
  if (c->d_chunk_pure2) cudaFree(c->d_chunk_pure2);
  if (c->d_fix_ptr2) cudaFree(c->d_fix_ptr2);
  if (c->d_fix2) cudaFree(c->d_fix2);
  if (c->d_fix) cudaFree(c->d_fix);
  cudaFree(c->T.state);
  cudaFree((void *)c->T.vars_old);
  cudaFree((void *)c->T.vars_new);
  cudaFree((void *)c->T.u_n);
  cudaFree((void *)c->T.u_k);
  cudaFree(c->T.eps);
  cudaFree(c->T.stress);
  cudaFree(c->T.partial);
  cudaFree(c->T.red);
  for (int l = 0; l < NLIST; ++l) cudaFree(c->d_list[l]);
  cudaFree(c->d_count);
  cudaFreeHost(c->h_count);
  cudaFree(c->d_cnt2);
  for (auto &kv : c->step_graphs) cudaGraphExecDestroy(kv.second.exec);
  cudaFree(c->d_elem_type);
  cudaFree(c->d_ke);
  if (c->d_ctan) cudaFree(c->d_ctan);
  if (c->d_be) cudaFree(c->d_be);
  if (c->d_ustore) cudaFree(c->d_ustore);
  for (auto p : c->var_chunks) cudaFree(p);
  for (int w = 0; w < 2; ++w)
    for (auto p : c->stage_vars[w])
      if (p) cudaFree(p);
  cudaEventDestroy(c->t0);
  cudaEventDestroy(c->t1);
  cudaStreamDestroy(c->stream);
  delete c;
}

int mgpu_wave_size(const mgpu_ctx *c) { return c->W; }
int mgpu_nn_pad(const mgpu_ctx *c) { return c->mc.nn_pad; }
int mgpu_nelem_pad(const mgpu_ctx *c) { return c->mc.nelem_pad; }
int mgpu_nvar(const mgpu_ctx *c) { return c->mc.nvar; }
void mgpu_sync(mgpu_ctx *c) {
  CK(cudaSetDevice(c->device));
  CK(cudaStreamSynchronize(c->stream));
}
unsigned long long mgpu_launch_count(const mgpu_ctx *c) { return c->launches; }

// ---- per-GP persistent state ---------------------------------------------------------------
void mgpu_gp_swap(mgpu_ctx *c, int gp) {
  std::swap(c->u_n[gp], c->u_k[gp]);
  std::swap(c->vars_n[gp], c->vars_k[gp]);
}
int mgpu_gp_has_vars(const mgpu_ctx *c, int gp) { return c->vars_n[gp] != nullptr; }
void mgpu_gp_alloc_vars(mgpu_ctx *c, int gp) {
  CK(cudaSetDevice(c->device));
  if (c->vars_n[gp]) return;
  c->vars_n[gp] = vars_buffer_alloc(c);
  c->vars_k[gp] = vars_buffer_alloc(c);
}
void mgpu_gp_free_vars(mgpu_ctx *c, int gp) {
  if (!c->vars_n[gp]) return;
  CK(cudaSetDevice(c->device));
  CK(cudaStreamSynchronize(c->stream));
  c->var_free.push_back(c->vars_n[gp]);
  c->var_free.push_back(c->vars_k[gp]);
  c->vars_n[gp] = c->vars_k[gp] = nullptr;
}
void mgpu_gp_get_u(mgpu_ctx *c, int gp, int which, double *host_aos) {
  CK(cudaSetDevice(c->device));
  std::vector<double> tmp((size_t)3 * c->mc.nn_pad);
  CK(cudaStreamSynchronize(c->stream));
  CK(cudaMemcpy(tmp.data(), which ? c->u_k[gp] : c->u_n[gp], tmp.size() * sizeof(double), cudaMemcpyDeviceToHost));
  soa_to_aos(c, tmp, host_aos);
}
void mgpu_gp_set_u(mgpu_ctx *c, int gp, int which, const double *host_aos) {
  CK(cudaSetDevice(c->device));
  std::vector<double> tmp;
  aos_to_soa(c, host_aos, tmp);
  CK(cudaStreamSynchronize(c->stream));
  h2d_sync(c, which ? c->u_k[gp] : c->u_n[gp], tmp.data(), tmp.size() * sizeof(double));
}
void mgpu_gp_get_vars(mgpu_ctx *c, int gp, int which, double *ref) {
  CK(cudaSetDevice(c->device));
  const double *src = which ? c->vars_k[gp] : c->vars_n[gp];
  if (!src) {
    memset(ref, 0, sizeof(double) * (size_t)c->mc.nelem * 56);
    return;
  }
  std::vector<double> tmp(c->var_len);
  CK(cudaStreamSynchronize(c->stream));
  CK(cudaMemcpy(tmp.data(), src, tmp.size() * sizeof(double), cudaMemcpyDeviceToHost));
  vars_internal_to_ref(c, tmp, ref);
}
void mgpu_gp_set_vars(mgpu_ctx *c, int gp, int which, const double *ref) {
  CK(cudaSetDevice(c->device));
  mgpu_gp_alloc_vars(c, gp);
  std::vector<double> tmp;
  vars_ref_to_internal(c, ref, tmp);
  tmp.resize(c->var_len, 0.0);
  CK(cudaStreamSynchronize(c->stream));
  h2d_sync(c, which ? c->vars_k[gp] : c->vars_n[gp], tmp.data(), tmp.size() * sizeof(double));
}

// ---- wave set-up ----------------------------------------------------------------------------
void mgpu_bind_slots(mgpu_ctx *c, int n, const int *slots, const int *gps, const int *use_vars_old) {
  CK(cudaSetDevice(c->device));
  for (int i = 0; i < n; ++i) {
    const int s = slots[i], g = gps[i];
    c->slot_gp[s] = g;
    if (g >= 0) {
      c->h_un[s] = c->u_n[g];
      c->h_uk[s] = c->u_k[g];
      c->h_vars_old[s] = (use_vars_old && use_vars_old[i]) ? c->vars_n[g] : nullptr;
      c->h_vars_new[s] = c->vars_k[g];
    } else {
      c->h_un[s] = c->h_uk[s] = nullptr;
      c->h_vars_old[s] = nullptr;
      c->h_vars_new[s] = nullptr;
    }
  }
  CK(cudaStreamSynchronize(c->stream));  // the pageable host mirrors are re-used across calls
  upload_slot_tables(c);
  CK(cudaStreamSynchronize(c->stream));
}

void mgpu_set_slot_strain(mgpu_ctx *c, int n, const int *slots, const double *eps6) {
  CK(cudaSetDevice(c->device));
  CK(cudaStreamSynchronize(c->stream));
  // contiguous runs are the common case; fall back to per-slot copies otherwise
  bool contiguous = true;
  for (int i = 1; i < n; ++i) contiguous &= (slots[i] == slots[0] + i);
  if (contiguous && n > 0) {
    h2d_sync(c, c->T.eps + (size_t)slots[0] * 6, eps6, sizeof(double) * 6 * n);
  } else {
    for (int i = 0; i < n; ++i)
      h2d_sync(c, c->T.eps + (size_t)slots[i] * 6, eps6 + (size_t)i * 6, sizeof(double) * 6);
  }
}

void mgpu_set_list(mgpu_ctx *c, int which_list, int n, const int *slots) {
  CK(cudaSetDevice(c->device));
  CK(cudaStreamSynchronize(c->stream));
  h2d_sync(c, c->d_list[which_list], slots, sizeof(int) * n);
}

// ---- kernels --------------------------------------------------------------------------------
void mgpu_load_u(mgpu_ctx *c, int l, int n, int which_u) {
  if (n <= 0) return;
  ProfScope ps(c, 9, n);
  const int len = 3 * c->mc.nn_pad;
  dim3 g(std::min((len + 255) / 256, 1024), n);
  k_load_u<<<g, 256, 0, c->stream>>>(c->mc, lst_of(c, l), c->T, c->V.u, c->V.vstride, which_u);
  CK(cudaGetLastError());
}
void mgpu_store_u(mgpu_ctx *c, int l, int n, int which_u) {
  if (n <= 0) return;
  ProfScope ps(c, 9, n);
  const int len = 3 * c->mc.nn_pad;
  dim3 g(std::min((len + 255) / 256, 1024), n);
  k_store_u<<<g, 256, 0, c->stream>>>(c->mc, lst_of(c, l), c->T, c->V.u, c->V.vstride, which_u);
  CK(cudaGetLastError());
}
void mgpu_zero_u(mgpu_ctx *c, int l, int n) {
  if (n <= 0) return;
  ProfScope ps(c, 9, n);
  const int len = 3 * c->mc.nn_pad;
  dim3 g(std::min((len + 255) / 256, 1024), n);
  k_zero_u<<<g, 256, 0, c->stream>>>(c->mc, lst_of(c, l), c->V.u, c->V.vstride);
  CK(cudaGetLastError());
}
void mgpu_set_bc(mgpu_ctx *c, int l, int n) {
  if (n <= 0) return;
  ProfScope ps(c, 9, n);
  k_set_bc<<<node_grid(c, n), NT, 0, c->stream>>>(c->mc, lst_of(c, l), c->T, c->V.u, c->V.vstride);
  CK(cudaGetLastError());
}
void mgpu_asm_rhs(mgpu_ctx *c, int l, int n, int mode) {
  if (n <= 0) return;
  ProfScope ps(c, 2, n);
  const size_t bstride = (size_t)24 * c->mc.nelem_pad;
  for (int off = 0; off < n; off += c->be_chunk) {
    const int cnt = std::min(c->be_chunk, n - off);
    const Lst lst = lst_of(c, l, off);
    k_elem_rhs<<<elem_grid(c, cnt), NT, 0, c->stream>>>(c->mc, lst, c->T, c->V.u, c->V.vstride, c->d_elem_type,
                                                       c->d_be, bstride, mode);
    k_asm_rhs<<<node_grid(c, cnt), NT, 0, c->stream>>>(c->mc, lst, c->T, c->V.b, c->V.vstride, c->d_be, bstride,
                                                      mode);
    c->launches += 1;
  }
  CK(cudaGetLastError());
}
void mgpu_asm_mat(mgpu_ctx *c, int l, int n, int to_shared) {
  if (n <= 0) return;
  double *shared = nullptr;
  if (to_shared) {
    if (!c->V.mat_shared) {
      CK(cudaMalloc(&c->V.mat_shared, sizeof(double) * c->V.mstride));
      CK(cudaMemsetAsync(c->V.mat_shared, 0, sizeof(double) * c->V.mstride, c->stream));
    }
    shared = c->V.mat_shared;
    n = 1;
  }
  if (!to_shared && n > c->mat_slots) {
    fprintf(stderr, "micropp-b200: explicit assembly of %d slots, but the matrix pool holds %d (implicit operator)\n",
            n, c->mat_slots);
    abort();
  }
  ProfScope ps(c, 1, n);
  if (c->all_elastic) {
    k_asm_mat_elastic<<<int_grid(c, n), NT, 0, c->stream>>>(c->mc, lst_of(c, l), c->V.mat, c->V.mstride, shared,
                                                            c->d_elem_type, c->d_ke);
  } else {
    const size_t cstride = (size_t)CTAN_LEN * c->mc.nelem_pad;
    for (int off = 0; off < n; off += c->ctan_chunk) {
      const int cnt = std::min(c->ctan_chunk, n - off);
      const Lst lst = lst_of(c, l, off);
      k_elem_ctan<<<elem_grid(c, cnt), NT, 0, c->stream>>>(c->mc, lst, c->T, c->V.u, c->V.vstride, c->d_elem_type,
                                                          c->d_ctan, cstride);
      dim3 g(std::max((c->mc.nint + GN - 1) / GN, 1), cnt);
      k_asm_mat_general<<<g, NT, GN * NPLANE * sizeof(double), c->stream>>>(
          c->mc, lst, c->V.mat, c->V.mstride, shared, c->d_elem_type, c->d_ke, c->d_ctan, cstride);
      c->launches += 1;
    }
  }
  CK(cudaGetLastError());
}
int mgpu_implicit(const mgpu_ctx *c) { return c->implicit ? 1 : 0; }
int mgpu_implicit_rows(const mgpu_ctx *c) { return c->nrows; }
int mgpu_implicit_fix_nodes(const mgpu_ctx *c) { return c->nfix; }

// the implicit SpMV kernel a request resolves to: 0 simple, 1 tiled (cp.async), 2 tiled (TMA)
static int imp_kernel_of(const mgpu_ctx *c, int kern) {
  const TileInfo &ti = c->tile;
  const int ntiles = ti.tiles_x * ti.tiles_y * ti.tiles_z;
  return std::min(kern, (c->imp_kernel == 2 && ntiles * ti.cb <= c->T.nblk_max) ? 2 : (ntiles <= c->T.nblk_max ? 1 : 0));
}
// partials of p.Ap that k_cg_update_imp would fold itself: measured slower (0.2 ms per DPCG iteration of 1024 RVEs at
// 30^3: every block repeats the fold before its stores) than the one-warp-per-slot k_fold_spmv launch => disabled
static int imp_fold_count(const mgpu_ctx *c) {
  (void)c;
  return 0;
}
// Ap = A p of the implicit operator over n entries of list l
static void launch_imp_spmv(mgpu_ctx *c, int l, int n, int force, int kern) {
  const TileInfo &ti = c->tile;
  const int ntiles = ti.tiles_x * ti.tiles_y * ti.tiles_z;
  // kern 2: the context's TMA kernel; kern 10 + v: TMA variant v (0 = k_spmv_dot_tma, v >= 1 = k_spmv_dot_tmac)
  int variant = c->tma_variant;
  if (kern >= 100) {  // measurement only: 100 = skip the compute, 200 = skip the loads (k_spmv_dot_tmac)
    force |= (kern / 100) << 1;
    kern %= 100;
  }
  if (kern >= 10) {
    variant = std::min(kern - 10, N_TMAC);
    kern = 2;
  }
  if (c->tile2.ntiles == 0) variant = 0;
  kern = imp_kernel_of(c, kern);
  if (kern == 2) {
    int rs = std::min(TMA_MAX_RS, n), ntl = rs >= 4 ? 1 : TMA_MAX_RS / rs;
    if (variant >= 1) {
      const TileInfo2 &t2 = c->tile2;
      // small groups of slots (MICROPP_CG_GROUP): fewer slots per block so that the grid still fills 148 SMs
      const long want_blocks = 148L * 4 * 2;
      rs = (int)std::max(1L, std::min((long)rs, (long)n * t2.ntiles / want_blocks));
      ntl = 1;
      if (n < 4) {  // one (or a few) large RVEs, e.g. a z-slab: several tiles per block, but at least ~4 waves of blocks
        rs = n;
        ntl = (int)std::max(1L, std::min((long)(TMA_MAX_RS / rs), (long)n * t2.ntiles / (148L * 4 * 4)));
      }
      const dim3 grid((t2.ntiles + ntl - 1) / ntl, (n + rs - 1) / rs);
      const int smem = kTmac[variant - 1].nstage * ((c->tile2_smem + 127) / 128 * 128) + 128;
      tmac_kernel(variant, t2.cb, t2.tn)<<<grid, 32 * t2.cb, smem, c->stream>>>(
          c->mc, lst_of(c, l), n, c->T, c->V, t2, c->tmap_a, c->tmap_b, c->pure_rows, ntl, rs, force);
      k_fold_spmv<<<dim3(1, n), 32, 0, c->stream>>>(c->mc, lst_of(c, l), c->T, t2.ntiles * t2.cb, force);
      c->launches++;
      return;
    }
    const dim3 grid((ntiles + ntl - 1) / ntl, (n + rs - 1) / rs);
    tma_kernel(ti.cb)<<<grid, 32 * ti.cb, c->tma_smem, c->stream>>>(c->mc, lst_of(c, l), n, c->T, c->V, ti, c->tmap_p,
                                                                   ntl, rs, force);
    // p.Ap: one warp per slot folds the per-(tile, warp) partials in a fixed order and runs the scalar tail
    k_fold_spmv<<<dim3(1, n), 32, 0, c->stream>>>(c->mc, lst_of(c, l), c->T, ntiles * ti.cb, force);
    c->launches++;
  } else if (kern == 1) {
    tile_kernel(ti.cb)<<<dim3(ntiles, n), 32 * ti.cb, c->tile_smem, c->stream>>>(c->mc, lst_of(c, l), c->T, c->V, ti,
                                                                                 force);
  } else {
    k_spmv_dot_imp<MR><<<int_grid(c, (n + MR - 1) / MR), NT, 0, c->stream>>>(c->mc, lst_of(c, l), n, c->T, c->V, force);
  }
}

// -1: no implicit operator; else the SpMV kernel it runs (0 simple, 1 tiled cp.async, 2 k_spmv_dot_tma, 3 k_spmv_dot_tmac)
int mgpu_implicit_kernel(const mgpu_ctx *c) {
  if (!c->implicit) return -1;
  const int k = imp_kernel_of(c, c->imp_kernel);
  return (k == 2 && c->tma_variant >= 1 && c->tile2.ntiles > 0) ? 3 : k;  // 3: k_spmv_dot_tmac
}

// OP_SLOT indexes the explicit matrix pool by slot: a context that serves all-elastic RVEs from the implicit operator
// keeps only mat_slots (<= 2) explicit matrices for the host-pointer API, so a wider OP_SLOT launch would read past it
static void require_mat_pool(const mgpu_ctx *c, int n, int op, const char *who) {
  if (op == OP_SLOT && n > c->mat_slots) {
    fprintf(stderr, "micropp-b200: %s over %d slots with per-slot matrices, but the matrix pool holds %d "
                    "(implicit operator)\n", who, n, c->mat_slots);
    abort();
  }
}

void mgpu_cg_init(mgpu_ctx *c, int l, int n, int use_shared) {
  if (n <= 0) return;
  require_mat_pool(c, n, use_shared, "mgpu_cg_init");
  if (use_shared == OP_IMPLICIT && !c->implicit) {
    fprintf(stderr, "micropp-b200: implicit operator requested for an RVE that is not all-elastic\n");
    abort();
  }
  c->cg_op = use_shared;
  ProfScope ps(c, 3, n);
  k_cg_init<<<node_grid(c, n), NT, 0, c->stream>>>(c->mc, lst_of(c, l), c->T, c->V, use_shared);
  CK(cudaGetLastError());
}
void mgpu_cg_spmv_dot(mgpu_ctx *c, int l, int n, int use_shared) {
  if (n <= 0) return;
  require_mat_pool(c, n, use_shared, "mgpu_cg_spmv_dot");
  ProfScope ps(c, 0, n);
  if (use_shared == OP_IMPLICIT) {
    launch_imp_spmv(c, l, n, 0, c->imp_kernel);
  } else {
    k_spmv_dot<<<int_grid(c, n), NT, 0, c->stream>>>(c->mc, lst_of(c, l), c->T, c->V, use_shared, 0);
  }
  CK(cudaGetLastError());
}
void mgpu_cg_update(mgpu_ctx *c, int l, int n) {
  if (n <= 0) return;
  ProfScope ps(c, 3, n);
  if (c->cg_op == OP_IMPLICIT)
    k_cg_update_imp<<<upd_grid(c, n), NT, 0, c->stream>>>(c->mc, lst_of(c, l), c->T, c->V, imp_fold_count(c));
  else
    k_cg_update<<<upd_grid(c, n), NT, 0, c->stream>>>(c->mc, lst_of(c, l), c->T, c->V);
  k_fold_update<<<dim3(1, n), 32, 0, c->stream>>>(c->mc, lst_of(c, l), c->T, (int)upd_grid(c, n).x);
  c->launches++;
  CK(cudaGetLastError());
}
void mgpu_cg_pupdate(mgpu_ctx *c, int l, int n) {
  if (n <= 0) return;
  ProfScope ps(c, 3, n);
  if (c->cg_op == OP_IMPLICIT)
    k_cg_pupdate_imp<<<node_grid(c, n), NT, 0, c->stream>>>(c->mc, lst_of(c, l), c->T, c->V);
  else
    k_cg_pupdate<<<node_grid(c, n), NT, 0, c->stream>>>(c->mc, lst_of(c, l), c->T, c->V);
  CK(cudaGetLastError());
}
void mgpu_cg_finish(mgpu_ctx *c, int l, int n) {
  if (n <= 0) return;
  ProfScope ps(c, 3, n);
  k_cg_finish<<<node_grid(c, n), NT, 0, c->stream>>>(c->mc, lst_of(c, l), c->T, c->V);
  CK(cudaGetLastError());
}
void mgpu_axpy_u(mgpu_ctx *c, int l, int n) {
  if (n <= 0) return;
  ProfScope ps(c, 9, n);
  k_axpy_u<<<node_grid(c, n), NT, 0, c->stream>>>(c->mc, lst_of(c, l), c->T, c->V);
  CK(cudaGetLastError());
}
void mgpu_ave_stress(mgpu_ctx *c, int l, int n) {
  if (n <= 0) return;
  ProfScope ps(c, 9, n);
  k_ave_stress<<<elem_grid(c, n), NT, 0, c->stream>>>(c->mc, lst_of(c, l), c->T, c->V.u, c->V.vstride,
                                                     c->d_elem_type);
  CK(cudaGetLastError());
}
void mgpu_vars_new(mgpu_ctx *c, int l, int n, int write) {
  if (n <= 0) return;
  ProfScope ps(c, 9, n);
  k_vars_new<<<elem_grid(c, n), NT, 0, c->stream>>>(c->mc, lst_of(c, l), c->T, c->V.u, c->V.vstride,
                                                   c->d_elem_type, write);
  CK(cudaGetLastError());
}
// element-averaged strain / stress of the RVE staged in `slot` (host arrays of 6*nelem doubles each)
void mgpu_elem_fields(mgpu_ctx *c, int slot, double ivol, double *elem_strain, double *elem_stress) {
  CK(cudaSetDevice(c->device));
  const size_t len = (size_t)c->mc.nelem * 6;
  double *d = nullptr;
  CK(cudaMalloc(&d, sizeof(double) * 2 * len));
  c->launches++;
  k_elem_fields<<<(c->mc.nelem + NT - 1) / NT, NT, 0, c->stream>>>(c->mc, slot, c->T, c->V.u, c->V.vstride,
                                                                  c->d_elem_type, ivol, d);
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(c->stream));
  CK(cudaMemcpy(elem_strain, d, sizeof(double) * len, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(elem_stress, d + len, sizeof(double) * len, cudaMemcpyDeviceToHost));
  CK(cudaFree(d));
}
void mgpu_clear_nl_flags(mgpu_ctx *c, int l, int n) {
  if (n <= 0) return;
  c->launches++;
  k_clear_nl<<<(n + 255) / 256, 256, 0, c->stream>>>(c->d_list[l], n, c->T);
  CK(cudaGetLastError());
}
int mgpu_compact(mgpu_ctx *c, int list_in, int n_in, int list_out, int mode) {
  return mgpu_compact_range(c, list_in, 0, n_in, list_out, mode);
}
// same over entries [off, off + n_in) of list_in (a group of the wave)
int mgpu_compact_range(mgpu_ctx *c, int list_in, int off, int n_in, int list_out, int mode) {
  if (n_in <= 0) return 0;
  c->launches++;
  // lists 1 (Newton) and 2 (CG) keep their length on the device too: the step graphs start from it
  int *count2 = list_out == 1 ? c->d_cnt2 : (list_out == 2 ? c->d_cnt2 + 1 : nullptr);
  k_compact<<<1, 1024, 0, c->stream>>>(c->d_list[list_in] + off, n_in, nullptr, c->d_list[list_out], c->d_count, count2,
                                       c->T, mode, cudaGraphConditionalHandle(), 0, nullptr);
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(c->h_count, c->d_count, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  return *c->h_count;
}

// ---- one Newton step as ONE CUDA graph --------------------------------------------------------------
// assembly_mat -> cg_init -> WHILE(any slot still iterating){ SpMV+dot ; x,r,z update ; p update ; compaction }
// -> u += du -> assembly_rhs (+ Newton loop-head test) -> compaction of the Newton list, for all slots of the
// Newton list at once (src/solve.cpp:49-77 with src/ell.cpp:93-119 inside).  The DPCG loop is a conditional WHILE
// node driven from the device (cudaGraphSetConditional in k_compact), so the host is not involved between the
// start of a Newton step and its end.  Launch shapes are fixed per graph (bucket = power of two >= active slots);
// blocks beyond the device-side list length leave immediately (Lst::dcount).
namespace {

void capture_begin(mgpu_ctx *c, cudaGraph_t g, const cudaGraphNode_t *deps, size_t ndeps) {
  CK(cudaStreamBeginCaptureToGraph(c->stream, g, deps, nullptr, ndeps, cudaStreamCaptureModeRelaxed));
}
// ends the capture and returns the leaf nodes captured so far (dependencies of whatever comes next)
std::vector<cudaGraphNode_t> capture_end(mgpu_ctx *c) {
  cudaStreamCaptureStatus st;
  const cudaGraphNode_t *deps = nullptr;
  size_t nd = 0;
  CK(cudaStreamGetCaptureInfo_v2(c->stream, &st, nullptr, nullptr, &deps, &nd));
  std::vector<cudaGraphNode_t> out(deps, deps + nd);
  cudaGraph_t g = nullptr;
  CK(cudaStreamEndCapture(c->stream, &g));
  return out;
}

mgpu_ctx::StepGraph build_step_graph(mgpu_ctx *c, int B, int use_shared) {
  const bool prof = c->prof;
  c->prof = false;  // no event records inside a capture
  const unsigned long long l0 = c->launches;
  cudaGraph_t g;
  CK(cudaGraphCreate(&g, 0));
  cudaGraphConditionalHandle cond;
  CK(cudaGraphConditionalHandleCreate(&cond, g, 0, cudaGraphCondAssignDefault));

  // head: Jacobian, CG start, CG list := Newton-list slots whose loop-head test says "iterate"
  capture_begin(c, g, nullptr, 0);
  c->dyn_count = c->d_cnt2;
  if (use_shared == OP_SLOT) mgpu_asm_mat(c, 1, B, 0);
  mgpu_cg_init(c, 1, B, use_shared);
  k_compact<<<1, 1024, 0, c->stream>>>(c->d_list[1], 0, c->d_cnt2, c->d_list[2], c->d_cnt2 + 1, nullptr, c->T, 1, cond,
                                       1, nullptr);
  c->launches++;
  std::vector<cudaGraphNode_t> leaves = capture_end(c);
  const int head = (int)(c->launches - l0);

  // the DPCG loop
  cudaGraphNodeParams wp = {};
  wp.type = cudaGraphNodeTypeConditional;
  wp.conditional.handle = cond;
  wp.conditional.type = cudaGraphCondTypeWhile;
  wp.conditional.size = 1;
  cudaGraphNode_t wnode;
  CK(cudaGraphAddNode(&wnode, g, leaves.data(), leaves.size(), &wp));
  cudaGraph_t body = wp.conditional.phGraph_out[0];
  const unsigned long long l1 = c->launches;
  capture_begin(c, body, nullptr, 0);
  c->dyn_count = c->d_cnt2 + 1;
  mgpu_cg_spmv_dot(c, 2, B, use_shared);
  mgpu_cg_update(c, 2, B);
  mgpu_cg_pupdate(c, 2, B);
  k_compact<<<1, 1024, 0, c->stream>>>(c->d_list[2], 0, c->d_cnt2 + 1, c->d_list[2], c->d_cnt2 + 1, nullptr, c->T, 1,
                                       cond, 1, c->d_cnt2 + 2);
  c->launches++;
  capture_end(c);
  const int body_l = (int)(c->launches - l1);

  // tail: update, residual + Newton test, Newton list compaction, list lengths to the host
  const unsigned long long l2 = c->launches;
  capture_begin(c, g, &wnode, 1);
  c->dyn_count = c->d_cnt2;
  mgpu_cg_finish(c, 1, B);
  mgpu_axpy_u(c, 1, B);
  mgpu_asm_rhs(c, 1, B, 1);
  k_compact<<<1, 1024, 0, c->stream>>>(c->d_list[1], 0, c->d_cnt2, c->d_list[1], c->d_cnt2, nullptr, c->T, 0, cond, 0,
                                       nullptr);
  c->launches++;
  CK(cudaMemcpyAsync(c->h_count, c->d_cnt2, 3 * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  capture_end(c);
  const int tail = (int)(c->launches - l2);
  c->dyn_count = nullptr;
  c->launches = l0;  // captured, not executed
  c->prof = prof;

  mgpu_ctx::StepGraph sg;
  CK(cudaGraphInstantiate(&sg.exec, g, 0));
  CK(cudaGraphDestroy(g));
  sg.fixed_launches = head + tail;
  sg.body_launches = body_l;
  return sg;
}

}  // namespace

extern "C" int mgpu_newton_step_graph(mgpu_ctx *c, int n_active, int use_shared) {
  if (n_active <= 0) return 0;
  CK(cudaSetDevice(c->device));
  int B = 1;
  while (B < n_active) B <<= 1;
  B = std::min(B, c->W);
  const long long key = (long long)B * 4 + use_shared;
  auto it = c->step_graphs.find(key);
  if (it == c->step_graphs.end()) it = c->step_graphs.emplace(key, build_step_graph(c, B, use_shared)).first;
  CK(cudaMemsetAsync(c->d_cnt2 + 2, 0, sizeof(int), c->stream));
  CK(cudaGraphLaunch(it->second.exec, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  c->launches += it->second.fixed_launches + (unsigned long long)it->second.body_launches * c->h_count[2];
  return c->h_count[0];
}

// ---- slab mode ---------------------------------------------------------------------------------
extern "C" void mgpu_tail(mgpu_ctx *c, int l, int n, int kind, int mode) {
  if (n <= 0) return;
  c->launches++;
  k_tail<<<dim3(1, n), 32, 0, c->stream>>>(c->mc, lst_of(c, l), c->T, kind, mode);
  CK(cudaGetLastError());
}
extern "C" void *mgpu_dev_ptr(mgpu_ctx *c, int which) {
  if (which == 10) return c->T.red;
  if (which == 11) return c->T.stress;
  return vec_of(c, which);
}
extern "C" void *mgpu_stream(mgpu_ctx *c) { return (void *)c->stream; }

// ---- slab mode over peer memory ------------------------------------------------------------------
extern "C" {
// this rank's mailbox (device memory, zeroed); export it with mgpu_ipc_export
void *mgpu_slab_mail(mgpu_ctx *c) {
  CK(cudaSetDevice(c->device));
  if (!c->slab_mail) {
    CK(cudaMalloc(&c->slab_mail, sizeof(SlabMail)));
    CK(cudaMemset(c->slab_mail, 0, sizeof(SlabMail)));
    CK(cudaDeviceSynchronize());
  }
  return c->slab_mail;
}
// 64-byte CUDA IPC handle of a device allocation of this process / mapping of another process' handle
void mgpu_ipc_export(void *devptr, char *handle64) {
  cudaIpcMemHandle_t h;
  CK(cudaIpcGetMemHandle(&h, devptr));
  memcpy(handle64, &h, sizeof(h));
}
void *mgpu_ipc_open(int device, const char *handle64) {
  CK(cudaSetDevice(device));
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, sizeof(h));
  void *p = nullptr;
  CK(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
  return p;
}
void mgpu_ipc_close(void *p) { cudaIpcCloseMemHandle(p); }
// mails[r]: mailbox of rank r as mapped here (r == rank: the own one); p_lo / p_hi: the neighbours' p vectors
// (null at the ends), *_off: offset (doubles, component 0) of the neighbour's owned plane next to the cut,
// *_npad: the neighbour's component stride
void mgpu_slab_link(mgpu_ctx *c, int rank, int size, void *const *mails, const void *p_lo, long long lo_off,
                    long long lo_npad, const void *p_hi, long long hi_off, long long hi_npad) {
  if (size > SLAB_MAX_RANKS) {
    fprintf(stderr, "micropp-b200: at most %d slabs\n", SLAB_MAX_RANKS);
    abort();
  }
  mgpu_slab_mail(c);
  c->slab_rank = rank;
  c->slab_size = size;
  for (int r = 0; r < size; ++r) c->slab_peers.mail[r] = (SlabMail *)mails[r];
  c->slab_p_lo = (const double *)p_lo;
  c->slab_p_hi = (const double *)p_hi;
  c->slab_lo_off = lo_off;
  c->slab_lo_npad = lo_npad;
  c->slab_hi_off = hi_off;
  c->slab_hi_npad = hi_npad;
}
void mgpu_slab_publish_p(mgpu_ctx *c) {
  c->launches++;
  k_slab_publish_p<<<1, 1, 0, c->stream>>>(c->slab_mail);
  CK(cudaGetLastError());
}
void mgpu_slab_halo_pull(mgpu_ctx *c) {
  c->launches++;
  const int nb = std::min(64, (c->mc.nxny + NT - 1) / NT);
  k_slab_halo_pull<<<dim3(nb, 2), NT, 0, c->stream>>>(
      c->mc, c->V.p, c->slab_mail, c->slab_p_lo, c->slab_lo_off, c->slab_lo_npad,
      c->slab_rank > 0 ? c->slab_peers.mail[c->slab_rank - 1] : nullptr, c->slab_p_hi, c->slab_hi_off, c->slab_hi_npad,
      c->slab_rank + 1 < c->slab_size ? c->slab_peers.mail[c->slab_rank + 1] : nullptr);
  CK(cudaGetLastError());
}
void mgpu_slab_post(mgpu_ctx *c, int k) {
  c->launches++;
  k_slab_post<<<1, 32, 0, c->stream>>>(c->T.red, c->slab_mail, k);
  CK(cudaGetLastError());
}
void mgpu_slab_gather_tail(mgpu_ctx *c, int l, int k, int kind, int mode) {
  c->launches++;
  k_slab_gather_tail<<<dim3(1, 1), 32, 0, c->stream>>>(c->mc, lst_of(c, l), c->T, c->slab_peers, c->slab_mail,
                                                      c->slab_size, k, kind, mode);
  CK(cudaGetLastError());
}
int mgpu_slab_error(mgpu_ctx *c) {
  CK(cudaSetDevice(c->device));
  CK(cudaStreamSynchronize(c->stream));
  SlabMail m;
  CK(cudaMemcpy(&m, c->slab_mail, sizeof(m), cudaMemcpyDeviceToHost));
  return m.error;
}
// One DPCG iteration of this rank's slab (multi-process mode: every rank calls it; no host synchronisation):
// halo pull -> SpMV -> post/tail(p.Ap) -> r, z update -> post/tail(z.z, r.z) -> p, x update -> publish p
void mgpu_slab_cg_iteration(mgpu_ctx *c, int l, int op) {
  mgpu_slab_halo_pull(c);
  mgpu_cg_spmv_dot(c, l, 1, op);
  mgpu_slab_post(c, 1);
  mgpu_slab_gather_tail(c, l, 1, 2, 0);
  mgpu_cg_update(c, l, 1);
  mgpu_slab_post(c, 2);
  mgpu_slab_gather_tail(c, l, 2, 3, 0);
  mgpu_cg_pupdate(c, l, 1);
  mgpu_slab_publish_p(c);
}
// `iters` DPCG iterations of this rank's slab as ONE CUDA graph launch (captured once per (op, iters); the epochs of
// the cross-rank flags live on the device, so every kernel argument is constant).  Slots that converge inside the
// chunk skip their kernels (cg_active), exactly as in the single-domain chunked loop.
void mgpu_slab_cg_chunk(mgpu_ctx *c, int l, int op, int iters) {
  CK(cudaSetDevice(c->device));
  const int key = op * 1024 + iters;
  auto it = c->slab_chunk_graphs.find(key);
  if (it == c->slab_chunk_graphs.end()) {
    const bool prof = c->prof;
    c->prof = false;
    const unsigned long long l0 = c->launches;
    cudaGraph_t g = nullptr;
    CK(cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeRelaxed));
    for (int k = 0; k < iters; ++k) mgpu_slab_cg_iteration(c, l, op);
    CK(cudaStreamEndCapture(c->stream, &g));
    cudaGraphExec_t ex = nullptr;
    CK(cudaGraphInstantiate(&ex, g, 0));
    CK(cudaGraphDestroy(g));
    c->slab_launches_per_chunk = (int)(c->launches - l0);
    c->launches = l0;
    c->prof = prof;
    it = c->slab_chunk_graphs.emplace(key, ex).first;
  }
  CK(cudaGraphLaunch(it->second, c->stream));
  c->launches += c->slab_launches_per_chunk;
}
}

// ---- results ----------------------------------------------------------------------------------
void mgpu_fetch_state(mgpu_ctx *c, int n, const int *slots, mgpu_slot_state *out) {
  CK(cudaSetDevice(c->device));
  CK(cudaStreamSynchronize(c->stream));
  bool contiguous = true;
  for (int i = 1; i < n; ++i) contiguous &= (slots[i] == slots[0] + i);
  if (contiguous && n > 0) {
    CK(cudaMemcpy(out, c->T.state + slots[0], sizeof(mgpu_slot_state) * n, cudaMemcpyDeviceToHost));
  } else {
    for (int i = 0; i < n; ++i)
      CK(cudaMemcpy(out + i, c->T.state + slots[i], sizeof(mgpu_slot_state), cudaMemcpyDeviceToHost));
  }
}
void mgpu_fetch_stress(mgpu_ctx *c, int n, const int *slots, double *sig6) {
  CK(cudaSetDevice(c->device));
  CK(cudaStreamSynchronize(c->stream));
  bool contiguous = true;
  for (int i = 1; i < n; ++i) contiguous &= (slots[i] == slots[0] + i);
  if (contiguous && n > 0) {
    CK(cudaMemcpy(sig6, c->T.stress + (size_t)slots[0] * 6, sizeof(double) * 6 * n, cudaMemcpyDeviceToHost));
  } else {
    for (int i = 0; i < n; ++i)
      CK(cudaMemcpy(sig6 + (size_t)i * 6, c->T.stress + (size_t)slots[i] * 6, sizeof(double) * 6,
                    cudaMemcpyDeviceToHost));
  }
}

// ---- staging ------------------------------------------------------------------------------------
void mgpu_stage_put_vec(mgpu_ctx *c, int slot, int which, const double *host_aos) {
  CK(cudaSetDevice(c->device));
  std::vector<double> tmp;
  aos_to_soa(c, host_aos, tmp);
  CK(cudaStreamSynchronize(c->stream));
  h2d_sync(c, vec_of(c, which) + (size_t)slot * c->V.vstride, tmp.data(), tmp.size() * sizeof(double));
}
void mgpu_stage_get_vec(mgpu_ctx *c, int slot, int which, double *host_aos) {
  CK(cudaSetDevice(c->device));
  std::vector<double> tmp((size_t)3 * c->mc.nn_pad);
  CK(cudaStreamSynchronize(c->stream));
  CK(cudaMemcpy(tmp.data(), vec_of(c, which) + (size_t)slot * c->V.vstride, tmp.size() * sizeof(double),
                cudaMemcpyDeviceToHost));
  soa_to_aos(c, tmp, host_aos);
}
void mgpu_stage_put_u(mgpu_ctx *c, int slot, const double *host_aos) { mgpu_stage_put_vec(c, slot, 4, host_aos); }
void mgpu_stage_get_u(mgpu_ctx *c, int slot, double *host_aos) { mgpu_stage_get_vec(c, slot, 4, host_aos); }

void mgpu_stage_put_vars(mgpu_ctx *c, int slot, int which, const double *ref) {
  CK(cudaSetDevice(c->device));
  CK(cudaStreamSynchronize(c->stream));
  auto &bufs = c->stage_vars[which];
  if ((int)bufs.size() < c->W) bufs.resize(c->W, nullptr);
  if (ref) {
    if (!bufs[slot]) CK(cudaMalloc(&bufs[slot], sizeof(double) * c->var_len));
    std::vector<double> tmp;
    vars_ref_to_internal(c, ref, tmp);
    tmp.resize(c->var_len, 0.0);
    h2d_sync(c, bufs[slot], tmp.data(), sizeof(double) * c->var_len);
  }
  if (which == 0)
    c->h_vars_old[slot] = ref ? bufs[slot] : nullptr;
  else
    c->h_vars_new[slot] = ref ? bufs[slot] : nullptr;
  upload_slot_tables(c);
  CK(cudaStreamSynchronize(c->stream));
}
void mgpu_stage_get_vars_new(mgpu_ctx *c, int slot, double *ref) {
  CK(cudaSetDevice(c->device));
  CK(cudaStreamSynchronize(c->stream));
  std::vector<double> tmp(c->var_len, 0.0);
  if (c->h_vars_new[slot])
    CK(cudaMemcpy(tmp.data(), c->h_vars_new[slot], sizeof(double) * c->var_len, cudaMemcpyDeviceToHost));
  vars_internal_to_ref(c, tmp, ref);
}
void mgpu_stage_get_mat(mgpu_ctx *c, int slot, double *vals) {
  CK(cudaSetDevice(c->device));
  CK(cudaStreamSynchronize(c->stream));
  const MeshConst &P = c->mc;
  std::vector<double> tmp(c->V.mstride);
  CK(cudaMemcpy(tmp.data(), c->V.mat + (size_t)slot * c->V.mstride, sizeof(double) * c->V.mstride,
                cudaMemcpyDeviceToHost));
  // reference layout: vals[(3n+fi)*81 + nbr*3 + fj]; boundary rows are the identity rows of ell_set_bc_3D
  memset(vals, 0, sizeof(double) * (size_t)P.nn * 3 * 81);
  for (int k = 0; k < P.nz; ++k)
    for (int j = 0; j < P.ny; ++j)
      for (int i = 0; i < P.nx; ++i) {
        const int n = k * P.nxny + j * P.nx + i;
        if (i == 0 || i == P.nx - 1 || j == 0 || j == P.ny - 1 || k == 0 || k == P.nz - 1) {
          for (int d = 0; d < 3; ++d) vals[((size_t)n * 3 + d) * 81 + 13 * 3 + d] = 1.0;
          continue;
        }
        const int m = ((k - 1) * P.niy + (j - 1)) * P.nix + (i - 1);
        for (int nbr = 0; nbr < 27; ++nbr)
          for (int fi = 0; fi < 3; ++fi)
            for (int fj = 0; fj < 3; ++fj)
              vals[((size_t)n * 3 + fi) * 81 + nbr * 3 + fj] = tmp[aidx(nbr * 9 + fi * 3 + fj, m)];
      }
}
// host matrix of the generic ELL API (ell_mvp / ell_solve_cgpd on caller-provided values): kept as it is
void mgpu_stage_put_mat(mgpu_ctx *c, int slot, const double *vals) {
  (void)slot;
  CK(cudaSetDevice(c->device));
  CK(cudaStreamSynchronize(c->stream));
  const size_t len = (size_t)c->mc.nn * 3 * 81;
  if (!c->V.gen) CK(cudaMalloc(&c->V.gen, sizeof(double) * len));
  h2d_sync(c, c->V.gen, vals, sizeof(double) * len);
}

void mgpu_ell_cols(int nx, int ny, int nz, int *cols, int device) {
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    fprintf(stderr, "micropp-b200: no CUDA device available; this library has no CPU path\n");
    abort();
  }
  CK(cudaSetDevice(device % ndev));
  const size_t nrow = (size_t)3 * nx * ny * nz;
  int *d = nullptr;
  CK(cudaMalloc(&d, sizeof(int) * nrow * 81));
  k_ell_cols<<<(unsigned)((nrow + 127) / 128), 128>>>(nx, ny, nz, d);
  CK(cudaGetLastError());
  CK(cudaMemcpy(cols, d, sizeof(int) * nrow * 81, cudaMemcpyDeviceToHost));
  CK(cudaFree(d));
}

// generic-matrix SpMV / CG steps for the host-pointer ELL API (arbitrary vals, no identity shortcut)
void mgpu_spmv_generic(mgpu_ctx *c, int l, int n, int force) {
  if (n <= 0) return;
  ProfScope ps(c, 0, n);
  k_spmv_generic<<<node_grid(c, n), NT, 0, c->stream>>>(c->mc, lst_of(c, l), c->T, c->V, force);
  CK(cudaGetLastError());
}

// One forced operator application on the slots of a list (parity tests of the SpMV kernels): Ap = A p with p as
// it stands in the pool; op as in mgpu_cg_init; imp_kernel: -1 the context's choice, 0 simple, 1 tiled
void mgpu_apply_operator(mgpu_ctx *c, int l, int n, int op, int imp_kernel) {
  if (n <= 0) return;
  CK(cudaSetDevice(c->device));
  c->launches++;
  const int kern = imp_kernel < 0 ? c->imp_kernel : imp_kernel;
  if (op == OP_IMPLICIT) {
    if (!c->implicit) {
      fprintf(stderr, "micropp-b200: implicit operator requested for an RVE that is not all-elastic\n");
      abort();
    }
    launch_imp_spmv(c, l, n, 1, kern);
  } else if (op == OP_GENERIC) {
    k_spmv_generic<<<node_grid(c, n), NT, 0, c->stream>>>(c->mc, lst_of(c, l), c->T, c->V, 1);
  } else {
    require_mat_pool(c, n, op, "mgpu_apply_operator");
    k_spmv_dot<<<int_grid(c, n), NT, 0, c->stream>>>(c->mc, lst_of(c, l), c->T, c->V, op, 1);
  }
  CK(cudaGetLastError());
}

// ---- measurement ---------------------------------------------------------------------------------
void mgpu_prof_enable(mgpu_ctx *c, int on) {
  prof_drain(c);
  c->prof = on != 0;
}
void mgpu_prof_read(mgpu_ctx *c, double *out6, int reset) {
  prof_drain(c);
  for (int i = 0; i < 6; ++i) out6[i] = c->prof_acc[i];
  if (reset)
    for (int i = 0; i < 6; ++i) c->prof_acc[i] = 0;
}
void mgpu_timer_start(mgpu_ctx *c) { CK(cudaEventRecord(c->t0, c->stream)); }
float mgpu_timer_stop(mgpu_ctx *c) {
  CK(cudaEventRecord(c->t1, c->stream));
  CK(cudaEventSynchronize(c->t1));
  float ms = 0.f;
  CK(cudaEventElapsedTime(&ms, c->t0, c->t1));
  return ms;
}
float mgpu_bench_spmv(mgpu_ctx *c, int n, int iters) {
  CK(cudaSetDevice(c->device));
  n = std::min(n, std::min(c->W, c->mat_slots));  // slots that own an explicit matrix
  std::vector<int> ids(n);
  for (int i = 0; i < n; ++i) ids[i] = i;
  mgpu_set_list(c, 5, n, ids.data());
  for (int w = 0; w < 2; ++w) {
    k_spmv_dot<<<int_grid(c, n), NT, 0, c->stream>>>(c->mc, lst_of(c, 5), c->T, c->V, 0, 1);
  }
  CK(cudaEventRecord(c->t0, c->stream));
  for (int it = 0; it < iters; ++it) {
    k_spmv_dot<<<int_grid(c, n), NT, 0, c->stream>>>(c->mc, lst_of(c, 5), c->T, c->V, 0, 1);
  }
  CK(cudaEventRecord(c->t1, c->stream));
  CK(cudaEventSynchronize(c->t1));
  CK(cudaGetLastError());
  float ms = 0.f;
  CK(cudaEventElapsedTime(&ms, c->t0, c->t1));
  c->launches += iters + 2;
  return ms / iters;
}

// isolated micro-benchmark of the implicit-operator SpMV (+ its p.Ap fold) on the first n slots; kern as in
// mgpu_apply_operator (2 = the context's kernel, 10 + v = TMA variant v).  p as it stands in the pool.
float mgpu_bench_imp_spmv(mgpu_ctx *c, int n, int iters, int kern) {
  CK(cudaSetDevice(c->device));
  if (!c->implicit) return -1.f;
  n = std::min(n, c->W);
  std::vector<int> ids(n);
  for (int i = 0; i < n; ++i) ids[i] = i;
  mgpu_set_list(c, 5, n, ids.data());
  for (int w = 0; w < 2; ++w) launch_imp_spmv(c, 5, n, 1, kern);
  CK(cudaEventRecord(c->t0, c->stream));
  for (int it = 0; it < iters; ++it) launch_imp_spmv(c, 5, n, 1, kern);
  CK(cudaEventRecord(c->t1, c->stream));
  CK(cudaEventSynchronize(c->t1));
  CK(cudaGetLastError());
  float ms = 0.f;
  CK(cudaEventElapsedTime(&ms, c->t0, c->t1));
  c->launches += 2 * (iters + 2);
  return ms / iters;
}

}  // extern "C"
