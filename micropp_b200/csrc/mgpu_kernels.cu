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
//   k_cg_update_imp, k_cg_pupdate_imp       DPCG vector kernels of the implicit operator (SpMV: spmv_implicit.cu)
//   k_axpy_u, k_ave_stress, k_vars_new, k_elem_fields, k_compact        Newton update, averages, history, lists
//   k_slab_*                                z-slab mode of one large RVE over NVLink peer memory
//   mgpu_*                                  the C ABI; mgpu_newton_step_graph = one Newton step as one CUDA graph
// The implicit-operator SpMV kernels live in spmv_implicit.cu; shared structures and helpers in mgpu_internal.cuh.
#include "mgpu_internal.cuh"

#include <execinfo.h>
#include <signal.h>
#include <unistd.h>

using namespace mgpu_int;

namespace {

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

// assembly_rhs of an ALL-ELASTIC RVE through the implicit operator: every material is linear, so the internal force is
// K u and the interior rows of -b are the ELL row blocks applied to u WITH its boundary values (the row block of an
// interior node holds the coupling to its boundary neighbours; the Dirichlet rows themselves are identity rows and b
// vanishes there, src/assembly.cpp:61-104).  u is copied into the operator's input vector, the implicit SpMV runs once
// (243 FMA and 48 B per node instead of 8 Gauss-point stress evaluations per element and a 192 B per element scratch
// round trip), and b = -A u with ||b||^2 and the Newton loop-head logic follows.  Same mathematics as the element loop
// of k_elem_rhs + k_asm_rhs, other rounding (1e-15 of |K||u|).
__global__ void __launch_bounds__(NT)
    k_u_to_p(const __grid_constant__ MeshConst P, const Lst L, SlotTables T, VecPool V, int mode) {
  const int slot = slot_of(L);
  if (slot < 0) return;
  if (mode == 1 && !T.state[slot].nr_active) return;
  const int n = blockIdx.x * NT + threadIdx.x;
  if (n >= P.nn) return;
  const size_t vo = (size_t)slot * V.vstride;
#pragma unroll
  for (int d = 0; d < 3; ++d) V.p[vo + (size_t)d * P.nn_pad + n] = V.u[vo + (size_t)d * P.nn_pad + n];
}
__global__ void __launch_bounds__(NT)
    k_rhs_from_ap(const __grid_constant__ MeshConst P, const Lst L, SlotTables T, VecPool V, int mode) {
  __shared__ double sm[NRED * (NT / 32)];
  __shared__ int sflag;
  const int slot = slot_of(L);
  if (slot < 0) return;
  mgpu_slot_state *st = &T.state[slot];
  if (mode == 1 && !st->nr_active) return;
  const size_t vo = (size_t)slot * V.vstride;
  const int n = blockIdx.x * NT + threadIdx.x;
  double nrm[1] = {0.0};
  if (n < P.nn) {
    int i, j, k;
    node_ijk(P, n, i, j, k);
    const bool bnd = on_boundary(P, i, j, k);
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      const size_t ix = vo + (size_t)d * P.nn_pad + n;
      const double bv = bnd ? 0.0 : -V.Ap[ix];
      V.b[ix] = bv;
      nrm[0] += bv * bv;
    }
  }
  double *partial = T.partial + (size_t)slot * NRED * T.nblk_max;
  if (grid_sum<1>(nrm, partial, T.nblk_max, &st->ticket, sm, &sflag) && threadIdx.x == 0) tail_rhs(P, st, nrm[0], mode);
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
// (1) k_elem_ctan: thread per (element, Gauss point).  Gathers the 24 element dofs and evaluates the reference's
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

__global__ void __launch_bounds__(NT, 4)
    k_elem_ctan(const __grid_constant__ MeshConst P, const Lst L, SlotTables T,
                const double *__restrict__ u_pool, size_t vstride, const int *__restrict__ elem_type,
                double *__restrict__ cbuf, size_t cstride, int hybrid) {
  const int slot = slot_of(L);
  if (slot < 0) return;
  const double *u = u_pool + (size_t)slot * vstride;
  const double *vars = T.vars_old[slot];
  double *cb = cbuf + (size_t)blockIdx.y * cstride;
  // thread per (element, Gauss point); the Gauss point is uniform over a block (grid.x = 8 * blocks per element sweep)
  const int nblk_e = (P.nelem + NT - 1) / NT;
  const int gp = blockIdx.x / nblk_e;
  const int e = (blockIdx.x - gp * nblk_e) * NT + threadIdx.x;
  if (e >= P.nelem) return;
  const int type = __ldg(&elem_type[e]);
  const mpp_material m = P.mat[type];
  if (m.type == MPP_ELASTIC) return;  // rows come from the per-material table
  // hybrid operator: elements inside their linear regime take the rows of their material's elastic law too
  if (hybrid && !T.enl[(size_t)slot * P.nelem_pad + e]) return;
  const int ez = e / (P.nex * P.ney);
  const int r = e - ez * P.nex * P.ney;
  const int ey = r / P.nex, ex = r - ey * P.nex;
  double ue[24];
  gather_ue(P, u, ex, ey, ez, ue);
  const int nv = mat_nvar(m.type);
  double eps[6], C[36], vbuf[7];
  gp_strain(P.dsh[gp], ue, eps);
  const double *v = fetch_vars(vars, P.nelem_pad, e, gp, nv, vbuf);
  mat_ctan(m, eps, v, C);
#pragma unroll
  for (int q = 0; q < 36; ++q) cb[(size_t)(gp * 36 + q) * P.nelem_pad + e] = C[q];
}

__global__ void __launch_bounds__(NT)
    k_asm_mat_general(const __grid_constant__ MeshConst P, const Lst L, double *mat_pool,
                      size_t mstride, double *mat_shared, const int *__restrict__ elem_type,
                      const double *__restrict__ ke_tab, const double *__restrict__ cbuf, size_t cstride,
                      SlotTables T, int hybrid) {
  extern __shared__ double s_acc[];  // [GN][243]
  const int slot = slot_of(L);
  if (slot < 0) return;
  const double *cb = cbuf + (size_t)blockIdx.y * cstride;
  double *A = mat_shared ? mat_shared : mat_pool + (size_t)slot * mstride;
  // hybrid operator: only the rows of the slot's compact node list are assembled (row of list position q at tile
  // position q); elements in their linear regime contribute the rows of their elastic law
  const int nrows_out = hybrid ? T.hcnt[slot] : P.nint;
  if ((int)blockIdx.x * GN >= nrows_out) return;
  const int *hnodes = hybrid ? T.hnodes + (size_t)slot * P.nint_pad : nullptr;
  const unsigned char *enl = hybrid ? T.enl + (size_t)slot * P.nelem_pad : nullptr;

  for (int t = threadIdx.x; t < GN * NPLANE; t += NT) s_acc[t] = 0.0;
  __syncthreads();

  // element-major thread order: the 16 lanes of a half-warp hold the SAME corner element of 16 x-consecutive nodes,
  // i.e. 16 consecutive elements => their tangent loads are 128-B segments (node-major order: 8 segments of 32 B)
  const int ln = threadIdx.x % GN, c = threadIdx.x / GN;
  const int q = blockIdx.x * GN + ln;  // row position (= interior-node index unless a node list is given)
  int i = 0, j = 0, k = 0;
  const bool work = q < nrows_out;
  if (work) interior_node(P, hnodes ? hnodes[q] : q, i, j, k);

  double R[72];
  int a = 0;
  if (work) {
    const int ax = (c >> 2) & 1, ay = (c >> 1) & 1, az = c & 1;
    const int ex = i - 1 + ax, ey = j - 1 + ay, ez = k - 1 + az;
    a = corner_of(1 - ax, 1 - ay, 1 - az);
    const int e = (ez * P.ney + ey) * P.nex + ex;
    const int type = __ldg(&elem_type[e]);
    if (P.mat[type].type == MPP_ELASTIC || (enl && !enl[e])) {
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
  for (int t = threadIdx.x; t < GN * NPLANE; t += NT) {
    const int pl = t / GN, l = t % GN;
    const int md = blockIdx.x * GN + l;
    if (md < nrows_out) A[aidx(pl, md)] = s_acc[l * NPLANE + pl];
  }
}

// ------------------------------------------------------------------------------------------------
// Hybrid operator of RVEs with a damage / plastic phase (OP_HYBRID).
//
// The Jacobian row of an interior node depends on u only through the elements around it that are PAST their material's
// linear regime; everywhere else it is the row of the elastic law -- for an elastic element by definition, for a damage
// / plastic element below its threshold because the reference's forward-difference tangent of a linear stress is the
// elastic tangent up to a rounding residue of about 2e-10 (src/material.cpp:49-63).  Measured on the damage50 load path
// (profiles/r03a_damage_fraction.log): until a Gauss point's matrix phase crosses the threshold as a whole, 0 .. 12 %
// of its rows touch a non-linear element.  So: (1) k_probe_lin flags the non-linear elements of every slot at the
// current iterate, (2) k_hyb_list turns them into the compact list of interior nodes that touch one, (3) only those
// rows are assembled (k_elem_ctan / k_asm_mat_general on the list) into the slot's matrix buffer, (4) the SpMV is the
// implicit elastic operator for ALL nodes followed by k_spmv_hyb, which overwrites the listed nodes with their
// explicit rows and corrects p.Ap.  Slots whose list exceeds MICROPP_HYBRID_MAX (default 0.7) of the rows take the
// fully assembled path.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(NT)
    k_probe_lin(const __grid_constant__ MeshConst P, const Lst L, SlotTables T, const double *__restrict__ u_pool,
                size_t vstride, const int *__restrict__ elem_type) {
  const int slot = slot_of(L);
  if (slot < 0) return;
  const int e = blockIdx.x * NT + threadIdx.x;
  if (e >= P.nelem) return;
  const mpp_material m = P.mat[__ldg(&elem_type[e])];
  bool nl = false;
  if (m.type != MPP_ELASTIC) {
    const double *u = u_pool + (size_t)slot * vstride;
    const double *vars = T.vars_old[slot];
    const int ez = e / (P.nex * P.ney);
    const int r = e - ez * P.nex * P.ney;
    const int ey = r / P.nex, ex = r - ey * P.nex;
    double ue[24];
    gather_ue(P, u, ex, ey, ez, ue);
    const int nv = mat_nvar(m.type);
#pragma unroll 1
    for (int gp = 0; gp < 8; ++gp) {
      double eps[6], vbuf[7];
      gp_strain(P.dsh[gp], ue, eps);
      const double *v = fetch_vars(vars, P.nelem_pad, e, gp, nv, vbuf);
      // past the threshold now (evolute's own test), or carrying damage from before (stress = (1 - D_old) sigma_lin:
      // not the elastic tangent even while unloading)
      nl |= mat_evolute(m, eps, v, nullptr);
      if (m.type == MPP_DAMAGE && v && v[1] != 0.0) nl = true;
    }
  }
  T.enl[(size_t)slot * P.nelem_pad + e] = nl ? 1 : 0;
}

// one block per slot: interior nodes that touch a flagged element -> ascending compact list + inverse map + count
__global__ void __launch_bounds__(1024) k_hyb_list(const __grid_constant__ MeshConst P, const Lst L, SlotTables T) {
  const int slot = slot_of(L);
  if (slot < 0) return;
  __shared__ int s_warp[32];
  __shared__ int s_base;
  if (threadIdx.x == 0) s_base = 0;
  __syncthreads();
  const unsigned char *enl = T.enl + (size_t)slot * P.nelem_pad;
  int *hnodes = T.hnodes + (size_t)slot * P.nint_pad, *hpos = T.hpos + (size_t)slot * P.nint_pad;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (int start = 0; start < P.nint; start += 1024) {
    const int m = start + threadIdx.x;
    int flag = 0;
    if (m < P.nint) {
      int i, j, k;
      interior_node(P, m, i, j, k);
#pragma unroll
      for (int c = 0; c < 8; ++c)
        flag |= enl[((k - 1 + (c & 1)) * P.ney + (j - 1 + ((c >> 1) & 1))) * P.nex + (i - 1 + ((c >> 2) & 1))];
    }
    const unsigned bal = __ballot_sync(0xffffffffu, flag);
    if (lane == 0) s_warp[w] = __popc(bal);
    __syncthreads();
    int off = s_base;
    for (int ww = 0; ww < w; ++ww) off += s_warp[ww];
    const int pos = off + __popc(bal & ((1u << lane) - 1u));
    if (m < P.nint) {
      hpos[m] = flag ? pos : -1;
      if (flag) hnodes[pos] = m;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      int tot = 0;
      for (int ww = 0; ww < 32; ++ww) tot += s_warp[ww];
      s_base += tot;
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) T.hcnt[slot] = s_base;
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
      // hybrid: the explicit row of a listed node, the row-block table for every other node
      const int hq = (use_shared == OP_HYBRID && !bnd) ? T.hpos[(size_t)slot * P.nint_pad + m] : -1;
      if (use_shared == OP_GENERIC)
        diag = V.gen[((size_t)n * 3 + d) * 81 + 13 * 3 + d];
      else if (use_shared == OP_HYBRID) {
        if (hq >= 0) diag = A[aidx(13 * 9 + d * 4, hq)];
      } else if (use_shared != OP_IMPLICIT && !bnd)
        diag = A[aidx(13 * 9 + d * 4, m)];
      // src/ell.cpp:73-76 (the implicit operator keeps 1/diag per distinct row block)
      const bool table = !bnd && (use_shared == OP_IMPLICIT || (use_shared == OP_HYBRID && hq < 0));
      const double kk = table ? __ldg(&V.rkinv[__ldg(&V.rowid[m]) * 3 + d]) : 1 / diag;
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


__device__ __forceinline__ double imp_kk(const MeshConst &P, const VecPool &V, int n, int d) {
  int i, j, k;
  node_ijk(P, n, i, j, k);
  if (on_boundary(P, i, j, k)) return 1.0;  // 1 / 1
  const int m = interior_index(P, i, j, k);
  return __ldg(&V.rkinv[__ldg(&V.rowid[m]) * 3 + d]);
}

// Hybrid SpMV, second half: the implicit elastic operator has produced Ap for every interior node; the nodes of the
// slot's list get their explicit row instead (same value order as k_spmv_dot: tile position = list position), and the
// per-block partial sums carry the CORRECTION p.(Ap_explicit - Ap_elastic) so that the fold of all partials is p.Ap.
__global__ void __launch_bounds__(NT)
    k_spmv_hyb(const __grid_constant__ MeshConst P, const Lst L, SlotTables T, VecPool V, int pbase, int force) {
  __shared__ double sm[NRED * (NT / 32)];
  const int slot = slot_of(L);
  if (slot < 0) return;
  mgpu_slot_state *st = &T.state[slot];
  if (!force && !st->cg_active) return;
  const double *A = V.mat + (size_t)slot * V.mstride;
  const size_t vo = (size_t)slot * V.vstride;
  const double *p = V.p + vo;
  double *Ap = V.Ap + vo;
  const int q = blockIdx.x * NT + threadIdx.x;
  double red[1] = {0.0};
  if (q < T.hcnt[slot]) {
    int i, j, k;
    const int n = interior_node(P, T.hnodes[(size_t)slot * P.nint_pad + q], i, j, k);
    double y0 = 0.0, y1 = 0.0, y2 = 0.0;
    const size_t npad = P.nn_pad;
    SpmvLoop<0>::run(A + aidx(0, q), p, npad, n, P.nx, P.nxny, y0, y1, y2);
    const double o0 = Ap[n], o1 = Ap[npad + n], o2 = Ap[2 * npad + n];
    Ap[n] = y0;
    Ap[npad + n] = y1;
    Ap[2 * npad + n] = y2;
    red[0] = p[n] * (y0 - o0) + p[npad + n] * (y1 - o1) + p[2 * npad + n] * (y2 - o2);
  }
  block_sum<1>(red, sm);
  if (threadIdx.x == 0) T.partial[((size_t)slot * NRED + FOLD_PLANE) * T.nblk_max + pbase + blockIdx.x] = red[0];
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

// Slab mode over peer memory: the p update PUSHES the new values of its two boundary planes (the bottom and top OWNED
// planes: local planes 1 and nz - 2) into the neighbours' receive buffers -- remote stores over NVLink, fire and
// forget -- and the LAST of the blocks that write such a plane (ticket in the own mailbox) releases the epoch flag in
// the NEIGHBOURS' mailboxes.  The receiver only waits on its own memory and copies the buffer into its halo plane
// (k_slab_halo_take): no remote load, no round trip on the critical path.  Tickets are taken also when the slot has already left the loop (all ranks run
// the same launch sequence).
__device__ __forceinline__ bool slab_halo_node(const MeshConst &P, int n) {
  const int k = n / P.nxny;
  return (k == 0 && P.halo_lo) || (k == P.nz - 1 && P.halo_hi);
}
__device__ __forceinline__ void slab_push_value(const MeshConst &P, const SlabHalo &H, int n, int d, double v) {
  const int k = n / P.nxny, r = n - k * P.nxny;
  if (k == 1 && H.in_lo) H.in_lo[(size_t)d * P.nxny + r] = v;
  if (k == P.nz - 2 && H.in_hi) H.in_hi[(size_t)d * P.nxny + r] = v;
}
__device__ __forceinline__ void slab_publish_after_update(const MeshConst &P, const SlabHalo &H, SlabMail *mail) {
  const int lo1 = P.nxny / NT, hi1 = (2 * P.nxny - 1) / NT;
  const int lo2 = ((P.nz - 2) * P.nxny) / NT, hi2 = ((P.nz - 1) * P.nxny - 1) / NT;
  const int b = blockIdx.x;
  if (!((b >= lo1 && b <= hi1) || (b >= lo2 && b <= hi2))) return;
  const int overlap = max(0, min(hi1, hi2) - max(lo1, lo2) + 1);
  const unsigned expected = (unsigned)((hi1 - lo1 + 1) + (hi2 - lo2 + 1) - overlap) * gridDim.y;
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned t = atomicAdd(&mail->pub_ticket, 1u);
    if (t == expected - 1) {
      mail->pub_ticket = 0u;
      const unsigned long long epoch = ++mail->p_local;
      __threadfence_system();
      // I am the UPPER neighbour of mail_lo's owner and the LOWER neighbour of mail_hi's owner
      if (H.mail_lo) asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(&H.mail_lo->p_in_epoch[1]), "l"(epoch) : "memory");
      if (H.mail_hi) asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(&H.mail_hi->p_in_epoch[0]), "l"(epoch) : "memory");
    }
  }
}

__global__ void __launch_bounds__(NT)
    k_cg_pupdate_imp(const __grid_constant__ MeshConst P, const Lst L, SlotTables T, VecPool V, SlabMail *publish,
                     const SlabHalo H) {
  const int slot = slot_of(L);
  const int n = blockIdx.x * NT + threadIdx.x;
  if (slot >= 0 && T.state[slot].cg_active && n < P.nn) {
    const mgpu_slot_state *st = &T.state[slot];
    const double beta = st->beta, alpha = st->alpha;
    const size_t vo = (size_t)slot * V.vstride;
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      const size_t ix = vo + (size_t)d * P.nn_pad + n;
      const double z = __dmul_rn(imp_kk(P, V, n, d), V.r[ix]);  // never fused into the FMA below
      const double pp = V.p[ix];
      V.du[ix] = fma(alpha, pp, V.du[ix]);  // x += alpha p of this iteration (src/ell.cpp:102)
      const double pn = z + beta * pp;
      if (!publish || !slab_halo_node(P, n)) V.p[ix] = pn;  // halo entries of p are the neighbour's values
      if (publish) slab_push_value(P, H, n, d, pn);
    }
  }
  if (publish) slab_publish_after_update(P, H, publish);
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
    k_cg_pupdate(const __grid_constant__ MeshConst P, const Lst L, SlotTables T, VecPool V, SlabMail *publish,
                 const SlabHalo H) {
  const int slot = slot_of(L);
  const int n = blockIdx.x * NT + threadIdx.x;
  if (slot >= 0 && T.state[slot].cg_active && n < P.nn) {
    const mgpu_slot_state *st = &T.state[slot];
    const double beta = st->beta, alpha = st->alpha;
    const size_t vo = (size_t)slot * V.vstride;
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      const size_t ix = vo + (size_t)d * P.nn_pad + n;
      const double pp = V.p[ix];
      V.du[ix] = fma(alpha, pp, V.du[ix]);  // x += alpha p of this iteration (src/ell.cpp:102)
      const double pn = V.z[ix] + beta * pp;
      if (!publish || !slab_halo_node(P, n)) V.p[ix] = pn;  // halo entries of p are the neighbour's values
      if (publish) slab_push_value(P, H, n, d, pn);
    }
  }
  if (publish) slab_publish_after_update(P, H, publish);
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
    k_axpy_u(const __grid_constant__ MeshConst P, const Lst L, SlotTables T, VecPool V, unsigned long long *apps,
             int op) {
  const int slot = slot_of(L);
  if (slot < 0) return;
  mgpu_slot_state *st = &T.state[slot];
  if (!st->nr_active) return;
  const size_t vo = (size_t)slot * V.vstride;
  const int n = blockIdx.x * NT + threadIdx.x;
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    st->solver_its += st->cg_its;
    atomicAdd(apps + (op & 7), (unsigned long long)st->cg_its);  // operator applications of this solve (measurement)
    if (op == OP_HYBRID)  // explicit rows those applications streamed
      atomicAdd(apps + 7, (unsigned long long)st->cg_its * (unsigned long long)T.hcnt[slot]);
  }
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
//  * the halo planes of p are PUSHED into the neighbours' receive buffers by the p update itself (k_cg_pupdate*),
//    followed by an epoch flag in the neighbours' mailboxes; the receiver waits on its own memory and copies the
//    buffer into its halo plane (k_slab_halo_take).
// Write-after-read safety needs no extra flag: a rank overwrites p (next p update) only after the tail of the
// following reduction, which cannot complete before every neighbour has posted its partial sum, i.e. has finished
// the SpMV that followed its pull.  Two mailbox buffers (epoch parity) are enough for the same reason.
// All waits are bounded (about 2 s): a lost rank raises SlabMail::error instead of hanging the GPU.
// ------------------------------------------------------------------------------------------------

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

// p as it stands (after cg_init: p = z) -> the neighbours' receive buffers + their epoch flags; one launch per solve
__global__ void __launch_bounds__(NT)
    k_slab_push_p(const __grid_constant__ MeshConst P, const double *p_own, SlabMail *own, const SlabHalo H) {
  // blockIdx.y = 0: bottom owned plane -> lower neighbour, 1: top owned plane -> upper neighbour
  const bool hi = blockIdx.y == 1;
  double *dst = hi ? H.in_hi : H.in_lo;
  if (dst) {
    const double *src = p_own + (size_t)(hi ? P.nz - 2 : 1) * P.nxny;
    for (int i = blockIdx.x * NT + threadIdx.x; i < P.nxny; i += gridDim.x * NT) {
#pragma unroll
      for (int d = 0; d < 3; ++d) dst[(size_t)d * P.nxny + i] = src[(size_t)d * P.nn_pad + i];
    }
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned t = atomicAdd(&own->pub_ticket, 1u);
    if (t == gridDim.x * gridDim.y - 1) {
      own->pub_ticket = 0u;
      const unsigned long long epoch = ++own->p_local;
      __threadfence_system();
      if (H.mail_lo) st_release_sys(&H.mail_lo->p_in_epoch[1], epoch);
      if (H.mail_hi) st_release_sys(&H.mail_hi->p_in_epoch[0], epoch);
    }
  }
}

// wait (on OWN memory) until the neighbour has pushed the plane that belongs to the p update this rank has done itself,
// then copy it from the receive buffer into the halo plane of p.  blockIdx.y = 0: lower halo plane, 1: upper one
__global__ void __launch_bounds__(NT) k_slab_halo_take(const __grid_constant__ MeshConst P, double *p_own, SlabMail *own) {
  const bool hi = blockIdx.y == 1;
  if (hi ? !P.halo_hi : !P.halo_lo) return;
  __shared__ int ok;
  if (threadIdx.x == 0) {
    ok = spin_until(&own->p_in_epoch[hi ? 1 : 0], own->p_local);
    if (!ok) own->error = 1;
  }
  __syncthreads();
  if (!ok) return;
  const double *src = reinterpret_cast<const double *>(reinterpret_cast<const char *>(own) + slab_mail_bytes()) +
                      (size_t)(hi ? 1 : 0) * 3 * P.nxny;
  double *dst = p_own + (size_t)(hi ? P.nz - 1 : 0) * P.nxny;
  for (int i = blockIdx.x * NT + threadIdx.x; i < P.nxny; i += gridDim.x * NT) {
#pragma unroll
    for (int d = 0; d < 3; ++d) dst[(size_t)d * P.nn_pad + i] = __ldcv(src + (size_t)d * P.nxny + i);
  }
}

// Fused cross-rank reduction of the slab mode: ONE warp folds the slab-local partial sums of the preceding kernel
// (kind 2: p.Ap partials of the SpMV; kind 3: z.z / r.z partials of the r update; other kinds: T.red as their ticket
// reductions left it), PUSHES them into every rank's mailbox (one lane per destination: the NVLink stores go out in
// parallel), waits on its own memory for every rank's contribution, adds them in RANK ORDER and runs the scalar tail.
// One launch per reduction, and no remote load on the critical path (a remote poll costs a NVLink round trip per
// rank: 8 ranks x 2 loads in sequence were ~25 us per reduction).
// sum of n partials by the whole block in a fixed order (strided per-thread sums, then a fixed tree): deterministic;
// result valid in thread 0.  A 200^3 slab leaves tens of thousands of partials -- too many for one warp.
__device__ __forceinline__ double block_fold(const double *partial, int n, double *sm /* [blockDim.x / 32] */) {
  double acc = 0.0;
  for (int q = threadIdx.x; q < n; q += blockDim.x) acc += __ldcg(&partial[q]);
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = acc;
  __syncthreads();
  double tot = 0.0;
  if (threadIdx.x == 0)
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) tot += sm[w];
  __syncthreads();
  return tot;
}

__global__ void k_slab_reduce_tail(const __grid_constant__ MeshConst P, const Lst L, SlotTables T,
                                   const __grid_constant__ SlabPeers peers, SlabMail *own, int myrank, int nranks,
                                   int k, int kind, int mode, int nfold) {
  __shared__ double sm[32];
  const int slot = slot_of(L);
  if (slot < 0) return;
  double *red = T.red + slot * 8;
  if (kind == 2 && nfold > 0) {  // nfold == 0: the assembled SpMV's own ticket reduction already left p.Ap in T.red
    const double s = block_fold(T.partial + ((size_t)slot * NRED + FOLD_PLANE) * T.nblk_max, nfold, sm);
    if (threadIdx.x == 0) red[0] = s;
  } else if (kind == 3) {
    const double *partial = T.partial + (size_t)slot * NRED * T.nblk_max;
    const double zz = block_fold(partial, nfold, sm), rz = block_fold(partial + T.nblk_max, nfold, sm);
    if (threadIdx.x == 0) {
      red[0] = zz;
      red[1] = rz;
    }
  }
  __syncthreads();
  if (threadIdx.x >= 32) return;
  // push: lane r writes this rank's sums into rank r's mailbox, then releases its flag there
  const int lane = threadIdx.x;
  unsigned long long epoch = 0;
  if (lane == 0) epoch = ++own->red_local;
  epoch = __shfl_sync(0xffffffffu, epoch, 0);
  if (lane < nranks) {
    SlabMail *dst = peers.mail[lane];
    for (int q = 0; q < k; ++q) dst->in[epoch & 1][myrank][q] = red[q];
    __threadfence_system();
    st_release_sys(&dst->in_epoch[myrank], epoch);
    // wait (on LOCAL memory) for rank `lane`'s contribution
    if (!spin_until(&own->in_epoch[lane], epoch)) own->error = 1;
  }
  __syncwarp();
  if (lane != 0) return;
  double tot[8] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
  for (int r = 0; r < nranks; ++r)  // rank order: identical bits on every rank
    for (int q = 0; q < k; ++q) tot[q] += __ldcv(&own->in[epoch & 1][r][q]);
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
                          SlotTables T, int mode, cudaGraphConditionalHandle cond, int cond_set, int *trips,
                          int hyb_max = 0) {
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
      // 0: Newton-active, 1: CG-active, 2 / 3: Newton-active slots for the hybrid / the fully assembled operator
      if (mode <= 1)
        keep = mode ? st->cg_active : st->nr_active;
      else
        keep = st->nr_active && ((T.hcnt[slot] <= hyb_max) == (mode == 2));
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
namespace {

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
      case 4:  // hybrid SpMV (implicit operator + explicit rows of the listed nodes)
        c->prof_acc[6] += ms;
        c->prof_acc[7] += e.slots;
        break;
      case 5: c->prof_res_ms += ms; break;  // cluster-resident DPCG solves
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

}  // namespace

// row blocks of `codes` (spmv_implicit.cu: implicit_row_ids) -> V.rows / V.rkinv, gathered by the same code as the
// explicit assembly (k_rows_build); the three pure-material blocks also go to the host copy (kernel parameter)
void mgpu_int::build_row_blocks(mgpu_ctx *c, const std::vector<int> &codes) {
  c->nrows = (int)codes.size();
  int *d_codes = nullptr;
  double *d_rows = nullptr, *d_rkinv = nullptr;
  CK(cudaMalloc(&d_codes, sizeof(int) * c->nrows));
  CK(cudaMalloc(&d_rows, sizeof(double) * RB_LEN * c->nrows));
  CK(cudaMalloc(&d_rkinv, sizeof(double) * 3 * c->nrows));
  h2d_sync(c, d_codes, codes.data(), sizeof(int) * c->nrows);
  k_rows_build<<<(c->nrows + NT - 1) / NT, NT, 0, c->stream>>>(d_codes, c->nrows, d_rows, d_rkinv, c->d_ke);
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(c->stream));
  CK(cudaFree(d_codes));
  c->launches++;
  c->V.rows = d_rows;
  c->V.rkinv = d_rkinv;
  memset(&c->pure_rows, 0, sizeof(PureRows));
  CK(cudaMemcpy(c->pure_rows.a, d_rows, sizeof(double) * RB_LEN * std::min(c->nrows, 3), cudaMemcpyDeviceToHost));
}

extern "C" {

int mgpu_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
  return n;
}

// MICROPP_SEGV_TRACE=1: native backtrace on SIGSEGV / SIGABRT (debugging aid; off by default)
static void segv_trace(int sig) {
  void *bt[64];
  const int n = backtrace(bt, 64);
  const char msg[] = "micropp-b200: fatal signal, native backtrace:\n";
  (void)!write(2, msg, sizeof(msg) - 1);
  backtrace_symbols_fd(bt, n, 2);
  _exit(128 + sig);
}
mgpu_ctx *mgpu_create(const mgpu_config *cfg) {
  static bool trace_installed = false;
  if (!trace_installed && getenv("MICROPP_SEGV_TRACE")) {
    trace_installed = true;
    signal(SIGSEGV, segv_trace);
    signal(SIGABRT, segv_trace);
  }
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
  int nblk_max = std::max((P.nn + NT - 1) / NT, (P.nelem + NT - 1) / NT);
  c->implicit = c->all_elastic && cfg->implicit_elastic && P.nint > 0;
  // RVE with a damage / plastic phase: the implicit tables are built as well (elastic law of every material) and the
  // hybrid operator serves the slots that are mostly inside their linear regime.  Needs the TMA kernel (nx even).
  c->hybrid = !c->all_elastic && cfg->implicit_elastic && !cfg->slab && P.nint > 0 && P.nx % 2 == 0;
  if (const char *env = getenv("MICROPP_HYBRID")) c->hybrid = c->hybrid && atoi(env) != 0;
  {
    double frac = 0.7;
    if (const char *env = getenv("MICROPP_HYBRID_MAX")) frac = atof(env);
    c->hyb_max = (int)(frac * P.nint);
  }
  if (c->hybrid) nblk_max += (P.nint + NT - 1) / NT;  // partial sums of k_spmv_hyb behind those of the implicit SpMV
  const size_t per_slot = sizeof(double) * ((c->implicit ? 0 : mlen) + 8 * vlen + (size_t)NRED * nblk_max + 12) + 256 +
                          (c->hybrid ? (size_t)P.nelem_pad + 8 * (size_t)P.nint_pad : 0);
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

  // implicit operator: row-block table, tilings and TMA descriptors of the SpMV kernels (spmv_implicit.cu); may raise
  // nblk_max (the per-slot partial-sum buffer also holds the per-(item, warp) partials of p.Ap)
  if (c->implicit || c->hybrid) {
    implicit_setup(c, cfg, &nblk_max);
    if (c->hybrid && c->imp_kernel != IMP_TMAC) c->hybrid = false;
  }
  // all-elastic RVEs evaluate the residual through the implicit operator (k_u_to_p); MICROPP_RHS_OPERATOR=0 keeps the
  // element loop (k_elem_rhs + k_asm_rhs), which every RVE with a damage / plastic phase and every z-slab uses
  c->rhs_operator = c->implicit && !P.slab;
  if (const char *env = getenv("MICROPP_RHS_OPERATOR")) c->rhs_operator = c->rhs_operator && atoi(env) != 0;

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
  T.enl = nullptr;
  T.hnodes = T.hpos = T.hcnt = nullptr;
  if (c->hybrid) {
    CK(cudaMalloc(&T.enl, (size_t)P.nelem_pad * W));
    CK(cudaMalloc(&T.hnodes, sizeof(int) * (size_t)P.nint_pad * W));
    CK(cudaMalloc(&T.hpos, sizeof(int) * (size_t)P.nint_pad * W));
    CK(cudaMalloc(&T.hcnt, sizeof(int) * W));
    CK(cudaMemset(T.hcnt, 0, sizeof(int) * W));
  }
  for (int l = 0; l < NLIST; ++l) CK(cudaMalloc(&c->d_list[l], sizeof(int) * W));
  CK(cudaMalloc(&c->d_count, sizeof(int)));
  CK(cudaMallocHost(&c->h_count, 8 * sizeof(int)));
  CK(cudaMalloc(&c->d_apps, 8 * sizeof(unsigned long long)));
  CK(cudaMemset(c->d_apps, 0, 8 * sizeof(unsigned long long)));
  CK(cudaMalloc(&c->d_cnt2, 8 * sizeof(int)));
  CK(cudaMemset(c->d_cnt2, 0, 8 * sizeof(int)));

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
  for (auto &kv : c->slab_chunk_graphs) cudaGraphExecDestroy(kv.second);
  if (c->slab_mail) cudaFree(c->slab_mail);
  implicit_destroy(c);
  cudaFree(c->T.state);
  cudaFree((void *)c->T.vars_old);
  cudaFree((void *)c->T.vars_new);
  cudaFree((void *)c->T.u_n);
  cudaFree((void *)c->T.u_k);
  cudaFree(c->T.eps);
  cudaFree(c->T.stress);
  cudaFree(c->T.partial);
  cudaFree(c->T.red);
  if (c->T.enl) cudaFree(c->T.enl);
  if (c->T.hnodes) cudaFree(c->T.hnodes);
  if (c->T.hpos) cudaFree(c->T.hpos);
  if (c->T.hcnt) cudaFree(c->T.hcnt);
  for (int l = 0; l < NLIST; ++l) cudaFree(c->d_list[l]);
  cudaFree(c->d_count);
  cudaFreeHost(c->h_count);
  cudaFree(c->d_cnt2);
  cudaFree(c->d_apps);
  if (c->d_cg_hist) cudaFree(c->d_cg_hist);
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
int mgpu_slot_state_size(void) { return (int)sizeof(mgpu_slot_state); }

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
  if (c->rhs_operator) {  // all-elastic RVE: b = -A u through the implicit operator (k_u_to_p above)
    k_u_to_p<<<node_grid(c, n), NT, 0, c->stream>>>(c->mc, lst_of(c, l), c->T, c->V, mode);
    launch_imp_spmv(c, l, n, 1, c->imp_kernel);
    k_rhs_from_ap<<<node_grid(c, n), NT, 0, c->stream>>>(c->mc, lst_of(c, l), c->T, c->V, mode);
    c->launches += 3;
    CK(cudaGetLastError());
    return;
  }
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
static void asm_mat_impl(mgpu_ctx *c, int l, int n, int to_shared, int hybrid);
void mgpu_asm_mat(mgpu_ctx *c, int l, int n, int to_shared) { asm_mat_impl(c, l, n, to_shared, 0); }
// hybrid operator: only the rows of every slot's node list (mgpu_hybrid_split)
void mgpu_asm_mat_hyb(mgpu_ctx *c, int l, int n) { asm_mat_impl(c, l, n, 0, 1); }
static void asm_mat_impl(mgpu_ctx *c, int l, int n, int to_shared, int hybrid) {
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
      const dim3 ge = elem_grid(c, cnt);
      k_elem_ctan<<<dim3(8 * ge.x, ge.y), NT, 0, c->stream>>>(c->mc, lst, c->T, c->V.u, c->V.vstride, c->d_elem_type,
                                                             c->d_ctan, cstride, hybrid);
      dim3 g(std::max((c->mc.nint + GN - 1) / GN, 1), cnt);
      k_asm_mat_general<<<g, NT, GN * NPLANE * sizeof(double), c->stream>>>(
          c->mc, lst, c->V.mat, c->V.mstride, shared, c->d_elem_type, c->d_ke, c->d_ctan, cstride, c->T, hybrid);
      c->launches += 1;
    }
  }
  CK(cudaGetLastError());
}
int mgpu_implicit(const mgpu_ctx *c) { return c->implicit ? 1 : 0; }
int mgpu_implicit_rows(const mgpu_ctx *c) { return c->nrows; }
int mgpu_implicit_fix_nodes(const mgpu_ctx *c) { return c->nfix; }

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
  if ((use_shared == OP_IMPLICIT && !c->implicit) || (use_shared == OP_HYBRID && !c->hybrid)) {
    fprintf(stderr, "micropp-b200: operator %d requested on a context that does not provide it\n", use_shared);
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
  ProfScope ps(c, use_shared == OP_HYBRID ? 4 : 0, n);
  if (use_shared == OP_IMPLICIT) {
    launch_imp_spmv(c, l, n, 0, c->imp_kernel);
  } else if (use_shared == OP_HYBRID) {
    // the implicit elastic operator on every node, then the explicit rows of the listed nodes + the p.Ap correction
    c->defer_fold = true;
    launch_imp_spmv(c, l, n, 0, c->imp_kernel);
    c->defer_fold = false;
    const dim3 g = int_grid(c, n);
    k_spmv_hyb<<<g, NT, 0, c->stream>>>(c->mc, lst_of(c, l), c->T, c->V, c->last_spmv_nfold, 0);
    launch_fold_spmv(c, l, n, c->last_spmv_nfold + (int)g.x, 0);
    c->launches += 2;
  } else {
    c->last_spmv_nfold = 0;
    k_spmv_dot<<<int_grid(c, n), NT, 0, c->stream>>>(c->mc, lst_of(c, l), c->T, c->V, use_shared, 0);
  }
  CK(cudaGetLastError());
}
void mgpu_cg_update(mgpu_ctx *c, int l, int n) {
  if (n <= 0) return;
  ProfScope ps(c, 3, n);
  if (c->cg_op == OP_IMPLICIT)
    k_cg_update_imp<<<upd_grid(c, n), NT, 0, c->stream>>>(c->mc, lst_of(c, l), c->T, c->V, 0);
  else
    k_cg_update<<<upd_grid(c, n), NT, 0, c->stream>>>(c->mc, lst_of(c, l), c->T, c->V);
  c->last_update_nblk = (int)upd_grid(c, n).x;
  if (!c->slab_fused) {  // fused slab path: k_slab_reduce_tail folds the partials itself
    k_fold_update<<<dim3(1, n), 32, 0, c->stream>>>(c->mc, lst_of(c, l), c->T, c->last_update_nblk);
    c->launches++;
  }
  CK(cudaGetLastError());
}
void mgpu_cg_pupdate(mgpu_ctx *c, int l, int n) {
  if (n <= 0) return;
  ProfScope ps(c, 3, n);
  if (c->cg_op == OP_IMPLICIT)
    k_cg_pupdate_imp<<<node_grid(c, n), NT, 0, c->stream>>>(c->mc, lst_of(c, l), c->T, c->V,
                                                            c->slab_fused ? c->slab_mail : nullptr, c->slab_halo);
  else
    k_cg_pupdate<<<node_grid(c, n), NT, 0, c->stream>>>(c->mc, lst_of(c, l), c->T, c->V,
                                                        c->slab_fused ? c->slab_mail : nullptr, c->slab_halo);
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
  k_axpy_u<<<node_grid(c, n), NT, 0, c->stream>>>(c->mc, lst_of(c, l), c->T, c->V, c->d_apps, c->cg_op);
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
  // lists 1 / 6 (Newton) and 2 / 7 (CG) keep their length on the device too: the step graphs start from it
  int *count2 = list_out == 1 ? c->d_cnt2 : list_out == 2 ? c->d_cnt2 + 1 : list_out == 6 ? c->d_cnt2 + 4 :
                list_out == 7 ? c->d_cnt2 + 5 : nullptr;
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

mgpu_ctx::StepGraph build_step_graph(mgpu_ctx *c, int B, int use_shared, int ls) {
  // list set: 0 = lists 1 (Newton) / 2 (CG) with counters d_cnt2[0..2], 1 = lists 6 / 7 with d_cnt2[4..6]
  const int LN = ls ? 6 : 1, LC = ls ? 7 : 2;
  int *cnt = c->d_cnt2 + 4 * ls;
  const bool prof = c->prof;
  c->prof = false;  // no event records inside a capture
  const unsigned long long l0 = c->launches;
  cudaGraph_t g;
  CK(cudaGraphCreate(&g, 0));
  const bool resident = use_shared == OP_IMPLICIT && c->res != nullptr;
  cudaGraphConditionalHandle cond = cudaGraphConditionalHandle();
  // (a handle that no conditional node uses makes cudaGraphInstantiate fail: only the looping graphs create one)
  if (!resident) CK(cudaGraphConditionalHandleCreate(&cond, g, 0, cudaGraphCondAssignDefault));

  // head: Jacobian, CG start, CG list := Newton-list slots whose loop-head test says "iterate"
  capture_begin(c, g, nullptr, 0);
  c->dyn_count = cnt;
  if (use_shared == OP_SLOT) mgpu_asm_mat(c, LN, B, 0);
  if (use_shared == OP_HYBRID) mgpu_asm_mat_hyb(c, LN, B);
  if (resident) {
    // the whole DPCG solve of every slot is one cluster-resident kernel: no loop node, the tail follows in stream order
    mgpu_cg_resident(c, LN, B);
    mgpu_axpy_u(c, LN, B);
    mgpu_asm_rhs(c, LN, B, 1);
    k_compact<<<1, 1024, 0, c->stream>>>(c->d_list[LN], 0, cnt, c->d_list[LN], cnt, nullptr, c->T, 0, cond, 0, nullptr);
    c->launches++;
    CK(cudaMemcpyAsync(c->h_count + 4 * ls, cnt, 3 * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    capture_end(c);
    mgpu_ctx::StepGraph sgr;
    sgr.fixed_launches = (int)(c->launches - l0);
    sgr.body_launches = 0;
    c->dyn_count = nullptr;
    c->launches = l0;  // captured, not executed
    c->prof = prof;
    CK(cudaGraphInstantiate(&sgr.exec, g, 0));
    CK(cudaGraphDestroy(g));
    return sgr;
  }
  mgpu_cg_init(c, LN, B, use_shared);
  k_compact<<<1, 1024, 0, c->stream>>>(c->d_list[LN], 0, cnt, c->d_list[LC], cnt + 1, nullptr, c->T, 1, cond, 1,
                                       nullptr);
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
  c->dyn_count = cnt + 1;
  mgpu_cg_spmv_dot(c, LC, B, use_shared);
  mgpu_cg_update(c, LC, B);
  mgpu_cg_pupdate(c, LC, B);
  k_compact<<<1, 1024, 0, c->stream>>>(c->d_list[LC], 0, cnt + 1, c->d_list[LC], cnt + 1, nullptr, c->T, 1, cond, 1,
                                       cnt + 2);
  c->launches++;
  capture_end(c);
  const int body_l = (int)(c->launches - l1);

  // tail: update, residual + Newton test, Newton list compaction, list lengths to the host
  const unsigned long long l2 = c->launches;
  capture_begin(c, g, &wnode, 1);
  c->dyn_count = cnt;
  mgpu_cg_finish(c, LN, B);
  mgpu_axpy_u(c, LN, B);
  mgpu_asm_rhs(c, LN, B, 1);
  k_compact<<<1, 1024, 0, c->stream>>>(c->d_list[LN], 0, cnt, c->d_list[LN], cnt, nullptr, c->T, 0, cond, 0, nullptr);
  c->launches++;
  CK(cudaMemcpyAsync(c->h_count + 4 * ls, cnt, 3 * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
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

extern "C" int mgpu_newton_step_graph_on(mgpu_ctx *c, int ls, int n_active, int use_shared) {
  if (n_active <= 0) return 0;
  CK(cudaSetDevice(c->device));
  int B = 1;
  while (B < n_active) B <<= 1;
  B = std::min(B, c->W);
  const long long key = ((long long)B * 8 + use_shared) * 2 + ls;
  auto it = c->step_graphs.find(key);
  if (it == c->step_graphs.end()) it = c->step_graphs.emplace(key, build_step_graph(c, B, use_shared, ls)).first;
  CK(cudaMemsetAsync(c->d_cnt2 + 4 * ls + 2, 0, sizeof(int), c->stream));
  CK(cudaGraphLaunch(it->second.exec, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  c->launches += it->second.fixed_launches + (unsigned long long)it->second.body_launches * c->h_count[4 * ls + 2];
  return c->h_count[4 * ls];
}
extern "C" int mgpu_newton_step_graph(mgpu_ctx *c, int n_active, int use_shared) {
  return mgpu_newton_step_graph_on(c, 0, n_active, use_shared);
}

// ---- hybrid operator: probe + split ------------------------------------------------------------------------
extern "C" int mgpu_hybrid_available(const mgpu_ctx *c) { return c->hybrid ? 1 : 0; }
// Flags the non-linear elements of the first n slots of list l at their current iterate, builds the per-slot node
// lists, and splits the slots into list 6 (hybrid operator: at most hyb_max listed rows) and list 1 (fully assembled
// operator).  l may be 1 (compacted in place).  Syncs.
extern "C" void mgpu_hybrid_split(mgpu_ctx *c, int l, int n, int *n_hybrid, int *n_full) {
  *n_hybrid = *n_full = 0;
  if (n <= 0) return;
  CK(cudaSetDevice(c->device));
  {
    ProfScope ps(c, 1, n);
    k_probe_lin<<<elem_grid(c, n), NT, 0, c->stream>>>(c->mc, lst_of(c, l), c->T, c->V.u, c->V.vstride, c->d_elem_type);
    k_hyb_list<<<dim3(1, n), 1024, 0, c->stream>>>(c->mc, lst_of(c, l), c->T);
  }
  k_compact<<<1, 1024, 0, c->stream>>>(c->d_list[l], n, nullptr, c->d_list[6], c->d_count, c->d_cnt2 + 4, c->T, 2,
                                       cudaGraphConditionalHandle(), 0, nullptr, c->hyb_max);
  CK(cudaMemcpyAsync(c->h_count + 3, c->d_count, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  k_compact<<<1, 1024, 0, c->stream>>>(c->d_list[l], n, nullptr, c->d_list[1], c->d_count, c->d_cnt2, c->T, 3,
                                       cudaGraphConditionalHandle(), 0, nullptr, c->hyb_max);
  CK(cudaMemcpyAsync(c->h_count + 7, c->d_count, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(c->stream));
  c->launches += 3;
  *n_hybrid = c->h_count[3];
  *n_full = c->h_count[7];
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
    // the mailbox and, in the same allocation (one IPC handle), the two receive buffers of the halo planes
    const size_t bytes = slab_mail_bytes() + sizeof(double) * 2 * 3 * (size_t)c->mc.nxny;
    CK(cudaMalloc((void **)&c->slab_mail, bytes));
    CK(cudaMemset(c->slab_mail, 0, bytes));
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
void mgpu_slab_link(mgpu_ctx *c, int rank, int size, void *const *mails) {
  if (size > SLAB_MAX_RANKS) {
    fprintf(stderr, "micropp-b200: at most %d slabs\n", SLAB_MAX_RANKS);
    abort();
  }
  mgpu_slab_mail(c);
  c->slab_rank = rank;
  c->slab_size = size;
  for (int r = 0; r < size; ++r) c->slab_peers.mail[r] = (SlabMail *)mails[r];
  SlabHalo &H = c->slab_halo;
  H.mail_lo = rank > 0 ? c->slab_peers.mail[rank - 1] : nullptr;
  H.mail_hi = rank + 1 < size ? c->slab_peers.mail[rank + 1] : nullptr;
  auto buf = [&](SlabMail *m, int side) {
    return reinterpret_cast<double *>(reinterpret_cast<char *>(m) + slab_mail_bytes()) + (size_t)side * 3 * c->mc.nxny;
  };
  H.in_lo = H.mail_lo ? buf(H.mail_lo, 1) : nullptr;  // I am the lower neighbour's UPPER neighbour
  H.in_hi = H.mail_hi ? buf(H.mail_hi, 0) : nullptr;
}
// p (as it stands) -> the neighbours' halo planes; once per solve, after cg_init
void mgpu_slab_push_p(mgpu_ctx *c) {
  c->launches++;
  const int nb = std::max(1, (c->mc.nxny + NT - 1) / NT);
  k_slab_push_p<<<dim3(nb, 2), NT, 0, c->stream>>>(c->mc, c->V.p, c->slab_mail, c->slab_halo);
  CK(cudaGetLastError());
}
void mgpu_slab_halo_take(mgpu_ctx *c) {
  if (!c->mc.halo_lo && !c->mc.halo_hi) return;
  c->launches++;
  const int nb = std::max(1, (c->mc.nxny + NT - 1) / NT);
  k_slab_halo_take<<<dim3(nb, 2), NT, 0, c->stream>>>(c->mc, c->V.p, c->slab_mail);
  CK(cudaGetLastError());
}
// fused path: on / off (slab_host.cpp switches it on for the whole life of a slab context)
void mgpu_slab_set_fused(mgpu_ctx *c, int on) { c->slab_fused = on != 0; }
// fold + post + gather + tail of the reduction that the preceding kernel started (kind / mode as mgpu_tail)
void mgpu_slab_reduce_tail(mgpu_ctx *c, int l, int k, int kind, int mode) {
  c->launches++;
  const int nfold = kind == 2 ? c->last_spmv_nfold : (kind == 3 ? c->last_update_nblk : 0);
  const int threads = nfold > 4096 ? 1024 : (nfold > 256 ? 256 : 32);
  k_slab_reduce_tail<<<dim3(1, 1), threads, 0, c->stream>>>(c->mc, lst_of(c, l), c->T, c->slab_peers, c->slab_mail,
                                                           c->slab_rank, c->slab_size, k, kind, mode, nfold);
  CK(cudaGetLastError());
}
int mgpu_slab_error(mgpu_ctx *c) {
  CK(cudaSetDevice(c->device));
  CK(cudaStreamSynchronize(c->stream));
  SlabMail m;
  CK(cudaMemcpy(&m, c->slab_mail, sizeof(m), cudaMemcpyDeviceToHost));
  return m.error;
}
// One DPCG iteration of this rank's slab (every rank calls it; no host synchronisation):
//   halo take -> SpMV -> reduce_tail(p.Ap) -> r, z update -> reduce_tail(z.z, r.z) -> p, x update (+ push of its boundary planes)
void mgpu_slab_cg_iteration(mgpu_ctx *c, int l, int op) {
  mgpu_slab_halo_take(c);
  mgpu_cg_spmv_dot(c, l, 1, op);
  mgpu_slab_reduce_tail(c, l, 1, 2, 0);
  mgpu_cg_update(c, l, 1);
  mgpu_slab_reduce_tail(c, l, 2, 3, 0);
  mgpu_cg_pupdate(c, l, 1);
}
// `iters` DPCG iterations of this rank's slab as ONE CUDA graph launch (captured once per (op, iters); the epochs of
// the cross-rank flags live on the device, so every kernel argument is constant).  Slots that converge inside the
// chunk skip their kernels (cg_active), exactly as in the single-domain chunked loop.
void mgpu_slab_cg_chunk(mgpu_ctx *c, int l, int op, int iters) {
  CK(cudaSetDevice(c->device));
  const int key = (c->slab_fused ? 4096 : 0) + op * 1024 + iters;
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
void mgpu_cg_history(mgpu_ctx *c, int k) {
  CK(cudaStreamSynchronize(c->stream));
  if (c->d_cg_hist) CK(cudaFree(c->d_cg_hist));
  c->d_cg_hist = nullptr;
  c->cg_hist_k = k > 0 ? k : 0;
  if (k > 0) {
    CK(cudaMalloc(&c->d_cg_hist, sizeof(double) * (size_t)k * c->W));
    CK(cudaMemset(c->d_cg_hist, 0, sizeof(double) * (size_t)k * c->W));
  }
  // the pointer lives in the slot state (device memory read at run time), so captured graphs need no rebuild
  for (int s = 0; s < c->W; ++s) {
    double *ptr = k > 0 ? c->d_cg_hist + (size_t)s * k : nullptr;
    CK(cudaMemcpy(&c->T.state[s].cg_hist, &ptr, sizeof(ptr), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(&c->T.state[s].cg_hist_k, &c->cg_hist_k, sizeof(int), cudaMemcpyHostToDevice));
  }
}
int mgpu_cg_history_read(mgpu_ctx *c, int slot, double *out, int k) {
  if (!c->d_cg_hist || slot < 0 || slot >= c->W) return 0;
  CK(cudaStreamSynchronize(c->stream));
  mgpu_slot_state st;
  CK(cudaMemcpy(&st, c->T.state + slot, sizeof(st), cudaMemcpyDeviceToHost));
  const int n = std::min(std::min(k, c->cg_hist_k), st.cg_its + 1);
  if (n > 0) CK(cudaMemcpy(out, c->d_cg_hist + (size_t)slot * c->cg_hist_k, sizeof(double) * n, cudaMemcpyDeviceToHost));
  return n;
}
void mgpu_prof_enable(mgpu_ctx *c, int on) {
  prof_drain(c);
  c->prof = on != 0;
}
void mgpu_prof_read(mgpu_ctx *c, double *out9, int reset) {
  double *out8 = out9;
  prof_drain(c);
  for (int i = 0; i < 8; ++i) out8[i] = c->prof_acc[i];
  // RVE applications of the operators, counted on the device (the DPCG iterations every slot really did)
  unsigned long long apps[8];
  CK(cudaStreamSynchronize(c->stream));
  CK(cudaMemcpy(apps, c->d_apps, sizeof(apps), cudaMemcpyDeviceToHost));
  out8[2] = (double)(apps[OP_SLOT] + apps[OP_SHARED] + apps[OP_GENERIC] + apps[OP_IMPLICIT]);
  out8[7] = (double)apps[OP_HYBRID];
  out9[8] = (double)apps[7];  // sum over hybrid applications of the explicit rows each one streamed
  if (reset) {
    CK(cudaMemset(c->d_apps, 0, sizeof(apps)));
    CK(cudaDeviceSynchronize());
  }
  if (reset)
    for (int i = 0; i < 8; ++i) c->prof_acc[i] = 0;
}
// accumulated time of the cluster-resident DPCG solves since the last reset (profiling mode)
double mgpu_prof_resident_ms(mgpu_ctx *c, int reset) {
  prof_drain(c);
  const double ms = c->prof_res_ms;
  if (reset) c->prof_res_ms = 0;
  return ms;
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

}  // extern "C"
