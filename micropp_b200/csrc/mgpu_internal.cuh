// mgpu_internal.cuh -- internal to micropp_b200/csrc/*.cu: the device-side data structures, the small device helpers
// every kernel file uses, and the host-side context behind the opaque `mgpu_ctx` of include/mgpu.h.
//
//   mgpu_kernels.cu    context, assembly, materials, DPCG vector kernels, Newton-step CUDA graph, the C ABI
//   spmv_implicit.cu   the implicit elastic operator: row-block table, tiling, k_spmv_dot_tmac / k_spmv_fix (+ the
//                      table-driven kernel for odd nx)
//   ell_generic.cu     the reference's ELL API for non-hot-path shapes (2-D, other field counts)
#pragma once
#include <cuda.h>  // CUtensorMap (type + enums only; the encoder is fetched through cudaGetDriverEntryPoint)
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <vector>

#include <cstdint>
#include <string>

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

namespace mgpu_int {

constexpr int NT = 128;      // threads per block of node/element kernels
constexpr int NPLANE = 243;  // 27 neighbours x 3 x 3
// row blocks of the implicit operator: [27 neighbours][10] doubles, the 3x3 block of a neighbour in the first 9 --
// 80-B groups are 16-B aligned, so a neighbour's block is five 128-bit loads (global or shared)
constexpr int RB_NBR = 10, RB_LEN = 27 * RB_NBR;
constexpr int NLIST = 8;  // 0 caller, 1 Newton, 2 / 4 CG, 3 caller subset, 5 bench, 6 Newton (hybrid set), 7 CG (hybrid set)
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
  // hybrid operator (RVEs with a damage / plastic phase, see OP_HYBRID): per slot the elements that are past their
  // material's linear regime, the interior nodes that touch one (compact list + inverse map) and their number
  unsigned char *enl;  // [W][nelem_pad]
  int *hnodes;         // [W][nint_pad]  interior-node indices, ascending
  int *hpos;           // [W][nint_pad]  position in hnodes, or -1
  int *hcnt;           // [W]
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
// OP_HYBRID: the implicit elastic row blocks for every node whose 8 elements are all in their linear regime, explicit
// ELL rows (compact list, same plane-major tile layout) only for the nodes that touch a non-linear element
enum { OP_SLOT = 0, OP_SHARED = 1, OP_GENERIC = 2, OP_IMPLICIT = 3, OP_HYBRID = 4 };

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
  if (st->cg_hist && st->cg_hist_k > 0) st->cg_hist[0] = pn;
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
  if (st->cg_hist && its < st->cg_hist_k) st->cg_hist[its] = pn;
  st->cg_active = (its < P.cg_max_its) && !(pn < P.cg_abs_tol || pn < st->pnorm0 * P.cg_rel_tol);
}


constexpr int FOLD_PLANE = 2;  // planes 0/1 of the partial-sum buffer belong to the consumer's own grid_sum<2>

// sum of the n per-(tile, warp) partials of a slot, by ONE warp, in a fixed order; result in every lane
__device__ __forceinline__ double fold_partials(const double *partial, int n) {
  double acc = 0.0;
  for (int q = threadIdx.x & 31; q < n; q += 32) acc += __ldcg(&partial[q]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  return acc;
}

// right-hand sides (slots) per thread of the kernels that fetch a row block per node (k_spmv_dot_imp, k_spmv_fix)
constexpr int MR = 8;
constexpr int UPD_VPT = 2;  // nodes per thread of k_cg_update / k_cg_update_imp

// ---- mbarrier / TMA (cp.async.bulk.tensor) ------------------------------------------------------------------
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
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"((unsigned)__cvta_generic_to_shared(bar)) : "memory");
}

// the three pure-material row blocks, passed BY VALUE as a __grid_constant__ kernel parameter: a row-block value is
// then an operand of the DFMA itself (uniform register filled from the constant bank: no LSU traffic, no vector
// registers)
struct PureRows {
  double a[3 * RB_LEN];
};

// tiling of k_spmv_dot_tmac
constexpr int TILE_Y = 8, TILE_Z = 4, BRICK_ROWS = (TILE_Y + 2) * (TILE_Z + 2);
struct TileInfo2 {
  int cb, tn, nchunk, pitch, ntiles;
  const int4 *tiles;      // [ntiles] x: first chunk, y / z: interior coordinates of the tile origin, w: lane shape
  const int *chunk_pure;  // [niz][niy][nchunk]: majority pure id | (TN-bit mask of the nodes the chunk does NOT keep) << 8
};

// ---- slab mode over peer memory ---------------------------------------------------------------------------
constexpr int SLAB_MAX_RANKS = 16;
// Mailbox of one rank, mapped by every rank (CUDA IPC).  Cross-rank sums are PUSHED: rank s writes its slab-local sums
// into in[parity][s][*] of EVERY rank's mailbox (remote stores over NVLink, one lane per destination), then releases
// in_epoch[s]; a rank waits on its OWN memory only and adds the contributions in rank order.  Two buffers (epoch parity)
// suffice: a rank can post epoch e + 2 only after reduction e + 1 completed, which needs every rank's post of e + 1,
// i.e. every rank has finished reading epoch e.
struct SlabMail {
  double in[2][SLAB_MAX_RANKS][8];
  unsigned long long in_epoch[SLAB_MAX_RANKS];  // epoch of the last post of rank s (written by s)
  unsigned long long p_in_epoch[2];             // epoch of the halo plane PUSHED by the lower [0] / upper [1] neighbour
  unsigned long long red_local, p_local;  // the owner's own counters: the epochs live on the device, so a whole chunk
                                          // of DPCG iterations is a replayable CUDA graph with constant arguments
  int error;
  unsigned pub_ticket;  // blocks of the p update that have finished (the last one publishes the epoch)
};

struct SlabPeers {
  SlabMail *mail[SLAB_MAX_RANKS];
};
// where this rank's boundary planes of p go: the neighbours' RECEIVE buffers (peer-mapped; they follow the SlabMail in
// the same allocation: [side][component][nx*ny], side 0 = from the lower neighbour, 1 = from the upper one) and the
// neighbours' mailboxes for the epoch flags.  The receiver copies a buffer into its halo plane of p in its own stream
// order (k_slab_halo_take) -- pushing into p itself would race with the receiver's own use of the halo values.
struct SlabHalo {
  double *in_lo, *in_hi;  // the lower neighbour's side-1 buffer / the upper neighbour's side-0 buffer (null at the ends)
  SlabMail *mail_lo, *mail_hi;
};
__host__ __device__ inline size_t slab_mail_bytes() { return (sizeof(SlabMail) + 255) / 256 * 256; }

}  // namespace mgpu_int

namespace mgpu_int {
struct ResState;  // cg_resident.cu: plan + device tables of the cluster-resident DPCG
}

struct mgpu_ctx {
  // (types of namespace mgpu_int)
  mgpu_int::ResState *res = nullptr;  // non-null: DPCG solves of the implicit operator run as ONE cluster-resident kernel
  double prof_res_ms = 0;             // profiling: time of those solves

  int device = 0;
  cudaStream_t stream = nullptr;
  mgpu_int::MeshConst mc;
  int ngp = 0, W = 0;
  bool all_elastic = true;
  bool implicit = false;  // all-elastic RVE served by the implicit operator (no per-slot matrices)
  bool rhs_operator = false;  // all-elastic RVE: assembly_rhs as b = -A u through the implicit operator
  bool hybrid = false;    // RVE with a damage / plastic phase: implicit tables built too, OP_HYBRID available
  int hyb_max = 0;        // a slot takes the hybrid operator while its node list holds at most this many rows
  bool defer_fold = false;  // launch_imp_spmv leaves the p.Ap fold to the caller (hybrid SpMV adds its correction first)
  int mat_slots = 0;      // slots of the explicit matrix pool
  int cg_op = mgpu_int::OP_SLOT;    // operator of the DPCG solve in flight (set by mgpu_cg_init)
  int nrows = 0;
  int nfix = 0;           // interior nodes whose row block is not a pure-material one (interface nodes)
  // SpMV kernel of the implicit operator: IMP_TMAC (TMA-tiled, nx even) or IMP_SIMPLE (table-driven; odd nx, announced)
  int imp_kernel = 0;
  mgpu_int::PureRows pure_rows;     // host copy of row blocks 0..2 (kernel parameter of the TMA kernels)
  // slab mode over peer memory (mgpu_slab_link)
  mgpu_int::SlabMail *slab_mail = nullptr;
  mgpu_int::SlabPeers slab_peers{};
  int slab_rank = -1, slab_size = 0;
  std::map<int, cudaGraphExec_t> slab_chunk_graphs;  // key: op * 1024 + iterations
  // fused slab path (slab_host.cpp): the reducing kernels leave their partial sums, k_slab_reduce_tail folds them,
  // posts / gathers the cross-rank sum and runs the scalar tail in ONE launch; the p update publishes its epoch itself
  bool slab_fused = false;
  int last_spmv_nfold = 0, last_update_nblk = 0;
  int slab_launches_per_chunk = 0;
  mgpu_int::SlabHalo slab_halo{};
  mgpu_int::TileInfo2 tile2;        // tiling of k_spmv_dot_tmac (7 or 8 nodes per thread, two lane shapes)
  CUtensorMap tmap_a, tmap_b;  // V.p with the boxes of lane shape 0 (pitch x 10 x 6) and 1 (pitch x 6 x 10)
  CUtensorMap smap_a, smap_b;  // V.Ap with the store boxes (cb*tn x 8 x 4) and (cb*tn x 4 x 8)
  int tile2_smem = 0;     // bytes of one brick
  int4 *d_tiles2 = nullptr;
  int *d_chunk_pure2 = nullptr;
  int2 *d_fixn = nullptr;  // the nfix nodes no chunk keeps: (node id, row-block id), served by k_spmv_fix
  int *d_elem_type = nullptr;
  double *d_ke = nullptr;
  double *d_be = nullptr;    // element residual scratch of assembly_rhs: [be_chunk][24][nelem_pad]
  int be_chunk = 0;
  double *d_ctan = nullptr;  // tangent scratch of the general Jacobian assembly: [ctan_chunk][288][nelem_pad]
  int ctan_chunk = 0;
  mgpu_int::VecPool V{};
  mgpu_int::SlotTables T{};
  int *d_list[mgpu_int::NLIST] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  int *d_count = nullptr;
  double *d_cg_hist = nullptr;           // [W][cg_hist_k] (mgpu_cg_history)
  int cg_hist_k = 0;
  unsigned long long *d_apps = nullptr;  // [8] DPCG iterations done per operator (measurement)
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
  double prof_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  cudaEvent_t t0 = nullptr, t1 = nullptr;
  unsigned long long launches = 0;
  struct StepGraph {
    cudaGraphExec_t exec;
    int fixed_launches, body_launches;
  };
  std::map<long long, StepGraph> step_graphs;  // key = (bucket * 8 + operator) * 2 + list set
};

namespace mgpu_int {

inline Lst lst_of(const mgpu_ctx *c, int l, int off = 0) { return Lst{c->d_list[l], c->dyn_count, off}; }
inline dim3 int_grid(const mgpu_ctx *c, int n) { return dim3(std::max((c->mc.nint + NT - 1) / NT, 1), n, 1); }
inline dim3 node_grid(const mgpu_ctx *c, int n) { return dim3((c->mc.nn + NT - 1) / NT, n, 1); }
inline dim3 upd_grid(const mgpu_ctx *c, int n) { return dim3((c->mc.nn + NT * UPD_VPT - 1) / (NT * UPD_VPT), n, 1); }
inline dim3 elem_grid(const mgpu_ctx *c, int n) { return dim3((c->mc.nelem + NT - 1) / NT, n, 1); }

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

// ---- implicit operator (spmv_implicit.cu) ----
void implicit_setup(mgpu_ctx *c, const mgpu_config *cfg, int *nblk_max);
void implicit_destroy(mgpu_ctx *c);
// Ap = A p (+ per-slot p.Ap and its scalar tail) over the first n entries of list l; kern: the context's kernel
// (c->imp_kernel) or an explicit IMP_* id (parity tests / A-B measurements); force: apply to inactive slots too
void launch_imp_spmv(mgpu_ctx *c, int l, int n, int force, int kern);
void launch_fold_spmv(mgpu_ctx *c, int l, int n, int nfold, int force);
void build_row_blocks(mgpu_ctx *c, const std::vector<int> &codes);  // mgpu_kernels.cu (k_rows_build)
enum { IMP_SIMPLE = 0, IMP_TMAC = 3 };

// ---- cluster-resident DPCG (cg_resident.cu) ----
void resident_setup(mgpu_ctx *c, const mgpu_config *cfg, const int *rowid_host);
void resident_destroy(mgpu_ctx *c);
void launch_cg_resident(mgpu_ctx *c, int l, int n);

}  // namespace mgpu_int
