// spmv_implicit.cu -- the implicit operator of all-elastic RVEs: Ap = A p fused with p.Ap WITHOUT a per-RVE matrix.
//
// Every material elastic => the Jacobian does not depend on u (src/material.cpp:84-94): it is the same for every macro
// Gauss point and every Newton step, and the ELL row block of an interior node (src/ell-common.cpp:166-198 applied to
// the element matrices of src/assembly.cpp:141-178) is a pure function of the materials of the node's 8 elements.  Only
// the DISTINCT row blocks are kept (rows[id][27][10], a few hundred KB per context; ids 0..2 = nodes surrounded by one
// material) plus one id per interior node.  Per RVE application the SpMV then moves 48 B/node (p in, Ap out) instead
// of 1992 B/node, and the bound moves from HBM to the FP64 pipe.
//
// Kernels (DESIGN.md section 5 has the measurements):
//   k_spmv_dot_tmac  default (nx even).  The nodes of a chunk (7 or 8 x-adjacent nodes of one grid row) that share the
//                    chunk's majority pure-material row block: the p brick of a tile arrives by ONE TMA load
//                    (cp.async.bulk.tensor + mbarrier), the row block is a __grid_constant__ kernel parameter (its values
//                    reach the DFMAs as uniform registers), and the tile of Ap leaves by ONE TMA store from shared memory
//                    (the per-lane stores of 8 / 16-byte pieces cost 27 % of the kernel: profiles/r02i).
//   k_spmv_fix       every other interior node (material interfaces; minority nodes of a chunk): thread = (node, 8 slots),
//                    row block from the L1/L2-resident table, all slots of a group share the fetch.
//   k_spmv_dot_imp   every node through the table, no TMA: used when nx is odd (TMA needs 16-B global strides; said on
//                    stderr at context creation) and as the yardstick of the parity tests.
// All three add the 243 terms of a row in the order of the assembled k_spmv_dot (and of src/ell.cpp:35-44): Ap is
// bit-identical to the assembled path.  p.Ap: per-(tile, warp) / per-block partial sums in plane FOLD_PLANE of the
// slot's partial buffer, folded in a fixed order by k_fold_spmv (one warp per slot) -- deterministic.
#include "mgpu_internal.cuh"

using namespace mgpu_int;

namespace {

// ------------------------------------------------------------------------------------------------
// k_spmv_dot_imp: a thread owns one interior node, fetches the node's row block from the table of distinct row blocks
// (L1-resident: almost every node of a warp uses the same block, so the 243 loads are broadcasts) and applies it to
// R right-hand sides (slots) at once.  Bound by L1 wavefronts (81 p loads per node and slot).
// ------------------------------------------------------------------------------------------------
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


// ------------------------------------------------------------------------------------------------
// k_spmv_dot_tmac
//   tile   = CB chunks of TN nodes in x  x  32 grid rows (lane shape 0: 8 y x 4 z, shape 1: 4 y x 8 z -- the strip that
//            covers a remainder of 1..4 y rows); item = (tile, slot); a block walks over ntl tiles x rs slots
//   brick  = the p values the tile needs: (pitch, 10, 6) or (pitch, 6, 10) nodes x 3 components, ONE
//            cp.async.bulk.tensor.5d load from the pool viewed as (x, y, z, component, slot); out-of-grid parts are
//            zero-filled by the TMA unit.  pitch is even with an odd number of 16-B units: the 128-bit loads of the 8
//            rows of a quarter-warp fall into 8 different bank groups.
//   thread = TN nodes of one row: 9 neighbour rows x 3 components x 5 LDS.128, 243 * TN DFMAs whose row-block operand is
//            a uniform register (LDCU from the kernel parameter); the material of the chunk selects one of three
//            compile-time copies of the loop (lanes in different materials take their copies one after the other)
//   Ap     = written back through the SAME shared memory: once every warp has left the brick, the threads lay their
//            TN x 3 results out as the dense (CB*TN - 2, 8, 4, 3) / (.., 4, 8, 3) box and one thread issues ONE TMA store
//            (16-B aligned origin: the first and last node of a tile row are stored by their own lanes).
//            Nodes the chunk does not keep (k_spmv_fix serves them afterwards, in stream order) and grid positions
//            outside the interior are written as zeros (boundary entries of Ap are zero by construction).
// One stage per block; 4 resident blocks per SM hide the load -> compute -> store chain of each other.
// ------------------------------------------------------------------------------------------------
// x pitch of the p brick: even (rows stay 16-B aligned) with an odd number of 16-B units
__host__ __device__ constexpr int tmac_pitch(int tn, int cb) {
  int p = (tn * cb + 2 + (tn & 1) + 1) & ~1;  // odd tn: the brick may start one node early (16-B aligned TMA origin)
  if (((p / 2) & 1) == 0) p += 2;
  return p;
}

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


__device__ __forceinline__ void tma_store_5d(const CUtensorMap *tmap, const void *smem_src, int c0, int c1, int c2, int c3,
                                             int c4) {
  asm volatile("cp.async.bulk.tensor.5d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5, %6}], [%1];\n" ::"l"(tmap),
               "r"((unsigned)__cvta_generic_to_shared(smem_src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
               : "memory");
}

template <int CB, int TN>
__global__ void __launch_bounds__(32 * CB, 4)
    k_spmv_dot_tmac(const __grid_constant__ MeshConst P, const Lst L, int n_list, SlotTables T, VecPool V, TileInfo2 ti,
                    const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                    const __grid_constant__ CUtensorMap smap_a, const __grid_constant__ CUtensorMap smap_b,
                    const __grid_constant__ PureRows R, int ntl, int rs, int force) {
  extern __shared__ unsigned char s_raw[];
  constexpr int pitch = tmac_pitch(TN, CB);
  constexpr int BRICK = 3 * BRICK_ROWS * pitch;  // doubles
  // Ap box of a tile: the TMA unit wants a 16-B aligned global origin, and a tile starts at an odd node (x = 1 + k CB TN),
  // so the box holds the CB*TN - 2 middle nodes of a row and the two end nodes are stored by their own lanes
  constexpr int PXO = CB * TN - 2;
  static_assert((CB * TN) % 2 == 0 && 3 * 32 * PXO <= BRICK, "Ap box must be a 16-B multiple wide and fit in the brick");
  __shared__ uint64_t s_full;
  __shared__ int s_slot[TMA_MAX_RS];
  unsigned char *s_base = s_raw + ((128u - ((unsigned)__cvta_generic_to_shared(s_raw) & 127u)) & 127u);
  double *s_brick = reinterpret_cast<double *>(s_base);
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
    mbar_init(&s_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
  }
  __syncthreads();

  auto next_active = [&](int i) {
    while (i < nitems && s_slot[i % rs] < 0) ++i;
    return i;
  };
  auto issue = [&](int item) {  // one thread
    const int4 td = __ldg(&ti.tiles[tile0 + item / rs]);
    mbar_expect_tx(&s_full, BRICK * 8);
    // the box starts at an even x (16-B aligned global address): with TN = 7 odd tile origins start one node early
    tma_load_5d(s_brick, td.w ? &tmap_b : &tmap_a, &s_full, (td.x * TN) & ~1, td.y, td.z, 0, s_slot[item % rs]);
  };

  int cur = next_active(0), k = 0;
  if (threadIdx.x == 0 && cur < nitems) issue(cur);
  // per-tile thread state (independent of the slot: with ntl == 1 it is computed once per block)
  int cur_tile = -1, pure = 0, keep = 0, ry = 0, rz = 0, by_rows = TILE_Y + 2, by_out = TILE_Y;
  int xoff = 0;  // 1: the brick starts one node before the tile (see issue())
  int4 td = make_int4(0, 0, 0, 0);
  bool work = false;
  while (cur < nitems) {
    const int nxt = next_active(cur + 1);
    const int tile = tile0 + cur / rs, slot = s_slot[cur % rs];
    if (tile != cur_tile) {
      cur_tile = tile;
      td = __ldg(&ti.tiles[tile]);
      // lane shape 0: 8 y-rows x 4 z-rows per warp (brick 10 x 6 rows); 1: 4 y-rows x 8 z-rows (brick 6 x 10 rows)
      ry = td.w ? (lane & 3) : (lane & 7);
      rz = td.w ? (lane >> 2) : (lane >> 3);
      by_rows = td.w ? TILE_Z + 2 : TILE_Y + 2;
      by_out = td.w ? TILE_Z : TILE_Y;
      const int c = td.x + w, jj = td.y + ry, kk = td.z + rz;
      work = c < ti.nchunk && jj < P.niy && kk < P.niz;
      const int info = work ? __ldg(&ti.chunk_pure[(kk * P.niy + jj) * ti.nchunk + c]) : 0;
      pure = info & 0xff;
      const int nvalid = min(TN, P.nix - c * TN);
      keep = work ? (((1 << nvalid) - 1) & ~(info >> 8)) : 0;  // nodes this thread produces itself
      xoff = (td.x * TN) & 1;
    }
    mbar_wait(&s_full, k & 1);

    double red0 = 0.0, red1 = 0.0, red2 = 0.0;
    double acc[8][3];
#pragma unroll
    for (int t = 0; t < 8; ++t) acc[t][0] = acc[t][1] = acc[t][2] = 0.0;
    if (keep) {
      const int xs = TN * w + xoff, bx0 = xs & ~1;
      if ((TN & 1) && (xs & 1)) {  // odd window start: loads begin one double earlier (warp-uniform)
        if (pure == 0)
          tile_rows_apply_const<0, 1, pitch, TN, TN & 1>(R, s_brick, bx0, ry, rz, by_rows, acc);
        else if (pure == 1)
          tile_rows_apply_const<1, 1, pitch, TN, TN & 1>(R, s_brick, bx0, ry, rz, by_rows, acc);
        else
          tile_rows_apply_const<2, 1, pitch, TN, TN & 1>(R, s_brick, bx0, ry, rz, by_rows, acc);
      } else {
        if (pure == 0)
          tile_rows_apply_const<0, 1, pitch, TN, 0>(R, s_brick, bx0, ry, rz, by_rows, acc);
        else if (pure == 1)
          tile_rows_apply_const<1, 1, pitch, TN, 0>(R, s_brick, bx0, ry, rz, by_rows, acc);
        else
          tile_rows_apply_const<2, 1, pitch, TN, 0>(R, s_brick, bx0, ry, rz, by_rows, acc);
      }
      const int cbase = ((rz + 1) * by_rows + (ry + 1)) * pitch + xs + 1;
#pragma unroll
      for (int t = 0; t < TN; ++t) {
        if ((keep >> t) & 1) {
          red0 += s_brick[cbase + t] * acc[t][0];
          red1 += s_brick[BRICK_ROWS * pitch + cbase + t] * acc[t][1];
          red2 += s_brick[2 * BRICK_ROWS * pitch + cbase + t] * acc[t][2];
        } else {
          acc[t][0] = acc[t][1] = acc[t][2] = 0.0;  // not this thread's node: a zero goes out (k_spmv_fix overwrites it)
        }
      }
    }
    // ---- Ap: through the brick's shared memory, one TMA store per item ----
    __syncthreads();  // every warp has left the brick
    {
      double *o = s_brick + (size_t)(rz * by_out + ry) * PXO + TN * w - 1;  // box column of node t: TN*w + t - 1
#pragma unroll
      for (int d = 0; d < 3; ++d)
#pragma unroll
        for (int t = 0; t < TN; ++t)
          if ((t > 0 || w > 0) && (t < TN - 1 || w < CB - 1)) o[d * 32 * PXO + t] = acc[t][d];
    }
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");  // generic-proxy writes -> visible to the TMA unit
    __syncthreads();
    if (threadIdx.x == 0) {
      tma_store_5d(td.w ? &smap_b : &smap_a, s_brick, td.x * TN + 2, td.y + 1, td.z + 1, 0, slot);
      asm volatile("cp.async.bulk.commit_group;\n" ::: "memory");
      if (nxt < nitems) {
        asm volatile("cp.async.bulk.wait_group.read 0;\n" ::: "memory");  // the store has read the tile: reuse the memory
        issue(nxt);
      }
    }
    // the two end nodes of the row (outside the 16-B aligned box)
    if (work) {
      double *Ap = V.Ap + (size_t)slot * V.vstride + (size_t)(td.z + rz + 1) * P.nxny + (td.y + ry + 1) * P.nx +
                   (td.x + w) * TN + 1;
      if (w == 0) {
#pragma unroll
        for (int d = 0; d < 3; ++d) Ap[d * P.nn_pad] = acc[0][d];
      }
      if (w == CB - 1 && (td.x + w) * TN + TN - 1 < P.nix) {
#pragma unroll
        for (int d = 0; d < 3; ++d) Ap[d * P.nn_pad + TN - 1] = acc[TN - 1][d];
      }
    }
    // p.Ap of this slot: one partial per (tile, warp) in plane FOLD_PLANE of the slot's partial-sum buffer, folded in
    // a fixed order by k_fold_spmv (one warp per slot), which the kernel boundary orders after these stores
    const double red = warp_sum((red0 + red1) + red2);
    if (lane == 0) T.partial[((size_t)slot * NRED + FOLD_PLANE) * T.nblk_max + tile * CB + w] = red;
    cur = nxt;
    ++k;
  }
  if (threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group 0;\n" ::: "memory");
}

// ------------------------------------------------------------------------------------------------
// k_spmv_fix: the interface nodes (row block != a pure-material one).  thread = (node of the list, R slots): the row
// block comes from the table (L1/L2; every slot shares it, so one fetch serves R right-hand sides), p is gathered from
// global memory.  All 243 terms in the order of k_spmv_dot (bit-identical to the assembled path).  Per-block partial
// sums of p.Ap go to entries [pbase + blockIdx.x] of the slot's FOLD_PLANE.
// ------------------------------------------------------------------------------------------------
template <int R>
__global__ void __launch_bounds__(NT, 3)
    k_spmv_fix(const __grid_constant__ MeshConst P, const Lst L, int n_list, SlotTables T, VecPool V,
               const int2 *__restrict__ fixn, int nfix, int pbase, int force) {
  __shared__ double sm[R * (NT / 32)];
  __shared__ int s_slot[R];
  if (threadIdx.x < R) {
    const int yy = (int)blockIdx.y * R + (int)threadIdx.x;
    const int cnt = L.dcount ? min(*L.dcount - L.yoff, n_list) : n_list;
    int slot = yy < cnt ? L.list[yy + L.yoff] : -1;
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

  const int f = blockIdx.x * NT + threadIdx.x;
  double red[R];
#pragma unroll
  for (int r = 0; r < R; ++r) red[r] = 0.0;
  if (f < nfix) {
    const int2 e = __ldg(&fixn[f]);
    const int n = e.x;
    const size_t npad = P.nn_pad;
    double y[R][3];
#pragma unroll
    for (int r = 0; r < R; ++r) y[r][0] = y[r][1] = y[r][2] = 0.0;
    const double *a = V.rows + (size_t)e.y * RB_LEN;
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
  block_sum<R>(red, sm);
  if (threadIdx.x == 0) {
#pragma unroll
    for (int r = 0; r < R; ++r)
      if (s_slot[r] >= 0)
        T.partial[((size_t)s_slot[r] * NRED + FOLD_PLANE) * T.nblk_max + pbase + blockIdx.x] = red[r];
  }
}



}  // namespace

// ================================================================================================
// host side
// ================================================================================================
namespace {

typedef void (*tmac_kernel_t)(const MeshConst, const Lst, int, SlotTables, VecPool, TileInfo2, const CUtensorMap,
                              const CUtensorMap, const CUtensorMap, const CUtensorMap, const PureRows, int, int, int);
inline tmac_kernel_t tmac_kernel(int cb, int tn) {
  if (tn == 7) return cb == 4 ? k_spmv_dot_tmac<4, 7> : k_spmv_dot_tmac<2, 7>;
  switch (cb) {
    case 1: return k_spmv_dot_tmac<1, 8>;
    case 2: return k_spmv_dot_tmac<2, 8>;
    case 3: return k_spmv_dot_tmac<3, 8>;
    default: return k_spmv_dot_tmac<4, 8>;
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


// Tiling of k_spmv_dot_tmac: nodes per thread (7 or 8; 7 only with an even number of warps, so that the Ap box is a
// 16-B multiple wide), warps per block, tile descriptors (two lane shapes), the pure row block + keep mask of every
// chunk, and the nodes no chunk keeps (the list of k_spmv_fix).  Pure host code (also reachable from the CPU tests).
struct TmacTiling {
  int tn = 8, cb = 1, nchunk = 0, pitch = 0;
  std::vector<int4> tiles;      // x: first chunk, y / z: interior coordinates of the tile origin, w: lane shape
  std::vector<int> chunk_pure;  // [niz][niy][nchunk]: majority pure id | (TN-bit mask of the nodes NOT kept) << 8
  std::vector<int2> fixn;       // nodes not kept by their chunk: x node id, y row-block id
};
static TmacTiling tmac_tiling(int nx, int ny, int nix, int niy, int niz, const std::vector<int> &rowid) {
  TmacTiling t2;
  long best = -1;
  for (int tn = 8; tn >= 7; --tn)
    for (int cb = 4; cb >= 1; --cb) {
      if (tn == 7 && (cb & 1)) continue;
      const int nch = (nix + tn - 1) / tn;
      if (cb > nch && !(tn == 8 && cb == 1)) continue;
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
  const int tiles_x = (t2.nchunk + t2.cb - 1) / t2.cb;
  for (int z0 = 0; z0 < niz; z0 += TILE_Z)
    for (int y0 = 0; y0 < y_a; y0 += TILE_Y)
      for (int tx = 0; tx < tiles_x; ++tx) t2.tiles.push_back(make_int4(tx * t2.cb, y0, z0, 0));
  if (y_a < niy)
    for (int z0 = 0; z0 < niz; z0 += TILE_Y)
      for (int tx = 0; tx < tiles_x; ++tx) t2.tiles.push_back(make_int4(tx * t2.cb, y_a, z0, 1));
  t2.chunk_pure.assign((size_t)niz * niy * t2.nchunk, 0);
  for (int kk = 0; kk < niz; ++kk)
    for (int jj = 0; jj < niy; ++jj)
      for (int cc = 0; cc < t2.nchunk; ++cc) {
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
            t2.fixn.push_back(make_int2((kk + 1) * nx * ny + (jj + 1) * nx + cc * TN + t + 1, rowid[m0 + t]));
          }
        t2.chunk_pure[((size_t)kk * niy + jj) * t2.nchunk + cc] = pure | (mask << 8);
      }
  return t2;
}

typedef CUresult (*tmap_encode_t)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// a pool vector (p or Ap) as a rank-5 tensor (x, y, z, component, slot) with the given box; the encoder comes from the
// driver through the runtime (no -lcuda)
bool encode_pool_map(mgpu_ctx *c, CUtensorMap *out, const double *base, const cuuint32_t box[5]) {
  static tmap_encode_t fn = nullptr;
  if (!fn) {
    void *f = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &qres) != cudaSuccess || !f ||
        qres != cudaDriverEntryPointSuccess)
      return false;
    fn = (tmap_encode_t)f;
  }
  const MeshConst &P = c->mc;
  const cuuint64_t gdim[5] = {(cuuint64_t)P.nx, (cuuint64_t)P.ny, (cuuint64_t)P.nz, 3, (cuuint64_t)c->W};
  const cuuint64_t gstr[4] = {(cuuint64_t)P.nx * 8, (cuuint64_t)P.nxny * 8, (cuuint64_t)P.nn_pad * 8,
                              (cuuint64_t)c->V.vstride * 8};
  const cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  return fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 5, (void *)base, gdim, gstr, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <class T>
T *to_device(mgpu_ctx *c, const std::vector<T> &v) {
  T *d = nullptr;
  CK(cudaMalloc(&d, sizeof(T) * std::max<size_t>(v.size(), 1)));
  if (!v.empty()) h2d_sync(c, d, v.data(), sizeof(T) * v.size());
  return d;
}

}  // namespace

void mgpu_int::implicit_setup(mgpu_ctx *c, const mgpu_config *cfg, int *nblk_max) {
  const MeshConst &P = c->mc;
  VecPool &V = c->V;
  // distinct row blocks: code = sum_c type_c 3^c over the 8 elements around an interior node
  std::vector<int> codes, rowid;
  implicit_row_ids(P.nix, P.niy, P.niz, P.nint_pad, cfg->elem_type, codes, rowid);
  build_row_blocks(c, codes);
  V.rowid = to_device(c, rowid);
  c->nfix = 0;
  c->imp_kernel = IMP_SIMPLE;
  resident_setup(c, cfg, rowid.data());  // whole solves inside one cluster when the RVE fits (cg_resident.cu)
  if (P.nx % 2 != 0) {
    // TMA needs 16-B global strides (nx even).  Said out loud, not chosen silently: an odd nx runs the table-driven
    // kernel, which is correct but L1-bound (1.12 ms instead of 0.6 ms per application of 1024 RVEs at 30^3)
    fprintf(stderr, "micropp-b200: nx = %d is odd: the implicit operator runs k_spmv_dot_imp (no TMA: 16-B row strides "
                    "need an even nx); an even nx gets k_spmv_dot_tmac\n", P.nx);
    return;
  }
  TileInfo2 &t2 = c->tile2;
  const TmacTiling tt = tmac_tiling(P.nx, P.ny, P.nix, P.niy, P.niz, rowid);
  t2.tn = tt.tn;
  t2.cb = tt.cb;
  t2.nchunk = tt.nchunk;
  t2.pitch = tt.pitch;
  t2.ntiles = (int)tt.tiles.size();
  c->tile2_smem = (int)(sizeof(double) * 3 * BRICK_ROWS * t2.pitch);
  c->nfix = (int)tt.fixn.size();
  c->d_tiles2 = to_device(c, tt.tiles);
  c->d_chunk_pure2 = to_device(c, tt.chunk_pure);
  c->d_fixn = to_device(c, tt.fixn);
  t2.tiles = c->d_tiles2;
  t2.chunk_pure = c->d_chunk_pure2;
  const cuuint32_t pxo = (cuuint32_t)(t2.cb * t2.tn - 2);  // the middle nodes of a tile row (see k_spmv_dot_tmac)
  const cuuint32_t box_a[5] = {(cuuint32_t)t2.pitch, TILE_Y + 2, TILE_Z + 2, 3, 1};
  const cuuint32_t box_b[5] = {(cuuint32_t)t2.pitch, TILE_Z + 2, TILE_Y + 2, 3, 1};
  const cuuint32_t out_a[5] = {pxo, TILE_Y, TILE_Z, 3, 1};
  const cuuint32_t out_b[5] = {pxo, TILE_Z, TILE_Y, 3, 1};
  if (encode_pool_map(c, &c->tmap_a, V.p, box_a) && encode_pool_map(c, &c->tmap_b, V.p, box_b) &&
      encode_pool_map(c, &c->smap_a, V.Ap, out_a) && encode_pool_map(c, &c->smap_b, V.Ap, out_b)) {
    CK(cudaFuncSetAttribute(tmac_kernel(t2.cb, t2.tn), cudaFuncAttributeMaxDynamicSharedMemorySize,
                            (c->tile2_smem + 127) / 128 * 128 + 128));
    *nblk_max = std::max(*nblk_max, t2.ntiles * t2.cb + (c->nfix + NT - 1) / NT);
    c->imp_kernel = IMP_TMAC;
  } else {
    fprintf(stderr, "micropp-b200: cuTensorMapEncodeTiled failed; the implicit operator runs k_spmv_dot_imp\n");
    t2.ntiles = 0;
  }
  if (const char *env = getenv("MICROPP_IMP_KERNEL"))  // parity tests: the table-driven kernel on request
    if (atoi(env) == IMP_SIMPLE) c->imp_kernel = IMP_SIMPLE;
}

void mgpu_int::implicit_destroy(mgpu_ctx *c) {
  resident_destroy(c);
  if (c->V.rows) cudaFree((void *)c->V.rows);
  if (c->V.rkinv) cudaFree((void *)c->V.rkinv);
  if (c->V.rowid) cudaFree((void *)c->V.rowid);
  if (c->d_tiles2) cudaFree(c->d_tiles2);
  if (c->d_chunk_pure2) cudaFree(c->d_chunk_pure2);
  if (c->d_fixn) cudaFree(c->d_fixn);
  c->V.rows = c->V.rkinv = nullptr;
  c->V.rowid = nullptr;
}

// Ap = A p of the implicit operator over n entries of list l
void mgpu_int::launch_imp_spmv(mgpu_ctx *c, int l, int n, int force, int kern) {
  if (kern == IMP_TMAC && c->tile2.ntiles == 0) kern = IMP_SIMPLE;  // odd nx (announced at context creation)
  if (kern == IMP_TMAC) {
    const TileInfo2 &t2 = c->tile2;
    int rs = std::min(TMA_MAX_RS, n), ntl = 1;
    // small groups of slots: fewer slots per block so that the grid still fills 148 SMs
    const long want_blocks = 148L * 4 * 2;
    rs = (int)std::max(1L, std::min((long)rs, (long)n * t2.ntiles / want_blocks));
    if (n < 4) {  // one (or a few) large RVEs, e.g. a z-slab: several tiles per block, but at least ~4 waves of blocks
      rs = n;
      ntl = (int)std::max(1L, std::min((long)(TMA_MAX_RS / rs), (long)n * t2.ntiles / (148L * 4 * 4)));
    }
    const dim3 grid((t2.ntiles + ntl - 1) / ntl, (n + rs - 1) / rs);
    const int smem = (c->tile2_smem + 127) / 128 * 128 + 128;
    tmac_kernel(t2.cb, t2.tn)<<<grid, 32 * t2.cb, smem, c->stream>>>(c->mc, lst_of(c, l), n, c->T, c->V, t2, c->tmap_a,
                                                                    c->tmap_b, c->smap_a, c->smap_b, c->pure_rows, ntl,
                                                                    rs, force);
    const int nfb = (c->nfix + NT - 1) / NT;
    if (nfb > 0) {  // the nodes no chunk keeps (stream order: after the zeros the tile stores put there)
      // few interface nodes x few slots: 2 slots per thread instead of 8, so that the grid still covers the SMs (the
      // kernel is a latency chain of 243 gathers per slot; the sums of a slot do not depend on the grouping)
      if ((long)nfb * ((n + MR - 1) / MR) < 148L * 3)
        k_spmv_fix<2><<<dim3(nfb, (n + 1) / 2), NT, 0, c->stream>>>(c->mc, lst_of(c, l), n, c->T, c->V, c->d_fixn,
                                                                    c->nfix, t2.ntiles * t2.cb, force);
      else
        k_spmv_fix<MR><<<dim3(nfb, (n + MR - 1) / MR), NT, 0, c->stream>>>(c->mc, lst_of(c, l), n, c->T, c->V, c->d_fixn,
                                                                          c->nfix, t2.ntiles * t2.cb, force);
      c->launches++;
    }
    // p.Ap: one warp per slot folds the per-(tile, warp) and per-block partials in a fixed order + the scalar tail
    // (fused slab path: k_slab_reduce_tail does the fold together with the cross-rank sum)
    c->last_spmv_nfold = t2.ntiles * t2.cb + nfb;
    if ((!c->slab_fused && !c->defer_fold) || force) {
      k_fold_spmv<<<dim3(1, n), 32, 0, c->stream>>>(c->mc, lst_of(c, l), c->T, c->last_spmv_nfold, force);
      c->launches++;
    }
  } else {
    k_spmv_dot_imp<MR><<<int_grid(c, (n + MR - 1) / MR), NT, 0, c->stream>>>(c->mc, lst_of(c, l), n, c->T, c->V, force);
  }
}

// the p.Ap fold on its own (hybrid operator: after k_spmv_hyb has added its correction partials)
void mgpu_int::launch_fold_spmv(mgpu_ctx *c, int l, int n, int nfold, int force) {
  k_fold_spmv<<<dim3(1, n), 32, 0, c->stream>>>(c->mc, lst_of(c, l), c->T, nfold, force);
}

extern "C" {

// Host-only view of the tiling of k_spmv_dot_tmac / k_spmv_fix for an nx x ny x nz RVE (tests/test_tiling.py, no GPU).
// meta[6] = {nodes per thread, warps per block, chunks per x row, brick pitch, tiles, nodes of k_spmv_fix}; the arrays
// may be null (size query): rowid[(nx-2)(ny-2)(nz-2)], tiles[4 * ntiles], chunk_pure[niz * niy * nchunk],
// fix_nodes[2 * nfix] (node id, row-block id).  Returns the number of distinct row blocks.
int mgpu_tmac_tiling_host(int nx, int ny, int nz, const int *elem_type, int *meta, int *rowid_out, int *tiles,
                          int *chunk_pure, int *fix_nodes) {
  const int nix = nx - 2, niy = ny - 2, niz = nz - 2;
  if (nix < 1 || niy < 1 || niz < 1) return 0;
  std::vector<int> codes, rowid;
  implicit_row_ids(nix, niy, niz, nix * niy * niz, elem_type, codes, rowid);
  const TmacTiling t = tmac_tiling(nx, ny, nix, niy, niz, rowid);
  if (meta) {
    meta[0] = t.tn;
    meta[1] = t.cb;
    meta[2] = t.nchunk;
    meta[3] = t.pitch;
    meta[4] = (int)t.tiles.size();
    meta[5] = (int)t.fixn.size();
  }
  if (rowid_out) memcpy(rowid_out, rowid.data(), sizeof(int) * (size_t)nix * niy * niz);
  if (tiles) memcpy(tiles, t.tiles.data(), sizeof(int4) * t.tiles.size());
  if (chunk_pure) memcpy(chunk_pure, t.chunk_pure.data(), sizeof(int) * t.chunk_pure.size());
  if (fix_nodes && !t.fixn.empty()) memcpy(fix_nodes, t.fixn.data(), sizeof(int2) * t.fixn.size());
  return (int)codes.size();
}

// -1: no implicit operator; else the SpMV kernel it runs: 0 k_spmv_dot_imp (table-driven), 3 k_spmv_dot_tmac
int mgpu_implicit_kernel(const mgpu_ctx *c) { return c->implicit ? c->imp_kernel : -1; }

// isolated micro-benchmark of the implicit-operator SpMV (+ k_spmv_fix + the p.Ap fold) on the first n slots; kern: -1
// the context's kernel, else 0 / 3 as above.  p as it stands in the pool.
float mgpu_bench_imp_spmv(mgpu_ctx *c, int n, int iters, int kern) {
  CK(cudaSetDevice(c->device));
  if (!c->implicit) return -1.f;
  if (kern < 0) kern = c->imp_kernel;
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
  c->launches += 3 * (iters + 2);
  return ms / iters;
}

}  // extern "C"
