// cg_resident.cu -- the WHOLE DPCG solve (src/ell.cpp:66-122) of an all-elastic RVE as ONE kernel launch whose working
// set never leaves the SMs: one thread-block CLUSTER per RVE.
//
// The interior (y, z) rows of the RVE are split py x pz over the CTAs of a cluster (<= 8).  A CTA keeps, for the whole
// solve,
//   p    its rows plus a one-row halo ring, in shared memory (the "brick": [3 components][(nyl+2)(nzl+2) rows + 1][pitch],
//        odd row pitch + skewed plane pitch: 16 consecutive rows fall into 16 different banks); boundary nodes of the RVE
//        are zeros that are never written (b and x0 vanish there: identity rows)
//   du   its rows in shared memory,  r  in REGISTERS (a thread owns TN x-adjacent nodes of one row for the whole solve;
//        blocks of more than 12 warps keep the third component of r in shared memory),
//   Ap   only ever exists as the thread's accumulators.
// Per iteration:
//   * the operator: the pure-material row block of the chunk as uniform-register DFMA operands (kernel parameter, as in
//     k_spmv_dot_tmac), 153 of its 243 terms when the row blocks are mirror-symmetric (res_rows_sparse); nodes next to a
//     material interface receive the element-wise correction sum_e (Ke_type(e) - Ke_m) p_e from three difference
//     tables that are a kernel parameter too (res_fix_piece), computed in pieces that level the SM quarters;
//   * two dot-product reductions (p.Ap; z.z and r.z): warp shuffle -> per-warp partials -> one per-CTA pair sent into the
//     mailbox of every CTA -> every thread adds the per-CTA partials in rank order: identical bits in every CTA
//     (identical loop decisions), identical for every slot;
//   * the halo rows of the new p are PUSHED into the neighbour CTAs' bricks.
//   Mailboxes and halo rows travel as st.async ... mbarrier::complete_tx through distributed shared memory; a CTA waits
//   on its OWN mbarriers: there is no cluster barrier and no cluster-scope fence inside the loop.
// HBM traffic of a solve: b in, du out (48 B per node), instead of 248 B per node and ITERATION of the three-kernel
// loop.  The scalar logic (alpha, beta, loop-head test, iteration count, residual history) is the reference's, as in
// tail_cg_init / tail_spmv / tail_cg_update.
//
// Availability: all-elastic RVE (implicit operator), not a z-slab, and a decomposition whose brick + du + tables fit
// the 227 KB of one SM (30^3: 8 CTAs of 14 x 7 rows, 222 KB).  Larger RVEs keep the three-kernel loop (k_spmv_dot_tmac +
// k_cg_update_imp + k_cg_pupdate_imp); said on stderr at context creation when MICROPP_VERBOSE is set.
// Measurements, the optimisation log and the dead ends: profiles/r04_resident_dpcg.md; DESIGN.md section 5.1b.
#include <mutex>

#include "mgpu_internal.cuh"

using namespace mgpu_int;

namespace mgpu_int {

constexpr int RES_MAX_CS = 8;
constexpr int RES_SMEM_LIMIT = 232448;  // 227 KB: the opt-in dynamic shared memory of one sm_100 CTA
constexpr int RES_MAX_PAIRS = 3;        // unordered material pairs {a < b}: u = a + b - 1, table D_u = Ke_b - Ke_a
constexpr int RES_MAX_GROUPS = 16;      // interface entries per CTA <= RES_GROUP * RES_MAX_GROUPS (2 pieces per group: 32 bits)
constexpr int RES_GROUP = 32;           // entries of an interface group: one per lane
constexpr int RES_DLEN = 8 * 8 * 9;     // doubles of one difference table: [element position][element node][3x3]

// the difference tables (Ke_t - Ke_m) of the pairs, passed BY VALUE as a __grid_constant__ kernel parameter like the
// pure row blocks: their values reach the DFMAs of the interface pass as uniform registers (no shared-memory traffic)
struct ResDtab {
  double d[RES_MAX_PAIRS][RES_DLEN];
};

struct ResGeom {
  int cs, py, pz;      // CTAs per cluster = py * pz; rank = qz * py + qy
  int tn, nchunk;      // nodes per thread (x-adjacent), chunks per row
  int nyb, nzb;        // rows per brick plane (max owned y rows + 2), brick planes (max owned z planes + 2)
  int pitch, zp, cstride;  // doubles: x pitch of a brick row (odd: 16 consecutive rows fall into 16 different 8-B
                       // banks), pitch of a brick plane (rows + a skew that keeps the bank sequence going from the last own
                       // row of a plane to the first own row of the next), distance between the components of the brick
  int nrows_max, dpitch, dstride;  // du: [3][nrows_max][dpitch]
  int nthreads, nwarps;
  int fixcap;          // interface-node entries per CTA (max over the ranks)
  int ntask[RES_MAX_CS], taskcap;  // halo rows to push after every p update
  int halo_in[RES_MAX_CS];         // bytes of halo rows a CTA receives per p update (transaction count of its mbarrier)
  int y0[RES_MAX_CS], y1[RES_MAX_CS], z0[RES_MAX_CS], z1[RES_MAX_CS];  // owned INTERIOR (0-based) row ranges
  int nfix[RES_MAX_CS];
  int off_du, off_r2, off_task, off_time, off_mbar, off_fixe, off_fixk, off_fixout, off_red, off_wp, smem_bytes;
  double rkp[3][3];    // 1 / diagonal of the three pure-material row blocks (the Jacobi preconditioner of most nodes)
  const int4 *tinfo;   // [cs][nthreads]  x: ly | lz << 8 | chunk << 16 | material << 24 (-1: idle thread)
                       //                 y: valid-node mask | interface-node mask << 8, z: first interface entry,
                       //                 w: own row (index into du)
  const int4 *fixe;    // [cs][fixcap]    x: ly | lz << 8 | i << 16, y: 8 codes of 3 bits, one per element position (0: element
                       //                 of the chunk's material m; else (1 + u) | neg << 2 for an element of type t: u the
                       //                 pair {m, t}, neg = t < m, i.e. the correction is -D_u), z: row-block id,
                       //                 w: position in owner order (the list itself is in bank-aware work order)
  const unsigned *wpiece;  // [cs][16]       interface pieces of every warp: bit 2 g + h = element positions 4 h .. 4 h + 3 of group g
  const int2 *task;    // [cs][taskcap]   x: brick offset of the own row (node i = 1, component 0), y: destination rank
                       //                 << 24 | brick offset of its halo copy there
};

MPP_HD int res_split(int n, int parts, int q) { return (int)((long long)q * n / parts); }

// Which neighbour CTAs hold a halo copy of own interior row (jj, kk) of `rank`, and at which brick row.  At most 3
// because every part is at least 2 rows wide in a direction that is split.
MPP_HD int res_push_dests(const ResGeom &G, int rank, int jj, int kk, int (&drank)[3], int (&doff)[3]) {
  const int qy = rank % G.py, qz = rank / G.py;
  int n = 0;
  for (int dz = -1; dz <= 1; ++dz)
    for (int dy = -1; dy <= 1; ++dy) {
      if (!dy && !dz) continue;
      const int ny_ = qy + dy, nz_ = qz + dz;
      if (ny_ < 0 || ny_ >= G.py || nz_ < 0 || nz_ >= G.pz) continue;
      if (dy == -1 && jj != G.y0[rank]) continue;
      if (dy == 1 && jj != G.y1[rank] - 1) continue;
      if (dz == -1 && kk != G.z0[rank]) continue;
      if (dz == 1 && kk != G.z1[rank] - 1) continue;
      const int nr = nz_ * G.py + ny_;
      if (n < 3) {
        drank[n] = nr;
        doff[n] = (kk - G.z0[nr] + 1) * G.zp + (jj - G.y0[nr] + 1) * G.pitch;  // of node i = 0
      }
      ++n;
    }
  return n;
}

// Entry q = (c * 8 + jn) * 9 + fi * 3 + fj of the difference table of the pair m < t: (Ke_t - Ke_m)[rows of the node
// that sits at corner (1-ax, 1-ay, 1-az) of element position c = (ax, ay, az)][columns of element node jn]
// (element position and corner as in gather_block_elastic).
MPP_HD double res_dtab_entry(const double *ke, int m, int t, int q) {
  const int c = q / 72, jn = (q / 9) % 8, w = q % 9;
  const int ax = (c >> 2) & 1, ay = (c >> 1) & 1, az = c & 1;
  const int a = corner_of(1 - ax, 1 - ay, 1 - az), fi = w / 3, fj = w - fi * 3;
  const int o = (a * 3 + fi) * 24 + jn * 3 + fj;
  return ke[t * 576 + o] - ke[m * 576 + o];
}

// Correction of (A_m p) at a node whose 8 elements are not all of the chunk's material m: the sum over the elements of
// another material t of (Ke_t - Ke_m)[rows of the node] . p(nodes of the element).  `no` = offset of the node inside
// component 0 of the brick; codes: 3 bits per element position (see ResGeom::fixe).
// Plain version (host replay); the kernel's res_fix_group adds the same terms pair by pair with uniform operands.
inline void res_fix_corr_host(const double *D, const double *brick, int no, int pitch, int zp, int cstride,
                              unsigned codes, double (&y)[3]) {
  y[0] = y[1] = y[2] = 0.0;
  for (int c = 0; c < 8; ++c) {
    const int pc = (int)((codes >> (3 * c)) & 7u);
    if (pc == 0) continue;
    const int ax = (c >> 2) & 1, ay = (c >> 1) & 1, az = c & 1;
    const double sg = (pc & 4) ? -1.0 : 1.0;
    const double *Dc = D + (size_t)((pc & 3) - 1) * RES_DLEN + c * 72;
    const int eo = no + (az - 1) * zp + (ay - 1) * pitch + (ax - 1);  // first node of the element
    for (int jn = 0; jn < 8; ++jn) {
      const int q = eo + corner_z(jn) * zp + corner_y(jn) * pitch + corner_x(jn);
      for (int fj = 0; fj < 3; ++fj)
        for (int fi = 0; fi < 3; ++fi) y[fi] += sg * Dc[jn * 9 + fi * 3 + fj] * brick[fj * cstride + q];
    }
  }
}

}  // namespace mgpu_int

// ================================================================================================
// host: the plan (pure host code, also reachable from the CPU tests)
// ================================================================================================
namespace {

struct ResPlan {
  ResGeom g;
  std::vector<int4> tinfo, fixe;
  std::vector<unsigned> wpiece;
  std::vector<int2> task;
  bool ok = false;
};

constexpr int RES_MAX_THREADS = 512;

// material with the most (node, element) incidences among the nodes of a chunk
int chunk_material(const int *elem_type, int nex, int ney, int i0, int nv, int j, int k) {
  int cnt[3] = {0, 0, 0};
  for (int t = 0; t < nv; ++t)
    for (int c = 0; c < 8; ++c) {
      const int ex = i0 + t - 1 + ((c >> 2) & 1), ey = j - 1 + ((c >> 1) & 1), ez = k - 1 + (c & 1);
      cnt[elem_type[(ez * ney + ey) * nex + ex]]++;
    }
  int m = 0;
  for (int q = 1; q < 3; ++q)
    if (cnt[q] > cnt[m]) m = q;
  return m;
}

bool res_plan_try(int nx, int ny, int nz, const int *elem_type, const int *rowid, int cs, int py, int pz, ResPlan &out) {
  const int nix = nx - 2, niy = ny - 2, niz = nz - 2, nex = nx - 1, ney = ny - 1;
  if (nix < 1 || niy < 1 || niz < 1 || nx > 4096) return false;
  ResGeom &G = out.g;
  memset(&G, 0, sizeof(G));
  G.cs = cs;
  G.py = py;
  G.pz = pz;
  // nodes per thread: fewest executed node slots per row, then the larger chunk
  long best = -1;
  for (int tn = 8; tn >= 6; --tn) {
    const long exec = (long)((nix + tn - 1) / tn) * tn;
    if (best < 0 || exec < best) {
      best = exec;
      G.tn = tn;
    }
  }
  G.nchunk = (nix + G.tn - 1) / G.tn;
  int nyl = 0, nzl = 0;
  for (int qz = 0; qz < pz; ++qz)
    for (int qy = 0; qy < py; ++qy) {
      const int r = qz * py + qy;
      G.y0[r] = res_split(niy, py, qy);
      G.y1[r] = res_split(niy, py, qy + 1);
      G.z0[r] = res_split(niz, pz, qz);
      G.z1[r] = res_split(niz, pz, qz + 1);
      const int a = G.y1[r] - G.y0[r], b = G.z1[r] - G.z0[r];
      if (a < 1 || b < 1 || (py > 1 && a < 2) || (pz > 1 && b < 2)) return false;
      nyl = std::max(nyl, a);
      nzl = std::max(nzl, b);
    }
  if (nyl + 2 > 255 || nzl + 2 > 255 || G.nchunk > 255) return false;
  G.nyb = nyl + 2;
  G.nzb = nzl + 2;
  // odd pitches: the 64-bit loads of 16 consecutive rows (a half-warp) fall into 16 different banks
  G.pitch = std::max(nx, G.nchunk * G.tn + 2) | 1;
  // plane pitch: the first own row of the next plane takes the banks that own row nyl + 1 of this plane would take, so
  // 16 consecutive own rows (a half-warp) stay conflict-free across a plane boundary
  G.zp = G.nyb * G.pitch;
  while ((G.zp - nyl * G.pitch) % 16 != 0) ++G.zp;
  G.cstride = G.nzb * G.zp + G.pitch;  // + one scratch row: idle threads carry zeros there (the phases are branch-free)
  G.nrows_max = nyl * nzl;
  G.dpitch = (G.nchunk * G.tn) | 1;
  G.dstride = (G.nrows_max + 1) * G.dpitch;  // + one scratch row (idle threads)

  // work items (own row, chunk) of every CTA, sorted by (material, chunk, row) and dealt to the threads in that order:
  // a warp holds one material (its operator pass takes ONE of the three compile-time copies), mostly one chunk, and
  // consecutive rows; every material class is padded to whole warps
  struct Item {
    int m, c, rr;
  };
  std::vector<std::vector<Item>> items(cs);
  int nthreads = 32;
  for (int r = 0; r < cs; ++r) {
    const int nyr = G.y1[r] - G.y0[r], nzr = G.z1[r] - G.z0[r];
    std::vector<Item> cls[3];
    for (int c = 0; c < G.nchunk; ++c)
      for (int rr = 0; rr < nyr * nzr; ++rr) {
        const int j = G.y0[r] + rr % nyr + 1, k = G.z0[r] + rr / nyr + 1;
        const int m = chunk_material(elem_type, nex, ney, c * G.tn + 1, std::min(G.tn, nix - c * G.tn), j, k);
        cls[m].push_back(Item{m, c, rr});
      }
    for (int m = 0; m < 3; ++m) {
      items[r].insert(items[r].end(), cls[m].begin(), cls[m].end());
      while (items[r].size() % 32) items[r].push_back(Item{-1, 0, 0});
    }
    nthreads = std::max(nthreads, (int)items[r].size());
  }
  G.nthreads = nthreads;
  G.nwarps = nthreads / 32;
  if (G.nthreads > RES_MAX_THREADS) return false;

  // per-thread and interface tables
  out.tinfo.assign((size_t)cs * G.nthreads, make_int4(-1, 0, 0, 0));
  std::vector<std::vector<int4>> fix(cs);
  for (int r = 0; r < cs; ++r) {
    const int nyr = G.y1[r] - G.y0[r];
    for (int tid = 0; tid < (int)items[r].size(); ++tid) {
      const Item &it = items[r][tid];
      if (it.m < 0) continue;
      const int c = it.c, rr = it.rr, m = it.m;
      const int ly = rr % nyr + 1, lz = rr / nyr + 1;
      const int j = G.y0[r] + ly, k = G.z0[r] + lz, i0 = c * G.tn + 1;  // grid coordinates of the first node
      const int nv = std::min(G.tn, nix - c * G.tn);
      int fixmask = 0;
      const int fixbase = (int)fix[r].size();
      for (int t = 0; t < nv; ++t) {
        unsigned codes = 0;
        for (int cc = 0; cc < 8; ++cc) {
          const int ex = i0 + t - 1 + ((cc >> 2) & 1), ey = j - 1 + ((cc >> 1) & 1), ez = k - 1 + (cc & 1);
          const int ty = elem_type[(ez * ney + ey) * nex + ex];
          if (ty == m) continue;
          codes |= (unsigned)((m + ty) | (ty < m ? 4 : 0)) << (3 * cc);  // 1 + u = m + ty
        }
        if (codes) {
          fixmask |= 1 << t;
          const int mi = ((k - 1) * niy + (j - 1)) * nix + (i0 + t - 1);
          fix[r].push_back(make_int4(ly | (lz << 8) | ((i0 + t) << 16), (int)codes, rowid[mi], 0));
        }
      }
      out.tinfo[(size_t)r * G.nthreads + tid] =
          make_int4(ly | (lz << 8) | (c << 16) | (m << 24), ((1 << nv) - 1) | (fixmask << 8), fixbase, rr);
    }
    // Work order of the interface entries (w = position in owner order, where the owner thread looks its correction and
    // its 1 / diagonal up).  Every shared load of the interface pass sits at a fixed offset from the entry's node, so
    // 16 lanes (a half-warp) whose NODES fall into 16 different 8-B banks are conflict-free in all of them.
    {
      // groups of RES_GROUP: entries sorted by the set of element positions they need (a warp executes a position if ANY of its
      // lanes needs it), so the groups skip most positions; inside a group the half-warps are packed by bank
      std::vector<int4> sorted;
      for (int q = 0; q < (int)fix[r].size(); ++q) {
        int4 e = fix[r][q];
        e.w = q;
        sorted.push_back(e);
      }
      auto posmask = [](const int4 &e) {
        int m8 = 0;
        for (int c = 0; c < 8; ++c)
          if (((unsigned)e.y >> (3 * c)) & 7u) m8 |= 1 << c;
        return m8;
      };
      std::stable_sort(sorted.begin(), sorted.end(), [&](const int4 &a, const int4 &b) {
        const int ma = posmask(a), mb = posmask(b);
        const int pa = __builtin_popcount(ma), pb = __builtin_popcount(mb);
        return pa != pb ? pa < pb : ma < mb;
      });
      std::vector<int4> order;
      for (size_t g0 = 0; g0 < sorted.size(); g0 += RES_GROUP) {
        const size_t g1 = std::min(sorted.size(), g0 + RES_GROUP);
        std::vector<std::vector<int4>> bybank(16);
        for (size_t q = g0; q < g1; ++q) {
          const int4 &e = sorted[q];
          const int no = ((e.x >> 8) & 0xff) * G.zp + (e.x & 0xff) * G.pitch + (e.x >> 16);
          bybank[no & 15].push_back(e);
        }
        size_t left = g1 - g0;
        while (left > 0) {
          int taken = 0;
          for (int b = 0; b < 16 && taken < 16; ++b)
            if (!bybank[b].empty()) {
              order.push_back(bybank[b].back());
              bybank[b].pop_back();
              ++taken;
              --left;
            }
          while (taken < 16 && left > 0) {  // fill the half-warp with whatever is left (conflicts, but no idle lanes)
            int bb = 0;
            for (int b = 1; b < 16; ++b)
              if (bybank[b].size() > bybank[bb].size()) bb = b;
            order.push_back(bybank[bb].back());
            bybank[bb].pop_back();
            ++taken;
            --left;
          }
        }
      }
      fix[r] = order;
    }
    G.nfix[r] = (int)fix[r].size();
    G.fixcap = std::max(G.fixcap, G.nfix[r]);
    if (G.nfix[r] > RES_GROUP * RES_MAX_GROUPS || G.nwarps > 16 || nyr * (G.z1[r] - G.z0[r]) > 0xffff) return false;
    // Interface PIECES -> warps.  A piece = (group of 32 entries, half of the 8 element positions): about 0.4 of an
    // operator pass in issued instructions, whatever the number of active lanes.  A warp that computed a whole group on
    // top of its operator pass was the straggler of its CTA and of the cluster (15 k cycles per group against 8 - 13 k
    // for the operator pass, measured per warp with clock64); the pieces level the four SM quarters instead (warp w
    // issues on quarter w % 4; an operator pass counts 1).
    {
      std::vector<double> wload(G.nwarps, 0.0);
      std::vector<unsigned> pm(16, 0u);
      for (int w = 0; w < G.nwarps; ++w)
        for (int l = 0; l < 32; ++l)
          if (w * 32 + l < (int)items[r].size() && items[r][w * 32 + l].m >= 0) wload[w] = 1.0;
      for (int pc = 0; pc < 2 * ((G.nfix[r] + RES_GROUP - 1) / RES_GROUP); ++pc) {
        double q4[4] = {0, 0, 0, 0};
        for (int w = 0; w < G.nwarps; ++w) q4[w % 4] += wload[w];
        int bq = 0;
        for (int q = 1; q < std::min(4, G.nwarps); ++q)
          if (q4[q] < q4[bq]) bq = q;
        int bw = bq;
        for (int w = bq; w < G.nwarps; w += 4)
          if (wload[w] < wload[bw]) bw = w;
        wload[bw] += 0.4;
        pm[bw] |= 1u << pc;
      }
      if (out.wpiece.empty()) out.wpiece.assign((size_t)cs * 16, 0u);
      for (int w = 0; w < 16; ++w) out.wpiece[(size_t)r * 16 + w] = pm[w];
    }
  }
  out.fixe.assign((size_t)cs * std::max(G.fixcap, 1), make_int4(0, 0, 0, 0));
  for (int r = 0; r < cs; ++r) std::copy(fix[r].begin(), fix[r].end(), out.fixe.begin() + (size_t)r * G.fixcap);
  // halo rows to push: (own row, neighbour CTA) pairs
  std::vector<std::vector<int2>> task(cs);
  for (int r = 0; r < cs; ++r) {
    for (int kk = G.z0[r]; kk < G.z1[r]; ++kk)
      for (int jj = G.y0[r]; jj < G.y1[r]; ++jj) {
        int drank[3], doff[3];
        const int nd = res_push_dests(G, r, jj, kk, drank, doff);
        if (nd > 3) return false;
        for (int q = 0; q < nd; ++q)
          task[r].push_back(make_int2((kk - G.z0[r] + 1) * G.zp + (jj - G.y0[r] + 1) * G.pitch + 1,
                                      (drank[q] << 24) | (doff[q] + 1)));
      }
    G.ntask[r] = (int)task[r].size();
    G.taskcap = std::max(G.taskcap, G.ntask[r]);
    for (const int2 &tk : task[r]) G.halo_in[(unsigned)tk.y >> 24] += 3 * nix * (int)sizeof(double);
  }
  if (3 * G.cstride >= (1 << 24)) return false;
  out.task.assign((size_t)cs * std::max(G.taskcap, 1), make_int2(0, 0));
  for (int r = 0; r < cs; ++r) std::copy(task[r].begin(), task[r].end(), out.task.begin() + (size_t)r * G.taskcap);

  // shared-memory layout (bytes); every array 128-B aligned
  auto up = [](size_t b) { return (b + 127) / 128 * 128; };
  size_t off = up(sizeof(double) * 3 * G.cstride);
  G.off_du = (int)off;
  off += up(sizeof(double) * 3 * G.dstride);
  // blocks of more than 12 warps have 128 registers per thread: the third component of r lives in shared memory
  G.off_r2 = (int)off;
  if (G.nthreads > 384) off += up(sizeof(double) * G.dstride);
  G.off_task = (int)off;
  off += up(sizeof(int2) * G.taskcap);
  G.off_mbar = (int)off;
  off += up(sizeof(uint64_t) * 4);  // mbarriers: dot products A, dot products B, halo rows
  G.off_time = (int)off;
  off += up(sizeof(long long) * 16 * 9);  // phase counters of the timeline instrument (dbg bit 256)
  G.off_fixe = (int)off;
  off += up(sizeof(int4) * G.fixcap);
  G.off_fixk = (int)off;
  off += up(sizeof(double) * 3 * G.fixcap);
  G.off_fixout = (int)off;
  off += up(sizeof(double) * 2 * 3 * G.fixcap);  // [half][entry][3]
  G.off_red = (int)off;
  off += up(sizeof(double) * 2 * 2 * RES_MAX_CS);
  G.off_wp = (int)off;
  off += up(sizeof(double) * 2 * 32);
  G.smem_bytes = (int)off;
  if (off > (size_t)RES_SMEM_LIMIT) return false;
  out.ok = true;
  return true;
}

// smallest cluster that fits; among its decompositions the one with the fewest warps, then rows, on the busiest CTA
bool res_plan(int nx, int ny, int nz, const int *elem_type, const int *rowid, ResPlan &out, int force_cs = 0) {
  for (int cs = 1; cs <= RES_MAX_CS; cs *= 2) {
    if (force_cs > 0 && cs != force_cs) continue;
    ResPlan bestp;
    for (int py = 1; py <= cs; py *= 2) {
      ResPlan p;
      if (!res_plan_try(nx, ny, nz, elem_type, rowid, cs, py, cs / py, p)) continue;
      auto key = [](const ResPlan &q) { return std::make_pair(q.g.nthreads * 100000L + q.g.nrows_max, q.g.smem_bytes); };
      if (!bestp.ok || key(p) < key(bestp)) bestp = std::move(p);
    }
    if (bestp.ok) {
      out = std::move(bestp);
      return true;
    }
  }
  return false;
}

}  // namespace

// ================================================================================================
// device
// ================================================================================================
namespace {

__device__ __forceinline__ unsigned cluster_ctarank() {
  unsigned r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;\n" : "=r"(r));
  return r;
}
__device__ __forceinline__ unsigned cluster_id_x() {
  unsigned r;
  asm volatile("mov.u32 %0, %%clusterid.x;\n" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;\n" ::: "memory");
}
// shared::cluster address of the same variable in CTA `rank` of the cluster
__device__ __forceinline__ unsigned map_to_rank(const void *local_smem, unsigned rank) {
  unsigned la = (unsigned)__cvta_generic_to_shared(local_smem), ra;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;\n" : "=r"(ra) : "r"(la), "r"(rank));
  return ra;
}
// Asynchronous store into the shared memory of another CTA of the cluster that signals the DESTINATION's mbarrier with its
// byte count (st.async ... mbarrier::complete_tx): the receiver waits on its own mbarrier -- no cluster barrier, no
// cluster-scope fence (which ptxas turns into MEMBAR.ALL.GPU + an L1 invalidation) anywhere in the DPCG loop.
__device__ __forceinline__ void st_async_f64(unsigned addr, double v, unsigned mbar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.f64 [%0], %1, [%2];\n" ::"r"(addr), "d"(v),
               "r"(mbar)
               : "memory");
}
__device__ __forceinline__ void st_async_v2f64(unsigned addr, double a, double b, unsigned mbar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v2.f64 [%0], {%1, %2}, [%3];\n" ::"r"(addr),
               "d"(a), "d"(b), "r"(mbar)
               : "memory");
}

// mbarrier wait with a bound: a transaction that never arrives (a bug in the byte accounting, a CTA of the cluster lost)
// becomes a launch failure with a CUDA error on the host instead of a hung device (2^26 polls of a try_wait that itself
// waits a hardware time slice: seconds, against microseconds for a DPCG iteration).
__device__ __forceinline__ void mbar_wait_bounded(uint64_t *bar, unsigned parity) {
  const unsigned a = (unsigned)__cvta_generic_to_shared(bar);
  unsigned ok, spins = 0;
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
    if (!ok && ++spins > (1u << 26)) __trap();
  } while (!ok);
}

// Sums a and b over every thread of the cluster; the same bits in every thread of every CTA.  s_wp: [2][16] per-warp
// partials (unused entries stay zero), s_red: [RES_MAX_CS][2] per-CTA sums (entries of absent ranks stay zero).
__device__ __forceinline__ void cluster_sum2(double &a, double &b, double *s_wp, double *s_red, uint64_t *mbar,
                                             unsigned &parity, int cs, unsigned rank) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  a = warp_sum(a);
  b = warp_sum(b);
  if (lane == 0) {
    s_wp[w] = a;
    s_wp[16 + w] = b;
  }
  __syncthreads();
  if (threadIdx.x == 0) mbar_expect_tx(mbar, 16u * (unsigned)cs);  // this phase: one arrival + 16 B from every CTA
  if ((int)threadIdx.x < cs) {  // thread q delivers this CTA's sums to CTA q
    double sa = 0.0, sb = 0.0;
#pragma unroll
    for (int ww = 0; ww < 16; ++ww) {
      sa += s_wp[ww];
      sb += s_wp[16 + ww];
    }
    st_async_v2f64(map_to_rank(&s_red[2 * rank], threadIdx.x), sa, sb, map_to_rank(mbar, threadIdx.x));
  }
  mbar_wait_bounded(mbar, parity);
  parity ^= 1u;
  a = 0.0;
  b = 0.0;
#pragma unroll
  for (int q = 0; q < RES_MAX_CS; ++q) {
    a += s_red[2 * q];
    b += s_red[2 * q + 1];
  }
}

// The thread's TN nodes x 3 rows of A_m p: 9 neighbour rows x 3 components x (TN + 2) 64-bit shared loads, DFMAs whose
// row-block operand is a uniform register (kernel parameter).
// SPARSE: a pure-material row block of an axis-aligned orthotropic (e.g. isotropic) material on the regular grid is
// mirror-symmetric, so the coupling of components fi != fj towards the neighbour at offset o vanishes unless o_fi != 0
// and o_fj != 0: 153 of the 243 entries are structurally non-zero (verified on the actual row blocks at set-up; the
// dense copy runs otherwise).  The 9 neighbour rows (oy, oz) fall into 4 classes with the same pattern -- oy, oz zero
// or not -- and each class is a rolled loop over its rows with a compile-time body: 9 / 13 / 13 / 23 DFMAs per node.
template <int MAT, int TN, bool OY, bool OZ>
__device__ __forceinline__ void res_apply_row(const PureRows &R, const double *__restrict__ rb, int row, int cstride,
                                              double (&acc)[TN][3]) {
#pragma unroll
  for (int fj = 0; fj < 3; ++fj) {
    double pv[TN + 2];
#pragma unroll
    for (int h = 0; h < TN + 2; ++h) pv[h] = rb[fj * cstride + h];
#pragma unroll
    for (int di = 0; di < 3; ++di) {
      const double *a = &R.a[MAT * RB_LEN + (row * 3 + di) * RB_NBR];
      const bool ox = di != 1;
#pragma unroll
      for (int fi = 0; fi < 3; ++fi) {
        // offsets of the neighbour along the axes of fi and fj
        const bool ofi = fi == 0 ? ox : fi == 1 ? OY : OZ, ofj = fj == 0 ? ox : fj == 1 ? OY : OZ;
        if (fi != fj && !(ofi && ofj)) continue;
#pragma unroll
        for (int t = 0; t < TN; ++t) acc[t][fi] += a[fi * 3 + fj] * pv[t + di];
      }
    }
  }
}

template <int MAT, int TN, bool SPARSE>
__device__ __forceinline__ void res_apply(const PureRows &R, const double *__restrict__ brick, int base, int pitch,
                                          int zp, int cstride, double (&acc)[TN][3]) {
  if (SPARSE) {
    res_apply_row<MAT, TN, false, false>(R, brick + base + zp + pitch, 4, cstride, acc);  // (oy, oz) = (0, 0)
#pragma unroll 1
    for (int q = 0; q < 2; ++q) {  // oy != 0, oz = 0: rows 3 and 5
      const int row = 3 + 2 * q;
      res_apply_row<MAT, TN, true, false>(R, brick + base + zp + (2 * q) * pitch, row, cstride, acc);
    }
#pragma unroll 1
    for (int q = 0; q < 2; ++q) {  // oy = 0, oz != 0: rows 1 and 7
      const int row = 1 + 6 * q;
      res_apply_row<MAT, TN, false, true>(R, brick + base + (2 * q) * zp + pitch, row, cstride, acc);
    }
#pragma unroll 1
    for (int q = 0; q < 4; ++q) {  // oy != 0, oz != 0: rows 0, 2, 6, 8
      const int dk = 2 * (q >> 1), dj = 2 * (q & 1);
      res_apply_row<MAT, TN, true, true>(R, brick + base + dk * zp + dj * pitch, dk * 3 + dj, cstride, acc);
    }
  } else {
#pragma unroll 1
    for (int row = 0; row < 9; ++row) {
      const int dk = row / 3, dj = row - dk * 3;
      const double *rb = brick + base + dk * zp + dj * pitch;
#pragma unroll
      for (int fj = 0; fj < 3; ++fj) {
        double pv[TN + 2];
#pragma unroll
        for (int h = 0; h < TN + 2; ++h) pv[h] = rb[fj * cstride + h];
#pragma unroll
        for (int di = 0; di < 3; ++di) {
          const double *a = &R.a[MAT * RB_LEN + (row * 3 + di) * RB_NBR];
#pragma unroll
          for (int t = 0; t < TN; ++t) {
            const double pval = pv[t + di];
            acc[t][0] += a[fj] * pval;
            acc[t][1] += a[3 + fj] * pval;
            acc[t][2] += a[6 + fj] * pval;
          }
        }
      }
    }
  }
}

// Interface pass of one PIECE: 32 entries (one per lane; `on`: the lane holds an entry) x the element positions
// 4 half .. 4 half + 3.  For every pair that occurs in the warp and every position of the half that some lane needs,
// the lanes that need it add (Ke_t - Ke_m)[node rows] . p(element nodes): 72 DFMAs whose coefficient is a uniform
// register (kernel parameter at a compile-time address), 24 shared loads of p.
__device__ __forceinline__ void res_fix_piece(const ResDtab &DT, const double *__restrict__ brick, int no, int pitch,
                                              int zp, int cstride, unsigned codes, bool on, int half, double (&yout)[3]) {
  yout[0] = yout[1] = yout[2] = 0.0;
#pragma unroll
  for (int pr = 0; pr < RES_MAX_PAIRS; ++pr) {  // unrolled: the table entries are compile-time constant-bank addresses
                                                // (rolled, the coefficients came through indexed LDC: 5 x slower)
    unsigned m8 = 0, neg = 0;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      const unsigned cd = (codes >> (3 * c)) & 7u;
      if ((cd & 3u) == (unsigned)(pr + 1)) {
        m8 |= 1u << c;
        neg = cd >> 2;  // the same for every element of this pair around the node (m is the chunk's material)
      }
    }
    if (!on) m8 = 0;
    m8 &= 0xfu << (4 * half);
    const unsigned any8 = __reduce_or_sync(0xffffffffu, m8);
    if (!any8) continue;
    double y[4][3];  // 12 independent accumulation chains
#pragma unroll
    for (int q = 0; q < 4; ++q) y[q][0] = y[q][1] = y[q][2] = 0.0;
    const double *Dp = &DT.d[pr][0];
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      if (!((any8 >> c) & 1)) continue;  // warp-uniform (also skips the other half)
      if ((m8 >> c) & 1) {
        const int ax = (c >> 2) & 1, ay = (c >> 1) & 1, az = c & 1;
        const int eo = no + (az - 1) * zp + (ay - 1) * pitch + (ax - 1);  // first node of the element
#pragma unroll
        for (int jn = 0; jn < 8; ++jn) {
          const int q = eo + corner_z(jn) * zp + corner_y(jn) * pitch + corner_x(jn);
          const double p0 = brick[q], p1 = brick[cstride + q], p2 = brick[2 * cstride + q];
          const double *d = Dp + (c * 8 + jn) * 9;
          double(&yy)[3] = y[jn & 3];
          yy[0] += d[0] * p0;
          yy[1] += d[3] * p0;
          yy[2] += d[6] * p0;
          yy[0] += d[1] * p1;
          yy[1] += d[4] * p1;
          yy[2] += d[7] * p1;
          yy[0] += d[2] * p2;
          yy[1] += d[5] * p2;
          yy[2] += d[8] * p2;
        }
      }
    }
    const double sg = neg ? -1.0 : 1.0;
#pragma unroll
    for (int d = 0; d < 3; ++d) yout[d] += sg * ((y[0][d] + y[1][d]) + (y[2][d] + y[3][d]));
  }
}

// the halo copies of the own rows: warp per (row, neighbour) task, lanes along x -- coalesced remote stores through
// distributed shared memory
__device__ __forceinline__ void res_push_rows(const double *s_p, const int2 *s_task, int ntask, int nix, int cstride,
                                              int nwarps, uint64_t *mbar) {
  const int lane = threadIdx.x & 31;
  for (int k = threadIdx.x >> 5; k < ntask; k += nwarps) {
    const int2 tk = s_task[k];
    const unsigned dr = (unsigned)tk.y >> 24;
    const unsigned base = map_to_rank(s_p + (tk.y & 0xffffff), dr), rmb = map_to_rank(mbar, dr);
    for (int x = lane; x < nix; x += 32) {
#pragma unroll
      for (int d = 0; d < 3; ++d)
        st_async_f64(base + (unsigned)((d * cstride + x) * 8), s_p[tk.x + d * cstride + x], rmb);
    }
  }
}

// The per-thread vector state of a solve and its element-wise phases.  FIX = the thread owns at least one interface node
// (1 / diagonal of such a node comes from shared memory, volatile so that it is re-read where it is used rather than
// hoisted and spilled); every loop is branch-free and fully unrolled: the shared loads of a phase go out together.
template <int TN, bool RSM>
struct ResThread {
  double r[TN][RSM ? 2 : 3];
  double kk[3];
  double *s_p, *s_du, *s_r2;
  const double *s_fixk;
  int cstride, dstride, fixmask;

  __device__ __forceinline__ double r_get(int t, int d) const { return (RSM && d == 2) ? s_r2[t] : r[t][RSM ? (d & 1) : d]; }
  __device__ __forceinline__ void r_set(int t, int d, double v) {
    if (RSM && d == 2)
      s_r2[t] = v;
    else
      r[t][RSM ? (d & 1) : d] = v;
  }
  // The 1 / diagonal table of the interface nodes behind a pointer the compiler cannot trace: taken anew at the head of
  // every phase, the loads of a phase are ordinary (batched) loads but cannot be hoisted out of the DPCG loop, where the
  // 3 TN selected values would be spilled to local memory.
  __device__ __forceinline__ const double *fixk_now() const {
    const double *q = s_fixk;
    asm volatile("" : "+l"(q));
    return q;
  }
  // 1 / diagonal of the thread's TN nodes, component d (one batch of loads per component: 2 TN registers)
  template <bool FIX>
  __device__ __forceinline__ void kinv(const double *fk, int d, double (&kd)[TN]) const {
    int fi = 0;
#pragma unroll
    for (int t = 0; t < TN; ++t) {
      const bool isfix = FIX && ((fixmask >> t) & 1);
      kd[t] = isfix ? fk[fi * 3 + d] : kk[d];
      if (isfix) ++fi;
    }
  }
  template <bool FIX>
  __device__ __forceinline__ void init(const double (&bv)[TN][3], double &rzs, double &zzs) {
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      double kd[TN];
      kinv<FIX>(fixk_now(), d, kd);
#pragma unroll
      for (int t = 0; t < TN; ++t) {
        const double rv = bv[t][d];
        const double z = __dmul_rn(kd[t], rv);
        r_set(t, d, rv);
        rzs += rv * z;
        zzs += z * z;
        s_p[d * cstride + t] = z;
      }
    }
  }
  __device__ __forceinline__ double p_dot(const double (&acc)[TN][3]) const {
    double pv[TN][3];
#pragma unroll
    for (int t = 0; t < TN; ++t)
#pragma unroll
      for (int d = 0; d < 3; ++d) pv[t][d] = s_p[d * cstride + t];
    double s = 0.0;
#pragma unroll
    for (int t = 0; t < TN; ++t)
#pragma unroll
      for (int d = 0; d < 3; ++d) s += pv[t][d] * acc[t][d];
    return s;
  }
  template <bool FIX>
  __device__ __forceinline__ void update(const double (&acc)[TN][3], double alpha, double &zzs, double &rzs) {
    zzs = 0.0;
    rzs = 0.0;
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      double kd[TN], rr[TN];
      kinv<FIX>(fixk_now(), d, kd);
#pragma unroll
      for (int t = 0; t < TN; ++t) rr[t] = r_get(t, d);
#pragma unroll
      for (int t = 0; t < TN; ++t) {
        const double rv = rr[t] - alpha * acc[t][d];
        r_set(t, d, rv);
        const double z = __dmul_rn(kd[t], rv);
        zzs += z * z;
        rzs += rv * z;
      }
    }
  }
  // (the loads of a component go out together before its stores: the compiler cannot reorder them itself, the three
  // arrays might alias)
  __device__ __forceinline__ void du_update(double alpha) {
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      double pp[TN], du[TN];
#pragma unroll
      for (int t = 0; t < TN; ++t) {
        pp[t] = s_p[d * cstride + t];
        du[t] = s_du[d * dstride + t];
      }
#pragma unroll
      for (int t = 0; t < TN; ++t) s_du[d * dstride + t] = fma(alpha, pp[t], du[t]);
    }
  }
  template <bool FIX>
  __device__ __forceinline__ void p_update(double alpha, double beta) {
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      double pp[TN], du[TN], rr[TN], kd[TN];
      kinv<FIX>(fixk_now(), d, kd);
#pragma unroll
      for (int t = 0; t < TN; ++t) {
        pp[t] = s_p[d * cstride + t];
        du[t] = s_du[d * dstride + t];
        rr[t] = r_get(t, d);
      }
#pragma unroll
      for (int t = 0; t < TN; ++t) {
        s_du[d * dstride + t] = fma(alpha, pp[t], du[t]);
        const double z = __dmul_rn(kd[t], rr[t]);
        s_p[d * cstride + t] = z + beta * pp[t];
      }
    }
  }
};

// Registers: the register file is 4 x 16384 (one quarter per SM sub-partition, warps are dealt round-robin), so a block
// of 13..16 warps puts 4 warps on a quarter: 128 registers per thread (MAXT = 512); up to 12 warps leave 168 (MAXT = 384).
template <int TN, int MAXT, bool SPARSE>
__global__ void __launch_bounds__(MAXT, 1)
    k_cg_resident(const __grid_constant__ MeshConst P, const Lst L, int n_list, SlotTables T, VecPool V,
                  const __grid_constant__ ResGeom G, const __grid_constant__ PureRows R,
                  const __grid_constant__ ResDtab DT, int dbg) {
  extern __shared__ __align__(128) unsigned char s_raw[];
  double *s_p = reinterpret_cast<double *>(s_raw);
  double *s_du = reinterpret_cast<double *>(s_raw + G.off_du);
  double *s_r2 = reinterpret_cast<double *>(s_raw + G.off_r2);  // [nrows_max][dpitch], MAXT = 512 only
  int2 *s_task = reinterpret_cast<int2 *>(s_raw + G.off_task);
  int4 *s_fixe = reinterpret_cast<int4 *>(s_raw + G.off_fixe);
  double *s_fixk = reinterpret_cast<double *>(s_raw + G.off_fixk);
  double *s_fixout = reinterpret_cast<double *>(s_raw + G.off_fixout);
  double *s_redA = reinterpret_cast<double *>(s_raw + G.off_red);
  double *s_redB = s_redA + 2 * RES_MAX_CS;
  double *s_wp = reinterpret_cast<double *>(s_raw + G.off_wp);
  uint64_t *s_mbar = reinterpret_cast<uint64_t *>(s_raw + G.off_mbar);  // [0] dot products A, [1] B, [2] halo rows

  const int tid = threadIdx.x;
  const unsigned rank = cluster_ctarank();
  const int entry = (int)cluster_id_x() + L.yoff;
  const int cnt = L.dcount ? min(*L.dcount, n_list + L.yoff) : n_list + L.yoff;
  if (entry >= cnt) return;  // the whole cluster leaves
  const int slot = L.list[entry];
  mgpu_slot_state *st = &T.state[slot];
  const size_t vo = (size_t)slot * V.vstride;
  const int pitch = G.pitch, nyb = G.nyb, zp = G.zp, cstride = G.cstride;
  const int nfix = G.nfix[rank], ntask = (dbg & 2) ? 0 : G.ntask[rank];
  const unsigned halo_in = (dbg & 2) ? 0u : (unsigned)G.halo_in[rank];
  unsigned parA = 0, parB = 0, parC = 0;

  // ---- set-up: zero brick, du and the reduction mailboxes; tables into shared memory ----
  for (int q = tid; q < 3 * cstride; q += blockDim.x) s_p[q] = 0.0;
  for (int q = tid; q < 3 * G.dstride; q += blockDim.x) s_du[q] = 0.0;
  if (tid == 0) {
    mbar_init(&s_mbar[0], 1);
    mbar_init(&s_mbar[1], 1);
    mbar_init(&s_mbar[2], 1);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  if (tid < 4 * RES_MAX_CS) s_redA[tid] = 0.0;
  if (tid < 32) s_wp[tid] = 0.0;
  for (int q = tid; q < G.ntask[rank]; q += blockDim.x) s_task[q] = __ldg(&G.task[(size_t)rank * G.taskcap + q]);
  for (int q = tid; q < nfix; q += blockDim.x) {
    const int4 e = __ldg(&G.fixe[(size_t)rank * G.fixcap + q]);
    s_fixe[q] = e;
#pragma unroll
    for (int d = 0; d < 3; ++d) s_fixk[e.w * 3 + d] = __ldg(&V.rkinv[e.z * 3 + d]);
  }
  const int4 ti = __ldg(&G.tinfo[(size_t)rank * G.nthreads + tid]);
  const bool has = ti.x >= 0;
  const int ly = ti.x & 0xff, lz = (ti.x >> 8) & 0xff, ch = (ti.x >> 16) & 0xff, mat = has ? (ti.x >> 24) & 0x3 : 0;
  const int vmask = has ? (ti.y & 0xff) : 0, fixmask = has ? ((ti.y >> 8) & 0xff) : 0, fixbase = ti.z;
  const unsigned pmask = __ldg(&G.wpiece[rank * 16 + (tid >> 5)]);  // interface pieces of this warp
  const int i0 = ch * TN + 1;
  const int po = has ? lz * zp + ly * pitch + i0 : G.nzb * zp;  // own first node, component 0 of the brick (idle: scratch row)
  const int dof = has ? (ti.w & 0xffff) * G.dpitch + ch * TN : G.nrows_max * G.dpitch;  // own first node in du
  double kk[3] = {0.0, 0.0, 0.0};  // 1 / diagonal of the chunk's pure-material row block
#pragma unroll
  for (int d = 0; d < 3; ++d) kk[d] = G.rkp[mat][d];
  const bool lead = rank == 0 && tid == 0;
  double *hist = nullptr;
  int hist_k = 0;
  if (lead) {
    hist = st->cg_hist;
    hist_k = hist ? st->cg_hist_k : 0;
  }
  cluster_sync_all();  // every brick of the cluster is zeroed before the first push arrives

  // ---- r = b, z = k r, p = z; r.z, z.z (src/ell.cpp:73-91) ----
  // Nodes of the chunk beyond the last interior node (t >= nv) are carried along with r = p = du = 0 (their b is read
  // as 0 and their operator result is masked to 0), so no phase below needs a per-node branch; threads that own an
  // interface node take the <true> copy of a phase, which re-reads that node's 1 / diagonal from shared memory.
  constexpr bool RSM = MAXT > 384;  // r[.][2] in shared memory (register budget, see above)
  ResThread<TN, RSM> th;
  th.s_p = s_p + po;
  th.s_du = s_du + dof;
  th.s_r2 = s_r2 + dof;
  th.s_fixk = s_fixk + fixbase * 3;
  th.cstride = cstride;
  th.dstride = G.dstride;
  th.fixmask = fixmask;
#pragma unroll
  for (int d = 0; d < 3; ++d) th.kk[d] = kk[d];
  // warp-uniform choice of the phase copies: a warp with an interface node takes the <true> copy as a whole
  const bool wfix = __any_sync(0xffffffffu, fixmask != 0);
  double s0 = 0.0, s1 = 0.0;
  {
    const size_t gnode = vo + (size_t)(G.z0[rank] + lz) * P.nxny + (size_t)(G.y0[rank] + ly) * P.nx + i0;
    double bv[TN][3];
#pragma unroll
    for (int t = 0; t < TN; ++t)
#pragma unroll
      for (int d = 0; d < 3; ++d) bv[t][d] = ((vmask >> t) & 1) ? V.b[gnode + (size_t)d * P.nn_pad + t] : 0.0;
    if (wfix)
      th.template init<true>(bv, s0, s1);
    else
      th.template init<false>(bv, s0, s1);
  }
  // halo rows: pushed into the neighbours' bricks with their byte count signalled on the neighbours' mbarrier [2]; a
  // CTA starts the operator when its own mbarrier says that all its halo bytes have landed.  No hazard on the bricks:
  // a CTA pushes the rows of iteration i + 1 only after the dot products of iteration i, i.e. after every CTA of the
  // cluster has finished reading the halo of iteration i.
  auto halo_exchange = [&]() {
    __syncthreads();  // the new p of the own rows is complete (the operator reads the rows of other threads)
    if (tid == 0 && halo_in > 0) mbar_expect_tx(&s_mbar[2], halo_in);
    if (ntask > 0) res_push_rows(s_p, s_task, ntask, P.nix, cstride, G.nwarps, &s_mbar[2]);
    if (halo_in > 0) {  // CTA-uniform
      mbar_wait_bounded(&s_mbar[2], parC);
      parC ^= 1u;
    }
  };
  halo_exchange();
  cluster_sum2(s0, s1, s_wp, s_redB, &s_mbar[1], parB, G.cs, rank);
  double rz = s0, pn = sqrt(s1), alpha = 0.0, beta = 0.0, pAp = 0.0;
  const double pnorm0 = pn;
  int its = 0;
  if (hist_k > 0) hist[0] = pn;
  bool active = (0 < P.cg_max_its) && !(pn < P.cg_abs_tol || pn < pn * P.cg_rel_tol);  // src/ell.cpp:93-94

  // dbg bit 256 (tools/resident_timeline.py): per-warp cycle counters of the phases of an iteration, summed over the
  // solve and left in the slot's partial-sum buffer: [rank][warp][8] = interface pass, operator, wait for the block,
  // cluster barrier A, update, cluster barrier B, du / p update + push, cluster barrier C
  // (the counters live in shared memory: in registers they cost 18 registers of every thread, timing or not)
  const bool timing = (dbg & 256) != 0;
  long long *s_time = reinterpret_cast<long long *>(s_raw + G.off_time) + (tid >> 5) * 9;  // [warp][8 phases + last]
  if (timing && (tid & 31) == 0)
    for (int k = 0; k < 9; ++k) s_time[k] = 0;
  auto lap = [&](int k) {
    if (timing && (tid & 31) == 0) {
      const long long now = clock64();
      if (k >= 0) s_time[k] += now - s_time[8];
      s_time[8] = now;
    }
  };
  while (active) {
    lap(-1);
    // ---- Ap = A p ----
    // interface corrections first: the few warps that compute a group start with it while the other warps are already in
    // their operator pass, so its latency-bound instruction stream fills bubbles instead of running alone at the end
    auto fix_pass = [&]() {
      unsigned pm = pmask;
      while (pm) {  // warp-uniform
        const int pc = __ffs(pm) - 1, half = pc & 1;
        pm &= pm - 1;
        const int e = (pc >> 1) * 32 + (tid & 31);
        const bool on = e < nfix;
        const int4 fe = on ? s_fixe[e] : make_int4(0, 0, 0, 0);
        const int no = ((fe.x >> 8) & 0xff) * zp + (fe.x & 0xff) * pitch + (fe.x >> 16);
        double y[3];
        res_fix_piece(DT, s_p, no, pitch, zp, cstride, (unsigned)fe.y, on, half, y);
        if (on) {
          double *o = s_fixout + (half * G.fixcap + fe.w) * 3;
          o[0] = y[0];
          o[1] = y[1];
          o[2] = y[2];
        }
      }
    };
    const bool dofix = nfix > 0 && !(dbg & 1);  // CTA-uniform
    if (dofix) fix_pass();
    lap(0);
    double acc[TN][3];
#pragma unroll
    for (int t = 0; t < TN; ++t) acc[t][0] = acc[t][1] = acc[t][2] = 0.0;
    if (has && !(dbg & 8)) {
      const int abase = po - zp - pitch - 1;  // the window of the first neighbour row
      if (mat == 0)
        res_apply<0, TN, SPARSE>(R, s_p, abase, pitch, zp, cstride, acc);
      else if (mat == 1)
        res_apply<1, TN, SPARSE>(R, s_p, abase, pitch, zp, cstride, acc);
      else
        res_apply<2, TN, SPARSE>(R, s_p, abase, pitch, zp, cstride, acc);
    }
    lap(1);
    if (dofix) {
      __syncthreads();
      lap(2);
      if (fixmask) {
        int fi = fixbase;
#pragma unroll
        for (int t = 0; t < TN; ++t)
          if ((fixmask >> t) & 1) {
#pragma unroll
            for (int d = 0; d < 3; ++d) acc[t][d] += s_fixout[fi * 3 + d] + s_fixout[(G.fixcap + fi) * 3 + d];
            ++fi;
          }
      }
    }
#pragma unroll
    for (int t = 0; t < TN; ++t)
      if (!((vmask >> t) & 1)) acc[t][0] = acc[t][1] = acc[t][2] = 0.0;  // selects, not branches
    // ---- p.Ap, alpha (src/ell.cpp:97-100) ----
    s0 = th.p_dot(acc);
    s1 = 0.0;
    cluster_sum2(s0, s1, s_wp, s_redA, &s_mbar[0], parA, G.cs, rank);
    lap(3);
    pAp = s0;
    alpha = rz / pAp;
    // ---- r -= alpha Ap, z = k r, z.z, r.z (src/ell.cpp:103-110) ----
    if (wfix)
      th.template update<true>(acc, alpha, s0, s1);
    else
      th.template update<false>(acc, alpha, s0, s1);
    lap(4);
    cluster_sum2(s0, s1, s_wp, s_redB, &s_mbar[1], parB, G.cs, rank);
    lap(5);
    pn = sqrt(s0);
    beta = s1 / rz;
    rz = s1;
    ++its;
    if (its < hist_k) hist[its] = pn;
    active = (its < P.cg_max_its) && !(pn < P.cg_abs_tol || pn < pnorm0 * P.cg_rel_tol);  // src/ell.cpp:93-94
    if (dbg & 4) active = its < 60;
    // ---- x += alpha p (src/ell.cpp:102); p = z + beta p (src/ell.cpp:113) while the loop goes on ----
    if (!(dbg & 16)) {
      if (!active)
        th.du_update(alpha);
      else if (wfix)
        th.template p_update<true>(alpha, beta);
      else
        th.template p_update<false>(alpha, beta);
    }
    if (active) {
      lap(6);
      halo_exchange();  // every halo row has arrived before the next operator
      lap(7);
    }
  }
  if (timing && (tid & 31) == 0) {
    long long *o = reinterpret_cast<long long *>(T.partial + (size_t)slot * NRED * T.nblk_max) + ((size_t)rank * 16 + (tid >> 5)) * 8;
    for (int k = 0; k < 8; ++k) o[k] = s_time[k];
  }

  // ---- du back to the pool: own rows, and zeros on the boundary nodes of the ring (no other writer) ----
  __syncthreads();
  {
    const int nyr = G.y1[rank] - G.y0[rank], nzr = G.z1[rank] - G.z0[rank];
    const int total = nyb * G.nzb * P.nx;
    for (int q = tid; q < total; q += blockDim.x) {
      const int row = q / P.nx, x = q - row * P.nx;
      const int bz = row / nyb, by = row - bz * nyb;
      if (by > nyr + 1 || bz > nzr + 1) continue;
      const int j = G.y0[rank] + by, k = G.z0[rank] + bz;
      const bool own = by >= 1 && by <= nyr && bz >= 1 && bz <= nzr;
      const bool brow = j == 0 || j == P.ny - 1 || k == 0 || k == P.nz - 1;
      if (!own && !brow) continue;
      const bool inner = own && x >= 1 && x <= P.nix;
      const int rr = (bz - 1) * nyr + (by - 1);
      const size_t gi = vo + (size_t)k * P.nxny + (size_t)j * P.nx + x;
#pragma unroll
      for (int d = 0; d < 3; ++d)
        V.du[gi + (size_t)d * P.nn_pad] = inner ? s_du[d * G.dstride + rr * G.dpitch + (x - 1)] : 0.0;
    }
  }
  if (lead) {
    st->rz = rz;
    st->pAp = pAp;
    st->alpha = alpha;
    st->beta = beta;
    st->pnorm0 = pnorm0;
    st->pnorm = pn;
    st->cg_its = its;
    st->cg_active = 0;
  }
}

typedef void (*res_kernel_t)(const MeshConst, const Lst, int, SlotTables, VecPool, const ResGeom, const PureRows,
                             const ResDtab, int);
template <bool SPARSE>
res_kernel_t res_kernel_of(int tn, int nthreads) {
  if (nthreads <= 384) switch (tn) {
      case 6: return k_cg_resident<6, 384, SPARSE>;
      case 7: return k_cg_resident<7, 384, SPARSE>;
      default: return k_cg_resident<8, 384, SPARSE>;
    }
  switch (tn) {
    case 6: return k_cg_resident<6, 512, SPARSE>;
    case 7: return k_cg_resident<7, 512, SPARSE>;
    default: return k_cg_resident<8, 512, SPARSE>;
  }
}
res_kernel_t res_kernel(int tn, int nthreads, bool sparse) {
  return sparse ? res_kernel_of<true>(tn, nthreads) : res_kernel_of<false>(tn, nthreads);
}

// true when every entry of the three pure-material row blocks that the mirror symmetry of an axis-aligned orthotropic
// material makes vanish (components fi != fj towards a neighbour whose offset along fi or fj is zero) is zero up to
// rounding (1e-13 of the block's largest entry): the SPARSE operator copy then skips those 90 of 243 entries
bool res_rows_sparse(const PureRows &R) {
  for (int m = 0; m < 3; ++m) {
    double amax = 0.0;
    for (int q = 0; q < RB_LEN; ++q) amax = std::max(amax, fabs(R.a[m * RB_LEN + q]));
    for (int nbr = 0; nbr < 27; ++nbr) {
      const int o[3] = {nbr % 3 - 1, (nbr / 3) % 3 - 1, nbr / 9 - 1};
      for (int fi = 0; fi < 3; ++fi)
        for (int fj = 0; fj < 3; ++fj)
          if (fi != fj && !(o[fi] != 0 && o[fj] != 0) &&
              fabs(R.a[m * RB_LEN + nbr * RB_NBR + fi * 3 + fj]) > 1e-13 * amax)
            return false;
    }
  }
  return true;
}

template <class T>
T *res_to_device(mgpu_ctx *c, const std::vector<T> &v) {
  T *d = nullptr;
  CK(cudaMalloc(&d, sizeof(T) * std::max<size_t>(v.size(), 1)));
  if (!v.empty()) h2d_sync(c, d, v.data(), sizeof(T) * v.size());
  return d;
}

}  // namespace

struct mgpu_int::ResState {
  ResGeom g;
  int4 *d_tinfo = nullptr, *d_fixe = nullptr;
  int2 *d_task = nullptr;
  unsigned *d_wpiece = nullptr;
  ResDtab dtab;
  int max_clusters = 0;
  bool sparse = false;  // the pure-material row blocks have the mirror-symmetry zero pattern (res_rows_sparse)
  int dbg = 0;  // MICROPP_RES_DBG: timing experiments only (tools/resident_phases.py) -- results are wrong with any bit set
};

// called after implicit_setup (needs V.rowid's host image: rebuilt here from the element types)
void mgpu_int::resident_setup(mgpu_ctx *c, const mgpu_config *cfg, const int *rowid_host) {
  c->res = nullptr;
  if (!c->implicit || c->mc.slab) return;
  if (const char *env = getenv("MICROPP_RESIDENT"))
    if (atoi(env) == 0) return;
  int force_cs = 0;
  if (const char *env = getenv("MICROPP_RESIDENT_CS")) force_cs = atoi(env);
  const MeshConst &P = c->mc;
  ResPlan plan;
  const bool verbose = getenv("MICROPP_VERBOSE") != nullptr;
  if (!res_plan(P.nx, P.ny, P.nz, cfg->elem_type, rowid_host, plan, force_cs)) {
    if (verbose)
      fprintf(stderr, "micropp-b200: %dx%dx%d does not fit a cluster-resident DPCG; the three-kernel loop runs\n", P.nx,
              P.ny, P.nz);
    return;
  }
  ResState *rs = new ResState();
  rs->g = plan.g;
  rs->d_tinfo = res_to_device(c, plan.tinfo);
  rs->d_fixe = res_to_device(c, plan.fixe);
  rs->d_task = res_to_device(c, plan.task);
  if (plan.wpiece.empty()) plan.wpiece.assign((size_t)plan.g.cs * 16, 0u);
  rs->d_wpiece = res_to_device(c, plan.wpiece);
  rs->g.wpiece = rs->d_wpiece;
  rs->g.tinfo = rs->d_tinfo;
  rs->g.fixe = rs->d_fixe;
  rs->g.task = rs->d_task;
  memset(&rs->dtab, 0, sizeof(rs->dtab));
  for (int a = 0; a < 3; ++a)
    for (int b = a + 1; b < 3; ++b)
      for (int q = 0; q < RES_DLEN; ++q) rs->dtab.d[a + b - 1][q] = res_dtab_entry(cfg->ke_elastic, a, b, q);
  CK(cudaStreamSynchronize(c->stream));  // k_rows_build has filled rkinv: ids 0..2 are the pure-material row blocks
  CK(cudaMemcpy(&rs->g.rkp[0][0], c->V.rkinv, sizeof(double) * 9, cudaMemcpyDeviceToHost));
  rs->sparse = res_rows_sparse(c->pure_rows);
  if (const char *env = getenv("MICROPP_RESIDENT_DENSE"))
    if (atoi(env) != 0) rs->sparse = false;
  res_kernel_t kern = res_kernel(rs->g.tn, rs->g.nthreads, rs->sparse);
  {
    // the attribute belongs to the kernel, not to this context: only ever raise it (another live context may run the
    // same instantiation with a larger plan)
    static std::map<const void *, int> granted;
    static std::mutex granted_lock;
    std::lock_guard<std::mutex> guard(granted_lock);
    int &have = granted[(const void *)kern];
    if (rs->g.smem_bytes > have) {
      CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, rs->g.smem_bytes));
      have = rs->g.smem_bytes;
    }
  }
  cudaLaunchConfig_t cfgl = {};
  cfgl.gridDim = dim3(rs->g.cs, 1, 1);
  cfgl.blockDim = dim3(rs->g.nthreads, 1, 1);
  cfgl.dynamicSmemBytes = rs->g.smem_bytes;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = rs->g.cs;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfgl.attrs = at;
  cfgl.numAttrs = 1;
  int ncl = 0;
  const cudaError_t e = cudaOccupancyMaxActiveClusters(&ncl, kern, &cfgl);
  if (e != cudaSuccess || ncl < 1) {
    cudaGetLastError();
    fprintf(stderr, "micropp-b200: the cluster-resident DPCG (%d CTAs x %d threads, %d B) cannot be scheduled (%s); the "
                    "three-kernel loop runs\n", rs->g.cs, rs->g.nthreads, rs->g.smem_bytes, cudaGetErrorString(e));
    cudaFree(rs->d_tinfo);
    cudaFree(rs->d_fixe);
    cudaFree(rs->d_task);
    cudaFree(rs->d_wpiece);
    delete rs;
    return;
  }
  rs->max_clusters = ncl;
  if (const char *env = getenv("MICROPP_RES_DBG")) rs->dbg = atoi(env);
  c->res = rs;
  if (verbose)
    fprintf(stderr, "micropp-b200: cluster-resident DPCG: %d CTAs (%d x %d) x %d threads, %d nodes per thread, %d B of "
                    "shared memory, %d clusters in flight, %s operator\n", rs->g.cs, rs->g.py, rs->g.pz, rs->g.nthreads,
            rs->g.tn, rs->g.smem_bytes, ncl, rs->sparse ? "153-term (mirror-symmetric row blocks)" : "243-term");
}

void mgpu_int::resident_destroy(mgpu_ctx *c) {
  if (!c->res) return;
  cudaFree(c->res->d_tinfo);
  cudaFree(c->res->d_fixe);
  cudaFree(c->res->d_task);
  cudaFree(c->res->d_wpiece);
  delete c->res;
  c->res = nullptr;
}

void mgpu_int::launch_cg_resident(mgpu_ctx *c, int l, int n) {
  const ResState *rs = c->res;
  cudaLaunchConfig_t cfgl = {};
  cfgl.gridDim = dim3((unsigned)rs->g.cs * (unsigned)n, 1, 1);
  cfgl.blockDim = dim3(rs->g.nthreads, 1, 1);
  cfgl.dynamicSmemBytes = rs->g.smem_bytes;
  cfgl.stream = c->stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = rs->g.cs;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfgl.attrs = at;
  cfgl.numAttrs = 1;
  CK(cudaLaunchKernelEx(&cfgl, res_kernel(rs->g.tn, rs->g.nthreads, rs->sparse), c->mc, lst_of(c, l), n, c->T, c->V, rs->g, c->pure_rows,
                        rs->dtab, rs->dbg));
}

extern "C" {

// 1 when DPCG solves of this context run as the cluster-resident kernel
int mgpu_resident(const mgpu_ctx *c) { return c->res ? 1 : 0; }

// meta[8] = {CTAs per cluster, py, pz, nodes per thread, threads per CTA, bytes of shared memory, interface entries
// of the busiest CTA, clusters in flight}
void mgpu_resident_info(const mgpu_ctx *c, int *meta) {
  for (int q = 0; q < 8; ++q) meta[q] = 0;
  if (!c->res) return;
  const ResGeom &g = c->res->g;
  meta[0] = g.cs;
  meta[1] = g.py;
  meta[2] = g.pz;
  meta[3] = g.tn;
  meta[4] = g.nthreads;
  meta[5] = g.smem_bytes;
  meta[6] = g.fixcap;
  meta[7] = c->res->max_clusters;
}

// The DPCG solve of the first n entries of list l (cg_init + loop + finish of the three-kernel path in one launch).
void mgpu_cg_resident(mgpu_ctx *c, int l, int n) {
  if (n <= 0) return;
  if (!c->res) {
    fprintf(stderr, "micropp-b200: mgpu_cg_resident on a context without the cluster-resident DPCG\n");
    abort();
  }
  c->cg_op = OP_IMPLICIT;
  ProfScope ps(c, 5, n);
  launch_cg_resident(c, l, n);
  CK(cudaGetLastError());
}

// Isolated timing of the kernel on the first n slots (b as it stands in the pool), `dbg` as MICROPP_RES_DBG (bit 4
// = exactly 60 iterations whatever the numbers: a fixed amount of work); ms per launch.  Measurement only.
float mgpu_bench_resident(mgpu_ctx *c, int n, int reps, int dbg) {
  if (!c->res) return -1.f;
  CK(cudaSetDevice(c->device));
  n = std::min(n, c->W);
  std::vector<int> ids(n);
  for (int i = 0; i < n; ++i) ids[i] = i;
  mgpu_set_list(c, 5, n, ids.data());
  const int keep = c->res->dbg;
  c->res->dbg = dbg;
  launch_cg_resident(c, 5, n);
  CK(cudaEventRecord(c->t0, c->stream));
  for (int it = 0; it < reps; ++it) launch_cg_resident(c, 5, n);
  CK(cudaEventRecord(c->t1, c->stream));
  CK(cudaEventSynchronize(c->t1));
  CK(cudaGetLastError());
  c->res->dbg = keep;
  float ms = 0.f;
  CK(cudaEventElapsedTime(&ms, c->t0, c->t1));
  c->launches += reps + 1;
  return ms / reps;
}

// dbg bit 256: the per-warp phase counters the kernel left for `slot` ([8 ranks][16 warps][8 phases] cycles)
void mgpu_resident_timeline(mgpu_ctx *c, int slot, long long *out1024) {
  CK(cudaSetDevice(c->device));
  CK(cudaStreamSynchronize(c->stream));
  if ((size_t)NRED * c->T.nblk_max < 1024) {
    memset(out1024, 0, 1024 * sizeof(long long));
    return;
  }
  CK(cudaMemcpy(out1024, c->T.partial + (size_t)slot * NRED * c->T.nblk_max, 1024 * sizeof(long long),
                cudaMemcpyDeviceToHost));
}

// Host-only: do the three pure-material row blocks rows[3][27][9] have the mirror-symmetry zero pattern that lets the
// kernel run its 153-term operator copy (res_rows_sparse)?  (tests/test_resident_plan.py, no GPU)
int mgpu_resident_rows_sparse_host(const double *rows_pure) {
  PureRows R;
  memset(&R, 0, sizeof(R));
  for (int m = 0; m < 3; ++m)
    for (int nbr = 0; nbr < 27; ++nbr)
      for (int q = 0; q < 9; ++q) R.a[m * RB_LEN + nbr * RB_NBR + q] = rows_pure[((size_t)m * 27 + nbr) * 9 + q];
  return res_rows_sparse(R) ? 1 : 0;
}

// Host-only replay of the plan (tests/test_resident_plan.py, no GPU): applies the operator to p exactly as the kernel
// does -- per CTA a zeroed brick, own rows stored, halo rows delivered by res_push_dests, the chunk's pure-material row
// block on every node of a chunk and res_fix_corr on the interface entries -- and returns Ap.  p / Ap: [3][nn]
// (component-major, unpadded).  rows_pure: [3][27][9] row blocks of the three materials, ke: [3][576].
// meta[8] as mgpu_resident_info (clusters in flight = 0).  Returns 0 when no plan fits, else the number of CTAs.
int mgpu_resident_replay_host(int nx, int ny, int nz, const int *elem_type, const double *rows_pure, const double *ke,
                              const double *p, double *Ap, int *meta, int force_cs) {
  const int nix = nx - 2, niy = ny - 2, niz = nz - 2, nn = nx * ny * nz;
  if (nix < 1 || niy < 1 || niz < 1) return 0;
  std::vector<int> rowid((size_t)nix * niy * niz, 0);
  ResPlan plan;
  if (!res_plan(nx, ny, nz, elem_type, rowid.data(), plan, force_cs)) return 0;
  const ResGeom &G = plan.g;
  if (meta) {
    meta[0] = G.cs;
    meta[1] = G.py;
    meta[2] = G.pz;
    meta[3] = G.tn;
    meta[4] = G.nthreads;
    meta[5] = G.smem_bytes;
    meta[6] = G.fixcap;
    meta[7] = 0;
  }
  const int TN = G.tn;
  // difference tables, as the kernel builds them in shared memory
  std::vector<double> Dbuf((size_t)RES_MAX_PAIRS * RES_DLEN);
  double *D = Dbuf.data();
  for (int a = 0; a < 3; ++a)
    for (int b = a + 1; b < 3; ++b)
      for (int q = 0; q < RES_DLEN; ++q) D[(size_t)(a + b - 1) * RES_DLEN + q] = res_dtab_entry(ke, a, b, q);
  // every interface piece (group of 32 entries x half of the element positions) belongs to exactly one warp of its CTA
  for (int r = 0; r < G.cs; ++r) {
    unsigned seen = 0;
    for (int w = 0; w < 16; ++w) {
      const unsigned pm = plan.wpiece.empty() ? 0u : plan.wpiece[(size_t)r * 16 + w];
      if ((seen & pm) || (pm && w >= G.nwarps)) return -4;
      seen |= pm;
    }
    const int npc = 2 * ((G.nfix[r] + RES_GROUP - 1) / RES_GROUP);
    if (seen != (npc >= 32 ? 0xffffffffu : (1u << npc) - 1u)) return -4;
  }
  std::vector<std::vector<double>> brick(G.cs, std::vector<double>((size_t)3 * G.cstride, 0.0));
  // own values, then the push tasks (row copies between the bricks)
  for (int r = 0; r < G.cs; ++r)
    for (int tid = 0; tid < G.nthreads; ++tid) {
      const int4 ti = plan.tinfo[(size_t)r * G.nthreads + tid];
      if (ti.x < 0) continue;
      const int ly = ti.x & 0xff, lz = (ti.x >> 8) & 0xff, ch = (ti.x >> 16) & 0xff, i0 = ch * TN + 1;
      const int j = G.y0[r] + ly, k = G.z0[r] + lz;
      for (int t = 0; t < TN; ++t)
        if ((ti.y >> t) & 1)
          for (int d = 0; d < 3; ++d)
            brick[r][(size_t)d * G.cstride + lz * G.zp + ly * G.pitch + i0 + t] =
                p[(size_t)d * nn + (size_t)k * nx * ny + j * nx + i0 + t];
    }
  std::vector<int> incoming(G.cs, 0);
  for (int r = 0; r < G.cs; ++r)
    for (int k = 0; k < G.ntask[r]; ++k) {
      const int2 tk = plan.task[(size_t)r * G.taskcap + k];
      const int dr = (int)((unsigned)tk.y >> 24), doff = tk.y & 0xffffff;
      if (dr >= G.cs || dr == r) return -1;
      incoming[dr] += 3 * nix * (int)sizeof(double);
      for (int d = 0; d < 3; ++d)
        for (int x = 0; x < nix; ++x) brick[dr][(size_t)d * G.cstride + doff + x] = brick[r][(size_t)d * G.cstride + tk.x + x];
    }
  for (int r = 0; r < G.cs; ++r)  // the transaction count a CTA arms its halo mbarrier with = the bytes pushed to it
    if (incoming[r] != G.halo_in[r]) return -6;
  for (int q = 0; q < 3 * nn; ++q) Ap[q] = 0.0;
  std::vector<int> written(nn, 0);
  for (int r = 0; r < G.cs; ++r)
    for (int tid = 0; tid < G.nthreads; ++tid) {
      const int4 ti = plan.tinfo[(size_t)r * G.nthreads + tid];
      if (ti.x < 0) continue;
      const int ly = ti.x & 0xff, lz = (ti.x >> 8) & 0xff, ch = (ti.x >> 16) & 0xff, m = (ti.x >> 24) & 3;
      const int i0 = ch * TN + 1, j = G.y0[r] + ly, k = G.z0[r] + lz;
      const int vmask = ti.y & 0xff, fixmask = (ti.y >> 8) & 0xff;
      int fi = ti.z;
      const double *bk = brick[r].data();
      for (int t = 0; t < TN; ++t) {
        if (!((vmask >> t) & 1)) continue;
        double y[3] = {0.0, 0.0, 0.0};
        for (int row = 0; row < 9; ++row) {
          const int dk = row / 3, dj = row % 3;
          for (int di = 0; di < 3; ++di) {
            const double *a = rows_pure + ((size_t)m * 27 + row * 3 + di) * 9;
            const int q = (lz - 1 + dk) * G.zp + (ly - 1 + dj) * G.pitch + i0 + t - 1 + di;
            for (int fj = 0; fj < 3; ++fj) {
              const double pval = bk[(size_t)fj * G.cstride + q];
              for (int f = 0; f < 3; ++f) y[f] += a[f * 3 + fj] * pval;
            }
          }
        }
        if ((fixmask >> t) & 1) {
          int4 fe = make_int4(0, 0, 0, -1);
          for (int q = 0; q < G.nfix[r]; ++q)  // the list is in work order: look the owner index up
            if (plan.fixe[(size_t)r * G.fixcap + q].w == fi) fe = plan.fixe[(size_t)r * G.fixcap + q];
          if (fe.w != fi) return -5;
          const int no = ((fe.x >> 8) & 0xff) * G.zp + (fe.x & 0xff) * G.pitch + (fe.x >> 16);
          if (no != lz * G.zp + ly * G.pitch + i0 + t) return -2;
          double cr[3];
          res_fix_corr_host(D, bk, no, G.pitch, G.zp, G.cstride, (unsigned)fe.y, cr);
          for (int f = 0; f < 3; ++f) y[f] += cr[f];
          ++fi;
        }
        const int n = k * nx * ny + j * nx + i0 + t;
        written[n]++;
        for (int d = 0; d < 3; ++d) Ap[(size_t)d * nn + n] = y[d];
      }
    }
  // every interior node produced exactly once, no boundary node touched
  for (int k = 0; k < nz; ++k)
    for (int j = 0; j < ny; ++j)
      for (int i = 0; i < nx; ++i) {
        const bool bnd = i == 0 || i == nx - 1 || j == 0 || j == ny - 1 || k == 0 || k == nz - 1;
        if (written[(k * ny + j) * nx + i] != (bnd ? 0 : 1)) return -3;
      }
  return G.cs;
}

}  // extern "C"
