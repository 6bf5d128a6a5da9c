// slab_host.cpp -- C++ host of the z-slab mode: ONE large RVE (BASELINE configs[4]) split into contiguous ranges of node
// planes, one per GPU.  The reference has no such mode (its only parallelism is over independent Gauss points, SURVEY
// 2.4); the arithmetic of every slab is that of homogenize() (src/homogenize.cpp:112-187 without history:
// set_displ_bc -> Newton-Raphson src/solve.cpp:29-82 -> DPCG src/ell.cpp:66-122 -> calc_ave_stress src/average.cpp:58-82).
//
// A macro code (C++, C, Fortran through the C ABI) drives it without Python:
//     s = micropp3x_slab_new(&params, rank, size, device);          // this rank's planes + one halo plane per neighbour
//     micropp3x_slab_export(s, &mine);                              // CUDA IPC handle of the mailbox (+ receive buffers)
//     MPI_Allgather(&mine, sizeof mine, MPI_BYTE, all, sizeof mine, MPI_BYTE, comm);      // the ONLY collective: set-up
//     micropp3x_slab_connect(s, all);
//     micropp3x_slab_homogenize(s, eps, sig, out);                  // every rank calls it; no host-side communication
// Inside a solve the ranks talk through NVLink peer memory only, and only by REMOTE STORES: the p update pushes its
// boundary planes into the neighbours' receive buffers, the slab-local dot products are pushed into every rank's mailbox,
// each followed by a device-side epoch flag; every wait polls local memory (k_cg_pupdate*, k_slab_halo_take,
// k_slab_reduce_tail in mgpu_kernels.cu).  A chunk of DPCG iterations is one replayed CUDA graph.  Several slabs
// may also live in ONE process (micropp3x_slab_connect_local): the same kernels with plain device pointers, used by
// the single-GPU tests.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "mgpu.h"
#include "micropp_b200_ext.h"

struct micropp3x_slab {
  mgpu_ctx *ctx = nullptr;
  int rank = 0, size = 1, device = 0;
  int nx = 0, ny = 0, nz = 0, z0 = 0, z1 = 0, halo_lo = 0, halo_hi = 0, nzl = 0;
  int op = 0;           // DPCG operator: 3 implicit (all-elastic), 0 the slab's own ELL matrix
  int elastic = 1;      // every material elastic (no history needed)
  int solves = 0;
  int cg_chunk = 8;
  bool linked = false;
  std::vector<void *> mapped;  // IPC mappings to close
};

namespace {
const int kSlot = 0, kList = 0;

void plane_range(int nz, int nslabs, int s, int *z0, int *z1) {
  // contiguous, remainder to the low slabs (the rule the reference's MPI drivers use for Gauss points,
  // test/multi-gpu-mpi.cpp:60)
  int b = 0;
  for (int r = 0; r <= s; ++r) {
    const int cnt = nz / nslabs + (nz % nslabs > r ? 1 : 0);
    if (r == s) {
      *z0 = b;
      *z1 = b + cnt;
    }
    b += cnt;
  }
}

mgpu_slot_state state_of(micropp3x_slab *s) {
  mgpu_slot_state st;
  mgpu_fetch_state(s->ctx, 1, &kSlot, &st);
  return st;
}

// `f(slab)` for every slab of an in-process group, or for the one slab of this rank
template <class F>
void each(micropp3x_slab *const *g, int n, F f) {
  for (int i = 0; i < n; ++i) f(g[i]);
}

void link(micropp3x_slab *s, const std::vector<void *> &mails, const std::vector<int> &ops) {
  mgpu_slab_link(s->ctx, s->rank, s->size, mails.data());
  // one operator for the whole RVE: the implicit one only if EVERY slab is all-elastic (a rank deciding from its own
  // elements alone could pick another operator -- and another launch sequence -- than its neighbours)
  s->op = *std::min_element(ops.begin(), ops.end()) >= 3 ? 3 : 0;
  mgpu_slab_set_fused(s->ctx, 1);
  s->linked = true;
}

// reducing kernel already enqueued on every slab -> cross-rank sum + scalar tail
void reduce_tail(micropp3x_slab *const *g, int n, int k, int kind, int mode) {
  each(g, n, [&](micropp3x_slab *s) { mgpu_slab_reduce_tail(s->ctx, kList, k, kind, mode); });
}

void cg_solve(micropp3x_slab *const *g, int n) {
  each(g, n, [&](micropp3x_slab *s) { mgpu_cg_init(s->ctx, kList, 1, s->op); });
  reduce_tail(g, n, 2, 1, 0);
  each(g, n, [&](micropp3x_slab *s) { mgpu_slab_push_p(s->ctx); });
  // every rank takes the same decisions (rank-ordered sums => identical bits), so slab 0 speaks for all
  while (state_of(g[0]).cg_active)
    each(g, n, [&](micropp3x_slab *s) { mgpu_slab_cg_chunk(s->ctx, kList, s->op, s->cg_chunk); });
  each(g, n, [&](micropp3x_slab *s) { mgpu_cg_finish(s->ctx, kList, 1); });  // deferred x += alpha p of the last iteration
}

int homogenize(micropp3x_slab *const *g, int n, const double *eps, double *stress, int *out3) {
  for (int i = 0; i < n; ++i) {
    if (!g[i]->linked) {
      fprintf(stderr, "micropp-b200: micropp3x_slab_homogenize before micropp3x_slab_connect\n");
      return -1;
    }
    if (!g[i]->elastic && g[i]->solves > 0) {
      // no internal-variable state is kept between calls: a damage / plastic RVE is only valid for its first load step
      fprintf(stderr, "micropp-b200: the z-slab mode keeps no history; a non-elastic RVE can be solved once (virgin state)\n");
      return -2;
    }
  }
  each(g, n, [&](micropp3x_slab *s) {
    mgpu_set_slot_strain(s->ctx, 1, &kSlot, eps);
    mgpu_zero_u(s->ctx, kList, 1);
    mgpu_set_bc(s->ctx, kList, 1);
    mgpu_asm_rhs(s->ctx, kList, 1, 0);
  });
  reduce_tail(g, n, 1, 0, 0);
  while (state_of(g[0]).nr_active) {
    each(g, n, [&](micropp3x_slab *s) {
      if (s->op == 0) mgpu_asm_mat(s->ctx, kList, 1, 0);
    });
    cg_solve(g, n);
    each(g, n, [&](micropp3x_slab *s) {
      mgpu_axpy_u(s->ctx, kList, 1);
      mgpu_asm_rhs(s->ctx, kList, 1, 1);
    });
    reduce_tail(g, n, 1, 0, 1);
  }
  each(g, n, [&](micropp3x_slab *s) { mgpu_ave_stress(s->ctx, kList, 1); });
  reduce_tail(g, n, 6, 4, 0);
  mgpu_fetch_stress(g[0]->ctx, 1, &kSlot, stress);
  const mgpu_slot_state st = state_of(g[0]);
  if (out3) {
    out3[0] = st.nr_its;
    out3[1] = st.solver_its;
    out3[2] = st.converged;
  }
  int err = 0;
  for (int i = 0; i < n; ++i) {
    g[i]->solves++;
    err |= mgpu_slab_error(g[i]->ctx);
  }
  return err ? -3 : 0;  // -3: a device-side wait for a peer timed out
}
}  // namespace

extern "C" {

micropp3x_slab *micropp3x_slab_new(const micropp3_params *q, int rank, int size, int device) {
  if (size < 1 || rank < 0 || rank >= size || q->size[2] < size) {
    fprintf(stderr, "micropp-b200: bad slab request (rank %d of %d, %d node planes)\n", rank, size, q->size[2]);
    return nullptr;
  }
  micropp3x_slab *s = new micropp3x_slab();
  s->rank = rank;
  s->size = size;
  s->device = device;
  s->nx = q->size[0];
  s->ny = q->size[1];
  s->nz = q->size[2];
  plane_range(s->nz, size, rank, &s->z0, &s->z1);
  s->halo_lo = s->z0 > 0;
  s->halo_hi = s->z1 < s->nz;
  s->nzl = s->z1 + s->halo_hi - (s->z0 - s->halo_lo);
  s->ctx = micropp3x_slab_create(q, s->z0, s->z1, device);
  for (int i = 0; i < 3; ++i) s->elastic &= q->mat_type[i] == 0;
  s->op = mgpu_implicit(s->ctx) ? 3 : 0;
  const int none = -1;
  mgpu_bind_slots(s->ctx, 1, &kSlot, &none, nullptr);
  mgpu_set_list(s->ctx, kList, 1, &kSlot);
  if (const char *env = getenv("MICROPP_SLAB_CHUNK")) s->cg_chunk = std::max(1, atoi(env));
  return s;
}

void micropp3x_slab_free(micropp3x_slab *s) {
  if (!s) return;
  mgpu_sync(s->ctx);
  for (void *m : s->mapped) mgpu_ipc_close(m);
  mgpu_destroy(s->ctx);
  delete s;
}

void micropp3x_slab_export(micropp3x_slab *s, micropp3x_slab_handle *out) {
  memset(out, 0, sizeof(*out));
  mgpu_ipc_export(mgpu_slab_mail(s->ctx), out->mail);
  out->op = s->op;
}

void micropp3x_slab_connect(micropp3x_slab *s, const micropp3x_slab_handle *all) {
  std::vector<void *> mails(s->size);
  std::vector<int> ops(s->size);
  for (int r = 0; r < s->size; ++r) {
    ops[r] = all[r].op;
    if (r == s->rank) {
      mails[r] = mgpu_slab_mail(s->ctx);
      continue;
    }
    mails[r] = mgpu_ipc_open(s->device, all[r].mail);
    s->mapped.push_back(mails[r]);
  }
  link(s, mails, ops);
}

void micropp3x_slab_connect_local(micropp3x_slab *const *group, int n) {
  std::vector<void *> mails(n);
  std::vector<int> ops(n);
  for (int r = 0; r < n; ++r) {
    mails[r] = mgpu_slab_mail(group[r]->ctx);
    ops[r] = group[r]->op;
  }
  for (int r = 0; r < n; ++r) link(group[r], mails, ops);
}

int micropp3x_slab_homogenize(micropp3x_slab *s, const double *eps, double *stress, int *out3) {
  return homogenize(&s, 1, eps, stress, out3);
}
int micropp3x_slab_homogenize_local(micropp3x_slab *const *group, int n, const double *eps, double *stress, int *out3) {
  return homogenize(group, n, eps, stress, out3);
}

void micropp3x_slab_planes(const micropp3x_slab *s, int *z0, int *z1) {
  *z0 = s->z0;
  *z1 = s->z1;
}
// displacements of the slab's local planes (halo planes included), reference layout [node][3]
void micropp3x_slab_get_u(micropp3x_slab *s, double *u_local) { mgpu_stage_get_u(s->ctx, kSlot, u_local); }
unsigned long long micropp3x_slab_launch_count(const micropp3x_slab *s) { return mgpu_launch_count(s->ctx); }
int micropp3x_slab_operator(const micropp3x_slab *s) { return s->op; }
void micropp3x_slab_cg_history(micropp3x_slab *s, int k) { mgpu_cg_history(s->ctx, k); }
int micropp3x_slab_cg_history_read(micropp3x_slab *s, double *out, int k) { return mgpu_cg_history_read(s->ctx, 0, out, k); }
}
