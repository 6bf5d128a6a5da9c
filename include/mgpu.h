/*
 * mgpu.h -- the thin C-ABI between the C++ host side of micropp-b200 and its sm_100a CUDA kernels.
 *
 * North star (BASELINE.json): "Host code stays C++ and calls CUDA through a thin C-ABI layer."
 * Every function below is `extern "C"`, takes plain pointers / sizes, and either manages device
 * memory of a context or enqueues exactly one kernel (or one small copy) on the context stream.
 * The orchestration (wave scheduling, Newton loop, CG loop, GP state machine) lives in C++
 * (micropp_b200/csrc/micropp_host.cpp) and only talks to the GPU through these entry points.
 *
 * Reference functions each launcher replaces (file:line under the reference tree):
 *   mgpu_set_bc          set_displ_bc            src/micro3D.cpp:27-78
 *   mgpu_asm_rhs         assembly_rhs            src/assembly.cpp:28-103, get_elem_rhs :124-138
 *   mgpu_asm_mat         assembly_mat            src/assembly.cpp:106-121, get_elem_mat :141-178,
 *                        ell_add_3D / ell_set_bc_3D  src/ell-common.cpp:166-198, :238-297
 *   mgpu_cg_init / mgpu_cg_spmv_dot / mgpu_cg_update / mgpu_cg_pupdate
 *                        ell_solve_cgpd, ell_mvp, get_dot   src/ell.cpp:35-122
 *   mgpu_axpy_u          u += du                 src/solve.cpp:73
 *   mgpu_ave_stress      calc_ave_stress         src/average.cpp:58-82
 *   mgpu_vars_new        calc_vars_new           src/update.cpp:33-56
 *
 * Device data layout (all FP64):
 *   vectors  : per slot, structure-of-arrays by displacement component: v[d*nn_pad + node]
 *   matrix   : per slot, INTERIOR rows only (boundary rows are the identity rows of ell_set_bc_3D and are neither
 *              stored nor read), in tiles of 32 consecutive interior nodes:
 *              A[((m/32)*243 + nbr*9 + fi*3 + fj)*32 + m%32], m = interior-node index (x fastest),
 *              nbr = 27-point stencil slot of src/ell-common.cpp:102-130; column indices are implicit
 *   int.vars : per Gauss point of the macro mesh, v[(var*8 + gp)*nelem_pad + elem], var < nvar
 */
#ifndef MGPU_H
#define MGPU_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct mgpu_ctx mgpu_ctx;

/* Per-slot solver state, mirrored on the host after mgpu_fetch_state. */
typedef struct mgpu_slot_state {
  double norm0, norm;           /* Newton residual norms (src/solve.cpp:39-41,75) */
  double rz, pAp, alpha, beta;  /* CG scalars (src/ell.cpp:86-116) */
  double pnorm0, pnorm;
  int nr_its, solver_its, nr_active, converged; /* newton_t (include/types.hpp:29-41) */
  int cg_its, cg_active;
  int nl_flag;                  /* OR of material->evolute() (src/update.cpp:49) */
  unsigned ticket;              /* last-block-done counter */
  double *cg_hist;              /* device, [cg_hist_k]: |z| at the head of DPCG iteration i of the LAST solve (NULL = off) */
  int cg_hist_k, pad_;
} mgpu_slot_state;

typedef struct mgpu_config {
  int nx, ny, nz;
  int device;                 /* CUDA device ordinal */
  int ngp;                    /* macro Gauss points that own FE state (u_n,u_k) */
  const int *elem_type;       /* [nelem] host, values 0..2 (src/micropp.cpp:117-124) */
  double dsh[8][24];          /* shape derivatives per Gauss point (src/micro3D.cpp:82-98) */
  double wg;                  /* Gauss weight (src/micropp.cpp:57) */
  double dx, dy, dz;
  double mat[3][8];           /* E,nu,Ka,Sy,k,mu,lambda,Xt per material */
  int mat_type[3];
  const double *ke_elastic;   /* [3][576] element matrices of the elastic materials (host) */
  int nr_max_its;
  double nr_max_tol, nr_rel_tol;
  int cg_max_its;
  double cg_abs_tol, cg_rel_tol;
  int wave_cap;               /* max slots (0 = size from free HBM) */
  /* z-slab of a larger RVE (all zero = a whole RVE).  nz above is the LOCAL plane count including halo planes;
     local plane k is global plane k + koff of nz_glob; halo_lo/halo_hi: the first/last local plane belongs to the
     neighbour rank; element layers [ez_own_lo, ez_own_hi) (local numbering) are the ones this rank averages. */
  int slab, koff, nz_glob, halo_lo, halo_hi, ez_own_lo, ez_own_hi;
  /* all-elastic RVEs only: serve the Jacobian from the table of distinct ELL row blocks (no per-slot matrix, DPCG
     operator 3) instead of assembling one matrix per slot; ignored when a damage/plastic material is in use */
  int implicit_elastic;
} mgpu_config;

/* ---- context ---- */
int mgpu_device_count(void);
mgpu_ctx *mgpu_create(const mgpu_config *cfg);
void mgpu_destroy(mgpu_ctx *);
int mgpu_wave_size(const mgpu_ctx *);
int mgpu_nn_pad(const mgpu_ctx *);
int mgpu_nelem_pad(const mgpu_ctx *);
int mgpu_nvar(const mgpu_ctx *);
void mgpu_sync(mgpu_ctx *);
unsigned long long mgpu_launch_count(const mgpu_ctx *);

/* ---- per-GP persistent state ---- */
void mgpu_gp_swap(mgpu_ctx *, int gp);              /* update_vars: pointer swaps (include/gp.hpp:95-105) */
int mgpu_gp_has_vars(const mgpu_ctx *, int gp);
void mgpu_gp_alloc_vars(mgpu_ctx *, int gp);        /* gp_t::allocate (include/gp.hpp:86-93): zeroed */
void mgpu_gp_free_vars(mgpu_ctx *, int gp);         /* back to "no history" (buffers return to the pool) */
/* reference-layout (AoS) import/export of u_n/u_k (which: 0=n,1=k) and vars_n/vars_k */
void mgpu_gp_get_u(mgpu_ctx *, int gp, int which, double *host_aos);
void mgpu_gp_set_u(mgpu_ctx *, int gp, int which, const double *host_aos);
void mgpu_gp_get_vars(mgpu_ctx *, int gp, int which, double *host_ref_layout);
void mgpu_gp_set_vars(mgpu_ctx *, int gp, int which, const double *host_ref_layout);

/* ---- wave set-up: bind slots to GPs / strains ---- */
/* slot_gp[i] >= 0 binds vars_old/new of that GP (vars_mode: 0 none(nullptr), 1 vars_n, ...) */
void mgpu_bind_slots(mgpu_ctx *, int n, const int *slots, const int *gps, const int *use_vars_old);
void mgpu_set_slot_strain(mgpu_ctx *, int n, const int *slots, const double *eps6);
void mgpu_set_list(mgpu_ctx *, int which_list, int n, const int *slots); /* upload an explicit slot list */

/* ---- kernels (each: one launch over list `which_list`, first `n` entries) ---- */
void mgpu_load_u(mgpu_ctx *, int which_list, int n, int which_u);   /* u_slot <- u_n / u_k of bound GP */
void mgpu_store_u(mgpu_ctx *, int which_list, int n, int which_u);  /* u_k of bound GP <- u_slot */
void mgpu_zero_u(mgpu_ctx *, int which_list, int n);                /* u_slot <- 0 */
void mgpu_set_bc(mgpu_ctx *, int which_list, int n);
void mgpu_asm_rhs(mgpu_ctx *, int which_list, int n, int mode);     /* mode 0: first of a Newton solve, 1: after update, 2: plain */
void mgpu_asm_mat(mgpu_ctx *, int which_list, int n, int to_shared); /* to_shared: assemble slot list[0] into the shared A0 buffer */
/* use_shared selects the operator of the solve: 0 the slot's own matrix, 1 the shared A0, 2 the generic host matrix,
   3 the implicit operator of an all-elastic RVE (mgpu_implicit() != 0), 4 the hybrid operator (mgpu_hybrid_available());
   mgpu_cg_update/pupdate follow the operator of the last mgpu_cg_init */
/* host-only (no GPU): tiling of the implicit operator's TMA kernel for an nx x ny x nz RVE; see mgpu_kernels.cu */
int mgpu_tmac_tiling_host(int nx, int ny, int nz, const int *elem_type, int *meta6, int *rowid, int *tiles4,
                          int *chunk_pure, int *fix_nodes2);
int mgpu_implicit(const mgpu_ctx *);
int mgpu_implicit_rows(const mgpu_ctx *);     /* distinct ELL row blocks of the implicit operator */
int mgpu_implicit_fix_nodes(const mgpu_ctx *); /* interior nodes served by k_spmv_fix (material interfaces + minority nodes of a chunk) */
int mgpu_implicit_kernel(const mgpu_ctx *);   /* -1 none; SpMV kernel of the implicit operator: 0 k_spmv_dot_imp (table-driven; odd nx), 3 k_spmv_dot_tmac (TMA load + TMA store) + k_spmv_fix */
void mgpu_cg_init(mgpu_ctx *, int which_list, int n, int use_shared);
void mgpu_cg_spmv_dot(mgpu_ctx *, int which_list, int n, int use_shared);
void mgpu_spmv_generic(mgpu_ctx *, int which_list, int n, int force); /* arbitrary matrix: boundary rows read too */
/* one forced application Ap = A p (+ p.Ap in the slot state) for kernel parity tests; op as in mgpu_cg_init;
   imp_kernel: -1 the context's choice, 0 k_spmv_dot_imp, 3 k_spmv_dot_tmac + k_spmv_fix */
void mgpu_apply_operator(mgpu_ctx *, int which_list, int n, int op, int imp_kernel);
void mgpu_cg_update(mgpu_ctx *, int which_list, int n);
void mgpu_cg_pupdate(mgpu_ctx *, int which_list, int n);
/* x += alpha p of the last iteration (deferred from mgpu_cg_update to mgpu_cg_pupdate, which the last iteration
   of a slot skips): call once after the DPCG loop, before du is used */
void mgpu_cg_finish(mgpu_ctx *, int which_list, int n);
/* Cluster-resident DPCG (cg_resident.cu): the WHOLE solve of src/ell.cpp:66-122 -- cg_init, every iteration, the
   deferred x update -- of each slot of the list in ONE launch, one thread-block cluster per RVE with p, du in shared
   memory and r in registers.  Available (mgpu_resident != 0) for all-elastic RVEs that fit a cluster (30^3 does);
   MICROPP_RESIDENT=0 keeps the three-kernel loop.  info: meta[8] = {CTAs per cluster, py, pz, nodes per thread,
   threads per CTA, shared-memory bytes, interface entries of the busiest CTA, clusters in flight}. */
int mgpu_resident(const mgpu_ctx *);
void mgpu_resident_info(const mgpu_ctx *, int *meta8);
void mgpu_cg_resident(mgpu_ctx *, int which_list, int n);
double mgpu_prof_resident_ms(mgpu_ctx *, int reset);
float mgpu_bench_resident(mgpu_ctx *, int n, int reps, int dbg); /* isolated timing, see cg_resident.cu */
void mgpu_resident_timeline(mgpu_ctx *, int slot, long long *out1024); /* per-warp phase cycles after a dbg-256 run */
/* host-only (CPU tests): see cg_resident.cu */
int mgpu_resident_rows_sparse_host(const double *rows_pure /* [3][27][9] */);
int mgpu_resident_replay_host(int nx, int ny, int nz, const int *elem_type, const double *rows_pure, const double *ke,
                              const double *p, double *Ap, int *meta8, int force_cs);
void mgpu_axpy_u(mgpu_ctx *, int which_list, int n);
void mgpu_ave_stress(mgpu_ctx *, int which_list, int n);
void mgpu_vars_new(mgpu_ctx *, int which_list, int n, int write);
/* calc_fields (src/average.cpp:85-112) of the RVE staged in `slot`: element averages, [nelem][6] each (host) */
void mgpu_elem_fields(mgpu_ctx *, int slot, double ivol, double *elem_strain, double *elem_stress);
/* compaction: list_out <- entries of list_in whose (mode 0: nr_active, 1: cg_active) flag is set; returns count (syncs) */
int mgpu_compact(mgpu_ctx *, int list_in, int n_in, int list_out, int mode);
int mgpu_compact_range(mgpu_ctx *, int list_in, int off, int n_in, int list_out, int mode); /* entries [off, off+n_in) */

/* One whole Newton step (assembly_mat -> DPCG as a device-driven WHILE node -> u += du -> assembly_rhs) as one
   CUDA graph over list 1 (the Newton list, whose device-side length mgpu_compact(.., 1, ..) maintains).  Returns
   the number of slots that need another step (syncs once). */
int mgpu_newton_step_graph(mgpu_ctx *, int n_active, int use_shared);
/* same on list set ls: 0 = lists 1 / 2 (as above), 1 = lists 6 / 7 (the hybrid-operator slots of mgpu_hybrid_split) */
int mgpu_newton_step_graph_on(mgpu_ctx *, int ls, int n_active, int use_shared);

/* ---- hybrid operator (use_shared = 4) of RVEs with a damage / plastic phase: implicit elastic row blocks for every node
   whose 8 elements are inside their linear regime, explicit ELL rows only for the others ---- */
int mgpu_hybrid_available(const mgpu_ctx *);
/* probe the first n slots of list l at their current iterate, build the per-slot node lists, split: list 6 <- slots for
   the hybrid operator, list 1 <- slots for the fully assembled one (l may be 1).  Syncs. */
void mgpu_hybrid_split(mgpu_ctx *, int which_list, int n, int *n_hybrid, int *n_full);
void mgpu_asm_mat_hyb(mgpu_ctx *, int which_list, int n); /* assembles the listed rows only */

/* ---- slab mode (one RVE split in z-slabs over several GPUs): the caller all-reduces the slab-local sums between a
   reducing kernel and its scalar tail, and exchanges the halo planes of p before every SpMV ---- */
void mgpu_tail(mgpu_ctx *, int which_list, int n, int kind /*0 rhs,1 cg_init,2 spmv,3 cg_update,4 ave_stress*/, int mode);
/* ---- slab mode over peer memory (NVLink P2P; no collective library in the DPCG loop) ----
   Every rank owns a mailbox that all ranks map (mgpu_ipc_export / mgpu_ipc_open); slab-local sums are PUSHED into every
   rank's mailbox with an epoch and summed in rank order by every rank's reduce_tail kernel (waiting on local memory
   only); the p update pushes its boundary planes into the neighbours' receive buffers and releases an epoch flag in
   their mailboxes.  No remote load anywhere on the critical path.  All device-side waits are bounded. */
void *mgpu_slab_mail(mgpu_ctx *);
void mgpu_ipc_export(void *devptr, char *handle64);
void *mgpu_ipc_open(int device, const char *handle64);
void mgpu_ipc_close(void *mapped);
/* mails[r]: mailbox of rank r as mapped here (r == rank: the own one).  A mailbox allocation also holds the two receive
   buffers of the halo planes, so one IPC handle per rank is all that travels */
void mgpu_slab_link(mgpu_ctx *, int rank, int size, void *const *mails);
void mgpu_slab_push_p(mgpu_ctx *);     /* p as it stands -> the neighbours' receive buffers + epoch flags (once per solve) */
void mgpu_slab_halo_take(mgpu_ctx *);  /* wait on own memory for the neighbours' pushes, copy them into the halo planes */
/* fused path (one launch per cross-rank reduction: fold + post + gather + tail; the p update publishes its epoch) */
void mgpu_slab_set_fused(mgpu_ctx *, int on);
void mgpu_slab_reduce_tail(mgpu_ctx *, int which_list, int k, int kind, int mode);
int mgpu_slab_error(mgpu_ctx *);  /* != 0: a wait for a peer timed out (syncs) */
void mgpu_slab_cg_iteration(mgpu_ctx *, int which_list, int op);
void mgpu_slab_cg_chunk(mgpu_ctx *, int which_list, int op, int iters); /* iters iterations as one CUDA graph launch */
/* raw device pointers for the exchange: which 0..5 = b,du,Ap,p,u,r of slot 0 ([3][nn_pad]); 10 = slab sums
   ([W][8] doubles); 11 = averaged stress ([W][6]) */
void *mgpu_dev_ptr(mgpu_ctx *, int which);
void *mgpu_stream(mgpu_ctx *);   /* the cudaStream_t every launch of this context goes to */

/* ---- results ---- */
void mgpu_fetch_state(mgpu_ctx *, int n, const int *slots, mgpu_slot_state *out); /* syncs */
int mgpu_slot_state_size(void); /* sizeof(mgpu_slot_state): bindings that mirror the struct check it when they load */
void mgpu_fetch_stress(mgpu_ctx *, int n, const int *slots, double *sig6);        /* syncs */
void mgpu_clear_nl_flags(mgpu_ctx *, int which_list, int n);

/* ---- staging (host-pointer versions of the protected micropp<3> kernels; slot 0) ---- */
void mgpu_stage_put_u(mgpu_ctx *, int slot, const double *host_aos);
void mgpu_stage_get_u(mgpu_ctx *, int slot, double *host_aos);
void mgpu_stage_get_vec(mgpu_ctx *, int slot, int which /*0=b,1=du,2=Ap,3=p*/, double *host_aos);
void mgpu_stage_put_vec(mgpu_ctx *, int slot, int which, const double *host_aos);
void mgpu_stage_put_vars(mgpu_ctx *, int slot, int which /*0=old,1=new*/, const double *host_ref_layout); /* NULL unbinds */
void mgpu_stage_get_vars_new(mgpu_ctx *, int slot, double *host_ref_layout);
void mgpu_stage_get_mat(mgpu_ctx *, int slot, double *vals_ref_layout /* [3nn][81] */);
void mgpu_stage_put_mat(mgpu_ctx *, int slot, const double *vals_ref_layout);
void mgpu_ell_cols(int nx, int ny, int nz, int *cols /* [3nn][81] */, int device); /* regenerates src/ell-common.cpp:34-139 on the GPU */

/* ---- measurement ---- */
/* Residual history of the DPCG solves (test instrument): from now on every slot records |z| = sqrt(z.z) -- the quantity
   src/ell.cpp:93-94 tests at the loop head -- of the first k iterations of its latest solve.  read: copies min(k, its+1)
   values of `slot`, returns how many. */
void mgpu_cg_history(mgpu_ctx *, int k);
int mgpu_cg_history_read(mgpu_ctx *, int slot, double *out, int k);
void mgpu_prof_enable(mgpu_ctx *, int on);
/* accumulated since last reset: [0]=spmv ms, [1]=spmv launches, [2]=spmv slot-applications, [3]=asm_mat ms,
   [4]=asm_rhs ms, [5]=cg_update+pupdate ms, [6]=hybrid-operator spmv ms, [7]=its slot-applications ([0]..[2] then count
   the other operators only), [8]=explicit rows streamed by those applications (sum of the listed-row counts) */
void mgpu_prof_read(mgpu_ctx *, double *out9, int reset);
void mgpu_timer_start(mgpu_ctx *);
float mgpu_timer_stop(mgpu_ctx *); /* ms on the context stream (syncs) */
/* isolated SpMV micro-benchmark on the first n slots of the pool (matrix contents as they are) */
float mgpu_bench_spmv(mgpu_ctx *, int n, int iters);
/* same for the implicit elastic operator; kern as in mgpu_apply_operator */
float mgpu_bench_imp_spmv(mgpu_ctx *, int n, int iters, int kern);

#ifdef __cplusplus
}
#endif
#endif
