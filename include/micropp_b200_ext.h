/*
 * micropp_b200_ext.h -- additive C entry points next to the reference-compatible micropp_c.h.
 *
 * The reference C wrapper cannot set most of micropp_params_t (src/micropp_c.cpp:37-60 hard-codes
 * them), and the FE stages of micropp<3> are `protected` C++ members that the reference's own tests
 * reach by subclassing (test/test_cg.cpp:38-82).  These functions expose both to C / ctypes so that
 * the parity tests and bench.py can drive every stage of the hot path THROUGH THE C-ABI.
 * All pointers are host pointers in the reference's layouts (u: node*3+d; vars: e*56+gp*7+v;
 * ELL values: row*81+slot).
 */
#ifndef MICROPP_B200_EXT_H
#define MICROPP_B200_EXT_H

#include "micropp_c.h"

#ifdef __cplusplus
extern "C" {
#endif

/* Flat mirror of micropp_params_t (include/types.hpp). */
struct micropp3_params {
  int ngp;
  int size[3];
  int type;
  double geo_params[4];
  int mat_type[3];
  double mat_E[3], mat_nu[3], mat_Ka[3], mat_Sy[3], mat_Xt[3];
  const int *coupling; /* NULL => every GP is FE_ONE_WAY (src/micropp.cpp:88-96) */
  int subiterations;
  int nsubiterations;
  int mpi_rank;
  int nr_max_its;
  double nr_max_tol;
  double nr_rel_tol;
  int calc_ctan_lin;
  int use_A0;
  int its_with_A0;
  int lin_stress;
  int write_log;
};

void micropp3_new_ext(struct micropp3 *self, const struct micropp3_params *params);

/* batched boundary crossing: one call for all GPs (strain: ngp*6; stress: ngp*6; ctan: ngp*36) */
void micropp3_set_strains(struct micropp3 *self, const double *strain);
void micropp3_get_stresses(const struct micropp3 *self, double *stress);
void micropp3_get_ctans(const struct micropp3 *self, double *ctan);

/* inspection */
int micropp3x_nelem(const struct micropp3 *self);
int micropp3x_nndim(const struct micropp3 *self);
int micropp3x_wave_size(const struct micropp3 *self);
/* distinct ELL row blocks of the implicit operator of an all-elastic RVE; 0 = one assembled matrix per slot */
int micropp3x_implicit_rows(const struct micropp3 *self);
/* -1: assembled matrices; else the implicit SpMV kernel in use: 0 k_spmv_dot_imp (odd nx), 3 k_spmv_dot_tmac */
int micropp3x_implicit_kernel(const struct micropp3 *self);
/* cluster-resident DPCG (whole solve in one launch, one thread-block cluster per RVE; see mgpu_cg_resident):
   returns the CTAs per cluster (0 = the three-kernel loop runs) and fills meta8 as mgpu_resident_info (may be NULL) */
int micropp3x_resident_info(const struct micropp3 *self, int *meta8);
/* profiling mode: accumulated time of the cluster-resident DPCG solves */
double micropp3x_prof_resident_ms(struct micropp3 *self, int reset);
void micropp3x_get_elem_type(const struct micropp3 *self, int *out);
void micropp3x_get_bmat(const struct micropp3 *self, double *out /* [8][6][24] */);
void micropp3x_get_ctan_lin(const struct micropp3 *self, double *out36);
int micropp3x_get_u(const struct micropp3 *self, int gp, int which /*0=u_n,1=u_k*/, double *out);
int micropp3x_get_vars(const struct micropp3 *self, int gp, int which /*0=vars_n,1=vars_k*/, double *out);

/* FE stages (protected members of the reference class) */
void micropp3x_set_displ_bc(struct micropp3 *self, const double *eps, double *u);
double micropp3x_assembly_rhs(struct micropp3 *self, const double *u, const double *vars_old, double *b);
void micropp3x_assembly_mat(struct micropp3 *self, const double *u, const double *vars_old, double *vals);
void micropp3x_newton(struct micropp3 *self, const double *eps, const double *vars_old, double *u, int *out3);
void micropp3x_ave_stress(struct micropp3 *self, const double *u, const double *vars_old, double *sig);
int micropp3x_vars_new(struct micropp3 *self, const double *u, const double *vars_old, double *vars_new);
/* Ap = A p with the Jacobian at u = 0 without history (ell_mvp of src/ell.cpp:35-44 on the matrix assembly_mat
   builds), through the chosen DPCG operator: op 0 = assembled ELL matrix, 3 = implicit operator of an all-elastic
   RVE (kernel 0 table-driven / 3 TMA-tiled / -1 default).  p, Ap in the reference's layout [node][3]; returns p.Ap */
double micropp3x_apply_operator(struct micropp3 *self, const double *p, double *Ap, int op, int kernel);

/* ELL pieces */
void micropp3x_ell_cols(int nx, int ny, int nz, int *cols);
void micropp3x_ell_mvp(int nx, int ny, int nz, const double *vals, const double *x, double *y);
int micropp3x_ell_solve_cgpd(int nx, int ny, int nz, const double *vals, const double *b, double *x, double *err);
void micropp3x_elem_nodes(int nx, int ny, int ex, int ey, int ez, int *n8);
/* colour of an element in the 8-colour structured ordering: (ex&1) + 2(ey&1) + 4(ez&1) */
int micropp3x_elem_colour(int ex, int ey, int ez);

/* z-slab of one large RVE (single RVE split over several GPUs; no counterpart in the reference): device context
   of the node planes [z0, z1) of the nx x ny x nz grid in `params->size` plus one halo plane towards each
   neighbour.  Returns an `mgpu_ctx *` (include/mgpu.h) that the caller drives; see micropp_b200/slab.py. */
struct mgpu_ctx *micropp3x_slab_create(const struct micropp3_params *params, int z0, int z1, int device);

/* ---- z-slab mode behind the C ABI (host logic in micropp_b200/csrc/slab_host.cpp) --------------------------------
   One rank per GPU:  s = micropp3x_slab_new(&params, rank, size, device);  micropp3x_slab_export(s, &mine);
   all-gather the handles with whatever the macro code has (MPI_Allgather of bytes; torch.distributed in bench.py);
   micropp3x_slab_connect(s, all);  then every rank calls micropp3x_slab_homogenize(s, eps, sig, out3) -- inside a solve
   the ranks exchange halo planes and dot products over NVLink peer memory only (device-side flags, no collective).
   out3 = {Newton iterations, DPCG iterations, converged}.  Returns 0, or < 0 on misuse / a lost peer. */
struct micropp3x_slab;
struct micropp3x_slab_handle {
  char mail[64];      /* CUDA IPC handle of this rank's mailbox + halo receive buffers (one allocation) */
  int op, pad;        /* DPCG operator this rank would choose alone (3 implicit, 0 assembled) */
};
struct micropp3x_slab *micropp3x_slab_new(const struct micropp3_params *params, int rank, int size, int device);
void micropp3x_slab_free(struct micropp3x_slab *);
void micropp3x_slab_export(struct micropp3x_slab *, struct micropp3x_slab_handle *mine);
void micropp3x_slab_connect(struct micropp3x_slab *, const struct micropp3x_slab_handle *all /* [size] */);
int micropp3x_slab_homogenize(struct micropp3x_slab *, const double *eps6, double *stress6, int *out3);
/* several slabs of one RVE in ONE process / on one GPU (tests): plain device pointers instead of IPC handles */
void micropp3x_slab_connect_local(struct micropp3x_slab *const *group, int n);
int micropp3x_slab_homogenize_local(struct micropp3x_slab *const *group, int n, const double *eps6, double *stress6,
                                    int *out3);
void micropp3x_slab_planes(const struct micropp3x_slab *, int *z0, int *z1); /* owned node planes [z0, z1) */
void micropp3x_slab_get_u(struct micropp3x_slab *, double *u_local /* [nzl*ny*nx][3] */);
unsigned long long micropp3x_slab_launch_count(const struct micropp3x_slab *);
int micropp3x_slab_operator(const struct micropp3x_slab *);
/* test instrument: |z| at the head of the first k DPCG iterations of the latest solve (every slab holds the same
   globally summed values); see mgpu_cg_history in mgpu.h */
void micropp3x_slab_cg_history(struct micropp3x_slab *, int k);
int micropp3x_slab_cg_history_read(struct micropp3x_slab *, double *out, int k);

/* measurement (CUDA events on the library's own stream) */
void micropp3x_prof_enable(struct micropp3 *self, int on);
void micropp3x_prof_read(struct micropp3 *self, double *out9, int reset);
void micropp3x_cg_history(struct micropp3 *self, int k);   /* test instrument, see mgpu_cg_history (slot = GP index
                                                              inside the last wave) */
int micropp3x_cg_history_read(struct micropp3 *self, int slot, double *out, int k);
int micropp3x_hybrid_available(const struct micropp3 *self); /* hybrid operator for RVEs with a damage / plastic phase */
double micropp3x_last_homogenize_ms(const struct micropp3 *self);
unsigned long long micropp3x_launch_count(const struct micropp3 *self);
double micropp3x_bench_spmv(struct micropp3 *self, int nslots, int iters); /* ms per launch */
/* implicit elastic operator; kern: -1 the context's kernel, 0 k_spmv_dot_imp, 3 k_spmv_dot_tmac + k_spmv_fix */
double micropp3x_bench_imp_spmv(struct micropp3 *self, int nslots, int iters, int kern);
/* isolated timing of the cluster-resident DPCG kernel (mgpu_bench_resident); ms per launch, -1 when unavailable */
double micropp3x_bench_resident(struct micropp3 *self, int nslots, int reps, int dbg);
/* per-warp phase cycle counters of the last dbg-256 run of slot `slot`: [8 CTAs][16 warps][8 phases] */
void micropp3x_resident_timeline(struct micropp3 *self, int slot, long long *out1024);

#ifdef __cplusplus
}
#endif
#endif
