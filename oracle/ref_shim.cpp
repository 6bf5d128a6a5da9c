/*
 * oracle/ref_shim.cpp -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * A thin C-ABI window onto the UNMODIFIED reference CPU implementation
 * (gagiuntoli/Micropp, compiled from the sources where they lie under
 * /root/reference by oracle/Makefile into oracle/_ref/libmicropp_ref.so).
 *
 * The reference keeps its FE kernels `protected` (include/micropp.hpp:125-169);
 * its own tests reach them by subclassing (test/test_cg.cpp:38-82).  This shim
 * does the same, and re-exports every stage of the homogenization hot path as
 * plain `extern "C"` functions so that Python (ctypes) can
 *   (1) generate the golden fixtures under tests/golden/,
 *   (2) validate the C restatement oracle/micropp_oracle.c,
 *   (3) act as the checker in the `-m gpu` parity tests and
 *   (4) be timed as the `--impl reference` / `cpu_baseline` arm of bench.py.
 *
 * Nothing here re-implements reference arithmetic: every function forwards to
 * a reference symbol.  No reference source text is copied.
 */
#include <cstdlib>
#include <cstring>

#include "micropp.hpp"

namespace {

class ref_probe : public micropp<3> {
 public:
  explicit ref_probe(const micropp_params_t &p) : micropp<3>(p) {}

  int p_nx() const { return nx; }
  int p_ny() const { return ny; }
  int p_nz() const { return nz; }
  int p_nelem() const { return nelem; }
  int p_nndim() const { return nndim; }
  int p_nvars() const { return nvars; }
  const int *p_elem_type() const { return elem_type; }
  const double *p_bmat() const { return &bmat[0][0][0]; }
  const double *p_ctan_lin() const { return ctan_lin_fe; }
  double p_wg() const { return wg; }
  gp_t<3> *p_gp(int i) { return &gp_list[i]; }

  void p_set_displ_bc(const double *eps, double *u) { set_displ_bc(eps, u); }
  double p_assembly_rhs(const double *u, const double *vars, double *b) { return assembly_rhs(u, vars, b); }
  void p_assembly_mat(ell_matrix *A, const double *u, const double *vars) { assembly_mat(A, u, vars); }
  newton_t p_newton(ell_matrix *A, double *b, double *u, double *du, const double *eps, const double *vars) {
    return newton_raphson(A, b, u, du, eps, vars);
  }
  void p_ave_stress(const double *u, double *sig, const double *vars) const { calc_ave_stress(u, sig, vars); }
  void p_ave_strain(const double *u, double *eps) const { calc_ave_strain(u, eps); }
  bool p_vars_new(const double *u, const double *vo, double *vn) const { return calc_vars_new(u, vo, vn); }
};

inline ref_probe *P(void *h) { return static_cast<ref_probe *>(h); }

}  // namespace

extern "C" {

/* Flat mirror of micropp_params_t (include/types.hpp:43-83) for ctypes. */
struct ref_params {
  int ngp;
  int size[3];
  int type;
  double geo_params[4];
  /* per material: type, E, nu, Ka, Sy, Xt -> passed through material_set (src/material.c:26-38) */
  int mat_type[3];
  double mat_E[3], mat_nu[3], mat_Ka[3], mat_Sy[3], mat_Xt[3];
  const int *coupling; /* may be NULL */
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

void *ref_new(const ref_params *rp) {
  micropp_params_t p;
  p.ngp = rp->ngp;
  memcpy(p.size, rp->size, sizeof(p.size));
  p.type = rp->type;
  memcpy(p.geo_params, rp->geo_params, sizeof(p.geo_params));
  for (int i = 0; i < 3; ++i)
    material_set(&p.materials[i], rp->mat_type[i], rp->mat_E[i], rp->mat_nu[i], rp->mat_Ka[i], rp->mat_Sy[i],
                 rp->mat_Xt[i]);
  p.coupling = const_cast<int *>(rp->coupling);
  p.subiterations = rp->subiterations != 0;
  p.nsubiterations = rp->nsubiterations;
  p.mpi_rank = rp->mpi_rank;
  p.nr_max_its = rp->nr_max_its;
  p.nr_max_tol = rp->nr_max_tol;
  p.nr_rel_tol = rp->nr_rel_tol;
  p.calc_ctan_lin = rp->calc_ctan_lin != 0;
  p.use_A0 = rp->use_A0 != 0;
  p.its_with_A0 = rp->its_with_A0;
  p.lin_stress = rp->lin_stress != 0;
  p.write_log = rp->write_log != 0;
  return new ref_probe(p);
}

void ref_free(void *h) { delete P(h); }

/* ---- public API of the reference class (include/micropp.hpp:176-217) ---- */
void ref_set_strain(void *h, int gp, const double *eps) { P(h)->set_strain(gp, eps); }
void ref_get_stress(void *h, int gp, double *sig) { P(h)->get_stress(gp, sig); }
void ref_get_ctan(void *h, int gp, double *c) { P(h)->get_ctan(gp, c); }
void ref_homogenize(void *h) { P(h)->homogenize(); }
void ref_homogenize_linear(void *h) { P(h)->homogenize_linear(); }
void ref_update_vars(void *h) { P(h)->update_vars(); }
int ref_is_non_linear(void *h, int gp) { return P(h)->is_non_linear(gp); }
int ref_get_non_linear_gps(void *h) { return P(h)->get_non_linear_gps(); }
int ref_get_cost(void *h, int gp) { return P(h)->get_cost(gp); }
int ref_has_converged(void *h, int gp) { return P(h)->has_converged(gp) ? 1 : 0; }
int ref_has_subiterated(void *h, int gp) { return P(h)->has_subiterated(gp) ? 1 : 0; }
void ref_write_restart(void *h, int id) { P(h)->write_restart(id); }
void ref_read_restart(void *h, int id) { P(h)->read_restart(id); }
void ref_output(void *h, int gp, const char *fname) { P(h)->output(gp, fname); }

/* ---- state inspection ---- */
int ref_nelem(void *h) { return P(h)->p_nelem(); }
int ref_nndim(void *h) { return P(h)->p_nndim(); }
int ref_nvars(void *h) { return P(h)->p_nvars(); }
double ref_wg(void *h) { return P(h)->p_wg(); }
void ref_get_elem_type(void *h, int *out) { memcpy(out, P(h)->p_elem_type(), sizeof(int) * P(h)->p_nelem()); }
void ref_get_bmat(void *h, double *out) { memcpy(out, P(h)->p_bmat(), sizeof(double) * 8 * 6 * 24); }
void ref_get_ctan_lin(void *h, double *out) { memcpy(out, P(h)->p_ctan_lin(), sizeof(double) * 36); }
/* which: 0 = u_n, 1 = u_k ; returns 0 when the buffer does not exist */
int ref_get_u(void *h, int gp, int which, double *out) {
  gp_t<3> *g = P(h)->p_gp(gp);
  const double *src = which ? g->u_k : g->u_n;
  if (!src) return 0;
  memcpy(out, src, sizeof(double) * g->nndim);
  return 1;
}
/* which: 0 = vars_n, 1 = vars_k ; returns 0 when the GP has no internal variables yet */
int ref_get_vars(void *h, int gp, int which, double *out) {
  gp_t<3> *g = P(h)->p_gp(gp);
  if (!g->allocated) return 0;
  memcpy(out, which ? g->vars_k : g->vars_n, sizeof(double) * g->nvars);
  return 1;
}

/* ---- index structures ---- */
void ref_elem_nodes(int nx, int ny, int ex, int ey, int ez, int *n8) { get_elem_nodes(n8, nx, ny, ex, ey, ez); }

/* 3-D ELL column table exactly as ell_init builds it (src/ell-common.cpp:34-139) */
void ref_ell_cols(int nx, int ny, int nz, int *cols) {
  ell_matrix A;
  const int ns[3] = {nx, ny, nz};
  ell_init(&A, 3, 3, ns);
  memcpy(cols, A.cols, sizeof(int) * (size_t)A.nrow * A.nnz);
  ell_free(&A);
}

/* scatter one 24x24 element matrix into a zeroed ELL and return vals (pins cols_row, src/ell-common.cpp:166-198) */
void ref_ell_add_one(int nx, int ny, int nz, int ex, int ey, int ez, const double *Ae, double *vals) {
  ell_matrix A;
  const int ns[3] = {nx, ny, nz};
  ell_init(&A, 3, 3, ns);
  ell_set_zero_mat(&A);
  ell_add_3D(&A, ex, ey, ez, Ae);
  memcpy(vals, A.vals, sizeof(double) * (size_t)A.nrow * A.nnz);
  ell_free(&A);
}

/* ---- FE stages (protected members) ---- */
void ref_set_displ_bc(void *h, const double *eps, double *u) { P(h)->p_set_displ_bc(eps, u); }
double ref_assembly_rhs(void *h, const double *u, const double *vars, double *b) {
  return P(h)->p_assembly_rhs(u, vars, b);
}
void ref_assembly_mat(void *h, const double *u, const double *vars, double *vals) {
  ref_probe *p = P(h);
  ell_matrix A;
  const int ns[3] = {p->p_nx(), p->p_ny(), p->p_nz()};
  ell_init(&A, 3, 3, ns);
  p->p_assembly_mat(&A, u, vars);
  memcpy(vals, A.vals, sizeof(double) * (size_t)A.nrow * A.nnz);
  ell_free(&A);
}
void ref_ave_stress(void *h, const double *u, const double *vars, double *sig) { P(h)->p_ave_stress(u, sig, vars); }
void ref_ave_strain(void *h, const double *u, double *eps) { P(h)->p_ave_strain(u, eps); }
int ref_vars_new(void *h, const double *u, const double *vo, double *vn) { return P(h)->p_vars_new(u, vo, vn) ? 1 : 0; }

/* y = A x with the reference SpMV (src/ell.cpp:35-44) on caller-provided vals (reference layout) */
void ref_ell_mvp(int nx, int ny, int nz, const double *vals, const double *x, double *y) {
  ell_matrix A;
  const int ns[3] = {nx, ny, nz};
  ell_init(&A, 3, 3, ns);
  memcpy(A.vals, vals, sizeof(double) * (size_t)A.nrow * A.nnz);
  ell_mvp(&A, x, y);
  ell_free(&A);
}

/* Jacobi-PCG (src/ell.cpp:66-122) on caller-provided vals; returns iterations, *err = final r.z */
int ref_ell_solve_cgpd(int nx, int ny, int nz, const double *vals, const double *b, double *x, double *err) {
  ell_matrix A;
  const int ns[3] = {nx, ny, nz};
  ell_init(&A, 3, 3, ns);
  memcpy(A.vals, vals, sizeof(double) * (size_t)A.nrow * A.nnz);
  const int its = ell_solve_cgpd(&A, b, x, err);
  ell_free(&A);
  return its;
}

/* one Newton-Raphson solve (src/solve.cpp:29-82); u is in/out; out3 = {its, solver_its, converged} */
void ref_newton(void *h, const double *eps, const double *vars, double *u, int *out3) {
  ref_probe *p = P(h);
  ell_matrix A;
  const int ns[3] = {p->p_nx(), p->p_ny(), p->p_nz()};
  ell_init(&A, 3, 3, ns);
  const int nd = p->p_nndim();
  double *b = (double *)calloc(nd, sizeof(double));
  double *du = (double *)calloc(nd, sizeof(double));
  newton_t r = p->p_newton(&A, b, u, du, eps, vars);
  out3[0] = r.its;
  out3[1] = r.solver_its;
  out3[2] = r.converged ? 1 : 0;
  free(b);
  free(du);
  ell_free(&A);
}

/* ---- material laws (src/material.cpp) through the reference factory ---- */
static material_t *mk(int type, double E, double nu, double Ka, double Sy, double Xt) {
  material_base mb;
  material_set(&mb, type, E, nu, Ka, Sy, Xt);
  return material_t::make_material(mb);
}
void ref_mat_stress(int type, double E, double nu, double Ka, double Sy, double Xt, const double *eps,
                    const double *vars, double *sig) {
  material_t *m = mk(type, E, nu, Ka, Sy, Xt);
  m->get_stress(eps, sig, vars);
  delete m;
}
void ref_mat_ctan(int type, double E, double nu, double Ka, double Sy, double Xt, const double *eps,
                  const double *vars, double *c36) {
  material_t *m = mk(type, E, nu, Ka, Sy, Xt);
  m->get_ctan(eps, c36, vars);
  delete m;
}
int ref_mat_evolute(int type, double E, double nu, double Ka, double Sy, double Xt, const double *eps,
                    const double *vars_old, double *vars_new) {
  material_t *m = mk(type, E, nu, Ka, Sy, Xt);
  const bool nl = m->evolute(eps, vars_old, vars_new);
  delete m;
  return nl ? 1 : 0;
}

/* util.hpp mvp<double,3> known-answer hook (test/test_util_1.cpp:35-56) */
void ref_mvp3(const double *m9, const double *x3, double *y3) {
  double m[3][3];
  memcpy(m, m9, sizeof(m));
  mvp<double, 3>(m, x3, y3);
}

#ifdef _OPENMP
int ref_omp_max_threads(void) { return omp_get_max_threads(); }
#else
int ref_omp_max_threads(void) { return 1; }
#endif
}
