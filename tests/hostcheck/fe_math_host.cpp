// tests/hostcheck/fe_math_host.cpp -- TEST-ONLY host build of the inline device math
// (micropp_b200/csrc/fe_math.cuh) so that the arithmetic the CUDA kernels execute can be checked
// against the oracle on a machine without a GPU.  Never loaded by the product.
#include <cstring>

#include "fe_math.cuh"

static mpp_material mk(int type, double E, double nu, double Ka, double Sy, double Xt) {
  mpp_material m;
  m.type = type;
  m.E = E;
  m.nu = nu;
  m.Ka = Ka;
  m.Sy = Sy;
  m.Xt = Xt;
  m.k = E / (3. * (1. - 2. * nu));
  m.mu = E / (2. * (1. + nu));
  m.lambda = nu * E / ((1. + nu) * (1. - 2. * nu));
  return m;
}

extern "C" {
void hc_mat_stress(int type, double E, double nu, double Ka, double Sy, double Xt, const double *eps,
                   const double *vars, double *sig) {
  mat_stress(mk(type, E, nu, Ka, Sy, Xt), eps, vars, sig);
}
void hc_mat_ctan(int type, double E, double nu, double Ka, double Sy, double Xt, const double *eps, const double *vars,
                 double *c) {
  mat_ctan(mk(type, E, nu, Ka, Sy, Xt), eps, vars, c);
}
int hc_mat_evolute(int type, double E, double nu, double Ka, double Sy, double Xt, const double *eps,
                   const double *vars_old, double *vars_new) {
  const mpp_material m = mk(type, E, nu, Ka, Sy, Xt);
  // same write rule as k_vars_new
  const bool wr = (m.type == MPP_DAMAGE) || (vars_old != nullptr);
  return mat_evolute(m, eps, vars_old, wr ? vars_new : nullptr) ? 1 : 0;
}
void hc_strain(const double *dsh24, const double *ue24, double *eps6) { gp_strain(dsh24, ue24, eps6); }
void hc_bc(int i, int j, int k, int nx, int ny, int nz, const double *eps, double *u3) {
  double c[3];
  bc_coords(i, j, k, nx, ny, nz, 1.0 / (nx - 1), 1.0 / (ny - 1), 1.0 / (nz - 1), c);
  bc_displacement(eps, c, u3);
}
int hc_cols_row(int a, int j) { return cols_row(a, j); }
int hc_corner_of(int x, int y, int z) { return corner_of(x, y, z); }
}
