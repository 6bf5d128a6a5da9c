// fe_math.cuh -- hex8 / Gauss-point arithmetic shared by every CUDA kernel of the hot path.
//
// Everything here is `__host__ __device__` so that the same expressions can be unit-tested on
// the CPU (tests/hostcheck) against the oracle before any GPU time is spent; the product only
// ever calls them from kernels.
//
// Parity notes (reference = gagiuntoli/Micropp, file:line under /root/reference):
//   * material laws follow src/material.cpp:49-307 operation by operation and are evaluated with
//     non-contracted IEEE arithmetic (mul_/add_/sub_ below): the forward-difference tangent
//     (src/material.cpp:49-63) amplifies rounding by 1/D_EPS_CTAN = 1e8, so FMA contraction inside
//     a stress evaluation would show up at ~1e-10 in the Jacobian.
//   * strains follow src/common.cpp:58-72 (sum over the 24 element dofs in ascending order); the
//     structurally-zero B entries (src/micro3D.cpp:100-119) are skipped, which is exact.
#pragma once

#include <math.h>

#ifdef __CUDACC__
#define MPP_HD __host__ __device__ __forceinline__
#else
#define MPP_HD inline
#endif

// include/material_base.h:27-29
#define MPP_D_EPS_CTAN 1.0e-8
#define MPP_SQRT_2DIV3 0.816496581
// include/params.hpp:34
#define MPP_CONSTXG 0.577350269189626

enum { MPP_ELASTIC = 0, MPP_PLASTIC = 1, MPP_DAMAGE = 2 };

// Same field order as `struct material_base` (include/material_base.h:38-43).
struct mpp_material {
  double E, nu, Ka, Sy;
  double k, mu, lambda;
  double Xt;
  int type;
};

// ---- non-contracted arithmetic -------------------------------------------------------------
MPP_HD double mul_(double a, double b) {
#ifdef __CUDA_ARCH__
  return __dmul_rn(a, b);
#else
  return a * b;
#endif
}
MPP_HD double add_(double a, double b) {
#ifdef __CUDA_ARCH__
  return __dadd_rn(a, b);
#else
  return a + b;
#endif
}
MPP_HD double sub_(double a, double b) {
#ifdef __CUDA_ARCH__
  return __dsub_rn(a, b);
#else
  return a - b;
#endif
}
MPP_HD double div_(double a, double b) {
#ifdef __CUDA_ARCH__
  return __ddiv_rn(a, b);
#else
  return a / b;
#endif
}
MPP_HD double sqrt_(double a) {
#ifdef __CUDA_ARCH__
  return __dsqrt_rn(a);
#else
  return sqrt(a);
#endif
}

// ---- hex8 topology ---------------------------------------------------------------------------
// Local node order of src/common.cpp:30-41 : (0,0,0),(1,0,0),(1,1,0),(0,1,0),(0,0,1),(1,0,1),(1,1,1),(0,1,1).
MPP_HD constexpr int corner_x(int a) { return ((a & 3) == 1 || (a & 3) == 2) ? 1 : 0; }
MPP_HD constexpr int corner_y(int a) { return ((a & 3) >= 2) ? 1 : 0; }
MPP_HD constexpr int corner_z(int a) { return (a >> 2) & 1; }
MPP_HD constexpr int corner_of(int lx, int ly, int lz) { return lz * 4 + (ly ? (3 - lx) : lx); }
// 27-point stencil slot of the neighbour at offset (di,dj,dk) (src/ell-common.cpp:102-130).
MPP_HD constexpr int nbr_slot(int di, int dj, int dk) { return (dk + 1) * 9 + (dj + 1) * 3 + (di + 1); }
// Slot that local node j occupies in the ELL row of local node a: equals the literal table
// cols_row[8][8] of src/ell-common.cpp:175-178 (checked in tests against the oracle scatter).
MPP_HD constexpr int cols_row(int a, int j) {
  return nbr_slot(corner_x(j) - corner_x(a), corner_y(j) - corner_y(a), corner_z(j) - corner_z(a));
}

// ---- strain at a Gauss point -------------------------------------------------------------------
// dsh[a*3+d] = dN_a/dx_d at this Gauss point (src/micro3D.cpp:82-98); ue[a*3+d] element dofs.
// eps = B ue with the B layout of src/micro3D.cpp:100-119, summed as src/common.cpp:67-71.
template <typename DSH>
MPP_HD void gp_strain(const DSH dsh, const double *ue, double *eps) {
  double e0 = 0.0, e1 = 0.0, e2 = 0.0, e3 = 0.0, e4 = 0.0, e5 = 0.0;
#pragma unroll
  for (int a = 0; a < 8; ++a) {
    const double gx = dsh[a * 3 + 0], gy = dsh[a * 3 + 1], gz = dsh[a * 3 + 2];
    const double ux = ue[a * 3 + 0], uy = ue[a * 3 + 1], uz = ue[a * 3 + 2];
    e0 = add_(e0, mul_(gx, ux));
    e1 = add_(e1, mul_(gy, uy));
    e2 = add_(e2, mul_(gz, uz));
    e3 = add_(add_(e3, mul_(gy, ux)), mul_(gx, uy));
    e4 = add_(add_(e4, mul_(gz, ux)), mul_(gx, uz));
    e5 = add_(add_(e5, mul_(gz, uy)), mul_(gy, uz));
  }
  eps[0] = e0;
  eps[1] = e1;
  eps[2] = e2;
  eps[3] = e3;
  eps[4] = e4;
  eps[5] = e5;
}

// ---- material laws -----------------------------------------------------------------------------
// Linear isotropic stress (src/material.cpp:76-81).
MPP_HD void elastic_stress(const mpp_material &m, const double *eps, double *sig) {
  const double tr = add_(add_(eps[0], eps[1]), eps[2]);
  const double ltr = mul_(m.lambda, tr);
  const double mu2 = mul_(2.0, m.mu);
#pragma unroll
  for (int i = 0; i < 3; ++i) sig[i] = add_(ltr, mul_(mu2, eps[i]));
#pragma unroll
  for (int i = 3; i < 6; ++i) sig[i] = mul_(m.mu, eps[i]);
}

// Isotropic tangent (src/material.cpp:84-94), row-major 6x6.
MPP_HD void elastic_ctan(const mpp_material &m, double *c) {
#pragma unroll
  for (int i = 0; i < 36; ++i) c[i] = 0.0;
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) c[i * 6 + j] += m.lambda;
#pragma unroll
  for (int i = 0; i < 3; ++i) c[i * 6 + i] = add_(c[i * 6 + i], mul_(2.0, m.mu));
#pragma unroll
  for (int i = 3; i < 6; ++i) c[i * 6 + i] = m.mu;
}

// Deviator of the normal components only (src/material.cpp:65-69).
MPP_HD void dev_part(const double *t, double *d) {
  const double third_tr = mul_(1 / 3.0, add_(add_(t[0], t[1]), t[2]));
#pragma unroll
  for (int i = 0; i < 3; ++i) d[i] = sub_(t[i], third_tr);
#pragma unroll
  for (int i = 3; i < 6; ++i) d[i] = t[i];
}

// J2 radial return with linear isotropic hardening (src/material.cpp:111-148).
// vars = {eps_p[6], alpha} or nullptr (=> zeros).  Returns true when yielding.
MPP_HD bool plastic_law(const mpp_material &m, const double *eps, const double *vars, double *dl, double *normal,
                        double *s_trial) {
  double epsp[6];
  double alpha = 0.0;
#pragma unroll
  for (int i = 0; i < 6; ++i) epsp[i] = vars ? vars[i] : 0.0;
  if (vars) alpha = vars[6];

  double eps_dev[6], epsp_dev[6];
  dev_part(epsp, epsp_dev);
  dev_part(eps, eps_dev);

  const double mu2 = mul_(2.0, m.mu);
#pragma unroll
  for (int i = 0; i < 3; ++i) s_trial[i] = mul_(mu2, sub_(eps_dev[i], epsp_dev[i]));
#pragma unroll
  for (int i = 3; i < 6; ++i) s_trial[i] = mul_(m.mu, sub_(eps_dev[i], epsp_dev[i]));

  double tmp = 0.0;
#pragma unroll
  for (int i = 0; i < 6; ++i) tmp = add_(tmp, mul_(s_trial[i], s_trial[i]));
  const double s_norm = sqrt_(tmp);

  const double f_trial = sub_(s_norm, mul_(MPP_SQRT_2DIV3, add_(m.Sy, mul_(m.Ka, alpha))));

  if (f_trial > 0) {
#pragma unroll
    for (int i = 0; i < 6; ++i) normal[i] = div_(s_trial[i], s_norm);
    *dl = div_(f_trial, mul_(mu2, add_(1., div_(m.Ka, mul_(3., m.mu)))));
    return true;
  }
#pragma unroll
  for (int i = 0; i < 6; ++i) normal[i] = 0.0;
  *dl = 0.0;
  return false;
}

// src/material.cpp:151-164
MPP_HD void plastic_stress(const mpp_material &m, const double *eps, const double *vars, double *sig) {
  double dl, normal[6], s_trial[6];
  plastic_law(m, eps, vars, &dl, normal, s_trial);
  const double ktr = mul_(m.k, add_(add_(eps[0], eps[1]), eps[2]));
  const double f = mul_(mul_(2.0, m.mu), dl);
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    double s = s_trial[i];
    if (i < 3) s = add_(s, ktr);
    sig[i] = sub_(s, mul_(f, normal[i]));
  }
}

// src/material.cpp:171-186.  NB: nothing is written when vars_old == nullptr (reference quirk).
MPP_HD bool plastic_evolute(const mpp_material &m, const double *eps, const double *vars_old, double *vars_new) {
  double dl, normal[6], s_trial[6];
  const bool nl = plastic_law(m, eps, vars_old, &dl, normal, s_trial);
  if (vars_old && vars_new) {
#pragma unroll
    for (int i = 0; i < 6; ++i) vars_new[i] = add_(vars_old[i], mul_(dl, normal[i]));
    vars_new[6] = add_(add_(vars_old[6], mul_(MPP_SQRT_2DIV3, dl)), 0.0);
  }
  return nl;
}

// src/material.cpp:206-225 (Ey, H0, H1 are hard-coded in the reference).
MPP_HD double damage_hardening(const mpp_material &m, double r) {
  const double Ey = 10.0e4;
  const double inf_Ey = mul_(10., Ey);
  const double H0 = 10.0, H1 = 5.0;
  const double sE = sqrt_(m.E);
  const double r0 = div_(Ey, sE);
  const double q0 = r0;
  const double q1 = div_(inf_Ey, sE);
  const double r1 = add_(r0, div_(sub_(q1, q0), H0));
  if (r < r0) return 0.0;
  if (r >= r0 && r < r1) return add_(q0, mul_(H0, sub_(r, r0)));
  return add_(q1, mul_(H1, sub_(r, r1)));
}

// src/material.cpp:228-269.  sig_lin receives the undamaged stress.
MPP_HD bool damage_law(const mpp_material &m, const double *eps, double r_old, double D_old, double *r_new,
                       double *D_new, double *sig_lin) {
  elastic_stress(m, eps, sig_lin);
  double product = 0.0;
#pragma unroll
  for (int i = 0; i < 6; ++i) product = add_(product, mul_(sig_lin[i], eps[i]));
  const double r = (product >= 0) ? sqrt_(product) : 0;
  const double r_thr = div_(m.Xt, sqrt_(m.E));
  const double r_old_a = (r_old < r_thr) ? r_thr : r_old;
  if (r <= r_old_a) {
    *r_new = r_old_a;
    *D_new = D_old;
    return false;
  }
  const double q = damage_hardening(m, r);
  *r_new = r;
  *D_new = sub_(1., div_(q, r));
  return true;
}

// src/material.cpp:272-287
MPP_HD void damage_stress(const mpp_material &m, const double *eps, const double *vars, double *sig) {
  const double r_old = vars ? vars[0] : 0.0;
  const double D_old = vars ? vars[1] : 0.0;
  double D, r;
  damage_law(m, eps, r_old, D_old, &r, &D, sig);
  const double w = sub_(1, D);
#pragma unroll
  for (int i = 0; i < 6; ++i) sig[i] = mul_(sig[i], w);
}

// src/material.cpp:294-307
MPP_HD bool damage_evolute(const mpp_material &m, const double *eps, const double *vars_old, double *vars_new) {
  const double r_old = vars_old ? vars_old[0] : 0.0;
  const double D_old = vars_old ? vars_old[1] : 0.0;
  double sig[6], r, D;
  const bool nl = damage_law(m, eps, r_old, D_old, &r, &D, sig);
  if (vars_new) {
    vars_new[0] = r;
    vars_new[1] = D;
  }
  return nl;
}

// Dispatch on the POD type tag (replaces the virtual calls of include/material.hpp:36-63).
MPP_HD void mat_stress(const mpp_material &m, const double *eps, const double *vars, double *sig) {
  if (m.type == MPP_ELASTIC)
    elastic_stress(m, eps, sig);
  else if (m.type == MPP_PLASTIC)
    plastic_stress(m, eps, vars, sig);
  else
    damage_stress(m, eps, vars, sig);
}

// Forward-difference tangent (src/material.cpp:49-63) for plastic/damage; closed form for elastic.
MPP_HD void mat_ctan(const mpp_material &m, const double *eps, const double *vars, double *c) {
  if (m.type == MPP_ELASTIC) {
    elastic_ctan(m, c);
    return;
  }
  double s0[6];
  mat_stress(m, eps, vars, s0);
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    double e1[6];
#pragma unroll
    for (int q = 0; q < 6; ++q) e1[q] = eps[q];
    e1[i] = add_(e1[i], MPP_D_EPS_CTAN);
    double s1[6];
    mat_stress(m, e1, vars, s1);
#pragma unroll
    for (int j = 0; j < 6; ++j) c[j * 6 + i] = div_(sub_(s1[j], s0[j]), MPP_D_EPS_CTAN);
  }
}

MPP_HD bool mat_evolute(const mpp_material &m, const double *eps, const double *vars_old, double *vars_new) {
  if (m.type == MPP_ELASTIC) return false;
  if (m.type == MPP_PLASTIC) return plastic_evolute(m, eps, vars_old, vars_new);
  return damage_evolute(m, eps, vars_old, vars_new);
}

// Number of internal variables a material actually carries (reference always stores NUM_VAR_GP = 7,
// include/params.hpp:28; plastic uses 7, damage 2, elastic 0).
MPP_HD constexpr int mat_nvar(int type) { return type == MPP_PLASTIC ? 7 : (type == MPP_DAMAGE ? 2 : 0); }

// ---- boundary displacement u = eps_bar . x (src/micro3D.cpp:27-78) ---------------------------
// The six face loops of the reference overwrite each other in the order z0,z1,y0,y1,x0,x1, and the
// "max" faces use the literal lengths lx=ly=lz=1.0 instead of (n-1)*d.  The winning face decides
// which coordinate expression each component sees.
MPP_HD void bc_coords(int i, int j, int k, int nx, int ny, int nz, double dx, double dy, double dz, double *c) {
  double cx = i * dx, cy = j * dy, cz = k * dz;
  if (i == nx - 1) {
    cx = 1.0;
  } else if (i == 0) {
    cx = 0;
  } else if (j == ny - 1) {
    cy = 1.0;
  } else if (j == 0) {
    cy = 0;
  } else if (k == nz - 1) {
    cz = 1.0;
  } else if (k == 0) {
    cz = 0;
  }
  c[0] = cx;
  c[1] = cy;
  c[2] = cz;
}

MPP_HD void bc_displacement(const double *eps, const double *c, double *u3) {
  const double e[3][3] = {{eps[0], mul_(0.5, eps[3]), mul_(0.5, eps[4])},
                          {mul_(0.5, eps[3]), eps[1], mul_(0.5, eps[5])},
                          {mul_(0.5, eps[4]), mul_(0.5, eps[5]), eps[2]}};
#pragma unroll
  for (int r = 0; r < 3; ++r) {  // mvp<double,3>, include/util.hpp:58-66
    double t = 0.0;
#pragma unroll
    for (int q = 0; q < 3; ++q) t = add_(t, mul_(e[r][q], c[q]));
    u3[r] = t;
  }
}
