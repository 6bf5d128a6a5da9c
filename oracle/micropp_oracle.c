/*
 * oracle/micropp_oracle.c -- TEST INFRASTRUCTURE ONLY (never linked into, imported by or shipped with
 * the product; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may use it).
 *
 * A plain-C99 CPU restatement of the reference's RVE-homogenization hot path
 * (gagiuntoli/Micropp; file:line citations are relative to /root/reference), serial, FP64, written
 * to be compiled with `gcc -O2 -ffp-contract=off` so that every floating-point operation happens in
 * the order the reference performs it.
 *
 * PARITY PINNED: tests/test_oracle.py checks this file against
 *   - the golden tables of the reference's own tests (test/benchmark-elastic.cpp:40-51,
 *     test/benchmark-plastic.cpp:40-51, test/benchmark-damage.cpp:40-51, test/test_get_elem_nodes.cpp:63-85,
 *     test/test_util_1.cpp:35-56, test/test3d_1.cpp known answers recorded in SURVEY.md appendix B),
 *   - the committed fixtures under tests/golden/ (generated from the compiled reference by
 *     tests/golden/make_golden.py), and
 *   - the compiled reference itself (oracle/_ref/libmicropp_ref.so) on seeded random inputs, whenever
 *     that file is present.
 *
 * Layouts are the reference's: u[node*3+d]; vars[e*56+gp*7+v] (include/params.hpp:41);
 * ELL values vals[row*81+slot], row = node*3+fi, slot = nbr*3+fj (src/ell-common.cpp:166-198).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define NPE 8
#define NVOI 6
#define NNZ 81
#define D_EPS_CTAN 1.0e-8      /* include/material_base.h:29 */
#define SQRT_2DIV3 0.816496581 /* include/material_base.h:30 (nine digits in the reference) */
#define CONSTXG 0.577350269189626 /* include/params.hpp:34 */
#define CG_ABS_TOL 1.0e-50     /* include/ell.hpp:33-35 */
#define CG_MAX_ITS 1000
#define CG_REL_TOL 1.0e-5

enum { ORC_ELASTIC = 0, ORC_PLASTIC = 1, ORC_DAMAGE = 2 };

typedef struct orc_material {
  double E, nu, Ka, Sy, k, mu, lambda, Xt;
  int type;
} orc_material;

typedef struct orc_problem {
  int nx, ny, nz, nn, nndim, nex, ney, nez, nelem;
  double dx, dy, dz, wg;
  double bmat[NPE][NVOI][NPE * 3];
  const int *elem_type; /* [nelem], borrowed */
  orc_material mat[3];
  int nr_max_its;
  double nr_max_tol, nr_rel_tol;
} orc_problem;

/* ------------------------------------------------------------------------------------------ indices */

/* src/common.cpp:30-41 */
void orc_elem_nodes(int nx, int ny, int ex, int ey, int ez, int n[8]) {
  const int plane = nx * ny;
  const int base = ez * plane + ey * nx + ex;
  const int quad[4] = {base, base + 1, base + nx + 1, base + nx};
  for (int a = 0; a < 4; ++a) {
    n[a] = quad[a];
    n[a + 4] = quad[a] + plane;
  }
}

/* 8-colour structured ordering (defined by this project): no two elements of one colour
 * share a node. */
int orc_elem_colour(int ex, int ey, int ez) { return (ex & 1) + 2 * (ey & 1) + 4 * (ez & 1); }

/* src/ell-common.cpp:86-137: explicit column table of the 27-point / 3-field ELL matrix; neighbours outside
 * the grid point at node 0. */
void orc_ell_cols(int nx, int ny, int nz, int *cols) {
  const int plane = nx * ny;
  for (int zi = 0; zi < nz; ++zi)
    for (int yi = 0; yi < ny; ++yi)
      for (int xi = 0; xi < nx; ++xi) {
        const int ni = zi * plane + yi * nx + xi;
        for (int fi = 0; fi < 3; ++fi) {
          int *row = cols + ((size_t)ni * 3 + fi) * NNZ;
          for (int dz = -1; dz <= 1; ++dz)
            for (int dy = -1; dy <= 1; ++dy)
              for (int dx = -1; dx <= 1; ++dx) {
                const int slot = (dz + 1) * 9 + (dy + 1) * 3 + (dx + 1);
                const int x = xi + dx, y = yi + dy, z = zi + dz;
                const int inside = x >= 0 && x < nx && y >= 0 && y < ny && z >= 0 && z < nz;
                const int nb = inside ? ni + dz * plane + dy * nx + dx : 0;
                for (int fj = 0; fj < 3; ++fj) row[slot * 3 + fj] = nb * 3 + fj;
              }
        }
      }
}

/* slot of local node j in the row of local node i: the literal table of src/ell-common.cpp:175-178 */
static const int k_cols_row[8][8] = {{13, 14, 17, 16, 22, 23, 26, 25}, {12, 13, 16, 15, 21, 22, 25, 24},
                                     {9, 10, 13, 12, 18, 19, 22, 21},  {10, 11, 14, 13, 19, 20, 23, 22},
                                     {4, 5, 8, 7, 13, 14, 17, 16},     {3, 4, 7, 6, 12, 13, 16, 15},
                                     {0, 1, 4, 3, 9, 10, 13, 12},      {1, 2, 5, 4, 10, 11, 14, 13}};
int orc_cols_row(int i, int j) { return k_cols_row[i][j]; }

static int on_boundary(const orc_problem *P, int n) {
  const int k = n / (P->nx * P->ny), r = n % (P->nx * P->ny), j = r / P->nx, i = r % P->nx;
  return i == 0 || i == P->nx - 1 || j == 0 || j == P->ny - 1 || k == 0 || k == P->nz - 1;
}

/* ------------------------------------------------------------------------------------------ set-up */

/* src/material.c:26-38 + include/material.hpp:69-75 */
void orc_material_set(orc_material *m, int type, double E, double nu, double Ka, double Sy, double Xt) {
  m->type = type;
  m->E = E;
  m->nu = nu;
  m->Ka = Ka;
  m->Sy = Sy;
  m->Xt = Xt;
  m->k = E / (3. * (1. - 2. * nu));
  m->mu = E / (2. * (1. + nu));
  m->lambda = nu * E / ((1. + nu) * (1. - 2. * nu));
}

/* src/micro3D.cpp:81-120 with the Gauss points of include/micropp.hpp:86-89 */
static void fill_bmat(orc_problem *P) {
  static const double sgn[8][3] = {{-1, -1, -1}, {+1, -1, -1}, {+1, +1, -1}, {-1, +1, -1},
                                   {-1, -1, +1}, {+1, -1, +1}, {+1, +1, +1}, {-1, +1, +1}};
  for (int gp = 0; gp < 8; ++gp) {
    const double xg[3] = {sgn[gp][0] * CONSTXG, sgn[gp][1] * CONSTXG, sgn[gp][2] * CONSTXG};
    memset(P->bmat[gp], 0, sizeof(P->bmat[gp]));
    for (int a = 0; a < 8; ++a) {
      /* the reference writes e.g. -(1 - xg1)*(1 - xg2)/8.*2./dx; (1 + s*x) reproduces (1 - x) exactly for s=-1 */
      const double fx = 1 + sgn[a][0] * xg[0], fy = 1 + sgn[a][1] * xg[1], fz = 1 + sgn[a][2] * xg[2];
      const double gx = sgn[a][0] * (fy * fz) / 8. * 2. / P->dx;
      const double gy = sgn[a][1] * (fx * fz) / 8. * 2. / P->dy;
      const double gz = sgn[a][2] * (fx * fy) / 8. * 2. / P->dz;
      double(*B)[24] = P->bmat[gp];
      B[0][a * 3 + 0] = gx;
      B[1][a * 3 + 1] = gy;
      B[2][a * 3 + 2] = gz;
      B[3][a * 3 + 0] = gy;
      B[3][a * 3 + 1] = gx;
      B[4][a * 3 + 0] = gz;
      B[4][a * 3 + 2] = gx;
      B[5][a * 3 + 1] = gz;
      B[5][a * 3 + 2] = gy;
    }
  }
}

/* mesh constants of src/micropp.cpp:30-62 */
orc_problem *orc_new(int nx, int ny, int nz, const int *elem_type, const orc_material *mats, int nr_max_its,
                     double nr_max_tol, double nr_rel_tol) {
  orc_problem *P = (orc_problem *)calloc(1, sizeof(orc_problem));
  P->nx = nx;
  P->ny = ny;
  P->nz = nz;
  P->nn = nx * ny * nz;
  P->nndim = 3 * P->nn;
  P->nex = nx - 1;
  P->ney = ny - 1;
  P->nez = nz - 1;
  P->nelem = P->nex * P->ney * P->nez;
  P->dx = 1.0 / P->nex;
  P->dy = 1.0 / P->ney;
  P->dz = 1.0 / P->nez;
  P->wg = (P->dx * P->dy * P->dz) / NPE;
  P->elem_type = elem_type;
  memcpy(P->mat, mats, 3 * sizeof(orc_material));
  P->nr_max_its = nr_max_its;
  P->nr_max_tol = nr_max_tol;
  P->nr_rel_tol = nr_rel_tol;
  fill_bmat(P);
  return P;
}
void orc_free(orc_problem *P) { free(P); }
void orc_get_bmat(const orc_problem *P, double *out) { memcpy(out, P->bmat, sizeof(P->bmat)); }

/* include/util.hpp:115-143 + src/micropp.cpp:339-372: the micro-structures used by the BASELINE configs
 * (0 homogeneous, 1 sphere, 2 layer_y, 3/4 cylindrical fibre along x/z).  Returns -1 for the others. */
static double norm3(const double v[3]) {
  double acc = 0;
  for (int i = 0; i < 3; ++i) acc += v[i] * v[i];
  return sqrt(acc);
}
static int inside_cyl(const double dir[3], const double c[3], double rad, const double p[3]) {
  const double d[3] = {p[0] - c[0], p[1] - c[1], p[2] - c[2]};
  double along = 0;
  for (int i = 0; i < 3; ++i) along += dir[i] * d[i];
  const double cs = along / (norm3(dir) * norm3(d));
  return norm3(d) * sqrt(1 - cs * cs) <= rad; /* NaN on the axis => outside, as in the reference */
}
int orc_elem_type(int micro_type, const double geo[4], int nx, int ny, int nz, int ex, int ey, int ez) {
  const double dx = 1.0 / (nx - 1), dy = 1.0 / (ny - 1), dz = 1.0 / (nz - 1);
  const double p[3] = {ex * dx + dx / 2., ey * dy + dy / 2., ez * dz + dz / 2.};
  const double mid[3] = {1.0 / 2, 1.0 / 2, 1.0 / 2};
  const double ax[3] = {1, 0, 0}, az[3] = {0, 0, 1};
  switch (micro_type) {
    case 0: return 0;
    case 1: {
      const double d[3] = {p[0] - mid[0], p[1] - mid[1], p[2] - mid[2]};
      return norm3(d) < geo[0];
    }
    case 2: return p[1] < geo[0];
    case 3: return inside_cyl(ax, mid, geo[0], p);
    case 4: return inside_cyl(az, mid, geo[0], p);
    default: return -1;
  }
}

/* ------------------------------------------------------------------------------------------ materials */

/* src/material.cpp:76-81 */
static void stress_elastic(const orc_material *m, const double *e, double *s) {
  for (int i = 0; i < 3; ++i) s[i] = m->lambda * (e[0] + e[1] + e[2]) + 2 * m->mu * e[i];
  for (int i = 3; i < 6; ++i) s[i] = m->mu * e[i];
}

/* src/material.cpp:65-69 */
static void deviator(const double t[6], double d[6]) {
  memcpy(d, t, 6 * sizeof(double));
  for (int i = 0; i < 3; ++i) d[i] -= (1 / 3.0) * (t[0] + t[1] + t[2]);
}

/* src/material.cpp:111-148 */
static int plastic_law(const orc_material *m, const double *eps, const double *vars, double *dl, double nrm[6],
                       double st[6]) {
  const double zero[6] = {0, 0, 0, 0, 0, 0};
  const double alpha = vars ? vars[6] : 0;
  const double *ep = vars ? vars : zero;
  double ed[6], epd[6];
  deviator(ep, epd);
  deviator(eps, ed);
  for (int i = 0; i < 3; ++i) st[i] = 2 * m->mu * (ed[i] - epd[i]);
  for (int i = 3; i < 6; ++i) st[i] = m->mu * (ed[i] - epd[i]);
  double acc = 0.0;
  for (int i = 0; i < 6; ++i) acc += st[i] * st[i];
  const double sn = sqrt(acc);
  const double f = sn - SQRT_2DIV3 * (m->Sy + m->Ka * alpha);
  if (f > 0) {
    for (int i = 0; i < 6; ++i) nrm[i] = st[i] / sn;
    *dl = f / (2. * m->mu * (1. + m->Ka / (3. * m->mu)));
    return 1;
  }
  memset(nrm, 0, 6 * sizeof(double));
  *dl = 0;
  return 0;
}

/* src/material.cpp:151-164 */
static void stress_plastic(const orc_material *m, const double *eps, const double *vars, double *s) {
  double dl, nrm[6], st[6];
  plastic_law(m, eps, vars, &dl, nrm, st);
  memcpy(s, st, 6 * sizeof(double));
  for (int i = 0; i < 3; ++i) s[i] += m->k * (eps[0] + eps[1] + eps[2]);
  for (int i = 0; i < 6; ++i) s[i] -= 2 * m->mu * dl * nrm[i];
}

/* src/material.cpp:206-225 */
static double hardening(const orc_material *m, double r) {
  const double Ey = 10.0e4, inf_Ey = 10. * Ey, H0 = 10.0, H1 = 5.0;
  const double r0 = Ey / sqrt(m->E), q0 = r0, q1 = inf_Ey / sqrt(m->E);
  const double r1 = r0 + (q1 - q0) / H0;
  if (r < r0) return 0.0;
  if (r >= r0 && r < r1) return q0 + H0 * (r - r0);
  return q1 + H1 * (r - r1);
}

/* src/material.cpp:228-269 */
static int damage_law(const orc_material *m, const double *eps, double r_old, double D_old, double *r_new,
                      double *D_new, double *s_lin) {
  stress_elastic(m, eps, s_lin);
  double prod = 0.0;
  for (int i = 0; i < 6; ++i) prod += s_lin[i] * eps[i];
  const double r = (prod >= 0) ? sqrt(prod) : 0;
  const double floor_r = (r_old < m->Xt / sqrt(m->E)) ? m->Xt / sqrt(m->E) : r_old;
  if (r <= floor_r) {
    *r_new = floor_r;
    *D_new = D_old;
    return 0;
  }
  const double q = hardening(m, r);
  *r_new = r;
  *D_new = 1. - q / r;
  return 1;
}

/* src/material.cpp:272-287 */
static void stress_damage(const orc_material *m, const double *eps, const double *vars, double *s) {
  double r, D;
  damage_law(m, eps, vars ? vars[0] : 0.0, vars ? vars[1] : 0.0, &r, &D, s);
  for (int i = 0; i < 6; ++i) s[i] *= (1 - D);
}

void orc_mat_stress(const orc_material *m, const double *eps, const double *vars, double *s) {
  if (m->type == ORC_ELASTIC)
    stress_elastic(m, eps, s);
  else if (m->type == ORC_PLASTIC)
    stress_plastic(m, eps, vars, s);
  else
    stress_damage(m, eps, vars, s);
}

/* elastic: src/material.cpp:84-94; plastic/damage: forward difference, src/material.cpp:49-63 */
void orc_mat_ctan(const orc_material *m, const double *eps, const double *vars, double *c) {
  if (m->type == ORC_ELASTIC) {
    memset(c, 0, 36 * sizeof(double));
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) c[i * 6 + j] += m->lambda;
    for (int i = 0; i < 3; ++i) c[i * 6 + i] += 2 * m->mu;
    for (int i = 3; i < 6; ++i) c[i * 6 + i] = m->mu;
    return;
  }
  double s0[6];
  orc_mat_stress(m, eps, vars, s0);
  for (int i = 0; i < 6; ++i) {
    double e1[6], s1[6];
    memcpy(e1, eps, sizeof(e1));
    e1[i] += D_EPS_CTAN;
    orc_mat_stress(m, e1, vars, s1);
    for (int j = 0; j < 6; ++j) c[j * 6 + i] = (s1[j] - s0[j]) / D_EPS_CTAN;
  }
}

/* src/material.cpp:96-99 (elastic), :171-186 (plastic: writes only with an old state), :294-307 (damage) */
int orc_mat_evolute(const orc_material *m, const double *eps, const double *vold, double *vnew) {
  if (m->type == ORC_ELASTIC) return 0;
  if (m->type == ORC_PLASTIC) {
    double dl, nrm[6], st[6];
    const int nl = plastic_law(m, eps, vold, &dl, nrm, st);
    if (vold && vnew) {
      for (int i = 0; i < 6; ++i) vnew[i] = vold[i] + dl * nrm[i];
      vnew[6] = vold[6] + SQRT_2DIV3 * dl + 0;
    }
    return nl;
  }
  double s[6], r, D;
  const int nl = damage_law(m, eps, vold ? vold[0] : 0, vold ? vold[1] : 0, &r, &D, s);
  if (vnew) {
    vnew[0] = r;
    vnew[1] = D;
  }
  return nl;
}

/* ------------------------------------------------------------------------------------------ FE stages */

/* include/util.hpp:58-66 */
void orc_mvp3(const double m[9], const double x[3], double y[3]) {
  for (int r = 0; r < 3; ++r) {
    double acc = 0.0;
    for (int c = 0; c < 3; ++c) acc += m[r * 3 + c] * x[c];
    y[r] = acc;
  }
}

/* src/micro3D.cpp:27-78: six face sweeps, later faces overwrite earlier ones; "max" faces use the literal 1.0 */
void orc_set_displ_bc(const orc_problem *P, const double eps[6], double *u) {
  const double et[9] = {eps[0], 0.5 * eps[3], 0.5 * eps[4], 0.5 * eps[3], eps[1],
                        0.5 * eps[5], 0.5 * eps[4], 0.5 * eps[5], eps[2]};
  const int nx = P->nx, ny = P->ny, nz = P->nz;
  for (int face = 0; face < 6; ++face)
    for (int a = 0; a < (face < 4 ? nx : ny); ++a)
      for (int b = 0; b < (face < 2 ? ny : nz); ++b) {
        int i, j, k;
        double c[3];
        if (face < 2) { /* z = 0 / z = lz */
          i = a, j = b, k = face == 0 ? 0 : nz - 1;
          c[0] = i * P->dx, c[1] = j * P->dy, c[2] = face == 0 ? 0 : 1.0;
        } else if (face < 4) { /* y = 0 / y = ly */
          i = a, k = b, j = face == 2 ? 0 : ny - 1;
          c[0] = i * P->dx, c[1] = face == 2 ? 0 : 1.0, c[2] = k * P->dz;
        } else { /* x = 0 / x = lx */
          j = a, k = b, i = face == 4 ? 0 : nx - 1;
          c[0] = face == 4 ? 0 : 1.0, c[1] = j * P->dy, c[2] = k * P->dz;
        }
        orc_mvp3(et, c, &u[((size_t)k * nx * ny + j * nx + i) * 3]);
      }
}

/* src/common.cpp:45-72 */
static void gp_strain(const orc_problem *P, const double *u, int gp, int ex, int ey, int ez, double eps[6]) {
  int n[8];
  double ue[24];
  orc_elem_nodes(P->nx, P->ny, ex, ey, ez, n);
  for (int a = 0; a < 8; ++a)
    for (int d = 0; d < 3; ++d) ue[a * 3 + d] = u[n[a] * 3 + d];
  for (int v = 0; v < 6; ++v) {
    double acc = 0;
    for (int i = 0; i < 24; ++i) acc += P->bmat[gp][v][i] * ue[i];
    eps[v] = acc;
  }
}

static const double *gp_vars(const double *vars, int e, int gp) { return vars ? vars + (size_t)e * 56 + gp * 7 : NULL; }

/* src/assembly.cpp:28-103 (+ get_elem_rhs :124-138) */
double orc_assembly_rhs(const orc_problem *P, const double *u, const double *vars, double *b) {
  memset(b, 0, P->nndim * sizeof(double));
  for (int ez = 0; ez < P->nez; ++ez)
    for (int ey = 0; ey < P->ney; ++ey)
      for (int ex = 0; ex < P->nex; ++ex) {
        const int e = (ez * P->ney + ey) * P->nex + ex;
        const orc_material *m = &P->mat[P->elem_type[e]];
        double be[24];
        memset(be, 0, sizeof(be));
        for (int gp = 0; gp < 8; ++gp) {
          double eps[6], sig[6];
          gp_strain(P, u, gp, ex, ey, ez, eps);
          orc_mat_stress(m, eps, gp_vars(vars, e, gp), sig);
          for (int i = 0; i < 24; ++i)
            for (int j = 0; j < 6; ++j) be[i] += P->bmat[gp][j][i] * sig[j] * P->wg;
        }
        int n[8];
        orc_elem_nodes(P->nx, P->ny, ex, ey, ez, n);
        for (int a = 0; a < 8; ++a)
          for (int d = 0; d < 3; ++d) b[n[a] * 3 + d] += be[a * 3 + d];
      }
  for (int n = 0; n < P->nn; ++n)
    if (on_boundary(P, n)) b[n * 3] = b[n * 3 + 1] = b[n * 3 + 2] = 0.0;
  for (int i = 0; i < P->nndim; ++i) b[i] = -b[i];
  double acc = 0.0;
  for (int i = 0; i < P->nndim; ++i) acc += b[i] * b[i];
  return sqrt(acc);
}

/* src/assembly.cpp:141-178 */
void orc_elem_mat(const orc_problem *P, const double *u, const double *vars, int ex, int ey, int ez, double *Ke) {
  const int e = (ez * P->ney + ey) * P->nex + ex;
  const orc_material *m = &P->mat[P->elem_type[e]];
  memset(Ke, 0, 576 * sizeof(double));
  for (int gp = 0; gp < 8; ++gp) {
    double eps[6], C[36], cb[6][24];
    gp_strain(P, u, gp, ex, ey, ez, eps);
    orc_mat_ctan(m, eps, gp_vars(vars, e, gp), C);
    for (int i = 0; i < 6; ++i)
      for (int j = 0; j < 24; ++j) {
        double acc = 0.0;
        for (int k = 0; k < 6; ++k) acc += C[i * 6 + k] * P->bmat[gp][k][j];
        cb[i][j] = acc * P->wg;
      }
    for (int r = 0; r < 6; ++r)
      for (int i = 0; i < 24; ++i) {
        const double bri = P->bmat[gp][r][i];
        for (int j = 0; j < 24; ++j) Ke[i * 24 + j] += bri * cb[r][j];
      }
  }
}

/* src/assembly.cpp:106-121 with ell_add_3D (src/ell-common.cpp:166-198) and ell_set_bc_3D (:238-297).
 * Element visiting order: ex outermost, ez innermost. */
void orc_assembly_mat(const orc_problem *P, const double *u, const double *vars, double *vals) {
  memset(vals, 0, (size_t)P->nndim * NNZ * sizeof(double));
  double Ke[576];
  for (int ex = 0; ex < P->nex; ++ex)
    for (int ey = 0; ey < P->ney; ++ey)
      for (int ez = 0; ez < P->nez; ++ez) {
        orc_elem_mat(P, u, vars, ex, ey, ez, Ke);
        int n[8];
        orc_elem_nodes(P->nx, P->ny, ex, ey, ez, n);
        for (int fi = 0; fi < 3; ++fi)
          for (int fj = 0; fj < 3; ++fj)
            for (int i = 0; i < 8; ++i)
              for (int j = 0; j < 8; ++j)
                vals[((size_t)n[i] * 3 + fi) * NNZ + k_cols_row[i][j] * 3 + fj] += Ke[(i * 3 + fi) * 24 + j * 3 + fj];
      }
  for (int n = 0; n < P->nn; ++n)
    if (on_boundary(P, n))
      for (int d = 0; d < 3; ++d) {
        double *row = vals + ((size_t)n * 3 + d) * NNZ;
        memset(row, 0, NNZ * sizeof(double));
        row[13 * 3 + d] = 1;
      }
}

/* src/ell.cpp:35-44 with the column table rebuilt on the fly */
void orc_ell_mvp(int nx, int ny, int nz, const double *vals, const double *x, double *y) {
  const int nn = nx * ny * nz;
  int *cols = (int *)malloc((size_t)nn * 3 * NNZ * sizeof(int));
  orc_ell_cols(nx, ny, nz, cols);
  for (int r = 0; r < nn * 3; ++r) {
    double acc = 0;
    for (int s = 0; s < NNZ; ++s) acc += vals[(size_t)r * NNZ + s] * x[cols[(size_t)r * NNZ + s]];
    y[r] = acc;
  }
  free(cols);
}

static double dot(const double *a, const double *b, int n) {
  double acc = 0.0;
  for (int i = 0; i < n; ++i) acc += a[i] * b[i];
  return acc;
}

/* src/ell.cpp:66-122: Jacobi-preconditioned CG; the convergence test sits at the head of the loop */
int orc_ell_solve_cgpd_hist(int nx, int ny, int nz, const double *vals, const double *b, double *x, double *err,
                            double *hist, int nhist);
int orc_ell_solve_cgpd(int nx, int ny, int nz, const double *vals, const double *b, double *x, double *err) {
  return orc_ell_solve_cgpd_hist(nx, ny, nz, vals, b, x, err, NULL, 0);
}
/* the same solver; hist[i] = the |z| that the loop-head test of iteration i sees (i < nhist) */
int orc_ell_solve_cgpd_hist(int nx, int ny, int nz, const double *vals, const double *b, double *x, double *err,
                            double *hist, int nhist) {
  const int nn = nx * ny * nz, nrow = nn * 3;
  int *cols = (int *)malloc((size_t)nrow * NNZ * sizeof(int));
  double *w = (double *)malloc((size_t)nrow * 5 * sizeof(double));
  double *kk = w, *r = w + nrow, *z = w + 2 * nrow, *p = w + 3 * nrow, *Ap = w + 4 * nrow;
  orc_ell_cols(nx, ny, nz, cols);
#define MVP(in, out)                                                                            \
  for (int r_ = 0; r_ < nrow; ++r_) {                                                           \
    double acc_ = 0;                                                                            \
    for (int s_ = 0; s_ < NNZ; ++s_) acc_ += vals[(size_t)r_ * NNZ + s_] * (in)[cols[(size_t)r_ * NNZ + s_]]; \
    (out)[r_] = acc_;                                                                           \
  }
  for (int n = 0; n < nn; ++n)
    for (int d = 0; d < 3; ++d) kk[n * 3 + d] = 1 / vals[((size_t)n * 3 + d) * NNZ + 13 * 3 + d];
  for (int i = 0; i < nrow; ++i) x[i] = 0.0;
  MVP(x, r);
  for (int i = 0; i < nrow; ++i) r[i] = b[i] - r[i];
  for (int i = 0; i < nrow; ++i) z[i] = kk[i] * r[i];
  for (int i = 0; i < nrow; ++i) p[i] = z[i];
  double rz = dot(r, z, nrow);
  const double pnorm0 = sqrt(dot(z, z, nrow));
  double pnorm = pnorm0;
  int its = 0;
  while (its < CG_MAX_ITS) {
    if (hist && its < nhist) hist[its] = pnorm;
    if (pnorm < CG_ABS_TOL || pnorm < pnorm0 * CG_REL_TOL) break;
    MVP(p, Ap);
    const double alpha = rz / dot(p, Ap, nrow);
    for (int i = 0; i < nrow; ++i) x[i] += alpha * p[i];
    for (int i = 0; i < nrow; ++i) r[i] -= alpha * Ap[i];
    for (int i = 0; i < nrow; ++i) z[i] = kk[i] * r[i];
    pnorm = sqrt(dot(z, z, nrow));
    const double rz_n = dot(r, z, nrow);
    const double beta = rz_n / rz;
    for (int i = 0; i < nrow; ++i) p[i] = z[i] + beta * p[i];
    rz = rz_n;
    its++;
  }
#undef MVP
  *err = rz;
  free(cols);
  free(w);
  return its;
}

/* src/solve.cpp:29-82 (use_A0 = false).  out3 = {its, solver_its, converged} */
void orc_newton(const orc_problem *P, const double eps[6], const double *vars, double *u, int out3[3]) {
  double *b = (double *)calloc(P->nndim, sizeof(double));
  double *du = (double *)calloc(P->nndim, sizeof(double));
  double *vals = (double *)malloc((size_t)P->nndim * NNZ * sizeof(double));
  orc_set_displ_bc(P, eps, u);
  int its = 0, solver_its = 0, converged = 0;
  double norm = orc_assembly_rhs(P, u, vars, b);
  const double norm0 = norm;
  while (its < P->nr_max_its) {
    if (norm < P->nr_max_tol || norm < norm0 * P->nr_rel_tol) {
      converged = 1;
      break;
    }
    orc_assembly_mat(P, u, vars, vals);
    double err;
    solver_its += orc_ell_solve_cgpd(P->nx, P->ny, P->nz, vals, b, du, &err);
    for (int i = 0; i < P->nndim; ++i) u[i] += du[i];
    norm = orc_assembly_rhs(P, u, vars, b);
    its++;
  }
  out3[0] = its;
  out3[1] = solver_its;
  out3[2] = converged;
  free(b);
  free(du);
  free(vals);
}

/* src/average.cpp:58-82 (vol_tot = 1) */
void orc_ave_stress(const orc_problem *P, const double *u, const double *vars, double sig[6]) {
  memset(sig, 0, 6 * sizeof(double));
  for (int ez = 0; ez < P->nez; ++ez)
    for (int ey = 0; ey < P->ney; ++ey)
      for (int ex = 0; ex < P->nex; ++ex) {
        const int e = (ez * P->ney + ey) * P->nex + ex;
        const orc_material *m = &P->mat[P->elem_type[e]];
        double part[6] = {0, 0, 0, 0, 0, 0};
        for (int gp = 0; gp < 8; ++gp) {
          double eps[6], s[6];
          gp_strain(P, u, gp, ex, ey, ez, eps);
          orc_mat_stress(m, eps, gp_vars(vars, e, gp), s);
          for (int v = 0; v < 6; ++v) part[v] += s[v] * P->wg;
        }
        for (int v = 0; v < 6; ++v) sig[v] += part[v];
      }
  for (int v = 0; v < 6; ++v) sig[v] /= 1.0;
}

/* src/update.cpp:33-56 */
int orc_vars_new(const orc_problem *P, const double *u, const double *vold, double *vnew) {
  int nl = 0;
  for (int ez = 0; ez < P->nez; ++ez)
    for (int ey = 0; ey < P->ney; ++ey)
      for (int ex = 0; ex < P->nex; ++ex) {
        const int e = (ez * P->ney + ey) * P->nex + ex;
        const orc_material *m = &P->mat[P->elem_type[e]];
        for (int gp = 0; gp < 8; ++gp) {
          double eps[6];
          gp_strain(P, u, gp, ex, ey, ez, eps);
          nl |= orc_mat_evolute(m, eps, gp_vars(vold, e, gp), vnew + (size_t)e * 56 + gp * 7);
        }
      }
  return nl;
}

/* ------------------------------------------------------------------------------------------ one Gauss point */

/* State of one macro Gauss point (include/gp.hpp:33-125) and homogenize_fe_one_way
 * (src/homogenize.cpp:112-186) without sub-iterations; lin_stress selects sigma = ctan_lin * eps. */
typedef struct orc_gp {
  double strain_old[6], strain[6], stress[6], ctan[36];
  int allocated, cost, converged;
  double *u_n, *u_k, *vars_n, *vars_k;
} orc_gp;

orc_gp *orc_gp_new(const orc_problem *P) {
  orc_gp *g = (orc_gp *)calloc(1, sizeof(orc_gp));
  g->u_n = (double *)calloc(P->nndim, sizeof(double));
  g->u_k = (double *)calloc(P->nndim, sizeof(double));
  g->converged = 1;
  return g;
}
void orc_gp_free(orc_gp *g) {
  free(g->u_n);
  free(g->u_k);
  free(g->vars_n);
  free(g->vars_k);
  free(g);
}
void orc_gp_set_strain(orc_gp *g, const double *e) { memcpy(g->strain, e, sizeof(g->strain)); }
void orc_gp_get_stress(const orc_gp *g, double *s) { memcpy(s, g->stress, sizeof(g->stress)); }
void orc_gp_set_ctan(orc_gp *g, const double *c) { memcpy(g->ctan, c, sizeof(g->ctan)); }
int orc_gp_cost(const orc_gp *g) { return g->cost; }
int orc_gp_converged(const orc_gp *g) { return g->converged; }
int orc_gp_allocated(const orc_gp *g) { return g->allocated; }

void orc_gp_homogenize(const orc_problem *P, orc_gp *g, int lin_stress) {
  const size_t nvars = (size_t)P->nelem * 56;
  double *u = (double *)malloc(P->nndim * sizeof(double));
  double *scratch = (double *)calloc(nvars ? nvars : 1, sizeof(double));
  double *vnew = g->allocated ? g->vars_k : scratch;
  int res[3];
  memcpy(u, g->u_n, P->nndim * sizeof(double));
  orc_newton(P, g->strain, g->vars_n, u, res);
  memcpy(g->u_k, u, P->nndim * sizeof(double));
  g->cost = res[1];
  g->converged = res[2];
  if (lin_stress) {
    for (int i = 0; i < 6; ++i) {
      g->stress[i] = 0.0;
      for (int j = 0; j < 6; ++j) g->stress[i] += g->ctan[i * 6 + j] * g->strain[j];
    }
  } else {
    orc_ave_stress(P, g->u_k, g->vars_n, g->stress);
  }
  if (orc_vars_new(P, g->u_k, g->vars_n, vnew) && !g->allocated) {
    g->vars_n = (double *)calloc(nvars, sizeof(double));
    g->vars_k = (double *)calloc(nvars, sizeof(double));
    g->allocated = 1;
    memcpy(g->vars_k, vnew, nvars * sizeof(double));
  }
  free(u);
  free(scratch);
}

/* include/gp.hpp:95-105 */
void orc_gp_update_vars(orc_gp *g) {
  double *t = g->vars_n;
  g->vars_n = g->vars_k;
  g->vars_k = t;
  t = g->u_n;
  g->u_n = g->u_k;
  g->u_k = t;
  memcpy(g->strain_old, g->strain, sizeof(g->strain));
}
