// ell.hpp -- host-visible ELLPACK matrix of the structured 27-point / 9-point stencil.
//
// API-compatible with the reference's include/ell.hpp:33-94.  Inside the B200 hot path the matrix
// never exists in this form: values are stored as 243 node-major planes and the column table is
// implicit (see include/mgpu.h).  This struct and the free functions below are the reference's
// *host* API, kept because its tests call them directly (test/test_ell_2.cpp, test/test_cg.cpp);
// the solver entry points stage the host matrix through the GPU kernels.
#pragma once

#include <cassert>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <string>

#define CG_ABS_TOL 1.0e-50
#define CG_MAX_ITS 1000
#define CG_REL_TOL 1.0e-5

#define nod_index(i, j, k) ((k) * nx * ny + (j) * nx + (i))
#define nod_index3D(i, j, k) ((k) * nx * ny + (j) * nx + (i))
#define nod_index2D(i, j) ((j) * nx + (i))

typedef struct {
  int n[3];    // nodes per direction
  int nn;      // nodes
  int dim;     // 2 or 3
  int nfield;  // unknowns per node
  int shift;   // stencil slot of the diagonal block (4 in 2-D, 13 in 3-D)
  int nrow;
  int ncol;
  int nnz;  // stored entries per row (9*nfield or 27*nfield)
  int *cols = NULL;
  double *vals = NULL;

  int max_its;
  double min_err;
  double rel_err;
  double *k, *r, *z, *p, *Ap;  // CG work vectors (host copies; the GPU solver has its own)
} ell_matrix;

// Allocates and fills the column table (src/ell-common.cpp:34-139).
void ell_init(ell_matrix *m, const int nfield, const int dim, const int ns[3], const double min_err = CG_ABS_TOL,
              const double rel_err = CG_REL_TOL, const int max_its = CG_MAX_ITS);

// y = A x on the GPU (3-D, nfield = 3); replaces src/ell.cpp:35-44.
void ell_mvp(const ell_matrix *m, const double *x, double *y);
// Jacobi-preconditioned CG on the GPU; returns iterations, *err = final r.z; replaces src/ell.cpp:66-122.
int ell_solve_cgpd(const ell_matrix *m, const double *b, double *x, double *err_);
void ell_add_2D(ell_matrix *m, int ex, int ey, const double *Ae);
void ell_add_3D(ell_matrix *m, int ex, int ey, int ez, const double *Ae);
void ell_set_zero_mat(ell_matrix *m);
void ell_set_bc_2D(ell_matrix *m);
void ell_set_bc_3D(ell_matrix *m);
void ell_free(ell_matrix *m);

double get_norm(const double *vector, const int n);
double get_dot(const double *v1, const double *v2, const int n);
double ell_get_norm(const ell_matrix *m);

int ell_write(std::string filename, const ell_matrix *A);
int ell_read(std::string filename, ell_matrix *A);
void print_ell(const ell_matrix *A);
