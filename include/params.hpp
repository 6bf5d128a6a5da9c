// params.hpp -- compile-time constants of the micro-scale solver.
// Same names and values as the reference's include/params.hpp:26-41 (tests and macro codes use them).
#pragma once

#define MAX_DIM 3
#define NUM_VAR_GP 7     // internal variables per element Gauss point: eps_p[6] + alpha
#define MAX_MATERIALS 3  // materials addressable by elem_type

#define FILTER_REL_TOL 1.0e-5

#define D_EPS_CTAN_AVE 1.0e-8  // macro-strain step of the homogenized tangent (src/homogenize.cpp:256-275)

#define CONSTXG 0.577350269189626  // 1/sqrt(3) as truncated by the reference

#define NR_MAX_TOL 1.0e-10
#define NR_MAX_ITS 4
#define NR_REL_TOL 1.0e-3  // relative to the first residual of the Newton solve

// element / internal-variable numbering of the reference layout (used by restart files and output)
#define glo_elem(ex, ey, ez) ((ez) * (nx - 1) * (ny - 1) + (ey) * (nx - 1) + (ex))
#define intvar_ix(e, gp, var) ((e) * npe * NUM_VAR_GP + (gp) * NUM_VAR_GP + (var))
