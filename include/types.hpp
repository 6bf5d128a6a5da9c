// types.hpp -- run-time configuration and result PODs of micropp<3>.
// API-compatible with the reference's include/types.hpp:29-117: same type names, field names,
// defaults and enumerator values, so that a macro-scale code or a reference test compiles unchanged.
#pragma once

#include <iostream>
#include <map>
#include <string>

#include "ell.hpp"
#include "material_base.h"
#include "params.hpp"

// What one Newton-Raphson solve reports (src/solve.cpp:29-82).
typedef struct {
  int its = 0;         // Newton iterations spent
  int solver_its = 0;  // CG iterations summed over them
  bool converged = false;

  void print() {
    std::cout << "newton.its        : " << its << std::endl;
    std::cout << "newton.solver_its : " << solver_its << std::endl;
    std::cout << "newton.converged  : " << converged << std::endl;
  }
} newton_t;

typedef struct {
  int ngp = 1;                                  // macro Gauss points handled by this object
  int size[3];                                  // nodes per edge of the RVE grid
  int type = 0;                                 // micro-structure (MIC_* below)
  double geo_params[4] = {0.1, 0.1, 0.1, 0.1};  // radii / widths, meaning depends on `type`
  struct material_base materials[4];
  int *coupling = nullptr;  // per GP: FE_LINEAR / FE_ONE_WAY / FE_FULL / MIX_RULE_CHAMIS; NULL = all FE_ONE_WAY
  bool subiterations = false;
  int nsubiterations = 10;
  int mpi_rank = 0;  // selects the GPU (rank % #devices) and names log / restart files
  int nr_max_its = NR_MAX_ITS;
  double nr_max_tol = NR_MAX_TOL;
  double nr_rel_tol = NR_REL_TOL;
  // NB: as in the reference the three cg_* fields are accepted but the solver always runs with
  // CG_MAX_ITS / CG_ABS_TOL / CG_REL_TOL (src/homogenize.cpp:115).
  int cg_max_its = CG_MAX_ITS;
  double cg_abs_tol = CG_ABS_TOL;
  double cg_rel_tol = CG_REL_TOL;
  bool calc_ctan_lin = true;
  bool use_A0 = false;
  int its_with_A0 = 1;
  bool lin_stress = true;
  bool write_log = false;

  void print() {
    using std::cout;
    using std::endl;
    cout << "ngp  : " << ngp << endl;
    cout << "size : " << size[0] << endl;
    cout << "type  : " << type << endl;
    cout << "geo_params : " << geo_params[0] << endl;
    cout << "subiterations : " << subiterations << endl;
    cout << "nsubiterations : " << nsubiterations << endl;
    cout << "mpi_rank : " << mpi_rank << endl;
    cout << "nr_max_its : " << nr_max_its << endl;
    cout << "nr_max_tol : " << nr_max_tol << endl;
    cout << "nr_rel_tol : " << nr_rel_tol << endl;
    cout << "calc_ctan_lin : " << calc_ctan_lin << endl;
    cout << "use_A0 : " << use_A0 << endl;
    cout << "its_with_A0 : " << its_with_A0 << endl;
    cout << "lin_stress : " << lin_stress << endl;
    cout << "write_log : " << write_log << endl;
  }
} micropp_params_t;

// Micro-structure catalogue (geometry in micropp<3>::get_elem_type).
enum {
  MIC_HOMOGENEOUS,
  MIC_SPHERE,
  MIC_LAYER_Y,
  MIC_CILI_FIB_X,
  MIC_CILI_FIB_Z,
  MIC_CILI_FIB_XZ,
  MIC_QUAD_FIB_XYZ,
  MIC_QUAD_FIB_XZ,
  MIC_QUAD_FIB_XZ_BROKEN_X,
  MIC3D_SPHERES,
  MIC3D_8,
  MIC3D_FIBS_20_ORDER,
  MIC3D_FIBS_20_DISORDER
};

static std::map<int, std::string> micro_names = {{MIC_HOMOGENEOUS, "MIC_HOMOGENEOUS"},
                                                 {MIC_SPHERE, "MIC_SPHERE"},
                                                 {MIC_LAYER_Y, "MIC_LAYER_Y"},
                                                 {MIC_CILI_FIB_X, "MIC_CILI_FIB_X"},
                                                 {MIC_CILI_FIB_Z, "MIC_CILI_FIB_Z"},
                                                 {MIC_CILI_FIB_XZ, "MIC_CILI_FIB_XZ"},
                                                 {MIC_QUAD_FIB_XYZ, "MIC_QUAD_FIB_XYZ"},
                                                 {MIC_QUAD_FIB_XZ, "MIC_QUAD_FIB_XZ"},
                                                 {MIC_QUAD_FIB_XZ_BROKEN_X, "MIC_QUAD_FIB_XZ_BROKEN_X"},
                                                 {MIC3D_SPHERES, "MIC3D_SPHERES"},
                                                 {MIC3D_8, "MIC3D_8"},
                                                 {MIC3D_FIBS_20_ORDER, "MIC3D_FIBS_20_ORDER"},
                                                 {MIC3D_FIBS_20_DISORDER, "MIC3D_FIBS_20_DISORDER"}};

// How a macro Gauss point is coupled to its RVE.
enum { FE_LINEAR, FE_ONE_WAY, FE_FULL, MIX_RULE_CHAMIS };

// Per-translation-unit tally of GPs by coupling (only print_info reads it), as in the reference.
static std::map<int, int> gp_counter = {{FE_LINEAR, 0}, {FE_ONE_WAY, 0}, {FE_FULL, 0}, {MIX_RULE_CHAMIS, 0}};
