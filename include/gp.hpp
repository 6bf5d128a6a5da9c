// gp.hpp -- host-side record of one macro Gauss point.
//
// API-compatible with the reference's gp_t (include/gp.hpp:33-125): same public fields.  The big
// per-GP arrays (u_n, u_k, vars_n, vars_k) live in B200 HBM inside the device context; the four
// pointers below are kept for source compatibility and stay nullptr on the host.  Swapping them
// (update_vars) and the restart I/O are done by micropp<3> against the device copies.
#pragma once

#include <cassert>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>

template <int dim>
class gp_t {
  static constexpr int nvoi = dim * (dim + 1) / 2;

 public:
  double strain_old[nvoi] = {0.0};
  double strain[nvoi] = {0.0};
  double stress[nvoi] = {0.0};
  double ctan[nvoi * nvoi] = {0.0};

  bool allocated = false;  // internal variables exist (the GP went non-linear at least once)

  double *vars_n = nullptr;  // device-resident; see header comment
  double *vars_k = nullptr;
  double *u_n = nullptr;
  double *u_k = nullptr;
  int nvars = 0;
  int nndim = 0;

  long int cost = 0;         // CG iterations spent by the last homogenize()
  bool converged = true;     // last Newton solve converged
  bool subiterated = false;  // sub-stepping was needed
  int coupling = 0;

  int fe_index = -1;  // index of this GP's state inside the device context (-1: no FE state)

  // strain history part of update_vars (include/gp.hpp:95-105); the array swaps happen on the device
  void update_vars() { memcpy(strain_old, strain, nvoi * sizeof(double)); }
};
