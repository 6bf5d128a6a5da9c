// micropp.hpp -- the C++ API of micropp-b200.
//
// `micropp<3>` keeps the public AND protected signatures of the reference class
// (include/micropp.hpp:54-218): macro-scale FEM codes use the public part
// (set_strain / homogenize / get_stress / get_ctan / update_vars ...), and the reference's own tests
// subclass it to reach the FE stages (test/test_cg.cpp:38-82), so both compile unchanged.
//
// What differs is where the work happens.  The object owns a device context (include/mgpu.h) that
// keeps a resident wave of RVEs in B200 HBM; homogenize() crosses the host/device boundary once
// (strains in; stress / tangent / flags out).  The protected FE stages take HOST pointers in the
// reference's layouts and stage them through the same CUDA kernels.  There is no CPU path:
// constructing a micropp<3> without a CUDA device aborts.
#pragma once

#include <cassert>
#include <cmath>
#include <cstring>
#include <ctime>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <sstream>
#include <string>
#include <vector>

#include "ell.hpp"
#include "gp.hpp"
#include "instrument.hpp"
#include "material.hpp"
#include "params.hpp"
#include "types.hpp"
#include "util.hpp"

using namespace std;  // the reference header leaks this and code written against it relies on it

struct mpp_engine;  // wave scheduler + Newton / CG drivers (micropp_b200/csrc/micropp_host.cpp)

template <int tdim>
class micropp {
 protected:
  static constexpr int dim = tdim;
  static constexpr int npe = mypow(2, dim);
  static constexpr int nvoi = dim * (dim + 1) / 2;
  double bmat[npe][nvoi][npe * dim];  // B matrices at the 8 Gauss points (src/micro3D.cpp:81-120)

  const int ngp, nx, ny, nz, nn, nndim;
  const int nex, ney, nez, nelem;
  const double lx, ly, lz;
  const double dx, dy, dz;
  const double vol_tot;
  const double wg, ivol, evol;

  const int micro_type, nvars;
  const int nsubiterations;
  const bool subiterations;
  const int mpi_rank;

  gp_t<tdim> *gp_list;

  static const int num_geo_params = 4;
  double geo_params[num_geo_params];

  material_t *material_list[MAX_MATERIALS];
  double ctan_lin_fe[nvoi * nvoi];

  int *elem_type;
  double *elem_stress;
  double *elem_strain;

  const double xg[8][3] = {{-CONSTXG, -CONSTXG, -CONSTXG}, {+CONSTXG, -CONSTXG, -CONSTXG},
                           {+CONSTXG, +CONSTXG, -CONSTXG}, {-CONSTXG, +CONSTXG, -CONSTXG},
                           {-CONSTXG, -CONSTXG, +CONSTXG}, {+CONSTXG, -CONSTXG, +CONSTXG},
                           {+CONSTXG, +CONSTXG, +CONSTXG}, {-CONSTXG, +CONSTXG, +CONSTXG}};

  const int nr_max_its;
  const double nr_max_tol;
  const double nr_rel_tol;
  const bool calc_ctan_lin_flag;
  const bool lin_stress;

  bool use_A0;
  int its_with_A0;
  ell_matrix *A0;  // unused on the host: the linear Jacobian lives in the device context

  double Vm;  // matrix volume fraction
  double Vf;  // fibre volume fraction

  const bool write_log_flag;
  int log_id = 0;
  ofstream ofstream_log;

  int gpu_id = 0;  // mpi_rank % visible devices (src/micropp.cpp:77-81)

  mpp_engine *engine = nullptr;
  double last_homogenize_ms = 0.0;

  void homogenize_linear(gp_t<tdim> *gp_ptr);
  // The two FE homogenizations of the reference take one GP; here they receive the whole batch that
  // shares the coupling mode (see homogenize()).
  void homogenize_fe_one_way(gp_t<tdim> *gp_ptr);
  void homogenize_fe_full(gp_t<tdim> *gp_ptr);
  void homogenize_fe_batch(const std::vector<int> &gp_ids);

  void calc_ctan_lin_fe_models();
  void calc_ctan_lin_mix_rule_Chamis(double ctan[nvoi * nvoi]);

  material_t *get_material(const int e) const;

  void get_stress(int gp, const double eps[nvoi], const double *vars_old, double stress_gp[nvoi], int ex, int ey,
                  int ez = 0) const;

  int get_elem_type(int ex, int ey, int ez = 0) const;

  void get_elem_rhs(const double *u, const double *vars_old, double be[npe * dim], int ex, int ey, int ez = 0) const;

  void calc_ave_stress(const double *u, double stress_ave[nvoi], const double *vars_old = nullptr) const;
  void calc_ave_strain(const double *u, double strain_ave[nvoi]) const;
  void calc_fields(double *u, double *vars_old);
  void calc_bmat(int gp, double bmat[nvoi][npe * dim]) const;
  void calc_volume_fractions();

  bool calc_vars_new(const double *u, const double *vars_old, double *vars_new) const;

  newton_t newton_raphson(ell_matrix *A, double *b, double *u, double *du, const double strain[nvoi],
                          const double *vars_old = nullptr);

  void get_elem_mat(const double *u, const double *vars_old, double Ae[npe * dim * npe * dim], int ex, int ey,
                    int ez = 0) const;

  void set_displ_bc(const double strain[nvoi], double *u);
  double assembly_rhs(const double *u, const double *vars_old, double *b);
  void assembly_mat(ell_matrix *A, const double *u, const double *vars_old);

  void write_vtu(double *u, double *vars_old, const char *filename);
  void write_log();

 public:
  micropp() = delete;
  micropp(const micropp_params_t &params);
  ~micropp();

  // macro-scale coupling (src/homogenize.cpp:31-99, :285-287)
  void set_strain(const int gp_id, const double *strain);
  void get_stress(const int gp_id, double *stress) const;
  void get_ctan(const int gp_id, double *ctan) const;
  void homogenize();
  void homogenize_linear();
  void update_vars();

  // bookkeeping (src/micropp.cpp:217-253)
  int is_non_linear(const int gp_id) const;
  int get_non_linear_gps(void) const;
  int get_cost(int gp_id) const;
  bool has_converged(int gp_id) const;
  bool has_subiterated(int gp_id) const;

  // files (src/output.cpp)
  void output(int gp_id, const char *filename);
  void output2(const int gp_id, const int elem_global, const int time_step);
  void write_restart(const int restart_id) const;
  void read_restart(const int restart_id) const;
  void print_info() const;

  friend struct mpp_access;  // C-ABI extension layer (include/micropp_b200_ext.h)
};

// Every member of micropp<3> is an explicit specialization defined in micropp_b200/csrc/*.cpp.
#define MPP_SPEC template <>
MPP_SPEC micropp<3>::micropp(const micropp_params_t &params);
MPP_SPEC micropp<3>::~micropp();
MPP_SPEC void micropp<3>::homogenize_linear(gp_t<3> *gp_ptr);
MPP_SPEC void micropp<3>::homogenize_fe_one_way(gp_t<3> *gp_ptr);
MPP_SPEC void micropp<3>::homogenize_fe_full(gp_t<3> *gp_ptr);
MPP_SPEC void micropp<3>::homogenize_fe_batch(const std::vector<int> &gp_ids);
MPP_SPEC void micropp<3>::calc_ctan_lin_fe_models();
MPP_SPEC void micropp<3>::calc_ctan_lin_mix_rule_Chamis(double ctan[36]);
MPP_SPEC material_t *micropp<3>::get_material(const int e) const;
MPP_SPEC void micropp<3>::get_stress(int gp, const double eps[6], const double *vars_old, double stress_gp[6], int ex,
                                     int ey, int ez) const;
MPP_SPEC int micropp<3>::get_elem_type(int ex, int ey, int ez) const;
MPP_SPEC void micropp<3>::calc_ave_stress(const double *u, double stress_ave[6], const double *vars_old) const;
MPP_SPEC void micropp<3>::calc_bmat(int gp, double bmat[6][24]) const;
MPP_SPEC void micropp<3>::calc_volume_fractions();
MPP_SPEC bool micropp<3>::calc_vars_new(const double *u, const double *vars_old, double *vars_new) const;
MPP_SPEC newton_t micropp<3>::newton_raphson(ell_matrix *A, double *b, double *u, double *du, const double strain[6],
                                             const double *vars_old);
MPP_SPEC void micropp<3>::set_displ_bc(const double strain[6], double *u);
MPP_SPEC double micropp<3>::assembly_rhs(const double *u, const double *vars_old, double *b);
MPP_SPEC void micropp<3>::assembly_mat(ell_matrix *A, const double *u, const double *vars_old);
MPP_SPEC void micropp<3>::write_log();
MPP_SPEC void micropp<3>::set_strain(const int gp_id, const double *strain);
MPP_SPEC void micropp<3>::get_stress(const int gp_id, double *stress) const;
MPP_SPEC void micropp<3>::get_ctan(const int gp_id, double *ctan) const;
MPP_SPEC void micropp<3>::homogenize();
MPP_SPEC void micropp<3>::homogenize_linear();
MPP_SPEC void micropp<3>::update_vars();
MPP_SPEC int micropp<3>::is_non_linear(const int gp_id) const;
MPP_SPEC int micropp<3>::get_non_linear_gps(void) const;
MPP_SPEC int micropp<3>::get_cost(int gp_id) const;
MPP_SPEC bool micropp<3>::has_converged(int gp_id) const;
MPP_SPEC bool micropp<3>::has_subiterated(int gp_id) const;
MPP_SPEC void micropp<3>::calc_fields(double *u, double *vars_old);
MPP_SPEC void micropp<3>::write_vtu(double *u, double *vars_old, const char *filename);
MPP_SPEC void micropp<3>::output(int gp_id, const char *filename);
MPP_SPEC void micropp<3>::output2(const int gp_id, const int elem_global, const int time_step);
MPP_SPEC void micropp<3>::write_restart(const int restart_id) const;
MPP_SPEC void micropp<3>::read_restart(const int restart_id) const;
MPP_SPEC void micropp<3>::print_info() const;
#undef MPP_SPEC
