// micropp_c.cpp -- the C ABI (include/micropp_c.h + the additive include/micropp_b200_ext.h).
// One-line forwards to micropp<3>, mirroring the reference wrapper src/micropp_c.cpp:33-141.
#include <cstring>

#include "micropp.hpp"
#include "micropp_b200_ext.h"

namespace {
inline micropp<3> *obj(const micropp3 *s) { return static_cast<micropp<3> *>(s->ptr); }
}

extern "C" {

void micropp3_new(struct micropp3 *self, int ngp, const int size[3], const int type, const double *geo_params,
                  const struct material_base *materials, const int *coupling, const int nsubiterations,
                  const int mpi_rank) {
  micropp_params_t p;
  p.ngp = ngp;
  memcpy(p.size, size, 3 * sizeof(int));
  p.type = type;
  memcpy(p.geo_params, geo_params, 4 * sizeof(double));
  for (int i = 0; i < MAX_MATERIALS; ++i) p.materials[i] = materials[i];
  // the reference copies ngp ints from `coupling` unconditionally (src/micropp_c.cpp:44-45)
  std::vector<int> cpl(coupling, coupling + ngp);
  p.coupling = cpl.data();
  p.subiterations = true;
  p.nsubiterations = nsubiterations;
  p.mpi_rank = mpi_rank;
  p.use_A0 = false;
  p.its_with_A0 = 1;
  p.write_log = false;
  self->ptr = new micropp<3>(p);
}

void micropp3_new_ext(struct micropp3 *self, const struct micropp3_params *q) {
  micropp_params_t p;
  p.ngp = q->ngp;
  memcpy(p.size, q->size, sizeof(p.size));
  p.type = q->type;
  memcpy(p.geo_params, q->geo_params, sizeof(p.geo_params));
  for (int i = 0; i < MAX_MATERIALS; ++i)
    material_set(&p.materials[i], q->mat_type[i], q->mat_E[i], q->mat_nu[i], q->mat_Ka[i], q->mat_Sy[i],
                 q->mat_Xt[i]);
  p.coupling = const_cast<int *>(q->coupling);
  p.subiterations = q->subiterations != 0;
  p.nsubiterations = q->nsubiterations;
  p.mpi_rank = q->mpi_rank;
  p.nr_max_its = q->nr_max_its;
  p.nr_max_tol = q->nr_max_tol;
  p.nr_rel_tol = q->nr_rel_tol;
  p.calc_ctan_lin = q->calc_ctan_lin != 0;
  p.use_A0 = q->use_A0 != 0;
  p.its_with_A0 = q->its_with_A0;
  p.lin_stress = q->lin_stress != 0;
  p.write_log = q->write_log != 0;
  self->ptr = new micropp<3>(p);
}

void micropp3_free(micropp3 *self) {
  delete obj(self);
  self->ptr = nullptr;
}
void micropp3_set_strain(micropp3 *self, const int gp_id, const double *strain) { obj(self)->set_strain(gp_id, strain); }
void micropp3_get_stress(const micropp3 *self, const int gp_id, double *stress) { obj(self)->get_stress(gp_id, stress); }
void micropp3_get_ctan(const micropp3 *self, int gp_id, double *ctan) { obj(self)->get_ctan(gp_id, ctan); }
void micropp3_homogenize(micropp3 *self) { obj(self)->homogenize(); }
void micropp3_homogenize_linear(micropp3 *self) { obj(self)->homogenize_linear(); }
int micropp3_get_cost(const micropp3 *self, int gp_id) { return obj(self)->get_cost(gp_id); }
bool micropp3_has_converged(const micropp3 *self, const int gp_id) { return obj(self)->has_converged(gp_id); }
bool micropp3_has_subiterated(const micropp3 *self, const int gp_id) { return obj(self)->has_subiterated(gp_id); }
void micropp3_update_vars(micropp3 *self) { obj(self)->update_vars(); }
void micropp3_output(micropp3 *self, const int gp_id, const char *filename) { obj(self)->output(gp_id, filename); }
void micropp3_output2(micropp3 *self, const int gp_id, const int elem_global, const int time_step) {
  obj(self)->output2(gp_id, elem_global, time_step);
}
void micropp3_print_info(micropp3 *self) { obj(self)->print_info(); }
bool micropp3_is_non_linear(const micropp3 *self, const int gp_id) { return obj(self)->is_non_linear(gp_id); }
int micropp3_get_non_linear_gps(const micropp3 *self) { return obj(self)->get_non_linear_gps(); }
void micropp3_write_restart(const micropp3 *self, const int restart_id) { obj(self)->write_restart(restart_id); }
void micropp3_read_restart(const micropp3 *self, const int restart_id) { obj(self)->read_restart(restart_id); }
}
