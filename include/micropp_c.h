/*
 * micropp_c.h -- the C ABI of micropp-b200: THE drop-in boundary for C and Fortran macro codes.
 *
 * Every entry point has the name, argument order and meaning of the reference wrapper
 * (include/micropp_c.h:33-70, implemented in src/micropp_c.cpp:33-141), which is also what the
 * Fortran interface module binds (src/micropp.f95:37-180).  Citations below are reference file:line.
 *
 * Conventions: strain / stress are 6 doubles in Voigt order
 * [11,22,33,12,13,23] with engineering shear strains; ctan is 36 doubles row-major,
 * ctan[v*6+i] = d sigma_v / d eps_i.  The caller owns every buffer.  There are no error codes:
 * failures surface as micropp3_has_converged() == false or NaNs, as in the reference.
 */
#ifndef MICROPP3_WRAPPER_H
#define MICROPP3_WRAPPER_H

#include "material_base.h"

#ifdef __cplusplus
extern "C" {
#endif

/* Opaque handle; layout mirrored by Fortran (src/micropp.f95:33-35). */
struct micropp3 {
  void *ptr;
};

/* src/micropp_c.cpp:37-60.  Forces subiterations=true, use_A0=false, write_log=false; everything it
 * does not set keeps the defaults of micropp_params_t (lin_stress=true, calc_ctan_lin=true, nr_max_its=4).
 * `coupling` must point to ngp ints (the reference dereferences it unconditionally). */
void micropp3_new(struct micropp3 *self, int ngp, const int size[3], const int micro_type,
                  const double *micro_params, const struct material_base *materials, const int *coupling,
                  const int nsubiterations, const int mpi_rank);

void micropp3_free(struct micropp3 *self);                                              /* :62-65  */
void micropp3_set_strain(struct micropp3 *self, const int gp_id, const double *strain); /* :67-70  */
void micropp3_get_stress(const struct micropp3 *self, const int gp_id, double *stress); /* :72-75  */
void micropp3_get_ctan(const struct micropp3 *self, const int gp_id, double *ctan);     /* :77-80  */
void micropp3_homogenize(struct micropp3 *self);                                        /* :82-85  */
void micropp3_homogenize_linear(struct micropp3 *self);                                 /* :87-90  */
void micropp3_update_vars(struct micropp3 *self);                                       /* :107-110 */
bool micropp3_is_non_linear(const struct micropp3 *self, const int gp_id);              /* :127-130 */
int micropp3_get_cost(const struct micropp3 *self, int gp_id);                          /* :92-95  */
bool micropp3_has_converged(const struct micropp3 *self, int gp_id);                    /* :97-100 */
bool micropp3_has_subiterated(const struct micropp3 *self, int gp_id);                  /* :102-105 */
void micropp3_output(struct micropp3 *self, const int gp_id, const char *filename);     /* :112-115 */
/* defined by the reference (:117-120) although its header forgets to declare it */
void micropp3_output2(struct micropp3 *self, const int gp_id, const int elem_global, const int time_step);
void micropp3_print_info(struct micropp3 *self);                             /* :122-125 */
int micropp3_get_non_linear_gps(const struct micropp3 *self);                /* :132-135 */
void micropp3_write_restart(const struct micropp3 *self, const int restart_id); /* :137-140 */
void micropp3_read_restart(const struct micropp3 *self, const int restart_id);  /* :142-145 */

#ifdef __cplusplus
}
#endif
#endif /* MICROPP3_WRAPPER_H */
