/* c_api_drive.c -- a plain-C macro-code stand-in that drives the drop-in boundary exactly as include/micropp_c.h
 * declares it (the reference's own C driver, test/test3d_6.c, reads a third material it never defines and passes a
 * NULL coupling array, which the reference dereferences -- it crashes against the reference itself -- so this file
 * restates its load path with valid arguments).  Compiled with gcc (C linkage, <stdbool.h>), linked against
 * libmicropp_b200.so by `make -C oracle ctests`, run by tests/test_gpu_ctests.py.
 *
 * Checks: two Gauss points with identical strains give identical results (test/test3d_4.cpp:109-115), the stress is
 * finite and grows with the load, the plastic matrix goes non-linear, ctan is returned.  Exit code 0 = pass. */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include "micropp_c.h"

#define D_EPS 4.0e-4

int main(int argc, char *argv[]) {
  const int n = argc > 1 ? atoi(argv[1]) : 5;
  const int dir = argc > 2 ? atoi(argv[2]) : 0;
  const int steps = argc > 3 ? atoi(argv[3]) : 10;
  int size[3] = {n, n, n};
  struct material_base mats[3];
  material_set(&mats[0], 1, 1.0e7, 0.25, 1.0e4, 1.0e4, 1.0); /* plastic matrix (test/test3d_6.c:50) */
  material_set(&mats[1], 0, 1.0e7, 0.25, 1.0e4, 1.0e7, 0.0); /* elastic inclusion */
  material_set(&mats[2], 0, 1.0e7, 0.25, 0.0, 0.0, 0.0);
  material_print(&mats[0]);
  struct micropp3 micro;
  const int ngp = 2;
  int coupling[2] = {1, 1}; /* FE_ONE_WAY */
  double params[4] = {.2, 0., 0., 0.}; /* sphere radius */
  micropp3_new(&micro, ngp, size, 1 /* MIC_SPHERE */, params, mats, coupling, 1, 0);
  micropp3_print_info(&micro);
  double eps[6] = {0.}, sig[2][6], ctan[36], prev = 0.0;
  int fail = 0;
  for (int t = 0; t < steps; ++t) {
    eps[dir] += D_EPS;
    for (int g = 0; g < ngp; ++g) micropp3_set_strain(&micro, g, eps);
    micropp3_homogenize(&micro);
    for (int g = 0; g < ngp; ++g) micropp3_get_stress(&micro, g, sig[g]);
    micropp3_get_ctan(&micro, 0, ctan);
    for (int i = 0; i < 6; ++i) {
      if (!isfinite(sig[0][i]) || sig[0][i] != sig[1][i]) fail = 1;
    }
    if (!(sig[0][dir] > prev)) fail = 1;
    prev = sig[0][dir];
    if (micropp3_get_cost(&micro, 0) != micropp3_get_cost(&micro, 1)) fail = 1;
    printf("step %d cost %d nl %d conv %d sig[%d] %e ctan00 %e\n", t, micropp3_get_cost(&micro, 0),
           (int)micropp3_is_non_linear(&micro, 0), (int)micropp3_has_converged(&micro, 0), dir, sig[0][dir], ctan[0]);
    micropp3_update_vars(&micro);
  }
  if (micropp3_get_non_linear_gps(&micro) != ngp) fail = 1; /* 4e-3 strain on Sy = 1e4, E = 1e7: the matrix yields */
  if (!(ctan[0] > 0.0)) fail = 1;
  micropp3_free(&micro);
  printf(fail ? "FAIL\n" : "OK\n");
  return fail;
}
