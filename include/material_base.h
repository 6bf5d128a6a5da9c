/*
 * material_base.h -- plain-C description of one micro-scale material.
 *
 * Drop-in replacement for the reference header include/material_base.h:27-50.
 * The struct is mirrored by the Fortran `bind(C)` type of src/material.f95:30-35, so the
 * field ORDER (E, nu, Ka, Sy, k, mu, lambda, Xt, type) is part of the ABI and must not change.
 */
#ifndef MATERIAL_BASE_H
#define MATERIAL_BASE_H

#ifdef __cplusplus
#include <cstdio>
extern "C" {
#else
#include <stdbool.h>
#include <stdio.h>
#endif

/* material law selector stored in material_base::type */
enum { MATERIAL_ELASTIC = 0, MATERIAL_PLASTIC, MATERIAL_DAMAGE };

/* strain step of the forward-difference tangent (src/material.cpp:49-63) */
#define D_EPS_CTAN 1.0e-8
/* sqrt(2/3) truncated to nine digits exactly as the reference has it -- parity depends on it */
#define SQRT_2DIV3 0.816496581

struct material_base {
  double E, nu, Ka, Sy; /* Young, Poisson, hardening modulus, yield stress */
  double k, mu, lambda; /* bulk, shear, Lame -- derived by material_set    */
  double Xt;            /* damage threshold stress                          */
  int type;             /* MATERIAL_ELASTIC | MATERIAL_PLASTIC | MATERIAL_DAMAGE */
};

/* Fills every field, deriving k, mu and lambda from (E, nu).  Replaces src/material.c:26-38. */
void material_set(struct material_base *self, const int type, const double E, const double nu, const double Ka,
                  const double Sy, const double Xt);

/* Human-readable dump to stdout.  Replaces src/material.c:40-45. */
void material_print(const struct material_base *self);

#ifdef __cplusplus
}
#endif
#endif /* MATERIAL_BASE_H */
