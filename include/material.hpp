// material.hpp -- host-side view of the three micro-scale material laws.
//
// API-compatible with the reference's include/material.hpp:36-157 (make_material factory, virtual
// get_stress / get_ctan / evolute / init_vars / print, the three concrete classes and their
// constructors deriving k, mu, lambda).  On the GPU the laws are NOT virtual: kernels switch on the
// POD type tag (micropp_b200/csrc/fe_math.cuh).  The host classes below call those same inline
// functions compiled for the host, and exist for API completeness (print_info, the mixture rule, and
// reference tests such as test/test_material.cpp); homogenize() never uses them.
#pragma once

#include <cmath>
#include <cstdio>
#include <cstring>
#include <iostream>

#include "common.hpp"
#include "material_base.h"

struct material_t : public material_base {
  // Builds the law selected by material.type (0 elastic, 1 plastic, 2 damage); nullptr otherwise.
  static material_t *make_material(const struct material_base material);

  virtual ~material_t() {}
  virtual void init_vars(double *vars_old) const = 0;
  // sigma(eps; history) -- history may be nullptr (virgin material)
  virtual void get_stress(const double *eps, double *stress, const double *history_params) const = 0;
  // 6x6 row-major tangent; forward differences for plastic / damage
  virtual void get_ctan(const double *eps, double *ctan, const double *history_params) const = 0;
  // writes the new internal variables; returns true when the point left the linear range
  virtual bool evolute(const double *eps, const double *vars_old, double *vars_new) const = 0;
  virtual void print() const = 0;

 protected:
  void apply_perturbation(const double *eps, double *ctan, const double *vars_old) const;
  void derive_moduli(double E_, double nu_) {
    E = E_;
    nu = nu_;
    k = E_ / (3. * (1. - 2. * nu_));
    mu = E_ / (2. * (1. + nu_));
    lambda = nu_ * E_ / ((1. + nu_) * (1. - 2. * nu_));
  }
};

class material_elastic : public material_t {
 public:
  material_elastic(double E_, double nu_) {
    derive_moduli(E_, nu_);
    Ka = Sy = Xt = -1.0;
    type = MATERIAL_ELASTIC;
  }
  void init_vars(double *vars_old) const override;
  void get_stress(const double *eps, double *stress, const double *history_params) const override;
  void get_ctan(const double *eps, double *ctan, const double *history_params) const override;
  bool evolute(const double *eps, const double *vars_old, double *vars_new) const override;
  void print() const override;
};

class material_plastic : public material_t {
 public:
  material_plastic(double E_, double nu_, double Ka_, double Sy_) {
    derive_moduli(E_, nu_);
    Ka = Ka_;
    Sy = Sy_;
    Xt = -1.0;
    type = MATERIAL_PLASTIC;
  }
  void init_vars(double *vars_old) const override;
  void get_stress(const double *eps, double *stress, const double *history_params) const override;
  void get_ctan(const double *eps, double *ctan, const double *vars_old) const override;
  bool evolute(const double *eps, const double *vars_old, double *vars_new) const override;
  void print() const override;
};

class material_damage : public material_t {
 public:
  material_damage(double E_, double nu_, double Xt_) {
    derive_moduli(E_, nu_);
    Ka = Sy = -1.0;
    Xt = Xt_;
    type = MATERIAL_DAMAGE;
  }
  void init_vars(double *vars_old) const override;
  void get_stress(const double *eps, double *stress, const double *vars_old) const override;
  void get_ctan(const double *eps, double *ctan, const double *vars_old) const override;
  bool evolute(const double *eps, const double *vars_old, double *vars_new) const override;
  void print() const override;
};
