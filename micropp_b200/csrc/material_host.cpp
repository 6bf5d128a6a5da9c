// material_host.cpp -- host-side material objects (include/material.hpp) and the C helpers of
// include/material_base.h.  The arithmetic is the SAME inline code the CUDA kernels use
// (fe_math.cuh compiled for the host), which restates src/material.cpp:27-307 of the reference.
// Not on the hot path: homogenize() evaluates materials only inside kernels.
#include <iostream>

#include "fe_math.cuh"
#include "material.hpp"

using std::cout;
using std::endl;

namespace {
inline mpp_material pod(const material_base &b) {
  mpp_material m;
  m.E = b.E;
  m.nu = b.nu;
  m.Ka = b.Ka;
  m.Sy = b.Sy;
  m.k = b.k;
  m.mu = b.mu;
  m.lambda = b.lambda;
  m.Xt = b.Xt;
  m.type = b.type;
  return m;
}
}  // namespace

extern "C" {

void material_set(struct material_base *self, const int type, const double E, const double nu, const double Ka,
                  const double Sy, const double Xt) {
  self->type = type;
  self->E = E;
  self->nu = nu;
  self->Ka = Ka;
  self->Sy = Sy;
  self->Xt = Xt;
  self->k = E / (3. * (1. - 2. * nu));
  self->mu = E / (2. * (1. + nu));
  self->lambda = nu * E / ((1. + nu) * (1. - 2. * nu));
}

void material_print(const struct material_base *self) {
  printf("E = %e\nnu = %e\nKa = %e\nSy = %e\nXt = %e\nk = %e\nmu = %e\nlambda = %e\ntype = %1d\n", self->E, self->nu,
         self->Ka, self->Sy, self->Xt, self->k, self->mu, self->lambda, self->type);
}
}

material_t *material_t::make_material(const struct material_base material) {
  switch (material.type) {
    case MATERIAL_ELASTIC: return new material_elastic(material.E, material.nu);
    case MATERIAL_PLASTIC: return new material_plastic(material.E, material.nu, material.Ka, material.Sy);
    case MATERIAL_DAMAGE: return new material_damage(material.E, material.nu, material.Xt);
    default: break;
  }
  return nullptr;
}

void material_t::apply_perturbation(const double *eps, double *ctan, const double *vars_old) const {
  mat_ctan(pod(*this), eps, vars_old, ctan);
}

// ---- elastic ----
void material_elastic::init_vars(double *) const {}
void material_elastic::get_stress(const double *eps, double *stress, const double *) const {
  elastic_stress(pod(*this), eps, stress);
}
void material_elastic::get_ctan(const double *, double *ctan, const double *) const { elastic_ctan(pod(*this), ctan); }
bool material_elastic::evolute(const double *, const double *, double *) const { return false; }
void material_elastic::print() const {
  cout << "Type : Elastic" << endl;
  cout << std::scientific << "E = " << E << " nu = " << nu << endl;
}

// ---- plastic ----
void material_plastic::init_vars(double *) const {}
void material_plastic::get_stress(const double *eps, double *stress, const double *h) const {
  plastic_stress(pod(*this), eps, h, stress);
}
void material_plastic::get_ctan(const double *eps, double *ctan, const double *vars_old) const {
  apply_perturbation(eps, ctan, vars_old);
}
bool material_plastic::evolute(const double *eps, const double *vars_old, double *vars_new) const {
  return plastic_evolute(pod(*this), eps, vars_old, vars_new);
}
void material_plastic::print() const {
  cout << "Type : Plastic" << endl;
  cout << "E = " << E << " nu = " << nu << " Ka = " << Ka << " Sy = " << Sy << endl;
}

// ---- damage ----
void material_damage::init_vars(double *vars_old) const {
  if (vars_old != nullptr) {  // never called by the solver
    vars_old[0] = 10.0e5 / sqrt(E);
    vars_old[1] = 0.0;
  }
}
void material_damage::get_stress(const double *eps, double *stress, const double *vars_old) const {
  damage_stress(pod(*this), eps, vars_old, stress);
}
void material_damage::get_ctan(const double *eps, double *ctan, const double *vars_old) const {
  apply_perturbation(eps, ctan, vars_old);
}
bool material_damage::evolute(const double *eps, const double *vars_old, double *vars_new) const {
  return damage_evolute(pod(*this), eps, vars_old, vars_new);
}
void material_damage::print() const {
  cout << "Type : Damage" << endl;
  cout << "E = " << E << " nu = " << nu << " Xt = " << Xt << endl;
}
