// micropp_host.cpp -- C++ host side of micropp-b200: the micropp<3> class of include/micropp.hpp.
//
// Everything numerical is done by the CUDA kernels behind include/mgpu.h; this file holds
//   * the construction-time set-up the reference does on the host (mesh constants, B matrices,
//     element classification, elastic element matrices)            src/micropp.cpp:30-183
//   * mpp_engine: the batched Newton-Raphson and DPCG drivers       src/solve.cpp:29-82, src/ell.cpp:66-122
//   * the per-Gauss-point state machine of homogenize()             src/homogenize.cpp:69-287
//   * host-pointer versions of the protected FE stages (staged through the same kernels).
// The reference loops over Gauss points with OpenMP and solves one RVE at a time per thread; here a
// whole wave of RVEs advances in lock-step, one kernel launch per algorithmic step for all of them.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "mgpu.h"
#include "micropp.hpp"
#include "mpp_engine.hpp"

// ================================================================================================
// construction
// ================================================================================================
int mpp_elem_type(int micro_type, const double *geo_params, double dx, double dy, double dz, int ex, int ey, int ez);

namespace {
void bmat_at(const double xg3[3], double dx, double dy, double dz, double b[6][24]);
}

template <>
void micropp<3>::calc_bmat(int gp, double b[nvoi][npe * dim]) const {
  bmat_at(xg[gp], dx, dy, dz, b);
}

namespace {
void bmat_at(const double xg3[3], double dx, double dy, double dz, double b[6][24]) {
  const int nvoi = 6, dim = 3;
  // Trilinear hex8 shape-function derivatives at Gauss point gp, scaled to the dx*dy*dz cell
  // (reference src/micro3D.cpp:81-120).  Node a sits at corner (cx,cy,cz) in {0,1}^3 with signs
  // s = 2c-1; dN_a/dx = sx (1 + sy eta)(1 + sz zeta)/8 * 2/dx.  Sign flips are exact, so the
  // products below round exactly like the reference's literal table.
  static const int corner[8][3] = {{0, 0, 0}, {1, 0, 0}, {1, 1, 0}, {0, 1, 0},
                                   {0, 0, 1}, {1, 0, 1}, {1, 1, 1}, {0, 1, 1}};
  const double xi = xg3[0], eta = xg3[1], zeta = xg3[2];
  for (int a = 0; a < 8; ++a) {
    const double sx = corner[a][0] ? 1.0 : -1.0, sy = corner[a][1] ? 1.0 : -1.0, sz = corner[a][2] ? 1.0 : -1.0;
    const double fx = 1 + sx * xi, fy = 1 + sy * eta, fz = 1 + sz * zeta;
    const double gx = sx * fy * fz / 8. * 2. / dx;
    const double gy = sy * fx * fz / 8. * 2. / dy;
    const double gz = sz * fx * fy / 8. * 2. / dz;
    for (int v = 0; v < nvoi; ++v)
      for (int d = 0; d < dim; ++d) b[v][a * dim + d] = 0;
    b[0][a * dim + 0] = gx;
    b[1][a * dim + 1] = gy;
    b[2][a * dim + 2] = gz;
    b[3][a * dim + 0] = gy;
    b[3][a * dim + 1] = gx;
    b[4][a * dim + 0] = gz;
    b[4][a * dim + 2] = gx;
    b[5][a * dim + 1] = gz;
    b[5][a * dim + 2] = gy;
  }
}

// 24x24 element matrix of an elastic material on the uniform grid: sum over Gauss points of
// B^T (C B wg), accumulated in the reference's loop order (src/assembly.cpp:141-178).
void elastic_element_matrix(const double bmat[8][6][24], const material_base &m, double wg, double *Ke) {
  double C[6][6];
  memset(C, 0, sizeof(C));
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) C[i][j] += m.lambda;
  for (int i = 0; i < 3; ++i) C[i][i] += 2 * m.mu;
  for (int i = 3; i < 6; ++i) C[i][i] = m.mu;

  for (int q = 0; q < 576; ++q) Ke[q] = 0.0;
  for (int gp = 0; gp < 8; ++gp) {
    double cb[6][24];
    for (int i = 0; i < 6; ++i)
      for (int j = 0; j < 24; ++j) {
        double acc = 0.0;
        for (int k = 0; k < 6; ++k) acc += C[i][k] * bmat[gp][k][j];
        cb[i][j] = acc * wg;
      }
    for (int mrow = 0; mrow < 6; ++mrow)
      for (int i = 0; i < 24; ++i) {
        const double bmi = bmat[gp][mrow][i];
        for (int j = 0; j < 24; ++j) Ke[i * 24 + j] += bmi * cb[mrow][j];
      }
  }
}

}  // namespace

template <>
micropp<3>::micropp(const micropp_params_t &params)
    : ngp(params.ngp),
      nx(params.size[0]),
      ny(params.size[1]),
      nz(params.size[2]),
      nn(nx * ny * nz),
      nndim(nn * dim),
      nex(nx - 1),
      ney(ny - 1),
      nez(nz - 1),
      nelem(nex * ney * nez),
      lx(1.0),
      ly(1.0),
      lz(1.0),
      dx(lx / nex),
      dy(ly / ney),
      dz(lz / nez),
      vol_tot(lx * ly * lz),
      wg((dx * dy * dz) / npe),
      ivol(1.0 / (wg * npe)),
      evol(dx * dy * dz),
      micro_type(params.type),
      nvars(nelem * npe * NUM_VAR_GP),
      nsubiterations(params.nsubiterations),
      subiterations(params.subiterations),
      mpi_rank(params.mpi_rank),
      nr_max_its(params.nr_max_its),
      nr_max_tol(params.nr_max_tol),
      nr_rel_tol(params.nr_rel_tol),
      calc_ctan_lin_flag(params.calc_ctan_lin),
      lin_stress(params.lin_stress),
      use_A0(params.use_A0),
      its_with_A0(params.its_with_A0),
      A0(nullptr),
      write_log_flag(params.write_log) {
  for (int gp = 0; gp < npe; ++gp) calc_bmat(gp, bmat[gp]);

  // Gauss points: coupling mode and which of them own FE state on the device (src/micropp.cpp:87-103)
  gp_list = new gp_t<3>[ngp]();
  int n_fe = 0;
  for (int gp = 0; gp < ngp; ++gp) {
    gp_t<3> &g = gp_list[gp];
    g.coupling = (params.coupling != nullptr) ? params.coupling[gp] : FE_ONE_WAY;
    gp_counter[g.coupling]++;
    g.nndim = nndim;
    g.nvars = nvars;
    if (g.coupling == FE_ONE_WAY || g.coupling == FE_FULL) g.fe_index = n_fe++;
  }

  elem_type = (int *)calloc(nelem > 0 ? nelem : 1, sizeof(int));
  elem_stress = (double *)calloc((nelem > 0 ? nelem : 1) * nvoi, sizeof(double));
  elem_strain = (double *)calloc((nelem > 0 ? nelem : 1) * nvoi, sizeof(double));

  for (int i = 0; i < num_geo_params; ++i) geo_params[i] = params.geo_params[i];
  for (int i = 0; i < MAX_MATERIALS; ++i) material_list[i] = material_t::make_material(params.materials[i]);

  for (int ez = 0; ez < nez; ++ez)
    for (int ey = 0; ey < ney; ++ey)
      for (int ex = 0; ex < nex; ++ex) elem_type[glo_elem(ex, ey, ez)] = get_elem_type(ex, ey, ez);

  calc_volume_fractions();

  // ---- device context ----
  mgpu_config cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.nx = nx;
  cfg.ny = ny;
  cfg.nz = nz;
  const int ndev = mgpu_device_count();
  gpu_id = ndev > 0 ? mpi_rank % ndev : 0;
  cfg.device = gpu_id;
  cfg.ngp = n_fe;
  cfg.elem_type = elem_type;
  for (int gp = 0; gp < 8; ++gp)
    for (int a = 0; a < 8; ++a) {
      cfg.dsh[gp][a * 3 + 0] = bmat[gp][0][a * 3 + 0];
      cfg.dsh[gp][a * 3 + 1] = bmat[gp][1][a * 3 + 1];
      cfg.dsh[gp][a * 3 + 2] = bmat[gp][2][a * 3 + 2];
    }
  cfg.wg = wg;
  cfg.dx = dx;
  cfg.dy = dy;
  cfg.dz = dz;
  std::vector<double> ke(3 * 576, 0.0);
  for (int i = 0; i < MAX_MATERIALS; ++i) {
    const material_base &m = params.materials[i];
    const double vals[8] = {m.E, m.nu, m.Ka, m.Sy, m.k, m.mu, m.lambda, m.Xt};
    memcpy(cfg.mat[i], vals, sizeof(vals));
    cfg.mat_type[i] = m.type;
    // element matrix of the material's ELASTIC law: what an elastic material always has, and what a damage / plastic
    // element has below its threshold (src/material.cpp:111-164, 206-287: the linear branch is lambda, mu elasticity)
    elastic_element_matrix(bmat, m, wg, &ke[i * 576]);
  }
  cfg.ke_elastic = ke.data();
  cfg.nr_max_its = nr_max_its;
  cfg.nr_max_tol = nr_max_tol;
  cfg.nr_rel_tol = nr_rel_tol;
  // the reference ignores params.cg_* and always solves with the macros (src/homogenize.cpp:115)
  cfg.cg_max_its = CG_MAX_ITS;
  cfg.cg_abs_tol = CG_ABS_TOL;
  cfg.cg_rel_tol = CG_REL_TOL;
  cfg.wave_cap = 0;
  // all-elastic RVEs run DPCG on the implicit operator (MICROPP_IMPLICIT=0: one assembled ELL matrix per slot)
  cfg.implicit_elastic = 1;
  if (const char *env = getenv("MICROPP_IMPLICIT")) cfg.implicit_elastic = atoi(env) != 0;

  engine = new mpp_engine();
  engine->ctx = mgpu_create(&cfg);
  engine->W = mgpu_wave_size(engine->ctx);
  engine->implicit = mgpu_implicit(engine->ctx) != 0;
  engine->hybrid = mgpu_hybrid_available(engine->ctx) != 0;
  engine->use_A0 = use_A0;
  engine->its_with_A0 = its_with_A0;
  if (const char *env = getenv("MICROPP_CG_CHUNK")) engine->cg_chunk = std::max(1, atoi(env));
  if (const char *env = getenv("MICROPP_CG_GROUP")) engine->cg_group = std::max(0, atoi(env));
  if (const char *env = getenv("MICROPP_GRAPHS")) engine->use_graphs = atoi(env) != 0;

  if (use_A0) {
    // linear Jacobian at u = 0 without history (src/micropp.cpp:128-143): one shared device matrix
    const int s0 = 0, g0 = -1;
    mgpu_bind_slots(engine->ctx, 1, &s0, &g0, nullptr);
    mgpu_set_list(engine->ctx, mpp_engine::L_SUB, 1, &s0);
    mgpu_zero_u(engine->ctx, mpp_engine::L_SUB, 1);
    mgpu_asm_mat(engine->ctx, mpp_engine::L_SUB, 1, 1);
    engine->A0_ready = true;
  }

  memset(ctan_lin_fe, 0, sizeof(ctan_lin_fe));
  if (calc_ctan_lin_flag) {
    const int num_fe_points = gp_counter[FE_LINEAR] + gp_counter[FE_ONE_WAY] + gp_counter[FE_FULL];
    if (num_fe_points > 0) calc_ctan_lin_fe_models();
  }

  for (int gp = 0; gp < ngp; ++gp) {
    gp_t<3> &g = gp_list[gp];
    if (g.coupling == FE_LINEAR || g.coupling == FE_ONE_WAY || g.coupling == FE_FULL) {
      memcpy(g.ctan, ctan_lin_fe, sizeof(ctan_lin_fe));
    } else if (g.coupling == MIX_RULE_CHAMIS) {
      double c[nvoi * nvoi];
      calc_ctan_lin_mix_rule_Chamis(c);
      memcpy(g.ctan, c, sizeof(c));
    }
  }

  if (write_log_flag) {
    std::stringstream name;
    name << "micropp-profiling-" << mpi_rank << ".log";
    ofstream_log.open(name.str(), ios::out);
    ofstream_log << "#<gp_id>  <non-linear>  <cost>  <converged>" << endl;
  }
}

template <>
micropp<3>::~micropp() {
  cout << "Calling micropp<" << dim << "> destructor" << endl;  // the reference prints this (src/micropp.cpp:189)
  if (engine) {
    mgpu_destroy(engine->ctx);
    delete engine;
  }
  free(elem_stress);
  free(elem_strain);
  free(elem_type);
  for (int i = 0; i < MAX_MATERIALS; ++i) delete material_list[i];
  delete[] gp_list;
}

// The six unit-strain solves that define the linear homogenized tangent (src/micropp.cpp:256-284),
// run as one batch of six RVEs.
template <>
void micropp<3>::calc_ctan_lin_fe_models() {
  mgpu_ctx *ctx = engine->ctx;
  const int W = engine->W;
  for (int first = 0; first < nvoi; first += W) {
    const int n = std::min(W, nvoi - first);
    std::vector<int> slots(n), none(n, -1);
    std::vector<double> eps(6 * n, 0.0);
    for (int i = 0; i < n; ++i) {
      slots[i] = i;
      eps[i * 6 + first + i] += D_EPS_CTAN_AVE;
    }
    mgpu_bind_slots(ctx, n, slots.data(), none.data(), nullptr);
    mgpu_set_list(ctx, mpp_engine::L_OUTER, n, slots.data());
    mgpu_set_slot_strain(ctx, n, slots.data(), eps.data());
    mgpu_zero_u(ctx, mpp_engine::L_OUTER, n);
    std::vector<newton_t> res;
    engine->newton_batch(mpp_engine::L_OUTER, n, slots.data(), res);
    mgpu_ave_stress(ctx, mpp_engine::L_OUTER, n);
    std::vector<double> sig(6 * n);
    mgpu_fetch_stress(ctx, n, slots.data(), sig.data());
    for (int i = 0; i < n; ++i)
      for (int v = 0; v < nvoi; ++v) ctan_lin_fe[v * nvoi + first + i] = sig[i * 6 + v] / D_EPS_CTAN_AVE;
  }
}

// Chamis mixture rule for a two-phase unidirectional composite (src/micropp.cpp:287-331).
template <>
void micropp<3>::calc_ctan_lin_mix_rule_Chamis(double ctan[nvoi * nvoi]) {
  const double Em = material_list[0]->E, nu_m = material_list[0]->nu;
  const double Ef = material_list[1]->E, nu_f = material_list[1]->nu;
  const double Gm = Em / (2 * (1 + nu_m));
  const double Gf = Ef / (2 * (1 + nu_f));

  const double E11 = Vf * Ef + Vm * Em;
  const double E22 = Em / (1 - sqrt(Vf) * (1 - Em / Ef));
  const double nu12 = Vf * nu_f + Vm * nu_m;
  const double G12 = Gm / (1 - sqrt(Vf) * (1 - Gm / Gf));
  const double nu23 = nu12;

  const double S[3][3] = {
      {1 / E11, -nu12 / E11, -nu12 / E11}, {-nu12 / E11, 1 / E22, -nu23 / E22}, {-nu12 / E11, -nu23 / E22, 1 / E22}};
  double Si[3][3];
  invert_3x3(S, Si);

  memset(ctan, 0, nvoi * nvoi * sizeof(double));
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) ctan[i * nvoi + j] = Si[i][j];
  for (int i = 3; i < 6; ++i) ctan[i * nvoi + i] = G12;
}

template <>
material_t *micropp<3>::get_material(const int e) const {
  return material_list[elem_type[e]];
}

template <>
void micropp<3>::calc_volume_fractions() {
  Vm = 0.0;
  Vf = 0.0;
  for (int e = 0; e < nelem; ++e) {
    if (elem_type[e] == 0)
      Vm += evol;
    else if (elem_type[e] >= 1)
      Vf += evol;
  }
  Vm /= vol_tot;
  Vf /= vol_tot;
}

// ================================================================================================
// macro-scale coupling
// ================================================================================================
template <>
void micropp<3>::set_strain(const int gp_id, const double *strain) {
  assert(gp_id >= 0 && gp_id < ngp);
  memcpy(gp_list[gp_id].strain, strain, nvoi * sizeof(double));
}

template <>
void micropp<3>::get_stress(const int gp_id, double *stress) const {
  assert(gp_id >= 0 && gp_id < ngp);
  memcpy(stress, gp_list[gp_id].stress, nvoi * sizeof(double));
}

template <>
void micropp<3>::get_ctan(const int gp_id, double *ctan) const {
  assert(gp_id >= 0 && gp_id < ngp);
  memcpy(ctan, gp_list[gp_id].ctan, nvoi * nvoi * sizeof(double));
}

// stress = ctan * strain (src/homogenize.cpp:102-109)
template <>
void micropp<3>::homogenize_linear(gp_t<3> *g) {
  for (int i = 0; i < nvoi; ++i) {
    double acc = 0.0;
    for (int j = 0; j < nvoi; ++j) acc += g->ctan[i * nvoi + j] * g->strain[j];
    g->stress[i] = acc;
  }
}

template <>
void micropp<3>::homogenize_linear() {
  for (int igp = 0; igp < ngp; ++igp) homogenize_linear(&gp_list[igp]);
}

// FE_ONE_WAY and FE_FULL Gauss points, wave by wave (src/homogenize.cpp:112-282).
template <>
void micropp<3>::homogenize_fe_batch(const std::vector<int> &ids) {
  mgpu_ctx *ctx = engine->ctx;
  const int W = engine->W;
  enum { L0 = mpp_engine::L_OUTER, LS = mpp_engine::L_SUB };

  for (size_t off = 0; off < ids.size(); off += W) {
    const int n = (int)std::min<size_t>(W, ids.size() - off);
    std::vector<int> slots(n), fe(n), had_vars(n);
    std::vector<double> eps(6 * n);
    for (int i = 0; i < n; ++i) {
      gp_t<3> &g = gp_list[ids[off + i]];
      slots[i] = i;
      fe[i] = g.fe_index;
      had_vars[i] = g.allocated ? 1 : 0;
      memcpy(&eps[i * 6], g.strain, 6 * sizeof(double));
      g.cost = 0;
      g.subiterated = false;
    }
    mgpu_bind_slots(ctx, n, slots.data(), fe.data(), had_vars.data());
    mgpu_set_list(ctx, L0, n, slots.data());
    mgpu_set_slot_strain(ctx, n, slots.data(), eps.data());

    // first Newton-Raphson from u_n
    mgpu_load_u(ctx, L0, n, 0);
    std::vector<newton_t> last;
    engine->newton_batch(L0, n, slots.data(), last);
    mgpu_store_u(ctx, L0, n, 1);
    for (int i = 0; i < n; ++i) {
      gp_t<3> &g = gp_list[ids[off + i]];
      g.cost += last[i].solver_its;
      g.converged = last[i].converged;
    }

    // sub-stepping of the strain increment for the GPs that failed (src/homogenize.cpp:138-157)
    if (subiterations) {
      std::vector<int> sub, sub_i;
      for (int i = 0; i < n; ++i)
        if (!last[i].converged) {
          sub.push_back(slots[i]);
          sub_i.push_back(i);
        }
      const int ns = (int)sub.size();
      if (ns > 0) {
        std::vector<double> eps_sub(6 * ns), deps(6 * ns);
        for (int q = 0; q < ns; ++q) {
          gp_t<3> &g = gp_list[ids[off + sub_i[q]]];
          g.subiterated = true;
          for (int j = 0; j < 6; ++j) {
            eps_sub[q * 6 + j] = g.strain_old[j];
            deps[q * 6 + j] = (g.strain[j] - g.strain_old[j]) / nsubiterations;
          }
        }
        mgpu_set_list(ctx, LS, ns, sub.data());
        mgpu_load_u(ctx, LS, ns, 0);
        std::vector<newton_t> res;
        for (int its = 0; its < nsubiterations; ++its) {
          for (int q = 0; q < 6 * ns; ++q) eps_sub[q] += deps[q];
          mgpu_set_slot_strain(ctx, ns, sub.data(), eps_sub.data());
          engine->newton_batch(LS, ns, sub.data(), res);
          for (int q = 0; q < ns; ++q) gp_list[ids[off + sub_i[q]]].cost += res[q].solver_its;
        }
        if (nsubiterations > 0) {
          for (int q = 0; q < ns; ++q) {
            last[sub_i[q]] = res[q];
            gp_list[ids[off + sub_i[q]]].converged = res[q].converged;
          }
        }
        mgpu_store_u(ctx, LS, ns, 1);
        // restore the target strain of those slots (the FE_FULL perturbations start from it)
        mgpu_set_slot_strain(ctx, n, slots.data(), eps.data());
      }
    }

    // homogenized stress (src/homogenize.cpp:159-169); the slot's u equals u_k at this point
    if (lin_stress) {
      for (int i = 0; i < n; ++i) homogenize_linear(&gp_list[ids[off + i]]);
    } else {
      mgpu_ave_stress(ctx, L0, n);
      std::vector<double> sig(6 * n);
      mgpu_fetch_stress(ctx, n, slots.data(), sig.data());
      for (int i = 0; i < n; ++i) memcpy(gp_list[ids[off + i]].stress, &sig[i * 6], 6 * sizeof(double));
    }

    // internal variables (src/homogenize.cpp:171-179): first find who is non-linear, allocate, then write
    if (mgpu_nvar(ctx) > 0) {
      mgpu_clear_nl_flags(ctx, L0, n);
      mgpu_vars_new(ctx, L0, n, 0);
      std::vector<mgpu_slot_state> st(n);
      mgpu_fetch_state(ctx, n, slots.data(), st.data());
      std::vector<int> wr;
      for (int i = 0; i < n; ++i) {
        gp_t<3> &g = gp_list[ids[off + i]];
        if (st[i].nl_flag && !g.allocated) {
          mgpu_gp_alloc_vars(ctx, g.fe_index);  // gp_t::allocate(): zero-filled vars_n / vars_k
          g.allocated = true;
          wr.push_back(i);
        } else if (had_vars[i]) {
          wr.push_back(i);
        }
      }
      if (!wr.empty()) {
        // had_vars keeps the *old* meaning: a GP that was just allocated evolved from "no history"
        std::vector<int> wslots(wr.size()), wfe(wr.size()), wold(wr.size());
        for (size_t q = 0; q < wr.size(); ++q) {
          wslots[q] = slots[wr[q]];
          wfe[q] = fe[wr[q]];
          wold[q] = had_vars[wr[q]];
        }
        mgpu_bind_slots(ctx, (int)wr.size(), wslots.data(), wfe.data(), wold.data());
        mgpu_set_list(ctx, LS, (int)wr.size(), wslots.data());
        mgpu_vars_new(ctx, LS, (int)wr.size(), 1);
      }
    }

    // FE_FULL: homogenized tangent by six perturbed solves chained on the same u (src/homogenize.cpp:252-276)
    std::vector<int> full, full_i;
    for (int i = 0; i < n; ++i) {
      const gp_t<3> &g = gp_list[ids[off + i]];
      if (g.coupling == FE_FULL && g.allocated) {
        full.push_back(slots[i]);
        full_i.push_back(i);
      }
    }
    const int nf = (int)full.size();
    if (nf > 0) {
      std::vector<int> ffe(nf), ones(nf, 1);
      for (int q = 0; q < nf; ++q) ffe[q] = fe[full_i[q]];
      mgpu_bind_slots(ctx, nf, full.data(), ffe.data(), ones.data());
      mgpu_set_list(ctx, LS, nf, full.data());
      std::vector<double> eps1(6 * nf), sig1(6 * nf);
      std::vector<newton_t> res;
      for (int i = 0; i < nvoi; ++i) {
        for (int q = 0; q < nf; ++q) {
          const gp_t<3> &g = gp_list[ids[off + full_i[q]]];
          memcpy(&eps1[q * 6], g.strain, 6 * sizeof(double));
          eps1[q * 6 + i] += D_EPS_CTAN_AVE;
        }
        mgpu_set_slot_strain(ctx, nf, full.data(), eps1.data());
        engine->newton_batch(LS, nf, full.data(), res);
        mgpu_ave_stress(ctx, LS, nf);
        mgpu_fetch_stress(ctx, nf, full.data(), sig1.data());
        for (int q = 0; q < nf; ++q) {
          gp_t<3> &g = gp_list[ids[off + full_i[q]]];
          // the reference adds the solver_its of the LAST ASSIGNED newton_t, not of this solve
          // (src/homogenize.cpp:267-269) -- reproduced, it is visible through get_cost()
          g.cost += last[full_i[q]].solver_its;
          for (int v = 0; v < nvoi; ++v) g.ctan[v * nvoi + i] = (sig1[q * 6 + v] - g.stress[v]) / D_EPS_CTAN_AVE;
        }
      }
    }
  }
}

template <>
void micropp<3>::homogenize_fe_one_way(gp_t<3> *g) {
  homogenize_fe_batch(std::vector<int>{(int)(g - gp_list)});
}
template <>
void micropp<3>::homogenize_fe_full(gp_t<3> *g) {
  homogenize_fe_batch(std::vector<int>{(int)(g - gp_list)});
}

template <>
void micropp<3>::homogenize() {
  std::vector<int> fe_ids;
  for (int igp = 0; igp < ngp; ++igp) {
    gp_t<3> *g = &gp_list[igp];
    if (g->coupling == FE_LINEAR || g->coupling == MIX_RULE_CHAMIS)
      homogenize_linear(g);
    else if (g->coupling == FE_ONE_WAY || g->coupling == FE_FULL)
      fe_ids.push_back(igp);
  }
  if (!fe_ids.empty()) {
    mgpu_timer_start(engine->ctx);
    homogenize_fe_batch(fe_ids);
    last_homogenize_ms = mgpu_timer_stop(engine->ctx);
  }
  if (write_log_flag) write_log();
}

template <>
void micropp<3>::update_vars() {
  for (int igp = 0; igp < ngp; ++igp) {
    gp_t<3> &g = gp_list[igp];
    if (g.fe_index >= 0) mgpu_gp_swap(engine->ctx, g.fe_index);
    g.update_vars();
  }
}

// ================================================================================================
// bookkeeping
// ================================================================================================
template <>
int micropp<3>::is_non_linear(const int gp_id) const {
  assert(gp_id >= 0 && gp_id < ngp);
  return (int)gp_list[gp_id].allocated;
}
template <>
int micropp<3>::get_cost(int gp_id) const {
  assert(gp_id >= 0 && gp_id < ngp);
  return gp_list[gp_id].cost;
}
template <>
bool micropp<3>::has_converged(int gp_id) const {
  assert(gp_id >= 0 && gp_id < ngp);
  return gp_list[gp_id].converged;
}
template <>
bool micropp<3>::has_subiterated(int gp_id) const {
  assert(gp_id >= 0 && gp_id < ngp);
  return gp_list[gp_id].subiterated;
}
template <>
int micropp<3>::get_non_linear_gps(void) const {
  int count = 0;
  for (int gp = 0; gp < ngp; ++gp) count += gp_list[gp].allocated ? 1 : 0;
  return count;
}

template <>
void micropp<3>::print_info() const {
  cout << "micropp" << dim << " (B200 build: batched RVEs on GPU " << gpu_id << ", wave = " << engine->W << ")"
       << endl;
  cout << "Micro-structure   : " << micro_names[micro_type] << endl;
  cout << "MATRIX [%]        : " << Vm << endl;
  cout << "FIBER  [%]        : " << Vf << endl;
  cout << "FE_LINEAR         : " << gp_counter[FE_LINEAR] << " GPs" << endl;
  cout << "FE_ONE_WAY        : " << gp_counter[FE_ONE_WAY] << " GPs" << endl;
  cout << "FE_FULL           : " << gp_counter[FE_FULL] << " GPs" << endl;
  cout << "MIX_RULE_CHAMIS   : " << gp_counter[MIX_RULE_CHAMIS] << " GPs" << endl;
  cout << "USE A0            : " << use_A0 << endl;
  cout << "NUM SUBITS        : " << nsubiterations << endl;
  cout << "MPI RANK          : " << mpi_rank << endl;
  cout << "ngp :" << ngp << " nx :" << nx << " ny :" << ny << " nz :" << nz << " nn :" << nn << endl
       << "lx : " << lx << " ly : " << ly << " lz : " << lz << endl;
  cout << "geo_params:";
  for (int i = 0; i < num_geo_params; ++i) cout << " " << geo_params[i];
  cout << endl;
  for (int i = 0; i < MAX_MATERIALS; ++i) {
    if (material_list[i]) material_list[i]->print();
    cout << endl;
  }
  cout << endl << "ctan_lin_fe = " << endl;
  for (int i = 0; i < 6; ++i) {
    for (int j = 0; j < 6; ++j) cout << ctan_lin_fe[i * 6 + j] << "\t";
    cout << endl;
  }
  cout << endl;
}

// <gp_id> <non-linear> <cost> <converged> per homogenize() (src/output.cpp:265-282)
template <>
void micropp<3>::write_log() {
  ofstream_log << "log_id : " << log_id << endl;
  for (int gp_id = 0; gp_id < ngp; ++gp_id)
    ofstream_log << "\t" << gp_id << "\t" << gp_list[gp_id].allocated << "\t" << gp_list[gp_id].cost << "\t"
                 << gp_list[gp_id].converged << endl;
  log_id++;
}

// Restart files keep the reference's on-disk format (src/output.cpp:217-262, include/gp.hpp:107-124):
// per GP one byte `allocated`, then -- if allocated -- vars_n[nvars] and u_n[nndim] as raw doubles in the
// reference's array-of-structs layouts.  The device copies are converted on the way.
template <>
void micropp<3>::write_restart(const int restart_id) const {
  std::stringstream name;
  name << "micropp-restart-" << mpi_rank << "-" << restart_id << ".bin";
  ofstream file(name.str(), ios::out | ios::binary);
  std::vector<double> vars(nvars), u(nndim);
  for (int igp = 0; igp < ngp; ++igp) {
    const gp_t<3> &g = gp_list[igp];
    file.write((const char *)&g.allocated, sizeof(bool));
    if (g.allocated) {
      mgpu_gp_get_vars(engine->ctx, g.fe_index, 0, vars.data());
      mgpu_gp_get_u(engine->ctx, g.fe_index, 0, u.data());
      file.write((const char *)vars.data(), nvars * sizeof(double));
      file.write((const char *)u.data(), nndim * sizeof(double));
    }
  }
}

template <>
void micropp<3>::read_restart(const int restart_id) const {
  std::stringstream name;
  name << "micropp-restart-" << mpi_rank << "-" << restart_id << ".bin";
  ifstream file(name.str(), ios::in | ios::binary);
  if (!file.good()) {
    cerr << "micropp-b200: cannot open restart file " << name.str() << endl;
    return;
  }
  std::vector<double> vars(nvars), u(nndim);
  for (int igp = 0; igp < ngp; ++igp) {
    gp_t<3> &g = gp_list[igp];
    bool allocated = false;
    file.read((char *)&allocated, sizeof(bool));
    if (allocated) {
      // the payload is always consumed, whatever this Gauss point is here, so that later ones stay aligned
      file.read((char *)vars.data(), nvars * sizeof(double));
      file.read((char *)u.data(), nndim * sizeof(double));
    }
    if (!file.good()) {
      cerr << "micropp-b200: restart file " << name.str() << " is truncated at Gauss point " << igp << endl;
      return;
    }
    if (g.fe_index < 0) continue;  // no FE state on this Gauss point (FE_LINEAR / mixture rule)
    if (allocated) {
      mgpu_gp_set_vars(engine->ctx, g.fe_index, 0, vars.data());
      mgpu_gp_set_u(engine->ctx, g.fe_index, 0, u.data());
    } else if (g.allocated) {
      mgpu_gp_free_vars(engine->ctx, g.fe_index);  // back to "no history": stale device buffers must not survive
    }
    g.allocated = allocated;
  }
}

// ================================================================================================
// host-pointer FE stages (the reference's protected kernels), staged through slot 0
// ================================================================================================
namespace {
const int kSlot0 = 0;
void stage_begin(mpp_engine *e, const double *u, const double *vars_old) {
  const int g0 = -1;
  mgpu_bind_slots(e->ctx, 1, &kSlot0, &g0, nullptr);
  mgpu_set_list(e->ctx, mpp_engine::L_SUB, 1, &kSlot0);
  if (u) mgpu_stage_put_u(e->ctx, kSlot0, u);
  mgpu_stage_put_vars(e->ctx, kSlot0, 0, vars_old);
}
}  // namespace

template <>
void micropp<3>::set_displ_bc(const double strain[nvoi], double *u) {
  stage_begin(engine, u, nullptr);
  mgpu_set_slot_strain(engine->ctx, 1, &kSlot0, strain);
  mgpu_set_bc(engine->ctx, mpp_engine::L_SUB, 1);
  mgpu_stage_get_u(engine->ctx, kSlot0, u);
}

template <>
double micropp<3>::assembly_rhs(const double *u, const double *vars_old, double *b) {
  stage_begin(engine, u, vars_old);
  mgpu_asm_rhs(engine->ctx, mpp_engine::L_SUB, 1, 2);
  mgpu_stage_get_vec(engine->ctx, kSlot0, 0, b);
  mgpu_slot_state st;
  mgpu_fetch_state(engine->ctx, 1, &kSlot0, &st);
  return st.norm;
}

template <>
void micropp<3>::assembly_mat(ell_matrix *A, const double *u, const double *vars_old) {
  stage_begin(engine, u, vars_old);
  mgpu_asm_mat(engine->ctx, mpp_engine::L_SUB, 1, 0);
  mgpu_stage_get_mat(engine->ctx, kSlot0, A->vals);
}

template <>
newton_t micropp<3>::newton_raphson(ell_matrix *A, double *b, double *u, double *du, const double strain[nvoi],
                                    const double *vars_old) {
  stage_begin(engine, u, vars_old);
  mgpu_set_slot_strain(engine->ctx, 1, &kSlot0, strain);
  std::vector<newton_t> res;
  engine->newton_batch(mpp_engine::L_SUB, 1, &kSlot0, res);
  mgpu_stage_get_u(engine->ctx, kSlot0, u);
  if (b) mgpu_stage_get_vec(engine->ctx, kSlot0, 0, b);
  if (du) mgpu_stage_get_vec(engine->ctx, kSlot0, 1, du);
  if (A && A->vals && res[0].its > 0) {
    // the implicit operator never materialises the Jacobian: assemble it for the caller who asked to see it
    if (engine->implicit) mgpu_asm_mat(engine->ctx, mpp_engine::L_SUB, 1, 0);
    mgpu_stage_get_mat(engine->ctx, kSlot0, A->vals);
  }
  return res[0];
}

template <>
void micropp<3>::calc_ave_stress(const double *u, double stress_ave[nvoi], const double *vars_old) const {
  stage_begin(engine, u, vars_old);
  mgpu_ave_stress(engine->ctx, mpp_engine::L_SUB, 1);
  mgpu_fetch_stress(engine->ctx, 1, &kSlot0, stress_ave);
}

template <>
bool micropp<3>::calc_vars_new(const double *u, const double *vars_old, double *vars_new) const {
  stage_begin(engine, u, vars_old);
  mgpu_stage_put_vars(engine->ctx, kSlot0, 1, vars_new);  // uploads the caller's buffer, results overwrite it
  mgpu_clear_nl_flags(engine->ctx, mpp_engine::L_SUB, 1);
  mgpu_vars_new(engine->ctx, mpp_engine::L_SUB, 1, 1);
  mgpu_slot_state st;
  mgpu_fetch_state(engine->ctx, 1, &kSlot0, &st);
  if (mgpu_nvar(engine->ctx) > 0) mgpu_stage_get_vars_new(engine->ctx, kSlot0, vars_new);
  return st.nl_flag != 0;
}

// ================================================================================================
// not on the hot path: kept as thin host utilities / explicit "not provided" stubs
// ================================================================================================
template <>
void micropp<3>::get_stress(int gp, const double eps[nvoi], const double *vars_old, double stress_gp[nvoi], int ex,
                            int ey, int ez) const {
  const int e = glo_elem(ex, ey, ez);
  const double *vars = (vars_old) ? &vars_old[intvar_ix(e, gp, 0)] : nullptr;
  get_material(e)->get_stress(eps, stress_gp, vars);
}

// ------------------------------------------------------------------------------------------------
// VTU output (src/output.cpp:30-213): the GP's u_k and vars_n come back from HBM in the reference's layouts, the
// element averages are computed on the device (k_elem_fields), the file is written in the reference's format.
// ------------------------------------------------------------------------------------------------
template <>
void micropp<3>::calc_fields(double *u, double *vars_old) {
  stage_begin(engine, u, vars_old);
  mgpu_elem_fields(engine->ctx, kSlot0, ivol, elem_strain, elem_stress);
}

namespace {
// one <DataArray> of the ASCII VTU file; `body` writes the values
template <class F>
void vtu_array(std::ostream &os, const char *type, const char *name, int ncomp, F body, const char *type_pad = "",
               const char *close_pad = "") {
  os << "<DataArray type=\"" << type << "\"" << type_pad << " Name=\"" << name << "\" NumberOfComponents=\"" << ncomp
     << "\" format=\"ascii\"" << close_pad << ">" << endl;
  body();
}
}  // namespace

template <>
void micropp<3>::write_vtu(double *u, double *vars_old, const char *filename) {
  ofstream os(std::string(filename) + ".vtu");
  os << "<?xml version=\"1.0\"?>\n<VTKFile type=\"UnstructuredGrid\" version=\"0.1\" byte_order=\"LittleEndian\">\n"
     << "<UnstructuredGrid>\n<Piece NumberOfPoints=\"" << nn << "\" NumberOfCells=\"" << nelem << "\">\n<Points>\n"
     << "<DataArray type=\"Float64\" Name=\"Position\" NumberOfComponents=\"3\" format=\"ascii\">" << endl;
  os << scientific;
  for (int n = 0; n < nn; ++n) {
    const int i = n % nx, j = (n / nx) % ny, k = n / (nx * ny);
    os << i * dx << " " << j * dy << " " << k * dz << " \n";
  }
  os << "</DataArray>\n</Points>\n<Cells>\n" << endl;

  vtu_array(os, "Int32", "connectivity", 1, [&] {
    for (int e = 0; e < nelem; ++e) {
      const int ex = e % nex, ey = (e / nex) % ney, ez = e / (nex * ney);
      int nodes[8];
      get_elem_nodes(nodes, nx, ny, ex, ey, ez);
      for (int a = 0; a < npe; ++a) os << nodes[a] << ' ';
      os << "\n";
    }
    os << "</DataArray>" << endl;
  });
  vtu_array(os, "Int32", "offsets", 1, [&] {
    for (int e = 1; e <= nelem; ++e) os << e * npe << " ";
    os << "\n</DataArray>" << endl;
  });
  vtu_array(os, "UInt8", "types", 1, [&] {
    for (int e = 0; e < nelem; ++e) os << 12 << " ";  // VTK_HEXAHEDRON
    os << "\n</DataArray>" << endl;
  }, " ");
  os << "</Cells>" << endl;

  os << "<PointData Vectors=\"displ\" >" << endl;
  vtu_array(os, "Float64", "displ", 3, [&] {
    for (int n = 0; n < nn; ++n) os << u[n * 3] << " " << u[n * 3 + 1] << " " << u[n * 3 + 2] << " \n";
    os << "</DataArray>" << endl;
  }, "", " ");
  os << "</PointData>" << endl;

  os << "<CellData>" << endl;
  auto per_elem6 = [&](const double *f) {
    for (int e = 0; e < nelem; ++e) {
      for (int v = 0; v < nvoi; ++v) os << f[e * nvoi + v] << " ";
      os << "\n";
    }
    os << "</DataArray>\n";
  };
  vtu_array(os, "Float64", "strain", nvoi, [&] { per_elem6(elem_strain); });
  vtu_array(os, "Float64", "stress", nvoi, [&] { per_elem6(elem_stress); });
  vtu_array(os, "Int32", "elem_type", 1, [&] {
    for (int e = 0; e < nelem; ++e) os << elem_type[e] << " ";
    os << "\n</DataArray>" << endl;
  });
  // Gauss-point means of internal variables (0 without a state).  "plasticity" keeps the reference's expression:
  // s = sum of the squared plastic strains, reported value (s + sqrt(s)) / 8 (src/output.cpp:160-168)
  auto var_mean = [&](const char *name, auto per_element) {
    vtu_array(os, "Float64", name, 1, [&] {
      for (int e = 0; e < nelem; ++e) os << per_element(e) / npe << " ";
      os << "\n</DataArray>" << endl;
    });
  };
  auto gp_sum = [&](int e, int v) {
    double s = 0.0;
    if (vars_old)
      for (int gp = 0; gp < npe; ++gp) s += vars_old[intvar_ix(e, gp, v)];
    return s;
  };
  var_mean("plasticity", [&](int e) {
    double s = 0.0;
    if (vars_old)
      for (int gp = 0; gp < npe; ++gp)
        for (int v = 0; v < nvoi; ++v) s += vars_old[intvar_ix(e, gp, v)] * vars_old[intvar_ix(e, gp, v)];
    return s + sqrt(s);
  });
  var_mean("damage_e", [&](int e) { return gp_sum(e, 0); });
  var_mean("damage_D", [&](int e) { return gp_sum(e, 1); });
  var_mean("hardening", [&](int e) { return gp_sum(e, 6); });
  os << "</CellData>\n</Piece>\n</UnstructuredGrid>\n</VTKFile>" << endl;
}

namespace {
// u_k and vars_n of a GP in the reference's host layouts (vars empty = the reference's nullptr)
void fetch_gp_state(mpp_engine *engine, const gp_t<3> &g, int nndim, int nvars, std::vector<double> &u,
                    std::vector<double> &vars) {
  u.assign(nndim, 0.0);
  vars.clear();
  if (g.fe_index < 0) return;
  mgpu_gp_get_u(engine->ctx, g.fe_index, 1, u.data());
  if (g.allocated) {
    vars.resize(nvars);
    mgpu_gp_get_vars(engine->ctx, g.fe_index, 0, vars.data());
  }
}
}  // namespace

template <>
void micropp<3>::output(int gp_id, const char *filename) {
  assert(gp_id >= 0 && gp_id < ngp);
  std::vector<double> u, vars;
  fetch_gp_state(engine, gp_list[gp_id], nndim, nvars, u, vars);
  double *vp = vars.empty() ? nullptr : vars.data();
  calc_fields(u.data(), vp);
  write_vtu(u.data(), vp, filename);
}

// writes "micropp-<elem_global>-<time_step>.vtu" (src/output.cpp:44-69)
template <>
void micropp<3>::output2(const int gp_id, const int elem_global, const int time_step) {
  std::stringstream name;
  name << "micropp-" << elem_global << "-" << time_step;
  output(gp_id, name.str().c_str());
}

// ================================================================================================
// access for the C-ABI extension layer
// ================================================================================================
struct mpp_access {
  static mpp_engine *engine(micropp<3> *m) { return m->engine; }
  static int nelem(const micropp<3> *m) { return m->nelem; }
  static int nndim(const micropp<3> *m) { return m->nndim; }
  static int nvars(const micropp<3> *m) { return m->nvars; }
  static int ngp(const micropp<3> *m) { return m->ngp; }
  static const int *elem_type(const micropp<3> *m) { return m->elem_type; }
  static const double *bmat(const micropp<3> *m) { return &m->bmat[0][0][0]; }
  static const double *ctan_lin(const micropp<3> *m) { return m->ctan_lin_fe; }
  static gp_t<3> *gp(micropp<3> *m, int i) { return &m->gp_list[i]; }
  static double last_ms(const micropp<3> *m) { return m->last_homogenize_ms; }
  static void set_displ_bc(micropp<3> *m, const double *eps, double *u) { m->set_displ_bc(eps, u); }
  static double assembly_rhs(micropp<3> *m, const double *u, const double *v, double *b) {
    return m->assembly_rhs(u, v, b);
  }
  static void assembly_mat(micropp<3> *m, ell_matrix *A, const double *u, const double *v) {
    m->assembly_mat(A, u, v);
  }
  static newton_t newton(micropp<3> *m, double *u, const double *eps, const double *v) {
    return m->newton_raphson(nullptr, nullptr, u, nullptr, eps, v);
  }
  static void ave_stress(micropp<3> *m, const double *u, double *s, const double *v) { m->calc_ave_stress(u, s, v); }
  static bool vars_new(micropp<3> *m, const double *u, const double *vo, double *vn) {
    return m->calc_vars_new(u, vo, vn);
  }
};

#include "micropp_b200_ext.h"

extern "C" {

int micropp3x_nelem(const micropp3 *s) { return mpp_access::nelem((micropp<3> *)s->ptr); }
int micropp3x_nndim(const micropp3 *s) { return mpp_access::nndim((micropp<3> *)s->ptr); }
int micropp3x_wave_size(const micropp3 *s) { return mpp_access::engine((micropp<3> *)s->ptr)->W; }
int micropp3x_implicit_rows(const micropp3 *s) {
  mpp_engine *e = mpp_access::engine((micropp<3> *)s->ptr);
  return e->implicit ? mgpu_implicit_rows(e->ctx) : 0;
}
void micropp3x_get_elem_type(const micropp3 *s, int *out) {
  const micropp<3> *m = (micropp<3> *)s->ptr;
  memcpy(out, mpp_access::elem_type(m), sizeof(int) * mpp_access::nelem(m));
}
void micropp3x_get_bmat(const micropp3 *s, double *out) {
  memcpy(out, mpp_access::bmat((micropp<3> *)s->ptr), sizeof(double) * 8 * 6 * 24);
}
void micropp3x_get_ctan_lin(const micropp3 *s, double *out) {
  memcpy(out, mpp_access::ctan_lin((micropp<3> *)s->ptr), sizeof(double) * 36);
}
int micropp3x_get_u(const micropp3 *s, int gp, int which, double *out) {
  micropp<3> *m = (micropp<3> *)s->ptr;
  gp_t<3> *g = mpp_access::gp(m, gp);
  if (g->fe_index < 0) return 0;
  mgpu_gp_get_u(mpp_access::engine(m)->ctx, g->fe_index, which, out);
  return 1;
}
int micropp3x_get_vars(const micropp3 *s, int gp, int which, double *out) {
  micropp<3> *m = (micropp<3> *)s->ptr;
  gp_t<3> *g = mpp_access::gp(m, gp);
  if (g->fe_index < 0 || !g->allocated) return 0;
  mgpu_gp_get_vars(mpp_access::engine(m)->ctx, g->fe_index, which, out);
  return 1;
}

void micropp3_set_strains(micropp3 *s, const double *strain) {
  micropp<3> *m = (micropp<3> *)s->ptr;
  const int n = mpp_access::ngp(m);
  for (int g = 0; g < n; ++g) m->set_strain(g, strain + (size_t)g * 6);
}
void micropp3_get_stresses(const micropp3 *s, double *stress) {
  micropp<3> *m = (micropp<3> *)s->ptr;
  const int n = mpp_access::ngp(m);
  for (int g = 0; g < n; ++g) m->get_stress(g, stress + (size_t)g * 6);
}
void micropp3_get_ctans(const micropp3 *s, double *ctan) {
  micropp<3> *m = (micropp<3> *)s->ptr;
  const int n = mpp_access::ngp(m);
  for (int g = 0; g < n; ++g) m->get_ctan(g, ctan + (size_t)g * 36);
}

void micropp3x_set_displ_bc(micropp3 *s, const double *eps, double *u) {
  mpp_access::set_displ_bc((micropp<3> *)s->ptr, eps, u);
}
double micropp3x_assembly_rhs(micropp3 *s, const double *u, const double *vars_old, double *b) {
  return mpp_access::assembly_rhs((micropp<3> *)s->ptr, u, vars_old, b);
}
void micropp3x_assembly_mat(micropp3 *s, const double *u, const double *vars_old, double *vals) {
  ell_matrix A;
  memset(&A, 0, sizeof(A));
  A.vals = vals;
  mpp_access::assembly_mat((micropp<3> *)s->ptr, &A, u, vars_old);
}
void micropp3x_newton(micropp3 *s, const double *eps, const double *vars_old, double *u, int *out3) {
  const newton_t r = mpp_access::newton((micropp<3> *)s->ptr, u, eps, vars_old);
  out3[0] = r.its;
  out3[1] = r.solver_its;
  out3[2] = r.converged ? 1 : 0;
}
void micropp3x_ave_stress(micropp3 *s, const double *u, const double *vars_old, double *sig) {
  mpp_access::ave_stress((micropp<3> *)s->ptr, u, sig, vars_old);
}
int micropp3x_vars_new(micropp3 *s, const double *u, const double *vars_old, double *vars_new) {
  return mpp_access::vars_new((micropp<3> *)s->ptr, u, vars_old, vars_new) ? 1 : 0;
}

void micropp3x_prof_enable(micropp3 *s, int on) {
  mpp_access::engine((micropp<3> *)s->ptr)->profiling = on != 0;  // per-kernel events need plain stream launches
  mgpu_prof_enable(mpp_access::engine((micropp<3> *)s->ptr)->ctx, on);
}
void micropp3x_cg_history(micropp3 *s, int k) { mgpu_cg_history(mpp_access::engine((micropp<3> *)s->ptr)->ctx, k); }
int micropp3x_cg_history_read(micropp3 *s, int slot, double *out, int k) {
  return mgpu_cg_history_read(mpp_access::engine((micropp<3> *)s->ptr)->ctx, slot, out, k);
}
int micropp3x_hybrid_available(const micropp3 *s) {
  return mgpu_hybrid_available(mpp_access::engine((micropp<3> *)s->ptr)->ctx);
}
void micropp3x_prof_read(micropp3 *s, double *out9, int reset) {
  mgpu_prof_read(mpp_access::engine((micropp<3> *)s->ptr)->ctx, out9, reset);
}
double micropp3x_last_homogenize_ms(const micropp3 *s) { return mpp_access::last_ms((micropp<3> *)s->ptr); }
unsigned long long micropp3x_launch_count(const micropp3 *s) {
  return mgpu_launch_count(mpp_access::engine((micropp<3> *)s->ptr)->ctx);
}
int micropp3x_implicit_kernel(const micropp3 *s) {
  return mgpu_implicit_kernel(mpp_access::engine((micropp<3> *)s->ptr)->ctx);
}
int micropp3x_resident_info(const micropp3 *s, int *meta8) {
  int meta[8];
  mgpu_resident_info(mpp_access::engine((micropp<3> *)s->ptr)->ctx, meta);
  if (meta8)
    for (int q = 0; q < 8; ++q) meta8[q] = meta[q];
  return meta[0];
}
double micropp3x_prof_resident_ms(micropp3 *s, int reset) {
  return mgpu_prof_resident_ms(mpp_access::engine((micropp<3> *)s->ptr)->ctx, reset);
}
double micropp3x_apply_operator(micropp3 *s, const double *p, double *Ap, int op, int kernel) {
  micropp<3> *m = (micropp<3> *)s->ptr;
  mpp_engine *e = mpp_access::engine(m);
  const int slot0 = 0, none = -1;
  mgpu_bind_slots(e->ctx, 1, &slot0, &none, nullptr);
  mgpu_set_list(e->ctx, mpp_engine::L_SUB, 1, &slot0);
  if (op == 0) {
    mgpu_zero_u(e->ctx, mpp_engine::L_SUB, 1);
    mgpu_asm_mat(e->ctx, mpp_engine::L_SUB, 1, 0);
  }
  std::vector<double> zero((size_t)mpp_access::nndim(m), 0.0);
  mgpu_stage_put_vec(e->ctx, slot0, 2, zero.data());  // Ap of boundary rows is never written
  mgpu_stage_put_vec(e->ctx, slot0, 3, p);
  mgpu_apply_operator(e->ctx, mpp_engine::L_SUB, 1, op, kernel);
  mgpu_stage_get_vec(e->ctx, slot0, 2, Ap);
  mgpu_slot_state st;
  mgpu_fetch_state(e->ctx, 1, &slot0, &st);
  return st.pAp;
}
double micropp3x_bench_spmv(micropp3 *s, int nslots, int iters) {
  return mgpu_bench_spmv(mpp_access::engine((micropp<3> *)s->ptr)->ctx, nslots, iters);
}
void micropp3x_resident_timeline(micropp3 *s, int slot, long long *out1024) {
  mgpu_resident_timeline(mpp_access::engine((micropp<3> *)s->ptr)->ctx, slot, out1024);
}
double micropp3x_bench_resident(micropp3 *s, int nslots, int reps, int dbg) {
  return mgpu_bench_resident(mpp_access::engine((micropp<3> *)s->ptr)->ctx, nslots, reps, dbg);
}
double micropp3x_bench_imp_spmv(micropp3 *s, int nslots, int iters, int kern) {
  return mgpu_bench_imp_spmv(mpp_access::engine((micropp<3> *)s->ptr)->ctx, nslots, iters, kern);
}
}

// ================================================================================================
// z-slab of one large RVE (BASELINE configs[4]: a single RVE split over several GPUs).  The reference has
// no such mode; this builds the per-rank device context: local node planes [z0-halo, z1+halo)
// of the global nx x ny x nz grid, with the global cell sizes, B matrices, element classification and
// elastic element matrices.  The solver loop (halo exchange of p, all-reduce of the dot products) is driven
// by the caller through include/mgpu.h -- see micropp_b200/slab.py.
// ================================================================================================
extern "C" mgpu_ctx *micropp3x_slab_create(const micropp3_params *q, int z0, int z1, int device) {
  const int nx = q->size[0], ny = q->size[1], nzg = q->size[2];
  if (z0 < 0 || z1 > nzg || z1 - z0 < 1) {
    fprintf(stderr, "micropp-b200: bad slab [%d,%d) of %d planes\n", z0, z1, nzg);
    abort();
  }
  const int halo_lo = z0 > 0 ? 1 : 0, halo_hi = z1 < nzg ? 1 : 0;
  const int koff = z0 - halo_lo, nzl = (z1 + halo_hi) - koff;
  const double dx = 1.0 / (nx - 1), dy = 1.0 / (ny - 1), dz = 1.0 / (nzg - 1);
  const double wg = (dx * dy * dz) / 8;
  const double g = CONSTXG;
  const double xg[8][3] = {{-g, -g, -g}, {+g, -g, -g}, {+g, +g, -g}, {-g, +g, -g},
                           {-g, -g, +g}, {+g, -g, +g}, {+g, +g, +g}, {-g, +g, +g}};
  double bmat[8][6][24];
  for (int gp = 0; gp < 8; ++gp) bmat_at(xg[gp], dx, dy, dz, bmat[gp]);

  const int nex = nx - 1, ney = ny - 1, nezl = nzl - 1;
  std::vector<int> et((size_t)std::max(nex * ney * nezl, 1), 0);
  for (int ez = 0; ez < nezl; ++ez)
    for (int ey = 0; ey < ney; ++ey)
      for (int ex = 0; ex < nex; ++ex)
        et[((size_t)ez * ney + ey) * nex + ex] = mpp_elem_type(q->type, q->geo_params, dx, dy, dz, ex, ey, ez + koff);

  mgpu_config cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.nx = nx;
  cfg.ny = ny;
  cfg.nz = nzl;
  cfg.device = device;
  cfg.ngp = 0;
  cfg.elem_type = et.data();
  for (int gp = 0; gp < 8; ++gp)
    for (int a = 0; a < 8; ++a) {
      cfg.dsh[gp][a * 3 + 0] = bmat[gp][0][a * 3 + 0];
      cfg.dsh[gp][a * 3 + 1] = bmat[gp][1][a * 3 + 1];
      cfg.dsh[gp][a * 3 + 2] = bmat[gp][2][a * 3 + 2];
    }
  cfg.wg = wg;
  cfg.dx = dx;
  cfg.dy = dy;
  cfg.dz = dz;
  std::vector<double> ke(3 * 576, 0.0);
  for (int i = 0; i < MAX_MATERIALS; ++i) {
    material_base m;
    material_set(&m, q->mat_type[i], q->mat_E[i], q->mat_nu[i], q->mat_Ka[i], q->mat_Sy[i], q->mat_Xt[i]);
    const double vals[8] = {m.E, m.nu, m.Ka, m.Sy, m.k, m.mu, m.lambda, m.Xt};
    memcpy(cfg.mat[i], vals, sizeof(vals));
    cfg.mat_type[i] = m.type;
    // element matrix of the material's ELASTIC law: what an elastic material always has, and what a damage / plastic
    // element has below its threshold (src/material.cpp:111-164, 206-287: the linear branch is lambda, mu elasticity)
    elastic_element_matrix(bmat, m, wg, &ke[i * 576]);
  }
  cfg.ke_elastic = ke.data();
  cfg.nr_max_its = q->nr_max_its;
  cfg.nr_max_tol = q->nr_max_tol;
  cfg.nr_rel_tol = q->nr_rel_tol;
  cfg.cg_max_its = CG_MAX_ITS;
  cfg.cg_abs_tol = CG_ABS_TOL;
  cfg.cg_rel_tol = CG_REL_TOL;
  cfg.wave_cap = 1;
  cfg.implicit_elastic = 1;
  if (const char *env = getenv("MICROPP_IMPLICIT")) cfg.implicit_elastic = atoi(env) != 0;
  cfg.slab = 1;
  cfg.koff = koff;
  cfg.nz_glob = nzg;
  cfg.halo_lo = halo_lo;
  cfg.halo_hi = halo_hi;
  // element layer e (global) is averaged by the rank that owns node plane e
  cfg.ez_own_lo = z0 - koff;
  cfg.ez_own_hi = std::min(z1, nzg - 1) - koff;
  return mgpu_create(&cfg);
}
