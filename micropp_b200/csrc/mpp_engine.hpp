// mpp_engine.hpp -- internal: the batched Newton-Raphson / DPCG drivers shared by micropp_host.cpp and
// ell_host.cpp.  Host C++ that only enqueues work through the thin C-ABI of include/mgpu.h.
#pragma once

#include <algorithm>
#include <vector>

#include "mgpu.h"
#include "types.hpp"

struct mpp_engine {
  mgpu_ctx *ctx = nullptr;
  int W = 0;
  bool use_A0 = false;
  int its_with_A0 = 1;
  bool A0_ready = false;
  bool implicit = false;  // all-elastic RVE: DPCG on the implicit operator (mgpu_implicit), no Jacobian assembly
  bool hybrid = false;    // damage / plastic phase: slots mostly inside their linear regime take the hybrid operator
  int cg_chunk = 8;
  int cg_group = 0;       // slots per L2-resident group of a Newton/DPCG solve (0: the whole wave at once)
  bool use_graphs = true;  // one CUDA graph per Newton step (MICROPP_GRAPHS=0: plain stream launches)
  bool profiling = false;  // per-kernel CUDA-event timing needs plain launches

  // device slot lists: 0 = caller's list, 1 = Newton-active, 2/4 = CG-active (ping-pong), 3 = caller's subset
  enum { L_OUTER = 0, L_NEWTON = 1, L_CG_A = 2, L_SUB = 3, L_CG_B = 4, L_HYB = 6 };

  // DPCG on the slots of `list` (src/ell.cpp:66-122).  The loop condition is evaluated on the device
  // per slot; the host only learns how many slots are still iterating, every cg_chunk iterations.
  void cg_solve(int list, int n, int use_shared, bool generic = false) {
    if (!generic && use_shared == 3 && mgpu_resident(ctx)) {  // the whole solve inside one cluster per RVE
      mgpu_cg_resident(ctx, list, n);
      return;
    }
    mgpu_cg_init(ctx, list, n, generic ? 2 : use_shared);
    int cur = L_CG_A, other = L_CG_B;
    int nc = mgpu_compact(ctx, list, n, cur, 1);
    while (nc > 0) {
      for (int k = 0; k < cg_chunk; ++k) {
        if (generic)
          mgpu_spmv_generic(ctx, cur, nc, 0);
        else
          mgpu_cg_spmv_dot(ctx, cur, nc, use_shared);
        mgpu_cg_update(ctx, cur, nc);
        mgpu_cg_pupdate(ctx, cur, nc);
      }
      nc = mgpu_compact(ctx, cur, nc, other, 1);
      std::swap(cur, other);
    }
    mgpu_cg_finish(ctx, list, n);  // the deferred x += alpha p of every slot's last iteration
  }

  // Newton-Raphson on the slots of `list` (src/solve.cpp:29-82); u and the strain of every slot must
  // already be in place.  All slots advance together; a slot that converged (or exhausted
  // nr_max_its) simply drops out of the active list.
  void newton_batch(int list, int n, const int *slots, std::vector<newton_t> &out) {
    mgpu_set_bc(ctx, list, n);
    mgpu_asm_rhs(ctx, list, n, 0);
    // The wave is solved group by group (cg_group slots at a time, 0 = all at once): the DPCG vectors of a group
    // (p, r, Ap, du: 96 B per node and slot) then stay in the 126 MB L2 from one iteration to the next instead of
    // streaming through HBM.  Slots are independent, so the grouping does not change any result.
    const int G = (cg_group > 0 && use_graphs && !profiling) ? cg_group : n;
    for (int off = 0; off < n; off += G) newton_group(list, off, std::min(G, n - off));
    std::vector<mgpu_slot_state> st(n);
    mgpu_fetch_state(ctx, n, slots, st.data());
    out.resize(n);
    for (int i = 0; i < n; ++i) {
      out[i].its = st[i].nr_its;
      out[i].solver_its = st[i].solver_its;
      out[i].converged = st[i].converged != 0;
    }
  }

  void newton_group(int list, int off, int n) {
    int it = 0;
    for (;;) {
      const int na = mgpu_compact_range(ctx, list, off, n, L_NEWTON, 0);
      if (na == 0) break;
      // linear Jacobian for the first its_with_A0 iterations (src/solve.cpp:56-66)
      // an all-elastic Jacobian does not depend on u (src/material.cpp:84-94): the implicit operator IS the
      // matrix assembly_mat would build, for every slot and every step (which also covers the A0 shortcut)
      auto op_of = [&](int step) { return implicit ? 3 : ((use_A0 && A0_ready && step <= its_with_A0 - 1) ? 1 : 0); };
      const int shared = op_of(it);
      if (hybrid && shared == 0) {
        // one Newton step at a time: every step re-probes which elements have left their linear regime, builds the
        // node lists and sends every slot to the hybrid or to the fully assembled operator
        int nh = 0, nf = 0;
        mgpu_hybrid_split(ctx, L_NEWTON, na, &nh, &nf);
        if (use_graphs && !profiling) {
          mgpu_newton_step_graph_on(ctx, 1, nh, 4);
          mgpu_newton_step_graph_on(ctx, 0, nf, 0);
        } else {
          if (nh > 0) {
            mgpu_asm_mat_hyb(ctx, L_HYB, nh);
            cg_solve(L_HYB, nh, 4);
            mgpu_axpy_u(ctx, L_HYB, nh);
            mgpu_asm_rhs(ctx, L_HYB, nh, 1);
          }
          if (nf > 0) {
            mgpu_asm_mat(ctx, L_NEWTON, nf, 0);
            cg_solve(L_NEWTON, nf, 0);
            mgpu_axpy_u(ctx, L_NEWTON, nf);
            mgpu_asm_rhs(ctx, L_NEWTON, nf, 1);
          }
        }
        ++it;
        continue;
      }
      if (use_graphs && !profiling) {
        int left = na;
        while (left > 0) {  // every graph launch is one Newton step of all still-active slots
          left = mgpu_newton_step_graph(ctx, left, op_of(it));
          ++it;
        }
        break;
      }
      if (!shared) mgpu_asm_mat(ctx, L_NEWTON, na, 0);
      cg_solve(L_NEWTON, na, shared);
      mgpu_axpy_u(ctx, L_NEWTON, na);
      mgpu_asm_rhs(ctx, L_NEWTON, na, 1);
      ++it;
    }
  }
};

