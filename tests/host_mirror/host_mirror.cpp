// host_mirror.cpp -- TEST INFRASTRUCTURE ONLY.
// Compiles the product's device headers (gorilla_b200/csrc/gb_*.cuh) for the HOST with g++
// (-ffp-contract=off == nvcc --fmad=false) so that the device algorithm can be checked bit-for-bit
// against the oracle on a machine without a GPU (`pytest -m "not gpu"`).  It is built into
// tests/_build/ by tests/conftest.py, is never shipped and is never loaded by the product: the
// product library has no CPU path.  The per-particle loop below restates orbit_kernel's lane logic
// (gb_internal.cuh) without the warp machinery.
#include <cstdint>
#include <cstring>
#include <vector>
#include "../../gorilla_b200/csrc/gb_find.cuh"
#include "../../gorilla_b200/csrc/gb_repack.hpp"

using namespace gb;

struct HostMirror {
  std::vector<double> geom, bpart, phi, cold, se, ham, skew, poly4;
  gb::FindBins bins;
  MeshDev m;
  int poly_order, boole_periodic_relocation, ipusher, adaptive;
  unsigned oq_mask;
};

// orbit events of one particle (mirrors the Batch event fields + emit_events of gb_internal.cuh)
struct HmEvents {
  int flags, nskip_p, nskip_v;
  double *J;
  int32_t *cnt_v, *cnt_p;
  gorilla_event *events;
  int64_t cap, *n_events, particle;
  int nskip_f;     // flags bit3: boole_full_orbit, every nskip_f-th push
  double t_step;
};
static void hm_emit(HmEvents *ev, int64_t push, const EvState &es, double t)
{
  for (int k = 0; k < es.n; k++) {
    const int64_t slot = (*ev->n_events)++;
    if (slot < ev->cap) {
      gorilla_event *e = ev->events + slot;
      e->particle = ev->particle; e->kind = es.e[k].kind; e->counter = es.e[k].counter; e->push = push;
      for (int i = 0; i < 3; i++) e->x[i] = es.e[k].x[i];
      e->value[0] = es.e[k].v[0]; e->value[1] = es.e[k].v[1];
      e->t = t;
    }
  }
}

template <int K, int PHI, int EXT = 0>
static void run_particle(const MeshDev &m, double *x, double *vpar_io, double *vperp_io, double t_step, int32_t *ind_io,
                         int32_t *iface_io, double *t_remain_out, int64_t *npush_out, int trace_cap, int32_t *tr_t,
                         int32_t *tr_f, int force_full, int64_t *fallback, unsigned oq_mask = 0, double *optq = nullptr,
                         HmEvents *ev = nullptr)
{
  double oq_acc[4] = {0.0, 0.0, 0.0, 0.0};
  int32_t ind_tetr = *ind_io, iface = *iface_io;
  double vpar = *vpar_io;
  const double vperp = *vperp_io;
  const double *pg = m.geom + ((int64_t)ind_tetr - 1) * GEOM_ND;
  double z_save[3] = {x[0] - pg[0], x[1] - pg[1], x[2] - pg[2]};
  const double perpinv = -0.5 * (vperp * vperp) / bmod_at<PHI>(m, ind_tetr, z_save);
  double t_remain = t_step;
  int32_t ind_save = ind_tetr;
  int64_t npush = 0;
  for (;;) {
    ind_save = ind_tetr;
    PushOut o;
    bool done = false;
    if constexpr (K == 0) {
      if (!force_full) {
        RkPusher<PHI, (EXT == 2 ? 2 : 0)> R;
        double stash[6];
        R.P.r.set_stash(stash, 1);
        R.init(&m, perpinv, ind_tetr, x, iface, vpar, t_remain);
        done = R.template push<true>(o);
      }
      if (!done) o = push_rk_full_call<PHI, (EXT == 2 ? 2 : 0)>(&m, perpinv, ind_tetr, iface, x[0], x[1], x[2], vpar, t_remain);
      if constexpr (EXT == 2) {
        if (ev && !o.finished) {
          EvState es;
          es.flags = ev->flags; es.nskip_p = ev->nskip_p; es.nskip_v = ev->nskip_v;
          es.J = *ev->J; es.cnt_v = *ev->cnt_v; es.cnt_p = *ev->cnt_p; es.n = 0;
          es = rk_events_call<PHI>(&m, perpinv, ind_tetr, iface, x[0], x[1], x[2], vpar, t_remain, o, es);
          *ev->J = es.J; *ev->cnt_v = es.cnt_v; *ev->cnt_p = es.cnt_p;
          if (es.n) hm_emit(ev, npush, es, ev->t_step - (t_remain - o.t_pass));
        }
      }
    } else {
      if constexpr (EXT != 5) {   // EXT = 5 (adaptive steps + list consumers) always takes the complete path
      if (!force_full) {
        PolyPusher<K, PHI, EXT> P;
        double stash[6];
        P.r.set_stash(stash, 1);
        P.mp = &m;
        P.perpinv = perpinv;
        if (EXT == 2) P.oq_mask = oq_mask;
        done = P.push_fast(ind_tetr, iface, x, vpar, t_remain, o);
        if (EXT == 2 && done && oq_mask)
          for (int q = 0; q < 4; q++) oq_acc[q] = oq_acc[q] + P.oq[q];
        if constexpr (EXT == 2 && K >= 2) {
          if (done && ev && !o.finished) {
            EvState es;
            es.flags = ev->flags; es.nskip_p = ev->nskip_p; es.nskip_v = ev->nskip_v;
            es.J = *ev->J; es.cnt_v = *ev->cnt_v; es.cnt_p = *ev->cnt_p;
            P.events_after_push(vpar, o, es);
            *ev->J = es.J; *ev->cnt_v = es.cnt_v; *ev->cnt_p = es.cnt_p;
            if (es.n) hm_emit(ev, npush, es, ev->t_step - (t_remain - o.t_pass));
          }
        }
      }
      }
      if (!done) {
        if constexpr (EXT == 2 || EXT == 5) {
          static thread_local std::vector<double> lst;
          const int lst_cap = EXT == 5 ? 3 * m.max_n_intermediate_steps : 0;
          if (EXT == 5 && lst.size() < (size_t)lst_cap * 5) lst.resize((size_t)lst_cap * 5);
          const PushOutX ox = push_full_call_x<K, PHI, EXT>(&m, perpinv, ind_tetr, iface, x[0], x[1], x[2], vpar, t_remain, oq_mask,
                                                            ev ? ev->flags : 0, ev ? ev->nskip_p : 1, ev ? ev->nskip_v : 1,
                                                            ev ? *ev->J : 0.0, ev ? *ev->cnt_v : 0, ev ? *ev->cnt_p : 0,
                                                            EXT == 5 ? lst.data() : nullptr, lst_cap);
          o = ox.o;
          for (int q = 0; q < 4; q++) oq_acc[q] = oq_acc[q] + ox.oq[q];
          if (ev) {
            *ev->J = ox.es.J; *ev->cnt_v = ox.es.cnt_v; *ev->cnt_p = ox.es.cnt_p;
            if (ox.es.n) hm_emit(ev, npush, ox.es, ev->t_step - (t_remain - o.t_pass));
          }
        } else {
          o = push_full_call<K, PHI, EXT>(&m, perpinv, ind_tetr, iface, x[0], x[1], x[2], vpar, t_remain);
        }
      }
    }
    x[0] = o.x[0]; x[1] = o.x[1]; x[2] = o.x[2];
    vpar = o.vpar;
    if (o.z_save_set) { z_save[0] = o.z_save[0]; z_save[1] = o.z_save[1]; z_save[2] = o.z_save[2]; }
    ind_tetr = o.ind_tetr;
    iface = o.iface;
    if (trace_cap > 0 && npush < trace_cap) { tr_t[npush] = ind_tetr; tr_f[npush] = iface; }
    npush++;
    for (int b = 0; b < 5; b++) if (o.fallback & (1 << b)) fallback[b]++;
    t_remain = t_remain - o.t_pass;
    if (ev && (ev->flags & 8)) {   // as lane_after_push (gb_internal.cuh): boole_full_orbit
      const long long cnt = npush;   // already incremented: counter_tetrahedron_passes
      if (cnt / ev->nskip_f * ev->nskip_f == cnt) {
        const int64_t slot = (*ev->n_events)++;
        if (slot < ev->cap) {
          gorilla_event *e = ev->events + slot;
          e->particle = ev->particle; e->kind = GORILLA_EVENT_FULL_ORBIT; e->counter = (int32_t)cnt; e->push = npush - 1;
          e->x[0] = o.x[0]; e->x[1] = o.x[1]; e->x[2] = o.x[2];
          orbit_point_invariants(m, ind_save, z_save, o.vpar, perpinv, e->value[0], e->value[1]);
          e->t = ev->t_step - t_remain;
        }
      }
    }
    if (o.finished || ind_tetr == -1) break;
  }
  double vperp_new = 0.0;
  if (perpinv != 0.0) vperp_new = sqrt(2.0 * fabs(perpinv) * bmod_at<PHI>(m, ind_save, z_save));
  *vpar_io = vpar;
  *vperp_io = vperp_new;
  *ind_io = ind_tetr;
  *iface_io = iface;
  if (t_remain_out) *t_remain_out = t_remain;
  if (npush_out) *npush_out = npush;
  if (optq)
    for (int q = 0; q < 4; q++) optq[q] = oq_acc[q];
}

extern "C" {

void *hm_create(const gorilla_mesh_desc *md, int poly_order, int boole_guess, int boole_periodic_relocation, int ipusher,
                int boole_strong_electric_field, int i_time_tracing_option, int oq_mask, int boole_adaptive_time_steps,
                double desired_delta_energy, int max_n_intermediate_steps, int handover_processing_kind, int i_precomp,
                int boole_newton_precalc, int boole_pusher_ode45, double rel_err_ode45)
{
  HostMirror *h = new HostMirror();
  bool has_phi = false;
  const bool strong = boole_strong_electric_field != 0;
  if (!repack_mesh(md, h->geom, h->bpart, h->phi, h->cold, has_phi, strong ? &h->se : nullptr)) { delete h; return nullptr; }
  MeshDev &m = h->m;
  memset(&m, 0, sizeof(m));
  m.ntetr = md->ntetr;
  m.geom = h->geom.data(); m.bpart = h->bpart.data(); m.phi = (has_phi || strong) ? h->phi.data() : nullptr; m.cold = h->cold.data();
  m.se = strong ? h->se.data() : nullptr;
  repack_hamiltonian_time(md, h->ham);
  m.ham = h->ham.data();
  m.time_tracing = i_time_tracing_option;
  h->oq_mask = (unsigned)oq_mask;
  h->adaptive = boole_adaptive_time_steps;
  if (handover_processing_kind == 2 && md->tetra_skew_coord) {
    repack_skew(md, h->skew);
    m.skew = h->skew.data();
  }
  m.desired_delta_energy = desired_delta_energy;
  m.max_n_intermediate_steps = max_n_intermediate_steps;
  if ((ipusher == 2 && i_precomp != 0) || (ipusher == 1 && boole_newton_precalc)) {
    make_precomp_poly4(md, h->poly4);
    m.poly4 = h->poly4.data();
    m.i_precomp = ipusher == 2 ? i_precomp : 0;
    m.newton_precalc = ipusher == 1 ? 1 : 0;
  }
  m.ode45 = (ipusher == 1 && boole_pusher_ode45) ? 1 : 0;
  m.rel_err_ode45 = rel_err_ode45;
  if (build_find_bins(md, h->bins)) {
    m.bin_start = h->bins.start.data(); m.bin_items = h->bins.items.data();
    m.bin_nu = h->bins.nu; m.bin_nv = h->bins.nv; m.bin_c0 = h->bins.c0; m.bin_c1 = h->bins.c1;
    m.bin_u0 = h->bins.u0; m.bin_v0 = h->bins.v0; m.bin_du_inv = h->bins.du_inv; m.bin_dv_inv = h->bins.dv_inv;
  }
  m.cm_over_e = md->cm_over_e; m.particle_mass = md->particle_mass; m.particle_charge = md->particle_charge;
  const double PI = 3.141592653589793238462643383;
  m.period_phi = 2.0 * PI / md->n_field_periods; m.period_theta = 2.0 * PI;
  m.sign_sqg = md->sign_sqg; m.coord_system = md->coord_system;
  m.grid_size1 = md->grid_size[0]; m.grid_size2 = md->grid_size[1]; m.grid_size3 = md->grid_size[2];
  m.boole_guess = boole_guess; m.grid_kind = md->grid_kind; m.n_field_periods = md->n_field_periods;
  m.Rmin = md->Rmin; m.Rmax = md->Rmax; m.Zmin = md->Zmin; m.Zmax = md->Zmax; m.sfc_s_min = md->sfc_s_min;
  h->poly_order = ipusher == 1 ? 0 : poly_order;
  h->ipusher = ipusher;
  h->boole_periodic_relocation = boole_periodic_relocation;
  return h;
}
void hm_free(void *p) { delete (HostMirror *)p; }
int hm_has_phi(void *p) { return ((HostMirror *)p)->m.phi != nullptr; }

// same contract as gorilla_b200_orbit_timestep_trace; returns number of domain errors
int64_t hm_orbit_timestep(void *p, int64_t n, double *x, double *vpar, double *vperp, double t_step, int32_t *binit,
                          int32_t *ind_tetr, int32_t *iface, double *t_remain_out, int64_t *n_pushes, int32_t trace_cap,
                          int32_t *tr_t, int32_t *tr_f, int force_full, int64_t *fallback /*[4]*/, double *optq /*[n][4] or NULL*/)
{
  HostMirror *h = (HostMirror *)p;
  const MeshDev &m = h->m;
  int64_t dom = 0;
  const int sign_t = signbit(t_step) ? -1 : 1;
  for (int64_t i = 0; i < n; i++) {
    double *xi = x + 3 * i;
    if (!binit[i]) {
      int32_t it = -1, ifc = -1;
      if (check_coordinate_domain(m, xi, h->boole_periodic_relocation) != 0) dom++;
      else if (m.se) find_tetra<2>(&m, xi, vpar[i], vperp[i], it, ifc, sign_t);
      else if (m.phi) find_tetra<1>(&m, xi, vpar[i], vperp[i], it, ifc, sign_t);
      else find_tetra<0>(&m, xi, vpar[i], vperp[i], it, ifc, sign_t);
      ind_tetr[i] = it; iface[i] = ifc;
      if (it != -1) binit[i] = 1;
    }
    if (n_pushes) n_pushes[i] = 0;
    if (t_remain_out) t_remain_out[i] = t_step;
    if (optq) optq[4 * i] = optq[4 * i + 1] = optq[4 * i + 2] = optq[4 * i + 3] = 0.0;
    if (!binit[i] || ind_tetr[i] < 1) continue;
    if (t_step == 0.0) { if (t_remain_out) t_remain_out[i] = 0.0; continue; }
    int32_t *tt = trace_cap > 0 ? tr_t + i * trace_cap : nullptr, *tf = trace_cap > 0 ? tr_f + i * trace_cap : nullptr;
#define HM_RUN(K, PHI) run_particle<K, PHI>(m, xi, vpar + i, vperp + i, t_step, ind_tetr + i, iface + i, \
      t_remain_out ? t_remain_out + i : nullptr, n_pushes ? n_pushes + i : nullptr, trace_cap, tt, tf, force_full, fallback)
#define HM_RUNX(K, PHI) run_particle<K, PHI, 2>(m, xi, vpar + i, vperp + i, t_step, ind_tetr + i, iface + i, \
      t_remain_out ? t_remain_out + i : nullptr, n_pushes ? n_pushes + i : nullptr, trace_cap, tt, tf, force_full, fallback, \
      optq ? h->oq_mask : 0u, optq ? optq + 4 * i : nullptr)
#define HM_RUNA(K, PHI) run_particle<K, PHI, 3>(m, xi, vpar + i, vperp + i, t_step, ind_tetr + i, iface + i, \
      t_remain_out ? t_remain_out + i : nullptr, n_pushes ? n_pushes + i : nullptr, trace_cap, tt, tf, force_full, fallback)
#define HM_RUNP(K, PHI) run_particle<K, PHI, 4>(m, xi, vpar + i, vperp + i, t_step, ind_tetr + i, iface + i, \
      t_remain_out ? t_remain_out + i : nullptr, n_pushes ? n_pushes + i : nullptr, trace_cap, tt, tf, force_full, fallback)
    if (h->ipusher == 2 && m.i_precomp != 0) {   // precomputed coefficients (no strong-E variant)
      if (m.phi) {
        switch (h->poly_order) { case 2: HM_RUNP(2, 1); break; case 3: HM_RUNP(3, 1); break; default: HM_RUNP(4, 1); }
      } else {
        switch (h->poly_order) { case 2: HM_RUNP(2, 0); break; case 3: HM_RUNP(3, 0); break; default: HM_RUNP(4, 0); }
      }
    } else
#define HM_RUN5(K, PHI) run_particle<K, PHI, 5>(m, xi, vpar + i, vperp + i, t_step, ind_tetr + i, iface + i, \
      t_remain_out ? t_remain_out + i : nullptr, n_pushes ? n_pushes + i : nullptr, trace_cap, tt, tf, force_full, fallback, \
      optq ? h->oq_mask : 0u, optq ? optq + 4 * i : nullptr)
    if (h->ipusher == 2 && h->adaptive && ((optq && h->oq_mask) || m.time_tracing == 2)) {   // same dispatch as launch_orbit_k
      if (m.se) {
        switch (h->poly_order) { case 1: HM_RUN5(1, 2); break; case 2: HM_RUN5(2, 2); break; case 3: HM_RUN5(3, 2); break; default: HM_RUN5(4, 2); }
      } else if (m.phi) {
        switch (h->poly_order) { case 1: HM_RUN5(1, 1); break; case 2: HM_RUN5(2, 1); break; case 3: HM_RUN5(3, 1); break; default: HM_RUN5(4, 1); }
      } else {
        switch (h->poly_order) { case 1: HM_RUN5(1, 0); break; case 2: HM_RUN5(2, 0); break; case 3: HM_RUN5(3, 0); break; default: HM_RUN5(4, 0); }
      }
    } else
    if (h->ipusher == 2 && h->adaptive) {
      if (m.se) {
        switch (h->poly_order) { case 1: HM_RUNA(1, 2); break; case 2: HM_RUNA(2, 2); break; case 3: HM_RUNA(3, 2); break; default: HM_RUNA(4, 2); }
      } else if (m.phi) {
        switch (h->poly_order) { case 1: HM_RUNA(1, 1); break; case 2: HM_RUNA(2, 1); break; case 3: HM_RUNA(3, 1); break; default: HM_RUNA(4, 1); }
      } else {
        switch (h->poly_order) { case 1: HM_RUNA(1, 0); break; case 2: HM_RUNA(2, 0); break; case 3: HM_RUNA(3, 0); break; default: HM_RUNA(4, 0); }
      }
    } else
#define HM_RUNT(K, PHI) run_particle<K, PHI, 1>(m, xi, vpar + i, vperp + i, t_step, ind_tetr + i, iface + i, \
      t_remain_out ? t_remain_out + i : nullptr, n_pushes ? n_pushes + i : nullptr, trace_cap, tt, tf, force_full, fallback)
    if (h->ipusher == 2 && m.time_tracing == 2 && !(optq && h->oq_mask) && !m.skew) {   // same dispatch as launch_orbit_k
      if (m.se) {
        switch (h->poly_order) { case 1: HM_RUNT(1, 2); break; case 2: HM_RUNT(2, 2); break; case 3: HM_RUNT(3, 2); break; default: HM_RUNT(4, 2); }
      } else if (m.phi) {
        switch (h->poly_order) { case 1: HM_RUNT(1, 1); break; case 2: HM_RUNT(2, 1); break; case 3: HM_RUNT(3, 1); break; default: HM_RUNT(4, 1); }
      } else {
        switch (h->poly_order) { case 1: HM_RUNT(1, 0); break; case 2: HM_RUNT(2, 0); break; case 3: HM_RUNT(3, 0); break; default: HM_RUNT(4, 0); }
      }
    } else
    if (h->ipusher == 1 && (m.skew || m.newton_precalc || m.ode45)) {
      if (m.se) HM_RUNX(0, 2); else if (m.phi) HM_RUNX(0, 1); else HM_RUNX(0, 0);
    } else
    if (h->ipusher == 2 && ((optq && h->oq_mask) || m.skew)) {
      if (m.se) {
        switch (h->poly_order) { case 1: HM_RUNX(1, 2); break; case 2: HM_RUNX(2, 2); break; case 3: HM_RUNX(3, 2); break; default: HM_RUNX(4, 2); }
      } else if (m.phi) {
        switch (h->poly_order) { case 1: HM_RUNX(1, 1); break; case 2: HM_RUNX(2, 1); break; case 3: HM_RUNX(3, 1); break; default: HM_RUNX(4, 1); }
      } else {
        switch (h->poly_order) { case 1: HM_RUNX(1, 0); break; case 2: HM_RUNX(2, 0); break; case 3: HM_RUNX(3, 0); break; default: HM_RUNX(4, 0); }
      }
    } else
    if (m.se) {
      switch (h->poly_order) { case 0: HM_RUN(0, 2); break; case 1: HM_RUN(1, 2); break; case 2: HM_RUN(2, 2); break; case 3: HM_RUN(3, 2); break; default: HM_RUN(4, 2); }
    } else if (m.phi) {
      switch (h->poly_order) { case 0: HM_RUN(0, 1); break; case 1: HM_RUN(1, 1); break; case 2: HM_RUN(2, 1); break; case 3: HM_RUN(3, 1); break; default: HM_RUN(4, 1); }
    } else {
      switch (h->poly_order) { case 0: HM_RUN(0, 0); break; case 1: HM_RUN(1, 0); break; case 2: HM_RUN(2, 0); break; case 3: HM_RUN(3, 0); break; default: HM_RUN(4, 0); }
    }
  }
  return dom;
}

// same contract as gorilla_b200_orbit_timestep_events (particles must be localised already); returns 0 or -1 (bad config)
int64_t hm_orbit_timestep_events(void *p, int64_t n, double *x, double *vpar, double *vperp, double t_step, int32_t *binit,
                                 int32_t *ind_tetr, int32_t *iface, double *t_remain_out, int64_t *n_pushes, int flags,
                                 int nskip_p, int nskip_v, double *J, int32_t *cnt_v, int32_t *cnt_p, gorilla_event *events,
                                 int64_t cap, int64_t *n_events, int force_full, int nskip_f)
{
  HostMirror *h = (HostMirror *)p;
  const MeshDev &m = h->m;
  if (h->ipusher == 2 && h->poly_order < 2) return -1;
  const int sign_t = signbit(t_step) ? -1 : 1;
  int64_t fallback[5] = {0, 0, 0, 0, 0};
  *n_events = 0;
  for (int64_t i = 0; i < n; i++) {
    double *xi = x + 3 * i;
    if (!binit[i]) {
      int32_t it = -1, ifc = -1;
      if (check_coordinate_domain(m, xi, h->boole_periodic_relocation) == 0) {
        if (m.se) find_tetra<2>(&m, xi, vpar[i], vperp[i], it, ifc, sign_t);
        else if (m.phi) find_tetra<1>(&m, xi, vpar[i], vperp[i], it, ifc, sign_t);
        else find_tetra<0>(&m, xi, vpar[i], vperp[i], it, ifc, sign_t);
      }
      ind_tetr[i] = it; iface[i] = ifc;
      if (it != -1) binit[i] = 1;
    }
    if (n_pushes) n_pushes[i] = 0;
    if (t_remain_out) t_remain_out[i] = t_step;
    if (!binit[i] || ind_tetr[i] < 1) continue;
    if (t_step == 0.0) { if (t_remain_out) t_remain_out[i] = 0.0; continue; }
    HmEvents ev = {flags, nskip_p, nskip_v, J + i, cnt_v + i, cnt_p + i, events, cap, n_events, i, nskip_f > 0 ? nskip_f : 1, t_step};
#define HM_RUNE(K, PHI) run_particle<K, PHI, 2>(m, xi, vpar + i, vperp + i, t_step, ind_tetr + i, iface + i, \
      t_remain_out ? t_remain_out + i : nullptr, n_pushes ? n_pushes + i : nullptr, 0, nullptr, nullptr, force_full, fallback, \
      0u, nullptr, &ev)
#define HM_RUNE5(K, PHI) run_particle<K, PHI, 5>(m, xi, vpar + i, vperp + i, t_step, ind_tetr + i, iface + i, \
      t_remain_out ? t_remain_out + i : nullptr, n_pushes ? n_pushes + i : nullptr, 0, nullptr, nullptr, force_full, fallback, \
      0u, nullptr, &ev)
    if (h->ipusher == 1) {
      if (m.se) HM_RUNE(0, 2); else if (m.phi) HM_RUNE(0, 1); else HM_RUNE(0, 0);
    } else if (h->adaptive) {
      if (m.se) {
        switch (h->poly_order) { case 2: HM_RUNE5(2, 2); break; case 3: HM_RUNE5(3, 2); break; default: HM_RUNE5(4, 2); }
      } else if (m.phi) {
        switch (h->poly_order) { case 2: HM_RUNE5(2, 1); break; case 3: HM_RUNE5(3, 1); break; default: HM_RUNE5(4, 1); }
      } else {
        switch (h->poly_order) { case 2: HM_RUNE5(2, 0); break; case 3: HM_RUNE5(3, 0); break; default: HM_RUNE5(4, 0); }
      }
    } else if (m.se) {
      switch (h->poly_order) { case 2: HM_RUNE(2, 2); break; case 3: HM_RUNE(3, 2); break; default: HM_RUNE(4, 2); }
    } else if (m.phi) {
      switch (h->poly_order) { case 2: HM_RUNE(2, 1); break; case 3: HM_RUNE(3, 1); break; default: HM_RUNE(4, 1); }
    } else {
      switch (h->poly_order) { case 2: HM_RUNE(2, 0); break; case 3: HM_RUNE(3, 0); break; default: HM_RUNE(4, 0); }
    }
  }
  return 0;
}

// device math / solver entry points for bit-level pinning against glibc and the oracle
void hm_csqrt(double re, double im, double *out) { cd r = csqrt_glibc(mk(re, im)); out[0] = r.re; out[1] = r.im; }
double hm_hypot(double a, double b) { return hypot_glibc(a, b); }
void hm_cdiv(double ar, double ai, double br, double bi, double *out) { cd r = cdiv(mk(ar, ai), mk(br, bi)); out[0] = r.re; out[1] = r.im; }
void hm_cmul(double ar, double ai, double br, double bi, double *out) { cd r = cmul(mk(ar, ai), mk(br, bi)); out[0] = r.re; out[1] = r.im; }
void hm_rmul(double r0, double br, double bi, double *out) { cd r = rmul(r0, mk(br, bi)); out[0] = r.re; out[1] = r.im; }
void hm_frac_jump_phase(int k, double *out) { cd r = frac_jump_phase(k); out[0] = r.re; out[1] = r.im; }
void hm_cmplx_roots_gen(int degree, const double *poly_re_im, double *roots_re_im)
{
  cd poly[5], roots[4];
  int iters = 0;
  for (int i = 0; i <= degree; i++) poly[i] = mk(poly_re_im[2 * i], poly_re_im[2 * i + 1]);
  sg_roots(degree, poly, roots, iters);
  for (int i = 0; i < degree; i++) { roots_re_im[2 * i] = roots[i].re; roots_re_im[2 * i + 1] = roots[i].im; }
}
double hm_quadratic_solver1(double a, double b, double c) { return quadratic_solver1(a, b, c); }
double hm_quadratic_solver2(double a, double b, double c) { int it = 0; return quadratic_solver2(a, b, c, it); }
double hm_cubic_solver(double a, double b, double c, double d) { int it = 0; return cubic_solver(a, b, c, d, it); }
double hm_quartic_solver(int s, double a, double b, double c, double d, double e) { int it = 0; return quartic_solver(s, a, b, c, d, e, it); }
// vectorised comparisons against glibc (returns mismatch count); n random operand pairs supplied by the test
}
