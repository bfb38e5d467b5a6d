// cfd_oracle_capi.cpp — extern "C" face of the CPU ORACLE (test infrastructure, NOT product code).
// Lets tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs drive
// cfd_oracle::Model<float|double> through ctypes.  See cfd_oracle.hpp for the restatement itself
// ("parity unpinned": no reference golden vectors exist for this path).
#include <chrono>
#include <cstring>
#include <memory>

#include "cfd_oracle.hpp"

using cfd_oracle::Model;

namespace {
struct Handle {
  int precision = 64;
  std::unique_ptr<Model<float>> f;
  std::unique_ptr<Model<double>> d;
  double last_step_seconds = 0;
};

template <class R>
std::vector<R>* field_ptr(Model<R>& m, int field) {
  switch (field) {
    case CFD_FIELD_P: return &m.p;
    case CFD_FIELD_U: return &m.u;
    case CFD_FIELD_V: return &m.v;
    case CFD_FIELD_U_STAR: return &m.u_star;
    case CFD_FIELD_V_STAR: return &m.v_star;
    case CFD_FIELD_RHS: return &m.rhs;
    case CFD_FIELD_P_PRIME: return &m.p_prime;
    case CFD_FIELD_U_OLD: return &m.u_old;
    case CFD_FIELD_V_OLD: return &m.v_old;
    case CFD_FIELD_MG_GUESS: return &m.mg_guess;
    case CFD_FIELD_MG_LAST: return &m.mg_last;
    case CFD_FIELD_MG_LAST2: return &m.mg_last2;
    default: return nullptr;
  }
}

template <class R>
uint64_t field_len_t(Model<R>& m, int field) {
  if (field == CFD_FIELD_MASK_U) return m.obstacle_mask_u.size();
  if (field == CFD_FIELD_MASK_V) return m.obstacle_mask_v.size();
  auto* v = field_ptr(m, field);
  return v ? v->size() : 0;
}

template <class R>
int get_field_t(Model<R>& m, int field, double* out) {
  if (field == CFD_FIELD_MASK_U || field == CFD_FIELD_MASK_V) {
    auto& mk = field == CFD_FIELD_MASK_U ? m.obstacle_mask_u : m.obstacle_mask_v;
    for (size_t k = 0; k < mk.size(); ++k) out[k] = double(mk[k]);
    return 0;
  }
  auto* v = field_ptr(m, field);
  if (!v) return 1;
  for (size_t k = 0; k < v->size(); ++k) out[k] = double((*v)[k]);
  return 0;
}

template <class R>
int set_field_t(Model<R>& m, int field, const double* in) {
  auto* v = field_ptr(m, field);
  if (!v) return 1;
  for (size_t k = 0; k < v->size(); ++k) (*v)[k] = R(in[k]);
  return 0;
}

template <class R>
double stage_t(Model<R>& m, int stage) {
  const R dt_sub = m.dt / R(m.substep_count);
  switch (stage) {
    case 0: m.predictor_u(dt_sub); return 0;
    case 1: m.predictor_v(dt_sub); return 0;
    case 2: m.recompute_divergence(dt_sub); return 0;
    case 3: return double(m.pressure_solve(dt_sub));
    case 4: m.apply_corrector(dt_sub); return 0;
    case 5: m.apply_boundary_conditions(); return 0;
    case 6: m.copy_star_from_current(); return 0;
    case 7: { const R e = m.jacobi_sweep(); m.jacobi_swap_and_bc(); return double(e); }
    default: return -1;
  }
}
}  // namespace

// strip-decomposed run of the same algorithm (tests of the multi-rank design on CPU, e.g. over gloo): the model
// keeps full-size arrays but computes only rows [ja, jb); the callbacks refresh halo rows / reduce scalars.
typedef void (*cfdo_exchange_cb)(void* user, void* field, uint64_t row_len, uint64_t nrows, int below, int above,
                                 int elem_bytes);
typedef double (*cfdo_allreduce_cb)(void* user, double x, int op /* 0 max, 1 sum */);
typedef void (*cfdo_gather_cb)(void* user, void* field, uint64_t row_len, uint64_t nrows, uint64_t lo, uint64_t hi,
                               int elem_bytes);

template <class R>
static void set_gather_t(Model<R>& m, cfdo_gather_cb cb, void* user) {
  m.hooks.gather_rows = [=](std::vector<R>& f, size_t row_len, size_t nrows, size_t lo, size_t hi) {
    cb(user, f.data(), row_len, nrows, lo, hi, int(sizeof(R)));
  };
}

template <class R>
static void set_strip_t(Model<R>& m, uint64_t ja, uint64_t jb, int owns_top, cfdo_exchange_cb ex, cfdo_allreduce_cb ar,
                        void* user) {
  m.ja = ja; m.jb = jb; m.owns_top = owns_top != 0;
  m.hooks.exchange = [=](std::vector<R>& f, size_t row_len, size_t nrows, int below, int above) {
    ex(user, f.data(), row_len, nrows, below, above, int(sizeof(R)));
  };
  m.hooks.allreduce_max = [=](R x) { return R(ar(user, double(x), 0)); };
  m.hooks.allreduce_sum = [=](R x) { return R(ar(user, double(x), 1)); };
}

extern "C" {

void cfd_solver_consts_default(cfd_solver_consts* out) { Model<double>::cfd_solver_consts_default_inline(out); }

void* cfdo_create(const cfd_grid* g, const cfd_params* p, const cfd_solver_consts* c, int precision) {
  if (!g || !p) return nullptr;
  if (g->nx % 8 != 0 || g->nx < 16 || g->ny < 4) return nullptr;  // the reference panics otherwise
  auto* h = new Handle();
  h->precision = precision;
  if (precision == 32) h->f.reset(new Model<float>(*g, *p, c));
  else h->d.reset(new Model<double>(*g, *p, c));
  return h;
}

void cfdo_destroy(void* hv) { delete static_cast<Handle*>(hv); }

void cfdo_update(void* hv) {
  auto* h = static_cast<Handle*>(hv);
  const auto t0 = std::chrono::steady_clock::now();
  if (h->f) h->f->update(); else h->d->update();
  h->last_step_seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

void cfdo_set_params(void* hv, const cfd_params* p) {
  auto* h = static_cast<Handle*>(hv);
  if (h->f) h->f->set_parameters(*p); else h->d->set_parameters(*p);
}

uint64_t cfdo_field_len(void* hv, int field) {
  auto* h = static_cast<Handle*>(hv);
  return h->f ? field_len_t(*h->f, field) : field_len_t(*h->d, field);
}

int cfdo_get_field_f64(void* hv, int field, double* out) {
  auto* h = static_cast<Handle*>(hv);
  return h->f ? get_field_t(*h->f, field, out) : get_field_t(*h->d, field, out);
}

int cfdo_set_field_f64(void* hv, int field, const double* in) {
  auto* h = static_cast<Handle*>(hv);
  return h->f ? set_field_t(*h->f, field, in) : set_field_t(*h->d, field, in);
}

void cfdo_get_residuals(void* hv, cfd_residuals* out) {
  auto* h = static_cast<Handle*>(hv);
  if (h->f) h->f->get_residuals(out, h->last_step_seconds); else h->d->get_residuals(out, h->last_step_seconds);
}

uint64_t cfdo_obstacle_count(void* hv) {
  auto* h = static_cast<Handle*>(hv);
  return h->f ? h->f->obstacle_coords.size() : h->d->obstacle_coords.size();
}

double cfdo_current_inlet_velocity(void* hv) {
  auto* h = static_cast<Handle*>(hv);
  return h->f ? double(h->f->current_inlet_velocity) : h->d->current_inlet_velocity;
}

// single stages, for bisecting a parity failure (0 predictor_u, 1 predictor_v, 2 divergence,
// 3 pressure solve, 4 corrector, 5 boundary conditions, 6 star<-current copies, 7 one Jacobi sweep + BC)
double cfdo_stage(void* hv, int stage) {
  auto* h = static_cast<Handle*>(hv);
  return h->f ? stage_t(*h->f, stage) : stage_t(*h->d, stage);
}

// restart support for the tests / bench: scalars that are not fields (simulation_step, time, dt)
void cfdo_set_scalars(void* hv, uint64_t simulation_step, double simulation_time, double dt) {
  auto* h = static_cast<Handle*>(hv);
  if (h->f) {
    h->f->simulation_step = simulation_step; h->f->simulation_time = float(simulation_time); h->f->dt = float(dt);
  } else {
    h->d->simulation_step = simulation_step; h->d->simulation_time = simulation_time; h->d->dt = dt;
  }
}

void cfdo_set_strip(void* hv, uint64_t ja, uint64_t jb, int owns_top, cfdo_exchange_cb ex, cfdo_allreduce_cb ar,
                    void* user) {
  auto* h = static_cast<Handle*>(hv);
  if (h->f) set_strip_t(*h->f, ja, jb, owns_top, ex, ar, user); else set_strip_t(*h->d, ja, jb, owns_top, ex, ar, user);
}

void cfdo_set_gather(void* hv, cfdo_gather_cb cb, void* user) {
  auto* h = static_cast<Handle*>(hv);
  if (h->f) set_gather_t(*h->f, cb, user); else set_gather_t(*h->d, cb, user);
}

uint64_t cfdo_total_sweeps(void* hv) {
  auto* h = static_cast<Handle*>(hv);
  return h->f ? h->f->total_sweeps : h->d->total_sweeps;
}

}  // extern "C"
