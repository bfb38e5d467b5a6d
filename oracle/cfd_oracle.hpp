// cfd_oracle.hpp — CPU ORACLE (test infrastructure, NOT product code).
//
// A literal C++ restatement of the reference solver `src/model.rs` of TSultanov/cfd-demo, templated on
// the scalar type R: R=float reproduces the reference's own f32 arithmetic, R=double is the oracle for
// the shipped fp64 CUDA path.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// --impl reference legs may use it; the product (cfd_demo_b200/) never links or calls it.
//
// PARITY UNPINNED: the reference has no tests, golden vectors or fixtures for model.rs
// and no Rust toolchain exists in this environment, so this restatement could not be checked against
// outputs of the reference itself.  It is pinned only by known answers that follow from reading the
// code (tests/test_oracle_kat.py) and by an independent numpy restatement (oracle/numpy_restatement.py, both velocity
// schemes) that must agree with it bit for bit in f32 and f64.
//
// Faithfulness rules: flat row-major indexing exactly as the Rust (so the
// "next row" wrap-around reads of the last 8-lane chunk happen naturally), the same 8-lane chunk /
// scalar-tail split with the tail's own rounding, no FMA contraction (build with -ffp-contract=off),
// true divisions, the same association of every expression, max-reductions that ignore NaN like
// f32::max, `u_star`/`v_star`/`p_prime` carried across calls and steps.
//
// Every function cites the reference lines it follows (paths relative to the reference repo).
#pragma once

#include <algorithm>
#include <cassert>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <functional>
#include <limits>
#include <utility>
#include <vector>

#include "../include/cfd_b200.h"

namespace cfd_oracle {

constexpr size_t LANES = 8;  // src/model.rs:11

#ifdef CFD_ORACLE_BOUNDS_CHECK
#define CFDO_AT(vec, idx) (vec).at(idx)
#else
#define CFDO_AT(vec, idx) (vec)[idx]
#endif

// Hooks for a strip-decomposed (multi-rank) run of the same algorithm; null for a single domain.
// Rows are rows of the named field; `below`/`above` are the halo depths to refresh.
template <class R>
struct StripHooks {
  // refresh halo rows of `field` (row length `row_len`, `nrows` rows in total) around owned rows [ja, jb)
  std::function<void(std::vector<R>& field, size_t row_len, size_t nrows, int below, int above)> exchange;
  std::function<R(R)> allreduce_max;
  std::function<R(R)> allreduce_sum;
  // every rank contributes rows [lo, hi) of `field` (nrows rows of row_len) and receives everybody else's
  std::function<void(std::vector<R>& field, size_t row_len, size_t nrows, size_t lo, size_t hi)> gather_rows;
};

template <class R>
class Model {
 public:
  // ---- Grid (src/model.rs:121-139); geometry kept in f32 as given, promoted copies for the solver ----
  size_t nx = 0, ny = 0;
  float f_lx = 0, f_ly = 0, f_dx = 0, f_dy = 0, f_cx = 0, f_cy = 0, f_radius = 0;
  bool has_obstacle = false;
  R lx = 0, ly = 0, dx = 0, dy = 0;

  // ---- Model fields (src/model.rs:166-214) ----
  R dt = 0, nu = 0;
  size_t substep_count = 1, simulation_step = 0, ramp_up_steps = 100;
  R current_inlet_velocity = 0, target_inlet_velocity = 0;
  int velocity_scheme = 0, pressure_solver = 0, inlet_profile = 0, scenario = 0;
  std::vector<R> u, v, p;
  std::vector<uint8_t> obstacle_mask_u, obstacle_mask_v;
  std::vector<std::pair<size_t, size_t>> obstacle_coords;
  std::vector<R> u_old, v_old, u_star, v_star, rhs, p_prime, p_prime_new;
  R last_pressure_residual = 0, last_u_residual = 0, last_v_residual = 0, simulation_time = 0;
  size_t last_piso_substeps_count = 0;

  // ---- additions: constants as data, counters, strip support ----
  cfd_solver_consts consts;
  uint64_t last_jacobi_calls = 0, last_sweeps = 0;  // K and S of the last update()
  int piso_solve_index = 0;                         // pressure solves so far in the current piso_step
  // Mode C bookkeeping of the last step's first solve: ||rhs||^2 over the unknowns (the reference of the relative
  // stopping rule, cg_relative), final ||r|| / ||rhs||, dt * rms(rhs), iterations
  R mg_bb = 0, last_p_rel = 0, last_rhs_rms = 0;
  uint64_t last_first_solve_iterations = 0;
  uint64_t total_sweeps = 0;
  size_t ja = 0, jb = 0;  // owned pressure rows [ja, jb); whole grid by default
  bool owns_top = true;   // owns v row ny
  StripHooks<R> hooks;
  // CG work vectors (extension)
  std::vector<R> cg_r, cg_d, cg_q;
  // multigrid-preconditioned CG work space (extension), see mgcg_pressure
  struct MgLevel {
    size_t mx = 0, my = 0;            // unknowns per direction
    std::vector<R> wx, hy;            // cell widths / heights in finest-cell units
    std::vector<R> WE, WW, CYW;       // per column: east / west link factors, wx / dy^2
    std::vector<R> WN, WS, CXH;       // per row: north / south link factors, hy / dx^2
    std::vector<R> e, rho, tmp;       // (mx + 2) x (my + 2) with a ring of zeros (levels >= 1)
  };
  std::vector<MgLevel> mg_levels;
  std::vector<R> mg_rho, mg_d, mg_w, mg_z, mg_z2;
  std::vector<R> mg_guess;  // start vector of the next step's first solve (carried state)
  std::vector<R> mg_last;   // p' the last first-solve ended with (carried state, mg_warm_start 2)
  std::vector<R> mg_last2;  // the one before that (carried state, mg_warm_start 3)

  // Model::new, src/model.rs:219-299
  Model(const cfd_grid& g, const cfd_params& prm, const cfd_solver_consts* c = nullptr) {
    nx = g.nx;
    ny = g.ny;
    f_lx = g.lx; f_ly = g.ly; f_dx = g.dx; f_dy = g.dy;
    lx = R(g.lx); ly = R(g.ly); dx = R(g.dx); dy = R(g.dy);
    has_obstacle = g.has_obstacle != 0;
    f_cx = g.center_x; f_cy = g.center_y; f_radius = g.radius;
    if (c) consts = *c; else cfd_solver_consts_default_inline(&consts);
    ramp_up_steps = size_t(consts.ramp_up_steps);
    ja = 0; jb = ny;

    const size_t size_u = (nx + 1) * ny, size_v = nx * (ny + 1), size_p = nx * ny;  // :223-225
    u.assign(size_u, R(0)); v.assign(size_v, R(0)); p.assign(size_p, R(0));
    obstacle_mask_u.assign(size_u, 0); obstacle_mask_v.assign(size_v, 0);
    if (has_obstacle) {  // :235-261, all in f32 like the reference (Grid and Cylinder are f32)
      for (size_t j = 0; j < ny; ++j) {
        for (size_t i = 0; i < nx; ++i) {
          const float x = (float(i) + 0.5f) * f_dx;
          const float y = (float(j) + 0.5f) * f_dy;
          const float ddx = x - f_cx;
          const float ddy = y - f_cy;
          const float distance = std::sqrt(ddx * ddx + ddy * ddy);
          if (distance < f_radius) {
            if (i > 0) obstacle_mask_u[i + j * (nx + 1)] = 1;
            if (i < nx) obstacle_mask_u[(i + 1) + j * (nx + 1)] = 1;
            if (j > 0) obstacle_mask_v[i + j * nx] = 1;
            if (j < ny) obstacle_mask_v[i + (j + 1) * nx] = 1;
            obstacle_coords.emplace_back(i, j);
          }
        }
      }
    }
    if (prm.scenario == CFD_SCENARIO_CAVITY) {
      // EXTENSION: the outermost ring of cells is solid in the reference's own mask semantics (:245-256:
      // a solid cell masks both of its u faces and both of its v faces), which puts impermeable walls on
      // the faces of the interior block; the ring is NOT added to obstacle_coords (the lid lives there).
      for (size_t j = 0; j < ny; ++j)
        for (size_t i = 0; i < nx; ++i) {
          if (!(i == 0 || i == nx - 1 || j == 0 || j == ny - 1)) continue;
          if (i > 0) obstacle_mask_u[i + j * (nx + 1)] = 1;
          obstacle_mask_u[(i + 1) + j * (nx + 1)] = 1;
          if (j > 0) obstacle_mask_v[i + j * nx] = 1;
          obstacle_mask_v[i + (j + 1) * nx] = 1;
        }
    }
    dt = R(prm.dt);                      // :265
    nu = R(prm.viscosity);               // :266
    target_inlet_velocity = R(prm.target_inlet_velocity);
    velocity_scheme = prm.velocity_scheme;
    pressure_solver = prm.pressure_solver;
    inlet_profile = prm.inlet_profile;
    scenario = prm.scenario;
    u_old = u; v_old = v; u_star = u; v_star = v;
    rhs.assign(size_p, R(0)); p_prime.assign(size_p, R(0)); p_prime_new.assign(size_p, R(0));
    mg_guess.assign(size_p, R(0));
    mg_last.assign(size_p, R(0));
    mg_last2.assign(size_p, R(0));
  }

  static void cfd_solver_consts_default_inline(cfd_solver_consts* c) {
    c->ramp_up_steps = 100;       // :269
    c->jacobi_iterations = 50;    // :737
    c->outer_rounds = 20;         // :696
    c->cg_max_iterations = 20000;
    c->jacobi_omega = 0.75;       // :735
    c->pressure_tolerance = 1e-4; // :736
    c->outer_tolerance = 1e-4;    // :721
    c->cfl = 0.2;                 // :885
    c->cg_tolerance = 1e-8;
    c->mg_omega = 0.8;
    c->mg_smoothing = 2;
    c->mg_warm_start = 3;
    c->cg_relative = 0;
    c->adaptive_substeps = 0;
  }

  // Model::set_parameters, src/model.rs:1250-1257
  void set_parameters(const cfd_params& prm) {
    nu = R(prm.viscosity);
    dt = R(prm.dt);
    target_inlet_velocity = R(prm.target_inlet_velocity);
    velocity_scheme = prm.velocity_scheme;
    pressure_solver = prm.pressure_solver;
    inlet_profile = prm.inlet_profile;
  }

  // Model::update, src/model.rs:304-379
  void update() {
    u_old = u;  // :307-308
    v_old = v;
    if (simulation_step < ramp_up_steps) {  // :311-316
      current_inlet_velocity = (R(simulation_step) / R(ramp_up_steps)) * target_inlet_velocity;
    } else {
      current_inlet_velocity = target_inlet_velocity;
    }
    const R dt_sub = dt / R(substep_count);  // :317
    last_piso_substeps_count = substep_count;
    last_jacobi_calls = 0;
    last_sweeps = 0;
    for (size_t s = 0; s < substep_count; ++s) piso_step(dt_sub);  // :322-329

    // :333-348  max |new-old| with f32::max semantics (NaN ignored)
    R max_residual_u = 0, max_residual_v = 0;
    {
      const size_t a = ja * (nx + 1), b = jb * (nx + 1);
      for (size_t k = a; k < b; ++k) {
        const R d = std::fabs(u[k] - u_old[k]);
        if (d > max_residual_u) max_residual_u = d;
      }
      const size_t va = ja * nx, vb = (owns_top ? jb + 1 : jb) * nx;
      for (size_t k = va; k < vb; ++k) {
        const R d = std::fabs(v[k] - v_old[k]);
        if (d > max_residual_v) max_residual_v = d;
      }
    }
    if (hooks.allreduce_max) {
      max_residual_u = hooks.allreduce_max(max_residual_u);
      max_residual_v = hooks.allreduce_max(max_residual_v);
    }
    last_u_residual = max_residual_u;
    last_v_residual = max_residual_v;
    simulation_step += 1;   // :350
    if (consts.adaptive_substeps) {
      // EXTENSION: the reference's own sub-step adaptation, commented out at src/model.rs:352-363,
      // made live.  f32 semantics of the Rust: `.ceil().min(20.0) as usize`, `(x / 2.0).floor() as usize`.
      const R error_norm = last_pressure_residual;
      const R tolerance = R(1e-3);
      if (error_norm > tolerance) {
        const R factor = error_norm / tolerance;
        const R grown = std::ceil(R(substep_count) * factor);
        substep_count = size_t(grown < R(20.0) ? grown : R(20.0));
      } else if (error_norm < tolerance / R(2.0) && substep_count > 1) {
        substep_count = size_t(std::floor(R(substep_count) / R(2.0)));
        if (substep_count < 1) substep_count = 1;
      }
    }
    simulation_time += dt;  // :365
    const R previous_dt = dt;  // :368-377
    const R new_dt = compute_automatic_time_step();
    const R max_increase_factor = R(1.1);
    dt = (new_dt > previous_dt) ? std::min(new_dt, previous_dt * max_increase_factor) : new_dt;
  }

  // piso_step, src/model.rs:529-730
  void piso_step(R dt_sub) {
    if (hooks.exchange) {  // predictor stencils reach 2 rows (second order), SURVEY §8e
      hooks.exchange(u, nx + 1, ny, 2, 2);
      hooks.exchange(v, nx, ny + 1, 2, 2);
    }
    piso_solve_index = 0;
    predictor_u(dt_sub);  // :538-580
    predictor_v(dt_sub);  // :586-670
    if (hooks.exchange) hooks.exchange(v_star, nx, ny + 1, 0, 1);
    recompute_divergence(dt_sub);                         // :676
    last_pressure_residual = pressure_solve(dt_sub);      // :682-687
    apply_corrector(dt_sub);                              // :693
    for (int iter = 0; iter < consts.outer_rounds; ++iter) {  // :696-724
      copy_star_from_current();                           // :698-699
      if (hooks.exchange) hooks.exchange(v_star, nx, ny + 1, 0, 1);
      recompute_divergence(dt_sub);                       // :704
      last_pressure_residual = pressure_solve(dt_sub);    // :708-713
      apply_corrector(dt_sub);                            // :718
      if (last_pressure_residual < R(consts.outer_tolerance)) break;  // :721
    }
    apply_boundary_conditions();  // :728
  }

  // u_star.copy_from_slice(&u); v_star.copy_from_slice(&v)  (src/model.rs:698-699), owned rows only
  void copy_star_from_current() {
    std::copy(u.begin() + ja * (nx + 1), u.begin() + jb * (nx + 1), u_star.begin() + ja * (nx + 1));
    const size_t vb = owns_top ? jb + 1 : jb;
    std::copy(v.begin() + ja * nx, v.begin() + vb * nx, v_star.begin() + ja * nx);
  }

  R pressure_solve(R dt_sub) {
    last_jacobi_calls += 1;
    const bool first_solve = piso_solve_index == 0;
    piso_solve_index += 1;
    if (pressure_solver == CFD_SOLVER_CG) return cg_pressure(dt_sub, first_solve);
    if (pressure_solver == CFD_SOLVER_MGCG) return mgcg_pressure(dt_sub, first_solve);
    return jacobi_pressure();
  }

  // ------------------------------------------------------------------------------------------------
  // u predictor: loop src/model.rs:538-580, compute_ustar :382-436, face helpers :893-1069
  // ------------------------------------------------------------------------------------------------
  void predictor_u(R dt_sub) {
    const size_t j_lo = std::max<size_t>(1, ja), j_hi = std::min(ny - 1, jb);
    const size_t W = nx + 1;
    // EXTENSION: QUICK takes the SecondOrder code path (scalar face helpers per lane) with its own face values
    const bool second = velocity_scheme != CFD_SCHEME_FIRST_ORDER;
    const bool quick = velocity_scheme == CFD_SCHEME_QUICK;
    for (size_t j = j_lo; j < j_hi; ++j) {
      for (size_t i = 1; i < nx; i += LANES) {  // (1..nx).step_by(LANES)
        R v_n[LANES], v_s[LANES], u_n[LANES], u_s[LANES], u_e[LANES], u_w[LANES];
        for (size_t k = 0; k < LANES; ++k) {
          const size_t c = i + k;
          v_n[k] = CFDO_AT(v, c + (j + 1) * nx);  // get_v_north :1056-1061 (flat index: wraps at c == nx)
          v_s[k] = CFDO_AT(v, c + j * nx);        // get_v_south :1064-1069
          if (!second) {
            const size_t idx = c + j * W;
            const R uc = CFDO_AT(u, idx);
            // u_face_n_first_order :966-981
            u_n[k] = (v_n[k] >= R(0)) ? uc : CFDO_AT(u, c + (j + 1) * W);
            // u_face_s_first_order :1011-1026
            u_s[k] = (v_s[k] >= R(0)) ? CFDO_AT(u, c + (j - 1) * W) : uc;
            // u_face_e_first_order :893-908
            {
              const R ul = uc, ur = CFDO_AT(u, idx + 1);
              const R avg = (ul + ur) * R(0.5);
              u_e[k] = (avg >= R(0)) ? ul : ur;
            }
            // u_face_w_first_order :929-941
            {
              const R uw = CFDO_AT(u, idx - 1);
              const R avg = (uw + uc) * R(0.5);
              u_w[k] = (avg >= R(0)) ? uw : uc;
            }
          } else if (quick) {
            u_n[k] = u_face_n_quick(c, j);
            u_s[k] = u_face_s_quick(c, j);
            u_e[k] = u_face_e_quick(c, j);
            u_w[k] = u_face_w_quick(c, j);
          } else {
            u_n[k] = u_face_n_second_order(c, j);
            u_s[k] = u_face_s_second_order(c, j);
            u_e[k] = u_face_e_second_order(c, j);
            u_w[k] = u_face_w_second_order(c, j);
          }
        }
        compute_ustar(dt_sub, i, j, v_n, v_s, u_n, u_s, u_e, u_w);
      }
    }
  }

  // compute_ustar, src/model.rs:382-436
  void compute_ustar(R dt_sub, size_t i, size_t j, const R* v_n, const R* v_s, const R* u_n, const R* u_s,
                     const R* u_e, const R* u_w) {
    const size_t W = nx + 1;
    for (size_t k = 0; k < LANES; ++k) {
      const size_t c = i + k;
      const size_t idx = c + j * W;
      const R f_e = u_e[k] * u_e[k];
      const R f_w = u_w[k] * u_w[k];
      const R f_n = v_n[k] * u_n[k];
      const R f_s = v_s[k] * u_s[k];
      const R convective = (f_e - f_w) / dx + (f_n - f_s) / dy;  // :414
      const R uc = CFDO_AT(u, idx);
      const R ue = CFDO_AT(u, idx + 1), uw = CFDO_AT(u, idx - 1);
      const R un = CFDO_AT(u, c + (j + 1) * W), us = CFDO_AT(u, c + (j - 1) * W);
      const R laplace = (ue - R(2.0) * uc + uw) / (dx * dx) + (un - R(2.0) * uc + us) / (dy * dy);  // :429-430
      R us_val = uc + dt_sub * (-convective + nu * laplace);  // :433
      if (CFDO_AT(obstacle_mask_u, idx) == 1) us_val = R(0);  // :434
      CFDO_AT(u_star, idx) = us_val;
    }
  }

  // u_face_e_second_order, src/model.rs:911-926
  R u_face_e_second_order(size_t i, size_t j) const {
    const size_t idx = i + j * (nx + 1), idx_e = (i + 1) + j * (nx + 1);
    if (CFDO_AT(u, idx) >= R(0)) {
      if (i > 1) return R(1.5) * CFDO_AT(u, idx) - R(0.5) * CFDO_AT(u, idx - 1);
      return CFDO_AT(u, idx);
    } else if ((idx_e + 1) < u.size() && i < nx - 1) {
      return R(1.5) * CFDO_AT(u, idx_e) - R(0.5) * CFDO_AT(u, idx_e + 1);
    }
    return CFDO_AT(u, idx_e);
  }
  // u_face_w_second_order, src/model.rs:944-963 (idx_ww is only formed when i > 2: the reference's
  // usize underflow at i == 1 is unobservable in release builds)
  R u_face_w_second_order(size_t i, size_t j) const {
    const size_t idx = i + j * (nx + 1), idx_w = (i - 1) + j * (nx + 1), idx_e = (i + 1) + j * (nx + 1);
    if (CFDO_AT(u, idx_w) >= R(0)) {
      if (i > 2) return R(1.5) * CFDO_AT(u, idx_w) - R(0.5) * CFDO_AT(u, (i - 2) + j * (nx + 1));
      return CFDO_AT(u, idx_w);
    }
    if (i < nx) return R(1.5) * CFDO_AT(u, idx) - R(0.5) * CFDO_AT(u, idx_e);
    return CFDO_AT(u, idx);
  }
  // get_v_north_scalar :984-989, get_v_south_scalar :1029-1034
  R get_v_north_scalar(size_t i, size_t j) const {
    const size_t idx_v_nw = (i > 0) ? (i - 1) + (j + 1) * nx : 0;
    return R(0.5) * (CFDO_AT(v, idx_v_nw) + CFDO_AT(v, i + (j + 1) * nx));
  }
  R get_v_south_scalar(size_t i, size_t j) const {
    const size_t idx_v_s = (i > 0) ? (i - 1) + j * nx : 0;
    return R(0.5) * (CFDO_AT(v, idx_v_s) + CFDO_AT(v, i + j * nx));
  }
  // u_face_n_second_order, src/model.rs:992-1008
  R u_face_n_second_order(size_t i, size_t j) const {
    const size_t W = nx + 1, idx = i + j * W, idx_n = i + (j + 1) * W;
    const R vn = get_v_north_scalar(i, j);
    if (vn >= R(0)) {
      if (j > 1) return R(1.5) * CFDO_AT(u, idx) - R(0.5) * CFDO_AT(u, i + (j - 1) * W);
      return CFDO_AT(u, idx);
    } else if ((i + (j + 2) * W) < u.size() && j < ny - 1) {
      return R(1.5) * CFDO_AT(u, idx_n) - R(0.5) * CFDO_AT(u, i + (j + 2) * W);
    }
    return CFDO_AT(u, idx_n);
  }
  // u_face_s_second_order, src/model.rs:1037-1053
  R u_face_s_second_order(size_t i, size_t j) const {
    const size_t W = nx + 1, idx = i + j * W, idx_s = i + (j - 1) * W;
    const R vs = get_v_south_scalar(i, j);
    if (vs >= R(0)) {
      if (j > 1) return R(1.5) * CFDO_AT(u, idx_s) - R(0.5) * CFDO_AT(u, i + (j - 2) * W);
      return CFDO_AT(u, idx_s);
    } else if (j < ny) {
      return R(1.5) * CFDO_AT(u, idx) - R(0.5) * CFDO_AT(u, i + (j + 1) * W);
    }
    return CFDO_AT(u, idx);
  }

  // ------------------------------------------------------------------------------------------------
  // EXTENSION (no Rust counterpart): the JS twin's QUICK face values, index.html:471-549 (u) and
  // :643-723 (v), evaluated inside the Rust predictor exactly where the SecondOrder helpers are (same loops, same
  // un-averaged flux velocities :1056-1069, same Laplacian and masks; the upwind side is picked like the Rust
  // SecondOrder helpers pick it).  JS expressions, left to right: (-a + 6 b + 3 c) / 8, (3 a + 6 b - c) / 8,
  // 1.5 a - 0.5 b.  Column nx of the u equation reads "next row" entries through the flat index like every other
  // scheme; all reads stay inside the arrays.
  // ------------------------------------------------------------------------------------------------
  static R quick3(R a, R b, R c) { return (-a + R(6) * b + R(3) * c) / R(8); }   // upstream-weighted, flow towards c
  static R quick3r(R a, R b, R c) { return (R(3) * a + R(6) * b - c) / R(8); }   // flow towards a
  R u_face_e_quick(size_t i, size_t j) const {  // index.html:473-488
    const size_t idx = i + j * (nx + 1);
    if (CFDO_AT(u, idx) >= R(0)) {
      if (i >= 2) return quick3(CFDO_AT(u, idx - 1), CFDO_AT(u, idx), CFDO_AT(u, idx + 1));
      return R(1.5) * CFDO_AT(u, idx) - R(0.5) * CFDO_AT(u, idx - 1);
    }
    if (i + 2 <= nx) return quick3r(CFDO_AT(u, idx), CFDO_AT(u, idx + 1), CFDO_AT(u, idx + 2));
    return CFDO_AT(u, idx + 1);
  }
  R u_face_w_quick(size_t i, size_t j) const {  // index.html:491-502
    const size_t idx = i + j * (nx + 1);
    if (CFDO_AT(u, idx - 1) >= R(0)) {
      if (i >= 3) return quick3(CFDO_AT(u, idx - 2), CFDO_AT(u, idx - 1), CFDO_AT(u, idx));
      return R(1.5) * CFDO_AT(u, idx - 1) - R(0.5) * CFDO_AT(u, idx);
    }
    return quick3r(CFDO_AT(u, idx - 1), CFDO_AT(u, idx), CFDO_AT(u, idx + 1));
  }
  R u_face_n_quick(size_t i, size_t j) const {  // index.html:505-523
    const size_t W = nx + 1, idx = i + j * W;
    const R u_north = CFDO_AT(u, idx + W);
    if (get_v_north_scalar(i, j) >= R(0)) {
      if (j >= 2) return quick3(CFDO_AT(u, idx - W), CFDO_AT(u, idx), u_north);
      return R(1.5) * CFDO_AT(u, idx) - R(0.5) * CFDO_AT(u, idx - W);
    }
    if (j + 2 < ny) return quick3r(CFDO_AT(u, idx), u_north, CFDO_AT(u, idx + 2 * W));
    return u_north;
  }
  R u_face_s_quick(size_t i, size_t j) const {  // index.html:526-544
    const size_t W = nx + 1, idx = i + j * W;
    const R u_south = CFDO_AT(u, idx - W);
    if (get_v_south_scalar(i, j) >= R(0)) {
      if (j >= 2) return quick3(CFDO_AT(u, idx - 2 * W), u_south, CFDO_AT(u, idx));
      return R(1.5) * u_south - R(0.5) * CFDO_AT(u, idx);
    }
    if (j + 1 < ny) return quick3r(u_south, CFDO_AT(u, idx), CFDO_AT(u, idx + W));
    return CFDO_AT(u, idx);
  }
  R v_face_e_quick(size_t i, size_t j) const {  // index.html:645-662
    const size_t idx = i + j * nx;
    if (CFDO_AT(u, (i + 1) + j * (nx + 1)) >= R(0)) {
      if (i >= 2) return quick3(CFDO_AT(v, idx - 1), CFDO_AT(v, idx), CFDO_AT(v, idx + 1));
      return R(1.5) * CFDO_AT(v, idx) - R(0.5) * CFDO_AT(v, idx - 1);
    }
    if (i + 2 < nx) return quick3r(CFDO_AT(v, idx), CFDO_AT(v, idx + 1), CFDO_AT(v, idx + 2));
    return CFDO_AT(v, idx + 1);
  }
  R v_face_w_quick(size_t i, size_t j) const {  // index.html:665-678
    const size_t idx = i + j * nx;
    if (CFDO_AT(u, i + j * (nx + 1)) >= R(0)) {
      if (i >= 3) return quick3(CFDO_AT(v, idx - 2), CFDO_AT(v, idx - 1), CFDO_AT(v, idx));
      return R(1.5) * CFDO_AT(v, idx - 1) - R(0.5) * CFDO_AT(v, idx);
    }
    return quick3r(CFDO_AT(v, idx - 1), CFDO_AT(v, idx), CFDO_AT(v, idx + 1));
  }
  R v_face_n_quick(size_t i, size_t j) const {  // index.html:681-698
    const size_t idx = i + j * nx, idx_n = idx + nx;
    const R avg = R(0.5) * (CFDO_AT(v, idx) + CFDO_AT(v, idx_n));
    if (avg >= R(0)) {
      if (j >= 2) return quick3(CFDO_AT(v, idx - nx), CFDO_AT(v, idx), CFDO_AT(v, idx_n));
      return R(1.5) * CFDO_AT(v, idx) - R(0.5) * CFDO_AT(v, idx - nx);
    }
    if (j + 1 < ny) return quick3r(CFDO_AT(v, idx), CFDO_AT(v, idx_n), CFDO_AT(v, idx + 2 * nx));
    return CFDO_AT(v, idx_n);
  }
  R v_face_s_quick(size_t i, size_t j) const {  // index.html:701-718
    const size_t idx = i + j * nx, idx_s = idx - nx;
    const R avg = R(0.5) * (CFDO_AT(v, idx_s) + CFDO_AT(v, idx));
    if (avg >= R(0)) {
      if (j >= 2) return quick3(CFDO_AT(v, idx_s - nx), CFDO_AT(v, idx_s), CFDO_AT(v, idx));
      return R(1.5) * CFDO_AT(v, idx_s) - R(0.5) * CFDO_AT(v, idx);
    }
    if (j + 1 < ny) return quick3r(CFDO_AT(v, idx_s), CFDO_AT(v, idx), CFDO_AT(v, idx + nx));
    return CFDO_AT(v, idx);
  }

  // ------------------------------------------------------------------------------------------------
  // v predictor: loop src/model.rs:586-670, compute_vstar :439-521, face helpers :1073-1248
  // ------------------------------------------------------------------------------------------------
  void predictor_v(R dt_sub) {
    const size_t j_lo = std::max<size_t>(1, ja), j_hi = std::min(ny, owns_top ? jb + 1 : jb);
    const size_t W = nx + 1;
    const bool second = velocity_scheme != CFD_SCHEME_FIRST_ORDER;
    const bool quick = velocity_scheme == CFD_SCHEME_QUICK;
    for (size_t j = j_lo; j < j_hi; ++j) {
      for (size_t i = 1; i < nx - 1; i += LANES) {  // (1..(nx-1)).step_by(LANES)
        R a_ue[LANES] = {0}, a_uw[LANES] = {0}, a_vn[LANES] = {0}, a_vs[LANES] = {0}, a_ve[LANES] = {0},
          a_vw[LANES] = {0};
        // lanes that gather inputs: first order fills nx-i lanes of the tail chunk (:599) or all 8 (:622-631);
        // second order stops at column nx-2 (:648-650), leaving zeros in the remaining lanes
        size_t lanes = LANES;
        if (!second) {
          if (i + LANES > nx - 1) lanes = nx - i;
        } else {
          if (i + LANES > nx - 1) lanes = (nx - 1) - i;
        }
        for (size_t k = 0; k < lanes; ++k) {
          const size_t c = i + k;
          const size_t idx = c + j * nx;
          a_ue[k] = CFDO_AT(u, (c + 1) + j * W);  // :600 / :622 / :651
          a_uw[k] = CFDO_AT(u, c + j * W);        // :601 / :625 / :652
          if (!second) {
            const R vc = CFDO_AT(v, idx);
            const R vnn = CFDO_AT(v, c + (j + 1) * nx), vss = CFDO_AT(v, c + (j - 1) * nx);
            // v_face_n_first_order(_scalar) :1163-1185
            a_vn[k] = (((vc + vnn) * R(0.5)) >= R(0)) ? vc : vnn;
            // v_face_s_first_order(_scalar) :1207-1229
            a_vs[k] = (((vc + vss) * R(0.5)) >= R(0)) ? vss : vc;
            // v_face_e_first_order(_scalar) :1073-1095
            a_ve[k] = (a_ue[k] >= R(0)) ? vc : CFDO_AT(v, idx + 1);
            // v_face_w_first_order(_scalar) :1116-1142
            a_vw[k] = (a_uw[k] >= R(0)) ? CFDO_AT(v, idx - 1) : vc;
          } else if (quick) {
            a_vn[k] = v_face_n_quick(c, j);
            a_vs[k] = v_face_s_quick(c, j);
            a_ve[k] = v_face_e_quick(c, j);
            a_vw[k] = v_face_w_quick(c, j);
          } else {
            a_vn[k] = v_face_n_second_order(c, j);
            a_vs[k] = v_face_s_second_order(c, j);
            a_ve[k] = v_face_e_second_order(c, j);
            a_vw[k] = v_face_w_second_order(c, j);
          }
        }
        compute_vstar(dt_sub, i, j, a_ue, a_uw, a_vn, a_vs, a_ve, a_vw);
      }
    }
  }

  // compute_vstar, src/model.rs:439-521 (the scalar tail :456-496 and the SIMD body :498-520 use the
  // same association, so one per-lane formula serves both; the tail writes nx-i lanes, the body 8)
  void compute_vstar(R dt_sub, size_t i, size_t j, const R* a_ue, const R* a_uw, const R* a_vn, const R* a_vs,
                     const R* a_ve, const R* a_vw) {
    const size_t lanes = (i + LANES > nx - 1) ? (nx - i) : LANES;
    for (size_t k = 0; k < lanes; ++k) {
      const size_t c = i + k;
      const size_t idx = c + j * nx;
      if (CFDO_AT(obstacle_mask_v, idx) == 1) {
        CFDO_AT(v_star, idx) = R(0);
        continue;
      }
      const R f_e = a_ue[k] * a_ve[k];
      const R f_w = a_uw[k] * a_vw[k];
      const R f_n = a_vn[k] * a_vn[k];
      const R f_s = a_vs[k] * a_vs[k];
      const R convective = (f_e - f_w) / dx + (f_n - f_s) / dy;
      const R vc = CFDO_AT(v, idx);
      const R ve = CFDO_AT(v, idx + 1), vw = CFDO_AT(v, idx - 1);
      const R vn = CFDO_AT(v, c + (j + 1) * nx), vs = CFDO_AT(v, c + (j - 1) * nx);
      const R laplace = (ve - R(2.0) * vc + vw) / (dx * dx) + (vn - R(2.0) * vc + vs) / (dy * dy);
      CFDO_AT(v_star, idx) = vc + dt_sub * (-convective + nu * laplace);
    }
  }

  // v_face_e_second_order, src/model.rs:1098-1113
  R v_face_e_second_order(size_t i, size_t j) const {
    const size_t idx = i + j * nx;
    const R ue = CFDO_AT(u, (i + 1) + j * (nx + 1));
    if (ue >= R(0)) {
      if (i > 0) return R(1.5) * CFDO_AT(v, idx) - R(0.5) * CFDO_AT(v, idx - 1);
      return CFDO_AT(v, idx);
    } else if ((idx + 2) < v.size() && i < nx - 2) {
      return R(1.5) * CFDO_AT(v, idx + 1) - R(0.5) * CFDO_AT(v, idx + 2);
    }
    return CFDO_AT(v, idx + 1);
  }
  // v_face_w_second_order, src/model.rs:1145-1160
  R v_face_w_second_order(size_t i, size_t j) const {
    const size_t idx = i + j * nx;
    const R uw = CFDO_AT(u, i + j * (nx + 1));
    if (uw >= R(0)) {
      if (i > 1) return R(1.5) * CFDO_AT(v, idx - 1) - R(0.5) * CFDO_AT(v, idx - 2);
      return CFDO_AT(v, idx - 1);
    } else if (i < nx - 1) {
      return R(1.5) * CFDO_AT(v, idx) - R(0.5) * CFDO_AT(v, idx + 1);
    }
    return CFDO_AT(v, idx);
  }
  // v_face_n_second_order, src/model.rs:1188-1204
  R v_face_n_second_order(size_t i, size_t j) const {
    const size_t idx = i + j * nx, idx_n = i + (j + 1) * nx;
    const R avg = R(0.5) * (CFDO_AT(v, idx) + CFDO_AT(v, idx_n));
    if (avg >= R(0)) {
      if (j > 1) return R(1.5) * CFDO_AT(v, idx) - R(0.5) * CFDO_AT(v, i + (j - 1) * nx);
      return CFDO_AT(v, idx);
    } else if ((i + (j + 2) * nx) < v.size() && j < ny - 1) {
      return R(1.5) * CFDO_AT(v, idx_n) - R(0.5) * CFDO_AT(v, i + (j + 2) * nx);
    }
    return CFDO_AT(v, idx_n);
  }
  // v_face_s_second_order, src/model.rs:1232-1248
  R v_face_s_second_order(size_t i, size_t j) const {
    const size_t idx = i + j * nx, idx_s = i + (j - 1) * nx;
    const R avg = R(0.5) * (CFDO_AT(v, idx_s) + CFDO_AT(v, idx));
    if (avg >= R(0)) {
      if (j > 1) return R(1.5) * CFDO_AT(v, idx_s) - R(0.5) * CFDO_AT(v, i + (j - 2) * nx);
      return CFDO_AT(v, idx_s);
    } else if (j < ny) {
      return R(1.5) * CFDO_AT(v, idx) - R(0.5) * CFDO_AT(v, i + (j + 1) * nx);
    }
    return CFDO_AT(v, idx);
  }

  // ------------------------------------------------------------------------------------------------
  // recompute_divergence, src/model.rs:1406-1440 (body :1427-1437 and tail :1413-1424 share one formula)
  // ------------------------------------------------------------------------------------------------
  void recompute_divergence(R dt_sub) {
    const size_t W = nx + 1;
    for (size_t j = ja; j < jb; ++j) {
      for (size_t i = 0; i < nx; ++i) {
        const R ue = CFDO_AT(u_star, (i + 1) + j * W), uw = CFDO_AT(u_star, i + j * W);
        const R vn = CFDO_AT(v_star, i + (j + 1) * nx), vs = CFDO_AT(v_star, i + j * nx);
        CFDO_AT(rhs, i + j * nx) = ((ue - uw) / dx + (vn - vs) / dy) / dt_sub;
      }
    }
  }

  // ------------------------------------------------------------------------------------------------
  // jacobi_pressure, src/model.rs:734-824
  // ------------------------------------------------------------------------------------------------
  // one sweep over owned rows: returns max |new - old| over the 8-lane body columns only (:795-798; the
  // scalar tail :757-771 never updates max_error — SURVEY N5)
  R jacobi_sweep() {
    const R omega = R(consts.jacobi_omega);
    const R one_minus = R(1.0) - omega;      // :745 / :768
    const R dx_sq = dx * dx, dy_sq = dy * dy;  // :740-742
    const R denom = R(2.0) / (dx * dx) + R(2.0) / (dy * dy);  // :746
    R max_error = 0;
    const size_t j_lo = std::max<size_t>(1, ja), j_hi = std::min(ny - 1, jb);
    const R* __restrict pp = p_prime.data();
    const R* __restrict rh = rhs.data();
    R* __restrict pn = p_prime_new.data();
    for (size_t j = j_lo; j < j_hi; ++j) {
      size_t i = 1;
      // 8-lane chunks (:774-801); written so that the compiler vectorises the lanes like std::simd does
      for (; i + LANES <= nx - 1; i += LANES) {
        const size_t stride = j * nx + i;
        R err[LANES];
#pragma omp simd
        for (size_t k = 0; k < LANES; ++k) {
          const size_t idx = stride + k;
          const R center = pp[idx];
          const R horizontal = (pp[idx + 1] + pp[idx - 1]) / dx_sq;
          const R vertical = (pp[idx + nx] + pp[idx - nx]) / dy_sq;
          const R p_update = (horizontal + vertical - rh[idx]) / denom;
          const R new_val = omega * p_update + one_minus * center;
          err[k] = std::fabs(new_val - center);
          pn[idx] = new_val;
        }
        for (size_t k = 0; k < LANES; ++k)  // reduce_max + `if error > max_error` (:795-798); NaN never wins
          if (err[k] > max_error) max_error = err[k];
      }
      // scalar tail (:755-771): the remaining nx - i columns, no contribution to max_error
      for (size_t k = 0; k < nx - i; ++k) {
        const size_t idx = j * nx + i + k;
        const R right = CFDO_AT(p_prime, idx + 1), left = CFDO_AT(p_prime, idx - 1);
        const R top = CFDO_AT(p_prime, idx + nx), bot = CFDO_AT(p_prime, idx - nx);
        const R center = CFDO_AT(p_prime, idx);
        const R horizontal = (right + left) / dx_sq;
        const R vertical = (top + bot) / dy_sq;
        const R p_update = (horizontal + vertical - CFDO_AT(rhs, idx)) / denom;
        CFDO_AT(p_prime_new, idx) = omega * p_update + one_minus * center;
      }
    }
    return max_error;
  }

  // swap + boundary values of p', src/model.rs:805-815 (rows first, then columns).  Cavity extension:
  // zero-gradient on the right wall instead of the outlet's p' = 0.
  void jacobi_swap_and_bc() {
    std::swap(p_prime, p_prime_new);
    for (size_t i = 0; i < nx; ++i) {
      if (ja == 0) p_prime[i] = p_prime[i + nx];
      if (jb == ny) p_prime[i + (ny - 1) * nx] = p_prime[i + (ny - 2) * nx];
    }
    for (size_t j = ja; j < jb; ++j) {
      p_prime[j * nx] = p_prime[1 + j * nx];
      if (scenario == CFD_SCENARIO_CAVITY) p_prime[(nx - 1) + j * nx] = p_prime[(nx - 2) + j * nx];
      else p_prime[(nx - 1) + j * nx] = R(0);
    }
    if (hooks.exchange) hooks.exchange(p_prime, nx, ny, 1, 1);
  }

  R jacobi_pressure() {
    R max_error = 0;
    for (int iter = 0; iter < consts.jacobi_iterations; ++iter) {
      max_error = jacobi_sweep();
      if (hooks.allreduce_max) max_error = hooks.allreduce_max(max_error);
      jacobi_swap_and_bc();
      last_sweeps += 1;
      total_sweeps += 1;
      if (max_error < R(consts.pressure_tolerance)) break;  // :816-819
    }
    last_pressure_residual = max_error;  // :822
    return max_error;
  }

  // ------------------------------------------------------------------------------------------------
  // EXTENSION (no reference counterpart; "Mode C" in DESIGN.md): conjugate gradients on the discrete
  // problem the Jacobi iteration relaxes.  Unknowns: p' on rows 1..ny-2, columns 1..nx-2.  Boundary
  // cells follow the Jacobi boundary rules (jacobi_swap_and_bc): mirrored on the left / bottom / top,
  // zero on the channel outlet column (cavity: mirrored too).  Eliminating them leaves the symmetric
  // positive (semi-)definite operator
  //   (A x)[i,j] = ((x - xE) + (x - xW))/dx^2 + ((x - xN) + (x - xS))/dy^2
  // and the system A x = -rhs.  Cold start x = 0 (p' is a correction), stop when the velocity
  // divergence the correction leaves behind, dt * ||b - A x||_2 / sqrt(#unknowns), is <= cg_tolerance.
  // Returns that quantity (it is what the outer loop compares with outer_tolerance).
  // Dot products are summed row by row, then over rows (the CUDA path sums in another order: Mode C
  // parity is to a tolerance).
  // ------------------------------------------------------------------------------------------------
  // boundary cells of a CG vector from its interior (same rules as the Jacobi boundary update)
  void cg_fill_boundary(std::vector<R>& x) {
    if (hooks.exchange) hooks.exchange(x, nx, ny, 1, 1);
    for (size_t i = 0; i < nx; ++i) {
      if (ja == 0) x[i] = x[i + nx];
      if (jb == ny) x[i + (ny - 1) * nx] = x[i + (ny - 2) * nx];
    }
    for (size_t j = ja; j < jb; ++j) {
      x[j * nx] = x[1 + j * nx];
      if (scenario == CFD_SCENARIO_CAVITY) x[(nx - 1) + j * nx] = x[(nx - 2) + j * nx];
      else x[(nx - 1) + j * nx] = R(0);
    }
    if (hooks.exchange) hooks.exchange(x, nx, ny, 1, 1);
  }

  // q = A d over owned interior rows, returns d.q (local part)
  R cg_apply(const std::vector<R>& d, std::vector<R>& q) {
    const R dx_sq = dx * dx, dy_sq = dy * dy;
    R acc = 0;
    const size_t j_lo = std::max<size_t>(1, ja), j_hi = std::min(ny - 1, jb);
    for (size_t j = j_lo; j < j_hi; ++j) {
      R row = 0;
      for (size_t i = 1; i < nx - 1; ++i) {
        const size_t idx = i + j * nx;
        const R c = d[idx];
        const R ax = ((c - d[idx + 1]) + (c - d[idx - 1])) / dx_sq + ((c - d[idx + nx]) + (c - d[idx - nx])) / dy_sq;
        q[idx] = ax;
        row += c * ax;
      }
      acc += row;
    }
    return acc;
  }

  // Mode C stopping rule (cg_tolerance): dt * rms(r) <= tol, or with cg_relative ||r||_2 <= tol * ||rhs||_2 where
  // rhs is the right-hand side of the piso_step's FIRST solve (bb = its sum of squares over the unknowns)
  bool mode_c_converged(R rr, R dt_sub, R n_unknowns) const {
    const R tol = R(consts.cg_tolerance);
    if (consts.cg_relative) return mode_c_relative(rr) <= tol;
    return dt_sub * std::sqrt(rr / n_unknowns) <= tol;
  }
  R mode_c_relative(R rr) const {
    if (mg_bb > R(0)) return std::sqrt(rr / mg_bb);
    return rr > R(0) ? std::numeric_limits<R>::infinity() : R(0);
  }

  R cg_pressure(R dt_sub, bool first_solve) {
    const size_t n = nx * ny;
    if (cg_r.size() != n) { cg_r.assign(n, R(0)); cg_d.assign(n, R(0)); cg_q.assign(n, R(0)); }
    const size_t j_lo = std::max<size_t>(1, ja), j_hi = std::min(ny - 1, jb);
    auto reduce = [&](R x) { return hooks.allreduce_sum ? hooks.allreduce_sum(x) : x; };
    const R n_unknowns = R((nx - 2) * (ny - 2));
    // x = 0, r = d = b = -rhs
    R rr = 0;
    for (size_t j = j_lo; j < j_hi; ++j) {
      R row = 0;
      for (size_t i = 1; i < nx - 1; ++i) {
        const size_t idx = i + j * nx;
        const R b = -rhs[idx];
        p_prime[idx] = R(0);
        cg_r[idx] = b;
        cg_d[idx] = b;
        row += b * b;
      }
      rr += row;
    }
    rr = reduce(rr);
    auto measure = [&](R rr_) { return dt_sub * std::sqrt(rr_ / n_unknowns); };
    if (first_solve) { mg_bb = rr; last_rhs_rms = measure(rr); }  // cold start: r = -rhs
    int it = 0;
    while (it < consts.cg_max_iterations && !mode_c_converged(rr, dt_sub, n_unknowns)) {
      cg_fill_boundary(cg_d);
      const R dq = reduce(cg_apply(cg_d, cg_q));
      const R alpha = rr / dq;
      R rr_new = 0;
      for (size_t j = j_lo; j < j_hi; ++j) {
        R row = 0;
        for (size_t i = 1; i < nx - 1; ++i) {
          const size_t idx = i + j * nx;
          p_prime[idx] = p_prime[idx] + alpha * cg_d[idx];
          const R r = cg_r[idx] - alpha * cg_q[idx];
          cg_r[idx] = r;
          row += r * r;
        }
        rr_new += row;
      }
      rr_new = reduce(rr_new);
      const R beta = rr_new / rr;
      for (size_t j = j_lo; j < j_hi; ++j)
        for (size_t i = 1; i < nx - 1; ++i) {
          const size_t idx = i + j * nx;
          cg_d[idx] = cg_r[idx] + beta * cg_d[idx];
        }
      rr = rr_new;
      ++it;
      last_sweeps += 1;
      total_sweeps += 1;
    }
    cg_fill_boundary(p_prime);
    if (first_solve) { last_p_rel = mode_c_relative(rr); last_first_solve_iterations = uint64_t(it); }
    const R res = measure(rr);
    last_pressure_residual = res;
    return res;
  }

  // ------------------------------------------------------------------------------------------------
  // EXTENSION (no reference counterpart; "Mode C" in DESIGN.md): conjugate gradients preconditioned by one
  // geometric-multigrid V-cycle, on the same discrete problem as cg_pressure, in the reference's own sign
  // convention: L x = rhs with (L x)[i,j] = ((xE - x) + (xW - x))/dx^2 + ((xN - x) + (xS - x))/dy^2 under the
  // Jacobi boundary rules (mirror left / bottom / top, zero outlet column; cavity: mirror there too).
  //  * Level 0 = the grid itself; its smoother is the reference's damped-Jacobi sweep (jacobi_sweep's
  //    formula + jacobi_swap_and_bc's boundary update) with damping mg_omega.
  //  * Level l+1 pairs the cells of level l per direction (a trailing single cell stays single when the
  //    count is odd), down to 1 x 1.  Coarse operators are finite-volume discretisations on that
  //    (non-uniform) tensor grid: a link between neighbours weighs (shared face) / (centre distance), in
  //    finest-cell units; the outlet's zero sits half a finest cell beyond the last column.
  //  * Transfer: residuals are summed over the (up to four) children, corrections are copied to them.
  //  * V(n,n) with n = mg_smoothing damped-Jacobi sweeps before and after; the first sweep starts from zero.
  // Stopping rule and return value as cg_pressure.  Start: the FIRST solve of a timestep starts from mg_guess —
  // the p' the first solve of the previous timestep ended with (mg_warm_start 1; the reference's Jacobi never
  // resets p' either, src/model.rs:734-824 — p' is the full pressure of this non-incremental projection and varies
  // slowly in time), or its linear extrapolation in time 2 p'_n - p'_(n-1) (mg_warm_start 2; the JS twin's
  // "extrapolated initial guess", index.html:262-270), or the quadratic one 3 p'_n - 3 p'_(n-1) + p'_(n-2)
  // (mg_warm_start 3, default).  The re-correction solves of the outer loop
  // (:696-724), whose solution is ~0, start from 0.
  // ------------------------------------------------------------------------------------------------
  void mg_build_levels() {
    mg_levels.clear();
    const R dx_sq = dx * dx, dy_sq = dy * dy;
    std::vector<R> wx(nx - 2, R(1)), hy(ny - 2, R(1));
    for (;;) {
      MgLevel L;
      L.mx = wx.size(); L.my = hy.size();
      L.wx = wx; L.hy = hy;
      L.WE.assign(L.mx, R(0)); L.WW.assign(L.mx, R(0)); L.CYW.assign(L.mx, R(0));
      L.WN.assign(L.my, R(0)); L.WS.assign(L.my, R(0)); L.CXH.assign(L.my, R(0));
      for (size_t i = 0; i < L.mx; ++i) {
        if (i + 1 < L.mx) L.WE[i] = R(1) / (R(0.5) * (wx[i] + wx[i + 1]));
        else if (scenario != CFD_SCENARIO_CAVITY) L.WE[i] = R(1) / (R(0.5) * wx[i] + R(0.5));
        if (i > 0) L.WW[i] = R(1) / (R(0.5) * (wx[i - 1] + wx[i]));
        L.CYW[i] = wx[i] / dy_sq;
      }
      for (size_t j = 0; j < L.my; ++j) {
        if (j + 1 < L.my) L.WN[j] = R(1) / (R(0.5) * (hy[j] + hy[j + 1]));
        if (j > 0) L.WS[j] = R(1) / (R(0.5) * (hy[j - 1] + hy[j]));
        L.CXH[j] = hy[j] / dx_sq;
      }
      if (!mg_levels.empty()) {
        const size_t n = (L.mx + 2) * (L.my + 2);
        L.e.assign(n, R(0)); L.rho.assign(n, R(0)); L.tmp.assign(n, R(0));
      }
      mg_levels.push_back(std::move(L));
      if (wx.size() == 1 && hy.size() == 1) break;
      auto pair_up = [](const std::vector<R>& w) {
        std::vector<R> o((w.size() + 1) / 2, R(0));
        for (size_t k = 0; k < o.size(); ++k) o[k] = w[2 * k] + (2 * k + 1 < w.size() ? w[2 * k + 1] : R(0));
        return o;
      };
      wx = pair_up(wx);
      hy = pair_up(hy);
    }
  }

  // one damped-Jacobi sweep of level 0: jacobi_sweep's formula on (in, rh) -> out over the owned rows, then the
  // boundary update (and, on strips, the neighbours' new edge rows)
  void mg_fine_sweep(const std::vector<R>& in, const std::vector<R>& rh, std::vector<R>& out) {
    const R omega = R(consts.mg_omega);
    const R one_minus = R(1.0) - omega;
    const R dx_sq = dx * dx, dy_sq = dy * dy;
    const R denom = R(2.0) / (dx * dx) + R(2.0) / (dy * dy);
    const size_t j_lo = std::max<size_t>(1, ja), j_hi = std::min(ny - 1, jb);
    for (size_t j = j_lo; j < j_hi; ++j)
      for (size_t i = 1; i < nx; ++i) {
        const size_t idx = i + j * nx;
        const R center = in[idx];
        const R horizontal = (in[idx + 1] + in[idx - 1]) / dx_sq;
        const R vertical = (in[idx + nx] + in[idx - nx]) / dy_sq;
        const R p_update = (horizontal + vertical - rh[idx]) / denom;
        out[idx] = omega * p_update + one_minus * center;
      }
    mg_fill_ring(out);
  }
  void mg_fill_ring(std::vector<R>& x) {  // jacobi_swap_and_bc's boundary update (rows, then columns)
    for (size_t i = 0; i < nx; ++i) {
      if (ja == 0) x[i] = x[i + nx];
      if (jb == ny) x[i + (ny - 1) * nx] = x[i + (ny - 2) * nx];
    }
    for (size_t j = ja; j < jb; ++j) {
      x[j * nx] = x[1 + j * nx];
      x[(nx - 1) + j * nx] = scenario == CFD_SCENARIO_CAVITY ? x[(nx - 2) + j * nx] : R(0);
    }
    if (hooks.exchange) hooks.exchange(x, nx, ny, 1, 1);
  }
  // (L x)[i,j] on an unknown, neighbours outside the unknowns replaced by the boundary rules
  R mg_fine_apply(const std::vector<R>& x, size_t i, size_t j) const {
    const R dx_sq = dx * dx, dy_sq = dy * dy;
    const size_t idx = i + j * nx;
    const R c = x[idx];
    const R xe = (i == nx - 2) ? (scenario == CFD_SCENARIO_CAVITY ? c : R(0)) : x[idx + 1];
    const R xw = (i == 1) ? c : x[idx - 1];
    const R xn = (j == ny - 2) ? c : x[idx + nx];
    const R xs = (j == 1) ? c : x[idx - nx];
    return ((xe - c) + (xw - c)) / dx_sq + ((xn - c) + (xs - c)) / dy_sq;
  }

  static R mg_coarse_apply(const MgLevel& L, const std::vector<R>& e, size_t I, size_t J) {
    const size_t W = L.mx + 2, idx = (I + 1) + (J + 1) * W;
    const R c = e[idx];
    return L.CXH[J] * (L.WE[I] * (e[idx + 1] - c) + L.WW[I] * (e[idx - 1] - c)) +
           L.CYW[I] * (L.WN[J] * (e[idx + W] - c) + L.WS[J] * (e[idx - W] - c));
  }
  static void mg_coarse_sweep(const MgLevel& L, const std::vector<R>& in, const std::vector<R>& rho, std::vector<R>& out,
                              R omega) {
    const size_t W = L.mx + 2;
    for (size_t J = 0; J < L.my; ++J)
      for (size_t I = 0; I < L.mx; ++I) {
        const size_t idx = (I + 1) + (J + 1) * W;
        const R diag = L.CXH[J] * (L.WE[I] + L.WW[I]) + L.CYW[I] * (L.WN[J] + L.WS[J]);
        const R c = in[idx];
        out[idx] = diag > R(0) ? c + omega * ((mg_coarse_apply(L, in, I, J) - rho[idx]) / diag) : R(0);
      }
  }

  // e <- approximately L_l^-1 rho on level l >= 1 (result in mg_levels[l].e)
  void mg_coarse_vcycle(size_t l) {
    MgLevel& L = mg_levels[l];
    const R omega = R(consts.mg_omega);
    const int nu_s = consts.mg_smoothing < 1 ? 1 : consts.mg_smoothing;
    std::fill(L.e.begin(), L.e.end(), R(0));
    if (L.mx == 1 && L.my == 1) {  // exact: e = -rho / diag (0 for the singular all-Neumann cavity)
      mg_coarse_sweep(L, L.e, L.rho, L.tmp, R(1));
      std::swap(L.e, L.tmp);
      return;
    }
    for (int s = 0; s < nu_s; ++s) { mg_coarse_sweep(L, L.e, L.rho, L.tmp, omega); std::swap(L.e, L.tmp); }
    MgLevel& C = mg_levels[l + 1];
    const size_t WC = C.mx + 2, W = L.mx + 2;
    for (size_t J = 0; J < C.my; ++J)
      for (size_t I = 0; I < C.mx; ++I) {
        R acc = R(0);
        for (size_t b = 0; b < 2; ++b)
          for (size_t a = 0; a < 2; ++a) {
            const size_t i = 2 * I + a, j = 2 * J + b;
            if (i < L.mx && j < L.my) acc += L.rho[(i + 1) + (j + 1) * W] - mg_coarse_apply(L, L.e, i, j);
          }
        C.rho[(I + 1) + (J + 1) * WC] = acc;
      }
    mg_coarse_vcycle(l + 1);
    for (size_t j = 0; j < L.my; ++j)
      for (size_t i = 0; i < L.mx; ++i) L.e[(i + 1) + (j + 1) * W] += C.e[(i / 2 + 1) + (j / 2 + 1) * WC];
    for (int s = 0; s < nu_s; ++s) { mg_coarse_sweep(L, L.e, L.rho, L.tmp, omega); std::swap(L.e, L.tmp); }
  }

  // mg_z <- V-cycle applied to mg_rho.  Strips (hooks set): level 0 is computed on the owned rows with one halo row
  // refreshed after every sweep; the residual of level 1 is gathered on every rank and the coarse levels run
  // replicated (the CUDA path keeps three more levels in strips — same per-cell arithmetic either way).  The strip
  // boundaries must pair up the unknown rows (the library's aligned partition, cfd_strip_rows).
  void mg_precondition() {
    const int nu_s = consts.mg_smoothing < 1 ? 1 : consts.mg_smoothing;
    const size_t j_lo = std::max<size_t>(1, ja), j_hi = std::min(ny - 1, jb);
    std::fill(mg_z.begin(), mg_z.end(), R(0));
    for (int s = 0; s < nu_s; ++s) { mg_fine_sweep(mg_z, mg_rho, mg_z2); std::swap(mg_z, mg_z2); }
    if (mg_levels.size() > 1) {
      MgLevel& C = mg_levels[1];
      const size_t WC = C.mx + 2;
      const size_t c_lo = (j_lo - 1) / 2, c_hi = (jb >= ny) ? C.my : (j_hi - 1) / 2;
      assert(((j_lo - 1) % 2 == 0) && (jb >= ny || (j_hi - 1) % 2 == 0) && "strip boundaries must pair up the unknown rows");
      for (size_t J = c_lo; J < c_hi; ++J)
        for (size_t I = 0; I < C.mx; ++I) {
          R acc = R(0);
          for (size_t b = 0; b < 2; ++b)
            for (size_t a = 0; a < 2; ++a) {
              const size_t i = 1 + 2 * I + a, j = 1 + 2 * J + b;
              if (i <= nx - 2 && j <= ny - 2) acc += mg_rho[i + j * nx] - mg_fine_apply(mg_z, i, j);
            }
          C.rho[(I + 1) + (J + 1) * WC] = acc;
        }
      if (hooks.gather_rows) hooks.gather_rows(C.rho, WC, C.my + 2, c_lo + 1, c_hi + 1);
      mg_coarse_vcycle(1);
      for (size_t j = j_lo; j < j_hi; ++j)
        for (size_t i = 1; i + 1 < nx; ++i) mg_z[i + j * nx] += C.e[((i - 1) / 2 + 1) + ((j - 1) / 2 + 1) * WC];
      mg_fill_ring(mg_z);
    }
    for (int s = 0; s < nu_s; ++s) { mg_fine_sweep(mg_z, mg_rho, mg_z2); std::swap(mg_z, mg_z2); }
  }

  R mgcg_pressure(R dt_sub, bool first_solve) {
    assert((!hooks.exchange || hooks.gather_rows) && "MGCG on strips needs the gather hook");
    const size_t n = nx * ny;
    if (mg_levels.empty()) mg_build_levels();
    if (mg_rho.size() != n) {
      mg_rho.assign(n, R(0)); mg_d.assign(n, R(0)); mg_w.assign(n, R(0)); mg_z.assign(n, R(0)); mg_z2.assign(n, R(0));
    }
    const size_t j_lo = std::max<size_t>(1, ja), j_hi = std::min(ny - 1, jb);  // owned rows of unknowns
    const R n_unknowns = R((nx - 2) * (ny - 2));
    auto measure = [&](R rr_) { return dt_sub * std::sqrt(rr_ / n_unknowns); };
    auto converged = [&](R rr_) { return mode_c_converged(rr_, dt_sub, n_unknowns); };
    // dot products: row sums, then over the owned rows, then over the ranks (the CUDA path sums in another order
    // -> tolerance parity)
    auto dot = [&](const std::vector<R>& a, const std::vector<R>& b) {
      R acc = 0;
      for (size_t j = j_lo; j < j_hi; ++j) {
        R row = 0;
        for (size_t i = 1; i + 1 < nx; ++i) row += a[i + j * nx] * b[i + j * nx];
        acc += row;
      }
      return hooks.allreduce_sum ? hooks.allreduce_sum(acc) : acc;
    };
    const bool warm = consts.mg_warm_start != 0 && first_solve;
    if (first_solve) { mg_bb = dot(rhs, rhs); last_rhs_rms = measure(mg_bb); }
    if (warm) p_prime = mg_guess; else std::fill(p_prime.begin(), p_prime.end(), R(0));
    if (warm && hooks.exchange) hooks.exchange(p_prime, nx, ny, 1, 1);
    std::fill(mg_rho.begin(), mg_rho.end(), R(0));
    std::fill(mg_d.begin(), mg_d.end(), R(0));
    for (size_t j = j_lo; j < j_hi; ++j)
      for (size_t i = 1; i + 1 < nx; ++i)
        mg_rho[i + j * nx] = warm ? rhs[i + j * nx] - mg_fine_apply(p_prime, i, j) : rhs[i + j * nx];
    R rr = dot(mg_rho, mg_rho);
    int it = 0;
    if (!converged(rr) && consts.cg_max_iterations > 0) {
      mg_precondition();
      R rz = dot(mg_rho, mg_z);
      for (size_t k = 0; k < n; ++k) mg_d[k] = mg_z[k] + R(0) * mg_d[k];
      for (;;) {
        if (hooks.exchange) hooks.exchange(mg_d, nx, ny, 1, 1);
        for (size_t j = j_lo; j < j_hi; ++j)
          for (size_t i = 1; i + 1 < nx; ++i) mg_w[i + j * nx] = mg_fine_apply(mg_d, i, j);
        const R dw = dot(mg_d, mg_w);
        const R alpha = rz / dw;
        for (size_t j = j_lo; j < j_hi; ++j)
          for (size_t i = 1; i + 1 < nx; ++i) {
            const size_t idx = i + j * nx;
            p_prime[idx] = p_prime[idx] + alpha * mg_d[idx];
            mg_rho[idx] = mg_rho[idx] - alpha * mg_w[idx];
          }
        rr = dot(mg_rho, mg_rho);
        ++it;
        last_sweeps += 1;
        total_sweeps += 1;
        if (converged(rr) || it >= consts.cg_max_iterations) break;
        mg_precondition();
        const R rz_new = dot(mg_rho, mg_z);
        const R beta = rz_new / rz;
        for (size_t j = j_lo; j < j_hi; ++j)
          for (size_t i = 1; i + 1 < nx; ++i) mg_d[i + j * nx] = mg_z[i + j * nx] + beta * mg_d[i + j * nx];
        rz = rz_new;
      }
    }
    cg_fill_boundary(p_prime);
    if (first_solve) {
      if (consts.mg_warm_start == 3) {  // quadratic extrapolation 3 p'_n - 3 p'_(n-1) + p'_(n-2)
        for (size_t k = 0; k < n; ++k) {
          mg_guess[k] = R(3) * p_prime[k] - R(3) * mg_last[k] + mg_last2[k];
          mg_last2[k] = mg_last[k];
          mg_last[k] = p_prime[k];
        }
      } else if (consts.mg_warm_start == 2) {
        for (size_t k = 0; k < n; ++k) {
          mg_guess[k] = R(2) * p_prime[k] - mg_last[k];
          mg_last[k] = p_prime[k];
        }
      } else {
        mg_guess = p_prime;
      }
    }
    if (first_solve) { last_p_rel = mode_c_relative(rr); last_first_solve_iterations = uint64_t(it); }
    const R res = measure(rr);
    last_pressure_residual = res;
    return res;
  }

  // ------------------------------------------------------------------------------------------------
  // apply_corrector, src/model.rs:1334-1404
  // ------------------------------------------------------------------------------------------------
  void apply_corrector(R dt_sub) {
    const size_t W = nx + 1;
    for (size_t j = ja; j < jb; ++j) {  // u, all rows (:1336)
      for (size_t i = 1; i < nx; i += LANES) {
        if (i + LANES > nx) {  // scalar tail :1338-1346: (dt*(pR-pL))/dx
          for (size_t k = 0; k < nx - i; ++k) {
            const size_t idx = i + k + j * W;
            const R p_right = CFDO_AT(p_prime, i + k + j * nx);
            const R p_left = CFDO_AT(p_prime, (i - 1) + k + j * nx);
            CFDO_AT(u, idx) = CFDO_AT(u_star, idx) - dt_sub * (p_right - p_left) / dx;
          }
          continue;
        }
        for (size_t k = 0; k < LANES; ++k) {  // SIMD body :1349-1362: dt*((pR-pL)/dx)
          const size_t idx = i + k + j * W;
          const R p_right = CFDO_AT(p_prime, i + k + j * nx);
          const R p_left = CFDO_AT(p_prime, (i - 1) + k + j * nx);
          const R correction = dt_sub * ((p_right - p_left) / dx);
          CFDO_AT(u, idx) = CFDO_AT(u_star, idx) - correction;
        }
      }
    }
    const size_t j_lo = std::max<size_t>(1, ja), j_hi = std::min(ny, owns_top ? jb + 1 : jb);
    for (size_t j = j_lo; j < j_hi; ++j) {  // v, rows 1..ny-1 (:1366)
      for (size_t i = 0; i < nx; i += LANES) {
        if (i + LANES > nx) {  // :1368-1375 (unreachable when nx % 8 == 0)
          for (size_t k = 0; k < nx - i; ++k) {
            const size_t idx = i + k + j * nx;
            const R p_top = CFDO_AT(p_prime, idx), p_bottom = CFDO_AT(p_prime, i + k + (j - 1) * nx);
            CFDO_AT(v, idx) = CFDO_AT(v_star, idx) - dt_sub * (p_top - p_bottom) / dy;
          }
          continue;
        }
        for (size_t k = 0; k < LANES; ++k) {  // :1378-1388
          const size_t idx = i + k + j * nx;
          const R p_top = CFDO_AT(p_prime, idx), p_bottom = CFDO_AT(p_prime, i + k + (j - 1) * nx);
          const R correction = dt_sub * ((p_top - p_bottom) / dy);
          CFDO_AT(v, idx) = CFDO_AT(v_star, idx) - correction;
        }
      }
    }
    for (size_t k = ja * nx; k < jb * nx; ++k) p[k] = p[k] + p_prime[k];  // :1392-1403
  }

  // ------------------------------------------------------------------------------------------------
  // apply_boundary_conditions, src/model.rs:827-875 (channel = reference).  Cavity = extension.
  // ------------------------------------------------------------------------------------------------
  void apply_boundary_conditions() {
    const size_t W = nx + 1;
    if (scenario == CFD_SCENARIO_CAVITY) {
      // EXTENSION: lid-driven box in the reference's style (wall values stored in the outermost rows /
      // columns of faces): lid speed = ramped target velocity on the top u row, everything else zero.
      for (size_t i = 0; i < W; ++i) {
        if (ja == 0) u[i] = R(0);
        if (jb == ny) u[i + (ny - 1) * W] = current_inlet_velocity;
      }
      for (size_t j = ja; j < jb; ++j) {
        u[0 + j * W] = R(0);
        u[nx + j * W] = R(0);
      }
      for (size_t i = 0; i < nx; ++i) {
        if (ja == 0) v[i] = R(0);
        if (owns_top) v[i + ny * nx] = R(0);
      }
      const size_t vb = owns_top ? jb + 1 : jb;
      for (size_t j = ja; j < vb; ++j) {
        v[0 + j * nx] = R(0);
        v[(nx - 1) + j * nx] = R(0);
      }
    } else {
      for (size_t j = ja; j < jb; ++j) {  // inlet :833-850
        const R y = (R(j) + R(0.5)) * dy;
        R inlet_val;
        if (inlet_profile == CFD_INLET_UNIFORM) {
          inlet_val = current_inlet_velocity;
        } else {
          const R center = ly / R(2.0), radius = ly / R(2.0);
          const R t = (y - center) / radius;
          const R val = current_inlet_velocity * (R(1.0) - t * t);  // powi(2)
          inlet_val = (val < R(0)) ? R(0) : val;
        }
        u[0 + j * W] = inlet_val;
      }
      for (size_t j = ja; j < jb; ++j) u[nx + j * W] = u[(nx - 1) + j * W];  // outlet :852-856
      for (size_t i = 0; i < W; ++i) {  // walls :858-861
        if (ja == 0) u[i] = R(0);
        if (jb == ny) u[i + (ny - 1) * W] = R(0);
      }
      for (size_t i = 0; i < nx; ++i) {  // :863-867
        if (ja == 0) v[i] = R(0);
        if (owns_top) v[i + ny * nx] = R(0);
      }
    }
    for (const auto& ij : obstacle_coords) {  // :869-874: west u face and south v face of solid cells
      if (ij.second < ja || ij.second >= jb) continue;
      u[ij.first + ij.second * W] = R(0);
      v[ij.first + ij.second * nx] = R(0);
    }
  }

  // compute_automatic_time_step, src/model.rs:878-889
  R compute_automatic_time_step() {
    R max_u = 0, max_v = 0;
    for (size_t k = ja * (nx + 1); k < jb * (nx + 1); ++k) {
      const R a = std::fabs(u[k]);
      if (a > max_u) max_u = a;
    }
    for (size_t k = ja * nx; k < (owns_top ? jb + 1 : jb) * nx; ++k) {
      const R a = std::fabs(v[k]);
      if (a > max_v) max_v = a;
    }
    R max_vel = (max_u > max_v) ? max_u : max_v;
    if (hooks.allreduce_max) max_vel = hooks.allreduce_max(max_vel);
    if (max_vel == R(0)) return dt;
    const R cfl = R(consts.cfl);
    const R dt_cfl = cfl * std::min(dx, dy) / max_vel;
    return std::min(dt_cfl, dt);
  }

  // Model::get_residuals, src/model.rs:1269-1280
  void get_residuals(cfd_residuals* out, double step_seconds) const {
    out->simulation_step = simulation_step;
    out->simulation_time = float(simulation_time);
    out->dt = float(dt);
    out->p = float(last_pressure_residual);
    out->u = float(last_u_residual);
    out->v = float(last_v_residual);
    out->step_seconds = step_seconds;
    out->piso_substeps = last_piso_substeps_count;
    out->jacobi_calls = last_jacobi_calls;
    out->sweeps = last_sweeps;
    out->simulation_time_f64 = double(simulation_time);
    out->dt_f64 = double(dt);
    out->p_f64 = double(last_pressure_residual);
    out->u_f64 = double(last_u_residual);
    out->v_f64 = double(last_v_residual);
    const bool mode_c = pressure_solver != CFD_SOLVER_JACOBI;
    out->p_rel_f64 = mode_c ? double(last_p_rel) : 0.0;
    out->rhs_rms_f64 = mode_c ? double(last_rhs_rms) : 0.0;
    out->first_solve_iterations = mode_c ? last_first_solve_iterations : 0;
  }
};

}  // namespace cfd_oracle
