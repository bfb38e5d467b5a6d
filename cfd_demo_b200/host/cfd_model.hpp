// cfd_model.hpp — C++ host-side mirror of the reference solver API (reference: src/model.rs) over the C ABI.
//
// Same names and meaning as the Rust: Grid (:121-131), Cylinder (:134-139), SimulationParams (:13-21, defaults
// :44-55), VelocityScheme / PressureSolver / InletProfile (:142-159), Residuals (:23-32), SimSnapshot (:36-42),
// Model::{new_, update, set_parameters, get_snapshot, get_residuals, run} (:219, :304, :1250, :1259, :1269, :1282),
// SimulationControlHandle (:65-117).  Where the reference panics this throws std::runtime_error.
// Header-only; link libcfd_b200.so.  No CPU fallback.
#pragma once

#include <atomic>
#include <chrono>
#include <condition_variable>
#include <deque>
#include <memory>
#include <mutex>
#include <optional>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

#include "../../include/cfd_b200.h"

namespace cfd {

enum class VelocityScheme { FirstOrder = CFD_SCHEME_FIRST_ORDER, SecondOrder = CFD_SCHEME_SECOND_ORDER, Quick = CFD_SCHEME_QUICK /* extension */ };
enum class PressureSolver { Jacobi = CFD_SOLVER_JACOBI, CG = CFD_SOLVER_CG /* extension */, MGCG = CFD_SOLVER_MGCG /* extension */ };
enum class InletProfile { Uniform = CFD_INLET_UNIFORM, Parabolic = CFD_INLET_PARABOLIC };
enum class Scenario { Channel = CFD_SCENARIO_CHANNEL, Cavity = CFD_SCENARIO_CAVITY /* extension */ };

struct Cylinder { float center_x, center_y, radius; };

struct Grid {
  size_t nx, ny;
  float lx, ly, dx, dy;
  std::optional<Cylinder> obstacle;
};

// default_grid(), src/app.rs:33-53
inline Grid default_grid() {
  const size_t nx = 800, ny = 264;
  const float lx = 30.0f, ly = 10.0f;
  return Grid{nx, ny, lx, ly, lx / float(nx), ly / float(ny), Cylinder{lx / 4.0f, ly / 2.0f, 0.75f}};
}

struct SimulationParams {  // Default, src/model.rs:44-55
  float dt = 0.005f, viscosity = 0.000001f, target_inlet_velocity = 1.0f;
  VelocityScheme velocity_scheme = VelocityScheme::FirstOrder;
  InletProfile inlet_profile = InletProfile::Uniform;
  PressureSolver pressure_solver = PressureSolver::Jacobi;
  Scenario scenario = Scenario::Channel;  // extension
};

struct Residuals {
  size_t simulation_step;
  float simulation_time, dt, p, u, v;
  std::chrono::duration<double> step_time;
  size_t piso_substeps;
  size_t jacobi_calls, sweeps;  // additions: K and S of the step
  double p_rel, rhs_rms;        // additions (Mode C): ||r|| / ||rhs|| and dt * rms(rhs) of the step's first solve
};

struct SimSnapshot {
  std::vector<float> p, u, v;
  float dt;
  bool paused;
};

inline void check(int rc) {
  if (rc != CFD_OK) throw std::runtime_error(std::string("cfd_b200: ") + cfd_last_error());
}

inline cfd_params to_c(const SimulationParams& p) {
  return cfd_params{p.dt, p.viscosity, p.target_inlet_velocity, int(p.velocity_scheme), int(p.inlet_profile),
                    int(p.pressure_solver), int(p.scenario)};
}

class SimulationControlHandle;

class Model {
 public:
  // Model::new(grid, &params), src/model.rs:219
  static Model new_(const Grid& grid, const SimulationParams& params) { return Model(grid, params); }
  Model(const Grid& grid, const SimulationParams& params) : grid_(grid) {
    cfd_grid g{grid.nx, grid.ny, grid.lx, grid.ly, grid.dx, grid.dy, grid.obstacle ? 1 : 0,
               grid.obstacle ? grid.obstacle->center_x : 0.0f, grid.obstacle ? grid.obstacle->center_y : 0.0f,
               grid.obstacle ? grid.obstacle->radius : 0.0f};
    const cfd_params p = to_c(params);
    cfd_model* m = nullptr;
    check(cfd_model_create(&g, &p, &m));
    h_.reset(m);
  }
  // extension: with the solver constants spelled out (cfd_solver_consts, e.g. cg_relative for the converged solvers)
  Model(const Grid& grid, const SimulationParams& params, const cfd_solver_consts& consts) : grid_(grid) {
    cfd_grid g{grid.nx, grid.ny, grid.lx, grid.ly, grid.dx, grid.dy, grid.obstacle ? 1 : 0,
               grid.obstacle ? grid.obstacle->center_x : 0.0f, grid.obstacle ? grid.obstacle->center_y : 0.0f,
               grid.obstacle ? grid.obstacle->radius : 0.0f};
    const cfd_params p = to_c(params);
    cfd_options o;
    cfd_options_default(&o);
    o.consts = consts;
    cfd_model* m = nullptr;
    check(cfd_model_create_ex(&g, &p, &o, &m));
    h_.reset(m);
  }
  Model(Model&&) = default;
  Model& operator=(Model&&) = default;

  void update() { check(cfd_model_update(h_.get())); }                       // :304
  void set_parameters(const SimulationParams& p) {                           // :1250
    const cfd_params c = to_c(p);
    check(cfd_model_set_params(h_.get(), &c));
  }
  SimSnapshot get_snapshot() const {                                         // :1259
    SimSnapshot s;
    s.p.resize(grid_.nx * grid_.ny);
    s.u.resize((grid_.nx + 1) * grid_.ny);
    s.v.resize(grid_.nx * (grid_.ny + 1));
    check(cfd_model_get_snapshot(h_.get(), s.p.data(), s.u.data(), s.v.data(), &s.dt));
    s.paused = false;
    return s;
  }
  Residuals get_residuals() const {                                          // :1269
    cfd_residuals r;
    check(cfd_model_get_residuals(h_.get(), &r));
    return Residuals{size_t(r.simulation_step), r.simulation_time, r.dt, r.p, r.u, r.v,
                     std::chrono::duration<double>(r.step_seconds), size_t(r.piso_substeps), size_t(r.jacobi_calls),
                     size_t(r.sweeps), r.p_rel_f64, r.rhs_rms_f64};
  }
  std::unique_ptr<SimulationControlHandle> run() &&;                         // :1282 (consumes the model)
  const Grid& grid() const { return grid_; }
  cfd_model* raw() const { return h_.get(); }

 private:
  struct Drop { void operator()(cfd_model* m) const { cfd_model_destroy(m); } };
  Grid grid_;
  std::unique_ptr<cfd_model, Drop> h_;
};

// SimulationControlHandle (src/model.rs:65-117) + the solver thread of Model::run (:1287-1325); three
// mutex-protected queues stand in for the three mpsc channels.
class SimulationControlHandle {
 public:
  explicit SimulationControlHandle(Model&& model) : model_(std::move(model)), thread_([this] { loop(); }) {}
  ~SimulationControlHandle() { stop(); }
  void stop() {
    push({Command::Stop, {}});
    if (thread_.joinable()) thread_.join();
  }
  std::optional<SimSnapshot> get_last_available_snapshot() {
    std::lock_guard<std::mutex> l(mu_);
    std::optional<SimSnapshot> last;
    while (!snapshots_.empty()) { last = std::move(snapshots_.front()); snapshots_.pop_front(); }
    return last;
  }
  std::vector<Residuals> get_new_log_messages() {
    std::lock_guard<std::mutex> l(mu_);
    std::vector<Residuals> out(residuals_.begin(), residuals_.end());
    residuals_.clear();
    return out;
  }
  void request_snapshot() { push({Command::GetSnapshot, {}}); }
  void set_params(const SimulationParams& p) { push({Command::SetParams, p}); }
  void pause() { push({Command::Pause, {}}); }
  void resume() { push({Command::Resume, {}}); }

 private:
  struct Command {
    enum Kind { Stop, GetSnapshot, SetParams, Pause, Resume } kind;
    SimulationParams params;
  };
  void push(Command c) { std::lock_guard<std::mutex> l(mu_); commands_.push_back(std::move(c)); }
  void loop() {
    bool paused = false;
    for (;;) {
      std::deque<Command> cmds;
      { std::lock_guard<std::mutex> l(mu_); cmds.swap(commands_); }
      bool snapshot_sent = false;
      for (auto& c : cmds) {
        switch (c.kind) {
          case Command::Stop: return;
          case Command::SetParams: model_.set_parameters(c.params); break;
          case Command::GetSnapshot:
            if (!snapshot_sent) {
              SimSnapshot s = model_.get_snapshot();
              s.paused = paused;
              std::lock_guard<std::mutex> l(mu_);
              snapshots_.push_back(std::move(s));
              snapshot_sent = true;
            }
            break;
          case Command::Pause: paused = true; break;
          case Command::Resume: paused = false; break;
        }
      }
      if (!paused) {
        model_.update();
        Residuals r = model_.get_residuals();
        std::lock_guard<std::mutex> l(mu_);
        residuals_.push_back(r);
      } else {
        std::this_thread::sleep_for(std::chrono::milliseconds(16));
      }
    }
  }
  Model model_;
  std::mutex mu_;
  std::deque<Command> commands_;
  std::deque<SimSnapshot> snapshots_;
  std::deque<Residuals> residuals_;
  std::thread thread_;
};

inline std::unique_ptr<SimulationControlHandle> Model::run() && {
  return std::make_unique<SimulationControlHandle>(std::move(*this));
}

}  // namespace cfd
