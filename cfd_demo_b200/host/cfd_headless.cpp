// cfd_headless — headless driver over the C++ mirror (host/cfd_model.hpp), i.e. over exactly the C ABI the Rust shim binds.
//   cfd_headless [steps=100] [nx ny]        the reference's default scenario (src/app.rs:33-53, src/model.rs:44-55), or a
//                                           channel of the given size, Mode R; prints the Residuals lines the reference UI
//                                           would log (src/app.rs:437-449)
//   cfd_headless cavity <n> <steps> [nu dt] the headline workload's family: lid-driven cavity n x n, MGCG, relative stopping
//                                           rule 1e-8 (BASELINE configs[1], [2]: n = 1024 nu 0.01 dt 2e-5; n = 4096 nu 1e-3 dt 1e-5)
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "cfd_model.hpp"

int main(int argc, char** argv) {
  try {
    if (argc > 3 && strcmp(argv[1], "cavity") == 0) {
      const size_t n = size_t(atoll(argv[2]));
      const int steps = atoi(argv[3]);
      cfd::Grid grid{n, n, 1.0f, 1.0f, 1.0f / float(n), 1.0f / float(n), std::nullopt};
      cfd::SimulationParams prm;
      prm.viscosity = argc > 4 ? float(atof(argv[4])) : 1.0e-3f;
      prm.dt = argc > 5 ? float(atof(argv[5])) : 1.0e-5f;
      prm.scenario = cfd::Scenario::Cavity;
      prm.pressure_solver = cfd::PressureSolver::MGCG;
      cfd_solver_consts consts;
      cfd_solver_consts_default(&consts);
      consts.cg_relative = 1;
      cfd::Model model(grid, prm, consts);
      double ms = 0;
      for (int s = 0; s < steps; ++s) {
        model.update();
        const cfd::Residuals r = model.get_residuals();
        ms += r.step_time.count() * 1e3;
        if (s < 3 || (s + 1) % 10 == 0 || s == steps - 1)
          printf("step %zu t=%.6f dt=%.3e  K=%zu iterations=%zu  rel=%.3e dt*rms(r)=%.3e dt*rms(rhs)=%.3e  u-res=%.3e\n",
                 r.simulation_step, r.simulation_time, r.dt, r.jacobi_calls, r.sweeps, r.p_rel, double(r.p), r.rhs_rms, r.u);
      }
      printf("cavity %zux%zu: %d steps, mean %.3f ms per update() call (host wall clock)\n", n, n, steps, ms / steps);
      return 0;
    }
    const int steps = argc > 1 ? atoi(argv[1]) : 100;
    cfd::Grid grid = cfd::default_grid();
    if (argc > 3) {
      grid.nx = size_t(atoll(argv[2]));
      grid.ny = size_t(atoll(argv[3]));
      grid.dx = grid.lx / float(grid.nx);
      grid.dy = grid.ly / float(grid.ny);
    }
    cfd::Model model = cfd::Model::new_(grid, cfd::SimulationParams{});
    for (int s = 0; s < steps; ++s) {
      model.update();
      const cfd::Residuals r = model.get_residuals();
      if (s < 5 || (s + 1) % 10 == 0)
        printf("step %zu t=%.5f dt=%.5f  res p=%.3e u=%.3e v=%.3e  K=%zu S=%zu  step computed in %.3f ms\n",
               r.simulation_step, r.simulation_time, r.dt, r.p, r.u, r.v, r.jacobi_calls, r.sweeps,
               r.step_time.count() * 1e3);
    }
    const cfd::SimSnapshot snap = model.get_snapshot();
    double sum = 0;
    for (float x : snap.u) sum += x;
    printf("snapshot: %zu p, %zu u, %zu v values; mean u = %.6f\n", snap.p.size(), snap.u.size(), snap.v.size(),
           sum / double(snap.u.size()));
  } catch (const std::exception& e) {
    fprintf(stderr, "%s\n", e.what());
    return 1;
  }
  return 0;
}
