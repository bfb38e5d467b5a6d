// cfd_headless — headless driver over the C++ mirror: runs the reference's default scenario (src/app.rs:33-53,
// src/model.rs:44-55) or a square channel of a given size for N timesteps and prints the Residuals lines the
// reference UI would log (src/app.rs:437-449).   usage: cfd_headless [steps=100] [nx ny]
#include <cstdio>
#include <cstdlib>

#include "cfd_model.hpp"

int main(int argc, char** argv) {
  const int steps = argc > 1 ? atoi(argv[1]) : 100;
  cfd::Grid grid = cfd::default_grid();
  if (argc > 3) {
    grid.nx = size_t(atoll(argv[2]));
    grid.ny = size_t(atoll(argv[3]));
    grid.dx = grid.lx / float(grid.nx);
    grid.dy = grid.ly / float(grid.ny);
  }
  try {
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
