import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cfd_demo_b200.model import Model
from cfd_demo_b200.types import Grid, SimulationParams, Scenario, PressureSolver
n = int(sys.argv[1]); nu = float(sys.argv[2]); dt = float(sys.argv[3]); steps = int(sys.argv[4]); solver = sys.argv[5] if len(sys.argv) > 5 else "cg"
prm = SimulationParams(dt=dt, viscosity=nu, scenario=Scenario.Cavity, pressure_solver=PressureSolver.CG if solver == "cg" else PressureSolver.Jacobi)
m = Model(Grid.uniform(n, n, 1.0, 1.0, None), prm)
for s in range(steps):
    m.update()
    r = m.get_residuals(); t = m.last_timing()
    if s < 6 or s % 10 == 9:
        print(f"n={n} {solver} step {r.simulation_step} K {r.jacobi_calls} S {r.sweeps} p_res {r.f64['p']:.3e} u_res {r.f64['u']:.3e} dt {r.f64['dt']:.2e} step_ms {t[0]:.2f} solve_ms {t[1]:.2f} per-iter us {t[1]*1e3/max(r.sweeps,1):.1f}", flush=True)
