"""Quick timing of the Mode C fast path (MGCG) on the BASELINE cavity config: python tools/mg_bench.py [n] [steps] [solver]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cfd_demo_b200.model import Model, default_options
from cfd_demo_b200.types import Grid, PressureSolver, Scenario, SimulationParams

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 8
solver = PressureSolver[sys.argv[3]] if len(sys.argv) > 3 else PressureSolver.MGCG
nu = {1024: 0.01, 4096: 1e-3}.get(n, 1e-3)
dt = 0.2 * (1.0 / n) ** 2 / nu / 1.0 * 0.85 if False else {1024: 2.0e-5, 4096: 1.0e-5}.get(n, 0.17 * (1.0 / n) ** 2 / nu)
prm = SimulationParams(dt=dt, viscosity=nu, target_inlet_velocity=1.0, scenario=Scenario.Cavity, pressure_solver=solver)
o = default_options()
if os.environ.get("MG_NU"):
    o.consts.mg_smoothing = int(os.environ["MG_NU"])
if os.environ.get("MG_OMEGA"):
    o.consts.mg_omega = float(os.environ["MG_OMEGA"])
o.consts.ramp_up_steps = int(os.environ.get("RAMP", "1"))
m = Model(Grid.uniform(n, n, 1.0, 1.0, None), prm, options=o)
for s in range(steps):
    t0 = time.perf_counter()
    m.update()
    wall = (time.perf_counter() - t0) * 1e3
    r = m.get_residuals()
    step_ms, solve_ms, launches = m.last_timing()
    print(f"step {s + 1}: K {r.jacobi_calls} iters {r.sweeps} p_res {r.f64['p']:.3e} u_res {r.f64['u']:.3e} step_ms {step_ms:.2f} "
          f"solve_ms {solve_ms:.2f} wall_ms {wall:.2f} launches {launches}", flush=True)
