import sys, time, json
import numpy as np
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cfd_demo_b200.model import Model
from cfd_demo_b200.types import Grid, SimulationParams, default_grid
for name, g, steps in [("800x264", default_grid(), 40), ("4096x4096", Grid.uniform(4096, 4096, 40.0, 40.0, None), int(sys.argv[1]) if len(sys.argv) > 1 else 14)]:
    m = Model(g, SimulationParams())
    for s in range(steps):
        m.update()
        r = m.get_residuals(); t = m.last_timing()
        if s >= steps - 3 or s % 5 == 0:
            N = g.nx * g.ny
            sw = t[1] / max(r.sweeps, 1)
            print(name, "step", r.simulation_step, "K", r.jacobi_calls, "S", r.sweeps, "step_ms %.3f sweep_ms %.3f launches %d" % t,
                  "per-sweep us %.2f  GB/s %.0f" % (sw * 1e3, 3 * 8 * N / (sw * 1e-3) / 1e9), "wall %.4f" % r.step_time)
    m.close()
