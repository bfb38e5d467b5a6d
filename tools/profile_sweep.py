"""4096x4096 Mode-R run for ncu: warm the flow up to the dense steady state (every step K=21, S=1050),
then open the profiler window for one more step.  Use with `ncu --profile-from-start off ...`."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cfd_demo_b200.model import Model
from cfd_demo_b200.types import Grid, SimulationParams
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
warm = int(sys.argv[2]) if len(sys.argv) > 2 else 26
cudart = ctypes.CDLL("libcudart.so.12")
m = Model(Grid.uniform(n, n, 40.0, 40.0, None), SimulationParams())
for s in range(warm):
    m.update()
cudart.cudaProfilerStart()
m.update()
cudart.cudaProfilerStop()
r = m.get_residuals()
print("done", r.simulation_step, r.jacobi_calls, r.sweeps, m.last_timing())
