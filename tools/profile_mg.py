"""BASELINE cavity 4096^2 Mode C (MGCG) run for ncu: spin up past the lid ramp, then open the profiler window for one
more timestep.  Use with `ncu --profile-from-start off ...`."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from cfd_demo_b200.model import Model
w = bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "cavity4096_modeC"]
warm = int(sys.argv[2]) if len(sys.argv) > 2 else 112
cudart = ctypes.CDLL("libcudart.so.12")
from cfd_demo_b200.model import default_options
opts = default_options()
opts.consts = bench.make_consts(w)  # the workload's solver constants (and CFD_BENCH_CONSTS overrides), like bench.py
m = Model(bench.make_grid(w), bench.make_params(w), options=opts)
for s in range(warm):
    m.update()
cudart.cudaProfilerStart()
m.update()
cudart.cudaProfilerStop()
r = m.get_residuals()
print("done", r.simulation_step, r.jacobi_calls, r.sweeps, m.last_timing())
