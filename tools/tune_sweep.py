"""A/B of the Jacobi sweep kernels and tile heights at 4096^2 in the dense steady state (fp64)."""
import os, sys, subprocess, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if len(sys.argv) > 1 and sys.argv[1] == "child":
    from cfd_demo_b200.model import Model, default_options
    from cfd_demo_b200.types import Grid, SimulationParams
    flags = int(sys.argv[2]); n = int(sys.argv[3]); steps = int(sys.argv[4])
    o = default_options(); o.flags = flags
    m = Model(Grid.uniform(n, n, 40.0, 40.0, None), SimulationParams(), options=o)
    best = 1e9
    for s in range(steps):
        m.update()
        r = m.get_residuals(); t = m.last_timing()
        if r.sweeps == 1050: best = min(best, t[1] / r.sweeps)
    N = n * n
    print(json.dumps({"flags": flags, "rows": os.environ.get("CFD_SWEEP_ROWS", "auto"), "sweep_us": best * 1e3,
                      "GBs": 3 * 8 * N / (best * 1e-3) / 1e9, "step_ms": t[0]}))
else:
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    rows_list = sys.argv[2].split(",") if len(sys.argv) > 2 else ["auto", "8", "16", "24", "32", "48", "64"]
    flag_list = [int(x) for x in sys.argv[3].split(",")] if len(sys.argv) > 3 else [0, 16]
    for flags in flag_list:
        for rows in rows_list:
            env = dict(os.environ)
            if rows != "auto": env["CFD_SWEEP_ROWS"] = rows
            out = subprocess.run([sys.executable, __file__, "child", str(flags), str(n), "30"], env=env, capture_output=True, text=True)
            print(out.stdout.strip() or out.stderr[-300:], flush=True)
