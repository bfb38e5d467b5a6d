"""bench.py's reference arm (the CPU port, no GPU needed) prints ONE JSON line with the keys the driver reads."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run_reference(workload, extra_env=None):
    env = dict(os.environ, CFD_BENCH_REF_BUDGET_S="20", OMP_NUM_THREADS="1")
    env.update(extra_env or {})
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", workload,
                        "--steps", "2", "--warmup", "3"], capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, lines
    return json.loads(lines[0])


def test_reference_arm_mode_c_line(oracle_built):
    d = run_reference("cavity1024_modeC")
    assert d["impl"] == "reference" and d["metric"] == "cell_updates_per_s" and d["unit"] == "cell-updates/s"
    assert d["higher_is_better"] is True and d["vs_baseline"] is None and d["dtype"] == "f64" and d["data"] == "synthetic"
    assert d["value"] > 0 and abs(d["value"] - 1024 * 1024 / (d["ms_per_step"] * 1e-3)) < 1e-6 * d["value"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] == 1 and cb["value"] == d["value"] and "MGCG iterations" in cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "cell-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["config"]["workload"].startswith("lid-driven cavity") and d["cg_iterations_per_step"] > 0
    assert d["stop"]["rel_residual"] <= 1e-8 and d["stop"]["dt_rms_rhs"] > 0  # the SURVEY norm, both norms printed
    # both arms print the same `config` for the same workload (the driver compares them)
    sys.path.insert(0, ROOT)
    import bench
    w = bench.WORKLOADS["cavity1024_modeC"]
    assert d["config"] == json.loads(json.dumps(bench.workload_config(w, w["nx"], w["ny"])))


def test_every_baseline_config_has_a_workload():
    sys.path.insert(0, ROOT)
    import bench
    W = bench.WORKLOADS
    assert (W["default800_modeR"]["nx"], W["default800_modeR"]["ny"]) == (800, 264)                      # configs[0]
    assert (W["cavity1024_modeC"]["nx"], W["cavity1024_modeR"]["nx"]) == (1024, 1024)                    # configs[1]: R and C
    assert W["cavity4096_modeC"]["consts"]["cg_relative"] == 1                                           # configs[2]
    c3 = W["channel8192x2048_modeR"]                                                                     # configs[3]
    assert (c3["nx"], c3["ny"], c3["lx"], c3["ly"], c3["cylinder"], c3.get("strong")) == (8192, 2048, 40.0, 10.0, (10.0, 5.0, 0.75), True)
    assert W["cavity16384_modeC"]["nx"] == 16384 and W["cavity16384_modeC"].get("strong") is True        # configs[4]


def test_reference_arm_is_rank_zero_only(oracle_built):
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                       capture_output=True, text=True, timeout=120, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_default_workload_is_the_baseline_config():
    sys.path.insert(0, ROOT)
    import bench
    w = bench.WORKLOADS[bench.DEFAULT_WORKLOAD]
    base = json.load(open(os.path.join(ROOT, "BASELINE.json")))
    assert "4096" in base["metric"] and "cavity Re=1000 on 4096" in base["configs"][2]
    assert (w["nx"], w["ny"], w["kind"]) == (4096, 4096, "modeC") and w["params"]["scenario"] == 1
    assert abs(1.0 * 1.0 / w["params"]["viscosity"] - 1000.0) < 1e-6          # Re = U L / nu
    assert w["params"]["dt"] < (1.0 / 4096) ** 2 / (4 * w["params"]["viscosity"])  # explicit diffusion limit


def test_mode_c_byte_accounting_matches_design():
    """stage-by-stage bytes of a V(nu,nu) iteration are 19.5 + 3 (2 nu - 1) sN on level 0 and 6.5 + 3 (2 nu - 1)
    s N_l on a coarse level (28.5 / 15.5 at nu = 2, round 1's figures); as the fused legs run, 16.5 and 5.5."""
    import bench
    n = 64 * 64
    base = bench.step_bytes_mode_c(64, 64, 1, 1, 0, nu=2)
    assert base == 8 * n * (8 + 14 + 3) + 2 * n
    per_it2 = bench.step_bytes_mode_c(64, 64, 1, 1, 1, nu=2) - base
    per_it3 = bench.step_bytes_mode_c(64, 64, 1, 1, 1, nu=3) - base
    fused = bench.step_bytes_mode_c(64, 64, 1, 1, 1, nu=3, fused=True) - base
    assert abs(per_it2 - 8 * n * (28.5 + 15.5 / 3)) < 1e-6
    assert abs(per_it3 - 8 * n * (34.5 + 21.5 / 3)) < 1e-6
    assert abs(fused - 8 * n * (16.5 + 5.5 / 3)) < 1e-6
    assert fused < per_it2 < per_it3
    # single domain: the elided round's divergence rides on the corrector before it (k_corrector_div): 1 sN instead of 3
    assert base - bench.step_bytes_mode_c(64, 64, 1, 1, 0, nu=2, fused=True, corrector_div=True) == 8 * n * 2
    assert bench.step_bytes_mode_c(64, 64, 1, 1, 0, nu=2, corrector_div=True) == base  # the stage-by-stage count is unchanged
