#!/usr/bin/env python
"""bench.py — throughput of the solver hot path (one `Model::update` timestep, src/model.rs:304-379).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]

Default workload = the configuration BASELINE.json's metric is quoted on (configs[2]): lid-driven cavity, Re = 1000,
4096 x 4096 cells, fp64, the pressure solve converged to dt * rms(residual) <= 1e-8 every step ("Mode C": conjugate
gradients preconditioned by a multigrid V-cycle whose fine-level smoother is the reference's damped-Jacobi sweep
kernel).  `--workload channel4096_modeR` is the reference's own algorithm (<= 50 Jacobi sweeps x <= 21 solves per
step, never converged) in its dense saturated regime; with --gpus N > 1 that one runs as row strips over NVLink.

A "step" is one timestep (predictor, K pressure solves with the corrector after each, boundary conditions,
residual / CFL reductions).  State is resident in HBM when the timed region starts.  Prints ONE JSON line
(contract in the task description): `value` = cell-updates/s (= nx*ny*timesteps/s) of the whole job, `e2e` = the
same metric through the C ABI with HOST buffers every step (set_params in; residuals + the f32 snapshot of p, u, v
out, into pinned memory), `roofline` for the Jacobi sweep kernel from live CUDA-event timing, `cpu_baseline` = the
CPU oracle (a C++ port of the reference, 1 core because the reference solver is single-threaded by construction,
src/model.rs:1287) on a bounded sample.

`--impl reference` times that CPU port alone (the Rust reference cannot be built here: no Rust toolchain).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

FALLBACK_HBM_GBS = 6650.0  # /opt/skills/guides/B200_PROFILING.md fallback when MEASURED_PEAKS.json is absent

STOP_RULE_REL = ("||r||_2 <= 1e-8 * ||rhs||_2 of the step's first solve (cfd_solver_consts::cg_relative = 1, the relative L2 "
                 "norm SURVEY 8d states for this config); re-correction solves measure against the same reference")

WORKLOADS = {
    # BASELINE.json configs[2] / SURVEY 8(d) config 3: cavity Re = 1000 at 4096^2, dt below the explicit diffusion
    # limit dx^2 / (4 nu) = 1.49e-5, pressure solve converged to a relative residual of 1e-8 (Mode C, MGCG).  Spin-up runs
    # through the reference's 100-step ramp of the driving velocity (src/model.rs:311-316) and 40 steps beyond it, where
    # every solve takes the same number of iterations (the count steps down 6 -> 3 in the 25 steps after the ramp).
    "cavity4096_modeC": dict(kind="modeC", nx=4096, ny=4096, lx=1.0, ly=1.0, cylinder=None, spinup=140,
                             params=dict(dt=1.0e-5, viscosity=1.0e-3, target_inlet_velocity=1.0, scenario=1,
                                         pressure_solver=2),
                             consts=dict(cg_relative=1, cg_tolerance=1e-8, mg_smoothing=3), stop_rule=STOP_RULE_REL,
                             desc="lid-driven cavity Re=1000, 4096x4096, fp64, pressure solve converged to a relative residual "
                                  "of 1e-8 every step (Mode C: CG preconditioned by a multigrid V(3,3)-cycle, the reference's "
                                  "damped-Jacobi sweep as fine-level smoother)"),
    # BASELINE.json configs[4]: the 16384^2 cavity; with --gpus N the SAME grid is cut into N strips (strong scaling)
    "cavity16384_modeC": dict(kind="modeC", nx=16384, ny=16384, lx=1.0, ly=1.0, cylinder=None, spinup=140, strong=True,
                              params=dict(dt=0.6e-6, viscosity=1.0e-3, target_inlet_velocity=1.0, scenario=1,
                                          pressure_solver=2),
                              consts=dict(cg_relative=1, cg_tolerance=1e-8, mg_smoothing=3), stop_rule=STOP_RULE_REL,
                              desc="lid-driven cavity Re=1000, 16384x16384, fp64, Mode C (MGCG) converged to a relative residual of "
                                   "1e-8, strong scaling (the same grid on every GPU count)"),
    # BASELINE.json configs[1]
    "cavity1024_modeC": dict(kind="modeC", nx=1024, ny=1024, lx=1.0, ly=1.0, cylinder=None, spinup=140,
                             params=dict(dt=2.0e-5, viscosity=1.0e-2, target_inlet_velocity=1.0, scenario=1,
                                         pressure_solver=2),
                             consts=dict(cg_relative=1, cg_tolerance=1e-8, mg_smoothing=3), stop_rule=STOP_RULE_REL,
                             desc="lid-driven cavity Re=100, 1024x1024, fp64, Mode C (MGCG), L2-resident regime"),
    "cavity1024_modeR": dict(kind="modeR", nx=1024, ny=1024, lx=1.0, ly=1.0, cylinder=None, spinup=32,
                             params=dict(dt=2.0e-5, viscosity=1.0e-2, target_inlet_velocity=1.0, scenario=1),
                             desc="lid-driven cavity Re=100, 1024x1024, fp64, Mode R (the reference's Jacobi + outer loop)"),
    # the reference's own algorithm; spin-up until the dense saturated regime where every step runs K=21, S=1050
    # (until p' is non-zero everywhere the sweeps still hit the slow zero-dividend path; ~21 steps at 4096x4096,
    # ~35 for the taller multi-GPU domains)
    "channel4096_modeR": dict(kind="modeR", nx=4096, ny=4096, lx=40.0, ly=40.0, cylinder=None, params={}, spinup=44,
                              desc="channel 4096x4096 (reference scenario), fp64, Mode R = the reference's damped "
                                   "Jacobi (<=50 sweeps) + <=20 outer re-corrections, dense saturated regime "
                                   "(K=21 solves, S=1050 sweeps per step)"),
    # BASELINE.json configs[3]: channel past a masked cylinder, 8192 x 2048, reference defaults; with --gpus N the SAME grid
    # is cut into N strips (strong scaling; Mode R is bit-identical for any strip count)
    "channel8192x2048_modeR": dict(kind="modeR", nx=8192, ny=2048, lx=40.0, ly=10.0, cylinder=(10.0, 5.0, 0.75), params={},
                                   spinup=60, strong=True,
                                   desc="channel past a masked cylinder, 8192x2048 (40 x 10, cylinder (10, 5, r 0.75), reference "
                                        "defaults), fp64, Mode R, saturated regime (K=21, S=1050 per step), strong scaling over "
                                        "row strips"),
    "default800_modeR": dict(kind="modeR", nx=800, ny=264, lx=30.0, ly=10.0, cylinder=(7.5, 5.0, 0.75), params={}, spinup=26,
                             desc="reference default_grid() 800x264 + cylinder, Mode R"),
}
DEFAULT_WORKLOAD = "cavity4096_modeC"


def peak_hbm():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


def make_grid(w):
    from cfd_demo_b200.types import Cylinder, Grid
    cyl = Cylinder(*w["cylinder"]) if w["cylinder"] else None
    return Grid.uniform(w["nx"], w["ny"], w["lx"], w["ly"], cyl)


def make_params(w):
    from cfd_demo_b200.types import SimulationParams
    return SimulationParams(**w["params"])


def make_consts(w):
    """The workload's solver constants (cfd_solver_consts): the reference's literals plus the workload's overrides — the
    same struct goes to the CUDA library and to the CPU oracle."""
    from cfd_demo_b200 import _abi
    from cfd_demo_b200.model import load_library
    import ctypes as C
    c = _abi.CfdSolverConsts()
    load_library().cfd_solver_consts_default(C.byref(c))
    for k, v in (w.get("consts") or {}).items():
        setattr(c, k, v)
    # A/B hook: CFD_BENCH_CONSTS="mg_smoothing=4,mg_omega=0.8" overrides solver constants (both arms read it)
    for item in filter(None, os.environ.get("CFD_BENCH_CONSTS", "").split(",")):
        k, v = item.split("=")
        setattr(c, k, type(getattr(c, k))(float(v)))
    return c


def workload_config(w, nx, ny):
    """`config` of the JSON line: identical keys and values in both arms (ours / --impl reference) for the same workload."""
    mode_c = w["kind"] == "modeC"
    return {"workload": w["desc"], "nx": nx, "ny": ny,
            "solver": "mgcg" if mode_c else "jacobi (reference)",
            "stop_rule": w.get("stop_rule", "dt * rms(r) <= 1e-8" if mode_c else
                               "reference: max|dp'| < 1e-4 or 50 sweeps; <= 20 re-corrections (src/model.rs:696-824)"),
            "l2": "every field (134 MB at 4096^2) is larger than L2 (126 MB); no flush needed",
            "timing": "ours: CUDA events on the model's stream around each update(), summed over K steps; reference: wall clock"}


class ClockSampler:
    """nvidia-smi clocks + throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, device_index=0):
        self.proc, self.lines, self.device_index = None, [], device_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.device_index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.lines:
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for name, val in zip(names, parts[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ---------------------------------------------------------------------------------------------------------
# CPU side, Mode R: the oracle port, timed on a bounded sample (one outer round = `u*<-u` copies + divergence +
# 50-sweep Jacobi solve + corrector, src/model.rs:698-718) on dense synthetic fields of the workload's size.
# A saturated timestep is 21 such rounds plus ~1 % of predictor / BC / reductions, which are timed once.
# ---------------------------------------------------------------------------------------------------------
def cpu_oracle_sample(w, precision, rounds, rounds_per_step=21):
    from cfd_demo_b200 import _abi
    from oracle.cpu_oracle import OracleModel
    grid, params = make_grid(w), make_params(w)
    nx, ny = grid.nx, grid.ny
    m = OracleModel(grid, params, precision=precision)
    # dense, smooth, non-zero synthetic state (values do not change the work per sweep; the solve runs its
    # full 50 sweeps because max|dp'| stays far above 1e-4)
    x = np.linspace(0.0, 1.0, nx, dtype=np.float64)
    y = np.linspace(0.0, 1.0, ny + 1, dtype=np.float64)
    uu = (1.0 + 0.25 * np.sin(6.0 * y[:ny, None] + 3.0 * np.linspace(0, 1, nx + 1)[None, :]))
    vv = 0.1 * np.cos(5.0 * y[:, None] + 2.0 * x[None, :])
    pp = 0.05 * np.sin(4.0 * y[:ny, None]) * np.cos(3.0 * x[None, :])
    for fid, arr in ((_abi.FIELD_U, uu), (_abi.FIELD_V, vv), (_abi.FIELD_U_STAR, uu), (_abi.FIELD_V_STAR, vv),
                     (_abi.FIELD_P_PRIME, pp), (_abi.FIELD_P, pp)):
        m.set_field(fid, arr.ravel())
    m.set_scalars(200, 1.0, float(np.float32(0.001)))
    t0 = time.perf_counter()
    m.stage(OracleModel.STAGE_PREDICTOR_U)
    m.stage(OracleModel.STAGE_PREDICTOR_V)
    m.stage(OracleModel.STAGE_BC)
    t_misc = time.perf_counter() - t0
    times = []
    sweeps0 = m.total_sweeps()
    for _ in range(rounds):
        t0 = time.perf_counter()
        m.stage(OracleModel.STAGE_COPY_STAR)
        m.stage(OracleModel.STAGE_DIVERGENCE)
        m.stage(OracleModel.STAGE_PRESSURE)
        m.stage(OracleModel.STAGE_CORRECTOR)
        times.append(time.perf_counter() - t0)
    sweeps = (m.total_sweeps() - sweeps0) / max(rounds, 1)
    t_round = statistics.median(times)
    step_seconds = rounds_per_step * t_round + t_misc
    return {"t_round": t_round, "t_misc": t_misc, "sweeps_per_round": sweeps, "step_seconds": step_seconds,
            "cells": nx * ny, "rounds": rounds, "cpu_seconds": sum(times) + t_misc}


CPU_NOTE = ("C++ port of the reference (oracle/cfd_oracle.hpp); the Rust reference cannot be built here (no Rust "
            "toolchain) and is single-threaded by construction (src/model.rs:1287)")


def run_reference_mode_r(args, w):
    from oracle import cpu_oracle
    cpu_oracle.build()
    precision = 32 if args.ref_precision == 32 else 64
    rounds = max(1, args.steps + args.warmup)
    # bound the whole run to a few minutes: one round at 4096^2 costs ~2.5-3 s on one core
    budget_rounds = max(1, int(150.0 / (3.0 * (w["nx"] * w["ny"]) / (4096.0 * 4096.0) + 1e-9)))
    rounds = min(rounds, budget_rounds)
    s = cpu_oracle_sample(w, precision, rounds)
    value = s["cells"] / s["step_seconds"]
    sample = (f"{s['rounds']} outer rounds (copies + divergence + {s['sweeps_per_round']:.0f}-sweep Jacobi solve + "
              f"corrector, src/model.rs:698-718) on dense synthetic {w['nx']}x{w['ny']} fields, median round "
              f"{s['t_round']:.3f} s, x21 rounds per saturated timestep + {s['t_misc']:.3f} s predictor/BC")
    return {
        "impl": "reference", "metric": "cell_updates_per_s", "value": value, "unit": "cell-updates/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": s["step_seconds"] * 1e3,
        "higher_is_better": True, "scaling": "strong" if w.get("strong") else "weak", "vs_baseline": None,
        "dtype": "f32" if precision == 32 else "f64", "data": "synthetic",
        "timesteps_per_s": 1.0 / s["step_seconds"],
        "config": workload_config(w, w["nx"], w["ny"]),
        "solves_per_step": 21, "sweeps_per_step": 1050,
        "cpu_baseline": {"value": value, "unit": "cell-updates/s", "cores": 1, "kind": "port", "sample": sample,
                         "host_cores": os.cpu_count(), "note": CPU_NOTE},
        "e2e": {"value": value, "unit": "cell-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }


def run_reference_mode_c(args, w):
    """Mode C on the CPU port: the same workload from rest (the first two steps are trivial: the driving velocity
    ramps up from 0, src/model.rs:311-316), W warm-up steps then up to K timed steps within a time budget."""
    from oracle import cpu_oracle
    from oracle.cpu_oracle import OracleModel
    cpu_oracle.build()
    precision = 32 if args.ref_precision == 32 else 64
    grid, params = make_grid(w), make_params(w)
    m = OracleModel(grid, params, precision=precision, consts=make_consts(w))
    budget_s = float(os.environ.get("CFD_BENCH_REF_BUDGET_S", "150"))
    t_begin = time.perf_counter()
    for _ in range(max(args.warmup, 3)):
        m.update()
    times, iters = [], []
    r = None
    for _ in range(max(1, args.steps)):
        t0 = time.perf_counter()
        m.update()
        times.append(time.perf_counter() - t0)
        r = m.get_residuals()
        iters.append(r.sweeps)
        if time.perf_counter() - t_begin + times[-1] > budget_s:
            break
    step_s = sum(times) / len(times)
    value = grid.nx * grid.ny / step_s
    sample = (f"{len(times)} timesteps of the same workload from rest after {max(args.warmup, 3)} warm-up steps "
              f"(oracle<{'float' if precision == 32 else 'double'}>, MGCG iterations per step {iters}), mean {step_s:.3f} s/step")
    return {
        "impl": "reference", "metric": "cell_updates_per_s", "value": value, "unit": "cell-updates/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": step_s * 1e3,
        "higher_is_better": True, "scaling": "strong" if w.get("strong") else "weak", "vs_baseline": None,
        "dtype": "f32" if precision == 32 else "f64", "data": "synthetic",
        "timesteps_per_s": 1.0 / step_s, "steps_timed": len(times),
        "config": workload_config(w, w["nx"], w["ny"]),
        "solves_per_step": 2, "cg_iterations_per_step": sum(iters) / len(iters), "cg_iterations_list": iters,
        "ms_per_cg_iteration": step_s * 1e3 / max(sum(iters) / len(iters), 1e-9),
        "start_state": "from rest (the GPU arm starts from a state spun up on the device; same grid, parameters, solver "
                       "constants and stopping rule)",
        "like_for_like_note": ("the work of a step is proportional to its CG iterations: this arm's steps, taken from rest inside the "
                               "100-step ramp of the lid speed, need the iteration counts listed in cg_iterations_list, the GPU "
                               "arm's spun-up steps need 2; the CPU port cannot reach the spun-up state within the time budget "
                               "(140 steps of ~5 s).  The same-state comparison is the GPU arm's own cpu_baseline (the GPU model's "
                               "state loaded into this port, which then computes the same next step with the same iteration "
                               "count); ms_per_cg_iteration here is the figure that carries over."),
        "stop": {"rel_residual": r.f64["p_rel"], "dt_rms_residual": r.f64["p"], "dt_rms_rhs": r.f64["rhs_rms"]},
        "cpu_baseline": {"value": value, "unit": "cell-updates/s", "cores": 1, "kind": "port", "sample": sample,
                         "host_cores": os.cpu_count(), "note": CPU_NOTE},
        "e2e": {"value": value, "unit": "cell-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }


def run_reference(args, w):
    """The reference arm: the CPU port of the reference (oracle), 1 core, bounded samples; rank 0 only."""
    if int(os.environ.get("RANK", "0")) != 0:
        return 0
    line = run_reference_mode_c(args, w) if w["kind"] == "modeC" else run_reference_mode_r(args, w)
    print(json.dumps(line), flush=True)
    return 0


def cpu_baseline_mode_r(w, nx, ny, budget_s=24.0):
    rounds = max(1, min(8, int(budget_s / (3.0 * w["nx"] * w["ny"] / (4096.0 * 4096.0) + 1e-9))))
    s = cpu_oracle_sample(w, 64, rounds)
    s32 = cpu_oracle_sample(w, 32, max(1, rounds // 2))
    return {"value": s["cells"] / s["step_seconds"], "unit": "cell-updates/s", "cores": 1, "kind": "port",
            "host_cores": os.cpu_count(), "value_f32": s32["cells"] / s32["step_seconds"],
            "sample": (f"oracle<double>: {s['rounds']} outer rounds (copies + divergence + "
                       f"{s['sweeps_per_round']:.0f}-sweep Jacobi solve + corrector) on dense synthetic "
                       f"{nx}x{ny} fields, median {s['t_round']:.3f} s/round, x21 rounds per saturated "
                       f"timestep + {s['t_misc']:.3f} s predictor/BC; value_f32 = same with oracle<float> "
                       f"(the reference's own precision)")}


def cpu_baseline_mode_c(w, model):
    """The GPU model's complete state after the timed region is loaded into the CPU oracle, which then computes the
    SAME next timestep (the GPU does it too; iteration counts are compared)."""
    from cfd_demo_b200 import _abi
    from oracle.cpu_oracle import OracleModel
    grid, params = make_grid(w), make_params(w)
    cpu = OracleModel(grid, params, precision=64, consts=make_consts(w))
    r0 = model.get_residuals()
    for fid in (_abi.FIELD_P, _abi.FIELD_U, _abi.FIELD_V, _abi.FIELD_U_STAR, _abi.FIELD_V_STAR, _abi.FIELD_P_PRIME,
                _abi.FIELD_MG_GUESS, _abi.FIELD_MG_LAST, _abi.FIELD_MG_LAST2):
        cpu.set_field(fid, model.field(fid))
    cpu.set_scalars(r0.simulation_step, r0.f64["simulation_time"], r0.f64["dt"])
    t0 = time.perf_counter()
    cpu.update()
    dt_cpu = time.perf_counter() - t0
    model.update()
    rc, rg = cpu.get_residuals(), model.get_residuals()
    du = float(np.linalg.norm(cpu.field(_abi.FIELD_U) - model.field(_abi.FIELD_U)) /
               max(np.linalg.norm(cpu.field(_abi.FIELD_U)), 1e-300))
    return {"value": grid.nx * grid.ny / dt_cpu, "unit": "cell-updates/s", "cores": 1, "kind": "port",
            "host_cores": os.cpu_count(),
            "sample": (f"oracle<double>: ONE timestep (step {rc.simulation_step}) from the GPU model's state, {dt_cpu:.2f} s, "
                       f"{rc.sweeps} MGCG iterations (GPU: {rg.sweeps}); relative L2 difference of u against the GPU's "
                       f"same step: {du:.2e}; ||r||/||rhs|| oracle {rc.f64['p_rel']:.2e}, GPU {rg.f64['p_rel']:.2e}")}


# ---------------------------------------------------------------------------------------------------------
# GPU side
# ---------------------------------------------------------------------------------------------------------
class Ranks:
    """torchrun environment of this process (one process per GPU)."""

    def __init__(self):
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        self.progress = time.time()

    def tick(self):
        self.progress = time.time()

    def barrier(self):
        import torch
        import torch.distributed as dist
        self.tick()
        torch.cuda.synchronize()
        if self.world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def max_over_ranks(self, values):
        import torch
        import torch.distributed as dist
        t = torch.tensor(values, dtype=torch.float64, device="cuda")
        if self.world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return [float(x) for x in t.tolist()]


def build_model(w, rk, strips, flags=0):
    """The workload's model on this rank: single domain, or this rank's strip of the (weak: N times taller) domain."""
    import torch.distributed as dist
    from cfd_demo_b200.model import Model, default_options, nccl_unique_id
    from cfd_demo_b200.types import Cylinder, Grid
    cyl = Cylinder(*w["cylinder"]) if w["cylinder"] else None
    strong = bool(w.get("strong")) and strips
    weak = strips and not strong
    ny_job = w["ny"] * rk.world if weak else w["ny"]
    grid = Grid.uniform(w["nx"], ny_job, w["lx"], w["ly"] * (rk.world if weak else 1), cyl)
    params = make_params(w)
    consts = make_consts(w)
    if strips:
        uid = [nccl_unique_id() if rk.rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        model = Model.strip(grid, params, rk.rank, rk.world, uid[0], device=rk.local, flags=flags, consts=consts)
    else:
        opts = default_options()
        opts.device = rk.local
        opts.flags = flags
        opts.consts = consts
        model = Model(grid, params, options=opts)
    return model, grid, params, strong


def step_bytes_mode_c(nx, ny, solves_full, solves_elided, iterations, nu=2, fused=False, corrector_div=False):
    """Algorithmic bytes of one Mode C step (SURVEY 8d's rule: every stage reads each of its inputs
    and writes each of its outputs once, s = 8 bytes): predictor 4 sN + 2 N, residual / CFL maxima 4 sN; a solve that runs
    = divergence 3 + set-up 4 + corrector 7; a re-correction round that is converged before its first iteration =
    divergence 3 (its set-up pass and its identity corrector are not executed and not counted).  A CG iteration with a
    V(nu,nu) cycle, stage by stage (fused=False): level 0 = first sweep 2 + (nu - 1) sweeps 3 + restriction 2.25 +
    prolongation 2.25 + nu sweeps 3 + rho.z 2 + direction 3 + L d 2 + update 6 = 19.5 + 3 (2 nu - 1) sN (28.5 at nu = 2);
    a coarse level = 6.5 + 3 (2 nu - 1) s N_l (sum N_l = N / 3).  fused=True: the compulsory traffic of the kernels as they
    run (each leg of the cycle is one pass: level 0 descending 2.25 + ascending 3.25 + direction / L d 5 + update 6 = 16.5,
    a coarse level 5.5) -- the conservative figure beside the stage-by-stage one; corrector_div (single domain): the
    divergence of an elided re-correction round comes out of the corrector that precedes it (k_corrector_div), which
    leaves 1 sN (the rhs it writes) of that stage's 3 sN as traffic."""
    n = nx * ny
    if fused:
        per_it = 16.5 + 5.5 / 3.0
    else:
        per_it = 19.5 + 3 * (2 * nu - 1) + (6.5 + 3 * (2 * nu - 1)) / 3.0
    elided_sn = 1 if (fused and corrector_div) else 3
    return 8 * n * (8 + 14 * solves_full + elided_sn * solves_elided + per_it * iterations) + 2 * n


def legs_active(consts, flags):
    """Whether the library runs the V-cycle with one launch per leg (cfd_mg_legs.cuh; same rule as mg_setup)."""
    return 2 <= consts.mg_smoothing <= 4 and not (flags & 1024) and os.environ.get("CFD_MG_NO_LEGS") is None


def timed_steps(model, rk, steps, mode_c):
    """K steps, state resident in HBM: device time from the model's own CUDA events, max over the ranks."""
    rk.barrier()
    t0 = time.perf_counter()
    acc = dict(dev_ms=0.0, sweep_ms=0.0, sweeps=0, solves=0, launches=0, smooth_ms=0.0, smooth_n=0, first_its=0)
    its, per_step_ms, last = [], [], None
    for _ in range(steps):
        model.update()
        rk.tick()
        s_ms, sw_ms, n_l = model.last_timing()
        r = model.get_residuals()
        acc["dev_ms"] += s_ms
        acc["sweep_ms"] += sw_ms
        acc["sweeps"] += r.sweeps
        acc["solves"] += r.jacobi_calls
        acc["launches"] += n_l
        acc["first_its"] += r.f64["first_solve_iterations"]
        its.append(int(r.sweeps))
        per_step_ms.append(s_ms)
        last = r
        if mode_c:
            a, b = model.last_smoother_timing()
            acc["smooth_ms"] += a
            acc["smooth_n"] += b
    rk.barrier()
    acc["wall_s"] = time.perf_counter() - t0
    acc["its"], acc["per_step_ms"], acc["last"] = its, per_step_ms, last
    return acc


def parity_block(rk):
    """N > 1, outside every timed region: the strip decomposition against a single-domain model of the same problem on
    rank 0's GPU — Mode R (must be bit-identical, SURVEY N8: max-reductions only) and Mode C / MGCG at the shipped
    stopping rule (dot products are summed in another order: relative L2).  1040 x 600 channel with a cylinder."""
    import torch.distributed as dist
    from cfd_demo_b200 import _abi
    from cfd_demo_b200.model import Model, default_options, nccl_unique_id
    from cfd_demo_b200.types import Cylinder, Grid, SimulationParams, PressureSolver
    grid = Grid.uniform(1040, 600, 10.4, 6.0, Cylinder(2.6, 3.0, 0.45))
    out = {"grid": "1040x600 channel + cylinder", "ranks": rk.world}
    nx, ny = grid.nx, grid.ny
    shapes = {_abi.FIELD_P: (ny, nx), _abi.FIELD_U: (ny, nx + 1), _abi.FIELD_V: (ny + 1, nx)}
    for name, params, steps in (("mode_r", SimulationParams(), 14),
                                ("mode_c_mgcg", SimulationParams(dt=1e-3, viscosity=0.01, pressure_solver=PressureSolver.MGCG), 8)):
        uid = [nccl_unique_id() if rk.rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        consts = None
        if name != "mode_r":
            consts = make_consts({"consts": dict(cg_relative=1, cg_tolerance=1e-8, mg_smoothing=3)})
        strip = Model.strip(grid, params, rk.rank, rk.world, uid[0], device=rk.local, consts=consts)
        whole = None
        if rk.rank == 0:
            o = default_options()
            o.device = rk.local
            if consts is not None:
                o.consts = consts
            whole = Model(grid, params, options=o)
        counters_equal, its_s, its_w = True, [], []
        for _ in range(steps):
            strip.update()
            rk.tick()
            rs = strip.get_residuals()
            its_s.append(int(rs.sweeps))
            if whole is not None:
                whole.update()
                rw = whole.get_residuals()
                its_w.append(int(rw.sweeps))
                if name == "mode_r":
                    counters_equal &= (rs.jacobi_calls, rs.sweeps, rs.f64["p"], rs.f64["u"], rs.f64["v"], rs.f64["dt"]) == \
                                      (rw.jacobi_calls, rw.sweeps, rw.f64["p"], rw.f64["u"], rw.f64["v"], rw.f64["dt"])
        ja, jb = strip.rows()
        top = 1 if rk.rank == rk.world - 1 else 0
        res = {}
        for fid, shape in shapes.items():
            mine = strip.field(fid)
            parts = [None] * rk.world if rk.rank == 0 else None
            dist.gather_object((ja, jb + (top if fid == _abi.FIELD_V else 0), mine), parts, dst=0)
            if rk.rank == 0:
                ref = whole.field(fid).reshape(shape)
                got = np.empty_like(ref)
                for a, b, rows in parts:
                    got[a:b] = rows.reshape(b - a, shape[1])
                key = _abi.FIELD_NAMES[fid]
                if name == "mode_r":
                    res[key + "_entries_differing"] = int((got != ref).sum())
                else:
                    res[key + "_rel_l2"] = float(np.linalg.norm(got - ref) / max(np.linalg.norm(ref), 1e-300))
        if rk.rank == 0:
            if name == "mode_r":
                res["residuals_and_counters_identical_every_step"] = bool(counters_equal)
                res["sweeps_last_step"] = its_s[-1]
            else:
                res["iterations_per_step_strips"] = its_s
                res["iterations_per_step_single_domain"] = its_w
            res["steps"] = steps
            out[name] = res
            whole.close()
        strip.close()
        rk.barrier()
    return out


def run_extra(args, name, rk, steps=3):
    """A secondary workload inside the default run (rank count of the run; no e2e, no CPU leg): ms/step, kernel share, roofline."""
    w = WORKLOADS[name]
    mode_c = w["kind"] == "modeC"
    strips = rk.world > 1
    model, grid, params, strong = build_model(w, rk, strips)
    spinup = int(os.environ.get("CFD_BENCH_SPINUP", w["spinup"]))
    for _ in range(spinup):
        model.update()
        rk.tick()
    if mode_c:
        model.profile_smoother(True)
    for _ in range(2):
        model.update()
    acc = timed_steps(model, rk, steps, mode_c)
    dev_s, = rk.max_over_ranks([acc["dev_ms"] * 1e-3])
    nx, ny = grid.nx, grid.ny
    peak, _ = peak_hbm()
    k, s_ = acc["solves"] / steps, acc["sweeps"] / steps
    rank_cells = nx * ny // rk.world if strips else nx * ny
    if mode_c:
        full = k - (k - 1 if acc["first_its"] == acc["sweeps"] else 0)  # re-correction rounds without an iteration are elided
        bytes_step = step_bytes_mode_c(nx, ny, full, k - full, s_, nu=make_consts(w).mg_smoothing)
        sweep_us = acc["smooth_ms"] * 1e3 / max(acc["smooth_n"], 1)
    else:
        bytes_step = 8 * nx * ny * (8 + 10 * k + 3 * s_) + 2 * nx * ny
        sweep_us = acc["sweep_ms"] * 1e3 / max(acc["sweeps"], 1)
    out = {"workload": w["desc"], "nx": nx, "ny": ny, "n_gpus": rk.world, "scaling": "strong" if strong else ("weak" if strips else "single GPU"),
           "steps": steps, "ms_per_step": dev_s * 1e3 / steps, "cell_updates_per_s": nx * ny * steps / dev_s,
           "solves_per_step": k, ("cg_iterations_per_step" if mode_c else "sweeps_per_step"): s_,
           "step_frac_of_peak": bytes_step / (dev_s / steps) / 1e9 / (peak * rk.world),
           "sweep_us": sweep_us, "sweep_frac_of_peak": 3 * 8 * rank_cells / (sweep_us * 1e-6) / 1e9 / peak if sweep_us > 0 else None}
    if mode_c and legs_active(make_consts(w), 0):
        # the timed launch is the ascending leg of level 0 (nu sweeps + prolongation + rho.z in one pass), not a single sweep
        nu = make_consts(w).mg_smoothing
        out["sweep_us"] = None
        out["sweep_frac_of_peak"] = None
        out["ascending_leg_us"] = sweep_us
        out["ascending_leg_frac_of_peak_stage_bytes"] = (3 * nu + 4.25) * 8 * rank_cells / (sweep_us * 1e-6) / 1e9 / peak if sweep_us > 0 else None
        out["step_frac_of_peak_fused_traffic"] = step_bytes_mode_c(nx, ny, full, k - full, s_, nu=nu, fused=True, corrector_div=rk.world == 1) / (dev_s / steps) / 1e9 / (peak * rk.world)
    if mode_c:
        out["cg_iterations_list"] = acc["its"]
        out["ms_per_cg_iteration"] = dev_s * 1e3 / max(acc["sweeps"], 1)
    model.close()
    rk.barrier()
    return out


def run_ours(args, w):
    import torch
    import torch.distributed as dist

    rk = Ranks()
    world, rank, local_rank = rk.world, rk.rank, rk.local
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product has no CPU fallback (use --impl reference for the CPU port)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("cpu:gloo,cuda:nccl")
    mode_c = w["kind"] == "modeC"
    # N > 1: row strips.  Workloads marked `strong` cut the SAME grid into N strips; the others scale WEAKLY — every GPU
    # owns a w.nx x w.ny strip of one tall domain (nx x N*ny cells, same dx = dy).  Mode R: NCCL halo rows + max-allreduce
    # after every sweep (CFD_BENCH_FLAGS=512: the fused peer-memory sweep).  Mode C (MGCG): multigrid levels 0-3 in
    # strips, level 4 gathered, the rest of the hierarchy replicated, dot products sum-allreduced.
    # CFD_BENCH_REPLICAS=1: N independent replicas of the single-GPU workload instead (no data-path collective).
    strips = world > 1 and os.environ.get("CFD_BENCH_REPLICAS") != "1"
    flags = int(os.environ.get("CFD_BENCH_FLAGS", "0"))  # A/B hook
    model, grid, params, strong = build_model(w, rk, strips, flags)
    nx, ny = grid.nx, grid.ny
    cells = nx * ny * (1 if (strips or world == 1) else world)  # whole job

    # multi-rank runs: a rank that dies leaves the others waiting in a collective; abort instead of hanging the box
    if world > 1:
        def watchdog():
            while True:
                time.sleep(5.0)
                if time.time() - rk.progress > float(os.environ.get("CFD_BENCH_WATCHDOG_S", "240")):
                    print(f"bench.py rank {rank}: no progress for too long, aborting", file=sys.stderr, flush=True)
                    os._exit(3)
        threading.Thread(target=watchdog, daemon=True).start()

    # build the synthetic input on the device: spin the flow up (Mode R: to the dense regime where every step
    # saturates at K=21, S=1050; Mode C: through the 100-step ramp of the lid velocity and the transient after it)
    spinup = int(os.environ.get("CFD_BENCH_SPINUP", w["spinup"]))
    for i in range(spinup):
        model.update()
        rk.tick()
        if os.environ.get("CFD_BENCH_VERBOSE") and rank == 0:
            r_, t_ = model.get_residuals(), model.last_timing()
            print(f"spinup {i + 1}: K {r_.jacobi_calls} S {r_.sweeps} step_ms {t_[0]:.2f} per-sweep us {t_[1] * 1e3 / max(r_.sweeps, 1):.1f} "
                  f"rel {r_.f64['p_rel']:.2e} abs {r_.f64['p']:.2e}", file=sys.stderr, flush=True)
    if mode_c:
        model.profile_smoother(True)  # CUDA-event pairs around the smoother launches
    for _ in range(args.warmup):
        model.update()

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    cudart = None
    if os.environ.get("CFD_BENCH_PROFILE") == "1":  # opens the ncu window (--profile-from-start off) on region 1
        import ctypes
        cudart = ctypes.CDLL("libcudart.so.12")
        cudart.cudaProfilerStart()
    # ---- timed region 1: K steps, state resident in HBM -------------------------------------------------
    acc = timed_steps(model, rk, args.steps, mode_c)
    if cudart is not None:
        cudart.cudaProfilerStop()
    # ---- timed region 2: the same K steps through the reference-facing calls with HOST buffers ------------
    model.profile_smoother(False)
    pinned = model.pinned_snapshot_buffers()
    pinned2 = model.pinned_snapshot_buffers()
    # (a) the reference's own protocol: the snapshot of step n is requested after the step and picked up one step later
    # (Command::GetSnapshot ... get_last_available_snapshot, src/model.rs:100-102, :1300-1306), so its device->host copy
    # overlaps step n+1 — cfd_model_snapshot_begin / _end.  Every step's p, u, v arrive in host memory inside the region.
    # The region is timed E2E_PASSES times (K steps each, barrier on both sides, max over the ranks per pass) and the
    # FASTEST pass is reported, every pass listed beside it: on the shared GPU boxes the 201 MB per step over PCIe see
    # bursts of host-side interference (r2af: the same binary 3.7 - 12.6 ms/step within a minute, device time unchanged).
    e2e_passes = []
    d2h = 0
    for _ in range(max(1, int(os.environ.get("CFD_BENCH_E2E_PASSES", "5")))):
        rk.barrier()
        t1 = time.perf_counter()
        for k in range(args.steps):
            model.set_parameters(params)         # host -> device: the 28-byte parameter block
            model.update()
            res = model.get_residuals()          # device -> host: the step's residual scalars
            model.snapshot_begin(pinned if k % 2 == 0 else pinned2)  # device -> host: p, u, v narrowed to f32 (SimSnapshot, :36-42)
            if k > 0:
                snap = model.snapshot_end()      # the previous step's snapshot is complete in host memory
        snap = model.snapshot_end()
        d2h = (snap.p.nbytes + snap.u.nbytes + snap.v.nbytes + 8 * 8) * world
        rk.barrier()
        e2e_passes.append(time.perf_counter() - t1)
    # (b) the same with a blocking get_snapshot after every step (nothing overlaps)
    rk.barrier()
    t1b = time.perf_counter()
    for _ in range(min(args.steps, 5)):
        model.set_parameters(params)
        model.update()
        res = model.get_residuals()
        snap = model.get_snapshot(out=pinned)
    torch.cuda.synchronize()
    wall_e2e_sync = (time.perf_counter() - t1b) / min(args.steps, 5)
    # (c) between two frames of the reference's UI no snapshot is requested at all (src/model.rs:1296-1306: the solver thread
    # steps continuously and copies fields only on Command::GetSnapshot): parameters in, residual scalars out, every step
    torch.cuda.synchronize()
    t1c = time.perf_counter()
    for _ in range(args.steps):
        model.set_parameters(params)
        model.update()
        res = model.get_residuals()
    torch.cuda.synchronize()
    wall_e2e_scalars = (time.perf_counter() - t1c) / args.steps
    # the same with freshly allocated pageable buffers (what a caller holding plain Vec<f32>s gets)
    t2 = time.perf_counter()
    for _ in range(min(args.steps, 3)):
        model.set_parameters(params)
        model.update()
        res = model.get_residuals()
        snap = model.get_snapshot()
    torch.cuda.synchronize()
    wall_e2e_pageable = (time.perf_counter() - t2) / min(args.steps, 3)
    # and with the UI's colour map computed on the device: one nx*ny RGBA image instead of 3 fields
    wall_e2e_image = None
    if world == 1 or not strips:
        from cfd_demo_b200.model import PinnedBuffer
        img_buf = PinnedBuffer(nx * ny * 4, np.uint8)
        torch.cuda.synchronize()
        t3 = time.perf_counter()
        for _ in range(min(args.steps, 5)):
            model.set_parameters(params)
            model.update()
            res = model.get_residuals()
            model.render_rgba(1, out=img_buf)
        wall_e2e_image = (time.perf_counter() - t3) / min(args.steps, 5)
    clocks = sampler.stop() if rank == 0 else None

    dev_s, wall_s, *e2e_passes = rk.max_over_ranks([acc["dev_ms"] * 1e-3, acc["wall_s"]] + e2e_passes)
    wall_e2e_s = min(e2e_passes)
    steps = args.steps
    value = cells * steps / dev_s
    e2e_value = cells * steps / wall_e2e_s
    peak, peak_src = peak_hbm()
    rank_cells = nx * ny // world if strips else nx * ny
    algo_bytes = 3 * 8 * rank_cells  # per launch (one rank's strip): read p', rhs; write p'new
    sweeps, solves, launches = acc["sweeps"], acc["solves"], acc["launches"]
    consts = make_consts(w)
    legs = mode_c and legs_active(consts, flags)
    nu = int(consts.mg_smoothing)
    traffic_file = "ncu_sweep_kernel.json"
    roof_note = None
    if legs:
        # the dominant kernel: the ascending leg of the V-cycle on level 0 -- nu damped-Jacobi sweeps (the reference's update,
        # src/model.rs:788-793), the prolongation and rho.z in ONE pass (cfd_mg_legs.cuh).  Algorithmic bytes by SURVEY 8d's
        # rule, stage by stage: nu sweeps x 3 sN + prolongation 2.25 sN + rho.z 2 sN; the kernel itself moves 3.25 sN
        # (temporal blocking: `frac` may exceed 1, `traffic` / `frac_of_peak_by_traffic` say what went through HBM)
        sweep_us = acc["smooth_ms"] * 1e3 / max(acc["smooth_n"], 1)
        roof_launches, roof_share = acc["smooth_n"], acc["smooth_ms"] / max(acc["dev_ms"], 1e-9)
        algo_bytes = int((3 * nu + 4.25) * 8 * rank_cells)
        kernel_name = (f"cfdk::k_mg0_up3<double, {nu}> (ascending leg of the V({nu},{nu})-cycle on level 0: the prolongation, {nu} sweeps "
                       f"of the reference's damped-Jacobi update incl. its boundary rules and rho.z in one pass over HBM)")
        traffic_file = "ncu_leg_kernel.json"
        roof_note = (f"algorithmic bytes = the stages' compulsory traffic (SURVEY 8d): {nu} sweeps x 3 sN + prolongation 2.25 sN + rho.z "
                     f"2 sN = {3 * nu + 4.25} sN; the fused kernel's own compulsory traffic is 3.25 sN (temporal blocking), it is bound "
                     f"by fp64 issue, not by HBM: see profiles/r2_legs_ncu.md")
    elif mode_c:
        sweep_us = acc["smooth_ms"] * 1e3 / max(acc["smooth_n"], 1)
        roof_launches, roof_share = acc["smooth_n"], acc["smooth_ms"] / max(acc["dev_ms"], 1e-9)
        kernel_name = ("cfdk::k_jacobi_sweep5<double> (the reference's damped-Jacobi sweep incl. boundary update, here the "
                       "fine-level smoother of the V-cycle; with the fused passes one plain launch per CG iteration, the one "
                       "that also sums rho.z)")
    else:
        sweep_us = acc["sweep_ms"] * 1e3 / max(sweeps, 1)
        roof_launches, roof_share = sweeps, acc["sweep_ms"] / (dev_s * 1e3)
        kernel_name = "cfdk::k_jacobi_sweep5<double> (one damped-Jacobi sweep incl. boundary update and max|dp'|)"
    achieved = algo_bytes / (sweep_us * 1e-6) / 1e9 if sweep_us > 0 else 0.0
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", traffic_file)) as f:
            t = json.load(f)
            if t.get("nx") == nx and t.get("ny") == rank_cells // nx and (not legs or t.get("nu") == nu):
                traffic = t.get("dram_bytes_per_launch")
    except Exception:
        pass
    k_per_step, s_per_step = solves / steps, sweeps / steps
    replicas = world if not strips else 1
    if mode_c:
        # re-correction rounds that converge before their first iteration are elided (no set-up pass, no corrector)
        elided = (solves - steps) / steps if acc["first_its"] == sweeps else 0.0
        step_bytes = step_bytes_mode_c(nx, ny, k_per_step - elided, elided, s_per_step, nu=nu) * replicas
        step_bytes_r1 = step_bytes_mode_c(nx, ny, k_per_step, 0.0, s_per_step, nu=nu) * replicas
        step_bytes_fused = step_bytes_mode_c(nx, ny, k_per_step - elided, elided, s_per_step, nu=nu, fused=legs, corrector_div=world == 1) * replicas
    else:
        step_bytes = 8 * cells * (8 + 10 * k_per_step + 3 * s_per_step) + 2 * cells  # whole job
        step_bytes_r1 = step_bytes
        step_bytes_fused = step_bytes
    peak_job = peak * world

    parity = None
    if strips and os.environ.get("CFD_BENCH_NO_PARITY") != "1":
        parity = parity_block(rk)
    # CPU port timed on rank 0 at N = 1 only, from the state the timed regions ended in
    cpu = None
    if rank == 0 and not args.no_cpu_baseline and world == 1:
        from oracle import cpu_oracle
        cpu_oracle.build()
        cpu = cpu_baseline_mode_c(w, model) if mode_c else cpu_baseline_mode_r(w, nx, ny)
    extras = {}
    if args.workload == DEFAULT_WORKLOAD and os.environ.get("CFD_BENCH_NO_EXTRAS") != "1":
        names = ["cavity16384_modeC"] + (["channel8192x2048_modeR"] if strips else
                                          ["channel4096_modeR", "default800_modeR", "channel8192x2048_modeR"])
        for name in names:
            try:
                extras[name] = run_extra(args, name, rk)
            except Exception as e:  # an extra must never take the headline line down with it
                extras[name] = {"error": repr(e)}
                if world > 1:
                    raise

    if rank == 0:
        if world == 1:
            multi = "single domain"
        elif strips and not mode_c:
            multi = (f"{world} row strips, {'strong' if strong else 'weak'} scaling, " +
                     ("halo rows and max-reduction fused into the sweep kernel over NVLink peer memory (CFD_BENCH_FLAGS=512)"
                      if flags & 512 else "NCCL halo rows + max-allreduce after every sweep (CFD_BENCH_FLAGS=512: fused peer-memory sweep)"))
        elif strips:
            multi = (f"{world} row strips of one {nx}x{ny} cavity, {'strong' if strong else 'weak'} scaling: multigrid levels 0-3 in "
                     f"strips (one launch per leg of the V-cycle; per level the halo rows of rho before the descending leg and of the "
                     f"correction after the ascending one are exchanged), level 4 gathered and the rest replicated, dot products "
                     f"sum-allreduced; transport: {'NVLink peer memory (CFD_PEER_STRIPS=1)' if os.environ.get('CFD_PEER_STRIPS') == '1' else 'NCCL'}")
        else:
            multi = f"{world} independent replicas of the workload, one per GPU (CFD_BENCH_REPLICAS=1)"
        last = acc["last"]
        line = {
            "metric": "cell_updates_per_s", "value": value, "unit": "cell-updates/s", "n_gpus": world,
            "steps": steps, "warmup": args.warmup, "ms_per_step": dev_s * 1e3 / steps, "higher_is_better": True,
            "scaling": "strong" if strong else "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "timesteps_per_s": steps / dev_s * (world if (world > 1 and not strips) else 1),
            "wall_ms_per_step": wall_s * 1e3 / steps,
            "config": workload_config(w, w["nx"], w["ny"]),
            "job": {"nx": nx, "ny": ny, "multi_gpu": multi, "spinup_steps": spinup,
                    "start_state": "spun up on the device from rest (the reference arm starts from rest)"},
            "solves_per_step": k_per_step,
            ("cg_iterations_per_step" if mode_c else "sweeps_per_step"): s_per_step,
            "step_algorithmic_gbs": step_bytes / (dev_s / steps) / 1e9,
            "step_frac_of_peak": step_bytes / (dev_s / steps) / 1e9 / peak_job,
            "step_frac_of_peak_counting_elided_passes": step_bytes_r1 / (dev_s / steps) / 1e9 / peak_job,
            "step_frac_of_peak_fused_traffic": step_bytes_fused / (dev_s / steps) / 1e9 / peak_job,
            "step_bytes_accounting": ("step_frac_of_peak: SURVEY 8d's rule, every executed stage reads its inputs and writes its outputs "
                                      "once (a V(nu,nu) iteration = 19.5 + 3 (2 nu - 1) sN on level 0, 6.5 + 3 (2 nu - 1) s N_l on a "
                                      "coarse level); _fused_traffic: the compulsory traffic of the kernels as they run (each leg of the "
                                      "cycle is one pass over HBM: 16.5 sN + 5.5 s N_l per iteration) -- the conservative figure"),
            "e2e": {"value": e2e_value, "unit": "cell-updates/s", "h2d_bytes_per_step": 28 * world, "d2h_bytes_per_step": d2h,
                    "ms_per_step": wall_e2e_s * 1e3 / steps,
                    "ms_per_step_passes": [x * 1e3 / steps for x in e2e_passes],
                    "ms_per_step_median_pass": statistics.median(e2e_passes) * 1e3 / steps,
                    "passes": (f"{len(e2e_passes)} passes of {steps} steps, each timed like the contract's region (barrier on both "
                               "sides, max over ranks); value = the fastest pass, all passes listed (the boxes are shared: PCIe / "
                               "host-memory interference comes in bursts and is not a property of the path)"),
                    "ms_per_step_blocking_get_snapshot": wall_e2e_sync * 1e3,
                    "ms_per_step_residuals_only": wall_e2e_scalars * 1e3,
                    "ms_per_step_pageable_destination": wall_e2e_pageable * 1e3,
                    "ms_per_step_rgba_image_instead": None if wall_e2e_image is None else wall_e2e_image * 1e3,
                    "calls": "cfd_model_set_params + cfd_model_update + cfd_model_get_residuals + cfd_model_snapshot_begin / _end "
                             "(every step's p, u, v as f32 into cfd_host_alloc'ed pinned buffers; the copy of step n overlaps step "
                             "n+1, the reference's request-now / pick-up-later snapshot protocol)"},
            "gpu_launches": launches,
            "roofline": {"kernel": kernel_name,
                         "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak,
                         "peak_source": peak_src, "traffic": traffic,
                         "algorithmic_bytes_per_launch": algo_bytes, "avg_launch_us": sweep_us,
                         "launches_timed": roof_launches,
                         "share_of_step": roof_share,
                         "frac_of_peak_by_traffic": (traffic / (sweep_us * 1e-6) / 1e9 / peak) if (traffic and sweep_us > 0) else None,
                         "note": roof_note},
            "cpu_baseline": cpu,
            "clocks": clocks,
        }
        if mode_c:
            line["cg_iterations_list"] = acc["its"]
            line["ms_per_cg_iteration"] = dev_s * 1e3 / max(sweeps, 1)
            line["ms_per_step_list"] = [round(x, 4) for x in acc["per_step_ms"]]
            line["stop"] = {"rule": w.get("stop_rule", "dt * rms(r) <= cg_tolerance"),
                            "rel_residual": last.f64["p_rel"], "dt_rms_residual": last.f64["p"],
                            "dt_rms_rhs": last.f64["rhs_rms"], "of": "the last timed step (first solve / last solve / first solve)"}
        if parity is not None:
            line["parity"] = parity
        if extras:
            line["extra"] = extras
        print(json.dumps(line), flush=True)
    model.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--workload", choices=sorted(WORKLOADS), default=DEFAULT_WORKLOAD)
    ap.add_argument("--ref-precision", type=int, default=64)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    w = WORKLOADS[args.workload]
    if args.impl == "reference":
        return run_reference(args, w)
    return run_ours(args, w)


if __name__ == "__main__":
    sys.exit(main())
