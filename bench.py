#!/usr/bin/env python
"""bench.py — throughput of the solver hot path (one `Model::update` timestep, src/model.rs:304-379).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]

A "step" is one timestep of the BASELINE 4096x4096 workload (predictor, K pressure solves of <= 50 damped
Jacobi sweeps each with the corrector after each, boundary conditions, residual / CFL reductions).  State is
resident in HBM when the timed region starts.  Prints ONE JSON line (contract in the task description):
`value` = cell-updates/s (= nx*ny*timesteps/s) of the whole job, `e2e` = the same metric through the C ABI
with HOST buffers every step (set_params in, residuals + f32 snapshot out), `roofline` for the Jacobi sweep
kernel from live CUDA-event timing, `cpu_baseline` = the CPU oracle (a C++ port of the reference, 1 core
because the reference solver is single-threaded by construction, src/model.rs:1287) on a bounded sample.

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

WORKLOADS = {
    # name: (nx, ny, lx, ly, cylinder, params kwargs, spin-up steps to reach the dense saturated regime: until
    # p' is non-zero everywhere the sweeps still hit the slow zero-dividend path; ~21 steps at 4096x4096, ~35 for
    # the taller multi-GPU domains)
    "channel4096_modeR": dict(nx=4096, ny=4096, lx=40.0, ly=40.0, cylinder=None, params={}, spinup=44,
                              desc="channel 4096x4096 (reference scenario), fp64, Mode R = the reference's damped "
                                   "Jacobi (<=50 sweeps) + <=20 outer re-corrections, dense saturated regime "
                                   "(K=21 solves, S=1050 sweeps per step)"),
    "default800_modeR": dict(nx=800, ny=264, lx=30.0, ly=10.0, cylinder=(7.5, 5.0, 0.75), params={}, spinup=26,
                             desc="reference default_grid() 800x264 + cylinder, Mode R"),
}
DEFAULT_WORKLOAD = "channel4096_modeR"


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
# CPU side: the oracle port, timed on a bounded sample (one outer round = `u*<-u` copies + divergence +
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


def run_reference(args, w):
    """The reference arm: the CPU port of the reference (oracle), 1 core, bounded samples."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
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
    line = {
        "impl": "reference", "metric": "cell_updates_per_s", "value": value, "unit": "cell-updates/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": s["step_seconds"] * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32" if precision == 32 else "f64", "data": "synthetic",
        "timesteps_per_s": 1.0 / s["step_seconds"],
        "config": {"workload": w["desc"], "nx": w["nx"], "ny": w["ny"], "solves_per_step": 21, "sweeps_per_step": 1050},
        "cpu_baseline": {"value": value, "unit": "cell-updates/s", "cores": 1, "kind": "port", "sample": sample,
                         "host_cores": os.cpu_count(),
                         "note": "C++ port of the reference (oracle/cfd_oracle.hpp); the Rust reference cannot be "
                                 "built here (no Rust toolchain) and is single-threaded by construction"},
        "e2e": {"value": value, "unit": "cell-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)
    return 0


def run_ours(args, w):
    import torch
    import torch.distributed as dist
    from cfd_demo_b200.model import Model

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product has no CPU fallback (use --impl reference for the CPU port)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("cpu:gloo,cuda:nccl")

    # N > 1: WEAK scaling over row strips — every GPU owns a w.nx x w.ny strip of one tall channel
    # (nx x N*ny cells, same dx = dy), halo rows and max-reductions over NCCL
    from cfd_demo_b200.types import Cylinder, Grid
    cyl = Cylinder(*w["cylinder"]) if w["cylinder"] else None
    grid = Grid.uniform(w["nx"], w["ny"] * world, w["lx"], w["ly"] * world, cyl)
    params = make_params(w)
    nx, ny = grid.nx, grid.ny
    cells = nx * ny  # whole job
    from cfd_demo_b200.model import default_options, nccl_unique_id
    if world > 1:
        uid = [nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        model = Model.strip(grid, params, rank, world, uid[0], device=local_rank,
                            flags=int(os.environ.get("CFD_BENCH_FLAGS", "0")))  # A/B hook (e.g. 32 = NCCL exchange)
    else:
        opts = default_options()
        opts.device = local_rank
        model = Model(grid, params, options=opts)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # build the synthetic input: spin the flow up to the dense regime where every step saturates (K=21,S=1050)
    spinup = int(os.environ.get("CFD_BENCH_SPINUP", w["spinup"]))
    for i in range(spinup):
        model.update()
        if os.environ.get("CFD_BENCH_VERBOSE") and rank == 0:
            r_, t_ = model.get_residuals(), model.last_timing()
            print(f"spinup {i + 1}: K {r_.jacobi_calls} S {r_.sweeps} step_ms {t_[0]:.2f} per-sweep us {t_[1] * 1e3 / max(r_.sweeps, 1):.1f}",
                  file=sys.stderr, flush=True)
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
    barrier()
    t0 = time.perf_counter()
    dev_ms, sweep_ms, sweeps, solves, launches = 0.0, 0.0, 0, 0, 0
    for _ in range(args.steps):
        model.update()
        s_ms, sw_ms, n_l = model.last_timing()
        r = model.get_residuals()
        dev_ms += s_ms
        sweep_ms += sw_ms
        sweeps += r.sweeps
        solves += r.jacobi_calls
        launches += n_l
    barrier()
    wall = time.perf_counter() - t0
    if cudart is not None:
        cudart.cudaProfilerStop()
    # ---- timed region 2: the same K steps through the reference-facing calls with HOST buffers ------------
    barrier()
    t1 = time.perf_counter()
    d2h = 0
    for _ in range(args.steps):
        model.set_parameters(params)      # host -> device: the 28-byte parameter block
        model.update()
        res = model.get_residuals()       # device -> host: the step's residual scalars
        snap = model.get_snapshot()       # device -> host: p, u, v narrowed to f32 (SimSnapshot, src/model.rs:36-42)
        d2h = (snap.p.nbytes + snap.u.nbytes + snap.v.nbytes + 8 * 8) * world
        launches_e2e = model.last_timing()[2] + 3
    barrier()
    wall_e2e = time.perf_counter() - t1
    clocks = sampler.stop() if rank == 0 else None

    # max over ranks
    times = torch.tensor([dev_ms * 1e-3, wall, wall_e2e], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    dev_s, wall_s, wall_e2e_s = [float(x) for x in times.tolist()]
    steps = args.steps
    value = cells * steps / dev_s
    e2e_value = cells * steps / wall_e2e_s
    peak, peak_src = peak_hbm()
    sweep_us = sweep_ms * 1e3 / max(sweeps, 1)
    algo_bytes = 3 * 8 * cells // world  # per launch (one rank's strip): read p', rhs; write p'new
    achieved = algo_bytes / (sweep_us * 1e-6) / 1e9
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_sweep_kernel.json")) as f:
            t = json.load(f)
            if t.get("nx") == nx and t.get("ny") == ny:
                traffic = t.get("dram_bytes_per_launch")
    except Exception:
        pass
    k_per_step, s_per_step = solves / steps, sweeps / steps
    step_bytes = 8 * cells * (8 + 10 * k_per_step + 3 * s_per_step) + 2 * cells  # whole job
    peak = peak * world

    if rank == 0:
        cpu = None
        if not args.no_cpu_baseline and world == 1:  # CPU port timed on rank 0 at N = 1 only
            from oracle import cpu_oracle
            cpu_oracle.build()
            rounds = max(1, min(8, int(24.0 / (3.0 * w["nx"] * w["ny"] / (4096.0 * 4096.0) + 1e-9))))
            s = cpu_oracle_sample(w, 64, rounds)
            s32 = cpu_oracle_sample(w, 32, max(1, rounds // 2))
            cpu = {"value": s["cells"] / s["step_seconds"], "unit": "cell-updates/s", "cores": 1, "kind": "port",
                   "host_cores": os.cpu_count(),
                   "value_f32": s32["cells"] / s32["step_seconds"],
                   "sample": (f"oracle<double>: {s['rounds']} outer rounds (copies + divergence + "
                              f"{s['sweeps_per_round']:.0f}-sweep Jacobi solve + corrector) on dense synthetic "
                              f"{nx}x{ny} fields, median {s['t_round']:.3f} s/round, x21 rounds per saturated "
                              f"timestep + {s['t_misc']:.3f} s predictor/BC; value_f32 = same with oracle<float> "
                              f"(the reference's own precision)")}
        line = {
            "metric": "cell_updates_per_s", "value": value, "unit": "cell-updates/s", "n_gpus": world,
            "steps": steps, "warmup": args.warmup, "ms_per_step": dev_s * 1e3 / steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "timesteps_per_s": steps / dev_s,
            "wall_ms_per_step": wall_s * 1e3 / steps,
            "config": {"workload": w["desc"], "nx": nx, "ny": ny, "spinup_steps": spinup,
                       "solves_per_step": k_per_step, "sweeps_per_step": s_per_step,
                       "l2": "every field (134 MB at 4096^2) is larger than L2 (126 MB); no flush needed",
                       "timing": "CUDA events on the model's stream around each update(), summed over K steps",
                       "multi_gpu": "single domain" if world == 1 else
                       f"{world} row strips of {w['nx']}x{w['ny']} cells each over NCCL (halo rows per sweep + max allreduce), weak scaling"},
            "step_algorithmic_gbs": step_bytes / (dev_s / steps) / 1e9,
            "step_frac_of_peak": step_bytes / (dev_s / steps) / 1e9 / peak,
            "e2e": {"value": e2e_value, "unit": "cell-updates/s", "h2d_bytes_per_step": 28, "d2h_bytes_per_step": d2h,
                    "ms_per_step": wall_e2e_s * 1e3 / steps,
                    "calls": "cfd_model_set_params + cfd_model_update + cfd_model_get_residuals + cfd_model_get_snapshot"},
            "gpu_launches": launches,
            "roofline": {"kernel": "cfdk::k_jacobi_sweep5<double> (one damped-Jacobi sweep incl. boundary update and max|dp'|)",
                         "bound": "hbm", "achieved": achieved, "peak": peak / world, "unit": "GB/s",
                         "frac": achieved / (peak / world),
                         "peak_source": peak_src, "traffic": traffic,
                         "algorithmic_bytes_per_launch": algo_bytes, "avg_launch_us": sweep_us,
                         "launches_timed": sweeps,
                         "share_of_step": sweep_ms / (dev_s * 1e3)},
            "cpu_baseline": cpu,
            "clocks": clocks,
        }
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
