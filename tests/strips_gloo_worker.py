"""Worker for tests/test_strips_gloo.py (run under torchrun, backend gloo, CPU only): the multi-rank design of
DESIGN.md section 7 — which halo rows are refreshed when, which scalars are max- / sum-reduced — executed with the
CPU oracle on every rank's strip, halos and reductions over torch.distributed, and compared with the
single-domain oracle.  Mode R must be bit-identical for any rank count (max-reductions only)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cfd_demo_b200 import _abi  # noqa: E402
from cfd_demo_b200.types import (Cylinder, Grid, InletProfile, PressureSolver, Scenario, SimulationParams,  # noqa: E402
                                 VelocityScheme)
from cfd_demo_b200.model import strip_rows  # noqa: E402  (host-only partition query; no GPU needed)
from oracle.cpu_oracle import OracleModel, default_consts  # noqa: E402


def split_rows(ny, rank, world):
    base, rem = divmod(ny, world)
    ja = rank * base + min(rank, rem)
    return ja, ja + base + (1 if rank < rem else 0)


def main():
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()

    def exchange(f, ja, jb, below, above):
        ops, keep = [], []
        if rank > 0:
            if above > 0:
                t = torch.from_numpy(f[ja:ja + above].copy()); keep.append(t)
                ops.append(dist.isend(t, rank - 1))
            if below > 0:
                r = torch.from_numpy(f[ja - below:ja]); keep.append(r)
                ops.append(dist.irecv(r, rank - 1))
        if rank < world - 1:
            if below > 0:
                t = torch.from_numpy(f[jb - below:jb].copy()); keep.append(t)
                ops.append(dist.isend(t, rank + 1))
            if above > 0:
                r = torch.from_numpy(f[jb:jb + above]); keep.append(r)
                ops.append(dist.irecv(r, rank + 1))
        for o in ops:
            o.wait()

    def allreduce(x, op):
        t = torch.tensor([x], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX if op == 0 else dist.ReduceOp.SUM)
        return float(t[0])

    def gather(f, lo, hi):
        parts = [None] * world
        dist.all_gather_object(parts, (lo, hi, f[lo:hi].copy()))
        for r, (a, b, rows) in enumerate(parts):
            if r != rank:
                f[a:b] = rows

    cyl = Cylinder(7.5, 5.0, 0.75)
    cg = default_consts()
    cg.cg_tolerance = 1e-13
    cg3 = default_consts()  # the cycle bench.py ships: V(3,3), relative stopping rule
    cg3.cg_tolerance, cg3.cg_relative, cg3.mg_smoothing = 1e-12, 1, 3
    cases = [
        ("modeR first order", Grid.uniform(64, 24, 30.0, 10.0, cyl), SimulationParams(), 64, None, 14, True),
        ("modeR second order f32", Grid.uniform(72, 26, 30.0, 10.0, cyl),
         SimulationParams(velocity_scheme=VelocityScheme.SecondOrder, inlet_profile=InletProfile.Parabolic), 32, None, 12, True),
        ("modeR cavity", Grid.uniform(48, 48, 1.0, 1.0, None),
         SimulationParams(dt=1e-3, viscosity=0.01, scenario=Scenario.Cavity), 64, None, 10, True),
        ("modeC channel", Grid.uniform(48, 32, 30.0, 10.0, cyl),
         SimulationParams(dt=1e-3, viscosity=0.01, pressure_solver=PressureSolver.CG), 64, cg, 8, False),
        # multigrid-preconditioned CG on strips: level 0 in strips, level-1 residual gathered, coarse levels replicated;
        # the strips are the library's aligned partition (cfd_strip_rows), which pairs up the unknown rows
        ("modeC mgcg cavity", Grid.uniform(64, 120, 64 / 64.0, 120 / 64.0, None),
         SimulationParams(dt=1e-3, viscosity=0.01, scenario=Scenario.Cavity, pressure_solver=PressureSolver.MGCG), 64, cg, 8, False),
        ("modeC mgcg channel", Grid.uniform(96, 103, 9.6, 10.3, Cylinder(2.4, 5.0, 0.9)),
         SimulationParams(dt=1e-3, viscosity=0.01, pressure_solver=PressureSolver.MGCG), 64, cg, 8, False),
        ("modeC mgcg V(3,3) relative rule", Grid.uniform(64, 120, 64 / 64.0, 120 / 64.0, None),
         SimulationParams(dt=1e-3, viscosity=0.01, scenario=Scenario.Cavity, pressure_solver=PressureSolver.MGCG), 64, cg3, 8, False),
    ]
    for name, grid, params, precision, consts, steps, exact in cases:
        nx, ny = grid.nx, grid.ny
        mgcg = params.pressure_solver == PressureSolver.MGCG
        ja, jb = strip_rows(ny, world, rank) if mgcg else split_rows(ny, rank, world)
        top = rank == world - 1
        strip = OracleModel(grid, params, precision=precision, consts=consts)
        strip.set_strip(ja, jb, top, exchange, allreduce, gather)
        whole = OracleModel(grid, params, precision=precision, consts=consts)
        for s in range(steps):
            strip.update()
            whole.update()
            rs, rw = strip.get_residuals(), whole.get_residuals()
            if mgcg:
                assert rs.jacobi_calls == rw.jacobi_calls == 2 and abs(rs.sweeps - rw.sweeps) <= 1, (name, s, rs.sweeps, rw.sweeps)
            if exact:
                assert (rs.jacobi_calls, rs.sweeps) == (rw.jacobi_calls, rw.sweeps), (name, s)
                for k in ("dt", "p", "u", "v"):
                    assert rs.f64[k] == rw.f64[k], (name, s, k, rs.f64[k], rw.f64[k])
        for fid in (_abi.FIELD_P, _abi.FIELD_U, _abi.FIELD_V, _abi.FIELD_U_STAR, _abi.FIELD_V_STAR, _abi.FIELD_P_PRIME):
            a, b = strip.field(fid), whole.field(fid)
            if fid in (_abi.FIELD_U, _abi.FIELD_U_STAR):
                a, b = a.reshape(ny, nx + 1)[ja:jb], b.reshape(ny, nx + 1)[ja:jb]
            elif fid in (_abi.FIELD_V, _abi.FIELD_V_STAR):
                a, b = a.reshape(ny + 1, nx)[ja:jb + top], b.reshape(ny + 1, nx)[ja:jb + top]
            else:
                a, b = a.reshape(ny, nx)[ja:jb], b.reshape(ny, nx)[ja:jb]
            if exact:
                assert np.array_equal(a, b), (name, rank, _abi.FIELD_NAMES[fid])
            else:
                d = np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)
                assert d <= 1e-9, (name, rank, _abi.FIELD_NAMES[fid], d)
        assert np.abs(whole.field(_abi.FIELD_U)).max() > 0
        dist.barrier()
    if rank == 0:
        print(f"gloo strips ok: {world} ranks, {len(cases)} cases")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
