"""Host-side mirror of the reference's public solver types (reference: src/model.rs).

`Grid` (:121-131), `Cylinder` (:134-139), `SimulationParams` (:13-21, defaults :44-55), the enums
`VelocityScheme` (:142-146), `PressureSolver` (:149-152), `InletProfile` (:155-159), `Residuals` (:23-32)
and `SimSnapshot` (:36-42) keep the reference's names and field meaning.  `Scenario` and
`PressureSolver.CG` are extensions (the reference hard-codes the channel and has Jacobi only).
All reals here are rounded to float32 on their way into the C ABI, because they are f32 in the reference.
"""
from __future__ import annotations

import enum
from dataclasses import dataclass, field
from typing import Optional

import numpy as np

from . import _abi

f32 = np.float32


class VelocityScheme(enum.IntEnum):
    FirstOrder = _abi.SCHEME_FIRST_ORDER
    SecondOrder = _abi.SCHEME_SECOND_ORDER
    Quick = _abi.SCHEME_QUICK  # extension: the JS twin's QUICK face values (index.html:471-549, :643-723)


class PressureSolver(enum.IntEnum):
    Jacobi = _abi.SOLVER_JACOBI
    CG = _abi.SOLVER_CG  # extension
    MGCG = _abi.SOLVER_MGCG  # extension: multigrid-preconditioned CG


class InletProfile(enum.IntEnum):
    Uniform = _abi.INLET_UNIFORM
    Parabolic = _abi.INLET_PARABOLIC


class Scenario(enum.IntEnum):
    Channel = _abi.SCENARIO_CHANNEL  # the reference's only scenario
    Cavity = _abi.SCENARIO_CAVITY    # extension


@dataclass
class Cylinder:
    center_x: float
    center_y: float
    radius: float


@dataclass
class Grid:
    nx: int
    ny: int
    lx: float
    ly: float
    dx: float
    dy: float
    obstacle: Optional[Cylinder] = None

    @staticmethod
    def uniform(nx: int, ny: int, lx: float, ly: float, obstacle: Optional[Cylinder] = None) -> "Grid":
        """dx, dy computed in f32 exactly as `default_grid()` does (reference: src/app.rs:38-39)."""
        dx = f32(lx) / f32(nx)
        dy = f32(ly) / f32(ny)
        return Grid(nx, ny, float(f32(lx)), float(f32(ly)), float(dx), float(dy), obstacle)

    def to_c(self) -> _abi.CfdGrid:
        g = _abi.CfdGrid()
        g.nx, g.ny = int(self.nx), int(self.ny)
        g.lx, g.ly = float(f32(self.lx)), float(f32(self.ly))
        g.dx, g.dy = float(f32(self.dx)), float(f32(self.dy))
        if self.obstacle is not None:
            g.has_obstacle = 1
            g.center_x = float(f32(self.obstacle.center_x))
            g.center_y = float(f32(self.obstacle.center_y))
            g.radius = float(f32(self.obstacle.radius))
        else:
            g.has_obstacle = 0
            g.center_x = g.center_y = g.radius = 0.0
        return g


def default_grid() -> Grid:
    """`default_grid()` of the reference UI (src/app.rs:33-53): 800x264 cells, 30x10 domain, cylinder."""
    lx, ly = f32(30.0), f32(10.0)
    return Grid.uniform(800, 264, 30.0, 10.0,
                        Cylinder(float(lx / f32(4.0)), float(ly / f32(2.0)), 0.75))


@dataclass
class SimulationParams:
    dt: float = 0.005
    viscosity: float = 0.000001
    target_inlet_velocity: float = 1.0
    velocity_scheme: VelocityScheme = VelocityScheme.FirstOrder
    inlet_profile: InletProfile = InletProfile.Uniform
    pressure_solver: PressureSolver = PressureSolver.Jacobi
    scenario: Scenario = Scenario.Channel  # extension; the reference has no such field

    def to_c(self) -> _abi.CfdParams:
        p = _abi.CfdParams()
        p.dt = float(f32(self.dt))
        p.viscosity = float(f32(self.viscosity))
        p.target_inlet_velocity = float(f32(self.target_inlet_velocity))
        p.velocity_scheme = int(self.velocity_scheme)
        p.inlet_profile = int(self.inlet_profile)
        p.pressure_solver = int(self.pressure_solver)
        p.scenario = int(self.scenario)
        return p


@dataclass
class Residuals:
    simulation_step: int
    simulation_time: float
    dt: float
    p: float
    u: float
    v: float
    step_time: float  # seconds (Duration in the reference)
    piso_substeps: int
    # additions
    jacobi_calls: int = 0
    sweeps: int = 0
    f64: dict = field(default_factory=dict)

    @staticmethod
    def from_c(r: _abi.CfdResiduals) -> "Residuals":
        return Residuals(int(r.simulation_step), float(r.simulation_time), float(r.dt), float(r.p),
                         float(r.u), float(r.v), float(r.step_seconds), int(r.piso_substeps),
                         int(r.jacobi_calls), int(r.sweeps),
                         {"simulation_time": r.simulation_time_f64, "dt": r.dt_f64, "p": r.p_f64,
                          "u": r.u_f64, "v": r.v_f64, "p_rel": r.p_rel_f64, "rhs_rms": r.rhs_rms_f64,
                          "first_solve_iterations": int(r.first_solve_iterations)})


@dataclass
class SimSnapshot:
    p: np.ndarray
    u: np.ndarray
    v: np.ndarray
    dt: float
    paused: bool = False
