"""cfd_demo_b200 — B200-native (sm_100a) implementation of cfd-demo's per-timestep solver hot path.

Host-side mirror of the reference `Model` API (src/model.rs) over the C ABI in include/cfd_b200.h.
The CUDA library is loaded lazily by `cfd_demo_b200.model`; there is no CPU fallback.
"""
from .types import (Cylinder, Grid, InletProfile, PressureSolver, Residuals, Scenario, SimSnapshot,
                    SimulationParams, VelocityScheme, default_grid)

__all__ = ["Cylinder", "Grid", "InletProfile", "PressureSolver", "Residuals", "Scenario", "SimSnapshot",
           "SimulationParams", "VelocityScheme", "default_grid"]
