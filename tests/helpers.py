"""Shared helpers for the parity tests: grids, and exact field comparison between the CUDA path and the oracle."""
import numpy as np

from cfd_demo_b200 import _abi
from cfd_demo_b200.types import Cylinder, Grid

STATE_FIELDS = [_abi.FIELD_P, _abi.FIELD_U, _abi.FIELD_V, _abi.FIELD_U_STAR, _abi.FIELD_V_STAR,
                _abi.FIELD_RHS, _abi.FIELD_P_PRIME, _abi.FIELD_U_OLD, _abi.FIELD_V_OLD]


def channel_grid(nx, ny, lx=None, ly=None, cylinder=True):
    """A channel like the reference's default_grid() (src/app.rs:33-53), scaled to nx x ny."""
    lx = 30.0 if lx is None else lx
    ly = 10.0 if ly is None else ly
    cyl = Cylinder(lx / 4.0, ly / 2.0, 0.075 * ly) if cylinder else None
    return Grid.uniform(nx, ny, lx, ly, cyl)


def box_grid(n, m=None):
    return Grid.uniform(n, n if m is None else m, 1.0, 1.0, None)


def assert_fields_identical(gpu, cpu, fields=STATE_FIELDS, context=""):
    """Bit-level agreement (numerically equal, NaNs in the same places; +0 == -0)."""
    for fid in fields:
        a, b = gpu.field(fid), cpu.field(fid)
        assert a.shape == b.shape, (context, _abi.FIELD_NAMES[fid])
        same = (a == b) | (np.isnan(a) & np.isnan(b))
        if not same.all():
            bad = np.flatnonzero(~same)
            k = bad[0]
            raise AssertionError(
                f"{context}: field {_abi.FIELD_NAMES[fid]} differs at {bad.size} of {a.size} entries; first at "
                f"flat index {k}: gpu={a[k]!r} oracle={b[k]!r} (max abs diff {np.nanmax(np.abs(a - b)):.3e})")


def assert_residuals_identical(rg, rc, context=""):
    assert rg.simulation_step == rc.simulation_step, context
    assert rg.jacobi_calls == rc.jacobi_calls, (context, rg.jacobi_calls, rc.jacobi_calls)
    assert rg.sweeps == rc.sweeps, (context, rg.sweeps, rc.sweeps)
    assert rg.piso_substeps == rc.piso_substeps, context
    for k in ("simulation_time", "dt", "p", "u", "v"):
        assert rg.f64[k] == rc.f64[k], (context, k, rg.f64[k], rc.f64[k])
    for k in ("simulation_time", "dt", "p", "u", "v"):
        assert getattr(rg, k) == getattr(rc, k), (context, k)


def rel_l2(a, b):
    d = np.linalg.norm(a - b)
    n = np.linalg.norm(b)
    return d / n if n > 0 else d
