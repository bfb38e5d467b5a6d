"""CPU tests of the oracle's SURVEY 8f extensions: QUICK face values (index.html:471-549, :643-723), the reference's
commented-out sub-step adaptation made live (src/model.rs:352-363), the relative Mode C stopping rule (SURVEY 8d
config 3) and the tracer particles (index.html:1472-1543).  No reference counterpart can be run here (no Rust, no
node): these pin the extensions to facts read off the code."""
import math

import numpy as np
import pytest

from cfd_demo_b200 import _abi
from cfd_demo_b200.types import (Grid, PressureSolver, Scenario, SimulationParams, VelocityScheme)
from oracle.cpu_oracle import OracleModel, default_consts

from helpers import box_grid, channel_grid, rel_l2


@pytest.fixture(scope="module", autouse=True)
def _built(oracle_built):
    return oracle_built


@pytest.mark.parametrize("precision", [32, 64])
@pytest.mark.parametrize("nx,ny", [(16, 4), (24, 7), (64, 33)])
def test_quick_reads_stay_in_bounds_and_differ_from_second_order(precision, nx, ny):
    """The bounds-checked build aborts on any out-of-range read: QUICK reaches two cells up- and downstream
    (index.html:481,514) and, on column nx of the u equation, through the flat index into the next row."""
    g = channel_grid(nx, ny, cylinder=ny >= 7)
    q = OracleModel(g, SimulationParams(velocity_scheme=VelocityScheme.Quick, target_inlet_velocity=2.0), precision=precision, checked=True)
    s = OracleModel(g, SimulationParams(velocity_scheme=VelocityScheme.SecondOrder, target_inlet_velocity=2.0), precision=precision, checked=True)
    for _ in range(12):
        q.update()
        s.update()
    uq, us = q.field(_abi.FIELD_U), s.field(_abi.FIELD_U)
    assert np.isfinite(uq).all() and np.abs(uq).max() > 0
    if ny >= 7:
        assert not np.array_equal(uq, us) and rel_l2(uq, us) < 0.2  # another scheme, the same flow


def test_quick_face_value_of_a_linear_profile_is_the_linear_interpolant():
    """(-a + 6 b + 3 c) / 8 and (3 a + 6 b - c) / 8 reproduce the face value b + (c - b) / 2 of any linear profile
    exactly (QUICK is third-order accurate), which 1.5 b - 0.5 a (second-order upwind) does too: on a uniform shear
    flow u = y (no x dependence, v = 0) the two schemes give bit-identical u predictors away from the walls."""
    n = 32
    g = box_grid(n)
    outs = []
    for scheme in (VelocityScheme.SecondOrder, VelocityScheme.Quick):
        m = OracleModel(g, SimulationParams(dt=1e-3, viscosity=0.0, velocity_scheme=scheme), precision=64, checked=True)
        u = np.tile(((np.arange(n) + 0.5) / n)[:, None], (1, n + 1))
        m.set_field(_abi.FIELD_U, u.ravel())
        m.stage(OracleModel.STAGE_PREDICTOR_U)
        outs.append(m.field(_abi.FIELD_U_STAR).reshape(n, n + 1))
    assert np.array_equal(outs[0][3:-3, 3:-3], outs[1][3:-3, 3:-3])


def test_adaptive_substeps_follow_the_commented_out_rule():
    """src/model.rs:352-363, live: error = last_pressure_residual; error > 1e-3 -> substeps = min(ceil(substeps * error /
    1e-3), 20); error < 5e-4 and substeps > 1 -> floor(substeps / 2).  Replayed here from the residual log."""
    c = default_consts()
    c.adaptive_substeps = 1
    m = OracleModel(channel_grid(96, 32), SimulationParams(dt=0.02, target_inlet_velocity=3.0), precision=64, consts=c)
    sub, seen = 1, set()
    for _ in range(32):
        m.update()
        r = m.get_residuals()
        assert r.piso_substeps == sub
        assert r.jacobi_calls <= 21 * sub and r.jacobi_calls >= 2 * sub
        e = r.f64["p"]
        if e > 1e-3:
            sub = int(min(math.ceil(sub * (e / 1e-3)), 20.0))
        elif e < 1e-3 / 2.0 and sub > 1:
            sub = max(int(math.floor(sub / 2.0)), 1)
        seen.add(sub)
    assert 20 in seen and 1 in seen  # the rule fired and saturated at its cap
    off = OracleModel(channel_grid(96, 32), SimulationParams(dt=0.02, target_inlet_velocity=3.0), precision=64)
    for _ in range(32):
        off.update()
    assert off.get_residuals().piso_substeps == 1  # reference behaviour: substep_count stays 1 (:267)


@pytest.mark.parametrize("solver", [PressureSolver.CG, PressureSolver.MGCG])
def test_relative_stopping_rule(solver):
    """cg_relative: stop on ||r||_2 <= tol * ||rhs||_2 of the step's first solve.  The reported p_rel is that ratio, the
    absolute measure dt * rms(r) equals p_rel * rhs_rms, and re-correction solves (measured against the same
    reference) converge at once."""
    c = default_consts()
    c.cg_relative = 1
    c.cg_tolerance = 1e-7
    g = box_grid(64)
    prm = SimulationParams(dt=1e-3, viscosity=0.01, scenario=Scenario.Cavity, pressure_solver=solver)
    m = OracleModel(g, prm, precision=64, consts=c)
    for s in range(8):
        m.update()
        r = m.get_residuals()
        if s >= 2:
            assert 0 < r.f64["p_rel"] <= 1e-7 and r.f64["rhs_rms"] > 0
            assert r.f64["first_solve_iterations"] == r.sweeps > 0  # the re-correction solve took none
            assert r.jacobi_calls == 2
    a = OracleModel(g, prm, precision=64)  # the absolute rule, same flow to the solver tolerance
    for _ in range(8):
        a.update()
    assert rel_l2(m.field(_abi.FIELD_U), a.field(_abi.FIELD_U)) < 1e-5


def test_tracers_restatement_known_answers():
    """Facts read off index.html:1472-1543: tracers start on the inlet at row centres; in a uniform flow u = U, v = 0 a
    tracer moves by U * dt per update (bilinear interpolation reproduces constants exactly) and is dropped once it is
    past x = lx; in a linear shear u = y the interpolated velocity at a cell centre is that cell's value."""
    from oracle import tracers as tr
    nx, ny, lx, ly = 16, 8, 4.0, 2.0
    dx, dy = lx / nx, ly / ny
    t = tr.inject(np.zeros((0, 2)), ny, dy)
    assert t.shape == (ny, 2) and not t[:, 0].any() and np.array_equal(t[:, 1], (np.arange(ny) + 0.5) * dy)
    u = np.full((ny, nx + 1), 1.5)
    v = np.zeros((ny + 1, nx))
    for k in range(1, 4):
        t = tr.update(t, u, v, nx, ny, dx, dy, lx, ly, 0.5)
        assert t.shape[0] == ny and np.array_equal(t[:, 0], np.full(ny, 0.75 * k))
    for _ in range(3):
        t = tr.update(t, u, v, nx, ny, dx, dy, lx, ly, 0.5)
    assert t.shape[0] == 0  # 6 * 0.75 = 4.5 > lx: all gone
    ushear = np.tile(((np.arange(ny) + 0.5) * dy)[:, None], (1, nx + 1))
    x = np.array([2.125]); y = np.array([(3 + 0.5) * dy])
    ui, vi = tr.velocity_at(ushear, v, nx, ny, dx, dy, x, y)
    # (x, y) = the centre of cell (8, 3) is the lower-left node of the interpolation square: rx = ry = 0.5 between cell centres
    # (3, 4) ... the JS interpolates between the values of cells (i, j), (i+1, j), (i, j+1), (i+1, j+1) with weights from the
    # position inside cell (i, j) — at its centre that is the mean of rows j and j+1
    assert np.allclose(ui, 0.5 * ((3 + 0.5) * dy + (4 + 0.5) * dy)) and vi[0] == 0.0
