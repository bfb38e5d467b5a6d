"""Known-answer tests that pin the CPU oracle (oracle/cfd_oracle.hpp) to facts read off the reference code.

The reference holds no golden vectors for src/model.rs ("parity unpinned"); these are the
answers that follow from the code itself: trivial first steps, boundary identities, the mask census of the
default grid, the early solver counters, and the in-bounds property of the wrap-around reads.
"""
import numpy as np
import pytest

from cfd_demo_b200 import _abi
from cfd_demo_b200.types import (Grid, InletProfile, Scenario, SimulationParams, VelocityScheme, default_grid)
from oracle.cpu_oracle import OracleModel, default_consts

from helpers import box_grid, channel_grid


@pytest.fixture(scope="module", autouse=True)
def _built(oracle_built):
    return oracle_built


def test_default_grid_matches_reference_default():
    g = default_grid()  # src/app.rs:33-53
    assert (g.nx, g.ny) == (800, 264)
    assert g.dx == float(np.float32(30.0) / np.float32(800))
    assert g.dy == float(np.float32(10.0) / np.float32(264))
    assert (g.obstacle.center_x, g.obstacle.center_y, g.obstacle.radius) == (7.5, 5.0, 0.75)


def test_solver_constants_are_the_reference_literals():
    c = default_consts()
    assert (c.ramp_up_steps, c.jacobi_iterations, c.outer_rounds) == (100, 50, 20)  # :269 :737 :696
    assert (c.jacobi_omega, c.pressure_tolerance, c.outer_tolerance, c.cfl) == (0.75, 1e-4, 1e-4, 0.2)
    # the f32 build rounds these to the reference's f32 literals
    assert np.float32(c.pressure_tolerance) == np.float32("1e-4") and np.float32(c.cfl) == np.float32("0.2")


@pytest.mark.parametrize("precision", [32, 64])
def test_mask_census_default_grid(precision):
    # pure function of Grid (src/model.rs:236-259); SURVEY §8c KAT 3
    m = OracleModel(default_grid(), SimulationParams(), precision=precision)
    assert m.obstacle_count() == 1248
    mu, mv = m.field(_abi.FIELD_MASK_U), m.field(_abi.FIELD_MASK_V)
    assert int(mu.sum()) == 1288 and int(mv.sum()) == 1288
    assert mu.reshape(264, 801)[:, 0].sum() == 0  # i > 0 guard (:245)


@pytest.mark.parametrize("precision", [32, 64])
def test_first_two_steps_are_trivial(precision):
    # SURVEY §8c KAT 1: ramp gives 0 at step 0; after the second update only the inlet column is non-zero
    g = default_grid()
    m = OracleModel(g, SimulationParams(), precision=precision)
    m.update()
    r = m.get_residuals()
    assert (r.simulation_step, r.jacobi_calls, r.sweeps) == (1, 2, 2)
    for fid in (_abi.FIELD_P, _abi.FIELD_U, _abi.FIELD_V):
        assert not m.field(fid).any()
    assert r.f64["u"] == 0.0 and r.f64["p"] == 0.0
    m.update()
    r = m.get_residuals()
    u = m.field(_abi.FIELD_U).reshape(g.ny, g.nx + 1)
    expect = float(np.float32(1) / np.float32(100)) if precision == 32 else (1.0 / 100.0) * 1.0
    assert (u[1:-1, 0] == expect).all() and u[0, 0] == 0 and u[-1, 0] == 0
    u[:, 0] = 0
    assert not u.any() and not m.field(_abi.FIELD_P).any() and not m.field(_abi.FIELD_V).any()
    assert r.f64["u"] == expect and (r.jacobi_calls, r.sweeps) == (2, 2)
    assert r.f64["dt"] == float(np.float32(0.005))  # CFL bound 0.2*0.0375/0.01 >> dt


def test_early_solver_counters_default_grid_f32():
    # SURVEY §8c KAT 4 (the survey's throw-away f32 emulation): (K,S) = (2,2) for steps 1-4, S=39 at step 6
    m = OracleModel(default_grid(), SimulationParams(), precision=32)
    ks = []
    for _ in range(8):
        m.update()
        r = m.get_residuals()
        ks.append((r.jacobi_calls, r.sweeps))
    assert ks[:4] == [(2, 2)] * 4
    assert ks[5] == (2, 39)


@pytest.mark.parametrize("scheme", [VelocityScheme.FirstOrder, VelocityScheme.SecondOrder])
@pytest.mark.parametrize("profile", [InletProfile.Uniform, InletProfile.Parabolic])
def test_boundary_identities_and_in_bounds_reads(scheme, profile):
    # SURVEY §8c KAT 2, on the bounds-checked build: any out-of-range flat index would abort the process
    g = channel_grid(64, 24)
    prm = SimulationParams(velocity_scheme=scheme, inlet_profile=profile)
    m = OracleModel(g, prm, precision=64, checked=True)
    for _ in range(12):
        m.update()
    nx, ny = g.nx, g.ny
    u = m.field(_abi.FIELD_U).reshape(ny, nx + 1)
    v = m.field(_abi.FIELD_V).reshape(ny + 1, nx)
    pp = m.field(_abi.FIELD_P_PRIME).reshape(ny, nx)
    assert (u[:, nx] == u[:, nx - 1]).all()
    assert not u[0].any() and not u[-1].any() and not v[0].any() and not v[-1].any()
    assert (pp[:, 0] == pp[:, 1]).all() and not pp[:, nx - 1].any()
    assert (pp[0] == pp[1]).all() and (pp[-1] == pp[-2]).all()
    assert np.isfinite(u).all() and np.abs(u).max() > 0
    if profile == InletProfile.Parabolic:
        inlet = u[1:-1, 0]
        assert inlet.argmax() in (ny // 2 - 1, ny // 2 - 2) and inlet.min() >= 0


def test_solid_faces_are_zero_after_a_step():
    g = channel_grid(64, 24)
    m = OracleModel(g, SimulationParams(), precision=64)
    for _ in range(10):
        m.update()
    nx, ny = g.nx, g.ny
    u = m.field(_abi.FIELD_U).reshape(ny, nx + 1)
    v = m.field(_abi.FIELD_V).reshape(ny + 1, nx)
    mu = m.field(_abi.FIELD_MASK_U).reshape(ny, nx + 1)
    assert m.obstacle_count() > 0
    # a cell is solid iff both its u faces ... cheaper: recompute the solid set as the reference does (:238-243)
    f = np.float32
    solid = []
    for j in range(ny):
        for i in range(nx):
            x, y = (f(i) + f(0.5)) * f(g.dx), (f(j) + f(0.5)) * f(g.dy)
            ddx, ddy = x - f(g.obstacle.center_x), y - f(g.obstacle.center_y)
            if np.sqrt(ddx * ddx + ddy * ddy) < f(g.obstacle.radius):
                solid.append((i, j))
    assert len(solid) == m.obstacle_count()
    for i, j in solid:
        assert u[j, i] == 0 and v[j, i] == 0 and mu[j, i + 1] == 1


def test_float_and_double_oracles_agree_loosely():
    g = channel_grid(64, 24)
    a = OracleModel(g, SimulationParams(), precision=32)
    b = OracleModel(g, SimulationParams(), precision=64)
    for _ in range(5):
        a.update()
        b.update()
    ua, ub = a.field(_abi.FIELD_U), b.field(_abi.FIELD_U)
    assert np.abs(ua - ub).max() < 1e-5 and np.abs(ub).max() > 1e-3


def test_invalid_grid_rejected():
    # the reference panics unless nx % 8 in {0, 1}; this build takes multiples of 8 only
    with pytest.raises(ValueError):
        OracleModel(Grid.uniform(20, 8, 1.0, 1.0), SimulationParams())


def test_set_parameters_changes_only_the_six_parameters():
    g = channel_grid(32, 16, cylinder=False)
    m = OracleModel(g, SimulationParams(), precision=64)
    m.update(); m.update()
    m.set_parameters(SimulationParams(dt=0.001, viscosity=0.01, target_inlet_velocity=2.0,
                                      velocity_scheme=VelocityScheme.SecondOrder,
                                      inlet_profile=InletProfile.Parabolic))
    assert m.get_residuals().f64["dt"] == float(np.float32(0.001))
    m.update()
    assert m.current_inlet_velocity() == (2.0 / 100.0) * float(np.float32(2.0))


def test_cavity_extension_is_closed_and_compatible():
    # extension: ring cells solid for the masks, lid on the top u row; the interior block has zero net flux
    g = box_grid(32)
    prm = SimulationParams(dt=1e-3, viscosity=0.01, scenario=Scenario.Cavity)
    m = OracleModel(g, prm, precision=64, checked=True)
    for _ in range(20):
        m.update()
    n = 32
    u = m.field(_abi.FIELD_U).reshape(n, n + 1)
    v = m.field(_abi.FIELD_V).reshape(n + 1, n)
    assert (u[n - 1, 1:n] == m.current_inlet_velocity()).all()
    assert not u[:n - 1, 1].any() and not u[:n - 1, n - 1].any() and not v[1].any() and not v[n - 1].any()
    rhs = m.field(_abi.FIELD_RHS).reshape(n, n)
    interior = rhs[1:n - 1, 1:n - 1]
    assert abs(interior.sum()) < 1e-6 * np.abs(interior).sum() + 1e-12
    assert np.abs(u[1:n - 1, 2:n - 1]).max() > 0


@pytest.mark.parametrize("scenario", [Scenario.Channel, Scenario.Cavity])
def test_mode_c_cg_leaves_a_divergence_free_interior(scenario):
    """Extension (no reference counterpart): the CG solve drives the divergence of the corrected velocity on the
    unknown cells below its tolerance, and the outer loop then stops after one re-correction (K = 2)."""
    from cfd_demo_b200.types import PressureSolver
    n = 32
    g = box_grid(n) if scenario == Scenario.Cavity else channel_grid(n, n, cylinder=False)
    consts = default_consts()
    consts.cg_tolerance = 1e-10
    prm = SimulationParams(dt=1e-3, viscosity=0.01, scenario=scenario, pressure_solver=PressureSolver.CG)
    m = OracleModel(g, prm, precision=64, consts=consts)
    for _ in range(12):
        m.update()
    r = m.get_residuals()
    assert r.jacobi_calls == 2 and 0 < r.sweeps < 400 and r.f64["p"] <= 1e-10
    u = m.field(_abi.FIELD_U).reshape(n, n + 1)
    v = m.field(_abi.FIELD_V).reshape(n + 1, n)
    # velocities BEFORE the boundary update are what the solve made divergence free; check rows/cols away from it
    div = (u[2:n - 2, 3:n - 1] - u[2:n - 2, 2:n - 2]) / g.dx + (v[3:n - 1, 2:n - 2] - v[2:n - 2, 2:n - 2]) / g.dy
    assert np.abs(div).max() < 1e-6 and np.abs(u).max() > 1e-3


@pytest.mark.parametrize("scenario", [Scenario.Channel, Scenario.Cavity])
@pytest.mark.parametrize("shape", [(64, 64), (48, 37), (40, 6)])
def test_mode_c_mgcg_agrees_with_cg_and_needs_few_iterations(scenario, shape):
    """Extension: multigrid-preconditioned CG solves the same discrete problem as CG (velocities agree to the
    solver tolerance; the all-Neumann cavity pressure is defined up to a constant) in an (almost) grid-independent
    handful of iterations, including odd unknown counts (trailing single-cell aggregates) and thin grids."""
    from cfd_demo_b200.types import PressureSolver
    nx, ny = shape
    g = Grid.uniform(nx, ny, nx / 64.0, ny / 64.0, None) if scenario == Scenario.Cavity else channel_grid(nx, ny, lx=nx / 16.0, ly=ny / 16.0, cylinder=nx >= 48)
    consts = default_consts()
    consts.cg_tolerance = 1e-12
    out = {}
    for solver in (PressureSolver.CG, PressureSolver.MGCG):
        prm = SimulationParams(dt=1e-3, viscosity=0.01, scenario=scenario, pressure_solver=solver)
        m = OracleModel(g, prm, precision=64, consts=consts)
        its = 0
        for _ in range(8):
            m.update()
            r = m.get_residuals()
            assert r.jacobi_calls == 2 and r.f64["p"] <= 1e-12
            its = max(its, r.sweeps)
        out[solver] = (m.field(_abi.FIELD_U), m.field(_abi.FIELD_V), m.field(_abi.FIELD_P), its)
    cg, mg = out[PressureSolver.CG], out[PressureSolver.MGCG]
    assert 0 < mg[3] <= 16 and mg[3] < cg[3], (mg[3], cg[3])
    assert np.abs(cg[0]).max() > 1e-4
    for k in range(2):
        assert np.linalg.norm(cg[k] - mg[k]) <= 1e-8 * np.linalg.norm(cg[0]), k
    dp = (cg[2] - mg[2]).reshape(ny, nx)[1:ny - 1, 1:nx - 1]
    if scenario == Scenario.Cavity:
        dp = dp - dp.mean()
    assert np.abs(dp).max() <= 1e-7 * max(np.abs(cg[2]).max(), 1e-30) + 1e-12


@pytest.mark.parametrize("precision", [32, 64])
@pytest.mark.parametrize("case", ["default_grid", "parabolic_small", "second_order_default_grid", "second_order_small"])
def test_cpp_oracle_agrees_bit_for_bit_with_the_independent_numpy_restatement(precision, case):
    """Two restatements of src/model.rs written independently and structured differently (C++: 8-lane chunks and
    scalar tails, loop by loop; numpy: whole rows, straight from the Rust) must produce identical bits: every state
    field, the residuals and the solver counters, through the start-up transient into the regime where the Jacobi
    solves saturate.  (Both remain unpinned against the Rust reference itself.)"""
    from oracle.numpy_restatement import NumpyModel
    if case == "default_grid":
        g, prm, steps = default_grid(), SimulationParams(), 11
    elif case == "second_order_default_grid":
        g, prm, steps = default_grid(), SimulationParams(velocity_scheme=VelocityScheme.SecondOrder), 11
    elif case == "second_order_small":
        g = channel_grid(48, 19)
        prm = SimulationParams(velocity_scheme=VelocityScheme.SecondOrder, inlet_profile=InletProfile.Parabolic, dt=0.02,
                               viscosity=1e-3)
        steps = 60
    else:
        g = channel_grid(40, 21)
        prm, steps = SimulationParams(inlet_profile=InletProfile.Parabolic, dt=0.02, viscosity=1e-3), 60
    dtype = np.float32 if precision == 32 else np.float64
    a = OracleModel(g, prm, precision=precision)
    b = NumpyModel(g, prm, dtype=dtype)
    saturated = False
    for s in range(steps):
        a.update()
        b.update()
        r = a.get_residuals()
        assert (r.jacobi_calls, r.sweeps) == (b.K, b.S), (s, r.jacobi_calls, r.sweeps, b.K, b.S)
        saturated |= r.sweeps >= 200
        for fid, arr in ((_abi.FIELD_U, b.u), (_abi.FIELD_V, b.v), (_abi.FIELD_P, b.p), (_abi.FIELD_U_STAR, b.u_star),
                         (_abi.FIELD_V_STAR, b.v_star), (_abi.FIELD_RHS, b.rhs), (_abi.FIELD_P_PRIME, b.pp)):
            x = a.field(fid)
            y = arr.astype(np.float64)
            same = (x == y) | (np.isnan(x) & np.isnan(y))
            assert same.all(), (s, _abi.FIELD_NAMES[fid], int((~same).sum()), float(np.nanmax(np.abs(x - y))))
        for k, val in (("dt", b.dt), ("p", b.last_p), ("u", b.last_u), ("v", b.last_v), ("simulation_time", b.time)):
            assert r.f64[k] == float(val), (s, k, r.f64[k], float(val))
    assert saturated and np.abs(b.u).max() > 0


def test_mgcg_start_vector_modes_converge_to_the_same_flow():
    """mg_warm_start only changes where the first solve of a step STARTS (0 cold, 1 previous p', 2 linear and 3
    quadratic extrapolation in time): every mode must reach the same velocities within the solver tolerance, and the
    extrapolated starts must need fewer iterations than the cold start once the flow evolves smoothly."""
    from cfd_demo_b200.types import PressureSolver
    n = 96
    g = Grid.uniform(n, n, 1.0, 1.0, None)
    nu = 1e-3
    prm = SimulationParams(dt=0.05 * (1.0 / n) ** 2 / nu, viscosity=nu, scenario=Scenario.Cavity,
                           pressure_solver=PressureSolver.MGCG)
    fields, iters = {}, {}
    for mode in (0, 1, 2, 3):
        c = default_consts()
        c.ramp_up_steps = 5
        c.mg_warm_start = mode
        m = OracleModel(g, prm, precision=64, consts=c)
        total = 0
        for s in range(40):
            m.update()
            r = m.get_residuals()
            assert r.jacobi_calls == 2 and r.f64["p"] <= 1e-8
            if s >= 20:
                total += r.sweeps
        fields[mode], iters[mode] = m.field(_abi.FIELD_U), total
    for mode in (1, 2, 3):
        assert np.linalg.norm(fields[mode] - fields[0]) <= 1e-5 * np.linalg.norm(fields[0]), mode
    assert iters[3] <= iters[2] <= iters[1] < iters[0], iters
