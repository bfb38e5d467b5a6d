"""BASELINE.json configs[1] and configs[3] on the GPU against the CPU oracle, and the state semantics of the Mode C
fast path's elided work (zero-iteration re-correction rounds, rotating start-vector history).

configs[1]: lid-driven cavity Re = 100 on 1024 x 1024, Mode R (bit-exact) and Mode C / MGCG (tolerance), SURVEY 8(d) row 2.
configs[3]: channel past a masked cylinder, 8192 x 2048 (src/app.rs:33-53 scaled; BCs src/model.rs:827-875; masks
:236-259): one dense saturated timestep per velocity scheme from the GPU's own state, bit-exact; the 2/4/8-strip runs
of the same grid are checked by tests/mgpu_strip_check.py (needs >= 2 GPUs) and by the `parity` block of every
`bench.py --gpus N > 1` line.
"""
import numpy as np
import pytest

from cfd_demo_b200 import _abi
from cfd_demo_b200.model import Model, default_options
from cfd_demo_b200.types import (Cylinder, Grid, PressureSolver, Scenario, SimulationParams, VelocityScheme)
from oracle.cpu_oracle import OracleModel, default_consts

from helpers import STATE_FIELDS, assert_fields_identical, assert_residuals_identical, box_grid, channel_grid, rel_l2

pytestmark = pytest.mark.gpu

RESTART_FIELDS = (_abi.FIELD_P, _abi.FIELD_U, _abi.FIELD_V, _abi.FIELD_U_STAR, _abi.FIELD_V_STAR, _abi.FIELD_P_PRIME)
MG_FIELDS = (_abi.FIELD_MG_GUESS, _abi.FIELD_MG_LAST, _abi.FIELD_MG_LAST2)


def cavity1024():
    # nu = 0.01 (Re 100), dt = 2e-5 (explicit diffusion limit dx^2 / 4 nu = 2.4e-5)
    return Grid.uniform(1024, 1024, 1.0, 1.0, None), dict(dt=2.0e-5, viscosity=1.0e-2, target_inlet_velocity=1.0,
                                                          scenario=Scenario.Cavity)


def test_config1_cavity1024_mode_r_bit_exact():
    """Mode R, 40 timesteps from rest: through the start-up transient ((K,S) = (2,2), (5,231), ...) into the saturated
    regime (21, 1050) that every later step repeats.  SURVEY asks for N = 200; the oracle needs 3.1 s per saturated
    step at this size (10 minutes for 200), so the test stops at 40 (100 s of oracle) — every residual and counter after
    every step, every state field every 10 steps, bit for bit."""
    g, kw = cavity1024()
    prm = SimulationParams(**kw)
    gpu, cpu = Model(g, prm), OracleModel(g, prm, precision=64)
    for s in range(40):
        gpu.update()
        cpu.update()
        assert_residuals_identical(gpu.get_residuals(), cpu.get_residuals(), f"cavity1024 Mode R step {s + 1}")
        if (s + 1) % 10 == 0:
            assert_fields_identical(gpu, cpu, STATE_FIELDS, f"cavity1024 Mode R step {s + 1}")
    r = gpu.get_residuals()
    assert (r.jacobi_calls, r.sweeps) == (21, 1050)
    for fid in (_abi.FIELD_U, _abi.FIELD_V, _abi.FIELD_P):
        assert rel_l2(gpu.field(fid), cpu.field(fid)) <= 1e-9  # north_star tolerance (met exactly)


def test_config1_cavity1024_mode_c_200_steps_at_the_benchmarked_tolerance():
    """Mode C (MGCG) at the SHIPPED stopping tolerance (cg_tolerance 1e-8, what bench.py runs), N = 200 timesteps.
    The dot products are summed in another order than the oracle's, so a solve whose stopping measure lands within
    rounding of the tolerance may take one iteration more or less on one side; both results are then converged to
    the tolerance but differ by about that much.  Asserted: the same iteration count on at least 97 % of the steps,
    never more than one apart, and u, v, p within 1e-7 relative L2 after 200 steps (1e-9 is asserted below, where
    both sides converge to the same, much smaller residual)."""
    g, kw = cavity1024()
    prm = SimulationParams(pressure_solver=PressureSolver.MGCG, **kw)
    gpu, cpu = Model(g, prm), OracleModel(g, prm, precision=64)
    differ = 0
    for s in range(200):
        gpu.update()
        cpu.update()
        rg, rc = gpu.get_residuals(), cpu.get_residuals()
        assert rg.jacobi_calls == rc.jacobi_calls == 2, s
        assert abs(rg.sweeps - rc.sweeps) <= 1 and rg.f64["p"] <= 1e-8, (s, rg.sweeps, rc.sweeps, rg.f64["p"])
        differ += rg.sweeps != rc.sweeps
    assert differ <= 6, differ
    for fid in (_abi.FIELD_U, _abi.FIELD_V, _abi.FIELD_P):
        e = rel_l2(gpu.field(fid), cpu.field(fid))
        assert e <= 1e-7, (_abi.FIELD_NAMES[fid], e, differ)


def test_config1_cavity1024_mode_c_converged_to_the_same_residual():
    """north_star's bar proper: both sides converge every Poisson solve to the same (tiny) residual, dt*rms(r) <= 1e-12;
    60 timesteps through the start of the lid ramp; identical iteration counts, u, v, p within 1e-9 relative L2."""
    g, kw = cavity1024()
    prm = SimulationParams(pressure_solver=PressureSolver.MGCG, **kw)
    consts = default_consts()
    consts.cg_tolerance = 1e-12
    o = default_options()
    o.consts = consts
    gpu, cpu = Model(g, prm, options=o), OracleModel(g, prm, precision=64, consts=consts)
    for s in range(60):
        gpu.update()
        cpu.update()
        rg, rc = gpu.get_residuals(), cpu.get_residuals()
        assert abs(rg.sweeps - rc.sweeps) <= 1 and rg.f64["p"] <= 1e-12, (s, rg.sweeps, rc.sweeps)
    for fid in (_abi.FIELD_U, _abi.FIELD_V, _abi.FIELD_P):
        assert rel_l2(gpu.field(fid), cpu.field(fid)) <= 1e-9, _abi.FIELD_NAMES[fid]


@pytest.mark.parametrize("scheme", [VelocityScheme.FirstOrder, VelocityScheme.SecondOrder])
def test_config3_channel_8192x2048_saturated_step_bit_exact(scheme):
    """BASELINE configs[3] on one GPU: 8192 x 2048 cells, 40 x 10 domain, cylinder (10, 5, r 0.75), reference defaults
    (dt 0.005, nu 1e-6, U 1, uniform inlet, Jacobi).  The GPU runs the flow into the saturated regime (K = 21,
    S = 1050 every step); its complete state goes into the CPU oracle and ONE more timestep on both sides — 1.8e10 cell
    updates through every kernel of the hot path, cylinder masks and solid-face boundary conditions included — must
    agree bit for bit (the oracle needs about a minute for that step)."""
    g = Grid.uniform(8192, 2048, 40.0, 10.0, Cylinder(10.0, 5.0, 0.75))
    prm = SimulationParams(velocity_scheme=scheme)
    gpu = Model(g, prm)
    r0 = None
    for _ in range(60):
        gpu.update()
        r0 = gpu.get_residuals()
        if (r0.jacobi_calls, r0.sweeps) == (21, 1050) and r0.simulation_step >= 26:
            break
    assert (r0.jacobi_calls, r0.sweeps) == (21, 1050), r0
    cpu = OracleModel(g, prm, precision=64)
    assert_fields_identical(gpu, cpu, [_abi.FIELD_MASK_U, _abi.FIELD_MASK_V], "masks 8192x2048")
    for fid in RESTART_FIELDS:
        cpu.set_field(fid, gpu.field(fid))
    cpu.set_scalars(r0.simulation_step, r0.f64["simulation_time"], r0.f64["dt"])
    gpu.update()
    cpu.update()
    assert_residuals_identical(gpu.get_residuals(), cpu.get_residuals(), "8192x2048 saturated step")
    assert_fields_identical(gpu, cpu, STATE_FIELDS, "8192x2048 saturated step")
    u = gpu.field(_abi.FIELD_U).reshape(2048, 8193)
    mask = gpu.field(_abi.FIELD_MASK_U).reshape(2048, 8193)
    # (masked faces are not all zero in u: the corrector does not mask and the boundary conditions zero only the west /
    # south faces of solid cells, src/model.rs:869-874 — the oracle agrees bit for bit, above)
    assert mask.sum() > 1000 and np.abs(u).max() > 0.2  # the inlet is still ramping up (step ~27 of 100, src/model.rs:311-316)


# ---- state semantics of the elided work (Mode C fast path) ------------------------------------------------------------

@pytest.mark.parametrize("scenario,shape", [(Scenario.Channel, (136, 61)), (Scenario.Cavity, (96, 96))])
def test_mgcg_complete_state_after_elided_recorrection_rounds(scenario, shape):
    """A re-correction round whose solve is converged before its first iteration has p' == 0; the CUDA path then skips
    that solve's set-up pass and the (identity) corrector, and keeps `u_star <- u` as an alias.  The COMPLETE state must
    still read back like the oracle's after every step: u_star / v_star (copies taken before the boundary conditions,
    src/model.rs:698-699 vs :728, solid faces included), p' (all zeros), rhs, u_old / v_old."""
    nx, ny = shape
    g = channel_grid(nx, ny, lx=nx / 16.0, ly=ny / 16.0, cylinder=True) if scenario == Scenario.Channel else box_grid(nx, ny)
    prm = SimulationParams(dt=1e-3, viscosity=0.01, scenario=scenario, pressure_solver=PressureSolver.MGCG,
                           velocity_scheme=VelocityScheme.SecondOrder)
    consts = default_consts()
    consts.cg_tolerance = 1e-11
    o = default_options()
    o.consts = consts
    gpu, cpu = Model(g, prm, options=o), OracleModel(g, prm, precision=64, consts=consts)
    elided = 0
    for s in range(14):
        gpu.update()
        cpu.update()
        rg, rc = gpu.get_residuals(), cpu.get_residuals()
        assert rg.jacobi_calls == rc.jacobi_calls and abs(rg.sweeps - rc.sweeps) <= 1, (s, rg, rc)
        pp = gpu.field(_abi.FIELD_P_PRIME)
        if not cpu.field(_abi.FIELD_P_PRIME).any():
            elided += 1
            assert not pp.any(), s
        if s % 3 == 2 or s == 13:  # reading u_star materialises the alias: alternate steps with and without
            for fid in STATE_FIELDS:
                if fid == _abi.FIELD_RHS:
                    continue  # the divergence left after the correction: rounding noise on both sides, nothing to compare
                a, b = gpu.field(fid), cpu.field(fid)
                assert a.shape == b.shape
                # relative to the field, or to the velocity scale where a field is still (numerically) zero
                scale = max(np.linalg.norm(b), 1e-6 * np.linalg.norm(cpu.field(_abi.FIELD_U)))
                assert np.linalg.norm(a - b) <= 1e-9 * scale, (s, _abi.FIELD_NAMES[fid], np.linalg.norm(a - b), np.linalg.norm(b))
            # the carried entries of u_star are copies taken BEFORE the boundary conditions: column 0 holds the
            # previous inlet value (exact), rows 0 / ny-1 what the corrector left there (not the zeros the BCs write into u)
            us_g, us_c = gpu.field(_abi.FIELD_U_STAR).reshape(ny, nx + 1), cpu.field(_abi.FIELD_U_STAR).reshape(ny, nx + 1)
            assert np.array_equal(us_g[:, 0], us_c[:, 0])
            assert np.allclose(us_g[0], us_c[0], rtol=1e-7, atol=1e-18) and np.allclose(us_g[-1], us_c[-1], rtol=1e-7, atol=1e-18)
            if scenario == Scenario.Channel and s >= 5:
                assert np.abs(us_c[0]).max() > 0  # i.e. the test would notice a star buffer filled from the post-BC fields
    assert elided >= 8, elided


def test_mgcg_restart_from_the_oracles_state():
    """The reverse of test_mgcg_warm_start_state_restarts_in_the_oracle: the ORACLE's complete state, start-vector
    history included, is written into a fresh GPU model through cfd_model_set_field_f64 (an explicitly set start
    vector is independent state, like the oracle's mg_guess), and both sides compute three more steps."""
    n = 192
    g = Grid.uniform(n, n, 1.0, 1.0, None)
    nu = 1e-3
    prm = SimulationParams(dt=0.02 * (1.0 / n) ** 2 / nu, viscosity=nu, scenario=Scenario.Cavity,
                           pressure_solver=PressureSolver.MGCG)
    consts = default_consts()
    consts.ramp_up_steps = 5
    consts.cg_tolerance = 1e-11
    o = default_options()
    o.consts = consts
    src = OracleModel(g, prm, precision=64, consts=consts)
    for _ in range(25):
        src.update()
    # both sides: a fresh model stepped past the (shortened) ramp, then every state field overwritten with the source's
    gpu, cpu = Model(g, prm, options=o), OracleModel(g, prm, precision=64, consts=consts)
    for _ in range(6):
        gpu.update()
        cpu.update()
    for fid in RESTART_FIELDS + MG_FIELDS:
        x = src.field(fid)
        gpu.set_field(fid, x)
        cpu.set_field(fid, x)
    for fid in RESTART_FIELDS + MG_FIELDS:
        assert np.array_equal(gpu.field(fid), src.field(fid)), _abi.FIELD_NAMES[fid]
    for s in range(3):
        gpu.update()
        cpu.update()
        rg, rc = gpu.get_residuals(), cpu.get_residuals()
        assert rg.sweeps == rc.sweeps and rg.jacobi_calls == rc.jacobi_calls, (s, rg.sweeps, rc.sweeps)
        assert 0 < rg.sweeps <= 6  # warm start: far fewer than the ~10 iterations of a cold start
    for fid in (_abi.FIELD_U, _abi.FIELD_V, _abi.FIELD_P) + MG_FIELDS:
        assert rel_l2(gpu.field(fid), cpu.field(fid)) <= 1e-9, _abi.FIELD_NAMES[fid]


def test_mgcg_switch_to_jacobi_mid_run_keeps_the_state():
    """set_parameters may change the pressure solver between steps (src/model.rs:1250-1257).  After MGCG steps whose
    last round was elided (p' logically zero, start-vector rotation pending), a Jacobi step must warm-start from the
    same p' as the oracle's (:734-824 never resets it)."""
    g = channel_grid(136, 61, lx=8.5, ly=3.8, cylinder=True)
    a = SimulationParams(dt=1e-3, viscosity=0.01, pressure_solver=PressureSolver.MGCG)
    b = SimulationParams(dt=1e-3, viscosity=0.01, pressure_solver=PressureSolver.Jacobi)
    consts = default_consts()
    consts.cg_tolerance = 1e-12
    o = default_options()
    o.consts = consts
    gpu, cpu = Model(g, a, options=o), OracleModel(g, a, precision=64, consts=consts)
    for _ in range(6):
        gpu.update()
        cpu.update()
    gpu.set_parameters(b)
    cpu.set_parameters(b)
    for s in range(4):
        gpu.update()
        cpu.update()
        rg, rc = gpu.get_residuals(), cpu.get_residuals()
        assert (rg.jacobi_calls, rg.sweeps) == (rc.jacobi_calls, rc.sweeps), (s, rg, rc)
    gpu.set_parameters(a)
    cpu.set_parameters(a)
    for _ in range(4):
        gpu.update()
        cpu.update()
    assert gpu.get_residuals().sweeps == cpu.get_residuals().sweeps
    for fid in (_abi.FIELD_U, _abi.FIELD_V, _abi.FIELD_P):
        assert rel_l2(gpu.field(fid), cpu.field(fid)) <= 1e-9, _abi.FIELD_NAMES[fid]
