"""SURVEY 8f extensions on the GPU against the CPU oracle: QUICK face values, adaptive sub-stepping, the relative Mode C
stopping rule, and the fused smoothing passes of the V-cycle against the separate kernels."""
import numpy as np
import pytest

from cfd_demo_b200 import _abi
from cfd_demo_b200.model import Model, default_options
from cfd_demo_b200.types import (Grid, InletProfile, PressureSolver, Scenario, SimulationParams, VelocityScheme)
from oracle.cpu_oracle import OracleModel, default_consts

from helpers import STATE_FIELDS, assert_fields_identical, assert_residuals_identical, box_grid, channel_grid, rel_l2

pytestmark = pytest.mark.gpu


def pair(grid, params, precision=64, consts=None, flags=0):
    o = default_options()
    o.precision = precision
    o.flags = flags
    if consts is not None:
        o.consts = consts
    return Model(grid, params, options=o), OracleModel(grid, params, precision=precision, consts=consts)


@pytest.mark.parametrize("precision", [64, 32])
@pytest.mark.parametrize("case", ["default", "ragged", "cavity"])
def test_quick_scheme_bit_exact(precision, case):
    """VelocityScheme::Quick (index.html:471-549, :643-723 inside the Rust predictor): Mode R, every state field,
    residual and counter bit for bit, through the start-up transient into the saturated regime."""
    if case == "default":
        from cfd_demo_b200.types import default_grid
        g, prm, steps = default_grid(), SimulationParams(velocity_scheme=VelocityScheme.Quick), 22
    elif case == "ragged":
        g = channel_grid(264, 41)
        prm, steps = SimulationParams(velocity_scheme=VelocityScheme.Quick, inlet_profile=InletProfile.Parabolic,
                                      target_inlet_velocity=2.0), 14
    else:
        g, prm, steps = box_grid(64), SimulationParams(dt=5e-4, viscosity=0.01, scenario=Scenario.Cavity,
                                                       velocity_scheme=VelocityScheme.Quick), 15
    gpu, cpu = pair(g, prm, precision)
    for s in range(steps):
        gpu.update()
        cpu.update()
        assert_residuals_identical(gpu.get_residuals(), cpu.get_residuals(), f"quick {case} step {s + 1}")
    assert_fields_identical(gpu, cpu, STATE_FIELDS, f"quick {case}")
    assert np.abs(gpu.field(_abi.FIELD_U)).max() > 0


@pytest.mark.parametrize("precision", [64, 32])
def test_adaptive_substeps_bit_exact(precision):
    """cfd_solver_consts::adaptive_substeps (the reference's commented-out rule, src/model.rs:352-363): the sub-step
    count grows to its cap of 20 on this case; every step's residuals, counters and sub-step count, and the complete
    state (u_old / v_old are the fields the STEP started from, not the last sub-step) must equal the oracle's."""
    c = default_consts()
    c.adaptive_substeps = 1
    prm = SimulationParams(dt=0.02, target_inlet_velocity=3.0)
    gpu, cpu = pair(channel_grid(96, 32), prm, precision, consts=c)
    seen = set()
    for s in range(27):
        gpu.update()
        cpu.update()
        rg, rc = gpu.get_residuals(), cpu.get_residuals()
        assert_residuals_identical(rg, rc, f"substeps step {s + 1}")
        seen.add(rg.piso_substeps)
        if s % 5 == 4:
            assert_fields_identical(gpu, cpu, STATE_FIELDS, f"substeps step {s + 1}")
    assert max(seen) > 4 and 1 in seen
    assert_fields_identical(gpu, cpu, STATE_FIELDS, "substeps end")


@pytest.mark.parametrize("solver", [PressureSolver.MGCG, PressureSolver.CG])
def test_relative_stopping_rule_matches_oracle(solver):
    """cg_relative = 1: ||r||_2 <= tol * ||rhs||_2 of the step's first solve (the norm SURVEY 8d states for the 4096^2
    cavity).  Same iteration counts, the reported ||r|| / ||rhs|| and dt * rms(rhs) agree with the oracle's, u, v, p within 1e-9."""
    c = default_consts()
    c.cg_relative = 1
    c.cg_tolerance = 1e-10
    g = Grid.uniform(128, 96, 2.0, 1.5, None)
    prm = SimulationParams(dt=1e-3, viscosity=0.01, scenario=Scenario.Cavity, pressure_solver=solver)
    gpu, cpu = pair(g, prm, consts=c)
    for s in range(10):
        gpu.update()
        cpu.update()
        rg, rc = gpu.get_residuals(), cpu.get_residuals()
        assert rg.jacobi_calls == rc.jacobi_calls
        assert abs(rg.sweeps - rc.sweeps) <= (1 if solver == PressureSolver.MGCG else 4), (s, rg.sweeps, rc.sweeps)
        if s >= 2:
            assert 0 < rg.f64["p_rel"] <= 1e-10
            assert abs(rg.f64["rhs_rms"] - rc.f64["rhs_rms"]) <= 1e-9 * rc.f64["rhs_rms"]
            assert rg.f64["first_solve_iterations"] == rg.sweeps
    for fid in (_abi.FIELD_U, _abi.FIELD_V, _abi.FIELD_P):
        assert rel_l2(gpu.field(fid), cpu.field(fid)) <= 1e-9, _abi.FIELD_NAMES[fid]


LEG_SHAPES = [(Scenario.Channel, (1040, 61)), (Scenario.Cavity, (264, 200)), (Scenario.Cavity, (16, 4)),
              (Scenario.Cavity, (520, 300)), (Scenario.Channel, (136, 37)), (Scenario.Cavity, (2056, 72))]


def _leg_grid(scenario, shape):
    nx, ny = shape
    if scenario == Scenario.Channel:
        return channel_grid(nx, ny, lx=nx / 16.0, ly=ny / 16.0, cylinder=ny >= 7)
    return Grid.uniform(nx, ny, nx / 64.0, ny / 64.0, None)


@pytest.mark.parametrize("nu", [2, 3, 4])
@pytest.mark.parametrize("precision", [64, 32])
@pytest.mark.parametrize("scenario,shape", LEG_SHAPES)
def test_mgcg_vcycle_legs_are_bit_identical_to_the_separate_kernels(scenario, shape, precision, nu):
    """cfd_mg_legs.cuh (each level's descending / ascending leg of the V(2,2)-cycle as one launch) performs the same
    per-cell arithmetic as the one-operation-per-launch kernels (CFD_FLAG_MG_UNFUSED): with one CG iteration per solve
    the V-cycle's result z (CFD_FIELD_MG_Z) must be bit-identical, ring included, on widths / heights that are not
    multiples of the tile, odd level sizes, channel (zero outlet column) and cavity (mirror) boundary rules, fp64 and
    fp32, V(2,2), V(3,3) and V(4,4).  Tolerance: 0."""
    from oracle.cpu_oracle import default_consts
    g = _leg_grid(scenario, shape)
    prm = SimulationParams(dt=1e-3, viscosity=0.01, scenario=scenario, pressure_solver=PressureSolver.MGCG)
    zs = []
    for flags in (0, _abi.FLAG_MG_UNFUSED):
        o = default_options()
        o.flags = flags
        o.consts = default_consts()
        o.consts.cg_max_iterations = 1
        o.consts.outer_rounds = 0
        o.consts.mg_warm_start = 0
        o.consts.mg_smoothing = nu
        o.precision = precision
        m = Model(g, prm, options=o)
        # the first V-cycle of the run's first non-trivial solve (the driving velocity ramps up from 0: the first steps have
        # a zero right-hand side); later solves start from fields that already carry the rounding of alpha
        for _ in range(6):
            m.update()
            if m.get_residuals().sweeps > 0:
                break
        assert m.get_residuals().sweeps == 1
        zs.append(m.field(_abi.FIELD_MG_Z))
        m.close()
    assert np.abs(zs[1]).max() > 0
    assert np.array_equal(zs[0], zs[1]), f"legs vs separate kernels: {np.count_nonzero(zs[0] != zs[1])} entries differ"


@pytest.mark.parametrize("scenario,shape", LEG_SHAPES[:4])
def test_mgcg_fused_smoothing_passes_equal_the_separate_kernels(scenario, shape):
    """The fused V-cycle (cfd_mg_legs.cuh) against the separate kernels (CFD_FLAG_MG_UNFUSED) over whole solves: the
    V-cycle itself is bit-identical (test above); rho.z is summed over other tiles, so alpha / beta differ in their last
    bits and the converged fields agree to rounding: same iteration counts, every state field within 1e-10 relative L2
    (cg_tolerance 1e-12)."""
    g = _leg_grid(scenario, shape)
    prm = SimulationParams(dt=1e-3, viscosity=0.01, scenario=scenario, pressure_solver=PressureSolver.MGCG)
    models = []
    for flags in (0, _abi.FLAG_MG_UNFUSED):
        o = default_options()
        o.flags = flags
        o.consts.cg_tolerance = 1e-12
        m = Model(g, prm, options=o)
        for _ in range(6):
            m.update()
        models.append(m)
    assert models[0].get_residuals().sweeps == models[1].get_residuals().sweeps > 0
    for fid in list(STATE_FIELDS) + [_abi.FIELD_MG_GUESS, _abi.FIELD_MG_LAST, _abi.FIELD_MG_LAST2]:
        a, b = models[0].field(fid), models[1].field(fid)
        assert np.linalg.norm(a - b) <= 1e-10 * max(np.linalg.norm(b), 1e-300), _abi.FIELD_NAMES[fid]


def test_set_parameters_validates_enums():
    from cfd_demo_b200.model import CfdError
    m = Model(channel_grid(32, 8), SimulationParams())
    bad = SimulationParams()
    for field, value in (("velocity_scheme", 7), ("inlet_profile", -1), ("pressure_solver", 9)):
        p = bad.to_c()
        setattr(p, field, value)
        import ctypes as C
        assert m._lib.cfd_model_set_params(m._handle(), C.byref(p)) == _abi.CFD_ERR_INVALID_ARGUMENT, field
    m.set_parameters(SimulationParams(velocity_scheme=VelocityScheme.Quick))
    m.update()


@pytest.mark.parametrize("precision", [64, 32])
def test_tracer_particles_match_the_js_restatement(precision):
    """cfd_model_tracers_* (index.html:1472-1543) against oracle/tracers.py on the GPU model's own fields: after every
    timestep the tracers are advected with the solver's dt (the JS loop, :1233-1237), every 10 steps a new row of tracers is
    injected; positions and their order must agree bit for bit (double arithmetic, no contraction), tracers that leave
    the domain disappear on both sides."""
    from oracle import tracers as tr
    g = channel_grid(96, 32, lx=6.0, ly=2.0)
    prm = SimulationParams(dt=0.02, target_inlet_velocity=3.0, viscosity=1e-3)
    o = default_options()
    o.precision = precision
    o.consts.ramp_up_steps = 4
    gpu = Model(g, prm, options=o)
    gpu.tracers_inject()
    ref = tr.inject(np.zeros((0, 2)), g.ny, g.dy)
    assert np.array_equal(gpu.tracers(), ref)
    dropped = False
    for s in range(60):
        gpu.update()
        dt = 25.0 * gpu.get_residuals().f64["dt"]  # the tracer step is the caller's choice; long enough to cross the domain
        u, v = gpu.field(_abi.FIELD_U), gpu.field(_abi.FIELD_V)
        gpu.tracers_update(dt)
        n_before = ref.shape[0]
        ref = tr.update(ref, u, v, g.nx, g.ny, g.dx, g.dy, g.lx, g.ly, dt)
        dropped |= ref.shape[0] < n_before
        if (s + 1) % 10 == 0:
            gpu.tracers_inject()
            ref = tr.inject(ref, g.ny, g.dy)
        got = gpu.tracers()
        assert got.shape == ref.shape, (s, got.shape, ref.shape)
        assert np.array_equal(got, ref), (s, np.abs(got - ref).max())
    assert ref.shape[0] > g.ny and dropped and ref[:, 0].max() > 1.0
    gpu.tracers_clear()
    assert gpu.tracers().shape == (0, 2)


def test_asynchronous_snapshot_is_the_state_at_begin():
    """cfd_model_snapshot_begin / _end: the snapshot is the state at `begin`, although two more timesteps run before `end`
    (the fields are narrowed on the model's stream before they change; the copy overlaps the steps); two snapshots may be
    in flight, a third `begin` is refused; pageable or mis-sized destinations are refused instead of being overrun."""
    from cfd_demo_b200.model import CfdError
    g = channel_grid(264, 96)
    m = Model(g, SimulationParams())
    for _ in range(6):
        m.update()
    a, b = m.pinned_snapshot_buffers(), m.pinned_snapshot_buffers()
    want_a = [m.field(f).astype(np.float32) for f in (_abi.FIELD_P, _abi.FIELD_U, _abi.FIELD_V)]
    dt_a = m.get_residuals().dt
    m.snapshot_begin(a)
    m.update()
    want_b = [m.field(f).astype(np.float32) for f in (_abi.FIELD_P, _abi.FIELD_U, _abi.FIELD_V)]
    m.snapshot_begin(b)
    with pytest.raises(CfdError):
        m.snapshot_begin(a)
    m.update()
    sa = m.snapshot_end()
    sb = m.snapshot_end()
    for got, want in zip((sa.p, sa.u, sa.v), want_a):
        assert np.array_equal(got, want)
    for got, want in zip((sb.p, sb.u, sb.v), want_b):
        assert np.array_equal(got, want)
    assert sa.dt == dt_a and not np.array_equal(sa.u, sb.u)
    with pytest.raises(CfdError):
        m.snapshot_end()
    sizes = m.snapshot_sizes()
    with pytest.raises(CfdError):  # pageable destinations cannot be written asynchronously
        m.snapshot_begin([np.empty(n, dtype=np.float32) for n in sizes])
    with pytest.raises(CfdError):  # wrong size / dtype: refused before anything is written
        m.get_snapshot(out=[np.empty(n - 1, dtype=np.float32) for n in sizes])
    with pytest.raises(CfdError):
        m.get_snapshot(out=[np.empty(n, dtype=np.float64) for n in sizes])
    with pytest.raises(CfdError):
        m.render_rgba(0, out=np.empty(10, dtype=np.uint8))


def test_cpp_headless_driver_runs_the_headline_workload_family():
    """The benchmarked path (lid-driven cavity, MGCG, relative stopping rule) through the C++ mirror of the reference API
    (host/cfd_model.hpp -> the same C ABI the Rust shim binds): its per-step log must equal the Python mirror's run of the
    same problem — iteration counts, the relative residual and dt * rms(rhs) to the printed precision."""
    import os
    import subprocess
    exe = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "cfd_demo_b200", "host", "cfd_headless")
    n, steps, nu, dt = 512, 30, 1.0e-2, 8.0e-5
    r = subprocess.run([exe, "cavity", str(n), str(steps), repr(nu), repr(dt)], capture_output=True, text=True, check=True)
    c = default_consts()
    c.cg_relative = 1
    o = default_options()
    o.consts = c
    m = Model(box_grid(n), SimulationParams(dt=dt, viscosity=nu, scenario=Scenario.Cavity, pressure_solver=PressureSolver.MGCG), options=o)
    for _ in range(steps):
        m.update()
    res = m.get_residuals()
    line = [l for l in r.stdout.splitlines() if l.startswith(f"step {steps} ")][0]
    assert f"K={res.jacobi_calls} iterations={res.sweeps} " in line, (line, res)
    assert f"rel={res.f64['p_rel']:.3e}" in line and f"dt*rms(rhs)={res.f64['rhs_rms']:.3e}" in line, (line, res.f64)
    assert res.sweeps > 0 and res.f64["p_rel"] <= 1e-8
