"""Parity tests proper: the CUDA path, called through the C ABI, against the CPU oracle on identical inputs.

Mode R (the reference's damped Jacobi + outer loop) has no order-dependent reduction, so the bar
is BIT-EXACT agreement of every state field, residual and solver counter — fp64 against oracle<double>, and
the fp32 build against oracle<float> (the reference's own arithmetic).  The north_star's rel-L2 <= 1e-9
bound is therefore met with margin; it is asserted explicitly as well.
"""
import numpy as np
import pytest

from cfd_demo_b200 import _abi
from cfd_demo_b200.model import CfdError, Model
from cfd_demo_b200.types import (Grid, InletProfile, Scenario, SimulationParams, VelocityScheme, default_grid)
from oracle.cpu_oracle import OracleModel

from helpers import (STATE_FIELDS, assert_fields_identical, assert_residuals_identical, box_grid, channel_grid, rel_l2)

pytestmark = pytest.mark.gpu


def run_pair(grid, params, precision, steps, check_every=1, mid=None):
    gpu = Model(grid, params, precision=precision)
    cpu = OracleModel(grid, params, precision=precision)
    # masks first (Model::new)
    assert_fields_identical(gpu, cpu, [_abi.FIELD_MASK_U, _abi.FIELD_MASK_V], "masks")
    for s in range(steps):
        if mid is not None and s == mid[0]:
            gpu.set_parameters(mid[1])
            cpu.set_parameters(mid[1])
        gpu.update()
        cpu.update()
        ctx = f"precision {precision} step {s + 1}"
        assert_residuals_identical(gpu.get_residuals(), cpu.get_residuals(), ctx)
        if (s + 1) % check_every == 0 or s == steps - 1:
            assert_fields_identical(gpu, cpu, STATE_FIELDS, ctx)
    return gpu, cpu


@pytest.mark.parametrize("precision", [64, 32])
@pytest.mark.parametrize("scheme", [VelocityScheme.FirstOrder, VelocityScheme.SecondOrder])
def test_default_scenario_bit_exact(precision, scheme):
    """BASELINE config 1: default_grid() 800x264 + cylinder, SimulationParams::default(), through the
    transition from (K,S)=(2,2) to the saturated (21,1050) regime."""
    steps = 24
    gpu, cpu = run_pair(default_grid(), SimulationParams(velocity_scheme=scheme), precision, steps, check_every=6)
    r = gpu.get_residuals()
    assert (r.jacobi_calls, r.sweeps) == (21, 1050)
    for fid in (_abi.FIELD_U, _abi.FIELD_V, _abi.FIELD_P):
        assert rel_l2(gpu.field(fid), cpu.field(fid)) <= 1e-9  # north_star tolerance (met exactly)


@pytest.mark.parametrize("precision", [64, 32])
@pytest.mark.parametrize("nx,ny", [(16, 4), (24, 7), (64, 33), (264, 40), (512, 9)])
def test_small_and_ragged_grids(precision, nx, ny):
    """Minimum size, widths that are not multiples of the block width, odd heights."""
    prm = SimulationParams(velocity_scheme=VelocityScheme.SecondOrder, inlet_profile=InletProfile.Parabolic)
    run_pair(channel_grid(nx, ny, cylinder=ny >= 7), prm, precision, 12)


def test_parabolic_first_order_no_cylinder():
    prm = SimulationParams(inlet_profile=InletProfile.Parabolic, viscosity=1e-3, target_inlet_velocity=2.0)
    run_pair(channel_grid(128, 48, cylinder=False), prm, 64, 15)


def test_set_parameters_mid_run():
    new = SimulationParams(dt=0.002, viscosity=1e-4, target_inlet_velocity=1.5,
                           velocity_scheme=VelocityScheme.SecondOrder, inlet_profile=InletProfile.Parabolic)
    run_pair(channel_grid(96, 32), SimulationParams(), 64, 10, mid=(4, new))


def test_cfl_limiter_shrinks_dt():
    """A large dt trips compute_automatic_time_step (src/model.rs:878-889) on both sides identically."""
    prm = SimulationParams(dt=0.05, target_inlet_velocity=4.0)
    gpu, cpu = run_pair(channel_grid(64, 24), prm, 64, 40, check_every=10)
    assert gpu.get_residuals().f64["dt"] < float(np.float32(0.05))


@pytest.mark.parametrize("precision", [64, 32])
def test_cavity_extension_bit_exact(precision):
    prm = SimulationParams(dt=5e-4, viscosity=0.01, scenario=Scenario.Cavity)
    gpu, cpu = run_pair(box_grid(64), prm, precision, 15, check_every=5)
    n = 64
    u = gpu.field(_abi.FIELD_U).reshape(n, n + 1)
    assert (u[n - 1, 1:n] == cpu.current_inlet_velocity()).all() and np.abs(u[1:n - 1]).max() > 0


def test_snapshot_is_f32_of_the_state():
    g = channel_grid(64, 24)
    gpu = Model(g, SimulationParams())
    for _ in range(5):
        gpu.update()
    s = gpu.get_snapshot()
    assert s.p.dtype == np.float32 and s.u.shape == ((g.nx + 1) * g.ny,) and s.v.shape == (g.nx * (g.ny + 1),)
    for arr, fid in ((s.p, _abi.FIELD_P), (s.u, _abi.FIELD_U), (s.v, _abi.FIELD_V)):
        assert np.array_equal(arr, gpu.field(fid).astype(np.float32))
    assert s.dt == gpu.get_residuals().dt


def test_set_field_round_trip_and_restart():
    """State written through the ABI restarts bit-identically (u_star, v_star, p_prime are state: SURVEY N6)."""
    g = channel_grid(64, 24)
    a = Model(g, SimulationParams())
    for _ in range(9):
        a.update()
    cpu = OracleModel(g, SimulationParams())
    for _ in range(9):
        cpu.update()
    b = OracleModel(g, SimulationParams())
    for fid in (_abi.FIELD_P, _abi.FIELD_U, _abi.FIELD_V, _abi.FIELD_U_STAR, _abi.FIELD_V_STAR, _abi.FIELD_P_PRIME):
        x = a.field(fid)
        a.set_field(fid, x)
        assert np.array_equal(a.field(fid), x)
    a.update()
    cpu.update()
    assert_fields_identical(a, cpu, STATE_FIELDS, "after round trip")


def test_control_handle_runs_the_model():
    import time
    g = channel_grid(64, 24)
    h = Model.new(g, SimulationParams()).run()
    h.request_snapshot()
    deadline = time.time() + 30
    logs, snap = [], None
    while time.time() < deadline and (len(logs) < 5 or snap is None):
        logs += h.get_new_log_messages()
        snap = h.get_last_available_snapshot() or snap
        time.sleep(0.01)
    h.pause()
    h.stop()
    assert len(logs) >= 5 and [r.simulation_step for r in logs[:5]] == [1, 2, 3, 4, 5]
    assert snap is not None and snap.p.size == g.nx * g.ny


def test_full_size_properties_4096():
    """BASELINE size 4096x4096 (too big for the oracle): size-independent properties of a step —
    boundary identities of u, v, p', solver counters in range, finite fields, trivial first step."""
    g = Grid.uniform(4096, 4096, 40.0, 40.0, None)
    m = Model(g, SimulationParams())
    m.update()
    r = m.get_residuals()
    assert (r.jacobi_calls, r.sweeps, r.f64["u"]) == (2, 2, 0.0)
    for _ in range(3):
        m.update()
    r = m.get_residuals()
    assert 2 <= r.jacobi_calls <= 21 and r.jacobi_calls <= r.sweeps <= 1050
    nx = ny = 4096
    u = m.field(_abi.FIELD_U).reshape(ny, nx + 1)
    assert np.isfinite(u).all()
    assert (u[1:-1, 0] == 0.03).all() and (u[:, nx] == u[:, nx - 1]).all() and not u[0].any() and not u[-1].any()
    pp = m.field(_abi.FIELD_P_PRIME).reshape(ny, nx)
    assert (pp[:, 0] == pp[:, 1]).all() and not pp[:, nx - 1].any() and (pp[0] == pp[1]).all() and (pp[-1] == pp[-2]).all()
    v = m.field(_abi.FIELD_V).reshape(ny + 1, nx)
    assert not v[0].any() and not v[-1].any()


def test_errors():
    with pytest.raises(CfdError):
        Model(Grid.uniform(20, 8, 1.0, 1.0), SimulationParams())
    m = Model(channel_grid(32, 8), SimulationParams())
    m.close()
    with pytest.raises(CfdError):
        m.update()


@pytest.mark.parametrize("mode", [0, 1, 2, 3])
def test_division_by_constant_is_exact(mode):
    """cfdk::div_c / div_fast (hoisted reciprocal, the compiler's own fast-path sequence) must return bit for
    bit what `x / y` returns: random bit patterns, the solver's working range, dividends next to exactly
    representable quotients and next to rounding midpoints; divisors = the actual Jacobi divisors of the
    BASELINE grids plus awkward significands."""
    from cfd_demo_b200.model import selftest_division
    f = np.float32
    divisors = []
    for nx, lx in ((800, 30.0), (1024, 1.0), (4096, 1.0), (8192, 40.0), (16384, 1.0)):
        dx = float(f(lx) / f(nx))
        divisors += [dx * dx, 2.0 / (dx * dx) + 2.0 / (dx * dx)]
    dy = float(f(10.0) / f(264))
    divisors += [dy * dy, 2.0 / (0.0375 * 0.0375) + 2.0 / (dy * dy), 3.0, 1.0 / 3.0, 1.9999999999999998, 1.0000000000000002,
                 7.0e-300, 3.0e300]
    for y in divisors:
        bad, fast = selftest_division(y, 20_000_000, seed=42 + mode, mode=mode)
        assert bad == 0, (y, mode, bad)
        if mode in (1, 2, 3) and 1e-200 < y < 1e200:
            assert fast > 0


@pytest.mark.parametrize("precision", [64, 32])
@pytest.mark.parametrize("other", [_abi.FLAG_NO_GRAPH, _abi.FLAG_BASELINE_SWEEP])
def test_tuned_sweep_equals_baseline_sweep(precision, other):
    """On this grid the default is the single cooperative launch per solve (k_jacobi_persist); FLAG_NO_GRAPH is the
    TMA-staged one-launch-per-sweep kernel that large grids get (k_jacobi_sweep5), FLAG_BASELINE_SWEEP the simple
    one-column kernel with the compiler's divisions.  All three: complete state bit-identical, on a grid wide enough to
    use several blocks per row and whose last 64-column strip is partial."""
    from cfd_demo_b200.model import default_options
    g = channel_grid(1040, 61)
    a = Model(g, SimulationParams(), precision=precision)
    o = default_options()
    o.precision = precision
    o.flags = other
    b = Model(g, SimulationParams(), options=o)
    for s in range(14):
        a.update()
        b.update()
    for fid in STATE_FIELDS:
        assert np.array_equal(a.field(fid), b.field(fid)), _abi.FIELD_NAMES[fid]
    ra, rb = a.get_residuals(), b.get_residuals()
    assert (ra.jacobi_calls, ra.sweeps, ra.f64["p"]) == (rb.jacobi_calls, rb.sweeps, rb.f64["p"])


@pytest.mark.parametrize("precision", [64, 32])
@pytest.mark.parametrize("flag", [_abi.FLAG_REGISTER_SWEEP, _abi.FLAG_BULK_SWEEP, _abi.FLAG_SWEEP4, _abi.FLAG_TEMPORAL, _abi.FLAG_PERSISTENT_SWEEP])
def test_ab_sweep_kernels_live_in_their_own_library(precision, flag):
    """The sweep kernels that lost their A/B (csrc/cfd_sweeps_ab.cuh) are not in the product library — it refuses their
    flags — but in libcfd_b200_ab.so, where each still matches the shipped kernel bit for bit (tests/ab_sweep_check.py)."""
    import os
    import subprocess
    import sys
    from cfd_demo_b200.model import default_options
    o = default_options()
    o.flags = flag
    with pytest.raises(CfdError) as e:
        Model(channel_grid(64, 24), SimulationParams(), options=o)
    assert e.value.code == _abi.CFD_ERR_UNSUPPORTED
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, CFD_B200_LIB=os.path.join(root, "cfd_demo_b200", "libcfd_b200_ab.so"))
    r = subprocess.run([sys.executable, os.path.join(root, "tests", "ab_sweep_check.py"), str(flag), str(precision)],
                       capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0 and "ab sweep ok" in r.stdout, r.stdout[-1000:] + r.stderr[-2000:]


def test_cpp_headless_driver_matches_python_mirror():
    """The C++ mirror (host/cfd_model.hpp) drives the same C ABI: its residual log equals the Python mirror's."""
    import os
    import subprocess
    exe = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "cfd_demo_b200", "host", "cfd_headless")
    r = subprocess.run([exe, "5"], capture_output=True, text=True, check=True)
    m = Model(default_grid(), SimulationParams())
    for _ in range(5):
        m.update()
    res = m.get_residuals()
    line = [l for l in r.stdout.splitlines() if l.startswith("step 5 ")][0]
    assert f"K={res.jacobi_calls} S={res.sweeps}" in line and f"u={res.u:.3e}" in line


@pytest.mark.parametrize("scenario", [Scenario.Channel, Scenario.Cavity])
@pytest.mark.parametrize("scheme", [VelocityScheme.FirstOrder, VelocityScheme.SecondOrder])
def test_mode_c_cg_matches_oracle_to_tolerance(scenario, scheme):
    """Mode C (CG, an extension): dot products are summed in a different order than the oracle's, so parity
    is to a tolerance.  North_star bar: relative L2 <= 1e-9 on u, v, p after N steps when both sides converge
    the Poisson solve to the same residual (here dt*rms(r) <= 1e-13)."""
    from cfd_demo_b200.model import default_options
    from cfd_demo_b200.types import PressureSolver
    from oracle.cpu_oracle import default_consts
    n = 64
    g = box_grid(n) if scenario == Scenario.Cavity else channel_grid(n, 48)
    prm = SimulationParams(dt=1e-3, viscosity=0.01, scenario=scenario, pressure_solver=PressureSolver.CG,
                           velocity_scheme=scheme)
    consts = default_consts()
    consts.cg_tolerance = 1e-13
    o = default_options()
    o.consts = consts
    gpu = Model(g, prm, options=o)
    cpu = OracleModel(g, prm, precision=64, consts=consts)
    for s in range(15):
        gpu.update()
        cpu.update()
        rg, rc = gpu.get_residuals(), cpu.get_residuals()
        assert rg.jacobi_calls == rc.jacobi_calls == 2
        assert abs(rg.sweeps - rc.sweeps) <= 4 and rg.f64["p"] <= 1e-13
    for fid in (_abi.FIELD_U, _abi.FIELD_V, _abi.FIELD_P):
        a, b = gpu.field(fid), cpu.field(fid)
        assert np.isfinite(a).all() and np.abs(b).max() > 0
        assert rel_l2(a, b) <= 1e-9, (_abi.FIELD_NAMES[fid], rel_l2(a, b))


def test_mode_c_is_deterministic_run_to_run():
    from cfd_demo_b200.types import PressureSolver
    prm = SimulationParams(dt=1e-3, viscosity=0.01, scenario=Scenario.Cavity, pressure_solver=PressureSolver.CG)
    outs = []
    for _ in range(2):
        m = Model(box_grid(128), prm)
        for _ in range(6):
            m.update()
        outs.append((m.field(_abi.FIELD_U), m.field(_abi.FIELD_P), m.get_residuals().sweeps))
    assert np.array_equal(outs[0][0], outs[1][0]) and np.array_equal(outs[0][1], outs[1][1]) and outs[0][2] == outs[1][2]


def test_strips_over_nccl_are_bit_identical():
    """world_size 2 (or more) strips against the single-domain model; needs >= 2 GPUs on the box
    (`gpurun --gpus 2`), skipped on a single-GPU box."""
    import os
    import subprocess
    import sys
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    n = 2 if n < 4 else 4
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}",
                        "--master-addr", "127.0.0.1", "--master-port", "29611",
                        os.path.join(root, "tests", "mgpu_strip_check.py")], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and "strips ok" in r.stdout, r.stdout[-2000:] + r.stderr[-3000:]


def test_full_size_saturated_step_bit_exact_2048():
    """A dense, saturated (K=21, S=1050) timestep at 2048x2048 fp64: the GPU spins the flow up, its complete
    state (u, v, p, u_star, v_star, p_prime + step/time/dt) is loaded into the CPU oracle, and ONE more step on
    both sides must agree bit for bit — 4.4e9 cell updates through every kernel of the hot path."""
    g = Grid.uniform(2048, 2048, 40.0, 40.0, None)
    prm = SimulationParams()
    gpu = Model(g, prm)
    for _ in range(40):
        gpu.update()
        r0 = gpu.get_residuals()
        if (r0.jacobi_calls, r0.sweeps) == (21, 1050) and r0.simulation_step >= 24:
            break
    assert (r0.jacobi_calls, r0.sweeps) == (21, 1050), r0
    cpu = OracleModel(g, prm, precision=64)
    for fid in (_abi.FIELD_P, _abi.FIELD_U, _abi.FIELD_V, _abi.FIELD_U_STAR, _abi.FIELD_V_STAR, _abi.FIELD_P_PRIME):
        cpu.set_field(fid, gpu.field(fid))
    cpu.set_scalars(r0.simulation_step, r0.f64["simulation_time"], r0.f64["dt"])
    gpu.update()
    cpu.update()
    assert_residuals_identical(gpu.get_residuals(), cpu.get_residuals(), "2048^2 saturated step")
    assert_fields_identical(gpu, cpu, STATE_FIELDS, "2048^2 saturated step")
    pp = gpu.field(_abi.FIELD_P_PRIME).reshape(2048, 2048)
    assert np.abs(pp[:, :-1]).min() > 0 and not pp[:, -1].any()  # dense: no untouched zero regions (outlet column is 0)


@pytest.mark.parametrize("scenario", [Scenario.Channel, Scenario.Cavity])
@pytest.mark.parametrize("shape", [(64, 64), (128, 93), (48, 6), (16, 4), (24, 5), (512, 512)])
def test_mode_c_mgcg_matches_oracle_to_tolerance(scenario, shape):
    """Mode C fast path (multigrid-preconditioned CG, an extension): smoother, transfers and coarse operators are
    bit-identical to the oracle's, only the dot products are summed in another order -> tolerance parity
    (north_star: relative L2 <= 1e-9 after N steps when both sides converge the Poisson solve to the same
    residual; here dt*rms(r) <= 1e-12), same iteration counts."""
    from cfd_demo_b200.model import default_options
    from cfd_demo_b200.types import PressureSolver
    from oracle.cpu_oracle import default_consts
    nx, ny = shape
    g = Grid.uniform(nx, ny, nx / 64.0, ny / 64.0, None) if scenario == Scenario.Cavity else channel_grid(nx, ny, lx=nx / 16.0, ly=ny / 16.0, cylinder=nx >= 64)
    prm = SimulationParams(dt=1e-3, viscosity=0.01, scenario=scenario, pressure_solver=PressureSolver.MGCG,
                           velocity_scheme=VelocityScheme.SecondOrder)
    consts = default_consts()
    consts.cg_tolerance = 1e-12
    o = default_options()
    o.consts = consts
    gpu = Model(g, prm, options=o)
    cpu = OracleModel(g, prm, precision=64, consts=consts)
    for s in range(6 if nx >= 512 else 12):
        gpu.update()
        cpu.update()
        rg, rc = gpu.get_residuals(), cpu.get_residuals()
        assert rg.jacobi_calls == rc.jacobi_calls == 2
        assert abs(rg.sweeps - rc.sweeps) <= 1 and rg.sweeps <= 20 and rg.f64["p"] <= 1e-12, (s, rg.sweeps, rc.sweeps)
    for fid in (_abi.FIELD_U, _abi.FIELD_V, _abi.FIELD_P):
        a, b = gpu.field(fid), cpu.field(fid)
        assert np.isfinite(a).all() and (fid != _abi.FIELD_U or np.abs(b).max() > 0)  # v is identically 0 on 4-row grids
        assert rel_l2(a, b) <= 1e-9, (_abi.FIELD_NAMES[fid], rel_l2(a, b))


def test_mode_c_mgcg_first_vcycle_is_bit_identical():
    """With cg_max_iterations = 1 the solve is ONE preconditioned step x = alpha * V(rhs): everything in V (the
    reference's Jacobi sweep as smoother, restriction, coarse sweeps, prolongation) is order-free arithmetic, so
    p' must agree with the oracle up to the single scalar alpha (two dot products)."""
    from cfd_demo_b200.model import default_options
    from cfd_demo_b200.types import PressureSolver
    from oracle.cpu_oracle import default_consts
    g = channel_grid(136, 61, lx=13.6, ly=6.1)
    prm = SimulationParams(dt=1e-3, viscosity=0.01, pressure_solver=PressureSolver.MGCG)
    consts = default_consts()
    consts.cg_max_iterations = 1
    consts.outer_rounds = 0
    consts.mg_warm_start = 0
    o = default_options()
    o.consts = consts
    gpu = Model(g, prm, options=o)
    cpu = OracleModel(g, prm, precision=64, consts=consts)
    for _ in range(3):
        gpu.update()
        cpu.update()
    a, b = gpu.field(_abi.FIELD_P_PRIME), cpu.field(_abi.FIELD_P_PRIME)
    k = np.argmax(np.abs(b))
    ratio = a[k] / b[k]
    assert abs(ratio - 1.0) < 1e-12
    assert np.abs(a - ratio * b).max() <= 4e-16 * np.abs(b).max()


def test_mgcg_rejects_bad_constants():
    from cfd_demo_b200.model import CfdError, default_options
    from cfd_demo_b200.types import PressureSolver
    o = default_options()
    o.consts.mg_smoothing = 0
    with pytest.raises(CfdError):
        Model(box_grid(32), SimulationParams(pressure_solver=PressureSolver.MGCG), options=o)


def test_mgcg_bottom_kernel_matches_per_level_launches():
    """The single-block kernel that runs the bottom of the V-cycle performs the same per-cell arithmetic as one
    launch per operation (CFD_FLAG_MG_NO_BOTTOM_KERNEL): complete state bit-identical."""
    from cfd_demo_b200.model import default_options
    from cfd_demo_b200.types import PressureSolver
    g = channel_grid(264, 200, lx=26.4, ly=20.0)
    prm = SimulationParams(dt=1e-3, viscosity=0.01, pressure_solver=PressureSolver.MGCG)
    models = []
    for flags in (0, 256):
        o = default_options()
        o.flags = flags
        m = Model(g, prm, options=o)
        for _ in range(6):
            m.update()
        models.append(m)
    assert models[0].get_residuals().sweeps == models[1].get_residuals().sweeps > 0
    assert_fields_identical(models[0], models[1], STATE_FIELDS, "bottom kernel vs per-level launches")


def test_mgcg_warm_start_state_restarts_in_the_oracle():
    """The start vector (extrapolated from the previous steps' first-solve results: CFD_FIELD_MG_GUESS / _LAST / _LAST2) is carried state: load the GPU's
    complete state into the oracle after a spin-up and compute one more step on both sides — same iteration count
    (far fewer than a cold start needs), fields within the Mode C tolerance; and a cold-start model converges to the
    same velocities."""
    from cfd_demo_b200.model import default_options
    from cfd_demo_b200.types import PressureSolver
    from oracle.cpu_oracle import default_consts
    n = 256
    g = Grid.uniform(n, n, 1.0, 1.0, None)
    nu = 1e-3
    prm = SimulationParams(dt=0.02 * (1.0 / n) ** 2 / nu, viscosity=nu, scenario=Scenario.Cavity,
                           pressure_solver=PressureSolver.MGCG)
    consts = default_consts()
    consts.ramp_up_steps = 5
    its = {}
    models = {}
    for warm in (3, 0):
        consts.mg_warm_start = warm
        o = default_options()
        o.consts = consts
        m = Model(g, prm, options=o)
        for _ in range(40):
            m.update()
        its[warm] = m.get_residuals().sweeps
        models[warm] = m
    assert 0 < its[3] < its[0], its
    assert rel_l2(models[3].field(_abi.FIELD_U), models[0].field(_abi.FIELD_U)) < 1e-6
    consts.mg_warm_start = 3
    gpu = models[3]
    cpu = OracleModel(g, prm, precision=64, consts=consts)
    r0 = gpu.get_residuals()
    for fid in (_abi.FIELD_P, _abi.FIELD_U, _abi.FIELD_V, _abi.FIELD_U_STAR, _abi.FIELD_V_STAR, _abi.FIELD_P_PRIME,
                _abi.FIELD_MG_GUESS, _abi.FIELD_MG_LAST, _abi.FIELD_MG_LAST2):
        cpu.set_field(fid, gpu.field(fid))
    cpu.set_scalars(r0.simulation_step, r0.f64["simulation_time"], r0.f64["dt"])
    gpu.update()
    cpu.update()
    rg, rc = gpu.get_residuals(), cpu.get_residuals()
    assert rg.sweeps == rc.sweeps and abs(rg.sweeps - its[3]) <= 1 and rg.jacobi_calls == rc.jacobi_calls == 2
    for fid in (_abi.FIELD_U, _abi.FIELD_V, _abi.FIELD_P, _abi.FIELD_MG_GUESS, _abi.FIELD_MG_LAST, _abi.FIELD_MG_LAST2):
        assert rel_l2(gpu.field(fid), cpu.field(fid)) <= 1e-9, _abi.FIELD_NAMES[fid]


@pytest.mark.parametrize("precision", [64, 32])
def test_nan_and_inf_propagate_exactly_like_the_oracle(precision):
    """A blown-up run "just shows NaNs" in the reference (no NaN checks); its reductions fold with
    f32::max, which ignores NaN (:338, :795-798, :879-880).  Poison one interior velocity with NaN and one with +inf:
    the NaN / inf pattern of every state field, the residuals and the solver counters must match the oracle for as
    long as the front spreads (this also drives the sweep through its out-of-line true-division path)."""
    g = channel_grid(64, 24)
    prm = SimulationParams()
    gpu, cpu = Model(g, prm, precision=precision), OracleModel(g, prm, precision=precision)
    for _ in range(8):
        gpu.update()
        cpu.update()
    u = gpu.field(_abi.FIELD_U)
    u[10 * 65 + 20] = np.nan
    u[15 * 65 + 40] = np.inf
    for m in (gpu, cpu):
        m.set_field(_abi.FIELD_U, u)
    for s in range(4):
        gpu.update()
        cpu.update()
        rg, rc = gpu.get_residuals(), cpu.get_residuals()
        assert (rg.jacobi_calls, rg.sweeps) == (rc.jacobi_calls, rc.sweeps), s
        for k in ("dt", "p", "u", "v"):
            a, b = rg.f64[k], rc.f64[k]
            assert a == b or (np.isnan(a) and np.isnan(b)), (s, k, a, b)
        assert_fields_identical(gpu, cpu, STATE_FIELDS, f"poisoned run, step {s}")
    assert np.isnan(gpu.field(_abi.FIELD_U)).sum() > 10
