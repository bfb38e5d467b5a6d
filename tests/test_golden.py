"""Known-answer vectors of the reference-pinned case (tests/golden/mode_r_default_grid.json, made by
tests/golden/make_golden.py from the two CPU restatements): per-step solver counters and residuals (hex floats), SHA-256
of p, u, v after the last step, and sampled cell values.  The CPU test pins the oracle against silent change; the GPU
test checks the CUDA path against the committed vectors directly, without running any CPU code next to it."""
import hashlib
import json
import os

import numpy as np
import pytest

from cfd_demo_b200 import _abi
from cfd_demo_b200.types import SimulationParams, VelocityScheme, default_grid

GOLDEN = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "mode_r_default_grid.json")))


def digest(a, precision):
    dt = np.float32 if precision == 32 else np.float64
    b = np.ascontiguousarray(a.astype(dt)) + dt(0)
    return hashlib.sha256(b.tobytes()).hexdigest()


def check(model, case):
    g = default_grid()
    for s, want in enumerate(case["per_step"]):
        model.update()
        r = model.get_residuals()
        got = {"K": r.jacobi_calls, "S": r.sweeps, "dt": float(r.f64["dt"]).hex(), "p": float(r.f64["p"]).hex(),
               "u": float(r.f64["u"]).hex(), "v": float(r.f64["v"]).hex()}
        assert got == want, (s, got, want)
    for fid in (_abi.FIELD_P, _abi.FIELD_U, _abi.FIELD_V):
        name = _abi.FIELD_NAMES[fid]
        assert digest(model.field(fid), case["precision"]) == case["sha256"][name], name
    p = model.field(_abi.FIELD_P).reshape(g.ny, g.nx)
    u = model.field(_abi.FIELD_U).reshape(g.ny, g.nx + 1)
    for smp in case["samples"]:
        assert float(p[smp["j"], smp["i"]]).hex() == smp["p"] and float(u[smp["j"], smp["i"]]).hex() == smp["u"], smp


def test_golden_file_describes_the_default_grid():
    g = default_grid()
    assert GOLDEN["grid"] == {"nx": g.nx, "ny": g.ny, "lx": g.lx, "ly": g.ly}
    assert set(GOLDEN["cases"]) == {"first_order_f32", "second_order_f32", "first_order_f64"}
    first = GOLDEN["cases"]["first_order_f32"]["per_step"]
    assert [(s["K"], s["S"]) for s in first[:4]] == [(2, 2)] * 4 and (first[-1]["K"], first[-1]["S"]) == (21, 1050)


@pytest.mark.parametrize("name", ["first_order_f32", "second_order_f32"])
def test_oracle_reproduces_the_golden_vectors(name, oracle_built):
    from oracle.cpu_oracle import OracleModel
    case = GOLDEN["cases"][name]
    check(OracleModel(default_grid(), SimulationParams(velocity_scheme=VelocityScheme(case["scheme"])),
                      precision=case["precision"]), case)


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(GOLDEN["cases"]))
def test_cuda_path_reproduces_the_golden_vectors(name):
    from cfd_demo_b200.model import Model
    case = GOLDEN["cases"][name]
    check(Model(default_grid(), SimulationParams(velocity_scheme=VelocityScheme(case["scheme"])),
                precision=case["precision"]), case)
