"""Generates tests/golden/mode_r_default_grid.json: known-answer vectors for the reference-pinned case (SURVEY 8c/8d
config 1: default_grid() 800 x 264 + cylinder, SimulationParams::default(), f32 = the reference's own arithmetic, and
fp64), produced by the C++ oracle and cross-checked here against the independent numpy restatement before writing.
PARITY UNPINNED against the Rust itself (it cannot be built here); these vectors pin the two restatements, and through
tests/test_gpu_parity.py the CUDA path, against silent change.

    python tests/golden/make_golden.py
"""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from cfd_demo_b200 import _abi  # noqa: E402
from cfd_demo_b200.types import SimulationParams, VelocityScheme, default_grid  # noqa: E402
from oracle.cpu_oracle import OracleModel  # noqa: E402
from oracle.numpy_restatement import NumpyModel  # noqa: E402

CASES = [("first_order_f32", 32, VelocityScheme.FirstOrder, 24), ("second_order_f32", 32, VelocityScheme.SecondOrder, 16),
         ("first_order_f64", 64, VelocityScheme.FirstOrder, 16)]
SAMPLES = [(1, 1), (100, 130), (200, 132), (399, 10), (640, 200), (798, 262)]  # (i, j) cells


def digest(a, precision):
    dt = np.float32 if precision == 32 else np.float64
    b = np.ascontiguousarray(a.astype(dt))
    b = b + dt(0)  # -0.0 -> +0.0, so that the digest does not depend on the sign of zero
    return hashlib.sha256(b.tobytes()).hexdigest()


def main():
    g = default_grid()
    out = {"grid": {"nx": g.nx, "ny": g.ny, "lx": g.lx, "ly": g.ly}, "cases": {}}
    for name, precision, scheme, steps in CASES:
        prm = SimulationParams(velocity_scheme=scheme)
        a = OracleModel(g, prm, precision=precision)
        b = NumpyModel(g, prm, dtype=np.float32 if precision == 32 else np.float64)
        per_step = []
        for _ in range(steps):
            a.update()
            b.update()
            r = a.get_residuals()
            assert (r.jacobi_calls, r.sweeps) == (b.K, b.S)
            per_step.append({"K": r.jacobi_calls, "S": r.sweeps, "dt": float(r.f64["dt"]).hex(), "p": float(r.f64["p"]).hex(),
                             "u": float(r.f64["u"]).hex(), "v": float(r.f64["v"]).hex()})
        fields = {}
        for fid, arr in ((_abi.FIELD_P, b.p), (_abi.FIELD_U, b.u), (_abi.FIELD_V, b.v)):
            x = a.field(fid)
            assert np.array_equal(x, arr.astype(np.float64)), (name, fid)
            fields[_abi.FIELD_NAMES[fid]] = digest(x, precision)
        p = a.field(_abi.FIELD_P).reshape(g.ny, g.nx)
        u = a.field(_abi.FIELD_U).reshape(g.ny, g.nx + 1)
        samples = [{"i": i, "j": j, "p": float(p[j, i]).hex(), "u": float(u[j, i]).hex()} for i, j in SAMPLES]
        out["cases"][name] = {"precision": precision, "scheme": int(scheme), "steps": steps, "per_step": per_step,
                              "sha256": fields, "samples": samples}
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "mode_r_default_grid.json"), "w") as f:
        json.dump(out, f, indent=1)
    print("written", {k: v["per_step"][-1] for k, v in out["cases"].items()})


if __name__ == "__main__":
    main()
