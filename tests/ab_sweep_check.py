"""Run with CFD_B200_LIB=<...>/libcfd_b200_ab.so (the product library plus the sweep kernels that lost their A/B,
csrc/cfd_sweeps_ab.cuh): every A/B sweep kernel against the shipped default, complete state bit-identical.
usage: python tests/ab_sweep_check.py <flag> <precision>"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from cfd_demo_b200 import _abi  # noqa: E402
from cfd_demo_b200.model import Model, default_options  # noqa: E402
from cfd_demo_b200.types import SimulationParams  # noqa: E402
from helpers import STATE_FIELDS, channel_grid  # noqa: E402


def main():
    flag, precision = int(sys.argv[1]), int(sys.argv[2])
    g = channel_grid(1040, 61)
    models = []
    for flags in (_abi.FLAG_NO_GRAPH, flag):  # the shipped one-launch-per-sweep kernel vs the A/B kernel
        o = default_options()
        o.precision = precision
        o.flags = flags
        m = Model(g, SimulationParams(), options=o)
        for _ in range(14):
            m.update()
        models.append(m)
    a, b = models
    for fid in STATE_FIELDS:
        assert np.array_equal(a.field(fid), b.field(fid)), _abi.FIELD_NAMES[fid]
    ra, rb = a.get_residuals(), b.get_residuals()
    assert (ra.jacobi_calls, ra.sweeps, ra.f64["p"]) == (rb.jacobi_calls, rb.sweeps, rb.f64["p"])
    print("ab sweep ok", flag, precision)


if __name__ == "__main__":
    main()
