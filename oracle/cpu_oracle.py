"""ctypes driver of the CPU ORACLE (oracle/cfd_oracle.hpp) — TEST INFRASTRUCTURE, not product code.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
"Parity unpinned": the reference holds no golden vectors for src/model.rs and cannot be compiled here.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from cfd_demo_b200 import _abi
from cfd_demo_b200.types import Grid, Residuals, SimulationParams

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIBS = {}


EXCHANGE_CB = C.CFUNCTYPE(None, C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint64, C.c_int, C.c_int, C.c_int)
ALLREDUCE_CB = C.CFUNCTYPE(C.c_double, C.c_void_p, C.c_double, C.c_int)
GATHER_CB = C.CFUNCTYPE(None, C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64, C.c_int)


def build(force: bool = False) -> None:
    """Compile the oracle shared objects with the committed Makefile (gcc only)."""
    args = ["make", "-C", _HERE]
    if force:
        args.append("-B")
    subprocess.run(args, check=True, capture_output=True)


def _load(checked: bool = False):
    name = "libcfd_oracle_checked.so" if checked else "libcfd_oracle.so"
    if name in _LIBS:
        return _LIBS[name]
    path = os.path.join(_HERE, name)
    if not os.path.exists(path):
        build()
    lib = C.CDLL(path)
    lib.cfdo_create.restype = C.c_void_p
    lib.cfdo_create.argtypes = [C.POINTER(_abi.CfdGrid), C.POINTER(_abi.CfdParams),
                                C.POINTER(_abi.CfdSolverConsts), C.c_int]
    lib.cfdo_destroy.argtypes = [C.c_void_p]
    lib.cfdo_update.argtypes = [C.c_void_p]
    lib.cfdo_set_params.argtypes = [C.c_void_p, C.POINTER(_abi.CfdParams)]
    lib.cfdo_field_len.restype = C.c_uint64
    lib.cfdo_field_len.argtypes = [C.c_void_p, C.c_int]
    lib.cfdo_get_field_f64.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
    lib.cfdo_set_field_f64.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
    lib.cfdo_get_residuals.argtypes = [C.c_void_p, C.POINTER(_abi.CfdResiduals)]
    lib.cfdo_obstacle_count.restype = C.c_uint64
    lib.cfdo_obstacle_count.argtypes = [C.c_void_p]
    lib.cfdo_current_inlet_velocity.restype = C.c_double
    lib.cfdo_current_inlet_velocity.argtypes = [C.c_void_p]
    lib.cfdo_stage.restype = C.c_double
    lib.cfdo_stage.argtypes = [C.c_void_p, C.c_int]
    lib.cfdo_set_scalars.argtypes = [C.c_void_p, C.c_uint64, C.c_double, C.c_double]
    lib.cfdo_set_strip.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_int, EXCHANGE_CB, ALLREDUCE_CB, C.c_void_p]
    lib.cfdo_set_gather.argtypes = [C.c_void_p, GATHER_CB, C.c_void_p]
    lib.cfdo_total_sweeps.restype = C.c_uint64
    lib.cfdo_total_sweeps.argtypes = [C.c_void_p]
    lib.cfd_solver_consts_default.argtypes = [C.POINTER(_abi.CfdSolverConsts)]
    _LIBS[name] = lib
    return lib


def default_consts() -> _abi.CfdSolverConsts:
    c = _abi.CfdSolverConsts()
    _load().cfd_solver_consts_default(C.byref(c))
    return c


class OracleModel:
    """CPU restatement of the reference `Model` (src/model.rs) in float (precision=32) or double (64)."""

    STAGE_PREDICTOR_U, STAGE_PREDICTOR_V, STAGE_DIVERGENCE, STAGE_PRESSURE = 0, 1, 2, 3
    STAGE_CORRECTOR, STAGE_BC, STAGE_COPY_STAR, STAGE_ONE_SWEEP = 4, 5, 6, 7

    def __init__(self, grid: Grid, params: SimulationParams, precision: int = 64, consts=None,
                 checked: bool = False):
        self._lib = _load(checked)
        self.grid, self.precision = grid, precision
        g, p = grid.to_c(), params.to_c()
        cptr = C.byref(consts) if consts is not None else None
        self._h = self._lib.cfdo_create(C.byref(g), C.byref(p), cptr, precision)
        if not self._h:
            raise ValueError("oracle: invalid grid (needs nx % 8 == 0, nx >= 16, ny >= 4)")

    def close(self):
        if getattr(self, "_h", None):
            self._lib.cfdo_destroy(self._h)
            self._h = None

    __del__ = close

    def update(self):
        self._lib.cfdo_update(self._h)

    def set_parameters(self, params: SimulationParams):
        p = params.to_c()
        self._lib.cfdo_set_params(self._h, C.byref(p))

    def field(self, fid: int) -> np.ndarray:
        n = self._lib.cfdo_field_len(self._h, fid)
        out = np.empty(n, dtype=np.float64)
        rc = self._lib.cfdo_get_field_f64(self._h, fid, out.ctypes.data)
        assert rc == 0
        return out

    def set_field(self, fid: int, values: np.ndarray):
        a = np.ascontiguousarray(values, dtype=np.float64)
        assert a.size == self._lib.cfdo_field_len(self._h, fid)
        assert self._lib.cfdo_set_field_f64(self._h, fid, a.ctypes.data) == 0

    def get_residuals(self) -> Residuals:
        r = _abi.CfdResiduals()
        self._lib.cfdo_get_residuals(self._h, C.byref(r))
        return Residuals.from_c(r)

    def obstacle_count(self) -> int:
        return int(self._lib.cfdo_obstacle_count(self._h))

    def current_inlet_velocity(self) -> float:
        return float(self._lib.cfdo_current_inlet_velocity(self._h))

    def stage(self, stage: int) -> float:
        return float(self._lib.cfdo_stage(self._h, stage))

    def set_scalars(self, simulation_step: int, simulation_time: float, dt: float):
        self._lib.cfdo_set_scalars(self._h, int(simulation_step), float(simulation_time), float(dt))

    def set_strip(self, ja: int, jb: int, owns_top: bool, exchange, allreduce, gather=None):
        """Restrict the model to rows [ja, jb) of a strip decomposition.  `exchange(field2d, ja, jb, below, above)`
        gets a writable (nrows, row_len) numpy view and must refresh `below` halo rows under ja and `above` over
        jb (sending the mirror-image rows to the neighbours); `allreduce(x, op)` reduces a float over the ranks
        (op 0 = max, 1 = sum); `gather(field2d, lo, hi)` (needed by MGCG) contributes rows [lo, hi) of a writable
        (nrows, row_len) view and must fill in every other rank's rows."""
        def _ex(_user, ptr, row_len, nrows, below, above, elem_bytes):
            dtype = np.float32 if elem_bytes == 4 else np.float64
            buf = (C.c_char * (row_len * nrows * elem_bytes)).from_address(ptr)
            exchange(np.frombuffer(buf, dtype=dtype).reshape(nrows, row_len), ja, jb, below, above)

        def _ar(_user, x, op):
            return float(allreduce(float(x), int(op)))

        self._cb = (EXCHANGE_CB(_ex), ALLREDUCE_CB(_ar))  # keep the thunks alive
        self._lib.cfdo_set_strip(self._h, int(ja), int(jb), 1 if owns_top else 0, self._cb[0], self._cb[1], None)
        if gather is not None:
            def _ga(_user, ptr, row_len, nrows, lo, hi, elem_bytes):
                dtype = np.float32 if elem_bytes == 4 else np.float64
                buf = (C.c_char * (row_len * nrows * elem_bytes)).from_address(ptr)
                gather(np.frombuffer(buf, dtype=dtype).reshape(nrows, row_len), int(lo), int(hi))

            self._cb_gather = GATHER_CB(_ga)
            self._lib.cfdo_set_gather(self._h, self._cb_gather, None)

    def total_sweeps(self) -> int:
        return int(self._lib.cfdo_total_sweeps(self._h))
