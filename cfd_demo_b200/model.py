"""Host-side mirror of the reference `Model` / `SimulationControlHandle` (reference: src/model.rs) over the
C ABI of include/cfd_b200.h (libcfd_b200.so, hand-written sm_100a kernels).

Same names, argument meaning and error behaviour as the reference: `Model.new(grid, params)` (:219),
`update()` (:304), `set_parameters()` (:1250), `get_snapshot()` (:1259), `get_residuals()` (:1269),
`run()` -> `SimulationControlHandle` (:1282) with `stop / pause / resume / set_params /
request_snapshot / get_last_available_snapshot / get_new_log_messages` (:71-117).  Where the reference
panics (invalid grid, a call on a dropped model) this raises `CfdError`.

There is NO CPU fallback: if libcfd_b200.so is missing or no CUDA device is usable, construction fails.
"""
from __future__ import annotations

import ctypes as C
import os
import queue
import threading
import time
from typing import List, Optional

import numpy as np

from . import _abi
from .types import Grid, Residuals, SimSnapshot, SimulationParams

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("CFD_B200_LIB") or os.path.join(_HERE, "libcfd_b200.so")  # override: A/B builds (tools/)
_lib = None


class CfdError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"cfd_b200 error {code}: {message}")
        self.code = code


def load_library():
    """Load libcfd_b200.so (built in-tree by `__graft_entry__.build()` / csrc/Makefile). Fails loudly."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise CfdError(_abi.CFD_ERR_CUDA,
                       f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                       "(there is no CPU fallback)")
    lib = C.CDLL(LIB_PATH)
    P = C.POINTER
    lib.cfd_abi_version.restype = C.c_int
    lib.cfd_last_error.restype = C.c_char_p
    lib.cfd_solver_consts_default.argtypes = [P(_abi.CfdSolverConsts)]
    lib.cfd_options_default.argtypes = [P(_abi.CfdOptions)]
    lib.cfd_model_create.argtypes = [P(_abi.CfdGrid), P(_abi.CfdParams), P(C.c_void_p)]
    lib.cfd_model_create_ex.argtypes = [P(_abi.CfdGrid), P(_abi.CfdParams), P(_abi.CfdOptions), P(C.c_void_p)]
    lib.cfd_model_destroy.argtypes = [C.c_void_p]
    lib.cfd_model_destroy.restype = None
    lib.cfd_model_update.argtypes = [C.c_void_p]
    lib.cfd_model_update_n.argtypes = [C.c_void_p, C.c_uint64]
    lib.cfd_model_set_params.argtypes = [C.c_void_p, P(_abi.CfdParams)]
    lib.cfd_model_get_snapshot.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, P(C.c_float)]
    lib.cfd_model_snapshot_begin.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.cfd_model_snapshot_end.argtypes = [C.c_void_p, P(C.c_float)]
    lib.cfd_model_get_residuals.argtypes = [C.c_void_p, P(_abi.CfdResiduals)]
    lib.cfd_model_render_rgba.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, P(C.c_float), P(C.c_float)]
    lib.cfd_model_field_len.argtypes = [C.c_void_p, C.c_int32, P(C.c_uint64)]
    lib.cfd_model_get_field_f64.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_uint64]
    lib.cfd_model_set_field_f64.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_uint64]
    lib.cfd_model_rows.argtypes = [C.c_void_p, P(C.c_uint64), P(C.c_uint64)]
    lib.cfd_strip_rows.argtypes = [C.c_uint64, C.c_int32, C.c_int32, P(C.c_uint64), P(C.c_uint64)]
    lib.cfd_model_last_timing.argtypes = [C.c_void_p, P(C.c_double), P(C.c_double), P(C.c_uint64)]
    lib.cfd_nccl_unique_id.argtypes = [C.c_void_p]
    lib.cfd_model_profile_smoother.argtypes = [C.c_void_p, C.c_int32]
    lib.cfd_model_last_smoother_timing.argtypes = [C.c_void_p, P(C.c_double), P(C.c_uint64)]
    lib.cfd_model_tracers_inject.argtypes = [C.c_void_p]
    lib.cfd_model_tracers_update.argtypes = [C.c_void_p, C.c_double]
    lib.cfd_model_tracers_count.argtypes = [C.c_void_p, P(C.c_uint64)]
    lib.cfd_model_tracers_get.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, P(C.c_uint64)]
    lib.cfd_model_tracers_clear.argtypes = [C.c_void_p]
    lib.cfd_host_alloc.argtypes = [C.c_uint64, P(C.c_void_p)]
    lib.cfd_host_free.argtypes = [C.c_void_p]
    lib.cfd_host_free.restype = None
    lib.cfd_selftest_division.argtypes = [C.c_double, C.c_uint64, C.c_uint64, C.c_int32, P(C.c_uint64),
                                          P(C.c_uint64)]
    if lib.cfd_abi_version() != _abi.CFD_ABI_VERSION:
        raise CfdError(_abi.CFD_ERR_UNSUPPORTED, "libcfd_b200.so ABI version mismatch; rebuild")
    _lib = lib
    return lib


def _check(lib, rc: int):
    if rc != 0:
        raise CfdError(rc, (lib.cfd_last_error() or b"").decode())


def default_options() -> _abi.CfdOptions:
    o = _abi.CfdOptions()
    load_library().cfd_options_default(C.byref(o))
    return o


def _checked_out(arr, count: int, dtype, what: str) -> np.ndarray:
    """A caller-supplied destination must be exactly what the C side writes: `count` C-contiguous entries of `dtype`."""
    a = arr.array if isinstance(arr, PinnedBuffer) else arr
    if not isinstance(a, np.ndarray) or a.dtype != np.dtype(dtype) or not a.flags["C_CONTIGUOUS"] or a.size != int(count):
        raise CfdError(_abi.CFD_ERR_INVALID_ARGUMENT,
                       f"{what}: need a C-contiguous {np.dtype(dtype).name} array of {int(count)} entries, got "
                       f"{getattr(a, 'dtype', type(a))} x {getattr(a, 'size', '?')}")
    if not a.flags["WRITEABLE"]:
        raise CfdError(_abi.CFD_ERR_INVALID_ARGUMENT, f"{what}: destination is read-only")
    return a


class PinnedBuffer:
    """Page-locked host memory (cfd_host_alloc) viewed as a numpy array; snapshots into it skip the bounce copy.
    The buffer owns the memory: `close()` (or garbage collection of the buffer) frees it, after which `array` and every
    view handed out from it (e.g. inside a `SimSnapshot`) must not be touched any more — copy what has to outlive it."""

    def __init__(self, count: int, dtype=np.float32):
        self._lib = load_library()
        self._ptr = C.c_void_p()
        nbytes = int(count) * np.dtype(dtype).itemsize
        _check(self._lib, self._lib.cfd_host_alloc(nbytes, C.byref(self._ptr)))
        raw = (C.c_char * max(nbytes, 1)).from_address(self._ptr.value)
        self.array = np.frombuffer(raw, dtype=dtype, count=int(count))

    def close(self):
        if self._ptr:
            self.array = None
            self._lib.cfd_host_free(self._ptr)
            self._ptr = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def strip_rows(ny: int, world_size: int, rank: int):
    """Rows [j0, j1) of pressure cells that `rank` of `world_size` owns (the library's strip partition)."""
    lib = load_library()
    j0, j1 = C.c_uint64(), C.c_uint64()
    _check(lib, lib.cfd_strip_rows(int(ny), int(world_size), int(rank), C.byref(j0), C.byref(j1)))
    return int(j0.value), int(j1.value)


def nccl_unique_id() -> bytes:
    lib = load_library()
    buf = C.create_string_buffer(128)
    _check(lib, lib.cfd_nccl_unique_id(buf))
    return buf.raw


def selftest_division(divisor: float, samples: int, seed: int = 1, mode: int = 0):
    """Counts dividends for which the kernels' hoisted-reciprocal division differs from `x / divisor`."""
    lib = load_library()
    bad, fast = C.c_uint64(), C.c_uint64()
    _check(lib, lib.cfd_selftest_division(float(divisor), int(samples), int(seed), int(mode), C.byref(bad),
                                          C.byref(fast)))
    return int(bad.value), int(fast.value)


class Model:
    """The simulation model (reference: `pub struct Model`, src/model.rs:166-214), state resident in HBM."""

    def __init__(self, grid: Grid, params: SimulationParams, precision: int = 64,
                 options: Optional[_abi.CfdOptions] = None):
        self._lib = load_library()
        self._h = C.c_void_p()
        self.grid = grid
        g, p = grid.to_c(), params.to_c()
        opts = options if options is not None else default_options()
        if options is None:
            opts.precision = precision
        self.precision = int(opts.precision)
        self._keepalive = opts
        _check(self._lib, self._lib.cfd_model_create_ex(C.byref(g), C.byref(p), C.byref(opts), C.byref(self._h)))
        self.nx, self.ny = int(grid.nx), int(grid.ny)

    @classmethod
    def new(cls, grid: Grid, params: SimulationParams, **kw) -> "Model":
        """`Model::new(grid, &params)` (src/model.rs:219)."""
        return cls(grid, params, **kw)

    @classmethod
    def strip(cls, grid: Grid, params: SimulationParams, rank: int, world_size: int, unique_id: bytes,
              device: int = -1, precision: int = 64, flags: int = 0, consts=None) -> "Model":
        """One rank of a row-strip decomposition over `world_size` GPUs (one process per GPU).  `unique_id` is
        the 128-byte id from `nccl_unique_id()` on rank 0, broadcast by the launcher.  Import torch BEFORE this
        module in such processes, so that the process shares one libnccl.so.2."""
        if len(unique_id) != 128:
            raise CfdError(_abi.CFD_ERR_INVALID_ARGUMENT, "unique_id must be 128 bytes")
        opts = default_options()
        opts.precision, opts.device, opts.rank, opts.world_size, opts.flags = precision, device, rank, world_size, flags
        if consts is not None:
            opts.consts = consts
        buf = C.create_string_buffer(unique_id, 128)
        opts.nccl_unique_id = C.cast(buf, C.c_void_p)
        m = cls(grid, params, options=opts)
        m._uid_buf = buf
        return m

    # -- lifecycle ---------------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            self._lib.cfd_model_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _handle(self):
        if not self._h:
            raise CfdError(_abi.CFD_ERR_INVALID_ARGUMENT, "model was dropped")
        return self._h

    # -- the reference API ---------------------------------------------------------------------------
    def update(self):
        """`Model::update(&mut self)` (src/model.rs:304-379): one timestep."""
        _check(self._lib, self._lib.cfd_model_update(self._handle()))

    def update_n(self, n: int):
        _check(self._lib, self._lib.cfd_model_update_n(self._handle(), int(n)))

    def set_parameters(self, params: SimulationParams):
        """`Model::set_parameters` (src/model.rs:1250-1257)."""
        p = params.to_c()
        _check(self._lib, self._lib.cfd_model_set_params(self._handle(), C.byref(p)))

    def snapshot_sizes(self):
        n = C.c_uint64()
        sizes = []
        for fid in (_abi.FIELD_P, _abi.FIELD_U, _abi.FIELD_V):  # whole fields, or this rank's rows of a strip run
            _check(self._lib, self._lib.cfd_model_field_len(self._handle(), fid, C.byref(n)))
            sizes.append(int(n.value))
        return sizes

    def pinned_snapshot_buffers(self):
        """Three page-locked f32 buffers sized for `get_snapshot(out=...)`."""
        return [PinnedBuffer(n, np.float32) for n in self.snapshot_sizes()]

    def get_snapshot(self, out=None) -> SimSnapshot:
        """`Model::get_snapshot` (src/model.rs:1259-1267): owned f32 copies of p, u, v in reference layout.
        `out` = three preallocated buffers (`pinned_snapshot_buffers()`) to fill instead of new arrays."""
        sizes = self.snapshot_sizes()
        if out is not None:
            p, u, v = (_checked_out(b, n, np.float32, "get_snapshot") for b, n in zip(out, sizes))
        else:
            p = np.empty(sizes[0], dtype=np.float32)
            u = np.empty(sizes[1], dtype=np.float32)
            v = np.empty(sizes[2], dtype=np.float32)
        dt = C.c_float()
        _check(self._lib, self._lib.cfd_model_get_snapshot(self._handle(), p.ctypes.data, u.ctypes.data,
                                                           v.ctypes.data, C.byref(dt)))
        return SimSnapshot(p=p, u=u, v=v, dt=float(dt.value), paused=False)

    def snapshot_begin(self, out):
        """First half of `get_snapshot`: narrows the current fields on the device and starts their copy into the three
        PINNED buffers `out` (`pinned_snapshot_buffers()`); returns at once, so the copy overlaps the next `update()`s.
        The reference's UI works the same way: it posts `Command::GetSnapshot` and picks the snapshot up on a later frame
        (src/model.rs:100-102, :1300-1306; src/app.rs:95-104).  At most two snapshots in flight."""
        sizes = self.snapshot_sizes()
        p, u, v = (_checked_out(b, n, np.float32, "snapshot_begin") for b, n in zip(out, sizes))
        _check(self._lib, self._lib.cfd_model_snapshot_begin(self._handle(), p.ctypes.data, u.ctypes.data, v.ctypes.data))
        self._snap_pending = getattr(self, "_snap_pending", []) + [(p, u, v, out)]

    def snapshot_end(self) -> SimSnapshot:
        """Second half: waits for the oldest snapshot in flight and returns it (views of the buffers given to `begin`)."""
        dt = C.c_float()
        _check(self._lib, self._lib.cfd_model_snapshot_end(self._handle(), C.byref(dt)))
        p, u, v, _keep = self._snap_pending.pop(0)
        return SimSnapshot(p=p, u=u, v=v, dt=float(dt.value), paused=False)

    def render_rgba(self, mode: int, out=None):
        """The UI's colour map (src/app.rs:235-404) computed on the device: (ny, nx, 4) uint8 RGBA, plus the (min, max)
        of the mapped quantity.  mode 0 pressure, 1 velocity magnitude, 2 vorticity."""
        if out is None:
            out = np.empty(self.nx * self.ny * 4, dtype=np.uint8)
        arr = _checked_out(out, self.nx * self.ny * 4, np.uint8, "render_rgba")
        lo, hi = C.c_float(), C.c_float()
        _check(self._lib, self._lib.cfd_model_render_rgba(self._handle(), int(mode), arr.ctypes.data, C.byref(lo), C.byref(hi)))
        return arr.reshape(self.ny, self.nx, 4), float(lo.value), float(hi.value)

    def get_residuals(self) -> Residuals:
        """`Model::get_residuals` (src/model.rs:1269-1280)."""
        r = _abi.CfdResiduals()
        _check(self._lib, self._lib.cfd_model_get_residuals(self._handle(), C.byref(r)))
        return Residuals.from_c(r)

    # -- tracer particles (the JS twin, index.html:1472-1543) ------------------------------
    def tracers_inject(self):
        """`initTracers` / `injectTracers`: one new tracer per cell row on the inlet."""
        _check(self._lib, self._lib.cfd_model_tracers_inject(self._handle()))

    def tracers_update(self, dt: float):
        """`updateTracers(dt)`: advect with the current fields, drop the tracers that left the domain."""
        _check(self._lib, self._lib.cfd_model_tracers_update(self._handle(), float(dt)))

    def tracers(self) -> np.ndarray:
        """(n, 2) float64 array of tracer positions, in injection order."""
        n = C.c_uint64()
        _check(self._lib, self._lib.cfd_model_tracers_count(self._handle(), C.byref(n)))
        out = np.empty((int(n.value), 2), dtype=np.float64)
        _check(self._lib, self._lib.cfd_model_tracers_get(self._handle(), out.ctypes.data, n.value, C.byref(n)))
        return out

    def tracers_clear(self):
        _check(self._lib, self._lib.cfd_model_tracers_clear(self._handle()))

    def run(self) -> "SimulationControlHandle":
        """`Model::run(self)` (src/model.rs:1282-1332): moves the model into one solver thread."""
        return SimulationControlHandle(self)

    # -- parity / measurement hooks --------------------------------------------------------------------
    def field(self, fid: int) -> np.ndarray:
        n = C.c_uint64()
        _check(self._lib, self._lib.cfd_model_field_len(self._handle(), fid, C.byref(n)))
        out = np.empty(n.value, dtype=np.float64)
        _check(self._lib, self._lib.cfd_model_get_field_f64(self._handle(), fid, out.ctypes.data, n.value))
        return out

    def set_field(self, fid: int, values: np.ndarray):
        a = np.ascontiguousarray(values, dtype=np.float64)
        _check(self._lib, self._lib.cfd_model_set_field_f64(self._handle(), fid, a.ctypes.data, a.size))

    def rows(self):
        j0, j1 = C.c_uint64(), C.c_uint64()
        _check(self._lib, self._lib.cfd_model_rows(self._handle(), C.byref(j0), C.byref(j1)))
        return int(j0.value), int(j1.value)

    def profile_smoother(self, enable: bool):
        _check(self._lib, self._lib.cfd_model_profile_smoother(self._handle(), 1 if enable else 0))

    def last_smoother_timing(self):
        ms, n = C.c_double(), C.c_uint64()
        _check(self._lib, self._lib.cfd_model_last_smoother_timing(self._handle(), C.byref(ms), C.byref(n)))
        return float(ms.value), int(n.value)

    def last_timing(self):
        step_ms, sweep_ms, launches = C.c_double(), C.c_double(), C.c_uint64()
        _check(self._lib, self._lib.cfd_model_last_timing(self._handle(), C.byref(step_ms), C.byref(sweep_ms),
                                                          C.byref(launches)))
        return float(step_ms.value), float(sweep_ms.value), int(launches.value)


class _Command:
    STOP, GET_SNAPSHOT, SET_PARAMS, PAUSE, RESUME = range(5)


class SimulationControlHandle:
    """`SimulationControlHandle` (src/model.rs:65-117) + the solver thread of `Model::run` (:1287-1325).

    Three queues stand in for the three mpsc channels.  Unlike the reference — whose `Command::Stop` only
    leaves the inner `for` and whose thread dies by panicking on a closed channel (:1296, :1319) — `stop()`
    ends the thread and releases the device memory.
    """

    def __init__(self, model: Model):
        self._commands: "queue.Queue" = queue.Queue()
        self._snapshots: "queue.Queue" = queue.Queue()
        self._residuals: "queue.Queue" = queue.Queue()
        self._model = model
        self._thread = threading.Thread(target=self._loop, name="cfd-solver", daemon=True)
        self._thread.start()

    def _loop(self):
        model, paused = self._model, False
        while True:
            snapshot_sent = False
            stop = False
            while True:  # command_receiver.try_iter()
                try:
                    cmd, arg = self._commands.get_nowait()
                except queue.Empty:
                    break
                if cmd == _Command.STOP:
                    stop = True
                    break
                if cmd == _Command.SET_PARAMS:
                    model.set_parameters(arg)
                elif cmd == _Command.GET_SNAPSHOT:
                    if not snapshot_sent:
                        snap = model.get_snapshot()
                        snap.paused = paused
                        self._snapshots.put(snap)
                        snapshot_sent = True
                elif cmd == _Command.PAUSE:
                    paused = True
                elif cmd == _Command.RESUME:
                    paused = False
            if stop:
                break
            if not paused:
                model.update()
                self._residuals.put(model.get_residuals())
            else:
                time.sleep(0.016)
        model.close()

    def stop(self):
        self._commands.put((_Command.STOP, None))
        self._thread.join()

    def get_last_available_snapshot(self) -> Optional[SimSnapshot]:
        last = None
        while True:
            try:
                last = self._snapshots.get_nowait()
            except queue.Empty:
                return last

    def get_new_log_messages(self) -> List[Residuals]:
        out = []
        while True:
            try:
                out.append(self._residuals.get_nowait())
            except queue.Empty:
                return out

    def request_snapshot(self):
        self._commands.put((_Command.GET_SNAPSHOT, None))

    def set_params(self, params: SimulationParams):
        self._commands.put((_Command.SET_PARAMS, params))

    def pause(self):
        self._commands.put((_Command.PAUSE, None))

    def resume(self):
        self._commands.put((_Command.RESUME, None))
