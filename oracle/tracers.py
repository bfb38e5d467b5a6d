"""CPU ORACLE (test infrastructure, not product code) for the tracer particles of the JS twin — a numpy restatement of
/root/reference/index.html:1472-1543: initTracers / injectTracers (:1475-1483, :1537-1543), updateTracers (:1485-1497),
getVelocityAt (:1499-1526).  The JS computes in double on its Float32Array fields; here the fields are whatever the
caller passes (float64 arrays of the model's u, v), all tracer arithmetic in float64, one numpy operation per JS
operation, in the JS's order.  "Parity unpinned": node is absent, the JS cannot be run here."""
import numpy as np


def inject(tracers: np.ndarray, ny: int, dy: float) -> np.ndarray:
    """injectTracers: append one tracer per cell row on the inlet, at (0, (j + 0.5) * dy)."""
    j = np.arange(ny, dtype=np.float64)
    new = np.stack([np.zeros(ny), (j + 0.5) * dy], axis=1)
    return np.concatenate([tracers.reshape(-1, 2), new], axis=0)


def velocity_at(u, v, nx, ny, dx, dy, x, y):
    """getVelocityAt: bilinear interpolation of the cell-centred velocity."""
    u = np.asarray(u, dtype=np.float64).reshape(ny, nx + 1)
    v = np.asarray(v, dtype=np.float64).reshape(ny + 1, nx)
    i = np.floor(x / dx)
    j = np.floor(y / dy)
    i = np.where(np.isnan(i), 0, i)  # (int) of NaN: such tracers are dropped by the bounds test anyway
    j = np.where(np.isnan(j), 0, j)
    i = np.clip(i, 0, nx - 2).astype(np.int64)
    j = np.clip(j, 0, ny - 2).astype(np.int64)
    rx = (x - i * dx) / dx
    ry = (y - j * dy) / dy

    def cu(ii, jj):
        return 0.5 * (u[jj, ii] + u[jj, ii + 1])

    def cv(ii, jj):
        return 0.5 * (v[jj, ii] + v[jj + 1, ii])

    u00, u10, u01, u11 = cu(i, j), cu(i + 1, j), cu(i, j + 1), cu(i + 1, j + 1)
    v00, v10, v01, v11 = cv(i, j), cv(i + 1, j), cv(i, j + 1), cv(i + 1, j + 1)
    ui = (1 - rx) * ((1 - ry) * u00 + ry * u01) + rx * ((1 - ry) * u10 + ry * u11)
    vi = (1 - rx) * ((1 - ry) * v00 + ry * v01) + rx * ((1 - ry) * v10 + ry * v11)
    return ui, vi


def update(tracers, u, v, nx, ny, dx, dy, lx, ly, dt):
    """updateTracers(dt): explicit Euler step, then keep (in order) the tracers inside [0, lx] x [0, ly]."""
    t = tracers.reshape(-1, 2).astype(np.float64)
    if t.shape[0] == 0:
        return t
    with np.errstate(invalid="ignore"):
        ui, vi = velocity_at(u, v, nx, ny, dx, dy, t[:, 0], t[:, 1])
        x = t[:, 0] + dt * ui
        y = t[:, 1] + dt * vi
        keep = (x >= 0) & (x <= lx) & (y >= 0) & (y <= ly)
    return np.stack([x[keep], y[keep]], axis=1)
