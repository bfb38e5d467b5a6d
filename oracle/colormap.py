"""CPU restatement (TEST INFRASTRUCTURE, not product code) of the reference UI's colour map, src/app.rs:235-404, in
numpy float32 — every operation is a single IEEE f32 op like the Rust.  Checks `cfd_model_render_rgba`.

pressure :238-279, velocity magnitude :281-330, vorticity :332-398; normalisation :239-259 (`max = min + 1` when the
range is below 1e-6), red-blue ramp `(norm * 255) as u8` / `((1 - norm) * 255) as u8` (Rust `as`: truncate, saturate,
NaN -> 0), grey overlay where the cell centre lies within the cylinder (`<=`, :262-268).  Parity unpinned: the
reference has no tests for this code either."""
import numpy as np

f32 = np.float32


def _as_u8(x):
    """Rust `f32 as u8`."""
    x = np.asarray(x, dtype=np.float32)
    out = np.zeros(x.shape, dtype=np.uint8)
    ok = x > 0  # NaN and negatives -> 0
    big = ok & (x >= 255)
    mid = ok & ~big
    out[big] = 255
    out[mid] = np.trunc(x[mid]).astype(np.uint8)
    return out


def mapped_quantity(mode, p, u, v, nx, ny, dx, dy):
    p = np.asarray(p, dtype=np.float32).reshape(ny, nx)
    u = np.asarray(u, dtype=np.float32).reshape(ny, nx + 1)
    v = np.asarray(v, dtype=np.float32).reshape(ny + 1, nx)
    half = f32(0.5)
    if mode == 0:
        return p.copy()
    if mode == 1:  # :287-303
        u_cell = half * (u[:, :-1] + u[:, 1:])
        v_cell = half * (v[:-1, :] + v[1:, :])
        return np.sqrt(u_cell * u_cell + v_cell * v_cell)
    vort = np.zeros((ny, nx), dtype=np.float32)  # :337-355, interior cells only
    uc = half * (u[:, :-1] + u[:, 1:])           # 0.5 * (u[i,j] + u[i+1,j]) for every cell
    vc = half * (v[:-1, :] + v[1:, :])           # 0.5 * (v[i,j] + v[i,j+1])
    du_dy = (uc[2:, 1:-1] - uc[1:-1, 1:-1]) / f32(dy)    # (u_top - u_bottom) / dy with u_top on row j+1
    dv_dx = (vc[1:-1, 2:] - vc[1:-1, 1:-1]) / f32(dx)    # (v_right - v_left) / dx with v_right on column i+1
    vort[1:-1, 1:-1] = dv_dx - du_dy
    return vort


def render(mode, p, u, v, grid):
    """Returns ((ny, nx, 4) uint8 RGBA, min, max)."""
    nx, ny = int(grid.nx), int(grid.ny)
    with np.errstate(all="ignore"):
        val = mapped_quantity(mode, p, u, v, nx, ny, grid.dx, grid.dy)
        finite_or_inf = val[~np.isnan(val)]
        lo = f32(finite_or_inf.min()) if finite_or_inf.size else f32(np.inf)
        hi = f32(finite_or_inf.max()) if finite_or_inf.size else f32(-np.inf)
        min_val, max_val = lo, hi
        if abs(f32(max_val - min_val)) < f32(1e-6):
            max_val = f32(min_val + f32(1.0))
        norm = (val - min_val) / f32(max_val - min_val)
        img = np.zeros((ny, nx, 4), dtype=np.uint8)
        img[..., 0] = _as_u8(norm * f32(255.0))
        img[..., 2] = _as_u8((f32(1.0) - norm) * f32(255.0))
        img[..., 3] = 255
        cyl = getattr(grid, "obstacle", None)
        if cyl is not None:
            x = (np.arange(nx, dtype=np.float32) + half_f32()) * f32(grid.dx)
            y = (np.arange(ny, dtype=np.float32) + half_f32()) * f32(grid.dy)
            ddx = x[None, :] - f32(cyl.center_x)
            ddy = y[:, None] - f32(cyl.center_y)
            inside = np.sqrt(ddx * ddx + ddy * ddy) <= f32(cyl.radius)
            img[inside, 0:3] = 128
    return img, float(lo), float(hi)


def half_f32():
    return f32(0.5)
