"""Throw-away numpy prototype of the multigrid-preconditioned CG used by Mode C (design study, not product).

Unknowns: p' on columns 1..nx-2, rows 1..ny-2 (mx x my); operator = the 5-point problem the reference's Jacobi
relaxes, with mirror (Neumann) sides and a zero (Dirichlet) outlet column for the channel.  Coarsening: pairs of
cells per direction (the last aggregate is a single cell when the count is odd), finite-volume coarse operators
on the resulting non-uniform tensor grid, piecewise-constant or linear transfer, damped-Jacobi smoothing.
"""
import sys
import time

import numpy as np


class Level:
    def __init__(self, wx, hy, cx, cy, dirichlet_right):
        self.wx, self.hy = wx, hy          # widths / heights of the cells in finest-cell units
        mx, my = len(wx), len(hy)
        self.mx, self.my = mx, my
        # link factors: 1 / (centre distance in finest-cell units)
        self.ax = 1.0 / (0.5 * (wx[:-1] + wx[1:])) if mx > 1 else np.zeros(0)
        self.ay = 1.0 / (0.5 * (hy[:-1] + hy[1:])) if my > 1 else np.zeros(0)
        self.cx, self.cy = cx, cy
        self.dir = (1.0 / (0.5 * wx[-1] + 0.5)) if dirichlet_right else 0.0
        # x-link weight between (I,J),(I+1,J) = cx*hy[J]*ax[I]; y-link = cy*wx[I]*ay[J]
        we = np.zeros(mx); we[:-1] = self.ax; we[-1] = self.dir
        ww = np.zeros(mx); ww[1:] = self.ax
        wn = np.zeros(my); wn[:-1] = self.ay
        ws = np.zeros(my); ws[1:] = self.ay
        self.we, self.ww, self.wn, self.ws = we, ww, wn, ws
        self.diag = cx * hy[:, None] * (we + ww)[None, :] + cy * wx[None, :] * (wn + ws)[:, None]
        # "full" diagonal as the reference uses on the finest level (all four links counted)
        self.diag_full = None

    def apply(self, e):
        r = self.diag * e
        cxh = self.cx * self.hy[:, None]
        cyw = self.cy * self.wx[None, :]
        r[:, :-1] -= cxh * self.ax[None, :] * e[:, 1:]
        r[:, 1:] -= cxh * self.ax[None, :] * e[:, :-1]
        r[:-1, :] -= cyw * self.ay[:, None] * e[1:, :]
        r[1:, :] -= cyw * self.ay[:, None] * e[:-1, :]
        return r


def coarsen_1d(w):
    n = len(w)
    m = (n + 1) // 2
    out = np.zeros(m)
    out[: n // 2] = w[0:2 * (n // 2):2] + w[1:2 * (n // 2):2]
    if n % 2:
        out[-1] = w[-1]
    return out


def build(nx, ny, dx, dy, cavity, min_size=1):
    cx, cy = 1.0 / (dx * dx), 1.0 / (dy * dy)
    wx, hy = np.ones(nx - 2), np.ones(ny - 2)
    levels = [Level(wx, hy, cx, cy, not cavity)]
    while max(len(wx), len(hy)) > min_size:
        wx, hy = coarsen_1d(wx), coarsen_1d(hy)
        levels.append(Level(wx, hy, cx, cy, not cavity))
    return levels


def restrict_sum(r, mxc, myc):
    my, mx = r.shape
    out = np.zeros((myc, mxc))
    fy, fx = my // 2, mx // 2
    out[:fy, :fx] = r[0:2 * fy:2, 0:2 * fx:2] + r[0:2 * fy:2, 1:2 * fx:2] + r[1:2 * fy:2, 0:2 * fx:2] + r[1:2 * fy:2, 1:2 * fx:2]
    if mx % 2:
        out[:fy, -1] = r[0:2 * fy:2, -1] + r[1:2 * fy:2, -1]
    if my % 2:
        out[-1, :fx] = r[-1, 0:2 * fx:2] + r[-1, 1:2 * fx:2]
    if mx % 2 and my % 2:
        out[-1, -1] = r[-1, -1]
    return out


def prolong_const(ec, mx, my):
    return np.repeat(np.repeat(ec, 2, axis=0), 2, axis=1)[:my, :mx]


def vcycle(levels, l, r, nu, omega, full_diag_fine):
    L = levels[l]
    if L.mx == 1 and L.my == 1:
        d = L.diag[0, 0]
        return r / d if d > 0 else np.zeros_like(r)
    dinv = omega / L.diag
    if l == 0 and full_diag_fine:
        dinv = np.full_like(L.diag, omega / (2 * L.cx + 2 * L.cy))
    e = dinv * r
    for _ in range(nu - 1):
        e = e + dinv * (r - L.apply(e))
    res = r - L.apply(e)
    C = levels[l + 1]
    rc = restrict_sum(res, C.mx, C.my)
    ec = vcycle(levels, l + 1, rc, nu, omega, full_diag_fine)
    e = e + prolong_const(ec, L.mx, L.my)
    for _ in range(nu):
        e = e + dinv * (r - L.apply(e))
    return e


def pcg(levels, b, tol_rel, nu, omega, full_diag_fine, maxit=200, precond=True):
    A = levels[0]
    x = np.zeros_like(b)
    r = b.copy()
    z = vcycle(levels, 0, r, nu, omega, full_diag_fine) if precond else r
    d = z.copy()
    rz = float((r * z).sum())
    r0 = np.sqrt(float((r * r).sum()))
    hist = []
    for it in range(1, maxit + 1):
        q = A.apply(d)
        alpha = rz / float((d * q).sum())
        x += alpha * d
        r -= alpha * q
        rn = np.sqrt(float((r * r).sum()))
        hist.append(rn / r0)
        if rn <= tol_rel * r0:
            return x, it, hist
        z = vcycle(levels, 0, r, nu, omega, full_diag_fine) if precond else r
        rz_new = float((r * z).sum())
        d = z + (rz_new / rz) * d
        rz = rz_new
    return x, maxit, hist


if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
    cavity = (sys.argv[2] == "cavity") if len(sys.argv) > 2 else True
    nu = int(sys.argv[3]) if len(sys.argv) > 3 else 2
    omega = float(sys.argv[4]) if len(sys.argv) > 4 else 0.8
    rng = np.random.default_rng(1)
    nx = ny = n
    dx = dy = 1.0 / n
    levels = build(nx, ny, dx, dy, cavity)
    print("levels", [(L.mx, L.my) for L in levels])
    b = rng.standard_normal((ny - 2, nx - 2))
    # lid-cavity-like: divergence concentrated under the lid + smooth part
    yy, xx = np.mgrid[0:ny - 2, 0:nx - 2] / float(n)
    b = 0.1 * b + np.sin(3 * xx) * np.cos(2 * yy)
    b[-1, :] += 50.0 * np.sign(xx[-1, :] - 0.5)
    if cavity:
        b -= b.mean()
    for full in (False, True):
        t0 = time.time()
        x, it, hist = pcg(levels, b, 1e-8, nu, omega, full)
        print(f"n={n} cavity={cavity} nu={nu} omega={omega} full_diag_fine={full}: iterations {it}  ({time.time() - t0:.1f}s) "
              f"last factors {[round(hist[k + 1] / hist[k], 3) for k in range(max(0, len(hist) - 4), len(hist) - 1)]}")
