"""numpy_restatement.py — a SECOND, independently written CPU restatement of the reference hot path (TEST
INFRASTRUCTURE, not product code): `Model::update` of src/model.rs, written row-vectorised in
numpy straight from the Rust source, with the dtype as a parameter (float32 = the reference's arithmetic).

Purpose: cross-check the C++ oracle (cfd_oracle.hpp), which was transliterated loop by loop.  The two share no code
and are structured differently (whole rows here, 8-lane chunks + scalar tails there); tests/test_oracle_kat.py
requires them to agree bit for bit.  PARITY UNPINNED all the same: neither has been compared with outputs of the
Rust reference (no Rust toolchain here, no golden vectors in the reference).

Scope: nx % 8 == 0 (the reference panics otherwise), both velocity schemes (FirstOrder :540-552, :590-632;
SecondOrder :553-579, :634-669 with the scalar face functions :911-1053, :1098-1248), both inlet profiles, optional
cylinder.  Every numpy expression keeps the Rust association; numpy evaluates each elementwise
operation as a single IEEE op of the array dtype (no FMA).  Flat indexing is kept (`u[i + j*(nx+1)]`) so that the
reference's reads past the end of a row land in the next row exactly as in the Rust.
"""
import numpy as np

LANES = 8  # src/model.rs:11


class NumpyModel:
    def __init__(self, grid, params, dtype=np.float32):
        T = self.T = np.dtype(dtype).type
        self.nx, self.ny = int(grid.nx), int(grid.ny)
        nx, ny = self.nx, self.ny
        assert nx % LANES == 0
        f = np.float32
        self.dx, self.dy, self.ly = T(f(grid.dx)), T(f(grid.dy)), T(f(grid.ly))
        self.dt, self.nu = T(f(params.dt)), T(f(params.viscosity))            # :265-266
        self.target = T(f(params.target_inlet_velocity))
        self.parabolic = int(params.inlet_profile) == 1
        assert int(params.pressure_solver) == 0 and int(params.scenario) == 0
        self.second = int(params.velocity_scheme) == 1
        self.current = T(0)
        self.step, self.ramp, self.time = 0, 100, T(0)                        # :267-269
        self.u = np.zeros((nx + 1) * ny, dtype)                               # :223-229
        self.v = np.zeros(nx * (ny + 1), dtype)
        self.p = np.zeros(nx * ny, dtype)
        self.u_star, self.v_star = self.u.copy(), self.v.copy()
        self.rhs, self.pp, self.pp_new = self.p.copy(), self.p.copy(), self.p.copy()
        self.mask_u = np.zeros(self.u.size, bool)
        self.mask_v = np.zeros(self.v.size, bool)
        self.solid = []
        if grid.obstacle is not None:                                         # :235-261, all f32
            cyl = grid.obstacle
            x = (np.arange(nx, dtype=f) + f(0.5)) * f(grid.dx)
            y = (np.arange(ny, dtype=f) + f(0.5)) * f(grid.dy)
            ddx, ddy = x[None, :] - f(cyl.center_x), y[:, None] - f(cyl.center_y)
            inside = np.sqrt(ddx * ddx + ddy * ddy) < f(cyl.radius)
            jj, ii = np.nonzero(inside)
            self.solid = list(zip(ii.tolist(), jj.tolist()))
            for i, j in self.solid:
                if i > 0:
                    self.mask_u[i + j * (nx + 1)] = True
                self.mask_u[i + 1 + j * (nx + 1)] = True
                if j > 0:
                    self.mask_v[i + j * nx] = True
                self.mask_v[i + (j + 1) * nx] = True
        self.last_p = self.last_u = self.last_v = T(0)
        self.K = self.S = 0

    @staticmethod
    def _g(a, idx):
        """gather with clamped indices: np.where evaluates both branches, the Rust only the one it takes"""
        return a[np.clip(idx, 0, a.size - 1)]

    # ---- predictor (:538-580 + :382-436; :586-670 + :439-521) ----
    def predictor_u(self, dt):
        nx, ny, W, T = self.nx, self.ny, self.nx + 1, self.T
        u, v, g = self.u, self.v, self._g
        i = np.arange(1, nx + 1)                    # chunks 1, 9, ... cover columns 1..nx when nx % 8 == 0
        for j in range(1, ny - 1):
            idx = i + j * W
            v_n, v_s = v[i + (j + 1) * nx], v[i + j * nx]                     # get_v_north / south :1056-1069
            uc = u[idx]
            ur = u[idx + 1]
            ul = u[idx - 1]
            if self.second:                                                   # scalar face functions, per column (:564-569)
                h, q = T(0.5), T(1.5)
                vn_s = h * (v[(i - 1) + (j + 1) * nx] + v[i + (j + 1) * nx])  # get_v_north_scalar :984-989 (i > 0)
                un_pos = (q * uc - h * u[idx - W]) if j > 1 else uc           # :992-1008
                far = i + (j + 2) * W
                un_neg = np.where((far < u.size) & (j < ny - 1), q * u[idx + W] - h * g(u, far), u[idx + W])
                u_n = np.where(vn_s >= 0, un_pos, un_neg)
                vs_s = h * (v[(i - 1) + j * nx] + v[i + j * nx])              # get_v_south_scalar :1029-1034
                us_pos = (q * u[idx - W] - h * g(u, idx - 2 * W)) if j > 1 else u[idx - W]      # :1037-1053
                us_neg = q * uc - h * u[idx + W]                              # j < ny always
                u_s = np.where(vs_s >= 0, us_pos, us_neg)
                ue_pos = np.where(i > 1, q * uc - h * ul, uc)                 # :911-926
                ue_neg = np.where(((idx + 2) < u.size) & (i < nx - 1), q * ur - h * g(u, idx + 2), ur)
                u_e = np.where(uc >= 0, ue_pos, ue_neg)
                uw_pos = np.where(i > 2, q * ul - h * g(u, idx - 2), ul)      # :944-963
                uw_neg = np.where(i < nx, q * uc - h * ur, uc)
                u_w = np.where(ul >= 0, uw_pos, uw_neg)
            else:
                u_n = np.where(v_n >= 0, uc, u[idx + W])                      # :966-981
                u_s = np.where(v_s >= 0, u[idx - W], uc)                      # :1011-1026
                u_e = np.where((uc + ur) * T(0.5) >= 0, uc, ur)               # :893-908
                u_w = np.where((ul + uc) * T(0.5) >= 0, ul, uc)               # :929-941
            conv = (u_e * u_e - u_w * u_w) / self.dx + (v_n * u_n - v_s * u_s) / self.dy      # :408-415
            lap = (ur - T(2) * uc + ul) / (self.dx * self.dx) + (u[idx + W] - T(2) * uc + u[idx - W]) / (self.dy * self.dy)
            res = uc + dt * (-conv + self.nu * lap)                           # :433
            self.u_star[idx] = np.where(self.mask_u[idx], T(0), res)          # :434

    def predictor_v(self, dt):
        nx, ny, W, T = self.nx, self.ny, self.nx + 1, self.T
        u, v, g = self.u, self.v, self._g
        i = np.arange(1, nx)                        # body chunks + the scalar tail reach column nx-1 (:591-620)
        for j in range(1, ny):
            idx = i + j * nx
            ue, uw = u[i + 1 + j * W], u[i + j * W]                           # :600-601, :622-626
            vc, vn_, vs_ = v[idx], v[idx + nx], v[idx - nx]
            if self.second:
                h, q = T(0.5), T(1.5)
                ve_pos = q * vc - h * v[idx - 1]                              # :1098-1113 (i > 0)
                ve_neg = np.where(((idx + 2) < v.size) & (i < nx - 2), q * v[idx + 1] - h * g(v, idx + 2), v[idx + 1])
                v_e = np.where(ue >= 0, ve_pos, ve_neg)
                vw_pos = np.where(i > 1, q * v[idx - 1] - h * g(v, idx - 2), v[idx - 1])       # :1145-1160
                vw_neg = np.where(i < nx - 1, q * vc - h * v[idx + 1], vc)
                v_w = np.where(uw >= 0, vw_pos, vw_neg)
                vn_pos = (q * vc - h * vs_) if j > 1 else vc                  # :1188-1204
                far = i + (j + 2) * nx
                vn_neg = np.where((far < v.size) & (j < ny - 1), q * vn_ - h * g(v, far), vn_)
                v_n = np.where(h * (vc + vn_) >= 0, vn_pos, vn_neg)
                vs_pos = (q * vs_ - h * g(v, idx - 2 * nx)) if j > 1 else vs_ # :1232-1248
                vs_neg = q * vc - h * vn_                                     # j < ny always
                v_s = np.where(h * (vs_ + vc) >= 0, vs_pos, vs_neg)
                # the lane of column nx-1 keeps zero fluxes (`break` at :647-650) but is still written (:456-496)
                last = i >= nx - 1
                z = T(0)
                ue, uw, v_e, v_w, v_n, v_s = (np.where(last, z, a) for a in (ue, uw, v_e, v_w, v_n, v_s))
            else:
                v_n = np.where((vc + vn_) * T(0.5) >= 0, vc, vn_)             # :1163-1185
                v_s = np.where((vc + vs_) * T(0.5) >= 0, vs_, vc)             # :1207-1229
                v_e = np.where(ue >= 0, vc, v[idx + 1])                       # :1073-1095
                v_w = np.where(uw >= 0, v[idx - 1], vc)                       # :1116-1142
            conv = (ue * v_e - uw * v_w) / self.dx + (v_n * v_n - v_s * v_s) / self.dy        # :501-505
            lap = (v[idx + 1] - T(2) * vc + v[idx - 1]) / (self.dx * self.dx) + (vn_ - T(2) * vc + vs_) / (self.dy * self.dy)
            res = vc + dt * (-conv + self.nu * lap)
            self.v_star[idx] = np.where(self.mask_v[idx], T(0), res)

    def divergence(self, dt):                                                  # :1406-1440
        nx, ny, W = self.nx, self.ny, self.nx + 1
        us = self.u_star.reshape(ny, W)
        vs = self.v_star.reshape(ny + 1, nx)
        self.rhs = (((us[:, 1:] - us[:, :-1]) / self.dx + (vs[1:, :] - vs[:-1, :]) / self.dy) / dt).ravel()

    def jacobi(self):                                                          # :734-824
        nx, ny, T = self.nx, self.ny, self.T
        omega, tol = T(0.75), T(1e-4)
        one_minus = T(1.0) - omega
        dx_sq, dy_sq = self.dx * self.dx, self.dy * self.dy
        denom = T(2.0) / (self.dx * self.dx) + T(2.0) / (self.dy * self.dy)
        # chunks start at 1, 9, ...; a chunk is SIMD iff i + 8 <= nx - 1 (:755); with nx % 8 == 0 the body is 1..nx-8
        body = 1 + ((nx - 2) // LANES) * LANES
        max_error = T(0)
        for _ in range(50):
            p = self.pp.reshape(ny, nx)
            r = self.rhs.reshape(ny, nx)
            c = p[1:-1, 1:]
            horizontal = (np.concatenate([p[1:-1, 2:], self.pp[(np.arange(1, ny - 1) + 1) * nx][:, None]], axis=1) + p[1:-1, :-1]) / dx_sq
            vertical = (p[2:, 1:] + p[:-2, 1:]) / dy_sq
            new = omega * ((horizontal + vertical - r[1:-1, 1:]) / denom) + one_minus * c
            err = np.abs(new[:, :body - 1] - c[:, :body - 1])                  # body columns 1..body-1 only (:795-798)
            max_error = T(np.fmax.reduce(err, axis=None, initial=0.0))
            pn = self.pp_new.reshape(ny, nx)
            pn[1:-1, 1:] = new
            self.pp, self.pp_new = self.pp_new, self.pp                        # :805
            q = self.pp.reshape(ny, nx)
            q[0, :] = q[1, :]                                                  # :807-811
            q[ny - 1, :] = q[ny - 2, :]
            q[:, 0] = q[:, 1]                                                  # :812-815
            q[:, nx - 1] = 0
            self.S += 1
            if max_error < tol:
                break
        self.last_p = max_error
        self.K += 1
        return max_error

    def corrector(self, dt):                                                   # :1334-1404
        nx, ny, W = self.nx, self.ny, self.nx + 1
        pp = self.pp.reshape(ny, nx)
        us, u = self.u_star.reshape(ny, W), self.u.reshape(ny, W)
        diff = pp[:, 1:] - pp[:, :-1]                                          # p'[i] - p'[i-1], i = 1..nx-1
        tail = nx - (nx - 1) % LANES if (nx - 1) % LANES else nx               # first column of the scalar tail (:1338)
        b = tail - 1                                                           # number of body columns (1..tail-1)
        u[:, 1:tail] = us[:, 1:tail] - dt * (diff[:, :b] / self.dx)            # SIMD: dt * ((pR - pL) / dx)   :1358
        u[:, tail:nx] = us[:, tail:nx] - dt * diff[:, b:] / self.dx            # tail: (dt * (pR - pL)) / dx   :1343
        vs, v = self.v_star.reshape(ny + 1, nx), self.v.reshape(ny + 1, nx)
        v[1:ny, :] = vs[1:ny, :] - dt * ((pp[1:, :] - pp[:-1, :]) / self.dy)   # :1366-1390 (no tail when nx % 8 == 0)
        self.p = self.p + self.pp                                              # :1392-1403

    def boundary_conditions(self):                                             # :827-875
        nx, ny, W, T = self.nx, self.ny, self.nx + 1, self.T
        u, v = self.u.reshape(ny, W), self.v.reshape(ny + 1, nx)
        if self.parabolic:
            y = (np.arange(ny).astype(self.u.dtype) + T(0.5)) * self.dy
            center = radius = self.ly / T(2.0)
            q = (y - center) / radius
            val = self.current * (T(1.0) - q * q)
            u[:, 0] = np.where(val < 0, T(0), val)
        else:
            u[:, 0] = self.current
        u[:, nx] = u[:, nx - 1]
        u[0, :] = 0
        u[ny - 1, :] = 0
        v[0, :] = 0
        v[ny, :] = 0
        for i, j in self.solid:
            u[j, i] = 0
            v[j, i] = 0

    def update(self):                                                          # :304-379
        T = self.T
        u_old, v_old = self.u.copy(), self.v.copy()
        if self.step < self.ramp:
            self.current = (T(self.step) / T(self.ramp)) * self.target
        else:
            self.current = self.target
        dt = self.dt / T(1)
        self.K = self.S = 0
        with np.errstate(all="ignore"):
            self.predictor_u(dt)
            self.predictor_v(dt)
            self.divergence(dt)
            self.jacobi()
            self.corrector(dt)
            for _ in range(20):                                                # :696-724
                self.u_star[:] = self.u
                self.v_star[:] = self.v
                self.divergence(dt)
                self.jacobi()
                self.corrector(dt)
                if self.last_p < T(1e-4):
                    break
            self.boundary_conditions()
            self.last_u = T(np.fmax.reduce(np.abs(self.u - u_old), initial=0.0))
            self.last_v = T(np.fmax.reduce(np.abs(self.v - v_old), initial=0.0))
            self.step += 1
            self.time = self.time + self.dt
            max_vel = max(T(np.fmax.reduce(np.abs(self.u), initial=0.0)), T(np.fmax.reduce(np.abs(self.v), initial=0.0)))
            if max_vel == 0:
                new_dt = self.dt
            else:
                dt_cfl = T(0.2) * min(self.dx, self.dy) / max_vel
                new_dt = min(dt_cfl, self.dt)
            prev = self.dt
            self.dt = min(new_dt, prev * T(1.1)) if new_dt > prev else new_dt
