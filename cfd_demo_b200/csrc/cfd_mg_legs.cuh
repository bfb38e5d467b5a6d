// cfd_mg_legs.cuh — the V(nu,nu)-cycle of the Mode C fast path (cfd_mg.cuh) with each level's descending and ascending
// leg as ONE launch.  EXTENSION (no reference counterpart; the level-0 smoother is the reference's damped-Jacobi
// update, src/model.rs:788-793 with its boundary rules :807-815, through the same jacobi_cell as k_jacobi_sweep5).
//
//   descending leg of level l:  x_1 = sweep(0), x_k = sweep(x_{k-1}) (k <= nu), rho_{l+1} = restrict(rho_l - L x_nu)
//        separate kernels: 2 + 3 (nu - 1) + 2.25 field passes, nu + 1 launches     here: read rho_l, write x_nu, rho_{l+1}: 2.25
//   ascending leg of level l:   x' = x_nu + prolong(e_{l+1}), nu sweeps [level 0: + rho.z]
//        separate kernels: 2.25 + 3 nu passes, nu + 1 launches (+ 1 reduce)         here: read x_nu, rho_l, e_{l+1}, write: 3.25
//
// A block owns a tile of TX x TY cells and stages the tile plus a ring of nu cells in shared memory; sweep k is
// recomputed on the ring cells the later sweeps still need (tile + nu + 1 - k descending, tile + nu - k ascending), so
// blocks never exchange anything and the memory traffic of a leg does not depend on nu.  Every cell goes through the
// SAME per-cell expressions as the one-operation-per-launch kernels (jacobi_cell / mg_lap on level 0, mgc_cell =
// mgc_sweep_cell's arithmetic on the coarse levels; no FMA, same association), so each leg is bit-identical to the
// sequence it replaces (CFD_FLAG_MG_UNFUSED keeps that one for the cross-check); only rho.z is summed in another
// order.  Divisions: DivTry per thread, recomputed with the compiler's division (DivSlow) by the thread whose window test failed.
//
// Three forms live in this file: the flat-indexed kernels right below (k_mg0_down/_up, k_mgc_down/_up: the plain statement of
// the idea, CFD_MG_LEGS=1), column strips (k_mg0_down2/_up2, CFD_MG_LEGS=2) and the register-tiled kernels that ship
// (k_mg0_down3/_up3, k_mgc_down3/_up3: section "register-tiled form"); the first two are kept for A/B and as cross-checks.
// What ncu says about each: profiles/r2_legs_ncu.md.
#pragma once

#include <type_traits>

namespace cfdk {

constexpr int kLegThreads = 256;

// the legs' rare exact path: the compiler's own division, out of line (a call per division instead of ~35 inlined
// instructions per division in every unrolled stage: the fall-back code would otherwise double the kernels' size)
template <class R>
struct DivSlow {
  __device__ __forceinline__ R operator()(R x, const DivG<R>& d) const { return x / d.y; }
};
template <>
struct DivSlow<double> {
  __device__ __forceinline__ double operator()(double x, const DivG<double>& d) const { return div_true(x, d.y); }
};

template <int I, int N, class F>
__device__ __forceinline__ void leg_static_for(F&& f) {
  if constexpr (I < N) {
    f(std::integral_constant<int, I>{});
    leg_static_for<I + 1, N>(f);
  }
}

// f(rx, ry) for every cell of the region (tile + NU) that lies at least M cells inside it, thread-strided
template <int RX, int RY, int M, class F>
__device__ __forceinline__ void leg_for_inset(int tid, F&& f) {
  constexpr int W = RX - 2 * M, H = RY - 2 * M;
#pragma unroll
  for (int k0 = 0; k0 < W * H; k0 += kLegThreads) {
    const int k = k0 + tid;
    if (k < W * H) {
      const int y = k / W, x = k - y * W;
      f(x + M, y + M);
    }
  }
}

template <class R, int TX, int TY, int NU>
struct LegSmem {
  static constexpr int RX = TX + 2 * NU, RY = TY + 2 * NU;
  R rho[RY][RX];
  R a[RY][RX], b[RY][RX];  // the sweeps alternate between them
};

// ---- level 0 -------------------------------------------------------------------------------------------------------
// Region cell (rx, ry) <-> array cell (i0 - NU + rx, j0 - NU + ry); the Jacobi boundary rules make a ring cell the image
// of the unknown next to it (column nx-1 of the channel: zero), so a ring value is formed by evaluating the unknown it
// mirrors: q = clamp(p).
struct LegGeom0 {
  int nx, ny, cavity, i0, j0, j_max;  // j_max: last array row that may be read (strips: the upper halo rows)
  __device__ __forceinline__ int qi(int i) const { return min(max(i, 1), nx - 2); }
  __device__ __forceinline__ int qj(int j) const { return min(min(max(j, 1), ny - 2), j_max); }
};

template <class R, int NU>
__device__ __forceinline__ LegGeom0 leg_geom0(const MgFine<R>& c, int TX, int TY) {
  LegGeom0 g;
  g.nx = c.nx; g.ny = c.ny; g.cavity = c.cavity;
  g.i0 = 1 + (int)blockIdx.x * TX; g.j0 = c.row_lo + (int)blockIdx.y * TY;
  g.j_max = c.row_hi + NU - 1 < c.ny - 1 ? c.row_hi + NU - 1 : c.ny - 1;
  return g;
}

// one Jacobi sweep of level 0 at the region cell (rx, ry): evaluated at the unknown it mirrors
template <class R, int NU, class Src, class Rho>
__device__ __forceinline__ R leg_sweep0(const LegGeom0& g, const JacobiConsts2<R>& c2, const Src& src, const Rho& rho, int rx, int ry) {
  const int i = g.i0 - NU + rx;
  const int ax = g.qi(i) - (g.i0 - NU), ay = g.qj(g.j0 - NU + ry) - (g.j0 - NU);
  R v = jacobi_cell<R>(c2, src[ay][ax - 1], src[ay][ax + 1], src[ay + 1][ax], src[ay - 1][ax], src[ay][ax], rho[ay][ax]);
  if (!g.cavity && i == g.nx - 1) v = R(0);
  return v;
}

template <class R, int TX, int TY, int NU>
__global__ void __launch_bounds__(kLegThreads) k_mg0_down(const MgFine<R> c, const JacobiConsts2<R> c2,
                                                           const R* __restrict__ rho, R* __restrict__ zout, int cmx,
                                                           R* __restrict__ crho, const MgScalars* __restrict__ sc) {
  using S = LegSmem<R, TX, TY, NU>;
  constexpr int RX = S::RX, RY = S::RY;
  extern __shared__ __align__(16) unsigned char leg_smem_raw[];
  S& s = *reinterpret_cast<S*>(leg_smem_raw);
  if (sc->done) return;
  const LegGeom0 g = leg_geom0<R, NU>(c, TX, TY);
  const int tid = threadIdx.x, nx = c.nx, ny = c.ny;
  const int j_end = min(g.j0 + TY, c.row_hi);  // owned unknown rows of the tile: [j0, j_end)
  // rho on the tile + NU (mirrored where the region leaves the unknowns)
  leg_for_inset<RX, RY, 0>(tid, [&](int rx, int ry) {
    s.rho[ry][rx] = rho[(size_t)g.qi(g.i0 - NU + rx) + (size_t)g.qj(g.j0 - NU + ry) * nx];
  });
  __syncthreads();
  // x_1 = first smoothing sweep applied to z = 0 (k_mg_first_sweep's expression), pointwise
  auto stage1 = [&](auto& dv) {
    leg_for_inset<RX, RY, 0>(tid, [&](int rx, int ry) {
      R v = c2.omega * dv((R(0) + R(0)) - s.rho[ry][rx], c2.denom) + c2.one_minus_omega * R(0);
      if (!g.cavity && g.i0 - NU + rx == nx - 1) v = R(0);
      s.a[ry][rx] = v;
    });
  };
  {
    DivTry<R> fast(c2.denom);
    stage1(fast);
    if (__builtin_expect(!fast.ok(), 0)) {
      DivSlow<R> exact;
      stage1(exact);
    }
  }
  __syncthreads();
  // x_k = sweep(x_{k-1}) on the tile + NU + 1 - k; the last one also goes to memory (the tile's cells and the ring cells
  // that mirror them)
  leg_static_for<2, NU + 1>([&](auto kc) {
    constexpr int K = decltype(kc)::value;
    auto& src = (K % 2 == 0) ? s.a : s.b;
    auto& dst = (K % 2 == 0) ? s.b : s.a;
    leg_for_inset<RX, RY, K - 1>(tid, [&](int rx, int ry) {
      const R v = leg_sweep0<R, NU>(g, c2, src, s.rho, rx, ry);
      dst[ry][rx] = v;
      if (K == NU) {
        const int i = g.i0 - NU + rx, j = g.j0 - NU + ry;
        const int qi = g.qi(i), qj = g.qj(j);
        if (i <= nx - 1 && j <= ny - 1 && qi >= g.i0 && qi < g.i0 + TX && qj >= g.j0 && qj < j_end) zout[(size_t)i + (size_t)j * nx] = v;
      }
    });
    __syncthreads();
  });
  auto& xn = (NU % 2 == 0) ? s.b : s.a;
  // rho_1 = sum over the children of (rho - L x_nu)  (k_mg_fine_restrict's order: b outer, a inner)
  constexpr int CX = TX / 2, CY = TY / 2;
  auto stage3 = [&](auto& dv) {
#pragma unroll
    for (int k0 = 0; k0 < CX * CY; k0 += kLegThreads) {
      const int k = k0 + tid;
      if (k < CX * CY) {
        const int cy = k / CX, cx = k - cy * CX;
        const int I = (int)blockIdx.x * CX + cx, J = (g.j0 - 1) / 2 + cy;
        if (1 + 2 * I <= nx - 2 && 1 + 2 * J < j_end) {
          R acc = R(0);
#pragma unroll
          for (int b = 0; b < 2; ++b)
#pragma unroll
            for (int a = 0; a < 2; ++a) {
              const int i = 1 + 2 * I + a, j = 1 + 2 * J + b;
              if (i <= nx - 2 && j <= ny - 2) {
                const int x = i - (g.i0 - NU), y = j - (g.j0 - NU);
                const R cc = xn[y][x];
                acc += s.rho[y][x] - mg_lap<R>(c, dv, cc, xn[y][x + 1], xn[y][x - 1], xn[y + 1][x], xn[y - 1][x]);
              }
            }
          crho[(size_t)(I + 1) + (size_t)(J + 1) * (cmx + 2)] = acc;
        }
      }
    }
  };
  {
    DivTry<R> fast(c.ddx_sq);
    fast.also(c.ddy_sq);
    stage3(fast);
    if (__builtin_expect(!fast.ok(), 0)) {
      DivSlow<R> exact;
      stage3(exact);
    }
  }
}

// ascending leg of level 0: z = NU sweeps of (x_nu + correction of the parents); rho.z -> one partial per block
template <class R, int TX, int TY, int NU>
__global__ void __launch_bounds__(kLegThreads) k_mg0_up(const MgFine<R> c, const JacobiConsts2<R> c2,
                                                         const R* __restrict__ zin, const R* __restrict__ rho, int cmx,
                                                         const R* __restrict__ ce, R* __restrict__ zout,
                                                         double* __restrict__ partials, const MgScalars* __restrict__ sc) {
  using S = LegSmem<R, TX, TY, NU>;
  constexpr int RX = S::RX, RY = S::RY;
  extern __shared__ __align__(16) unsigned char leg_smem_raw[];
  S& s = *reinterpret_cast<S*>(leg_smem_raw);
  __shared__ double s_red[kLegThreads / 32];
  if (sc->done) return;
  const LegGeom0 g = leg_geom0<R, NU>(c, TX, TY);
  const int tid = threadIdx.x, nx = c.nx, ny = c.ny;
  const int j_end = min(g.j0 + TY, c.row_hi);
  leg_for_inset<RX, RY, 0>(tid, [&](int rx, int ry) {
    const int i = g.i0 - NU + rx;
    const int qi = g.qi(i), qj = g.qj(g.j0 - NU + ry);
    const size_t idx = (size_t)qi + (size_t)qj * nx;
    R v = zin[idx] + ce[(size_t)((qi - 1) / 2 + 1) + (size_t)((qj - 1) / 2 + 1) * (cmx + 2)];
    if (!g.cavity && i == nx - 1) v = R(0);
    s.a[ry][rx] = v;
    s.rho[ry][rx] = rho[idx];
  });
  __syncthreads();
  // sweeps 1 .. NU-1 on the tile + NU - k
  leg_static_for<1, NU>([&](auto kc) {
    constexpr int K = decltype(kc)::value;
    auto& src = (K % 2 == 1) ? s.a : s.b;
    auto& dst = (K % 2 == 1) ? s.b : s.a;
    leg_for_inset<RX, RY, K>(tid, [&](int rx, int ry) { dst[ry][rx] = leg_sweep0<R, NU>(g, c2, src, s.rho, rx, ry); });
    __syncthreads();
  });
  // the last sweep: the tile's cells and the ring cells that mirror them, straight to memory; rho.z over the unknowns
  auto& src = (NU % 2 == 1) ? s.a : s.b;
  double acc = 0.0;
  leg_for_inset<RX, RY, NU - 1>(tid, [&](int rx, int ry) {
    const int i = g.i0 - NU + rx, j = g.j0 - NU + ry;
    const int qi = g.qi(i), qj = g.qj(j);
    if (i <= nx - 1 && j <= ny - 1 && qi >= g.i0 && qi < g.i0 + TX && qj >= g.j0 && qj < j_end) {
      const R v = leg_sweep0<R, NU>(g, c2, src, s.rho, rx, ry);
      if (i == qi && j == qj) acc += (double)(s.rho[ry][rx] * v);
      zout[(size_t)i + (size_t)j * nx] = v;
    }
  });
  const double t = block_sum<kLegThreads / 32>(acc, s_red);
  if (tid == 0) partials[blockIdx.y * gridDim.x + blockIdx.x] = t;
}

// ---- level 0, column-strip form ------------------------------------------------------------------------------------------
// Same legs, arranged for the instruction issue limit (the flat-indexed kernels above spend more than half of their
// instructions on index arithmetic, clamps and per-cell division guards: profiles/r2_legs_ncu.md).  The staged region is
// exactly 64 columns wide (tile = 64 - 2 NU columns x 32 rows, region rows = 32 + 2 NU); thread (tx, ty) owns column tx
// of one of 4 row strips and walks it upwards with the vertical neighbours rotating through registers: per cell and
// sweep 4 shared loads, 16 fp64 operations, one store, two integer instructions per division (DivTry; the thread redoes
// its strip with DivTrue if a dividend left the window).  No clamps: a tile that touches the boundary of the array
// rewrites the ring cells of each sweep's result from the unknowns they mirror (leg0_fix_ring) before the next sweep
// reads them.
template <class R, class Div>
__device__ __forceinline__ R jacobi_cell_dv(const JacobiConsts2<R>& c, Div& dv, R left, R right, R top, R bot, R cen, R rhs) {
  const R horizontal = dv(right + left, c.dx_sq), vertical = dv(top + bot, c.dy_sq);
  const R p_update = dv(horizontal + vertical - rhs, c.denom);
  return c.omega * p_update + c.one_minus_omega * cen;
}

template <int NU>
struct Leg0 {
  static constexpr int TX = 64 - 2 * NU, TY = 32, RX = 64, RY = TY + 2 * NU, kStrips = kLegThreads / 64;
};
template <class R, int NU>
struct Leg0Smem {
  R rho[Leg0<NU>::RY][64];
  R a[Leg0<NU>::RY][64], b[Leg0<NU>::RY][64];
};

// what a block knows about its tile: region cell (x, y) <-> array cell (i0 - NU + x, j0 - NU + y); [xa, xb] x [ya, yb] =
// the region cells that exist in the array (and, on strips, in this rank's rows + halo)
struct Leg0Box {
  int i0, j0, xa, xb, ya, yb;
  int x_ring_l, x_ring_r, y_ring_b, y_ring_t;  // region index of array column 0 / nx-1, row 0 / ny-1 (or out of [0, 64) / [0, RY))
  bool edge;
};
template <class R, int NU>
__device__ __forceinline__ Leg0Box leg0_box(const MgFine<R>& c) {
  using G = Leg0<NU>;
  Leg0Box b;
  b.i0 = 1 + (int)blockIdx.x * G::TX;
  b.j0 = c.row_lo + (int)blockIdx.y * G::TY;
  const int oi = b.i0 - NU, oj = b.j0 - NU;
  const int j_first = c.row_lo - NU > 0 ? c.row_lo - NU : 0;
  const int j_last = c.row_hi + NU - 1 < c.ny - 1 ? c.row_hi + NU - 1 : c.ny - 1;
  b.xa = max(0, -oi); b.xb = min(63, c.nx - 1 - oi);
  b.ya = max(0, j_first - oj); b.yb = min(G::RY - 1, j_last - oj);
  b.x_ring_l = -oi; b.x_ring_r = c.nx - 1 - oi; b.y_ring_b = -oj; b.y_ring_t = c.ny - 1 - oj;
  b.edge = b.x_ring_l >= 0 || b.x_ring_r < 64 || b.y_ring_b >= 0 || b.y_ring_t < G::RY;
  return b;
}

// ring cells of a sweep's result (cells of the region at least M inside it) <- the unknowns they mirror; column nx-1 of the
// channel is zero.  Call between two __syncthreads.
template <class R, int NU, int M>
__device__ __forceinline__ void leg0_fix_ring(const Leg0Box& b, int cavity, R (*f)[64], int tid) {
  using G = Leg0<NU>;
  constexpr int H = G::RY - 2 * M, W = 64 - 2 * M;
  auto col_src = [&](int x) { return x == b.x_ring_l ? x + 1 : (x == b.x_ring_r ? x - 1 : x); };
  // columns (rows that are not ring rows; the ring rows are written below, corners included)
  for (int k = tid; k < 2 * H; k += kLegThreads) {
    const int y = M + (k >> 1), x = (k & 1) ? b.x_ring_r : b.x_ring_l;
    if (x >= M && x < 64 - M && y != b.y_ring_b && y != b.y_ring_t) f[y][x] = (!(k & 1) || cavity) ? f[y][col_src(x)] : R(0);
  }
  for (int k = tid; k < 2 * W; k += kLegThreads) {
    const int x = M + (k >> 1), y = (k & 1) ? b.y_ring_t : b.y_ring_b;
    if (y >= M && y < G::RY - M) {
      const int ys = (k & 1) ? y - 1 : y + 1;
      f[y][x] = (x == b.x_ring_r && !cavity) ? R(0) : f[ys][col_src(x)];
    }
  }
}

// one Jacobi sweep over the cells at least M inside the region: thread (tx, ty) walks column tx of row strip ty;
// out(x, y, v) receives every value
template <class R, int NU, int M, class Div, class F>
__device__ __forceinline__ void leg0_sweep_strip(const JacobiConsts2<R>& c2, const Leg0Box& b, Div& dv, const R (*src)[64],
                                                 const R (*rho)[64], int tx, int ty, F&& out) {
  using G = Leg0<NU>;
  constexpr int H = G::RY - 2 * M, HS = (H + G::kStrips - 1) / G::kStrips;
  if (tx < max(M, b.xa) || tx > min(63 - M, b.xb)) return;
  const int y0 = M + ty * HS, y_end = min(min(y0 + HS, G::RY - M), b.yb + 1);
  R s_ = src[y0 - 1][tx], c_ = src[y0][tx];
#pragma unroll
  for (int r = 0; r < HS; ++r) {
    const int y = y0 + r;
    if (y < y_end) {
      const R n_ = src[y + 1][tx];
      if (y >= b.ya) out(tx, y, jacobi_cell_dv<R>(c2, dv, src[y][tx - 1], src[y][tx + 1], n_, s_, c_, rho[y][tx]));
      s_ = c_;
      c_ = n_;
    }
  }
}

template <class R, int NU>
__global__ void __launch_bounds__(kLegThreads, 3) k_mg0_down2(const MgFine<R> c, const JacobiConsts2<R> c2,
                                                               const R* __restrict__ rho, R* __restrict__ zout, int cmx,
                                                               R* __restrict__ crho, const MgScalars* __restrict__ sc) {
  using G = Leg0<NU>;
  using S = Leg0Smem<R, NU>;
  extern __shared__ __align__(16) unsigned char leg_smem_raw[];
  S& s = *reinterpret_cast<S*>(leg_smem_raw);
  if (sc->done) return;
  const Leg0Box b = leg0_box<R, NU>(c);
  const int tid = threadIdx.x, tx = tid & 63, ty = tid >> 6, nx = c.nx;
  const int oi = b.i0 - NU, oj = b.j0 - NU;
  const int j_end = min(b.j0 + G::TY, c.row_hi);  // owned unknown rows of the tile: [j0, j_end)
  // rho on the region and x_1 = the first smoothing sweep applied to z = 0 (k_mg_first_sweep's expression), pointwise
  {
    const bool col_ok = tx >= b.xa && tx <= b.xb;
    const R* __restrict__ src = rho + (size_t)(oi + tx);
    R v[(G::RY + G::kStrips - 1) / G::kStrips];
#pragma unroll
    for (int r = 0; r < (G::RY + G::kStrips - 1) / G::kStrips; ++r) {
      const int y = ty + r * G::kStrips;
      v[r] = (col_ok && y >= b.ya && y <= b.yb) ? src[(size_t)(oj + y) * nx] : R(0);
    }
    auto first = [&](auto& dv) {
#pragma unroll
      for (int r = 0; r < (G::RY + G::kStrips - 1) / G::kStrips; ++r) {
        const int y = ty + r * G::kStrips;
        if (y < G::RY) {
          s.rho[y][tx] = v[r];
          s.a[y][tx] = c2.omega * dv((R(0) + R(0)) - v[r], c2.denom) + c2.one_minus_omega * R(0);
        }
      }
    };
    DivTry<R> fast(c2.denom);
    first(fast);
    if (__builtin_expect(!fast.ok(), 0)) {
      DivSlow<R> exact;
      first(exact);
    }
  }
  __syncthreads();
  if (b.edge) {
    leg0_fix_ring<R, NU, 0>(b, c.cavity, s.a, tid);
    __syncthreads();
  }
  // x_k = sweep(x_{k-1}) on the cells at least k - 1 inside the region; the last one (the tile + 1) also goes to memory
  leg_static_for<2, NU + 1>([&](auto kc) {
    constexpr int K = decltype(kc)::value;
    auto& src = (K % 2 == 0) ? s.a : s.b;
    auto& dst = (K % 2 == 0) ? s.b : s.a;
    auto stage = [&](auto& dv) {
      leg0_sweep_strip<R, NU, K - 1>(c2, b, dv, src, s.rho, tx, ty, [&](int x, int y, R v) {
        dst[y][x] = v;
        if (K == NU) {
          const int i = oi + x, j = oj + y;
          if (x >= NU && x < 64 - NU && i <= nx - 2 && j >= b.j0 && j < j_end) zout[(size_t)i + (size_t)j * nx] = v;
        }
      });
    };
    DivTry<R> fast(c2.dx_sq);
    fast.also(c2.dy_sq).also(c2.denom);
    stage(fast);
    if (__builtin_expect(!fast.ok(), 0)) {
      DivSlow<R> exact;
      stage(exact);
    }
    __syncthreads();
    if (b.edge) {
      leg0_fix_ring<R, NU, K - 1>(b, c.cavity, dst, tid);
      __syncthreads();
    }
  });
  auto& xn = (NU % 2 == 0) ? s.b : s.a;
  auto& res = (NU % 2 == 0) ? s.a : s.b;  // free now: the residuals of the tile
  if (b.edge) {
    // the ring cells that mirror this tile's unknowns go to memory too (like every sweep kernel leaves them)
    constexpr int W = 64 - 2 * (NU - 1), H = G::RY - 2 * (NU - 1);
    for (int k = tid; k < W * H; k += kLegThreads) {
      const int y = NU - 1 + k / W, x = NU - 1 + k % W;
      const int i = oi + x, j = oj + y;
      const bool ring = i == 0 || i == nx - 1 || j == 0 || j == c.ny - 1;
      const int qi = min(max(i, 1), nx - 2), qj = min(max(j, 1), c.ny - 2);
      if (ring && i <= nx - 1 && j <= c.ny - 1 && qi >= b.i0 && qi < b.i0 + G::TX && qj >= b.j0 && qj < j_end)
        zout[(size_t)i + (size_t)j * nx] = xn[y][x];
    }
  }
  // residuals rho - L x_nu on the tile (mg_lap's expressions), then rho_1 = their sums over the children of a coarse cell
  // (k_mg_fine_restrict's order: row b outer, column a inner)
  {
    constexpr int HS = G::TY / G::kStrips;
    auto stage = [&](auto& dv) {
      if (tx >= NU && tx < 64 - NU && tx <= b.xb) {
        const int y0 = NU + ty * HS;
        R s_ = xn[y0 - 1][tx], c_ = xn[y0][tx];
#pragma unroll
        for (int r = 0; r < HS; ++r) {
          const int y = y0 + r;
          const R n_ = xn[y + 1][tx];
          res[y][tx] = s.rho[y][tx] - mg_lap<R>(c, dv, c_, xn[y][tx + 1], xn[y][tx - 1], n_, s_);
          s_ = c_;
          c_ = n_;
        }
      }
    };
    DivTry<R> fast(c.ddx_sq);
    fast.also(c.ddy_sq);
    stage(fast);
    if (__builtin_expect(!fast.ok(), 0)) {
      DivSlow<R> exact;
      stage(exact);
    }
  }
  __syncthreads();
  constexpr int CX = G::TX / 2, CY = G::TY / 2;
  for (int k = tid; k < CX * CY; k += kLegThreads) {
    const int cy = k / CX, cx = k - cy * CX;
    const int I = (int)blockIdx.x * CX + cx, J = (b.j0 - 1) / 2 + cy;
    if (1 + 2 * I <= nx - 2 && 1 + 2 * J < j_end) {
      R acc = R(0);
#pragma unroll
      for (int bb = 0; bb < 2; ++bb)
#pragma unroll
        for (int aa = 0; aa < 2; ++aa) {
          const int i = 1 + 2 * I + aa, j = 1 + 2 * J + bb;
          if (i <= nx - 2 && j <= c.ny - 2) acc += res[j - oj][i - oi];
        }
      crho[(size_t)(I + 1) + (size_t)(J + 1) * (cmx + 2)] = acc;
    }
  }
}

// ascending leg of level 0, column-strip form: z = NU sweeps of (x_nu + correction of the parents); rho.z -> one partial per block
template <class R, int NU>
__global__ void __launch_bounds__(kLegThreads, 3) k_mg0_up2(const MgFine<R> c, const JacobiConsts2<R> c2,
                                                             const R* __restrict__ zin, const R* __restrict__ rho, int cmx,
                                                             const R* __restrict__ ce, R* __restrict__ zout,
                                                             double* __restrict__ partials, const MgScalars* __restrict__ sc) {
  using G = Leg0<NU>;
  using S = Leg0Smem<R, NU>;
  extern __shared__ __align__(16) unsigned char leg_smem_raw[];
  S& s = *reinterpret_cast<S*>(leg_smem_raw);
  __shared__ double s_red[kLegThreads / 32];
  if (sc->done) return;
  const Leg0Box b = leg0_box<R, NU>(c);
  const int tid = threadIdx.x, tx = tid & 63, ty = tid >> 6, nx = c.nx;
  const int oi = b.i0 - NU, oj = b.j0 - NU;
  const int j_end = min(b.j0 + G::TY, c.row_hi);
  {
    const bool col_ok = tx >= b.xa && tx <= b.xb;
    const int i = oi + tx;
    const int qi = min(max(i, 1), nx - 2);
    const size_t pcol = (size_t)((qi - 1) / 2 + 1);
#pragma unroll
    for (int r = 0; r < (G::RY + G::kStrips - 1) / G::kStrips; ++r) {
      const int y = ty + r * G::kStrips;
      if (y < G::RY) {
        R v = R(0), q = R(0);
        if (col_ok && y >= b.ya && y <= b.yb) {
          const int j = oj + y;
          const int qj = min(max(j, 1), c.ny - 2);
          const size_t idx = (size_t)i + (size_t)j * nx;
          v = zin[idx] + ce[pcol + (size_t)((qj - 1) / 2 + 1) * (cmx + 2)];
          q = rho[idx];
        }
        s.a[y][tx] = v;
        s.rho[y][tx] = q;
      }
    }
  }
  __syncthreads();
  if (b.edge) {
    leg0_fix_ring<R, NU, 0>(b, c.cavity, s.a, tid);
    __syncthreads();
  }
  double acc = 0.0;
  leg_static_for<1, NU + 1>([&](auto kc) {
    constexpr int K = decltype(kc)::value;
    auto& src = (K % 2 == 1) ? s.a : s.b;
    auto& dst = (K % 2 == 1) ? s.b : s.a;
    auto stage = [&](auto& dv) {
      if (K == NU) acc = 0.0;
      auto out = [&](int x, int y, R v) {
        if (K == NU) {  // the tile's unknowns: straight to memory, rho.z; (edge tiles: everything also to dst for the ring cells)
          const int i = oi + x, j = oj + y;
          if (b.edge) dst[y][x] = v;
          if (x >= NU && x < 64 - NU && i <= nx - 2 && j >= b.j0 && j < j_end) {
            acc += (double)(s.rho[y][x] * v);
            zout[(size_t)i + (size_t)j * nx] = v;
          }
        } else {
          dst[y][x] = v;
        }
      };
      // the last sweep: the tile; on an edge tile the tile + 1, whose ring cells go to memory as well
      if (K == NU && !b.edge) leg0_sweep_strip<R, NU, NU>(c2, b, dv, src, s.rho, tx, ty, out);
      else leg0_sweep_strip<R, NU, (K == NU ? NU - 1 : K)>(c2, b, dv, src, s.rho, tx, ty, out);
    };
    DivTry<R> fast(c2.dx_sq);
    fast.also(c2.dy_sq).also(c2.denom);
    stage(fast);
    if (__builtin_expect(!fast.ok(), 0)) {
      DivSlow<R> exact;
      stage(exact);
    }
    if (K != NU || b.edge) __syncthreads();
    if (b.edge) {
      leg0_fix_ring<R, NU, (K == NU ? NU - 1 : K)>(b, c.cavity, dst, tid);
      __syncthreads();
    }
  });
  if (b.edge) {
    auto& zn = (NU % 2 == 1) ? s.b : s.a;
    constexpr int W = 64 - 2 * (NU - 1), H = G::RY - 2 * (NU - 1);
    for (int k = tid; k < W * H; k += kLegThreads) {
      const int y = NU - 1 + k / W, x = NU - 1 + k % W;
      const int i = oi + x, j = oj + y;
      const bool ring = i == 0 || i == nx - 1 || j == 0 || j == c.ny - 1;
      const int qi = min(max(i, 1), nx - 2), qj = min(max(j, 1), c.ny - 2);
      if (ring && i <= nx - 1 && j <= c.ny - 1 && qi >= b.i0 && qi < b.i0 + G::TX && qj >= b.j0 && qj < j_end)
        zout[(size_t)i + (size_t)j * nx] = zn[y][x];
    }
  }
  const double t = block_sum<kLegThreads / 32>(acc, s_red);
  if (tid == 0) partials[blockIdx.y * gridDim.x + blockIdx.x] = t;
}

// one cell of a damped-Jacobi sweep on a coarse level: mgc_sweep_cell's arithmetic (k_mgc_sweep_tab's form)
template <class R, class Div>
__device__ __forceinline__ R mgc_cell(R we, R ww, R cyw, R wn, R ws, R cxh, R cc, R ee, R ew, R en, R es, R rho,
                                      const DivG<R>& dg, R omega, Div& dv) {
  const R le = cxh * (we * (ee - cc) + ww * (ew - cc)) + cyw * (wn * (en - cc) + ws * (es - cc));
  const R res = dv(le - rho, dg);
  return dg.y > R(0) ? cc + omega * res : R(0);
}

// ---- register-tiled form (level 0 and coarse levels) ------------------------------------------------------------------
// The shipped legs.  The staged region is 64 columns x 36 rows (tile = (64 - 2 NU) x (36 - 2 NU)); thread (tx, ty) owns
// column tx of the FIXED row strip [9 ty, 9 ty + 9) through every sweep: its 9 cells and their right-hand sides live in
// registers, shared memory only carries each sweep's result to the neighbouring threads (per cell and sweep: 2 shared
// loads for the horizontal neighbours, 16 fp64 operations, 1 shared store, 2 integer instructions per division).  Every
// thread computes all of its cells in every sweep, without predicates: the cells a sweep does not need (closer than k to
// the region's edge) are computed from clamped neighbours and never read by a cell that matters.  Tiles that touch the
// boundary of the array rewrite the ring cells of each sweep's result from the unknowns they mirror (leg0_fix_ring) and
// reload their registers; cells outside the array start from rho = 1 (any harmless non-zero value).
template <int NU>
struct Leg3 {
  static constexpr int TX = 64 - 2 * NU, TY = 36 - 2 * NU, RY = 36, HS = 9;
};
template <class R>
struct Leg3Smem {
  R a[36][64], b[36][64];
};

template <class R, int NU>
__device__ __forceinline__ Leg0Box leg3_box(const MgFine<R>& c) {
  using G = Leg3<NU>;
  Leg0Box b;
  b.i0 = 1 + (int)blockIdx.x * G::TX;
  b.j0 = c.row_lo + (int)blockIdx.y * G::TY;
  const int oi = b.i0 - NU, oj = b.j0 - NU;
  const int j_first = c.row_lo - NU > 0 ? c.row_lo - NU : 0;
  const int j_last = c.row_hi + NU - 1 < c.ny - 1 ? c.row_hi + NU - 1 : c.ny - 1;
  b.xa = max(0, -oi); b.xb = min(63, c.nx - 1 - oi);
  b.ya = max(0, j_first - oj); b.yb = min(G::RY - 1, j_last - oj);
  b.x_ring_l = -oi; b.x_ring_r = c.nx - 1 - oi; b.y_ring_b = -oj; b.y_ring_t = c.ny - 1 - oj;
  b.edge = b.x_ring_l >= 0 || b.x_ring_r < 64 || b.y_ring_b >= 0 || b.y_ring_t < G::RY;
  return b;
}

// ring cells of a 64 x 36 staged field <- the unknowns they mirror (every ring cell inside the region; column nx-1 of the
// channel: zero).  Call between two __syncthreads.
template <class R>
__device__ __forceinline__ void leg3_fix_ring(const Leg0Box& b, int cavity, R (*f)[64], int tid) {
  auto col_src = [&](int x) { return x == b.x_ring_l ? x + 1 : (x == b.x_ring_r ? x - 1 : x); };
  for (int k = tid; k < 2 * 36; k += kLegThreads) {
    const int y = k >> 1, x = (k & 1) ? b.x_ring_r : b.x_ring_l;
    if (x >= 0 && x < 64 && y != b.y_ring_b && y != b.y_ring_t) f[y][x] = (!(k & 1) || cavity) ? f[y][col_src(x)] : R(0);
  }
  for (int k = tid; k < 2 * 64; k += kLegThreads) {
    const int x = k >> 1, y = (k & 1) ? b.y_ring_t : b.y_ring_b;
    if (y >= 0 && y < 36) {
      const int ys = (k & 1) ? y - 1 : y + 1;
      f[y][x] = (x == b.x_ring_r && !cavity) ? R(0) : f[ys][col_src(x)];
    }
  }
}

// what a thread of the register-tiled legs knows about its strip
struct Leg3Thread {
  int tx, y0, txl, txr, y_below, y_above;
  __device__ __forceinline__ explicit Leg3Thread(int tid) {
    tx = tid & 63; y0 = (tid >> 6) * 9;
    txl = max(tx - 1, 0); txr = min(tx + 1, 63);
    y_below = max(y0 - 1, 0); y_above = min(y0 + 9, 35);
  }
};

// one Jacobi sweep of the thread's 9 cells, in place in `cur` (the vertical neighbours inside the strip are the thread's own
// registers); the result goes to dst for the neighbours.  Redone with true divisions if a dividend left the window.
template <class R>
__device__ __forceinline__ void leg3_sweep0(const JacobiConsts2<R>& c2, const Leg3Thread& t, const R (*src)[64], R (*dst)[64],
                                            R (&cur)[9], const R (&q)[9]) {
  auto pass = [&](auto& dv) {
    R s_ = src[t.y_below][t.tx];
    const R above = src[t.y_above][t.tx];
#pragma unroll
    for (int r = 0; r < 9; ++r) {
      const R c_ = cur[r], n_ = r < 8 ? cur[r + 1] : above;
      cur[r] = jacobi_cell_dv<R>(c2, dv, src[t.y0 + r][t.txl], src[t.y0 + r][t.txr], n_, s_, c_, q[r]);
      s_ = c_;
    }
  };
  DivTry<R> fast(c2.dx_sq);
  fast.also(c2.dy_sq).also(c2.denom);
  pass(fast);
  if (__builtin_expect(!fast.ok(), 0)) {
#pragma unroll
    for (int r = 0; r < 9; ++r) cur[r] = src[t.y0 + r][t.tx];
    DivSlow<R> exact;
    pass(exact);
  }
#pragma unroll
  for (int r = 0; r < 9; ++r) dst[t.y0 + r][t.tx] = cur[r];
}

template <class R>
__device__ __forceinline__ void leg3_reload(const Leg3Thread& t, const R (*f)[64], R (&cur)[9]) {
#pragma unroll
  for (int r = 0; r < 9; ++r) cur[r] = f[t.y0 + r][t.tx];
}

// the ring cells that mirror this tile's unknowns go to memory too (like every sweep kernel leaves them)
template <class R, int NU>
__device__ __forceinline__ void leg3_store_ring(const Leg0Box& b, const MgFine<R>& c, const R (*f)[64], R* __restrict__ zout,
                                                int j_end, int tid) {
  using G = Leg3<NU>;
  constexpr int W = G::TX + 2, H = G::TY + 2;
  const int oi = b.i0 - NU, oj = b.j0 - NU;
  for (int k = tid; k < W * H; k += kLegThreads) {
    const int y = NU - 1 + k / W, x = NU - 1 + k % W;
    const int i = oi + x, j = oj + y;
    const bool ring = i == 0 || i == c.nx - 1 || j == 0 || j == c.ny - 1;
    const int qi = min(max(i, 1), c.nx - 2), qj = min(max(j, 1), c.ny - 2);
    if (ring && i <= c.nx - 1 && j <= c.ny - 1 && qi >= b.i0 && qi < b.i0 + G::TX && qj >= b.j0 && qj < j_end)
      zout[(size_t)i + (size_t)j * c.nx] = f[y][x];
  }
}

template <class R, int NU>
__global__ void __launch_bounds__(kLegThreads, 3) k_mg0_down3(const MgFine<R> c, const JacobiConsts2<R> c2,
                                                               const R* __restrict__ rho, R* __restrict__ zout, int cmx,
                                                               R* __restrict__ crho, const MgScalars* __restrict__ sc) {
  using G = Leg3<NU>;
  extern __shared__ __align__(16) unsigned char leg_smem_raw[];
  Leg3Smem<R>& s = *reinterpret_cast<Leg3Smem<R>*>(leg_smem_raw);
  if (sc->done) return;
  const Leg0Box b = leg3_box<R, NU>(c);
  const int tid = threadIdx.x, nx = c.nx;
  const Leg3Thread t(tid);
  const int oi = b.i0 - NU, oj = b.j0 - NU;
  const int j_end = min(b.j0 + G::TY, c.row_hi);  // owned unknown rows of the tile: [j0, j_end)
  R q[9], cur[9];
  // rho of the thread's cells, x_1 = the first smoothing sweep applied to z = 0 (k_mg_first_sweep's expression)
  {
    const bool col_ok = t.tx >= b.xa && t.tx <= b.xb;
    const R* __restrict__ src = rho + min(max(oi + t.tx, 0), nx - 1);
#pragma unroll
    for (int r = 0; r < 9; ++r) q[r] = src[(long)(oj + min(max(t.y0 + r, b.ya), b.yb)) * nx];  // clamped, branch-free
#pragma unroll
    for (int r = 0; r < 9; ++r) {
      const int y = t.y0 + r;
      if (!(col_ok && y >= b.ya && y <= b.yb)) q[r] = R(1);
    }
    auto first = [&](auto& dv) {
#pragma unroll
      for (int r = 0; r < 9; ++r) cur[r] = c2.omega * dv((R(0) + R(0)) - q[r], c2.denom) + c2.one_minus_omega * R(0);
    };
    DivTry<R> fast(c2.denom);
    first(fast);
    if (__builtin_expect(!fast.ok(), 0)) {
      DivSlow<R> exact;
      first(exact);
    }
#pragma unroll
    for (int r = 0; r < 9; ++r) s.a[t.y0 + r][t.tx] = cur[r];
  }
  __syncthreads();
  if (b.edge) {
    leg3_fix_ring<R>(b, c.cavity, s.a, tid);
    __syncthreads();
    leg3_reload<R>(t, s.a, cur);
  }
  leg_static_for<2, NU + 1>([&](auto kc) {
    constexpr int K = decltype(kc)::value;
    auto& src = (K % 2 == 0) ? s.a : s.b;
    auto& dst = (K % 2 == 0) ? s.b : s.a;
    leg3_sweep0<R>(c2, t, src, dst, cur, q);
    __syncthreads();
    if (b.edge) {
      leg3_fix_ring<R>(b, c.cavity, dst, tid);
      __syncthreads();
      leg3_reload<R>(t, dst, cur);
    }
  });
  auto& xn = (NU % 2 == 0) ? s.b : s.a;
  auto& res = (NU % 2 == 0) ? s.a : s.b;  // free now: the residuals of the tile
  // x_nu of the tile's unknowns
  if (t.tx >= NU && t.tx < 64 - NU && oi + t.tx <= nx - 2) {
    R* __restrict__ out = zout + (long)(oi + t.tx);
#pragma unroll
    for (int r = 0; r < 9; ++r) {
      const int y = t.y0 + r, j = oj + y;
      if (y >= NU && y < NU + G::TY && j < j_end) out[(long)j * nx] = cur[r];
    }
  }
  if (b.edge) leg3_store_ring<R, NU>(b, c, xn, zout, j_end, tid);
  // residuals rho - L x_nu (mg_lap's expressions), then rho_1 = their sums over the children of a coarse cell
  // (k_mg_fine_restrict's order: row b outer, column a inner)
  {
    R rs[9];
    auto pass = [&](auto& dv) {
      R s_ = xn[t.y_below][t.tx];
      const R above = xn[t.y_above][t.tx];
#pragma unroll
      for (int r = 0; r < 9; ++r) {
        const R n_ = r < 8 ? cur[r + 1] : above;
        rs[r] = q[r] - mg_lap<R>(c, dv, cur[r], xn[t.y0 + r][t.txr], xn[t.y0 + r][t.txl], n_, s_);
        s_ = cur[r];
      }
    };
    DivTry<R> fast(c.ddx_sq);
    fast.also(c.ddy_sq);
    pass(fast);
    if (__builtin_expect(!fast.ok(), 0)) {
      DivSlow<R> exact;
      pass(exact);
    }
#pragma unroll
    for (int r = 0; r < 9; ++r) res[t.y0 + r][t.tx] = rs[r];
  }
  __syncthreads();
  constexpr int CX = G::TX / 2, CY = G::TY / 2;
  for (int k = tid; k < CX * CY; k += kLegThreads) {
    const int cy = k / CX, cx = k - cy * CX;
    const int I = (int)blockIdx.x * CX + cx, J = (b.j0 - 1) / 2 + cy;
    if (1 + 2 * I <= nx - 2 && 1 + 2 * J < j_end) {
      R acc = R(0);
#pragma unroll
      for (int bb = 0; bb < 2; ++bb)
#pragma unroll
        for (int aa = 0; aa < 2; ++aa) {
          const int i = 1 + 2 * I + aa, j = 1 + 2 * J + bb;
          if (i <= nx - 2 && j <= c.ny - 2) acc += res[j - oj][i - oi];
        }
      crho[(size_t)(I + 1) + (size_t)(J + 1) * (cmx + 2)] = acc;
    }
  }
}

// ascending leg of level 0, register-tiled: z = NU sweeps of (x_nu + correction of the parents); rho.z -> one partial per block
template <class R, int NU>
__global__ void __launch_bounds__(kLegThreads, 3) k_mg0_up3(const MgFine<R> c, const JacobiConsts2<R> c2,
                                                             const R* __restrict__ zin, const R* __restrict__ rho, int cmx,
                                                             const R* __restrict__ ce, R* __restrict__ zout,
                                                             double* __restrict__ partials, const MgScalars* __restrict__ sc) {
  using G = Leg3<NU>;
  extern __shared__ __align__(16) unsigned char leg_smem_raw[];
  Leg3Smem<R>& s = *reinterpret_cast<Leg3Smem<R>*>(leg_smem_raw);
  __shared__ double s_red[kLegThreads / 32];
  if (sc->done) return;
  const Leg0Box b = leg3_box<R, NU>(c);
  const int tid = threadIdx.x, nx = c.nx;
  const Leg3Thread t(tid);
  const int oi = b.i0 - NU, oj = b.j0 - NU;
  const int j_end = min(b.j0 + G::TY, c.row_hi);
  R q[9], cur[9];
  {
    // every load of the strip first, from clamped (always valid) addresses and without branches: a load whose consumer sits
    // in the same conditional block costs one DRAM round trip per row (profiles/r2_legs_ncu.md)
    const bool col_ok = t.tx >= b.xa && t.tx <= b.xb;
    const int i = min(max(oi + t.tx, 0), nx - 1);
    const int qi = min(max(i, 1), nx - 2);
    const R* __restrict__ pc = ce + (long)((qi - 1) / 2 + 1);
    const R* __restrict__ pz = zin + i;
    const R* __restrict__ pr = rho + i;
    R zv[9], cv[9];
#pragma unroll
    for (int r = 0; r < 9; ++r) {
      const int j = oj + min(max(t.y0 + r, b.ya), b.yb);
      const int qj = min(max(j, 1), c.ny - 2);
      zv[r] = pz[(long)j * nx];
      cv[r] = pc[(long)((qj - 1) / 2 + 1) * (cmx + 2)];
      q[r] = pr[(long)j * nx];
    }
#pragma unroll
    for (int r = 0; r < 9; ++r) {
      const int y = t.y0 + r;
      const bool ok = col_ok && y >= b.ya && y <= b.yb;
      cur[r] = ok ? zv[r] + cv[r] : R(1);
      if (!ok) q[r] = R(1);
      s.a[y][t.tx] = cur[r];
    }
  }
  __syncthreads();
  if (b.edge) {
    leg3_fix_ring<R>(b, c.cavity, s.a, tid);
    __syncthreads();
    leg3_reload<R>(t, s.a, cur);
  }
  leg_static_for<1, NU + 1>([&](auto kc) {
    constexpr int K = decltype(kc)::value;
    auto& src = (K % 2 == 1) ? s.a : s.b;
    auto& dst = (K % 2 == 1) ? s.b : s.a;
    leg3_sweep0<R>(c2, t, src, dst, cur, q);
    if (K != NU || b.edge) __syncthreads();
    if (b.edge) {
      leg3_fix_ring<R>(b, c.cavity, dst, tid);
      __syncthreads();
      if (K != NU) leg3_reload<R>(t, dst, cur);
    }
  });
  // the tile's unknowns to memory, rho.z
  double acc = 0.0;
  if (t.tx >= NU && t.tx < 64 - NU && oi + t.tx <= nx - 2) {
    R* __restrict__ out = zout + (long)(oi + t.tx);
#pragma unroll
    for (int r = 0; r < 9; ++r) {
      const int y = t.y0 + r, j = oj + y;
      if (y >= NU && y < NU + G::TY && j < j_end) {
        acc += (double)(q[r] * cur[r]);
        out[(long)j * nx] = cur[r];
      }
    }
  }
  if (b.edge) leg3_store_ring<R, NU>(b, c, (NU % 2 == 1) ? s.b : s.a, zout, j_end, tid);
  const double tot = block_sum<kLegThreads / 32>(acc, s_red);
  if (tid == 0) partials[blockIdx.y * gridDim.x + blockIdx.x] = tot;
}

// ---- coarse levels, register-tiled ----
// Cells outside the level are the ring of zeros (and beyond): they get the null class of the diagonal table (divisor 0 ->
// mgc_cell returns 0), zero weights and rho = 1, so they stay zero through every sweep without a predicate.
template <class R>
struct Leg3SmemC {
  Leg3Smem<R> f;
  R wn[36], ws[36], cxh[36];
  R dy[kMgClasses * kMgClasses + 1], dr[kMgClasses * kMgClasses + 1];  // + the null class
  int rcls8[36];  // row class * kMgClasses; rows outside the level: >= the null class
  R c0[3];        // column 0's weights and class: what leg3_tile_uniform compares every column with
  int c0_cls;
};
constexpr int kLegNullClass = kMgClasses * kMgClasses;

template <class R, int NU>
__device__ __forceinline__ void leg3_load_level(Leg3SmemC<R>& s, const MgLevelDev<R>& L, int J0, int tid) {
  for (int k = tid; k < 36; k += kLegThreads) {
    const int J = J0 - NU + k;
    const bool in = J >= 0 && J < L.my;
    s.wn[k] = in ? L.WN[J] : R(0);
    s.ws[k] = in ? L.WS[J] : R(0);
    s.cxh[k] = in ? L.CXH[J] : R(0);
    s.rcls8[k] = in ? (int)L.row_class[J] * kMgClasses : kLegNullClass;
  }
  for (int k = tid; k <= kLegNullClass; k += kLegThreads) {
    s.dy[k] = k < kLegNullClass ? L.diag_table[k].y : R(0);
    s.dr[k] = k < kLegNullClass ? L.diag_table[k].r : R(0);
  }
}

// Row data of a coarse level as the sweeps see it: the general form looks every row up in shared memory (weights, class ->
// diagonal table); on a tile whose 36 rows and 64 columns all carry the same weights (every tile away from the level's
// boundary and from a trailing unpaired cell) the row weights and the one diagonal are block constants in registers — the
// same numbers through the same expressions, 2 instead of 8 shared loads per cell and sweep.
template <class R>
struct Leg3RowsSmem {
  const Leg3SmemC<R>& s;
  int col_cls;
  __device__ __forceinline__ R wn(int y) const { return s.wn[y]; }
  __device__ __forceinline__ R ws(int y) const { return s.ws[y]; }
  __device__ __forceinline__ R cxh(int y) const { return s.cxh[y]; }
  __device__ __forceinline__ DivG<R> dg(int y) const {
    DivG<R> d;
    const int cls = min(s.rcls8[y] + col_cls, kLegNullClass);
    d.y = s.dy[cls]; d.r = s.dr[cls]; d.lo = 0u; d.span = 0u;
    return d;
  }
};
template <class R>
struct Leg3RowsUniform {
  R wn_, ws_, cxh_;
  DivG<R> dg_;
  __device__ __forceinline__ R wn(int) const { return wn_; }
  __device__ __forceinline__ R ws(int) const { return ws_; }
  __device__ __forceinline__ R cxh(int) const { return cxh_; }
  __device__ __forceinline__ DivG<R> dg(int) const { return dg_; }
};

// one sweep of the thread's 9 cells of a coarse level, in place in `cur` (mgc_cell = mgc_sweep_cell's arithmetic)
template <class R, class Rows>
__device__ __forceinline__ void leg3_sweep_c(const Rows& rw, const DivG<R>& win, const Leg3Thread& t, R we, R ww, R cyw, R omega,
                                             const R (*src)[64], R (*dst)[64], R (&cur)[9], const R (&q)[9]) {
  auto pass = [&](auto& dv) {
    R s_ = src[t.y_below][t.tx];
    const R above = src[t.y_above][t.tx];
#pragma unroll
    for (int r = 0; r < 9; ++r) {
      const int y = t.y0 + r;
      const R c_ = cur[r], n_ = r < 8 ? cur[r + 1] : above;
      cur[r] = mgc_cell<R>(we, ww, cyw, rw.wn(y), rw.ws(y), rw.cxh(y), c_, src[y][t.txr], src[y][t.txl], n_, s_, q[r], rw.dg(y), omega, dv);
      s_ = c_;
    }
  };
  DivTry<R> fast(win);
  pass(fast);
  if (__builtin_expect(!fast.ok(), 0)) {
#pragma unroll
    for (int r = 0; r < 9; ++r) cur[r] = src[t.y0 + r][t.tx];
    DivSlow<R> exact;
    pass(exact);
  }
#pragma unroll
  for (int r = 0; r < 9; ++r) dst[t.y0 + r][t.tx] = cur[r];
}

// what a thread of the coarse legs knows about its column, and whether the whole tile is uniform
template <class R>
struct Leg3ColW {
  R we, ww, cyw;
  int cls;  // column class; outside the level: the null class
  bool in;
};
template <class R, int NU>
__device__ __forceinline__ Leg3ColW<R> leg3_column(const MgLevelDev<R>& L, int I) {
  Leg3ColW<R> c;
  c.in = I >= 0 && I < L.mx;
  c.cls = c.in ? (int)L.col_class[I] : kLegNullClass;
  c.we = c.in ? L.WE[I] : R(0);
  c.ww = c.in ? L.WW[I] : R(0);
  c.cyw = c.in ? L.CYW[I] : R(0);
  return c;
}
// call after leg3_load_level + __syncthreads (contains a barrier): do all 64 columns and all 36 rows carry the same numbers?
template <class R>
__device__ __forceinline__ bool leg3_tile_uniform(Leg3SmemC<R>& s, const Leg3ColW<R>& col, int tid) {
  if (tid == 0) { s.c0[0] = col.we; s.c0[1] = col.ww; s.c0[2] = col.cyw; s.c0_cls = col.cls; }
  __syncthreads();
  bool same = col.we == s.c0[0] && col.ww == s.c0[1] && col.cyw == s.c0[2] && col.cls == s.c0_cls && col.cls < kLegNullClass;
  if (tid < 36) same = same && s.wn[tid] == s.wn[0] && s.ws[tid] == s.ws[0] && s.cxh[tid] == s.cxh[0] && s.rcls8[tid] == s.rcls8[0] &&
                       s.rcls8[tid] < kLegNullClass;
  return __syncthreads_and(same) != 0;
}

template <class R, int NU, class Rows>
__device__ __forceinline__ void leg3_down_body(Leg3SmemC<R>& s, const Rows& rw, const MgLevelDev<R>& L, const Leg3ColW<R>& col,
                                               const Leg3Thread& t, const R* __restrict__ rho, R* __restrict__ xout, int cmx,
                                               R* __restrict__ crho, R omega, int I0, int J0, int J_end, int c_lo, int c_hi, int tid) {
  using G = Leg3<NU>;
  const int mx = L.mx, my = L.my;
  const long W = (long)mx + 2;
  const int I = I0 - NU + t.tx;
  const R we = col.we, ww = col.ww, cyw = col.cyw;
  R q[9], cur[9];
  {
    const R* __restrict__ src = rho + (long)(min(max(I, 0), mx - 1) + 1);
#pragma unroll
    for (int r = 0; r < 9; ++r) q[r] = src[(long)(min(max(J0 - NU + t.y0 + r, 0), my - 1) + 1) * W];  // clamped, branch-free
#pragma unroll
    for (int r = 0; r < 9; ++r) {
      const int J = J0 - NU + t.y0 + r;
      if (!(col.in && J >= 0 && J < my)) q[r] = R(1);
    }
  }
  const DivG<R> win = L.diag_table[kMgClasses * kMgClasses];  // intersection of the table's dividend windows
  // x_1 = sweep of the zero field: cc + omega * ((L 0 - rho) / diag) with cc = 0, L 0 = +0
  {
    auto first = [&](auto& dv) {
#pragma unroll
      for (int r = 0; r < 9; ++r) {
        const DivG<R> dg = rw.dg(t.y0 + r);
        const R res = dv(R(0) - q[r], dg);
        cur[r] = dg.y > R(0) ? R(0) + omega * res : R(0);
      }
    };
    DivTry<R> fast(win);
    first(fast);
    if (__builtin_expect(!fast.ok(), 0)) {
      DivSlow<R> exact;
      first(exact);
    }
#pragma unroll
    for (int r = 0; r < 9; ++r) s.f.a[t.y0 + r][t.tx] = cur[r];
  }
  __syncthreads();
  leg_static_for<2, NU + 1>([&](auto kc) {
    constexpr int K = decltype(kc)::value;
    auto& src = (K % 2 == 0) ? s.f.a : s.f.b;
    auto& dst = (K % 2 == 0) ? s.f.b : s.f.a;
    leg3_sweep_c<R>(rw, win, t, we, ww, cyw, omega, src, dst, cur, q);
    __syncthreads();
  });
  auto& xn = (NU % 2 == 0) ? s.f.b : s.f.a;
  auto& res = (NU % 2 == 0) ? s.f.a : s.f.b;
  if (t.tx >= NU && t.tx < 64 - NU && col.in) {
    R* __restrict__ out = xout + (long)(I + 1);
#pragma unroll
    for (int r = 0; r < 9; ++r) {
      const int y = t.y0 + r, J = J0 - NU + y;
      if (y >= NU && y < NU + G::TY && J < J_end && J < my) out[(long)(J + 1) * W] = cur[r];
    }
  }
  // residuals rho_l - L_l x_nu (mg_coarse_apply's expression), then rho_{l+1} = their sums over the children
  // (mgc_restrict_cell's order)
  {
    R s_ = xn[t.y_below][t.tx];
    const R above = xn[t.y_above][t.tx];
#pragma unroll
    for (int r = 0; r < 9; ++r) {
      const int y = t.y0 + r;
      const R cc = cur[r], n_ = r < 8 ? cur[r + 1] : above;
      const R le = rw.cxh(y) * (we * (xn[y][t.txr] - cc) + ww * (xn[y][t.txl] - cc)) + cyw * (rw.wn(y) * (n_ - cc) + rw.ws(y) * (s_ - cc));
      res[y][t.tx] = q[r] - le;
      s_ = cc;
    }
  }
  __syncthreads();
  constexpr int CX = G::TX / 2, CY = G::TY / 2;
  for (int k = tid; k < CX * CY; k += kLegThreads) {
    const int cy = k / CX, cx = k - cy * CX;
    const int Ic = I0 / 2 + cx, Jc = J0 / 2 + cy;
    if (2 * Ic < mx && 2 * Jc < my && 2 * Jc < J_end && Jc >= c_lo && Jc < c_hi) {
      R acc = R(0);
#pragma unroll
      for (int bb = 0; bb < 2; ++bb)
#pragma unroll
        for (int aa = 0; aa < 2; ++aa) {
          const int i = 2 * Ic + aa, j = 2 * Jc + bb;
          if (i < mx && j < my) acc += res[j - (J0 - NU)][i - (I0 - NU)];
        }
      crho[(size_t)(Ic + 1) + (size_t)(Jc + 1) * ((size_t)cmx + 2)] = acc;
    }
  }
}

// descending leg of level l >= 1, register-tiled (arguments as k_mgc_down)
template <class R, int NU>
__global__ void __launch_bounds__(kLegThreads, 2) k_mgc_down3(const MgLevelDev<R> L, const R* __restrict__ rho,
                                                               R* __restrict__ xout, int cmx, R* __restrict__ crho, R omega,
                                                               int row_lo, int row_hi, int c_lo, int c_hi,
                                                               const MgScalars* __restrict__ sc) {
  using G = Leg3<NU>;
  extern __shared__ __align__(16) unsigned char leg_smem_raw[];
  Leg3SmemC<R>& s = *reinterpret_cast<Leg3SmemC<R>*>(leg_smem_raw);
  if (sc->done) return;
  const int tid = threadIdx.x;
  const Leg3Thread t(tid);
  const int I0 = (int)blockIdx.x * G::TX, J0 = row_lo + (int)blockIdx.y * G::TY;
  const int J_end = min(J0 + G::TY, row_hi);
  const Leg3ColW<R> col = leg3_column<R, NU>(L, I0 - NU + t.tx);
  leg3_load_level<R, NU>(s, L, J0, tid);
  __syncthreads();
  if (leg3_tile_uniform<R>(s, col, tid)) {
    Leg3RowsUniform<R> rw;
    rw.wn_ = s.wn[0]; rw.ws_ = s.ws[0]; rw.cxh_ = s.cxh[0];
    const int cls = s.rcls8[0] + col.cls;
    rw.dg_.y = s.dy[cls]; rw.dg_.r = s.dr[cls]; rw.dg_.lo = 0u; rw.dg_.span = 0u;
    leg3_down_body<R, NU>(s, rw, L, col, t, rho, xout, cmx, crho, omega, I0, J0, J_end, c_lo, c_hi, tid);
  } else {
    const Leg3RowsSmem<R> rw{s, col.cls};
    leg3_down_body<R, NU>(s, rw, L, col, t, rho, xout, cmx, crho, omega, I0, J0, J_end, c_lo, c_hi, tid);
  }
}

template <class R, int NU, class Rows>
__device__ __forceinline__ void leg3_up_body(Leg3SmemC<R>& s, const Rows& rw, const MgLevelDev<R>& L, const Leg3ColW<R>& col,
                                             const Leg3Thread& t, const R* __restrict__ xin, const R* __restrict__ rho, int cmx,
                                             const R* __restrict__ ce, R* __restrict__ xout, R omega, int I0, int J0, int J_end) {
  using G = Leg3<NU>;
  const int mx = L.mx, my = L.my;
  const long W = (long)mx + 2;
  const int I = I0 - NU + t.tx;
  R q[9], cur[9];
  {
    // every load first, from clamped addresses and without branches (see k_mg0_up3)
    const int Ic = min(max(I, 0), mx - 1);
    const R* __restrict__ pc = ce + (long)(Ic / 2 + 1);
    const R* __restrict__ px = xin + (long)(Ic + 1);
    const R* __restrict__ pr = rho + (long)(Ic + 1);
    R xv[9], cv[9];
#pragma unroll
    for (int r = 0; r < 9; ++r) {
      const int Jc = min(max(J0 - NU + t.y0 + r, 0), my - 1);
      xv[r] = px[(long)(Jc + 1) * W];
      cv[r] = pc[(long)(Jc / 2 + 1) * ((long)cmx + 2)];
      q[r] = pr[(long)(Jc + 1) * W];
    }
#pragma unroll
    for (int r = 0; r < 9; ++r) {
      const int y = t.y0 + r, J = J0 - NU + y;
      const bool ok = col.in && J >= 0 && J < my;
      cur[r] = ok ? xv[r] + cv[r] : R(0);
      if (!ok) q[r] = R(1);
      s.f.a[y][t.tx] = cur[r];
    }
  }
  __syncthreads();
  const DivG<R> win = L.diag_table[kMgClasses * kMgClasses];
  leg_static_for<1, NU + 1>([&](auto kc) {
    constexpr int K = decltype(kc)::value;
    auto& src = (K % 2 == 1) ? s.f.a : s.f.b;
    auto& dst = (K % 2 == 1) ? s.f.b : s.f.a;
    leg3_sweep_c<R>(rw, win, t, col.we, col.ww, col.cyw, omega, src, dst, cur, q);
    if (K != NU) __syncthreads();
  });
  if (t.tx >= NU && t.tx < 64 - NU && col.in) {
    R* __restrict__ out = xout + (long)(I + 1);
#pragma unroll
    for (int r = 0; r < 9; ++r) {
      const int y = t.y0 + r, J = J0 - NU + y;
      if (y >= NU && y < NU + G::TY && J < J_end && J < my) out[(long)(J + 1) * W] = cur[r];
    }
  }
}

// ascending leg of level l >= 1, register-tiled (arguments as k_mgc_up)
template <class R, int NU>
__global__ void __launch_bounds__(kLegThreads, 2) k_mgc_up3(const MgLevelDev<R> L, const R* __restrict__ xin,
                                                             const R* __restrict__ rho, int cmx, const R* __restrict__ ce,
                                                             R* __restrict__ xout, R omega, int row_lo, int row_hi,
                                                             const MgScalars* __restrict__ sc) {
  using G = Leg3<NU>;
  extern __shared__ __align__(16) unsigned char leg_smem_raw[];
  Leg3SmemC<R>& s = *reinterpret_cast<Leg3SmemC<R>*>(leg_smem_raw);
  if (sc->done) return;
  const int tid = threadIdx.x;
  const Leg3Thread t(tid);
  const int I0 = (int)blockIdx.x * G::TX, J0 = row_lo + (int)blockIdx.y * G::TY;
  const int J_end = min(J0 + G::TY, row_hi);
  const Leg3ColW<R> col = leg3_column<R, NU>(L, I0 - NU + t.tx);
  leg3_load_level<R, NU>(s, L, J0, tid);
  __syncthreads();
  if (leg3_tile_uniform<R>(s, col, tid)) {
    Leg3RowsUniform<R> rw;
    rw.wn_ = s.wn[0]; rw.ws_ = s.ws[0]; rw.cxh_ = s.cxh[0];
    const int cls = s.rcls8[0] + col.cls;
    rw.dg_.y = s.dy[cls]; rw.dg_.r = s.dr[cls]; rw.dg_.lo = 0u; rw.dg_.span = 0u;
    leg3_up_body<R, NU>(s, rw, L, col, t, xin, rho, cmx, ce, xout, omega, I0, J0, J_end);
  } else {
    const Leg3RowsSmem<R> rw{s, col.cls};
    leg3_up_body<R, NU>(s, rw, L, col, t, xin, rho, cmx, ce, xout, omega, I0, J0, J_end);
  }
}

// ---- coarse levels (l >= 1, fields (mx + 2) x (my + 2) with a ring of zeros; levels with a diagonal table) -------------
template <class R, int TX, int TY, int NU>
struct LegSmemC {
  static constexpr int RX = TX + 2 * NU, RY = TY + 2 * NU;
  LegSmem<R, TX, TY, NU> f;
  R we[RX], ww[RX], cyw[RX], wn[RY], ws[RY], cxh[RY];
  R dy[kMgClasses * kMgClasses], dr[kMgClasses * kMgClasses];  // the diagonal table: divisors and hoisted reciprocals
  unsigned char ccls[RX], rcls[RY];
};

template <class R, int TX, int TY, int NU>
__device__ __forceinline__ void leg_load_level(LegSmemC<R, TX, TY, NU>& s, const MgLevelDev<R>& L, int I0, int J0, int tid) {
  constexpr int RX = TX + 2 * NU, RY = TY + 2 * NU;
  for (int k = tid; k < RX; k += kLegThreads) {
    const int I = I0 - NU + k;
    const bool in = I >= 0 && I < L.mx;
    s.we[k] = in ? L.WE[I] : R(0);
    s.ww[k] = in ? L.WW[I] : R(0);
    s.cyw[k] = in ? L.CYW[I] : R(0);
    s.ccls[k] = in ? L.col_class[I] : (unsigned char)0;
  }
  for (int k = tid; k < RY; k += kLegThreads) {
    const int J = J0 - NU + k;
    const bool in = J >= 0 && J < L.my;
    s.wn[k] = in ? L.WN[J] : R(0);
    s.ws[k] = in ? L.WS[J] : R(0);
    s.cxh[k] = in ? L.CXH[J] : R(0);
    s.rcls[k] = in ? L.row_class[J] : (unsigned char)0;
  }
  for (int k = tid; k < kMgClasses * kMgClasses; k += kLegThreads) {
    s.dy[k] = L.diag_table[k].y;
    s.dr[k] = L.diag_table[k].r;
  }
}

template <class R, int TX, int TY, int NU>
__device__ __forceinline__ DivG<R> leg_diag(const LegSmemC<R, TX, TY, NU>& s, int rx, int ry) {
  DivG<R> dg;
  const int cls = s.rcls[ry] * kMgClasses + s.ccls[rx];
  dg.y = s.dy[cls]; dg.r = s.dr[cls]; dg.lo = 0u; dg.span = 0u;
  return dg;
}

// one sweep of a coarse level at the region cell (rx, ry) from the staged field `src`
template <class R, int TX, int TY, int NU, class Src, class Div>
__device__ __forceinline__ R leg_coarse_sweep(const LegSmemC<R, TX, TY, NU>& s, const Src& src, int rx, int ry, R omega, Div& dv) {
  return mgc_cell<R>(s.we[rx], s.ww[rx], s.cyw[rx], s.wn[ry], s.ws[ry], s.cxh[ry], src[ry][rx], src[ry][rx + 1], src[ry][rx - 1],
                     src[ry + 1][rx], src[ry - 1][rx], s.f.rho[ry][rx], leg_diag<R, TX, TY, NU>(s, rx, ry), omega, dv);
}

// descending leg of level l: x_nu on the rows [row_lo, row_hi) of the level's unknowns (whole level: [0, my); strips: the
// owned rows plus NU on each side, recomputed from NU + 1 halo rows of rho so that the ascending leg needs no exchange
// of x_nu), the parents' rho on the coarse rows [c_lo, c_hi) (strips: the owned ones)
template <class R, int TX, int TY, int NU>
__global__ void __launch_bounds__(kLegThreads, 2) k_mgc_down(const MgLevelDev<R> L, const R* __restrict__ rho,
                                                              R* __restrict__ xout, int cmx, R* __restrict__ crho, R omega,
                                                              int row_lo, int row_hi, int c_lo, int c_hi,
                                                              const MgScalars* __restrict__ sc) {
  using S = LegSmemC<R, TX, TY, NU>;
  constexpr int RX = S::RX, RY = S::RY;
  extern __shared__ __align__(16) unsigned char leg_smem_raw[];
  S& s = *reinterpret_cast<S*>(leg_smem_raw);
  if (sc->done) return;
  const int tid = threadIdx.x, mx = L.mx, my = L.my;
  const int I0 = (int)blockIdx.x * TX, J0 = row_lo + (int)blockIdx.y * TY;
  const int J_end = min(J0 + TY, row_hi);
  const size_t W = (size_t)mx + 2;
  auto inside = [&](int rx, int ry) {
    const int I = I0 - NU + rx, J = J0 - NU + ry;
    return I >= 0 && I < mx && J >= 0 && J < my;
  };
  leg_load_level<R, TX, TY, NU>(s, L, I0, J0, tid);
  leg_for_inset<RX, RY, 0>(tid, [&](int rx, int ry) {
    s.f.rho[ry][rx] = inside(rx, ry) ? rho[(size_t)(I0 - NU + rx + 1) + (size_t)(J0 - NU + ry + 1) * W] : R(0);
  });
  __syncthreads();
  const DivG<R> win = L.diag_table[kMgClasses * kMgClasses];  // intersection of the table's dividend windows
  // x_1 = sweep of the zero field: cc + omega * ((L 0 - rho) / diag) with cc = 0, L 0 = +0
  auto stage1 = [&](auto& dv) {
    leg_for_inset<RX, RY, 0>(tid, [&](int rx, int ry) {
      R v = R(0);
      if (inside(rx, ry)) {  // (cells outside the level are the ring of zeros: no division, their zero dividend would fail the window)
        const DivG<R> dg = leg_diag<R, TX, TY, NU>(s, rx, ry);
        const R res = dv(R(0) - s.f.rho[ry][rx], dg);
        v = dg.y > R(0) ? R(0) + omega * res : R(0);
      }
      s.f.a[ry][rx] = v;
    });
  };
  {
    DivTry<R> fast(win);
    stage1(fast);
    if (__builtin_expect(!fast.ok(), 0)) {
      DivSlow<R> exact;
      stage1(exact);
    }
  }
  __syncthreads();
  leg_static_for<2, NU + 1>([&](auto kc) {
    constexpr int K = decltype(kc)::value;
    auto& src = (K % 2 == 0) ? s.f.a : s.f.b;
    auto& dst = (K % 2 == 0) ? s.f.b : s.f.a;
    auto stage = [&](auto& dv) {
      leg_for_inset<RX, RY, K - 1>(tid, [&](int rx, int ry) {
        R v = R(0);  // outside the level: the ring of zeros
        const bool in = inside(rx, ry);
        if (in) v = leg_coarse_sweep<R, TX, TY, NU>(s, src, rx, ry, omega, dv);
        dst[ry][rx] = v;
        if (K == NU) {
          const int I = I0 - NU + rx, J = J0 - NU + ry;
          if (in && I >= I0 && I < I0 + TX && J >= J0 && J < J_end) xout[(size_t)(I + 1) + (size_t)(J + 1) * W] = v;
        }
      });
    };
    DivTry<R> fast(win);
    stage(fast);
    if (__builtin_expect(!fast.ok(), 0)) {
      DivSlow<R> exact;
      stage(exact);
    }
    __syncthreads();
  });
  auto& xn = (NU % 2 == 0) ? s.f.b : s.f.a;
  // rho_{l+1}[I, J] = sum over the children of (rho_l - L_l x_nu)  (mgc_restrict_cell's order and expressions)
  constexpr int CX = TX / 2, CY = TY / 2;
#pragma unroll
  for (int k0 = 0; k0 < CX * CY; k0 += kLegThreads) {
    const int k = k0 + tid;
    if (k < CX * CY) {
      const int cy = k / CX, cx = k - cy * CX;
      const int I = I0 / 2 + cx, J = J0 / 2 + cy;
      if (2 * I < mx && 2 * J < my && 2 * J < J_end && J >= c_lo && J < c_hi) {
        R acc = R(0);
#pragma unroll
        for (int b = 0; b < 2; ++b)
#pragma unroll
          for (int a = 0; a < 2; ++a) {
            const int i = 2 * I + a, j = 2 * J + b;
            if (i < mx && j < my) {
              const int rx = i - (I0 - NU), ry = j - (J0 - NU);
              const R cc = xn[ry][rx];
              const R le = s.cxh[ry] * (s.we[rx] * (xn[ry][rx + 1] - cc) + s.ww[rx] * (xn[ry][rx - 1] - cc)) +
                           s.cyw[rx] * (s.wn[ry] * (xn[ry + 1][rx] - cc) + s.ws[ry] * (xn[ry - 1][rx] - cc));
              acc += s.f.rho[ry][rx] - le;
            }
          }
        crho[(size_t)(I + 1) + (size_t)(J + 1) * ((size_t)cmx + 2)] = acc;
      }
    }
  }
}

// ascending leg of level l: xout = NU sweeps of (x_nu + correction of the parents) on the rows [row_lo, row_hi)
template <class R, int TX, int TY, int NU>
__global__ void __launch_bounds__(kLegThreads, 2) k_mgc_up(const MgLevelDev<R> L, const R* __restrict__ xin,
                                                            const R* __restrict__ rho, int cmx, const R* __restrict__ ce,
                                                            R* __restrict__ xout, R omega, int row_lo, int row_hi,
                                                            const MgScalars* __restrict__ sc) {
  using S = LegSmemC<R, TX, TY, NU>;
  constexpr int RX = S::RX, RY = S::RY;
  extern __shared__ __align__(16) unsigned char leg_smem_raw[];
  S& s = *reinterpret_cast<S*>(leg_smem_raw);
  if (sc->done) return;
  const int tid = threadIdx.x, mx = L.mx, my = L.my;
  const int I0 = (int)blockIdx.x * TX, J0 = row_lo + (int)blockIdx.y * TY;
  const int J_end = min(J0 + TY, row_hi);
  const size_t W = (size_t)mx + 2;
  auto inside = [&](int rx, int ry) {
    const int I = I0 - NU + rx, J = J0 - NU + ry;
    return I >= 0 && I < mx && J >= 0 && J < my;
  };
  leg_load_level<R, TX, TY, NU>(s, L, I0, J0, tid);
  leg_for_inset<RX, RY, 0>(tid, [&](int rx, int ry) {
    R v = R(0), q = R(0);
    if (inside(rx, ry)) {
      const int I = I0 - NU + rx, J = J0 - NU + ry;
      const size_t idx = (size_t)(I + 1) + (size_t)(J + 1) * W;
      v = xin[idx] + ce[(size_t)(I / 2 + 1) + (size_t)(J / 2 + 1) * ((size_t)cmx + 2)];
      q = rho[idx];
    }
    s.f.a[ry][rx] = v;
    s.f.rho[ry][rx] = q;
  });
  __syncthreads();
  const DivG<R> win = L.diag_table[kMgClasses * kMgClasses];
  leg_static_for<1, NU + 1>([&](auto kc) {
    constexpr int K = decltype(kc)::value;
    auto& src = (K % 2 == 1) ? s.f.a : s.f.b;
    auto& dst = (K % 2 == 1) ? s.f.b : s.f.a;
    auto stage = [&](auto& dv) {
      leg_for_inset<RX, RY, K>(tid, [&](int rx, int ry) {
        R v = R(0);
        const bool in = inside(rx, ry);
        if (in) v = leg_coarse_sweep<R, TX, TY, NU>(s, src, rx, ry, omega, dv);
        if (K == NU) {  // the tile itself: straight to memory
          const int I = I0 - NU + rx, J = J0 - NU + ry;
          if (in && J < J_end) xout[(size_t)(I + 1) + (size_t)(J + 1) * W] = v;
        } else {
          dst[ry][rx] = v;
        }
      });
    };
    DivTry<R> fast(win);
    stage(fast);
    if (__builtin_expect(!fast.ok(), 0)) {
      DivSlow<R> exact;
      stage(exact);
    }
    if (K != NU) __syncthreads();
  });
}

}  // namespace cfdk
