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
// order.  Divisions: DivTry per thread, recomputed with DivTrue by the thread whose window test failed.
#pragma once

#include <type_traits>

namespace cfdk {

constexpr int kLegThreads = 256;

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
      DivTrue<R> exact;
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
      DivTrue<R> exact;
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

// ---- coarse levels (l >= 1, fields (mx + 2) x (my + 2) with a ring of zeros; levels with a diagonal table) -------------
template <class R, int TX, int TY, int NU>
struct LegSmemC {
  static constexpr int RX = TX + 2 * NU, RY = TY + 2 * NU;
  LegSmem<R, TX, TY, NU> f;
  R we[RX], ww[RX], cyw[RX], wn[RY], ws[RY], cxh[RY];
  R dy[kMgClasses * kMgClasses], dr[kMgClasses * kMgClasses];  // the diagonal table: divisors and hoisted reciprocals
  unsigned char ccls[RX], rcls[RY];
};

// one cell of a damped-Jacobi sweep on a coarse level: mgc_sweep_cell's arithmetic (k_mgc_sweep_tab's form)
template <class R, class Div>
__device__ __forceinline__ R mgc_cell(R we, R ww, R cyw, R wn, R ws, R cxh, R cc, R ee, R ew, R en, R es, R rho,
                                      const DivG<R>& dg, R omega, Div& dv) {
  const R le = cxh * (we * (ee - cc) + ww * (ew - cc)) + cyw * (wn * (en - cc) + ws * (es - cc));
  const R res = dv(le - rho, dg);
  return dg.y > R(0) ? cc + omega * res : R(0);
}

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
      DivTrue<R> exact;
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
      DivTrue<R> exact;
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
      DivTrue<R> exact;
      stage(exact);
    }
    if (K != NU) __syncthreads();
  });
}

}  // namespace cfdk
