// cfd_mg.cuh — EXTENSION, "Mode C" fast path: conjugate gradients preconditioned by one geometric-multigrid
// V-cycle (CFD_SOLVER_MGCG).  No reference counterpart (the reference has damped Jacobi only, src/model.rs:
// 734-824; its doc comment names multigrid as intended, :526); the CPU test oracle carries the same algorithm.
//
// Problem, in the reference's sign convention: L x = rhs, (L x)[i,j] = ((xE - x) + (xW - x))/dx^2 +
// ((xN - x) + (xS - x))/dy^2 on the unknowns (columns 1..nx-2, rows 1..ny-2) under the Jacobi boundary rules
// (:807-815: mirror left / bottom / top, zero outlet column; cavity extension: mirror there too).
//  * Level 0 = the grid itself.  Its smoother IS the reference's damped-Jacobi sweep (k_jacobi_sweep5, the
//    tensor-TMA kernel the roofline is quoted on), run with damping mg_omega on (z, rho) instead of (p', rhs).
//  * Level l+1 pairs the cells of level l per direction (a trailing single cell stays single when the count is
//    odd — 4096 - 2 = 4094 = 2 * 2047), down to 1 x 1.  Coarse operators are finite-volume discretisations on
//    that non-uniform tensor grid: a link weighs (shared face) / (centre distance) in finest-cell units, so a
//    level is described by six 1-D arrays (per column: WE, WW, CYW = width/dy^2; per row: WN, WS, CXH =
//    height/dx^2).  Coarse fields carry a ring of zeros, so no kernel branches on the boundary.
//  * Transfer: residuals are summed over the (up to four) children, corrections are copied to them.
// Every per-cell expression is written exactly like the oracle's (no FMA, same association), so everything but
// the dot products (summed in a different order) is bit-identical; Mode C parity is to a tolerance.
#pragma once

namespace cfdk {

constexpr int kMgClasses = 8;  // distinct (WE + WW, CYW) column / (WN + WS, CXH) row combinations a level's diagonal table holds
template <class R>
struct MgLevelDev {
  int mx, my;               // unknowns per direction; fields are (mx + 2) x (my + 2)
  const R *WE, *WW, *CYW;   // per column
  const R *WN, *WS, *CXH;   // per row
  // The diagonal CXH[J] (WE[I] + WW[I]) + CYW[I] (WN[J] + WS[J]) takes only a handful of distinct values on a level (the
  // grid is uniform except next to the boundary and where a trailing cell stayed single): they are tabulated per
  // (row class, column class) together with their hoisted reciprocals (DivG), so that the sweep's division by the
  // diagonal is the exact hoisted-reciprocal one instead of the compiler's 30-instruction sequence.  nullptr: no table
  // (more than kMgClasses classes), the generic per-cell kernel is used.
  const unsigned char *col_class, *row_class;
  const DivG<R>* diag_table;  // [row class * kMgClasses + column class]
};

constexpr int kMgThreads = 256;
constexpr int kMgRows = 8;  // rows per block of the level-0 vector kernels

// ---- level 0 vector kernels ----------------------------------------------------------------------------
// A thread owns the aligned column pair (2t, 2t+1) (16-byte loads / stores; the ring columns 0 and nx-1 are
// masked out) and walks kMgRows rows; grid = (ceil(nx / 512), ceil((ny - 2) / kMgRows)).  Dot products: every
// thread accumulates its cells in a fixed order, a block reduces to ONE partial, and the block that finishes
// last (ticket) sums the partials in index order and advances the CG scalars — deterministic, no extra launch.
template <class R>
struct MgPair {
  int c0, j0, j1;
  bool v0, v1, any;
};
template <class R>
__device__ __forceinline__ MgPair<R> mg_pair(const MgFine<R>& c) {
  MgPair<R> p;
  p.c0 = 2 * (blockIdx.x * blockDim.x + threadIdx.x);
  p.j0 = c.row_lo + blockIdx.y * kMgRows;
  p.j1 = min(p.j0 + kMgRows, c.row_hi);
  p.any = p.c0 < c.nx;
  p.v0 = p.any && p.c0 >= 1;
  p.v1 = p.any && p.c0 + 1 <= c.nx - 2;
  return p;
}

// (L x) on the unknowns of one column pair; l / r = the columns left / right of the pair, already replaced by the
// boundary rules where the pair touches the ring (mirror; 0 at the channel outlet)
// Div = DivTry (hoisted reciprocals, one window flag per tile) or DivTrue (`/`): cfd_kernels.cuh
template <class R, class Div>
__device__ __forceinline__ R mg_lap(const MgFine<R>& c, Div& dv, R cc, R xe, R xw, R xn, R xs) {
  return dv((xe - cc) + (xw - cc), c.ddx_sq) + dv((xn - cc) + (xs - cc), c.ddy_sq);
}

template <class R, class Div>
__device__ __forceinline__ R mg_fine_apply(const MgFine<R>& c, Div& dv, const R* __restrict__ x, int i, int j) {
  const size_t idx = (size_t)i + (size_t)j * c.nx;
  const R cc = x[idx];
  const R xe = (i == c.nx - 2) ? (c.cavity ? cc : R(0)) : x[idx + 1];
  const R xw = (i == 1) ? cc : x[idx - 1];
  const R xn = (j == c.ny - 2) ? cc : x[idx + c.nx];
  const R xs = (j == 1) ? cc : x[idx - c.nx];
  return mg_lap<R>(c, dv, cc, xe, xw, xn, xs);
}

// Start vector of a solve, derived on the fly from the p' the last first-solves ended with (the buffers rotate on the
// host, nothing is copied): mode 3 = 3 a - 3 b + c (quadratic extrapolation in time), 2 = 2 a - b (linear, the JS
// twin's "extrapolated initial guess", index.html:262-270), 1 = a (what the reference's Jacobi does by never resetting
// p', src/model.rs:734-824), 0 = cold start.  Same expressions, same association as the oracle's mg_guess update.
template <class R>
struct MgStart {
  const R *a, *b, *c;
  int mode;
};
template <class R>
__device__ __forceinline__ R mg_start_one(const MgStart<R>& g, size_t idx) {
  if (g.mode == 3) return R(3) * g.a[idx] - R(3) * g.b[idx] + g.c[idx];
  if (g.mode == 2) return R(2) * g.a[idx] - g.b[idx];
  return g.a[idx];
}
template <class R>
__device__ __forceinline__ typename Vec2<R>::type mg_start_pair(const MgStart<R>& g, size_t idx) {
  using V = typename Vec2<R>::type;
  const V a = *reinterpret_cast<const V*>(g.a + idx);
  V o = a;
  if (g.mode == 3) {
    const V b = *reinterpret_cast<const V*>(g.b + idx), c = *reinterpret_cast<const V*>(g.c + idx);
    o.x = R(3) * a.x - R(3) * b.x + c.x;
    o.y = R(3) * a.y - R(3) * b.y + c.y;
  } else if (g.mode == 2) {
    const V b = *reinterpret_cast<const V*>(g.b + idx);
    o.x = R(2) * a.x - b.x;
    o.y = R(2) * a.y - b.y;
  }
  return o;
}

// x = start vector (or 0) on the whole grid, rho = rhs - L x on the unknowns (0 on the ring), rho.rho; the search
// direction needs no initialisation (k_mg_dir_apply takes it as zero in a solve's first iteration).
// grid = (ceil(nx / 512), ceil(ny / rows_per_block)); a thread walks its column pair up the tile, the start vector's rows
// j-1, j, j+1 rotate through registers.
template <class R>
__global__ void __launch_bounds__(kMgThreads) k_mg_init(MgFine<R> c, MgScalars* __restrict__ sc,
                                                         const R* __restrict__ rhs, const MgStart<R> g,
                                                         R* __restrict__ x, R* __restrict__ rho,
                                                         double* __restrict__ partials, unsigned* __restrict__ ticket,
                                                         int rows_per_block) {
  using V = typename Vec2<R>::type;
  const int nx = c.nx;
  const int c0 = 2 * (blockIdx.x * blockDim.x + threadIdx.x);
  const int j0 = c.init_lo + blockIdx.y * rows_per_block, j1 = min(j0 + rows_per_block, c.init_hi);
  double acc = 0.0;
  if (c0 < nx && j0 < j1) {
    V zero;
    zero.x = R(0); zero.y = R(0);
    const bool warm = g.mode != 0;
    const int cl = max(c0 - 1, 0), cr = min(c0 + 2, nx - 1);
    // rows j0-1 and j0 of the start vector (row -1 does not exist: the stencil never reads below row 1's mirror)
    V south = zero, cen = zero, north = zero;
    if (warm) {
      if (j0 >= 1) south = mg_start_pair<R>(g, (size_t)c0 + (size_t)(j0 - 1) * nx);
      cen = mg_start_pair<R>(g, (size_t)c0 + (size_t)j0 * nx);
    }
    for (int j = j0; j < j1; ++j) {
      const size_t idx = (size_t)c0 + (size_t)j * nx;
      const bool row_ok = j >= 1 && j <= c.ny - 2;
      const bool ok0 = row_ok && c0 >= 1, ok1 = row_ok && c0 + 1 <= nx - 2;
      V b = *reinterpret_cast<const V*>(rhs + idx);
      if (warm) {
        if (j + 1 <= c.ny - 1) north = mg_start_pair<R>(g, idx + nx);
        if (row_ok) {
          const size_t row = (size_t)j * nx;
          const R gl = mg_start_one<R>(g, row + cl), gr = mg_start_one<R>(g, row + cr);
          const R xw0 = (c0 == 1) ? cen.x : gl;
          const R xe0 = (c0 == nx - 2) ? (c.cavity ? cen.x : R(0)) : cen.y;
          const R xn0 = (j == c.ny - 2) ? cen.x : north.x, xs0 = (j == 1) ? cen.x : south.x;
          const R xw1 = (c0 + 1 == 1) ? cen.y : cen.x;
          const R xe1 = (c0 + 1 == nx - 2) ? (c.cavity ? cen.y : R(0)) : gr;
          const R xn1 = (j == c.ny - 2) ? cen.y : north.y, xs1 = (j == 1) ? cen.y : south.y;
          DivTry<R> dv(c.ddx_sq);
          dv.also(c.ddy_sq);
          R l0 = mg_lap<R>(c, dv, cen.x, xe0, xw0, xn0, xs0), l1 = mg_lap<R>(c, dv, cen.y, xe1, xw1, xn1, xs1);
          if (__builtin_expect(!dv.ok(), 0)) {
            DivTrue<R> ex;
            l0 = mg_lap<R>(c, ex, cen.x, xe0, xw0, xn0, xs0);
            l1 = mg_lap<R>(c, ex, cen.y, xe1, xw1, xn1, xs1);
          }
          if (ok0) b.x = b.x - l0;
          if (ok1) b.y = b.y - l1;
        }
      }
      if (!ok0) b.x = R(0);
      if (!ok1) b.y = R(0);
      *reinterpret_cast<V*>(x + idx) = cen;
      *reinterpret_cast<V*>(rho + idx) = b;
      acc += (double)(b.x * b.x);
      acc += (double)(b.y * b.y);
      south = cen;
      cen = north;
    }
  }
  mg_finish_dot<R, kMgThreads>(c, sc, partials, ticket, acc, 0);
}

// the derived start vector, written out (state read-back: CFD_FIELD_MG_GUESS); whole owned rows, grid-stride
template <class R>
__global__ void __launch_bounds__(kMgThreads) k_mg_start_materialize(const MgStart<R> g, R* __restrict__ out, size_t n) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += stride) out[k] = mg_start_one<R>(g, k);
}

// The kernels below load every row of their tile before they compute (fully unrolled, rows past the tile's end
// predicated off): with one row in flight per thread they were latency-bound at ~3 TB/s.

// a.b over the unknowns -> mg_advance(mode)
template <class R>
__global__ void __launch_bounds__(kMgThreads) k_mg_dot(MgFine<R> c, MgScalars* __restrict__ sc, const R* __restrict__ a,
                                                        const R* __restrict__ b, double* __restrict__ partials,
                                                        unsigned* __restrict__ ticket, int mode) {
  using V = typename Vec2<R>::type;
  const MgPair<R> p = mg_pair<R>(c);
  double acc = 0.0;
  if (p.any) {
    V av[kMgRows], bv[kMgRows];
#pragma unroll
    for (int r = 0; r < kMgRows; ++r) {
      const int j = min(p.j0 + r, c.row_hi - 1);
      const size_t idx = (size_t)p.c0 + (size_t)j * c.nx;
      av[r] = *reinterpret_cast<const V*>(a + idx);
      bv[r] = *reinterpret_cast<const V*>(b + idx);
    }
#pragma unroll
    for (int r = 0; r < kMgRows; ++r) {
      if (p.j0 + r < p.j1) {
        if (p.v0) acc += (double)(av[r].x * bv[r].x);
        if (p.v1) acc += (double)(av[r].y * bv[r].y);
      }
    }
  }
  mg_finish_dot<R, kMgThreads>(c, sc, partials, ticket, acc, mode);
}

// d_new = z + beta d_old and w = L d_new in one pass (d_new goes to its own buffer: neighbours still read d_old),
// d_new.w -> alpha.  Tiles of kMgDirRows rows; grid = (ceil(nx / 512), ceil((ny - 2) / kMgDirRows)).
constexpr int kMgDirRows = 2;  // r2af, 4096^2, whole step: 2 rows x 256 threads 2.595 ms, 2 x 128 2.609, 4 x 256 2.613 (58 against 90 registers)
template <class R, int kRows = kMgDirRows, int kThreads = kMgThreads>
__global__ void __launch_bounds__(kThreads) k_mg_dir_apply(MgFine<R> c, MgScalars* __restrict__ sc,
                                                              const R* __restrict__ z, const R* __restrict__ d_old,
                                                              R* __restrict__ d_new, R* __restrict__ w,
                                                              double* __restrict__ partials, unsigned* __restrict__ ticket) {
  using V = typename Vec2<R>::type;
  if (sc->done) return;  // iterations are enqueued in batches; once the solve is over the rest are no-ops
  const R beta = (R)sc->beta;
  // first iteration of a solve: d_old is identically zero by definition (and beta is 0) — it is neither initialised
  // by k_mg_init nor read here; z + 0 * 0 keeps the oracle's arithmetic (mg_d = mg_z + 0 * mg_d over zeros)
  const bool first = sc->iterations == 0;
  const int nx = c.nx;
  const int c0 = 2 * (blockIdx.x * blockDim.x + threadIdx.x);
  const int j0 = c.row_lo + blockIdx.y * kRows, j1 = min(j0 + kRows, c.row_hi);
  const bool any = c0 < nx, v0 = any && c0 >= 1, v1 = any && c0 + 1 <= nx - 2;
  double acc = 0.0;
  if (any) {
    V dn[kRows + 2];            // d_new of the own pair on rows j0-1 .. j0+kRows
    R dl[kRows], dr[kRows];  // d_new left / right of the pair on the tile's rows
    const int cl = max(c0 - 1, 0), cr = min(c0 + 2, nx - 1);
#pragma unroll
    for (int m = 0; m < kRows + 2; ++m) {
      const int j = min(j0 - 1 + m, c.row_hi);  // row_hi is the ring / halo row above the owned unknowns
      const size_t idx = (size_t)c0 + (size_t)j * nx;
      const V zv = *reinterpret_cast<const V*>(z + idx);
      V dv;
      dv.x = R(0); dv.y = R(0);
      if (!first) dv = *reinterpret_cast<const V*>(d_old + idx);
      dn[m].x = zv.x + beta * dv.x;
      dn[m].y = zv.y + beta * dv.y;
    }
#pragma unroll
    for (int r = 0; r < kRows; ++r) {
      const int j = min(j0 + r, c.row_hi - 1);
      const size_t row = (size_t)j * nx;
      dl[r] = z[row + cl] + beta * (first ? R(0) : d_old[row + cl]);
      dr[r] = z[row + cr] + beta * (first ? R(0) : d_old[row + cr]);
    }
#pragma unroll
    V wv[kRows];
    auto apply_tile = [&](auto& dv) {
#pragma unroll
      for (int r = 0; r < kRows; ++r) {
        const int j = min(j0 + r, j1 - 1);
        const V cen = dn[r + 1], south = dn[r], north = dn[r + 2];
        const R xw0 = (c0 == 1) ? cen.x : dl[r];
        const R xe0 = (c0 == nx - 2) ? (c.cavity ? cen.x : R(0)) : cen.y;
        const R xn0 = (j == c.ny - 2) ? cen.x : north.x, xs0 = (j == 1) ? cen.x : south.x;
        const R xw1 = (c0 + 1 == 1) ? cen.y : cen.x;
        const R xe1 = (c0 + 1 == nx - 2) ? (c.cavity ? cen.y : R(0)) : dr[r];
        const R xn1 = (j == c.ny - 2) ? cen.y : north.y, xs1 = (j == 1) ? cen.y : south.y;
        wv[r].x = mg_lap<R>(c, dv, cen.x, xe0, xw0, xn0, xs0);
        wv[r].y = mg_lap<R>(c, dv, cen.y, xe1, xw1, xn1, xs1);
      }
    };
    DivTry<R> fast(c.ddx_sq);
    fast.also(c.ddy_sq);
    apply_tile(fast);
    if (__builtin_expect(!fast.ok(), 0)) {
      DivTrue<R> exact;
      apply_tile(exact);
    }
#pragma unroll
    for (int r = 0; r < kRows; ++r) {
      const int j = j0 + r;
      if (j < j1) {
        const V cen = dn[r + 1];
        V out;
        out.x = v0 ? wv[r].x : R(0);
        out.y = v1 ? wv[r].y : R(0);
        if (v0) acc += (double)(cen.x * out.x);
        if (v1) acc += (double)(cen.y * out.y);
        *reinterpret_cast<V*>(d_new + c0 + (size_t)j * nx) = cen;
        *reinterpret_cast<V*>(w + c0 + (size_t)j * nx) = out;
      }
    }
  }
  mg_finish_dot<R, kThreads>(c, sc, partials, ticket, acc, 2);
}

// x += alpha d, rho -= alpha w, rho.rho -> iteration count, stopping rule.  Tiles of kMgUpdRows rows.
constexpr int kMgUpdRows = 4;  // r2af: 2-row tiles change nothing here (2.616 against 2.613 ms/step)
template <class R, int kRows = kMgUpdRows, int kThreads = kMgThreads>
__global__ void __launch_bounds__(kThreads) k_mg_update(MgFine<R> c, MgScalars* __restrict__ sc,
                                                           const R* __restrict__ d, const R* __restrict__ w,
                                                           R* __restrict__ x, R* __restrict__ rho,
                                                           double* __restrict__ partials, unsigned* __restrict__ ticket) {
  using V = typename Vec2<R>::type;
  if (sc->done) return;
  const R alpha = (R)sc->alpha;
  const int c0 = 2 * (blockIdx.x * blockDim.x + threadIdx.x);
  const int j0 = c.row_lo + blockIdx.y * kRows, j1 = min(j0 + kRows, c.row_hi);
  const bool any = c0 < c.nx, v0 = any && c0 >= 1, v1 = any && c0 + 1 <= c.nx - 2;
  double acc = 0.0;
  if (any) {
    V dv[kRows], wv[kRows], xv[kRows], rv[kRows];
#pragma unroll
    for (int r = 0; r < kRows; ++r) {
      const int j = min(j0 + r, c.row_hi - 1);
      const size_t idx = (size_t)c0 + (size_t)j * c.nx;
      dv[r] = *reinterpret_cast<const V*>(d + idx);
      wv[r] = *reinterpret_cast<const V*>(w + idx);
      xv[r] = *reinterpret_cast<const V*>(x + idx);
      rv[r] = *reinterpret_cast<const V*>(rho + idx);
    }
#pragma unroll
    for (int r = 0; r < kRows; ++r) {
      const int j = j0 + r;
      if (j < j1) {
        if (v0) {
          xv[r].x = xv[r].x + alpha * dv[r].x;
          rv[r].x = rv[r].x - alpha * wv[r].x;
          acc += (double)(rv[r].x * rv[r].x);
        }
        if (v1) {
          xv[r].y = xv[r].y + alpha * dv[r].y;
          rv[r].y = rv[r].y - alpha * wv[r].y;
          acc += (double)(rv[r].y * rv[r].y);
        }
        const size_t idx = (size_t)c0 + (size_t)j * c.nx;
        *reinterpret_cast<V*>(x + idx) = xv[r];
        *reinterpret_cast<V*>(rho + idx) = rv[r];
      }
    }
  }
  mg_finish_dot<R, kThreads>(c, sc, partials, ticket, acc, 3);
}

// strips: advance the CG scalars from the sum-allreduced local_sum (the single-domain kernels do this themselves)
template <class R>
__global__ void k_mg_advance(MgFine<R> c, MgScalars* __restrict__ sc, int mode) {
  if (mode != 0 && sc->done) return;
  if (threadIdx.x == 0 && blockIdx.x == 0) mg_advance<R>(c, sc, sc->local_sum, mode);
}

// First smoothing sweep of a V-cycle: the reference's Jacobi update (src/model.rs:788-793) applied to z = 0,
//   z1 = omega * ((0 + 0 - rho) / denom) + (1 - omega) * 0,
// and its boundary update (:807-815) — pointwise, so it needs neither the zero field nor the stencil.
template <class R>
__global__ void __launch_bounds__(kMgThreads) k_mg_first_sweep(MgFine<R> c, R omega, R one_minus_omega, DivG<R> denom,
                                                                const R* __restrict__ rho, R* __restrict__ z,
                                                                const MgScalars* __restrict__ sc) {
  using V = typename Vec2<R>::type;
  if (sc->done) return;
  const MgPair<R> p = mg_pair<R>(c);
  if (!p.any) return;
  const int nx = c.nx;
  V rv[kMgRows];
#pragma unroll
  for (int r = 0; r < kMgRows; ++r)
    rv[r] = *reinterpret_cast<const V*>(rho + p.c0 + (size_t)min(p.j0 + r, c.row_hi - 1) * nx);
#pragma unroll
  for (int r = 0; r < kMgRows; ++r) {
    const int j = p.j0 + r;
    if (j < p.j1) {
      V o;
      o.x = omega * div_exact((R(0) + R(0)) - rv[r].x, denom) + one_minus_omega * R(0);
      o.y = omega * div_exact((R(0) + R(0)) - rv[r].y, denom) + one_minus_omega * R(0);
      if (p.c0 == 0) o.x = o.y;                             // p'[0,j] <- p'[1,j]
      if (p.c0 == nx - 2) o.y = c.cavity ? o.x : R(0);      // outlet 0 / cavity mirror
      *reinterpret_cast<V*>(z + p.c0 + (size_t)j * nx) = o;
      if (j == 1) *reinterpret_cast<V*>(z + p.c0) = o;                                   // bottom row <- row 1
      if (j == c.ny - 2) *reinterpret_cast<V*>(z + p.c0 + (size_t)(c.ny - 1) * nx) = o;  // top row <- row ny-2
    }
  }
}

// ---- level 0 <-> level 1 transfer ----

// rho_1[I,J] = sum over the children of (rho - L z); one thread per coarse cell
template <class R>
__global__ void __launch_bounds__(kMgThreads) k_mg_fine_restrict(MgFine<R> c, const R* __restrict__ z,
                                                                  const R* __restrict__ rho, int cmx, int c_lo,
                                                                  R* __restrict__ crho, const MgScalars* __restrict__ sc) {
  const int I = blockIdx.x * blockDim.x + threadIdx.x, J = c_lo + blockIdx.y;  // grid.y = owned rows of level 1
  if (I >= cmx || sc->done) return;
  auto children = [&](auto& dv) -> R {
    R acc = R(0);
#pragma unroll
    for (int b = 0; b < 2; ++b)
#pragma unroll
      for (int a = 0; a < 2; ++a) {
        const int i = 1 + 2 * I + a, j = 1 + 2 * J + b;
        if (i <= c.nx - 2 && j <= c.ny - 2) acc += rho[(size_t)i + (size_t)j * c.nx] - mg_fine_apply<R>(c, dv, z, i, j);
      }
    return acc;
  };
  DivTry<R> fast(c.ddx_sq);
  fast.also(c.ddy_sq);
  R acc = children(fast);
  if (__builtin_expect(!fast.ok(), 0)) {
    DivTrue<R> exact;
    acc = children(exact);
  }
  crho[(size_t)(I + 1) + (size_t)(J + 1) * (cmx + 2)] = acc;
}

// z += (correction of the parent), then the ring of z from its interior (the Jacobi boundary rules; corners are
// never read by a stencil on the unknowns and are left alone).  One thread per unknown.
template <class R>
__global__ void __launch_bounds__(kMgThreads) k_mg_fine_prolong(MgFine<R> c, R* __restrict__ z, int cmx,
                                                                 const R* __restrict__ ce, int j_lo,
                                                                 const MgScalars* __restrict__ sc) {
  // grid.y = rows handled: the owned unknown rows plus, on strips, the neighbours' edge rows (halo)
  const int i = 1 + blockIdx.x * blockDim.x + threadIdx.x, j = j_lo + blockIdx.y;
  if (i > c.nx - 2 || sc->done) return;
  const size_t idx = (size_t)i + (size_t)j * c.nx;
  const R v = z[idx] + ce[(size_t)((i - 1) / 2 + 1) + (size_t)((j - 1) / 2 + 1) * (cmx + 2)];
  z[idx] = v;
  if (i == 1) z[idx - 1] = v;
  if (i == c.nx - 2) z[idx + 1] = c.cavity ? v : R(0);
  if (j == 1) z[idx - c.nx] = v;
  if (j == c.ny - 2) z[idx + c.nx] = v;
}

// ---- fused smoothing passes of level 0 ---------------------------------------------------------------------------
// One damped-Jacobi sweep (the reference's update, src/model.rs:788-793, via jacobi_cell — the very function the
// tensor-TMA sweep kernel uses) whose INPUT field is formed on the fly instead of being read from memory:
//   kMode 0: input = the first smoothing sweep applied to z = 0, i.e. k_mg_first_sweep(rho)  -> replaces the pair
//            (k_mg_first_sweep, k_jacobi_sweep5) of a V(2,2) cycle by one pass that reads rho and writes z: 2 sN
//            instead of 5 sN of traffic;
//   kMode 1: input = z + (correction of the parent), i.e. k_mg_fine_prolong(z, ce)           -> replaces the pair
//            (k_mg_fine_prolong, k_jacobi_sweep5): 3.25 sN instead of 5.25 sN.
// The ring of the input follows the Jacobi boundary rules (:807-815: column 0 mirrors column 1, column nx-1 is zero at the
// channel outlet / mirrors column nx-2 in the cavity, rows 0 and ny-1 mirror rows 1 and ny-2), exactly as the kernels
// it replaces leave it; the output ring is written like the sweep kernel writes it.  Same per-cell arithmetic, same
// association: bit-identical to the unfused sequence (CFD_FLAG_MG_UNFUSED keeps that one for the cross-check).
// A thread owns the aligned column pair (2t, 2t+1) and kFsRows rows; grid = (ceil(nx / 512), ceil(rows / kFsRows)).
constexpr int kFsRows = 4;
template <class R, int kMode>
__global__ void __launch_bounds__(kMgThreads) k_mg_fused_sweep(const MgFine<R> c, const JacobiConsts2<R> c2,
                                                                const R* __restrict__ rho, const R* __restrict__ zin,
                                                                int cmx, const R* __restrict__ ce, R* __restrict__ zout,
                                                                const MgScalars* __restrict__ sc) {
  using V = typename Vec2<R>::type;
  if (sc->done) return;
  const int nx = c.nx, ny = c.ny;
  const int c0 = 2 * (blockIdx.x * blockDim.x + threadIdx.x);
  const int j0 = c.row_lo + blockIdx.y * kFsRows, j1 = min(j0 + kFsRows, c.row_hi);
  if (c0 >= nx || j0 >= j1) return;
  const bool ghost_l = c0 == 0, ghost_r = c0 == nx - 2;
  const R* __restrict__ src = kMode == 0 ? rho : zin;
  // every load of the tile first: the own pair on rows j0-1 .. j0+kFsRows (rows 0 and ny-1 mirror rows 1 and ny-2; rows
  // past the tile are loaded clamped and never used), the columns left / right of the pair and rho on the tile's rows
  V a[kFsRows + 2], ca[kFsRows + 2];
  R al[kFsRows], ar[kFsRows], cl[kFsRows], cr[kFsRows];
  V rh[kFsRows];
#pragma unroll
  for (int m = 0; m < kFsRows + 2; ++m) {
    const int j = min(max(min(j0 - 1 + m, j1), 1), ny - 2);
    a[m] = *reinterpret_cast<const V*>(src + c0 + (size_t)j * nx);
    if (kMode == 1) {  // correction of the parent cells of (c0, j) and (c0 + 1, j); column 0 has none (value unused)
      const size_t prow = (size_t)((j - 1) / 2 + 1) * (cmx + 2);
      ca[m].x = ce[prow + (size_t)((c0 - 1) / 2 + 1)];
      ca[m].y = ce[prow + (size_t)(c0 / 2 + 1)];
    }
  }
#pragma unroll
  for (int r = 0; r < kFsRows; ++r) {
    const int j = min(j0 + r, j1 - 1);
    const size_t row = (size_t)j * nx;
    al[r] = ghost_l ? R(0) : src[row + c0 - 1];   // unknown columns whenever they are used
    ar[r] = ghost_r ? R(0) : src[row + c0 + 2];
    if (kMode == 1) {
      const size_t prow = (size_t)((j - 1) / 2 + 1) * (cmx + 2);
      cl[r] = ghost_l ? R(0) : ce[prow + (size_t)((c0 - 2) / 2 + 1)];
      cr[r] = ghost_r ? R(0) : ce[prow + (size_t)((c0 + 1) / 2 + 1)];
    }
    rh[r] = *reinterpret_cast<const V*>(rho + c0 + row);
  }
  // the input field, ring rules applied (mode 0 divides: DivTry for the tile, DivTrue if a quotient left its window)
  V in[kFsRows + 2];
  R inl[kFsRows], inr[kFsRows];
  auto form = [&](auto& dv) {
    auto raw1 = [&](R x, R corr) -> R {
      if (kMode == 0) return c2.omega * dv((R(0) + R(0)) - x, c2.denom) + c2.one_minus_omega * R(0);
      return x + corr;
    };
#pragma unroll
    for (int m = 0; m < kFsRows + 2; ++m) {
      V o;
      o.x = raw1(a[m].x, ca[m].x);
      o.y = raw1(a[m].y, ca[m].y);
      if (ghost_l) o.x = o.y;                          // column 0 <- column 1
      if (ghost_r) o.y = c.cavity ? o.x : R(0);        // outlet column 0 / cavity mirror
      in[m] = o;
    }
#pragma unroll
    for (int r = 0; r < kFsRows; ++r) {
      inl[r] = raw1(al[r], cl[r]);
      inr[r] = raw1(ar[r], cr[r]);
    }
  };
  DivTry<R> fast(c2.denom);
  form(fast);
  if (kMode == 0 && __builtin_expect(!fast.ok(), 0)) {
    DivTrue<R> exact;
    form(exact);
  }
#pragma unroll
  for (int r = 0; r < kFsRows; ++r) {
    const int j = j0 + r;
    if (j >= j1) break;
    const V cen = in[r + 1], south = in[r], north = in[r + 2];
    R n0 = jacobi_cell<R>(c2, inl[r], cen.y, north.x, south.x, cen.x, rh[r].x);
    R n1 = jacobi_cell<R>(c2, cen.x, inr[r], north.y, south.y, cen.y, rh[r].y);
    if (ghost_l) n0 = n1;
    if (ghost_r) n1 = c.cavity ? n0 : R(0);
    V out;
    out.x = n0; out.y = n1;
    *reinterpret_cast<V*>(zout + c0 + (size_t)j * nx) = out;
    if (j == 1) *reinterpret_cast<V*>(zout + c0) = out;                                   // bottom row <- row 1
    if (j == ny - 2) *reinterpret_cast<V*>(zout + c0 + (size_t)(ny - 1) * nx) = out;      // top row <- row ny-2
  }
}

// ---- coarse levels (l >= 1): fields (mx + 2) x (my + 2) with a ring of zeros ----
// Per-cell operations, shared by the one-launch-per-operation kernels (large levels) and the single-block kernel
// that runs the whole bottom of the V-cycle (k_mg_bottom).  Plain pointers: inside k_mg_bottom the fields are
// written and re-read by the same launch, so no read-only (non-coherent) loads may be used on them.
template <class R>
__device__ __forceinline__ R mg_coarse_apply(const MgLevelDev<R>& L, const R* e, int I, int J, R cc, bool zero_in) {
  if (zero_in) cc = R(0);
  const size_t W = (size_t)L.mx + 2, idx = (size_t)(I + 1) + (size_t)(J + 1) * W;
  const R ee = zero_in ? R(0) : e[idx + 1], ew = zero_in ? R(0) : e[idx - 1];
  const R en = zero_in ? R(0) : e[idx + W], es = zero_in ? R(0) : e[idx - W];
  return L.CXH[J] * (L.WE[I] * (ee - cc) + L.WW[I] * (ew - cc)) + L.CYW[I] * (L.WN[J] * (en - cc) + L.WS[J] * (es - cc));
}

// one cell of a damped-Jacobi sweep: out = in + omega * ((L in - rho) / diag)  (0 where the diagonal vanishes: the
// 1 x 1 level of the all-Neumann cavity); zero_in: `in` is taken as 0 without being read
template <class R>
__device__ __forceinline__ void mgc_sweep_cell(const MgLevelDev<R>& L, const R* in, const R* rho, R* out, R omega,
                                               bool zero_in, int I, int J) {
  const size_t idx = (size_t)(I + 1) + (size_t)(J + 1) * ((size_t)L.mx + 2);
  const R diag = L.CXH[J] * (L.WE[I] + L.WW[I]) + L.CYW[I] * (L.WN[J] + L.WS[J]);
  const R cc = zero_in ? R(0) : in[idx];
  const R le = mg_coarse_apply<R>(L, in, I, J, cc, zero_in);
  out[idx] = diag > R(0) ? cc + omega * ((le - rho[idx]) / diag) : R(0);
}

// rho_{l+1}[I,J] = sum over the children of (rho_l - L_l e)
template <class R>
__device__ __forceinline__ void mgc_restrict_cell(const MgLevelDev<R>& L, const R* e, const R* rho, int cmx, R* crho,
                                                  int I, int J) {
  const size_t W = (size_t)L.mx + 2;
  R acc = R(0);
#pragma unroll
  for (int b = 0; b < 2; ++b)
#pragma unroll
    for (int a = 0; a < 2; ++a) {
      const int i = 2 * I + a, j = 2 * J + b;
      if (i < L.mx && j < L.my) {
        const size_t idx = (size_t)(i + 1) + (size_t)(j + 1) * W;
        acc += rho[idx] - mg_coarse_apply<R>(L, e, i, j, e[idx], false);
      }
    }
  crho[(size_t)(I + 1) + (size_t)(J + 1) * ((size_t)cmx + 2)] = acc;
}

template <class R>
__global__ void __launch_bounds__(kMgThreads) k_mgc_sweep(MgLevelDev<R> L, const R* in, const R* rho, R* out, R omega,
                                                           int zero_in, int row_lo, const MgScalars* __restrict__ sc) {
  const int I = blockIdx.x * blockDim.x + threadIdx.x, J = row_lo + blockIdx.y;
  if (sc->done) return;
  if (I < L.mx) mgc_sweep_cell<R>(L, in, rho, out, omega, zero_in != 0, I, J);
}

// fills a level's diagonal table from the diagonal values (one thread per entry)
// table[n] = the intersection of the dividend windows of the positive entries (y, r unused): what a DivTry over
// divisions by any of them tests against (cfd_mg_legs.cuh).  One block.
template <class R>
__global__ void k_mgc_diag_table(const R* __restrict__ diag, DivG<R>* __restrict__ table, int n) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k < n) table[k] = make_divg(diag[k]);
  __syncthreads();
  if (k == 0) {
    unsigned lo = 0u, end = 0xffffffffu;
    bool any = false;
    for (int m = 0; m < n; ++m)
      if (table[m].y > R(0)) {
        const unsigned e = table[m].span == 0u ? table[m].lo : table[m].lo + table[m].span;
        lo = table[m].lo > lo ? table[m].lo : lo;
        end = e < end ? e : end;
        any = true;
      }
    DivG<R> w;
    w.y = R(1); w.r = R(1);
    w.lo = lo;
    w.span = any && end > lo ? end - lo : 0u;
    table[n] = w;
  }
}

// k_mgc_sweep for a level with a diagonal table: same per-cell arithmetic (mgc_sweep_cell), a thread owns column I and
// kMgcRows rows — the column's weights are loaded once, the centre column's rows rotate through registers, every load of
// the tile is issued before the arithmetic, and the division by the diagonal is DivTry (DivTrue for the rare tile with a
// dividend outside the window).  grid = (ceil(mx / 256), ceil(rows / kMgcRows)).
constexpr int kMgcRows = 4;
template <class R>
__global__ void __launch_bounds__(kMgThreads) k_mgc_sweep_tab(MgLevelDev<R> L, const R* __restrict__ in,
                                                               const R* __restrict__ rho, R* __restrict__ out, R omega,
                                                               int zero_in, int row_lo, int row_hi,
                                                               const MgScalars* __restrict__ sc) {
  const int I = blockIdx.x * blockDim.x + threadIdx.x;
  const int J0 = row_lo + blockIdx.y * kMgcRows, J1 = min(J0 + kMgcRows, row_hi);
  if (sc->done || I >= L.mx || J0 >= J1) return;
  const size_t W = (size_t)L.mx + 2;
  const R we = L.WE[I], ww = L.WW[I], cyw = L.CYW[I];
  const int cc_cls = L.col_class[I];
  R ec[kMgcRows + 2], el[kMgcRows], er[kMgcRows], rh[kMgcRows];
#pragma unroll
  for (int m = 0; m < kMgcRows + 2; ++m) {
    const int J = min(J0 - 1 + m, J1);  // array row J + 1; rows -1 and my are the ring of zeros
    ec[m] = zero_in ? R(0) : in[(size_t)(I + 1) + (size_t)(J + 1) * W];
  }
#pragma unroll
  for (int r = 0; r < kMgcRows; ++r) {
    const int J = min(J0 + r, J1 - 1);
    const size_t idx = (size_t)(I + 1) + (size_t)(J + 1) * W;
    el[r] = zero_in ? R(0) : in[idx - 1];
    er[r] = zero_in ? R(0) : in[idx + 1];
    rh[r] = rho[idx];
  }
  R num[kMgcRows], res[kMgcRows];
  DivG<R> dg[kMgcRows];
#pragma unroll
  for (int r = 0; r < kMgcRows; ++r) {
    const int J = min(J0 + r, J1 - 1);
    const R cxh = L.CXH[J], wn = L.WN[J], ws = L.WS[J];
    const R cc = ec[r + 1];
    // mg_coarse_apply, same association
    const R le = cxh * (we * (er[r] - cc) + ww * (el[r] - cc)) + cyw * (wn * (ec[r + 2] - cc) + ws * (ec[r] - cc));
    num[r] = le - rh[r];
    dg[r] = L.diag_table[L.row_class[J] * kMgClasses + cc_cls];
  }
  DivTry<R> dv(dg[0]);
#pragma unroll
  for (int r = 1; r < kMgcRows; ++r) dv.also(dg[r]);
#pragma unroll
  for (int r = 0; r < kMgcRows; ++r) res[r] = dv(num[r], dg[r]);
  if (__builtin_expect(!dv.ok(), 0)) {
#pragma unroll
    for (int r = 0; r < kMgcRows; ++r) res[r] = num[r] / dg[r].y;
  }
#pragma unroll
  for (int r = 0; r < kMgcRows; ++r) {
    const int J = J0 + r;
    if (J < J1) out[(size_t)(I + 1) + (size_t)(J + 1) * W] = dg[r].y > R(0) ? ec[r + 1] + omega * res[r] : R(0);
  }
}

template <class R>
__global__ void __launch_bounds__(kMgThreads) k_mgc_restrict(MgLevelDev<R> L, const R* e, const R* rho, int cmx,
                                                              int c_lo, R* crho, const MgScalars* __restrict__ sc) {
  const int I = blockIdx.x * blockDim.x + threadIdx.x, J = c_lo + blockIdx.y;
  if (sc->done) return;
  if (I < cmx) mgc_restrict_cell<R>(L, e, rho, cmx, crho, I, J);
}

// e_l += (correction of the parent); one thread per cell of level l
template <class R>
__global__ void __launch_bounds__(kMgThreads) k_mgc_prolong(int mx, R* e, int cmx, const R* ce, int row_lo,
                                                             const MgScalars* __restrict__ sc) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x, j = row_lo + blockIdx.y;
  if (i >= mx || sc->done) return;
  const size_t idx = (size_t)(i + 1) + (size_t)(j + 1) * ((size_t)mx + 2);
  e[idx] += ce[(size_t)(i / 2 + 1) + (size_t)(j / 2 + 1) * ((size_t)cmx + 2)];
}

// The bottom of the V-cycle in ONE launch: every level from the first one that fits 64 x 64 down to 1 x 1 and
// back up, by a single block (the fields stay in L1/L2; phases are separated by __syncthreads).  Replaces ~45
// launches of a few microseconds each.  lv[0].rho is the input, the correction ends up in lv[0].e.
constexpr int kMgBottomMax = 8;
constexpr int kMgBottomThreads = 1024;
template <class R>
struct MgBottomLevel {
  MgLevelDev<R> dev;
  R *e, *rho, *tmp;
};
template <class R>
struct MgBottom {
  int n, nu;
  R omega;
  MgBottomLevel<R> lv[kMgBottomMax];
};

// kSmem: every field of every bottom level lives in dynamic shared memory (3 x (mx + 2) x (my + 2) values per level,
// ~145 KB for a 64 x 64 first level): a phase then costs a barrier and shared-memory latency instead of an L2 round trip
// (~1.2 us per phase, 7 levels x (2 nu + 2) phases).  Same per-cell functions on other pointers: bit-identical.
template <class R, bool kSmem>
__global__ void __launch_bounds__(kMgBottomThreads) k_mg_bottom(const MgBottom<R> B, const MgScalars* __restrict__ sc) {
  extern __shared__ __align__(16) unsigned char mg_bottom_raw[];
  if (sc->done) return;
  R* cur[kMgBottomMax];
  R* oth[kMgBottomMax];
  R *fe[kMgBottomMax], *frho[kMgBottomMax], *ftmp[kMgBottomMax];  // the fields the phases work on
  const int tid = threadIdx.x;
  if (kSmem) {
    R* p = reinterpret_cast<R*>(mg_bottom_raw);
    size_t total = 0;
    for (int l = 0; l < B.n; ++l) {
      const size_t n = (size_t)(B.lv[l].dev.mx + 2) * (size_t)(B.lv[l].dev.my + 2);
      fe[l] = p + total; frho[l] = p + total + n; ftmp[l] = p + total + 2 * n;
      total += 3 * n;
    }
    for (size_t k = tid; k < total; k += kMgBottomThreads) p[k] = R(0);  // the rings of zeros (and everything else)
    __syncthreads();
    const int mx = B.lv[0].dev.mx, my = B.lv[0].dev.my;
    for (int k = tid; k < mx * my; k += kMgBottomThreads) {
      const size_t idx = (size_t)(k % mx + 1) + (size_t)(k / mx + 1) * ((size_t)mx + 2);
      frho[0][idx] = B.lv[0].rho[idx];
    }
    __syncthreads();
  } else {
    for (int l = 0; l < B.n; ++l) { fe[l] = B.lv[l].e; frho[l] = B.lv[l].rho; ftmp[l] = B.lv[l].tmp; }
  }
  for (int l = 0; l < B.n; ++l) {
    const MgBottomLevel<R>& L = B.lv[l];
    const int mx = L.dev.mx, my = L.dev.my, cells = mx * my;
    R *a = fe[l], *b = ftmp[l];
    if (mx == 1 && my == 1) {  // exact
      if (tid == 0) mgc_sweep_cell<R>(L.dev, a, frho[l], b, R(1), true, 0, 0);
      cur[l] = b; oth[l] = a;
      __syncthreads();
      break;
    }
    for (int s = 0; s < B.nu; ++s) {
      for (int k = tid; k < cells; k += kMgBottomThreads) mgc_sweep_cell<R>(L.dev, a, frho[l], b, B.omega, s == 0, k % mx, k / mx);
      __syncthreads();
      R* t = a; a = b; b = t;
    }
    cur[l] = a; oth[l] = b;
    const MgBottomLevel<R>& C = B.lv[l + 1];
    const int cmx = C.dev.mx, ccells = cmx * C.dev.my;
    for (int k = tid; k < ccells; k += kMgBottomThreads) mgc_restrict_cell<R>(L.dev, a, frho[l], cmx, frho[l + 1], k % cmx, k / cmx);
    __syncthreads();
  }
  for (int l = B.n - 2; l >= 0; --l) {
    const MgBottomLevel<R>& L = B.lv[l];
    const int mx = L.dev.mx, my = L.dev.my, cells = mx * my;
    const int cmx = B.lv[l + 1].dev.mx;
    R *a = cur[l], *b = oth[l];
    const R* ce = cur[l + 1];
    for (int k = tid; k < cells; k += kMgBottomThreads) {
      const int i = k % mx, j = k / mx;
      a[(size_t)(i + 1) + (size_t)(j + 1) * ((size_t)mx + 2)] += ce[(size_t)(i / 2 + 1) + (size_t)(j / 2 + 1) * ((size_t)cmx + 2)];
    }
    __syncthreads();
    for (int s = 0; s < B.nu; ++s) {
      for (int k = tid; k < cells; k += kMgBottomThreads) mgc_sweep_cell<R>(L.dev, a, frho[l], b, B.omega, false, k % mx, k / mx);
      __syncthreads();
      R* t = a; a = b; b = t;
    }
    cur[l] = a; oth[l] = b;
  }
  // after nu + nu swaps the correction of a (non-trivial) level is back in its `e` buffer, where the caller expects it
  if (kSmem) {
    const int mx = B.lv[0].dev.mx, my = B.lv[0].dev.my;
    for (int k = tid; k < mx * my; k += kMgBottomThreads) {
      const size_t idx = (size_t)(k % mx + 1) + (size_t)(k / mx + 1) * ((size_t)mx + 2);
      B.lv[0].e[idx] = cur[0][idx];
    }
  }
}

}  // namespace cfdk
