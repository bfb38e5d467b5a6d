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

struct MgScalars {
  double rr, rz, dw, alpha, beta, measure;
  int done, iterations, max_iterations, pad;
};

template <class R>
struct MgFine {
  R dx_sq, dy_sq, dt, tol, n_unknowns;
  int nx, ny, cavity;
};

template <class R>
struct MgLevelDev {
  int mx, my;               // unknowns per direction; fields are (mx + 2) x (my + 2)
  const R *WE, *WW, *CYW;   // per column
  const R *WN, *WS, *CXH;   // per row
};

constexpr int kMgThreads = 256;

// ---- level 0 vector kernels (grid: (ceil(nx / 256), ny) for init, (ceil((nx - 2) / 256), ny - 2) otherwise) ----

// x = 0, d = 0, rho = rhs on the unknowns (0 on the ring), partial rho.rho per block
template <class R>
__global__ void __launch_bounds__(kMgThreads) k_mg_init(MgFine<R> c, const R* __restrict__ rhs, R* __restrict__ x,
                                                         R* __restrict__ rho, R* __restrict__ d,
                                                         double* __restrict__ partials) {
  __shared__ double s_red[kMgThreads / 32];
  const int i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y;
  double acc = 0.0;
  if (i < c.nx) {
    const size_t idx = (size_t)i + (size_t)j * c.nx;
    const bool unknown = (i >= 1 && i <= c.nx - 2 && j >= 1 && j <= c.ny - 2);
    const R b = unknown ? rhs[idx] : R(0);
    x[idx] = R(0);
    d[idx] = R(0);
    rho[idx] = b;
    acc = (double)(b * b);
  }
  const double t = block_sum<kMgThreads / 32>(acc, s_red);
  if (threadIdx.x == 0) partials[blockIdx.y * gridDim.x + blockIdx.x] = t;
}

template <class R>
__device__ __forceinline__ R mg_fine_apply(const MgFine<R>& c, const R* __restrict__ x, int i, int j) {
  const size_t idx = (size_t)i + (size_t)j * c.nx;
  const R cc = x[idx];
  const R xe = (i == c.nx - 2) ? (c.cavity ? cc : R(0)) : x[idx + 1];
  const R xw = (i == 1) ? cc : x[idx - 1];
  const R xn = (j == c.ny - 2) ? cc : x[idx + c.nx];
  const R xs = (j == 1) ? cc : x[idx - c.nx];
  return ((xe - cc) + (xw - cc)) / c.dx_sq + ((xn - cc) + (xs - cc)) / c.dy_sq;
}

// partial a.b over the unknowns
template <class R>
__global__ void __launch_bounds__(kMgThreads) k_mg_dot(MgFine<R> c, const R* __restrict__ a, const R* __restrict__ b,
                                                        double* __restrict__ partials) {
  __shared__ double s_red[kMgThreads / 32];
  const int i = 1 + blockIdx.x * blockDim.x + threadIdx.x, j = 1 + blockIdx.y;
  double acc = 0.0;
  if (i <= c.nx - 2) {
    const size_t idx = (size_t)i + (size_t)j * c.nx;
    acc = (double)(a[idx] * b[idx]);
  }
  const double t = block_sum<kMgThreads / 32>(acc, s_red);
  if (threadIdx.x == 0) partials[blockIdx.y * gridDim.x + blockIdx.x] = t;
}

// d = z + beta d
template <class R>
__global__ void __launch_bounds__(kMgThreads) k_mg_direction(MgFine<R> c, const MgScalars* __restrict__ sc,
                                                              const R* __restrict__ z, R* __restrict__ d) {
  const R beta = (R)sc->beta;
  const int i = 1 + blockIdx.x * blockDim.x + threadIdx.x, j = 1 + blockIdx.y;
  if (i <= c.nx - 2) {
    const size_t idx = (size_t)i + (size_t)j * c.nx;
    d[idx] = z[idx] + beta * d[idx];
  }
}

// w = L d, partial d.w
template <class R>
__global__ void __launch_bounds__(kMgThreads) k_mg_apply(MgFine<R> c, const R* __restrict__ d, R* __restrict__ w,
                                                          double* __restrict__ partials) {
  __shared__ double s_red[kMgThreads / 32];
  const int i = 1 + blockIdx.x * blockDim.x + threadIdx.x, j = 1 + blockIdx.y;
  double acc = 0.0;
  if (i <= c.nx - 2) {
    const size_t idx = (size_t)i + (size_t)j * c.nx;
    const R lw = mg_fine_apply<R>(c, d, i, j);
    w[idx] = lw;
    acc = (double)(d[idx] * lw);
  }
  const double t = block_sum<kMgThreads / 32>(acc, s_red);
  if (threadIdx.x == 0) partials[blockIdx.y * gridDim.x + blockIdx.x] = t;
}

// x += alpha d, rho -= alpha w, partial rho.rho
template <class R>
__global__ void __launch_bounds__(kMgThreads) k_mg_update(MgFine<R> c, const MgScalars* __restrict__ sc,
                                                           const R* __restrict__ d, const R* __restrict__ w,
                                                           R* __restrict__ x, R* __restrict__ rho,
                                                           double* __restrict__ partials) {
  __shared__ double s_red[kMgThreads / 32];
  const R alpha = (R)sc->alpha;
  const int i = 1 + blockIdx.x * blockDim.x + threadIdx.x, j = 1 + blockIdx.y;
  double acc = 0.0;
  if (i <= c.nx - 2) {
    const size_t idx = (size_t)i + (size_t)j * c.nx;
    x[idx] = x[idx] + alpha * d[idx];
    const R rn = rho[idx] - alpha * w[idx];
    rho[idx] = rn;
    acc = (double)(rn * rn);
  }
  const double t = block_sum<kMgThreads / 32>(acc, s_red);
  if (threadIdx.x == 0) partials[blockIdx.y * gridDim.x + blockIdx.x] = t;
}

// one block: sums the per-block partials in a fixed order, then advances the scalars.
// mode 0: rho.rho after init; 1: rho.z -> beta (0 before the first iteration); 2: d.w -> alpha;
// 3: rho.rho after the update -> iteration count, stopping rule (same measure as k_cg_reduce)
template <class R>
__global__ void __launch_bounds__(1024) k_mg_reduce(MgFine<R> c, MgScalars* __restrict__ sc,
                                                     const double* __restrict__ partials, int n, int mode) {
  __shared__ double s_red[32];
  double acc = 0.0;
  for (int k = threadIdx.x; k < n; k += blockDim.x) acc += partials[k];
  const double t = block_sum<32>(acc, s_red);
  if (threadIdx.x == 0) {
    const R sum = (R)t;
    if (mode == 1) {
      sc->beta = sc->iterations == 0 ? 0.0 : (double)(sum / (R)sc->rz);
      sc->rz = (double)sum;
    } else if (mode == 2) {
      sc->dw = (double)sum;
      sc->alpha = (double)((R)sc->rz / sum);
    } else {
      if (mode == 3) sc->iterations += 1;
      sc->rr = (double)sum;
      const R measure = c.dt * (R)sqrt((double)(sum / c.n_unknowns));
      sc->measure = (double)measure;
      if (measure <= c.tol || sc->iterations >= sc->max_iterations) sc->done = 1;
    }
  }
}

// ---- level 0 <-> level 1 transfer ----

// rho_1[I,J] = sum over the children of (rho - L z); one thread per coarse cell
template <class R>
__global__ void __launch_bounds__(kMgThreads) k_mg_fine_restrict(MgFine<R> c, const R* __restrict__ z,
                                                                  const R* __restrict__ rho, int cmx, int cmy,
                                                                  R* __restrict__ crho) {
  const int I = blockIdx.x * blockDim.x + threadIdx.x, J = blockIdx.y;
  if (I >= cmx) return;
  R acc = R(0);
#pragma unroll
  for (int b = 0; b < 2; ++b)
#pragma unroll
    for (int a = 0; a < 2; ++a) {
      const int i = 1 + 2 * I + a, j = 1 + 2 * J + b;
      if (i <= c.nx - 2 && j <= c.ny - 2) acc += rho[(size_t)i + (size_t)j * c.nx] - mg_fine_apply<R>(c, z, i, j);
    }
  crho[(size_t)(I + 1) + (size_t)(J + 1) * (cmx + 2)] = acc;
}

// z += (correction of the parent), then the ring of z from its interior (the Jacobi boundary rules; corners are
// never read by a stencil on the unknowns and are left alone).  One thread per unknown.
template <class R>
__global__ void __launch_bounds__(kMgThreads) k_mg_fine_prolong(MgFine<R> c, R* __restrict__ z, int cmx,
                                                                 const R* __restrict__ ce) {
  const int i = 1 + blockIdx.x * blockDim.x + threadIdx.x, j = 1 + blockIdx.y;
  if (i > c.nx - 2) return;
  const size_t idx = (size_t)i + (size_t)j * c.nx;
  const R v = z[idx] + ce[(size_t)((i - 1) / 2 + 1) + (size_t)((j - 1) / 2 + 1) * (cmx + 2)];
  z[idx] = v;
  if (i == 1) z[idx - 1] = v;
  if (i == c.nx - 2) z[idx + 1] = c.cavity ? v : R(0);
  if (j == 1) z[idx - c.nx] = v;
  if (j == c.ny - 2) z[idx + c.nx] = v;
}

// ---- coarse levels (l >= 1): fields (mx + 2) x (my + 2) with a ring of zeros ----
template <class R>
__device__ __forceinline__ R mg_coarse_apply(const MgLevelDev<R>& L, const R* __restrict__ e, int I, int J, R cc,
                                             bool zero_in) {
  if (zero_in) cc = R(0);
  const size_t W = (size_t)L.mx + 2, idx = (size_t)(I + 1) + (size_t)(J + 1) * W;
  const R ee = zero_in ? R(0) : e[idx + 1], ew = zero_in ? R(0) : e[idx - 1];
  const R en = zero_in ? R(0) : e[idx + W], es = zero_in ? R(0) : e[idx - W];
  return L.CXH[J] * (L.WE[I] * (ee - cc) + L.WW[I] * (ew - cc)) + L.CYW[I] * (L.WN[J] * (en - cc) + L.WS[J] * (es - cc));
}

// one damped-Jacobi sweep: out = in + omega * ((L in - rho) / diag)  (0 where the diagonal vanishes: the 1 x 1
// level of the all-Neumann cavity); zero_in: `in` is taken as 0 without being read
template <class R>
__global__ void __launch_bounds__(kMgThreads) k_mgc_sweep(MgLevelDev<R> L, const R* __restrict__ in,
                                                           const R* __restrict__ rho, R* __restrict__ out, R omega,
                                                           int zero_in) {
  const int I = blockIdx.x * blockDim.x + threadIdx.x, J = blockIdx.y;
  if (I >= L.mx) return;
  const size_t idx = (size_t)(I + 1) + (size_t)(J + 1) * ((size_t)L.mx + 2);
  const R diag = L.CXH[J] * (L.WE[I] + L.WW[I]) + L.CYW[I] * (L.WN[J] + L.WS[J]);
  const R cc = zero_in ? R(0) : in[idx];
  const R le = mg_coarse_apply<R>(L, in, I, J, cc, zero_in != 0);
  out[idx] = diag > R(0) ? cc + omega * ((le - rho[idx]) / diag) : R(0);
}

// rho_{l+1}[I,J] = sum over the children of (rho_l - L_l e); one thread per cell of level l+1
template <class R>
__global__ void __launch_bounds__(kMgThreads) k_mgc_restrict(MgLevelDev<R> L, const R* __restrict__ e,
                                                              const R* __restrict__ rho, int cmx, int cmy,
                                                              R* __restrict__ crho) {
  const int I = blockIdx.x * blockDim.x + threadIdx.x, J = blockIdx.y;
  if (I >= cmx) return;
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

// e_l += (correction of the parent); one thread per cell of level l
template <class R>
__global__ void __launch_bounds__(kMgThreads) k_mgc_prolong(int mx, R* __restrict__ e, int cmx, const R* __restrict__ ce) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y;
  if (i >= mx) return;
  const size_t idx = (size_t)(i + 1) + (size_t)(j + 1) * ((size_t)mx + 2);
  e[idx] += ce[(size_t)(i / 2 + 1) + (size_t)(j / 2 + 1) * ((size_t)cmx + 2)];
}

}  // namespace cfdk
