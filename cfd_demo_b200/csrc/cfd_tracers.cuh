// cfd_tracers.cuh — SURVEY 8f row 4: the JS twin's tracer particles (index.html:1472-1543) on the resident fields.
//
// A tracer is a point (x, y) advected with the bilinearly interpolated CELL-CENTRED velocity (getVelocityAt, :1499-1526;
// cell centre velocity = mean of the cell's two u faces / two v faces, :1513-1517) by explicit Euler (updateTracers,
// :1485-1497); tracers that leave [0, lx] x [0, ly] are dropped, the others keep their order.  New tracers start on the
// inlet, one per cell row at (0, (j + 1/2) dy) (initTracers / injectTracers, :1475-1483, :1537-1543).
// The JS does this in double arithmetic on its Float32Array fields; here the fields are the model's (R = double for the
// shipped path, float for the reference's own precision) and the tracer arithmetic is double either way — with
// precision 32 that is exactly the JS's.  No FMA contraction (this translation unit is built with -fmad=false).
#pragma once

#include <cuda_runtime.h>

namespace cfdk {

struct TracerGeom {
  int nx, ny;
  double dx, dy, lx, ly;
};

// getVelocityAt, index.html:1499-1526
template <class R>
__device__ __forceinline__ double2 tracer_velocity(const TracerGeom& g, const R* __restrict__ u, const R* __restrict__ v,
                                                   double x, double y) {
  int i = (int)floor(x / g.dx);
  int j = (int)floor(y / g.dy);
  if (i < 0) i = 0;
  if (i > g.nx - 2) i = g.nx - 2;
  if (j < 0) j = 0;
  if (j > g.ny - 2) j = g.ny - 2;
  const double rx = (x - i * g.dx) / g.dx;
  const double ry = (y - j * g.dy) / g.dy;
  const size_t W = (size_t)g.nx + 1;
  auto cu = [&](int ii, int jj) { return 0.5 * ((double)u[(size_t)ii + (size_t)jj * W] + (double)u[(size_t)(ii + 1) + (size_t)jj * W]); };
  auto cv = [&](int ii, int jj) { return 0.5 * ((double)v[(size_t)ii + (size_t)jj * g.nx] + (double)v[(size_t)ii + (size_t)(jj + 1) * g.nx]); };
  const double u00 = cu(i, j), u10 = cu(i + 1, j), u01 = cu(i, j + 1), u11 = cu(i + 1, j + 1);
  const double v00 = cv(i, j), v10 = cv(i + 1, j), v01 = cv(i, j + 1), v11 = cv(i + 1, j + 1);
  double2 o;
  o.x = (1 - rx) * ((1 - ry) * u00 + ry * u01) + rx * ((1 - ry) * u10 + ry * u11);
  o.y = (1 - rx) * ((1 - ry) * v00 + ry * v01) + rx * ((1 - ry) * v10 + ry * v11);
  return o;
}

// updateTracers, index.html:1485-1497: advance every tracer, flag the ones that stay inside the domain
template <class R>
__global__ void __launch_bounds__(256) k_tracers_advect(TracerGeom g, const R* __restrict__ u, const R* __restrict__ v,
                                                        double dt, double2* __restrict__ pos, unsigned char* __restrict__ keep,
                                                        unsigned n) {
  const unsigned t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  double2 p = pos[t];
  const double2 vel = tracer_velocity<R>(g, u, v, p.x, p.y);
  p.x += dt * vel.x;
  p.y += dt * vel.y;
  pos[t] = p;
  keep[t] = (p.x >= 0 && p.x <= g.lx && p.y >= 0 && p.y <= g.ly) ? 1 : 0;  // NaN fails every comparison: dropped, like the JS
}

// stable compaction (the JS rebuilds its array in order): one block, chunks of 1024 tracers, running offset
__global__ void __launch_bounds__(1024) k_tracers_compact(const double2* __restrict__ in, const unsigned char* __restrict__ keep,
                                                          unsigned n, double2* __restrict__ out, unsigned* __restrict__ n_out) {
  __shared__ unsigned s_warp[32];
  __shared__ unsigned s_base;
  const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) s_base = 0u;
  __syncthreads();
  for (unsigned start = 0; start < n; start += 1024u) {
    const unsigned t = start + threadIdx.x;
    const unsigned k = (t < n && keep[t]) ? 1u : 0u;
    const unsigned ballot = __ballot_sync(0xffffffffu, k != 0u);
    const unsigned before = __popc(ballot & ((1u << lane) - 1u));
    if (lane == 0) s_warp[warp] = __popc(ballot);
    __syncthreads();
    unsigned warp_off = 0u, total = 0u;
    for (unsigned w = 0; w < 32u; ++w) {
      const unsigned c = s_warp[w];
      if (w < warp) warp_off += c;
      total += c;
    }
    const unsigned base = s_base;
    if (k) out[base + warp_off + before] = in[t];
    __syncthreads();
    if (threadIdx.x == 0) s_base = base + total;
    __syncthreads();
  }
  if (threadIdx.x == 0) *n_out = s_base;
}

// initTracers / injectTracers, index.html:1475-1483, :1537-1543: append one tracer per cell row on the inlet
__global__ void k_tracers_inject(TracerGeom g, double2* __restrict__ pos, unsigned n_before) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= g.ny) return;
  double2 p;
  p.x = 0.0;
  p.y = (j + 0.5) * g.dy;
  pos[n_before + (unsigned)j] = p;
}

}  // namespace cfdk
