// cfd_sweeps_ab.cuh — the Jacobi sweep kernels that LOST their A/B against k_jacobi_sweep5 (cfd_kernels.cuh): kept for the
// record and for the cross-check test, compiled only into libcfd_b200_ab.so (-DCFD_WITH_AB_SWEEPS), not into the product
// library.  Same arithmetic as the shipped kernel (bit-identical results: tests/test_gpu_parity.py); measured slower at
// 4096^2 (profiles/r1_notes.md): k_jacobi_sweep2 register prefetch, k_jacobi_sweep3 row-by-row bulk copies, k_jacobi_sweep4
// one row per step, k_jacobi_sweep6 persistent warp queue (87.5 us vs 78), k_jacobi_sweep_t2 two sweeps per pass (issue-bound).
#pragma once

namespace cfdk {

// k_jacobi_sweep2 — register-prefetch kernel (CFD_FLAG_REGISTER_SWEEP)
template <class R>
__device__ __forceinline__ RowRegs<R> load_row(const R* __restrict__ p, int off_l, int off_r) {
  using V = typename Vec2<R>::type;
  RowRegs<R> o;
  const V c = __ldg(reinterpret_cast<const V*>(p));
  o.x = c.x;
  o.y = c.y;
  o.l = __ldg(p + off_l);
  o.r = __ldg(p + off_r);
  return o;
}

// NOTE: p and rhs must be readable up to 3 rows past row ny-1 (the prefetch runs ahead without clamping);
// the model allocates its p', p'new and rhs buffers with that slack.
template <class R>
__global__ void __launch_bounds__(128, 4) k_jacobi_sweep2(JacobiConsts2<R> c, const R* __restrict__ p,
                                                          const R* __restrict__ rhs, R* __restrict__ pn,
                                                          unsigned long long* __restrict__ err_slots, int sweep) {
  using V = typename Vec2<R>::type;
  __shared__ double s_red[4];
  if (sweep > 0) {
    const R prev = (R)bits_nonneg(err_slots[sweep - 1]);
    if (prev < c.tol) return;
  }
  const int nx = c.nx, ny = c.ny;
  const int c0 = 2 * (blockIdx.x * blockDim.x + threadIdx.x);
  const int j0 = c.row_begin + blockIdx.y * c.rows_per_block;
  const int j1 = min(j0 + c.rows_per_block, c.row_end);  // rows [j0, j1)
  double max_err = 0.0;
  if (c0 < nx && j0 < j1) {
    const bool ghost_l = (c0 == 0), ghost_r = (c0 == nx - 2);
    const bool cnt0 = (c0 >= 1) && (c0 <= nx - kLanes), cnt1 = (c0 + 1 <= nx - kLanes);
    const int off_l = ghost_l ? 0 : -1, off_r = ghost_r ? 1 : 2;
    const R* pc = p + c0 + (size_t)(j0 - 1) * nx;   // row j0-1; advances one row per load
    const R* rc = rhs + c0 + (size_t)j0 * nx;       // row j0
    R* oc = pn + c0 + (size_t)j0 * nx;
    R* const o_bottom = pn + c0;
    R* const o_top = pn + c0 + (size_t)(ny - 1) * nx;
    RowRegs<R> ring[5];
    V q[5];
    // slots: row j-1 -> (s+0), j -> (s+1), j+1 -> (s+2), j+2 -> (s+3), incoming j+3 -> (s+4)
    ring[0] = load_row<R>(pc, off_l, off_r); pc += nx;
    ring[1] = load_row<R>(pc, off_l, off_r); pc += nx;
    ring[2] = load_row<R>(pc, off_l, off_r); pc += nx;
    ring[3] = load_row<R>(pc, off_l, off_r); pc += nx;
    q[1] = __ldg(reinterpret_cast<const V*>(rc)); rc += nx;
    q[2] = __ldg(reinterpret_cast<const V*>(rc)); rc += nx;
    for (int jb = j0; jb < j1; jb += 5) {
#pragma unroll
      for (int s = 0; s < 5; ++s) {
        const int j = jb + s;
        if (j < j1) {
          const RowRegs<R>& bot = ring[s % 5];
          const RowRegs<R>& cen = ring[(s + 1) % 5];
          const RowRegs<R>& top = ring[(s + 2) % 5];
          // prefetch: p' row j+3 (the top row of row j+2) and rhs row j+2
          ring[(s + 4) % 5] = load_row<R>(pc, off_l, off_r); pc += nx;
          q[(s + 3) % 5] = __ldg(reinterpret_cast<const V*>(rc)); rc += nx;
          const V rr = q[(s + 1) % 5];
          R n0 = jacobi_cell<R>(c, cen.l, cen.y, top.x, bot.x, cen.x, rr.x);
          R n1 = jacobi_cell<R>(c, cen.x, cen.r, top.y, bot.y, cen.y, rr.y);
          if (ghost_l) n0 = n1;                  // p'[0,j] <- p'[1,j]                       (:813)
          if (ghost_r) n1 = c.cavity ? n0 : R(0);  // outlet p'[nx-1,j] <- 0 (:814); cavity: mirror p'[nx-2,j]
          if (cnt0) max_err = fmax(max_err, (double)r_abs<R>(n0 - cen.x));
          if (cnt1) max_err = fmax(max_err, (double)r_abs<R>(n1 - cen.y));
          V out;
          out.x = n0;
          out.y = n1;
          *reinterpret_cast<V*>(oc) = out;
          if (j == 1) *reinterpret_cast<V*>(o_bottom) = out;        // bottom row <- row 1     (:808)
          if (j == ny - 2) *reinterpret_cast<V*>(o_top) = out;      // top row <- row ny-2     (:809)
          oc += nx;
        }
      }
    }
  }
  block_atomic_max<4>(max_err, err_slots + sweep, s_red);
}

// k_jacobi_sweep3 — row-by-row cp.async.bulk kernel (CFD_FLAG_BULK_SWEEP)
constexpr int kSweepStages = 8;   // rows in flight per warp
template <class R>
struct SweepRing {
  static constexpr int kHalo = 16 / (int)sizeof(R);                     // elements: 2 (fp64) or 4 (fp32)
  static constexpr int kPRowBytes = (kStripCols + 2 * kHalo) * (int)sizeof(R);
  static constexpr int kQRowBytes = kStripCols * (int)sizeof(R);
  alignas(128) R prow[kSweepWarps][kSweepStages][kStripCols + 2 * kHalo];
  alignas(128) R qrow[kSweepWarps][kSweepStages][kStripCols];
  alignas(8) unsigned long long bar[kSweepWarps][kSweepStages];
};

template <class R>
__global__ void __launch_bounds__(kSweepWarps * 32) k_jacobi_sweep3(JacobiConsts2<R> c, const R* __restrict__ p,
                                                                   const R* __restrict__ rhs, R* __restrict__ pn,
                                                                   unsigned long long* __restrict__ err_slots,
                                                                   int sweep) {
  using V = typename Vec2<R>::type;
  using Ring = SweepRing<R>;
  constexpr int H = Ring::kHalo;
  __shared__ Ring ring;
  __shared__ double s_red[kSweepWarps];
  if (sweep > 0) {
    const R prev = (R)bits_nonneg(err_slots[sweep - 1]);
    if (prev < c.tol) return;
  }
  const int nx = c.nx, ny = c.ny;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int cw = (blockIdx.x * kSweepWarps + warp) * kStripCols;  // first column of this warp's strip
  const int c0 = cw + 2 * lane;
  const int j0 = c.row_begin + blockIdx.y * c.rows_per_block;
  const int j1 = min(j0 + c.rows_per_block, c.row_end);  // rows [j0, j1)
  double max_err = 0.0;
  if (cw < nx && j0 < j1) {
    const int total = (j1 - j0) + 2;  // staged rows: j0-1 .. j1
    const unsigned bar0 = tma::smem_addr(&ring.bar[warp][0]);
    const unsigned prow0 = tma::smem_addr(&ring.prow[warp][0][0]);
    const unsigned qrow0 = tma::smem_addr(&ring.qrow[warp][0][0]);
    const R* psrc = p + (size_t)(j0 - 1) * nx + cw - H;   // 16-byte aligned: cw % 64 == 0, nx % 8 == 0
    const R* qsrc = rhs + (size_t)(j0 - 1) * nx + cw;
    if (lane == 0) {
#pragma unroll
      for (int k = 0; k < kSweepStages; ++k) tma::mbar_init(bar0 + 8u * k, 1u);
      tma::fence_mbar_init();
    }
    __syncwarp();
    if (lane == 0) {
#pragma unroll
      for (int k = 0; k < kSweepStages; ++k) {
        if (k < total) {
          tma::mbar_expect_tx(bar0 + 8u * k, Ring::kPRowBytes + Ring::kQRowBytes);
          tma::bulk_g2s(prow0 + (unsigned)(k * Ring::kPRowBytes), psrc + (size_t)k * nx, Ring::kPRowBytes, bar0 + 8u * k);
          tma::bulk_g2s(qrow0 + (unsigned)(k * Ring::kQRowBytes), qsrc + (size_t)k * nx, Ring::kQRowBytes, bar0 + 8u * k);
        }
      }
    }
    const bool active = c0 < nx;
    const bool ghost_l = (c0 == 0), ghost_r = (c0 == nx - 2);
    const bool cnt0 = active && (c0 >= 1) && (c0 <= nx - kLanes), cnt1 = active && (c0 + 1 <= nx - kLanes);
    R* oc = pn + c0 + (size_t)j0 * nx;
    R* const o_bottom = pn + c0;
    R* const o_top = pn + c0 + (size_t)(ny - 1) * nx;
    const R* my_p = &ring.prow[warp][0][H + 2 * lane];
    const R* my_q = &ring.qrow[warp][0][2 * lane];
    constexpr int kPStride = kStripCols + 2 * H, kQStride = kStripCols;

    // consume staged row k: centre pair (+ neighbours and rhs when wanted), then hand the stage back to TMA
    auto consume = [&](int k, RowRegs<R>& row, V& q, bool want_lr_q) {
      const int st = k & (kSweepStages - 1);
      tma::mbar_wait(bar0 + 8u * st, (unsigned)(k / kSweepStages) & 1u);
      const V cpair = *reinterpret_cast<const V*>(my_p + st * kPStride);
      row.x = cpair.x;
      row.y = cpair.y;
      if (want_lr_q) {
        row.l = my_p[st * kPStride - 1];
        row.r = my_p[st * kPStride + 2];
        q = *reinterpret_cast<const V*>(my_q + st * kQStride);
      }
      __syncwarp();
      if (lane == 0 && k + kSweepStages < total) {
        const int kn = k + kSweepStages;
        tma::fence_proxy_async();
        tma::mbar_expect_tx(bar0 + 8u * st, Ring::kPRowBytes + Ring::kQRowBytes);
        tma::bulk_g2s(prow0 + (unsigned)(st * Ring::kPRowBytes), psrc + (size_t)kn * nx, Ring::kPRowBytes, bar0 + 8u * st);
        tma::bulk_g2s(qrow0 + (unsigned)(st * Ring::kQRowBytes), qsrc + (size_t)kn * nx, Ring::kQRowBytes, bar0 + 8u * st);
      }
    };

    RowRegs<R> r3[3];
    V q3[3];
    consume(0, r3[0], q3[0], false);  // row j0-1: only its centre pair is ever used (as `bot`)
    consume(1, r3[1], q3[1], true);   // row j0
    for (int jb = j0; jb < j1; jb += 3) {
#pragma unroll
      for (int s = 0; s < 3; ++s) {
        const int j = jb + s;
        if (j < j1) {
          const RowRegs<R>& bot = r3[s % 3];
          const RowRegs<R>& cen = r3[(s + 1) % 3];
          RowRegs<R>& top = r3[(s + 2) % 3];
          consume(j - j0 + 2, top, q3[(s + 2) % 3], true);  // row j+1 (and its rhs, used next step)
          const V rr = q3[(s + 1) % 3];
          if (active) {
            R n0 = jacobi_cell<R>(c, cen.l, cen.y, top.x, bot.x, cen.x, rr.x);
            R n1 = jacobi_cell<R>(c, cen.x, cen.r, top.y, bot.y, cen.y, rr.y);
            if (ghost_l) n0 = n1;                    // p'[0,j] <- p'[1,j]                     (:813)
            if (ghost_r) n1 = c.cavity ? n0 : R(0);  // outlet p'[nx-1,j] <- 0 (:814); cavity: mirror
            if (cnt0) max_err = fmax(max_err, (double)r_abs<R>(n0 - cen.x));
            if (cnt1) max_err = fmax(max_err, (double)r_abs<R>(n1 - cen.y));
            V out;
            out.x = n0;
            out.y = n1;
            *reinterpret_cast<V*>(oc) = out;
            if (j == 1) *reinterpret_cast<V*>(o_bottom) = out;      // bottom row <- row 1   (:808)
            if (j == ny - 2) *reinterpret_cast<V*>(o_top) = out;    // top row <- row ny-2   (:809)
          }
          oc += nx;
        }
      }
    }
  }
  block_atomic_max<kSweepWarps>(max_err, err_slots + sweep, s_red);
}

// k_jacobi_sweep4 — one-row-per-step tensor-TMA kernel (CFD_FLAG_SWEEP4), predecessor of k_jacobi_sweep5
template <class R>
__global__ void __launch_bounds__(kSweepWarps * 32) k_jacobi_sweep4(JacobiConsts2<R> c,
                                                                   const __grid_constant__ CUtensorMap map_p,
                                                                   const __grid_constant__ CUtensorMap map_rhs,
                                                                   R* __restrict__ pn,
                                                                   unsigned long long* __restrict__ err_slots,
                                                                   int sweep) {
  using V = typename Vec2<R>::type;
  using Ring = SweepChunkRing<R>;
  constexpr int H = Ring::kHalo;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  Ring& ring = *reinterpret_cast<Ring*>(smem_raw);
  __shared__ double s_red[kSweepWarps];
  if (sweep > 0) {
    const R prev = (R)bits_nonneg(err_slots[sweep - 1]);
    if (prev < c.tol) return;
  }
  const int nx = c.nx, ny = c.ny;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int cw = (blockIdx.x * kSweepWarps + warp) * kStripCols;  // first column of this warp's strip
  const int c0 = cw + 2 * lane;
  const int j0 = c.row_begin + blockIdx.y * c.rows_per_block;
  const int j1 = min(j0 + c.rows_per_block, c.row_end);  // rows [j0, j1)
  double max_err = 0.0;
  if (cw < nx && j0 < j1) {
    const int total = (j1 - j0) + 2;                               // staged rows k = 0..total-1 <-> rows j0-1 .. j1
    const int n_chunks = (total + kChunkRows - 1) / kChunkRows;
    const unsigned bar0 = tma::smem_addr(&ring.bar[warp][0]);
    const unsigned prow0 = tma::smem_addr(&ring.prow[warp][0][0][0]);
    const unsigned qrow0 = tma::smem_addr(&ring.qrow[warp][0][0][0]);
    if (lane == 0) {
#pragma unroll
      for (int st = 0; st < kSweepChunkStages; ++st) tma::mbar_init(bar0 + 8u * st, 1u);
      tma::fence_mbar_init();
    }
    __syncwarp();
    if (lane == 0) {
#pragma unroll
      for (int st = 0; st < kSweepChunkStages; ++st) {
        if (st < n_chunks) {
          tma::mbar_expect_tx(bar0 + 8u * st, Ring::kPBytes + Ring::kQBytes);
          tma::tensor_g2s_2d(prow0 + (unsigned)(st * Ring::kPBytes), &map_p, cw - H, j0 - 1 - c.row_shift + st * kChunkRows, bar0 + 8u * st);
          tma::tensor_g2s_2d(qrow0 + (unsigned)(st * Ring::kQBytes), &map_rhs, cw, j0 - 1 - c.row_shift + st * kChunkRows, bar0 + 8u * st);
        }
      }
    }
    const bool active = c0 < nx;
    const bool ghost_l = (c0 == 0), ghost_r = (c0 == nx - 2);
    const bool cnt0 = active && (c0 >= 1) && (c0 <= nx - kLanes), cnt1 = active && (c0 + 1 <= nx - kLanes);
    R* oc = pn + c0 + (size_t)j0 * nx;
    R* const o_bottom = pn + c0;
    R* const o_top = pn + c0 + (size_t)(ny - 1) * nx;
    const R* my_p = &ring.prow[warp][0][0][H + 2 * lane];
    const R* my_q = &ring.qrow[warp][0][0][2 * lane];
    RowRegs<R> r3[3];
    V q3[3];
    unsigned parity = 0;
    int j = j0 - 2;  // row computed when staged row k arrives: j0 + k - 2
    for (int kb = 0; kb < total; kb += kSweepChunkStages * kChunkRows) {
#pragma unroll
      for (int st = 0; st < kSweepChunkStages; ++st) {
        const int chunk = kb / kChunkRows + st;
        if (chunk < n_chunks) {
          tma::mbar_wait(bar0 + 8u * st, parity);
#pragma unroll
          for (int r = 0; r < kChunkRows; ++r) {
            constexpr int kDummy = 0;
            (void)kDummy;
            const int slot = (st * kChunkRows + r) % 3;          // == k % 3 (kb is a multiple of 12)
            const int k = kb + st * kChunkRows + r;
            if (k < total) {
              const R* sp = my_p + (st * kChunkRows + r) * Ring::kPCols;
              const V cpair = *reinterpret_cast<const V*>(sp);
              r3[slot].x = cpair.x;
              r3[slot].y = cpair.y;
              r3[slot].l = sp[-1];
              r3[slot].r = sp[2];
              q3[slot] = *reinterpret_cast<const V*>(my_q + (st * kChunkRows + r) * kStripCols);
              if (k >= 2) {
                const RowRegs<R>& bot = r3[(slot + 1) % 3];   // k-2
                const RowRegs<R>& cen = r3[(slot + 2) % 3];   // k-1
                const RowRegs<R>& top = r3[slot];
                const V rr = q3[(slot + 2) % 3];
                if (active) {
                  R n0 = jacobi_cell<R>(c, cen.l, cen.y, top.x, bot.x, cen.x, rr.x);
                  R n1 = jacobi_cell<R>(c, cen.x, cen.r, top.y, bot.y, cen.y, rr.y);
                  if (ghost_l) n0 = n1;                    // p'[0,j] <- p'[1,j]                     (:813)
                  if (ghost_r) n1 = c.cavity ? n0 : R(0);  // outlet p'[nx-1,j] <- 0 (:814); cavity: mirror
                  if (cnt0) max_err = fmax(max_err, (double)r_abs<R>(n0 - cen.x));
                  if (cnt1) max_err = fmax(max_err, (double)r_abs<R>(n1 - cen.y));
                  V out;
                  out.x = n0;
                  out.y = n1;
                  *reinterpret_cast<V*>(oc) = out;
                  if (j == 1) *reinterpret_cast<V*>(o_bottom) = out;      // bottom row <- row 1   (:808)
                  if (j == ny - 2) *reinterpret_cast<V*>(o_top) = out;    // top row <- row ny-2   (:809)
                }
                oc += nx;
              }
              ++j;
            }
          }
          // every lane has copied its values out of the stage: hand it back to the TMA unit
          __syncwarp();
          if (lane == 0 && chunk + kSweepChunkStages < n_chunks) {
            const int row = j0 - 1 - c.row_shift + (chunk + kSweepChunkStages) * kChunkRows;
            tma::fence_proxy_async();
            tma::mbar_expect_tx(bar0 + 8u * st, Ring::kPBytes + Ring::kQBytes);
            tma::tensor_g2s_2d(prow0 + (unsigned)(st * Ring::kPBytes), &map_p, cw - H, row, bar0 + 8u * st);
            tma::tensor_g2s_2d(qrow0 + (unsigned)(st * Ring::kQBytes), &map_rhs, cw, row, bar0 + 8u * st);
          }
        }
      }
      parity ^= 1u;
    }
  }
  block_atomic_max<kSweepWarps>(max_err, err_slots + sweep, s_red);
}

// ---------------------------------------------------------------------------------------------------
// k_jacobi_sweep6 — PERSISTENT version of k_jacobi_sweep5 (opt-in, CFD_FLAG_PERSISTENT_SWEEP; measured 87.5 us per
// sweep at 4096^2 against 78 us for sweep5: 128 registers with spills and a 5.05-units-per-warp tail).  The source-level profile of sweep5
// (profiles/r1_notes.md item 7) puts a quarter of all stall samples at tile boundaries: the wait for a tile's
// first TMA boxes (a full DRAM latency per 22-row tile and warp) and the drain at block exit.  Here the grid is
// one wave of resident blocks and every WARP pulls (tile, strip) units from an atomic counter; while it finishes
// a unit, the stages that fall free are re-armed with the first boxes of its NEXT unit, so the TMA ring never
// drains, and the dynamic hand-out keeps the load balance of the short tiles.  Arithmetic, boundary handling,
// strips protocol (edge units first, peer stores, mailbox) are those of k_jacobi_sweep5.
// Every unit stages exactly kUnitChunks boxes (the ring's phase bookkeeping then runs across units).
// ---------------------------------------------------------------------------------------------------
constexpr int kUnitChunks = 2 * kSweepChunkStages;                 // 6 boxes = 24 staged rows
constexpr int kUnitRows = kUnitChunks * kChunkRows - 2;            // 22 rows updated per unit

template <class R>
__global__ void __launch_bounds__(kSweepWarps * 32, kSweepBlocksPerSm) k_jacobi_sweep6(JacobiConsts2<R> c,
                                                                      const __grid_constant__ CUtensorMap map_p,
                                                                      const __grid_constant__ CUtensorMap map_rhs,
                                                                      R* __restrict__ pn,
                                                                      unsigned long long* __restrict__ err_slots,
                                                                      int sweep, const SweepPeer<R> peer,
                                                                      unsigned int* __restrict__ work_counter) {
  using V = typename Vec2<R>::type;
  using Ring = SweepChunkRing<R>;
  constexpr int H = Ring::kHalo;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  Ring& ring = *reinterpret_cast<Ring*>(smem_raw);
  const bool peers = peer.world > 1;
  const unsigned long long stamp = peer.stamp_base + (unsigned long long)sweep + 1ull;
  if (peers) {
    if (peer.tickets[768 + sweep] != 0u) {  // stop flag, written by an earlier launch
      if (blockIdx.x == 0 && threadIdx.x == 0) {
        peer_publish_max<R>(peer, stamp, sweep, 0ull);
        peer.tickets[768 + sweep + 1] = 1u;
      }
      return;
    }
  } else if (c.fix_pass >= 0) {
    const int s0 = 2 * c.fix_pass;
    const bool ran = s0 == 0 || ((R)bits_nonneg(err_slots[s0 - 1]) >= c.tol && (R)bits_nonneg(err_slots[s0 - 2]) >= c.tol);
    if (!ran || !((R)bits_nonneg(err_slots[s0]) < c.tol)) return;
  } else if (sweep >= 1) {
    if ((R)bits_nonneg(err_slots[sweep - 1]) < c.tol) return;
  }
  const int nx = c.nx, ny = c.ny;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_strips = (nx + kStripCols - 1) / kStripCols;
  const int n_tiles = (c.row_end - c.row_begin + kUnitRows - 1) / kUnitRows;
  const unsigned n_units = (unsigned)(n_strips * n_tiles);
  const unsigned bar0 = tma::smem_addr(&ring.bar[warp][0]);
  const unsigned prow0 = tma::smem_addr(&ring.prow[warp][0][0][0]);
  const unsigned qrow0 = tma::smem_addr(&ring.qrow[warp][0][0][0]);
  if (lane == 0) {
#pragma unroll
    for (int st = 0; st < kSweepChunkStages; ++st) tma::mbar_init(bar0 + 8u * st, 1u);
    tma::fence_mbar_init();
  }
  __syncwarp();

  auto fetch = [&]() -> unsigned {
    unsigned u = 0;
    if (lane == 0) u = atomicAdd(work_counter, 1u);
    return __shfl_sync(0xffffffffu, u, 0);
  };
  // unit -> tile (strips: both edge tiles are handed out first) and strip
  auto unit_tile = [&](unsigned u) -> int {
    int t = (int)(u / (unsigned)n_strips);
    if (peers && n_tiles > 2) t = t == 0 ? 0 : (t == 1 ? n_tiles - 1 : t - 1);
    return t;
  };
  auto unit_is_edge = [&](unsigned u) -> bool {
    if (!peers) return false;
    const int t = unit_tile(u);
    return (peer.down_out != nullptr && t == 0) || (peer.up_out != nullptr && t == n_tiles - 1);
  };
  // lane 0: arm stage `st` and fetch box `chunk` of unit u into it
  auto issue_box = [&](unsigned u, int chunk, int st) {
    const int t = unit_tile(u);
    const int cw = (int)(u % (unsigned)n_strips) * kStripCols;
    const int row = c.row_begin + t * kUnitRows - 1 - c.row_shift + chunk * kChunkRows;
    tma::mbar_expect_tx(bar0 + 8u * st, Ring::kPBytes + Ring::kQBytes);
    tma::tensor_g2s_2d(prow0 + (unsigned)(st * Ring::kPBytes), &map_p, cw - H, row, bar0 + 8u * st);
    tma::tensor_g2s_2d(qrow0 + (unsigned)(st * Ring::kQBytes), &map_rhs, cw, row, bar0 + 8u * st);
  };

  R max_err = R(0);
  double dot_dummy = 0.0;
  unsigned cur = fetch(), nxt = fetch();
  bool cur_issued = false;
  unsigned parity = 0;
  while (cur < n_units) {
    const int tile = unit_tile(cur);
    const int cw = (int)(cur % (unsigned)n_strips) * kStripCols;
    const int j0 = c.row_begin + tile * kUnitRows;
    const int j1 = min(j0 + kUnitRows, c.row_end);  // rows [j0, j1)
    const int total = (j1 - j0) + 2;
    const bool edge_lo = peers && peer.down_out != nullptr && j0 == c.row_begin;
    const bool edge_hi = peers && peer.up_out != nullptr && j1 == c.row_end;
    if (!cur_issued) {
      if (sweep > 0 && (edge_lo || edge_hi)) {  // the neighbours' previous sweep must have filled the halo rows
        if (lane == 0) {
          if (edge_lo) while (ld_acquire_sys(&peer.mine->halo_flag[0]) < stamp - 1ull) __nanosleep(64);
          if (edge_hi) while (ld_acquire_sys(&peer.mine->halo_flag[1]) < stamp - 1ull) __nanosleep(64);
          asm volatile("fence.proxy.async;" ::: "memory");
        }
        __syncwarp();
      }
      if (lane == 0) {
#pragma unroll
        for (int st = 0; st < kSweepChunkStages; ++st) issue_box(cur, st, st);
      }
    }
    const bool prefetch_next = nxt < n_units && !unit_is_edge(nxt);
    const int lane_eff = min(lane, (nx - 2 - cw) >> 1);
    const int c0 = cw + 2 * lane_eff;
    const bool active = lane_eff == lane;
    const bool ghost_l = (c0 == 0), ghost_r = (c0 == nx - 2), ghost = ghost_l | ghost_r;
    const bool cnt0 = active && (c0 >= 1) && (c0 <= nx - kLanes), cnt1 = active && (c0 + 1 <= nx - kLanes);
    R* oc = pn + c0 + (size_t)j0 * nx;
    R* const o_bottom = edge_lo ? peer.down_out + c0 + (size_t)j0 * nx : pn + c0;
    R* const o_top = edge_hi ? peer.up_out + c0 + (size_t)(j1 - 1) * nx : pn + c0 + (size_t)(ny - 1) * nx;
    const int m_bottom = (j0 == 1 || edge_lo) ? 1 : -1;
    const int m_top = (j1 == ny - 1 || edge_hi) ? total - 2 : -1;
    const R* my_p = &ring.prow[warp][0][0][H + 2 * lane_eff];
    const R* my_q = &ring.qrow[warp][0][0][2 * lane_eff];
    const size_t two_rows = 2 * (size_t)nx;
    RowRegs<R> r4[4];
    V q4[4];
#pragma unroll 1
    for (int body = 0; body < kUnitChunks / kSweepChunkStages; ++body) {
      const int kb = body * kSweepChunkStages * kChunkRows;
#pragma unroll
      for (int st = 0; st < kSweepChunkStages; ++st) {
        const int chunk = body * kSweepChunkStages + st;
        tma::mbar_wait(bar0 + 8u * st, parity);
#pragma unroll
        for (int h = 0; h < kChunkRows / 2; ++h) {
          const int sa = (st * kChunkRows + 2 * h) % 4, sb = sa + 1;
          const int k = kb + st * kChunkRows + 2 * h;  // even
          {
            const R* sp = my_p + (st * kChunkRows + 2 * h) * Ring::kPCols;
            const V ca = *reinterpret_cast<const V*>(sp);
            const V cb = *reinterpret_cast<const V*>(sp + Ring::kPCols);
            r4[sa].x = ca.x; r4[sa].y = ca.y; r4[sa].l = sp[-1]; r4[sa].r = sp[2];
            r4[sb].x = cb.x; r4[sb].y = cb.y; r4[sb].l = sp[Ring::kPCols - 1]; r4[sb].r = sp[Ring::kPCols + 2];
            const R* sq = my_q + (st * kChunkRows + 2 * h) * kStripCols;
            q4[sa] = *reinterpret_cast<const V*>(sq);
            q4[sb] = *reinterpret_cast<const V*>(sq + kStripCols);
          }
          const RowRegs<R>& rm2 = r4[(sa + 2) % 4];
          const RowRegs<R>& rm1 = r4[(sa + 3) % 4];
          if (k >= 2) {
            if (k + 1 < total) {
              sweep_row<R>(c, rm2, rm1, r4[sa], q4[(sa + 3) % 4], ghost, ghost_l, ghost_r, cnt0, cnt1, active, oc,
                           k - 1 == m_bottom, o_bottom, k - 1 == m_top, o_top, max_err, dot_dummy);
              sweep_row<R>(c, rm1, r4[sa], r4[sb], q4[sa], ghost, ghost_l, ghost_r, cnt0, cnt1, active, oc + nx,
                           false, o_bottom, k == m_top, o_top, max_err, dot_dummy);
            } else if (k < total) {
              sweep_row<R>(c, rm2, rm1, r4[sa], q4[(sa + 3) % 4], ghost, ghost_l, ghost_r, cnt0, cnt1, active, oc,
                           k - 1 == m_bottom, o_bottom, k - 1 == m_top, o_top, max_err, dot_dummy);
            }
            oc += two_rows;
          }
        }
        // the stage is free again: next box of this unit, or — when this unit has none left — of the next unit
        __syncwarp();
        if (lane == 0) {
          tma::fence_proxy_async();
          if (chunk + kSweepChunkStages < kUnitChunks) issue_box(cur, chunk + kSweepChunkStages, st);
          else if (prefetch_next) issue_box(nxt, chunk + kSweepChunkStages - kUnitChunks, st);
        }
      }
      parity ^= 1u;
    }
    if (peers && (edge_lo || edge_hi)) {  // this unit wrote a neighbour's halo row: signal when the whole row is there
      __threadfence_system();
      __syncwarp();
      if (lane == 0) {
        if (edge_lo && atomicAdd(&peer.tickets[256 + sweep], 1u) == (unsigned)n_strips - 1u) {
          __threadfence_system();
          st_release_sys(peer.down_flag, stamp);
        }
        if (edge_hi && atomicAdd(&peer.tickets[512 + sweep], 1u) == (unsigned)n_strips - 1u) {
          __threadfence_system();
          st_release_sys(peer.up_flag, stamp);
        }
      }
    }
    cur_issued = prefetch_next;
    cur = nxt;
    nxt = fetch();
  }
  // one atomicMax per warp; strips: the last warp of the launch publishes the local max and the next stop flag
  double m = warp_max((double)max_err);
  if (lane == 0) {
    if (m > 0.0) atomicMax(err_slots + (c.fix_pass >= 0 ? 255 : sweep), nonneg_bits(m));
    if (peers) {
      __threadfence();
      if (atomicAdd(&peer.tickets[sweep], 1u) == gridDim.x * kSweepWarps - 1u) {
        __threadfence();
        const unsigned long long local = atomicMax(err_slots + sweep, 0ull);
        peer_publish_max<R>(peer, stamp, sweep, local);
        if (sweep >= 1 && (R)peer_global_max(peer.mine, peer.world, stamp - 1ull, sweep - 1) < c.tol)
          peer.tickets[768 + sweep + 1] = 1u;
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// k_jacobi_sweep_t2 — TEMPORAL BLOCKING: two damped-Jacobi sweeps (s, s+1) per pass over HBM.
//
// One sweep is bound by HBM at 3*s*N bytes (profiles/r1_notes.md: 0.79 of the measured peak) while the fp64 pipe
// idles at ~25 %.  This kernel reads p' and rhs once, keeps the intermediate field p'(s) in registers, and writes
// p'(s+1): 3*s*N bytes for TWO sweeps.  Per warp (64-column strip, rows streamed through the same tensor-TMA ring
// as k_jacobi_sweep4): level-0 rows k-2..k in a 3-slot register ring -> level-1 row k-1 (sweep s) -> level-2 row k-2
// (sweep s+1) from level-1 rows k-3..k-1.  Horizontal level-1 neighbours come from the adjacent lanes (shuffles);
// the two strip-edge lanes compute one redundant level-1 cell each (columns cw-1 and cw+64, from the 2-column TMA
// halo); tiles overlap by one redundant level-1 row on each side (staged rows = tile rows + 4).
//
// EXACT reference semantics (src/model.rs:748-820):
//  * ghost cells after sweep s are images of new interior values (:807-815), so level-1 ghost rows / columns are
//    formed exactly like the stored ones (row 0 <- row 1, row ny-1 <- row ny-2, column 0 <- column 1, column nx-1 <-
//    0 or mirror) before level 2 consumes them;
//  * max|dp'| of BOTH sweeps is reduced, over the cells this tile owns only, into err_slots[s] and err_slots[s+1];
//  * the reference stops after the first sweep whose max is below the tolerance.  If that is sweep s+1, or a later
//    one, nothing special happens (later passes see it and return).  If it is sweep s — the first of the pair — the
//    pass has gone one sweep too far: a fix-up launch of k_jacobi_sweep5 (fix_pass) recomputes sweep s alone from the
//    pass's INPUT buffer, which is still intact, into the same output buffer.  Fields, counters and residuals stay
//    bit-identical to the one-sweep-per-launch path (tests/test_gpu_parity.py).
// Algorithmic bytes per launch: 2 x 3*s*N (two sweeps' worth); DRAM traffic ~1.6*s*N per sweep.
// MEASURED (profiles/r1_notes.md): 233 us per pass at 4096^2 = 117 us per sweep, SLOWER than k_jacobi_sweep5's
// 78 us: 100 M warp-instructions per pass (3.2x a single sweep, not 2x — the two strip-edge lanes' redundant cell
// costs a full warp instruction stream, plus ring copies and shuffles) at the same ~44 % issue-active.  Both
// kernels are issue/latency-limited, not HBM-limited, so halving the DRAM traffic buys nothing yet.  Opt-in
// (CFD_FLAG_TEMPORAL), kept bit-exact under test as the starting point for an instruction-leaner version.
// ---------------------------------------------------------------------------------------------------
template <class R>
struct SweepT2Ring {
  static constexpr int kHalo = 16 / (int)sizeof(R);
  static constexpr int kPCols = kStripCols + 2 * kHalo;
  static constexpr int kBoxBytes = kPCols * kChunkRows * (int)sizeof(R);
  alignas(128) R prow[kSweepWarps][kSweepChunkStages][kChunkRows][kPCols];
  alignas(128) R qrow[kSweepWarps][kSweepChunkStages][kChunkRows][kPCols];  // rhs, same halo box
  alignas(8) unsigned long long bar[kSweepWarps][kSweepChunkStages];
};

template <class R>
struct Lvl1 {
  R x, y, e;  // level-1 values at columns c0, c0+1 and (edge lanes) the redundant column next to the strip
};

template <class R>
__global__ void __launch_bounds__(kSweepWarps * 32, kSweepBlocksPerSm) k_jacobi_sweep_t2(JacobiConsts2<R> c,
                                                                        const __grid_constant__ CUtensorMap map_p,
                                                                        const __grid_constant__ CUtensorMap map_rhs_halo,
                                                                        R* __restrict__ pn,
                                                                        unsigned long long* __restrict__ err_slots,
                                                                        int sweep) {
  using V = typename Vec2<R>::type;
  using Ring = SweepT2Ring<R>;
  constexpr int H = Ring::kHalo;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  Ring& ring = *reinterpret_cast<Ring*>(smem_raw);
  __shared__ double s_red[kSweepWarps];
  if (sweep >= 2) {  // an earlier pass already contained the stopping sweep (or was itself skipped: slots stay 0)
    if ((R)bits_nonneg(err_slots[sweep - 1]) < c.tol || (R)bits_nonneg(err_slots[sweep - 2]) < c.tol) return;
  }
  const int nx = c.nx, ny = c.ny;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int cw = (blockIdx.x * kSweepWarps + warp) * kStripCols;
  const int j0 = c.row_begin + blockIdx.y * c.rows_per_block;
  const int j1 = min(j0 + c.rows_per_block, c.row_end);  // level-2 rows [j0, j1)
  R err1 = R(0), err2 = R(0);
  if (cw < nx && j0 < j1) {
    const int total = (j1 - j0) + 4;  // staged level-0 rows k = 0..total-1  <->  global rows j0-2 .. j1+1
    const int n_chunks = (total + kChunkRows - 1) / kChunkRows;
    const unsigned bar0 = tma::smem_addr(&ring.bar[warp][0]);
    const unsigned prow0 = tma::smem_addr(&ring.prow[warp][0][0][0]);
    const unsigned qrow0 = tma::smem_addr(&ring.qrow[warp][0][0][0]);
    const int row0 = j0 - 2 - c.row_shift;
    if (lane == 0) {
#pragma unroll
      for (int st = 0; st < kSweepChunkStages; ++st) tma::mbar_init(bar0 + 8u * st, 1u);
      tma::fence_mbar_init();
    }
    __syncwarp();
    if (lane == 0) {
#pragma unroll
      for (int st = 0; st < kSweepChunkStages; ++st) {
        if (st < n_chunks) {
          tma::mbar_expect_tx(bar0 + 8u * st, 2 * Ring::kBoxBytes);
          tma::tensor_g2s_2d(prow0 + (unsigned)(st * Ring::kBoxBytes), &map_p, cw - H, row0 + st * kChunkRows, bar0 + 8u * st);
          tma::tensor_g2s_2d(qrow0 + (unsigned)(st * Ring::kBoxBytes), &map_rhs_halo, cw - H, row0 + st * kChunkRows, bar0 + 8u * st);
        }
      }
    }
    const int lane_eff = min(lane, (nx - 2 - cw) >> 1);
    const int c0 = cw + 2 * lane_eff;
    const bool active = lane_eff == lane;
    const bool ghost_l = (c0 == 0), ghost_r = (c0 == nx - 2), ghost = ghost_l | ghost_r;
    const bool cnt0 = active && (c0 >= 1) && (c0 <= nx - kLanes), cnt1 = active && (c0 + 1 <= nx - kLanes);
    // the redundant level-1 cell of the two strip-edge lanes: column cw-1 (lane 0) / cw+64 (lane 31), if it exists
    const bool extra_l = (lane == 0) && (cw > 0), extra_r = (lane == 31) && (cw + kStripCols < nx);
    const bool has_extra = extra_l | extra_r;
    // shared-memory offsets of that cell's outer horizontal neighbour (cw-2 / cw+65) and of its rhs, relative to
    // this lane's centre pair; lanes without an extra cell read their own centre (harmless)
    const int off_outer = extra_l ? -2 : (extra_r ? 3 : 0);
    const int off_qe = extra_l ? -1 : (extra_r ? 2 : 0);
    R* oc = pn + c0 + (size_t)j0 * nx;
    R* const o_bottom = pn + c0;
    R* const o_top = pn + c0 + (size_t)(ny - 1) * nx;
    const bool tile_bottom = (j0 == 1), tile_top = (j1 == ny - 1);
    const R* my_p = &ring.prow[warp][0][0][H + 2 * lane_eff];
    const R* my_q = &ring.qrow[warp][0][0][H + 2 * lane_eff];
    RowRegs<R> L0[3];   // level-0 rows k-2, k-1, k
    R outer[3];         // level-0 value at the extra cell's outer neighbour column, same rows
    Lvl1<R> E[3];       // level-1 rows k-3, k-2, k-1
    V q[3];             // rhs rows k-2, k-1, k
    R qe[3];            // rhs at the extra cell's column
    unsigned parity = 0;
    for (int kb = 0; kb < total; kb += kSweepChunkStages * kChunkRows) {
#pragma unroll
      for (int st = 0; st < kSweepChunkStages; ++st) {
        const int chunk = kb / kChunkRows + st;
        if (chunk < n_chunks) {
          tma::mbar_wait(bar0 + 8u * st, parity);
#pragma unroll
          for (int r = 0; r < kChunkRows; ++r) {
            const int sl = (st * kChunkRows + r) % 3;  // == k % 3 (kb is a multiple of 12)
            const int k = kb + st * kChunkRows + r;
            if (k < total) {
              {  // level-0 row k and its rhs, from the staged box
                const R* sp = my_p + (st * kChunkRows + r) * Ring::kPCols;
                const R* sq = my_q + (st * kChunkRows + r) * Ring::kPCols;
                const V cp = *reinterpret_cast<const V*>(sp);
                L0[sl].x = cp.x; L0[sl].y = cp.y; L0[sl].l = sp[-1]; L0[sl].r = sp[2];
                outer[sl] = sp[off_outer];
                q[sl] = *reinterpret_cast<const V*>(sq);
                qe[sl] = sq[off_qe];
              }
              const int s_m1 = (sl + 2) % 3, s_m2 = (sl + 1) % 3;  // slots of k-1, k-2 (and of k-4 == k-1, k-5 == k-2 ...)
              // ---- A: level-1 row m = k-1 (global row j0-2+m): sweep s
              const int m = k - 1;
              if (k >= 2) {
                const int gm = j0 - 2 + m;
                if (gm >= 1 && gm <= ny - 2) {
                  const RowRegs<R>& bot = L0[s_m2];
                  const RowRegs<R>& cen = L0[s_m1];
                  const RowRegs<R>& top = L0[sl];
                  const V rr = q[s_m1];
                  R n0 = jacobi_cell<R>(c, cen.l, cen.y, top.x, bot.x, cen.x, rr.x);
                  R n1 = jacobi_cell<R>(c, cen.x, cen.r, top.y, bot.y, cen.y, rr.y);
                  R ne = R(0);
                  if (has_extra) {  // strip-edge lanes only (divergent on purpose: 2 of 32 lanes)
                    if (extra_l) ne = jacobi_cell<R>(c, outer[s_m1], cen.x, top.l, bot.l, cen.l, qe[s_m1]);
                    else ne = jacobi_cell<R>(c, cen.y, outer[s_m1], top.r, bot.r, cen.r, qe[s_m1]);
                  }
                  if (__builtin_expect(ghost, 0)) {
                    const V g = fix_ghost_columns<R>(n0, n1, ghost_l, ghost_r, c.cavity);
                    n0 = g.x;
                    n1 = g.y;
                  }
                  if (m >= 2 && m <= total - 3) {  // rows this tile owns (the two outer level-1 rows are redundant)
                    const R e0 = r_abs<R>(n0 - cen.x), e1 = r_abs<R>(n1 - cen.y);
                    if (cnt0 && e0 > err1) err1 = e0;
                    if (cnt1 && e1 > err1) err1 = e1;
                  }
                  E[s_m1].x = n0; E[s_m1].y = n1; E[s_m1].e = ne;
                  if (gm == 1) E[s_m2] = E[s_m1];  // bottom ghost row of p'(s): row 0 <- row 1 (:808)
                } else if (gm == ny - 1) {
                  E[s_m1] = E[s_m2];               // top ghost row: row ny-1 <- row ny-2 (:809)
                }
              }
              // ---- B: level-2 row n = k-2 (global row j0-2+n = an owned row): sweep s+1
              if (k >= 4) {  // n >= 2; n <= total-3 holds because k <= total-1
                const Lvl1<R>& bot = E[sl];      // n-1 = k-3
                const Lvl1<R>& cen = E[s_m2];    // n   = k-2
                const Lvl1<R>& top = E[s_m1];    // n+1 = k-1
                R left = __shfl_up_sync(0xffffffffu, cen.y, 1);
                R right = __shfl_down_sync(0xffffffffu, cen.x, 1);
                if (lane == 0) left = cen.e;
                if (lane == 31) right = cen.e;
                const V rr = q[s_m2];
                R n0 = jacobi_cell<R>(c, left, cen.y, top.x, bot.x, cen.x, rr.x);
                R n1 = jacobi_cell<R>(c, cen.x, right, top.y, bot.y, cen.y, rr.y);
                if (__builtin_expect(ghost, 0)) {
                  const V g = fix_ghost_columns<R>(n0, n1, ghost_l, ghost_r, c.cavity);
                  n0 = g.x;
                  n1 = g.y;
                }
                const R e0 = r_abs<R>(n0 - cen.x), e1 = r_abs<R>(n1 - cen.y);
                if (cnt0 && e0 > err2) err2 = e0;
                if (cnt1 && e1 > err2) err2 = e1;
                if (active) {
                  V out;
                  out.x = n0;
                  out.y = n1;
                  *reinterpret_cast<V*>(oc) = out;
                  if (tile_bottom && k == 4) *reinterpret_cast<V*>(o_bottom) = out;         // row 0 <- row 1
                  if (tile_top && k == total - 1) *reinterpret_cast<V*>(o_top) = out;       // row ny-1 <- row ny-2
                }
                oc += nx;
              }
            }
          }
          __syncwarp();
          if (lane == 0 && chunk + kSweepChunkStages < n_chunks) {
            const int row = row0 + (chunk + kSweepChunkStages) * kChunkRows;
            tma::fence_proxy_async();
            tma::mbar_expect_tx(bar0 + 8u * st, 2 * Ring::kBoxBytes);
            tma::tensor_g2s_2d(prow0 + (unsigned)(st * Ring::kBoxBytes), &map_p, cw - H, row, bar0 + 8u * st);
            tma::tensor_g2s_2d(qrow0 + (unsigned)(st * Ring::kBoxBytes), &map_rhs_halo, cw - H, row, bar0 + 8u * st);
          }
        }
      }
      parity ^= 1u;
    }
  }
  block_atomic_max<kSweepWarps>((double)err1, err_slots + sweep, s_red);
  __syncthreads();
  block_atomic_max<kSweepWarps>((double)err2, err_slots + sweep + 1, s_red);
}


}  // namespace cfdk
