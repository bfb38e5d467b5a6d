// cfd_peer.cuh — row strips over NVLink peer memory: halo rows, gathers and scalar reductions without NCCL.
//
// One process per GPU; every rank maps its neighbours' field buffers and every rank's mailbox with CUDA IPC
// (the 64-byte handles travel through one NCCL all-gather at set-up).  After that an exchange is ONE small launch:
//   * k_peer_push copies this rank's edge rows straight into the neighbours' halo rows (NVLink stores), its last block
//     raises a sequence number in each neighbour's mailbox (threadfence.system + st.release.sys) and then waits until
//     the neighbours' own pushes of the same sequence number have arrived (ld.acquire.sys on the local mailbox);
//   * k_peer_reduce publishes up to four scalars to every rank's mailbox and combines all ranks' contributions in rank
//     order (sum of doubles or max of non-negative bit patterns): deterministic, identical on every rank;
//   * k_peer_gather_push / k_peer_gather_wait do the same for "every rank needs every rank's rows" (the multigrid level
//     below which the hierarchy runs replicated).
// ~5 us per exchange instead of ~25-30 us for a grouped ncclSend/ncclRecv or an 8-byte ncclAllReduce at 8 ranks, which
// is what the 17 + 1 + 3 NCCL operations per CG iteration of round 1 cost.
//
// Ordering argument.  Every rank issues the same sequence of exchanges (same host control flow).  Sequence numbers
// only grow, so stale mailbox contents are harmless.  Exchange E's push is stream-ordered after this rank's wait of
// exchange E-1, i.e. after both neighbours finished everything they had enqueued before their push E-1 — in particular
// every kernel that read the halo rows which push E overwrites, because every stencil kernel of the solver writes to
// a buffer other than the one it reads (ping-pong), so a halo is never re-read after the exchange that follows its last
// reader.  The waits spin for at most kPeerTimeoutNs; a rank whose neighbour died raises `error` in its own mailbox
// instead of hanging, and the host turns that into CFD_ERR_PEER_TIMEOUT at the end of the step.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace cfdk {

constexpr int kPeerMaxRanks = 8;
constexpr int kPeerRedSlots = 4;
constexpr unsigned long long kPeerTimeoutNs = 20ull * 1000ull * 1000ull * 1000ull;

struct PeerRed {
  unsigned long long v[4];  // up to four scalars (doubles as bit patterns)
  unsigned long long seq;
  unsigned long long pad[3];
};

struct PeerBox {
  unsigned long long from_below;  // sequence number of the last halo push that arrived from the rank below
  unsigned long long from_above;
  unsigned long long error;       // != 0: a wait timed out at this sequence number
  unsigned long long pad0[5];
  unsigned long long gather[kPeerMaxRanks];  // [source rank]: sequence number of its last gather push
  PeerRed red[kPeerRedSlots][kPeerMaxRanks]; // [seq % slots][source rank]
};

__device__ __forceinline__ unsigned long long peer_ld_acquire(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void peer_st_release(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ void peer_st_relaxed(unsigned long long* p, unsigned long long v) {
  asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long peer_now_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// waits until *flag >= seq; false on time-out (the caller records it in the local mailbox)
__device__ __forceinline__ bool peer_wait_geq(const unsigned long long* flag, unsigned long long seq) {
  if (peer_ld_acquire(flag) >= seq) return true;
  const unsigned long long t0 = peer_now_ns();
  for (;;) {
    __nanosleep(100);
    if (peer_ld_acquire(flag) >= seq) return true;
    if (peer_now_ns() - t0 > kPeerTimeoutNs) return false;
  }
}

// ---- scalar reductions over all ranks ------------------------------------------------------------------------------------
struct PeerAll {
  PeerBox* box[kPeerMaxRanks];  // every rank's mailbox (box[rank] is the local one)
  int rank, world;
};

// data[0..n) (n <= 4) <- reduction over the ranks of data[0..n): op 0 = max of unsigned 64-bit patterns (non-negative doubles
// order like their bit patterns), op 1 = sum of doubles in rank order.  One thread.  `skip` (may be null): a
// device flag; when it is up the whole operation is a no-op ON EVERY RANK (the flag is itself a reduced quantity).
__device__ __forceinline__ void peer_reduce_values(const PeerAll& p, unsigned long long* data, int n, int op,
                                                   unsigned long long seq) {
  const int slot = (int)(seq % kPeerRedSlots);
  for (int r = 0; r < p.world; ++r) {
    PeerRed* rec = &p.box[r]->red[slot][p.rank];
    for (int k = 0; k < n; ++k) peer_st_relaxed(&rec->v[k], data[k]);
    peer_st_release(&rec->seq, seq);
  }
  PeerBox* mine = p.box[p.rank];
  unsigned long long acc_max[4] = {0ull, 0ull, 0ull, 0ull};
  double acc_sum[4] = {0.0, 0.0, 0.0, 0.0};
  bool ok = true;
  for (int r = 0; r < p.world; ++r) {
    const PeerRed* rec = &mine->red[slot][r];
    ok &= peer_wait_geq(&rec->seq, seq);
    for (int k = 0; k < n; ++k) {
      const unsigned long long v = peer_ld_acquire(&rec->v[k]);
      if (op == 0) acc_max[k] = v > acc_max[k] ? v : acc_max[k];
      else acc_sum[k] += __longlong_as_double((long long)v);
    }
  }
  for (int k = 0; k < n; ++k) data[k] = op == 0 ? acc_max[k] : (unsigned long long)__double_as_longlong(acc_sum[k]);
  if (!ok) mine->error = seq;
}

__global__ void k_peer_reduce(const PeerAll p, unsigned long long* __restrict__ data, int n, int op, unsigned long long seq,
                              const int* __restrict__ skip) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  if (skip != nullptr && *skip) return;
  peer_reduce_values(p, data, n, op, seq);
}

// MGCG on strips: the ranks' parts of a dot product (MgScalars::local_sum) summed over the ranks AND the CG scalars
// advanced (mg_advance) in one single-thread launch — every rank ends up with identical scalars, hence the same `done`
template <class R>
__global__ void k_mg_advance_peer(const PeerAll p, const MgFine<R> c, MgScalars* __restrict__ sc, int mode, unsigned long long seq) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  // (no early exit on sc->done: every rank must take part in every reduction it was enqueued for)
  unsigned long long v = (unsigned long long)__double_as_longlong(sc->local_sum);
  peer_reduce_values(p, &v, 1, 1, seq);
  sc->local_sum = __longlong_as_double((long long)v);
  if (mode != 0 && mode != 4 && mode != 5 && sc->done) return;
  mg_advance<R>(c, sc, sc->local_sum, mode);
}

// ---- halo rows: push to the neighbours, then wait for theirs --------------------------------------------------------------
struct PeerPush {
  const uint32_t* src[2];        // [0] towards the rank below, [1] towards the rank above (4-byte words)
  uint32_t* dst[2];              // the neighbours' halo rows (peer pointers)
  unsigned long long words[2];   // may be 0 with a neighbour present: the flag is still raised (a pure synchronisation)
  unsigned long long* flag[2];   // the neighbours' mailboxes: [0] the lower rank's from_above, [1] the upper rank's from_below;
                                 // nullptr = no neighbour on that side
  PeerBox* mine;
  unsigned int* ticket;
  unsigned long long seq;
  // optional: a max-reduction over ALL ranks riding on the same launch (Mode R: the sweep's max|dp'| next to its halo rows)
  unsigned long long* red;       // nullptr: none
  int red_n;
  unsigned long long red_seq;
  PeerAll all;
};

__global__ void __launch_bounds__(256) k_peer_push(const PeerPush a) {
  const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
  const unsigned long long t = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
#pragma unroll
  for (int k = 0; k < 2; ++k)
    if (a.flag[k] != nullptr)
      for (unsigned long long w = t; w < a.words[k]; w += stride) a.dst[k][w] = a.src[k][w];
  // the block's peer stores happen-before thread 0's fence through the barrier (fences are cumulative), so one
  // system-scope fence per block suffices; the last block's fence then covers every block through the ticket
  __syncthreads();
  if (threadIdx.x != 0) return;
  __threadfence_system();
  if (atomicAdd(a.ticket, 1u) != gridDim.x - 1) return;
  *a.ticket = 0u;
  __threadfence_system();
  if (a.flag[0] != nullptr) peer_st_release(a.flag[0], a.seq);
  if (a.flag[1] != nullptr) peer_st_release(a.flag[1], a.seq);
  if (a.red != nullptr) peer_reduce_values(a.all, a.red, a.red_n, 0, a.red_seq);
  bool ok = true;
  if (a.flag[0] != nullptr) ok &= peer_wait_geq(&a.mine->from_below, a.seq);
  if (a.flag[1] != nullptr) ok &= peer_wait_geq(&a.mine->from_above, a.seq);
  if (!ok) a.mine->error = a.seq;
}

// ---- gather: every rank's rows of a replicated array to every rank ---------------------------------------------------------
struct PeerGather {
  const uint32_t* src;             // this rank's rows (local array)
  uint32_t* dst[kPeerMaxRanks];    // the same rows inside every OTHER rank's copy of the array (nullptr for this rank)
  unsigned long long words;
  PeerAll all;
  unsigned int* ticket;
  unsigned long long seq;
};

__global__ void __launch_bounds__(256) k_peer_gather(const PeerGather a, const int* __restrict__ skip) {
  if (skip != nullptr && *skip) return;
  const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
  const unsigned long long t = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  for (unsigned long long w = t; w < a.words; w += stride) {
    const uint32_t v = a.src[w];
    for (int r = 0; r < a.all.world; ++r)
      if (a.dst[r] != nullptr) a.dst[r][w] = v;
  }
  __syncthreads();
  if (threadIdx.x != 0) return;
  __threadfence_system();
  if (atomicAdd(a.ticket, 1u) != gridDim.x - 1) return;
  *a.ticket = 0u;
  __threadfence_system();
  for (int r = 0; r < a.all.world; ++r)
    if (r != a.all.rank) peer_st_release(&a.all.box[r]->gather[a.all.rank], a.seq);
  PeerBox* mine = a.all.box[a.all.rank];
  bool ok = true;
  for (int r = 0; r < a.all.world; ++r)
    if (r != a.all.rank) ok &= peer_wait_geq(&mine->gather[r], a.seq);
  if (!ok) mine->error = a.seq;
}

}  // namespace cfdk
