// cfd_model.cu — host orchestration of one timestep + the C ABI of include/cfd_b200.h.
//
// Replaces the inside of the reference's `Model` (src/model.rs:166-379, 529-730, 1250-1280): the fields
// live in HBM as structure-of-arrays, every stage is a kernel from cfd_kernels.cuh, the full-field copies
// of the reference (u_old <- u :307-308, u_star <- u :698-699) are buffer rotations, and the only
// device->host traffic per step is a handful of scalars (sweep counts, max-reductions).
// There is no CPU fallback: every entry point fails with CFD_ERR_CUDA if the device is unusable.
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#include <dlfcn.h>
#include <nccl.h>

#include "../../include/cfd_b200.h"
#include "cfd_kernels.cuh"
#include "cfd_mg.cuh"
#include "cfd_mg_legs.cuh"
#ifdef CFD_WITH_AB_SWEEPS
#include "cfd_sweeps_ab.cuh"  // A/B kernels: libcfd_b200_ab.so only
#endif
#include "cfd_peer.cuh"
#include "cfd_tracers.cuh"

namespace {

thread_local std::string g_last_error;

int fail(int code, const std::string& msg) {
  g_last_error = msg;
  return code;
}

#define CFD_CUDA(expr)                                                                              \
  do {                                                                                              \
    cudaError_t _e = (expr);                                                                        \
    if (_e != cudaSuccess)                                                                          \
      return fail(CFD_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e) + " (" + __FILE__ + \
                                    ":" + std::to_string(__LINE__) + ")");                          \
  } while (0)

// NCCL is resolved at run time (dlopen by SONAME) instead of being a link-time dependency: a process that
// also hosts PyTorch must end up with ONE libnccl.so.2 — whichever is loaded first is shared — and a process that
// never goes multi-GPU (the reference UI, the single-GPU tests) never loads NCCL at all.
struct NcclApi {
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  bool ok = false;
};

NcclApi& nccl_api() {
  static NcclApi api;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_LOCAL);
    if (h) {
#define CFD_SYM(field, name) api.field = reinterpret_cast<decltype(api.field)>(dlsym(h, name))
      CFD_SYM(GetUniqueId, "ncclGetUniqueId");
      CFD_SYM(CommInitRank, "ncclCommInitRank");
      CFD_SYM(CommDestroy, "ncclCommDestroy");
      CFD_SYM(Send, "ncclSend");
      CFD_SYM(Recv, "ncclRecv");
      CFD_SYM(AllReduce, "ncclAllReduce");
      CFD_SYM(AllGather, "ncclAllGather");
      CFD_SYM(Broadcast, "ncclBroadcast");
      CFD_SYM(GroupStart, "ncclGroupStart");
      CFD_SYM(GroupEnd, "ncclGroupEnd");
      CFD_SYM(GetErrorString, "ncclGetErrorString");
#undef CFD_SYM
      api.ok = api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.Send && api.Recv && api.AllReduce && api.AllGather && api.Broadcast &&
               api.GroupStart && api.GroupEnd && api.GetErrorString;
    }
  }
  return api;
}

#define CFD_NCCL_READY()                                                                            \
  do {                                                                                              \
    if (!nccl_api().ok) return fail(CFD_ERR_NCCL, "libnccl.so.2 could not be loaded (needed for world_size > 1)"); \
  } while (0)

#define CFD_NCCL(expr)                                                                              \
  do {                                                                                              \
    ncclResult_t _r = (expr);                                                                       \
    if (_r != ncclSuccess)                                                                          \
      return fail(CFD_ERR_NCCL, std::string(#expr) + ": " + nccl_api().GetErrorString(_r) + " (" + __FILE__ + \
                                    ":" + std::to_string(__LINE__) + ")");                          \
  } while (0)

// A grouped NCCL section that is always closed: CFD_NCCL returns on the first failing call, and a group left open would
// swallow every later NCCL call of the thread (including ncclCommDestroy) instead of reporting the one error.
struct NcclGroupGuard {
  bool open = false;
  ncclResult_t begin() {
    const ncclResult_t r = nccl_api().GroupStart();
    open = r == ncclSuccess;
    return r;
  }
  ncclResult_t end() {
    open = false;
    return nccl_api().GroupEnd();
  }
  ~NcclGroupGuard() {
    if (open) nccl_api().GroupEnd();
  }
};

void consts_default(cfd_solver_consts* c) {
  c->ramp_up_steps = 100;        // src/model.rs:269
  c->jacobi_iterations = 50;     // :737
  c->outer_rounds = 20;          // :696
  c->cg_max_iterations = 20000;  // extension
  c->jacobi_omega = 0.75;        // :735
  c->pressure_tolerance = 1e-4;  // :736
  c->outer_tolerance = 1e-4;     // :721
  c->cfl = 0.2;                  // :885
  c->cg_tolerance = 1e-8;        // extension
  c->mg_omega = 0.8;             // extension (MGCG)
  c->mg_smoothing = 2;           // extension (MGCG)
  c->mg_warm_start = 3;          // extension (MGCG)
  c->cg_relative = 0;            // extension (CG, MGCG)
  c->adaptive_substeps = 0;      // extension (src/model.rs:352-363 is commented out in the reference)
}

constexpr int kMaxSweepSlots = 256;
constexpr int kJacobiRows = 32;

struct ModelBase {
  virtual ~ModelBase() {}
  virtual int init() = 0;
  virtual int update() = 0;
  virtual int set_params(const cfd_params& p) = 0;
  virtual int get_snapshot(float* p, float* u, float* v, float* dt) = 0;
  virtual int snapshot_begin(float* p, float* u, float* v) = 0;
  virtual int snapshot_end(float* dt) = 0;
  virtual int get_residuals(cfd_residuals* out) = 0;
  virtual int field_len(int field, uint64_t* len) = 0;
  virtual int get_field_f64(int field, double* out, uint64_t len) = 0;
  virtual int set_field_f64(int field, const double* in, uint64_t len) = 0;
  virtual int rows(uint64_t* j0, uint64_t* j1) = 0;
  virtual int last_timing(double* step_ms, double* sweep_ms, uint64_t* launches) = 0;
  virtual int render(int mode, unsigned char* rgba, float* min_out, float* max_out) = 0;
  virtual int tracers_inject() = 0;
  virtual int tracers_update(double dt) = 0;
  virtual int tracers_get(double* xy, uint64_t capacity, uint64_t* n) = 0;
  virtual int tracers_clear() = 0;
  virtual int profile_smoother(int enable) = 0;
  virtual int last_smoother_timing(double* ms, uint64_t* launches) = 0;
};

// One rotating / plain field: `base` is the allocation, `v` the VIRTUAL ORIGIN such that v[j * rowlen + i] is
// entry (i, j) in GLOBAL row numbering.  A rank of a strip decomposition allocates only its own rows plus kHalo
// rows on each side, and every kernel keeps the reference's global flat indexing (so the "next row" wrap-around
// reads, SURVEY N2, still fall out of the index arithmetic).
template <class T>
struct Field {
  T* base = nullptr;
  T* v = nullptr;
  size_t rowlen = 0;
  size_t count = 0;  // entries allocated
  T* row(long j) const { return v + j * (long)rowlen; }
};

// Row strips: rank r owns array rows [strip_row_start(r), strip_row_start(r + 1)).  Interior boundaries sit at
// 1 + (a multiple of kStripAlign) so that the unknown rows (array row - 1) of a strip pair up within the strip on the
// first four multigrid levels (cfd_mg.cuh); any partition gives bit-identical Mode R results.
constexpr int kStripAlign = 16;
inline int strip_row_start(int ny, int world, int r) {
  if (r <= 0) return 0;
  if (r >= world) return ny;
  const long unknowns = (long)ny - 2;
  const long b = (r * unknowns / world) / kStripAlign * kStripAlign;
  return (int)(1 + b);
}

constexpr int kPredHalo = 2;  // rows: the second-order predictor reaches j +- 2 (src/model.rs:999, 1044, 1195, 1239)
constexpr int kHalo = 4;      // halo rows allocated: the V(4,4) legs of the multigrid cycle read 4 (cfd_mg_legs.cuh)

template <class R>
struct ModelImpl final : ModelBase {
  // ---- problem definition ----
  cfd_grid grid;
  cfd_options opt;
  int nx, ny;
  size_t n_p, n_u, n_v;
  // ---- strip owned by this rank: pressure / u rows [ja, jb); v rows [ja, jb) (+ row ny on the last rank) ----
  int rank = 0, world = 1, ja = 0, jb = 0;
  bool owns_bottom = true, owns_top = true;
  int rows_alloc = 0;  // jb - ja + 2*kHalo + 1
  ncclComm_t comm = nullptr;
  // ---- scalars kept on the host in R precision, same arithmetic as the reference (update(), :304-379) ----
  R dx, dy, lx, ly, dt, nu;
  R current_inlet_velocity = 0, target_inlet_velocity = 0;
  R last_pressure_residual = 0, last_u_residual = 0, last_v_residual = 0, simulation_time = 0;
  uint64_t simulation_step = 0, substep_count = 1, last_piso_substeps = 0;
  int velocity_scheme = 0, pressure_solver = 0, inlet_profile = 0, scenario = 0;
  uint64_t last_K = 0, last_S = 0;
  double last_step_seconds = 0, last_step_ms = 0, last_sweep_ms = 0;
  uint64_t last_launches = 0, launches = 0;
  // ---- device state ----
  int device = 0;
  cudaStream_t stream = nullptr;
  static constexpr size_t kFront = 256 / sizeof(R);
  Field<R> ubuf[3], vbuf[3];
  int iu = 0, ius = 1, ifree = 2;  // roles of the three u (and v) buffers: current, star, free
  Field<R> p, rhs, pp[2];
  int ipp = 0;  // pp[ipp] is p_prime, the other one p_prime_new
  Field<uint8_t> mask_u, mask_v, solid;
  unsigned long long* err_slots = nullptr;   // kMaxSweepSlots
  unsigned long long* step_slots = nullptr;  // 4
  cfdk::JacobiResult* h_jres = nullptr;      // pinned, device-visible
  unsigned long long* h_step = nullptr;      // pinned, 4
  void* staging = nullptr;                   // device scratch for read-back conversions
  size_t staging_bytes = 0;
  void* h_staging = nullptr;                 // pinned host scratch
  size_t h_staging_bytes = 0;
  void* bounce[2] = {nullptr, nullptr};      // pinned bounce buffers for snapshots into pageable memory
  cudaEvent_t bounce_ev[2] = {nullptr, nullptr};
  // Mode C (CG) work space, allocated on first use
  Field<R> cg_r, cg_d;
  double* cg_partials = nullptr;
  cfdk::CgScalars* cg_scalars = nullptr;   // device
  cfdk::CgScalars* h_cg = nullptr;         // pinned host copy
  // Mode C fast path (MGCG) work space, allocated on first use (cfd_mg.cuh)
  struct MgLevelHost {
    int mx = 0, my = 0;
    R* weights = nullptr;                      // WE, WW, CYW (mx each), WN, WS, CXH (my each)
    R *e = nullptr, *rho = nullptr, *tmp = nullptr;  // (mx + 2) x (my + 2), ring of zeros; level 0 has none
    R* cur = nullptr;                          // whichever of e / tmp holds the level's correction
    unsigned char* classes = nullptr;          // column + row classes of the level's diagonal table (MgLevelDev)
    cfdk::DivG<R>* diag_table = nullptr;
    cfdk::MgLevelDev<R> dev;
  };
  std::vector<MgLevelHost> mg;
  Field<R> mg_rho, mg_b[3];  // mg_b: the search direction d and the two smoothing buffers of z, roles rotate
  int mg_id = 0;             // index of the buffer that holds d
  // Start vector of a step's first solve (carried state).  mg_hist[0..2] = the p' the last three first-solves ended
  // with, newest first; the start vector 3 h0 - 3 h1 + h2 (mg_warm_start 3; 2 h0 - h1; h0) is formed on the fly by
  // k_mg_init, and the buffers ROTATE with the two p' buffers instead of being copied: after a first solve its
  // result stays in pp[ipp] (mg_rotate_pending) until the next MGCG solve is about to overwrite p', which then takes
  // the retired h2 buffer for p' and promotes the old p' buffer to h0.  Anything else that looks at these fields
  // resolves the pending rotation by a copy first (resolve_mg_rotation).
  Field<R> mg_hist[3];
  CUtensorMap tmap_hist[3];
  bool mg_rotate_pending = false;
  Field<R> mg_guess;               // explicit start vector: only after cfd_model_set_field_f64(CFD_FIELD_MG_GUESS)
  bool mg_guess_explicit = false;
  // An outer round whose re-correction solve is converged before its first iteration has p' == 0: its corrector is the
  // identity and is elided together with the solve's set-up pass.  Logical state while the flags are up:
  bool pp_zero_pending = false;    // p' is identically zero although pp[ipp] still holds older data
  bool star_alias = false;         // u_star / v_star equal the pre-BC current fields; only the entries the next
                                   // predictor does not overwrite were actually copied (k_star_save_*)
  int mg_pred[2] = {3, 0};         // iterations the previous step's first / re-correction solves took (batch size)
  int solid_i0 = 0, solid_i1 = 0, solid_j0 = 0, solid_j1 = 0;  // bounding box of the solid cells (empty: no obstacle)
  CUtensorMap tmap_mg_b[3], tmap_mg_rho, tmap_mg_rho_halo;
  cfdk::MgScalars* mg_scalars = nullptr;  // device
  cfdk::MgScalars* h_mg = nullptr;        // pinned host copy
  double* mg_partials = nullptr;
  unsigned long long* mg_err = nullptr;   // scratch max|dz| slot of the smoothing sweeps
  unsigned* mg_ticket = nullptr;          // last-block ticket of the fused dot-product reductions
  bool mg_finish_launch = true;           // dot products of k_mg_init / k_mg_dir_apply / k_mg_update: per-block partials + k_mg_reduce
  bool fuse_corr_div = true;              // k_corrector_div (CFD_FUSED_CORRECTOR=0: the separate kernels)
  int corr_rows = cfdk::kCorrRows;        // its tile height and width (CFD_CORR_ROWS / CFD_CORR_THREADS: A/B forms)
  int corr_threads = cfdk::kCorrThreads;
  int dir_rows = cfdk::kMgDirRows, dir_threads = cfdk::kMgThreads;  // tiles of k_mg_dir_apply / k_mg_update (CFD_DIR_TILE /
  int upd_rows = cfdk::kMgUpdRows, upd_threads = cfdk::kMgThreads;  // CFD_UPD_TILE = "<rows>x<threads>": A/B forms)
  int pred_rows = cfdk::kPredRows;        // rows a block of k_predict_first walks (CFD_PRED_ROWS: A/B hook)
  int init_rows = cfdk::kMgRows;          // rows a block of k_mg_init walks (CFD_INIT_ROWS: A/B hook)
  int div_rows = cfdk::kDivRows;          // tile height of k_divergence (CFD_DIV_ROWS=8: A/B form)
  bool corr_div_ready = false;            // rhs and the rhs^2 partials of the fields the last corrector wrote are in place
  // measurement hook: CUDA-event pairs around every k_jacobi_sweep5 launch of the MGCG smoother (bench.py roofline)
  bool prof_smoother = false;
  std::vector<cudaEvent_t> ev_prof;
  size_t ev_prof_used = 0;
  double last_prof_ms = 0;
  uint64_t last_prof_launches = 0;
  int mg_bottom_level = 0;                // first level run by the single-block bottom kernel
  unsigned* persist_barrier = nullptr;    // two alternating arrival counters of k_jacobi_persist2
  unsigned long long persist_launches = 0;
  bool persist2_ready = false;
  size_t mg_bottom_smem = 0;              // > 0: the bottom kernel keeps its levels in this much shared memory
  int mg_last_z = -1;                     // mg_b index of the last V-cycle's result (CFD_FIELD_MG_Z)
  bool mg_legs = false;                   // V-cycle legs as single launches (cfd_mg_legs.cuh)
  // tiles of the leg kernels: large levels (by the number of smoothing sweeps NU: the staged region is the tile + NU) and
  // small levels (more blocks)
  template <int NU> struct LegTile { static constexpr int TX = NU >= 4 ? 64 : 128, TY = NU >= 4 ? 32 : 16; };
  static constexpr int kLegSX = 64, kLegSY = 8;
  // form of the leg kernels: 3 register-tiled (shipped), 2 column strips, 1 flat-indexed (A/B and cross-checks: CFD_MG_LEGS=1|2|3)
  static int leg_form() {
    static const int form = [] { const char* e = getenv("CFD_MG_LEGS"); return e ? atoi(e) : 3; }();
    return form;
  }
  int leg_tx() const { return mg_smoothing() >= 4 ? 64 : 128; }
  int leg_ty() const { return mg_smoothing() >= 4 ? 32 : 16; }
  // strips: halo rows of a coarse level's rho the legs need (x_nu is recomputed on nu rows beyond the owned ones)
  int leg_rho_halo() const { return 2 * mg_smoothing() - 1; }
  cfdk::MgBottom<R> mg_bottom;
  CUtensorMap tmap_pp[2], tmap_rhs;  // 2-D tiled views of p' (ping, pong) and rhs for the tensor-TMA sweep
  CUtensorMap tmap_rhs_halo;         // rhs with the same halo box as p' (two-sweep kernel)
  int t2_rows_per_block = 20;        // tile height of the two-sweep kernel: rows + 4 halo rows = whole 4-row boxes
  cfdk::DivG<R> div_dx_sq, div_dy_sq, div_denom;  // divisors of the Jacobi update with hoisted reciprocals
  cfdk::StepDivs<R>* d_divs = nullptr;            // device: dx, dy, dx^2, dy^2, dt (refreshed every step), denom
  cfdk::StepDivs<R> h_divs;                       // host mirror (dt entry as of model creation)
  int sweep_rows_per_block = 32;
  int sweep6_resident_blocks = 148 * 4;  // one wave of the persistent sweep
  cudaEvent_t ev_step0 = nullptr, ev_step1 = nullptr;
  std::vector<cudaEvent_t> ev_sweep;  // pairs, one per pressure solve of the step
  size_t ev_sweep_used = 0;
  // asynchronous snapshots: two device staging buffers, a copy stream, per-slot events
  cudaStream_t copy_stream = nullptr;
  float* snap_stage[2] = {nullptr, nullptr};
  size_t snap_stage_floats = 0;
  cudaEvent_t snap_ready[2] = {nullptr, nullptr}, snap_done[2] = {nullptr, nullptr};
  float snap_dt[2] = {0.0f, 0.0f};
  int snap_head = 0, snap_in_flight = 0;  // slots [head - in_flight, head) mod 2 are in flight
  // tracer particles (cfd_tracers.cuh): positions in injection order, double-buffered for the stable compaction
  double2* tr_pos[2] = {nullptr, nullptr};
  unsigned char* tr_keep = nullptr;
  unsigned* tr_count_dev = nullptr;
  unsigned* h_tr_count = nullptr;  // pinned
  size_t tr_capacity = 0;
  unsigned tr_n = 0;
  int tr_cur = 0;
  Field<R> uold_buf, vold_buf;        // u_old / v_old of a step with several sub-steps (adaptive_substeps)
  bool old_in_copy = false;
  // Mode C bookkeeping of the last step's first solve (cfd_residuals::p_rel_f64, rhs_rms_f64, first_solve_iterations)
  double step_bb = 0, last_p_rel = 0, last_rhs_rms = 0;
  uint64_t last_first_solve_iterations = 0;
  // strips over peer memory (CUDA IPC): the neighbours' p' buffers and every rank's mailbox (cfd_kernels.cuh)
  cfdk::Mailbox* mailbox = nullptr;              // this rank's, device memory
  cfdk::Mailbox* peer_mailbox[cfdk::kMaxRanks] = {};
  R* peer_pp_down[2] = {nullptr, nullptr};       // lower neighbour's pp[0], pp[1] (virtual origins)
  R* peer_pp_up[2] = {nullptr, nullptr};
  std::vector<void*> ipc_opened;
  unsigned int* tickets = nullptr;               // [4][260]
  unsigned long long solve_counter = 0;
  unsigned long long* peer_trace = nullptr;
  double dbg_work = 0, dbg_wait = 0, dbg_gap = 0, dbg_span = 0; unsigned long long dbg_n = 0, dbg_solves = 0;
  bool peer_ready = false;
  // generic strips over peer memory (cfd_peer.cuh): every exchanged buffer is registered, its IPC handle all-gathered once,
  // and exchange_rows / exchange_level / gather_level / the scalar all-reduces run as single small launches instead of NCCL
  struct PeerBuf {
    void* base = nullptr;
    bool want_all = false;                   // mapped on every rank (gathered arrays), not only on the two neighbours
    void* map[cfdk::kPeerMaxRanks] = {};     // [rank]: that rank's buffer mapped into this process (nullptr: not mapped)
  };
  std::vector<PeerBuf> peer_bufs;
  size_t peer_synced = 0;
  bool persist_ready = false;
  bool peer2_ready = false;
  cfdk::PeerBox* box = nullptr;
  cfdk::PeerBox* peer_box[cfdk::kPeerMaxRanks] = {};
  unsigned int* peer_ticket = nullptr;
  unsigned long long xseq = 0, rseq = 0, gseq = 0;
  bool ready = false;

  ModelImpl(const cfd_grid& g, const cfd_params& prm, const cfd_options& o) : grid(g), opt(o) {
    nx = (int)g.nx;
    ny = (int)g.ny;
    n_p = (size_t)nx * ny;
    n_u = (size_t)(nx + 1) * ny;
    n_v = (size_t)nx * (ny + 1);
    dx = R(g.dx); dy = R(g.dy); lx = R(g.lx); ly = R(g.ly);
    dt = R(prm.dt);          // src/model.rs:265
    nu = R(prm.viscosity);   // :266
    target_inlet_velocity = R(prm.target_inlet_velocity);
    velocity_scheme = prm.velocity_scheme;
    pressure_solver = prm.pressure_solver;
    inlet_profile = prm.inlet_profile;
    scenario = prm.scenario;
    rank = o.rank;
    world = o.world_size;
    ja = strip_row_start(ny, world, rank);
    jb = strip_row_start(ny, world, rank + 1);
    owns_bottom = rank == 0;
    owns_top = rank == world - 1;
    rows_alloc = (jb - ja) + 2 * kHalo + 1;
  }

  ~ModelImpl() override {
    if (device >= 0) cudaSetDevice(device);
    if (stream) cudaStreamSynchronize(stream);
    if (comm) nccl_api().CommDestroy(comm);
    for (auto& f : ubuf) cudaFree(f.base);
    for (auto& f : vbuf) cudaFree(f.base);
    cudaFree(uold_buf.base); cudaFree(vold_buf.base);
    if (copy_stream) cudaStreamSynchronize(copy_stream);
    for (int k = 0; k < 2; ++k) {
      cudaFree(snap_stage[k]);
      if (snap_ready[k]) cudaEventDestroy(snap_ready[k]);
      if (snap_done[k]) cudaEventDestroy(snap_done[k]);
    }
    if (copy_stream) cudaStreamDestroy(copy_stream);
    cudaFree(tr_pos[0]); cudaFree(tr_pos[1]); cudaFree(tr_keep); cudaFree(tr_count_dev);
    if (h_tr_count) cudaFreeHost(h_tr_count);
    cudaFree(p.base); cudaFree(rhs.base); cudaFree(pp[0].base); cudaFree(pp[1].base);
    cudaFree(mask_u.base); cudaFree(mask_v.base); cudaFree(solid.base);
    cudaFree(err_slots); cudaFree(step_slots); cudaFree(staging); cudaFree(tickets); cudaFree(d_divs);
    cudaFree(cg_r.base); cudaFree(cg_d.base); cudaFree(cg_partials); cudaFree(cg_scalars);
    for (auto& L : mg) { cudaFree(L.weights); cudaFree(L.e); cudaFree(L.rho); cudaFree(L.tmp); cudaFree(L.classes); cudaFree(L.diag_table); }
    cudaFree(mg_rho.base); cudaFree(mg_guess.base);
    for (auto& f : mg_hist) cudaFree(f.base);
    for (auto& f : mg_b) cudaFree(f.base);
    cudaFree(mg_scalars); cudaFree(mg_partials); cudaFree(mg_err); cudaFree(mg_ticket);
    if (h_mg) cudaFreeHost(h_mg);
    if (h_cg) cudaFreeHost(h_cg);
    if (h_jres) cudaFreeHost(h_jres);
    if (h_step) cudaFreeHost(h_step);
    if (h_staging) cudaFreeHost(h_staging);
    for (int k = 0; k < 2; ++k) {
      if (bounce[k]) cudaFreeHost(bounce[k]);
      if (bounce_ev[k]) cudaEventDestroy(bounce_ev[k]);
    }
    if (tickets && getenv("CFD_PEER_DEBUG")) {
      unsigned int h[16];
      cudaMemcpy(h, tickets + 1040, sizeof h, cudaMemcpyDeviceToHost);
      unsigned long long w0, w1;
      memcpy(&w0, h + 8, 8); memcpy(&w1, h + 10, 8);
      fprintf(stderr, "[cfd peer rank %d] halo waits %u, total %.3f ms (%.2f us each); last-block max waits %u, total %.3f ms\n",
              rank, h[4], w0 / 1.9e6, h[4] ? w0 / 1.9e3 / h[4] : 0.0, h[5], w1 / 1.9e6);
    }
    if (peer_trace) {
      std::vector<unsigned long long> t(768);
      cudaMemcpy(t.data(), peer_trace, 768 * 8, cudaMemcpyDeviceToHost);
      fprintf(stderr, "[cfd peer rank %d] saturated solves %llu: avg work %.1f us, exit wait %.1f us, gap %.1f us per sweep; avg solve span %.1f us\n",
              rank, dbg_solves, dbg_n ? dbg_work / dbg_n : 0.0, dbg_n ? dbg_wait / dbg_n : 0.0, dbg_n ? dbg_gap / dbg_n : 0.0,
              dbg_solves ? dbg_span / dbg_solves : 0.0);
      fprintf(stderr, "[cfd peer rank %d] last solve, per sweep: start(+us since sweep 0) work_us exit_wait_us gap_to_next_us\n", rank);
      for (int s2 = 0; s2 < 50; s2 += 1)
        fprintf(stderr, "[cfd peer rank %d] s=%2d start %9.1f work %7.1f wait %6.1f gap %6.1f\n", rank, s2,
                (t[s2] - t[0]) / 1e3, (t[256 + s2] - t[s2]) / 1e3, (t[512 + s2] - t[256 + s2]) / 1e3,
                s2 < 49 ? ((double)t[s2 + 1] - (double)t[512 + s2]) / 1e3 : 0.0);
      cudaFree(peer_trace);
    }
    for (void* ptr : ipc_opened) cudaIpcCloseMemHandle(ptr);
    cudaFree(mailbox);
    cudaFree(box); cudaFree(peer_ticket);
    if (ev_step0) cudaEventDestroy(ev_step0);
    if (ev_step1) cudaEventDestroy(ev_step1);
    for (auto e : ev_sweep) cudaEventDestroy(e);
    for (auto e : ev_prof) cudaEventDestroy(e);
    if (stream) cudaStreamDestroy(stream);
  }

  template <class T>
  int dalloc(T** ptr, size_t count) {
    CFD_CUDA(cudaMalloc((void**)ptr, count * sizeof(T)));
    CFD_CUDA(cudaMemsetAsync(*ptr, 0, count * sizeof(T), stream));
    return CFD_OK;
  }

  // rows [ja - kHalo, jb + kHalo] of a field with `rowlen` entries per row, zero-filled; 256 B of slack in
  // front (16-byte halo of the first TMA strip) and 4 rows behind (prefetch of the register / bulk sweeps)
  template <class T>
  int falloc(Field<T>* f, size_t rowlen) {
    const size_t front = 256 / sizeof(T);
    const size_t count = front + (size_t)rows_alloc * rowlen + 4 * rowlen;
    int rc;
    if ((rc = dalloc(&f->base, count))) return rc;
    f->rowlen = rowlen;
    f->count = count;
    f->v = f->base + front - (long)(ja - kHalo) * (long)rowlen;
    return CFD_OK;
  }

  // Model::new, src/model.rs:219-299: zero fields, masks from the cylinder
  int init() override {
    int ndev = 0;
    CFD_CUDA(cudaGetDeviceCount(&ndev));
    if (ndev <= 0) return fail(CFD_ERR_CUDA, "no CUDA device: cfd_b200 has no CPU fallback");
    if (opt.device >= 0) {
      CFD_CUDA(cudaSetDevice(opt.device));
      device = opt.device;
    } else {
      CFD_CUDA(cudaGetDevice(&device));
    }
    CFD_CUDA(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
    if (world > 1) {
      CFD_NCCL_READY();
      ncclUniqueId id;
      memcpy(&id, opt.nccl_unique_id, sizeof id);
      CFD_NCCL(nccl_api().CommInitRank(&comm, world, id, rank));
    }
    int rc;
    for (int k = 0; k < 3; ++k) {
      if ((rc = falloc(&ubuf[k], (size_t)nx + 1))) return rc;
      if ((rc = falloc(&vbuf[k], (size_t)nx))) return rc;
    }
    if ((rc = falloc(&p, (size_t)nx))) return rc;
    if ((rc = falloc(&rhs, (size_t)nx))) return rc;
    if ((rc = falloc(&pp[0], (size_t)nx))) return rc;
    if ((rc = falloc(&pp[1], (size_t)nx))) return rc;
    if ((rc = falloc(&mask_u, (size_t)nx + 1))) return rc;
    if ((rc = falloc(&mask_v, (size_t)nx))) return rc;
    if ((rc = falloc(&solid, (size_t)nx))) return rc;
    if ((rc = dalloc(&err_slots, (size_t)kMaxSweepSlots))) return rc;
    if ((rc = dalloc(&step_slots, (size_t)4))) return rc;
    // per-sweep tickets / stop flags / diagnostics (strips) and work counters (persistent sweep), cfd_kernels.cuh
    if ((rc = dalloc(&tickets, (size_t)1400))) return rc;
    CFD_CUDA(cudaHostAlloc((void**)&h_jres, sizeof(cfdk::JacobiResult), cudaHostAllocMapped));
    CFD_CUDA(cudaHostAlloc((void**)&h_step, 8 * sizeof(unsigned long long), cudaHostAllocMapped));
    CFD_CUDA(cudaEventCreate(&ev_step0));
    CFD_CUDA(cudaEventCreate(&ev_step1));
    const int n_pairs = 2 * (opt.consts.outer_rounds + 1);
    ev_sweep.resize(n_pairs);
    for (auto& e : ev_sweep) CFD_CUDA(cudaEventCreate(&e));

    cfdk::MaskGeom g;
    g.nx = nx; g.ny = ny;
    g.has_obstacle = grid.has_obstacle != 0;
    g.cavity = scenario == CFD_SCENARIO_CAVITY;
    g.dx = grid.dx; g.dy = grid.dy; g.cx = grid.center_x; g.cy = grid.center_y; g.radius = grid.radius;
    const int mj_lo = ja, mj_hi = owns_top ? jb + 1 : jb;
    dim3 blk(256), grd((nx + 1 + 255) / 256, mj_hi - mj_lo);
    cfdk::k_build_masks<<<grd, blk, 0, stream>>>(g, solid.v, mask_u.v, mask_v.v, mj_lo, mj_hi);
    CFD_CUDA(cudaGetLastError());
    CFD_CUDA(cudaStreamSynchronize(stream));
    if (grid.has_obstacle) {
      // bounding box of the cells whose centre can lie inside the cylinder (two cells of margin around the f32 test)
      const double x0 = ((double)grid.center_x - (double)grid.radius) / (double)grid.dx, x1 = ((double)grid.center_x + (double)grid.radius) / (double)grid.dx;
      const double y0 = ((double)grid.center_y - (double)grid.radius) / (double)grid.dy, y1 = ((double)grid.center_y + (double)grid.radius) / (double)grid.dy;
      auto clampi = [](double v, int lo, int hi) { return v < (double)lo ? lo : (v > (double)hi ? hi : (int)v); };
      solid_i0 = clampi(std::floor(x0) - 2.0, 0, nx); solid_i1 = clampi(std::ceil(x1) + 3.0, 0, nx);
      solid_j0 = clampi(std::floor(y0) - 2.0, 0, ny); solid_j1 = clampi(std::ceil(y1) + 3.0, 0, ny);
    }
    if ((rc = init_sweep_constants())) return rc;
    for (int k = 0; k < 3; ++k) { peer_register(ubuf[k].base); peer_register(vbuf[k].base); }
    peer_register(pp[0].base); peer_register(pp[1].base);
    if ((rc = peer_sync())) return rc;
    // Mode R strips: NCCL halo rows + allreduce after every sweep by default; CFD_FLAG_PEER_EXCHANGE opts into the fused
    // peer-memory path (faster; it hung at start-up in 2 of 8 two-GPU runs before its max records were double-buffered,
    // and has passed only one run since)
    if (world > 1 && (opt.flags & CFD_FLAG_PEER_EXCHANGE) && !(opt.flags & CFD_FLAG_NCCL_EXCHANGE) && (rc = init_peer_memory())) return rc;
    ready = true;
    return CFD_OK;
  }

  // Map the neighbours' p' buffers and every rank's mailbox into this process (CUDA IPC; the 64-byte handles
  // travel through one NCCL all-gather).  After this the sweep loop needs no NCCL call (k_jacobi_sweep5).
  int init_peer_memory() {
    if (world > cfdk::kMaxRanks) return fail(CFD_ERR_UNSUPPORTED, "peer-memory strips support at most 8 ranks");
    int rc;
    if ((rc = dalloc(&mailbox, (size_t)1))) return rc;
    if (getenv("CFD_PEER_DEBUG") && (rc = dalloc(&peer_trace, (size_t)3 * 256))) return rc;
    CFD_CUDA(cudaStreamSynchronize(stream));
    struct Handles { cudaIpcMemHandle_t pp0, pp1, box; };
    static_assert(sizeof(Handles) == 192, "three 64-byte IPC handles");
    std::vector<Handles> all((size_t)world);
    Handles mine;
    CFD_CUDA(cudaIpcGetMemHandle(&mine.pp0, pp[0].base));
    CFD_CUDA(cudaIpcGetMemHandle(&mine.pp1, pp[1].base));
    CFD_CUDA(cudaIpcGetMemHandle(&mine.box, mailbox));
    unsigned char* d = nullptr;
    CFD_CUDA(cudaMalloc((void**)&d, sizeof(Handles) * (size_t)(world + 1)));
    CFD_CUDA(cudaMemcpyAsync(d, &mine, sizeof mine, cudaMemcpyHostToDevice, stream));
    CFD_NCCL(nccl_api().AllGather(d, d + sizeof(Handles), sizeof(Handles), ncclChar, comm, stream));
    CFD_CUDA(cudaMemcpyAsync(all.data(), d + sizeof(Handles), sizeof(Handles) * (size_t)world, cudaMemcpyDeviceToHost, stream));
    CFD_CUDA(cudaStreamSynchronize(stream));
    CFD_CUDA(cudaFree(d));
    auto open = [&](const cudaIpcMemHandle_t& h, void** out) -> int {
      CFD_CUDA(cudaIpcOpenMemHandle(out, h, cudaIpcMemLazyEnablePeerAccess));
      ipc_opened.push_back(*out);
      return CFD_OK;
    };
    for (int r = 0; r < world; ++r) {
      if (r == rank) { peer_mailbox[r] = mailbox; continue; }
      void* ptr = nullptr;
      if ((rc = open(all[(size_t)r].box, &ptr))) return rc;
      peer_mailbox[r] = (cfdk::Mailbox*)ptr;
    }
    // a neighbour lays its buffers out exactly like this rank does (falloc): virtual origin = base + front -
    // (its first owned row - kHalo) * nx
    auto neighbour_origin = [&](void* base, int r) -> R* {
      const int ja_n = strip_row_start(ny, world, r);
      return (R*)base + kFront - (long)(ja_n - kHalo) * (long)nx;
    };
    if (rank > 0) {
      void *b0 = nullptr, *b1 = nullptr;
      if ((rc = open(all[(size_t)rank - 1].pp0, &b0))) return rc;
      if ((rc = open(all[(size_t)rank - 1].pp1, &b1))) return rc;
      peer_pp_down[0] = neighbour_origin(b0, rank - 1);
      peer_pp_down[1] = neighbour_origin(b1, rank - 1);
    }
    if (rank < world - 1) {
      void *b0 = nullptr, *b1 = nullptr;
      if ((rc = open(all[(size_t)rank + 1].pp0, &b0))) return rc;
      if ((rc = open(all[(size_t)rank + 1].pp1, &b1))) return rc;
      peer_pp_up[0] = neighbour_origin(b0, rank + 1);
      peer_pp_up[1] = neighbour_origin(b1, rank + 1);
    }
    peer_ready = true;
    return CFD_OK;
  }

  // ---- generic peer layer -----------------------------------------------------------------------------------------
  bool want_peer2() const {
    if (world <= 1 || world > cfdk::kPeerMaxRanks) return false;
    if (opt.flags & CFD_FLAG_NCCL_EXCHANGE) return false;
    if (const char* e = getenv("CFD_PEER_STRIPS")) return atoi(e) != 0;  // A/B hook
    return (opt.flags & CFD_FLAG_PEER_STRIPS) != 0;
  }
  void peer_register(void* base, bool want_all = false) {
    if (!want_peer2() || !base) return;
    PeerBuf b;
    b.base = base;
    b.want_all = want_all;
    peer_bufs.push_back(b);
  }
  const PeerBuf* peer_find(const void* base) const {
    if (!peer2_ready) return nullptr;
    for (size_t k = 0; k < peer_synced; ++k)
      if (peer_bufs[k].base == base) return &peer_bufs[k];
    return nullptr;
  }
  // collective: all-gathers the IPC handles of the buffers registered since the last call (every rank registers the same
  // buffers in the same order) and maps the neighbours' (or everyone's) copies; the first call also sets up the mailboxes
  int peer_sync() {
    if (!want_peer2()) return CFD_OK;
    int rc;
    if (!box) {
      if ((rc = dalloc(&box, (size_t)1))) return rc;
      if ((rc = dalloc(&peer_ticket, (size_t)4))) return rc;
      PeerBuf b;
      b.base = box;
      b.want_all = true;
      peer_bufs.insert(peer_bufs.begin() + (long)peer_synced, b);
    }
    const size_t n_new = peer_bufs.size() - peer_synced;
    if (n_new == 0) return CFD_OK;
    CFD_CUDA(cudaStreamSynchronize(stream));
    std::vector<cudaIpcMemHandle_t> mine(n_new), all(n_new * (size_t)world);
    for (size_t k = 0; k < n_new; ++k) CFD_CUDA(cudaIpcGetMemHandle(&mine[k], peer_bufs[peer_synced + k].base));
    unsigned char* d = nullptr;
    const size_t bytes = n_new * sizeof(cudaIpcMemHandle_t);
    CFD_CUDA(cudaMalloc((void**)&d, bytes * (size_t)(world + 1)));
    CFD_CUDA(cudaMemcpyAsync(d, mine.data(), bytes, cudaMemcpyHostToDevice, stream));
    CFD_NCCL(nccl_api().AllGather(d, d + bytes, bytes, ncclChar, comm, stream));
    CFD_CUDA(cudaMemcpyAsync(all.data(), d + bytes, bytes * (size_t)world, cudaMemcpyDeviceToHost, stream));
    CFD_CUDA(cudaStreamSynchronize(stream));
    CFD_CUDA(cudaFree(d));
    for (size_t k = 0; k < n_new; ++k) {
      PeerBuf& b = peer_bufs[peer_synced + k];
      for (int r = 0; r < world; ++r) {
        if (r == rank) { b.map[r] = b.base; continue; }
        if (!b.want_all && r != rank - 1 && r != rank + 1) continue;
        void* ptr = nullptr;
        CFD_CUDA(cudaIpcOpenMemHandle(&ptr, all[(size_t)r * n_new + k], cudaIpcMemLazyEnablePeerAccess));
        ipc_opened.push_back(ptr);
        b.map[r] = ptr;
      }
      if (b.base == box)
        for (int r = 0; r < world; ++r) peer_box[r] = (cfdk::PeerBox*)b.map[r];
    }
    peer_synced = peer_bufs.size();
    peer2_ready = true;
    // nobody pushes into a mailbox or a buffer before every rank has mapped everything: one more collective as a barrier
    CFD_NCCL(nccl_api().AllReduce(peer_ticket + 2, peer_ticket + 2, 1, ncclUint32, ncclMax, comm, stream));
    CFD_CUDA(cudaStreamSynchronize(stream));
    return CFD_OK;
  }
  cfdk::PeerAll peer_all() const {
    cfdk::PeerAll a;
    for (int r = 0; r < cfdk::kPeerMaxRanks; ++r) a.box[r] = r < world ? peer_box[r] : nullptr;
    a.rank = rank; a.world = world;
    return a;
  }
  // rows of a strip field or of a replicated coarse array to the two neighbours: src / dst as element offsets from the
  // (virtual or real) origins, which the caller computed for its own copy and for the neighbours' copies
  int peer_push(const void* src_down, void* dst_down, size_t bytes_down, const void* src_up, void* dst_up, size_t bytes_up,
                unsigned long long* max_data = nullptr, int max_n = 0) {
    cfdk::PeerPush p;
    memset(&p, 0, sizeof p);
    p.mine = box; p.ticket = peer_ticket; p.seq = ++xseq;
    if (max_data) { p.red = max_data; p.red_n = max_n; p.red_seq = ++rseq; p.all = peer_all(); }
    if (rank > 0) {
      p.flag[0] = &peer_box[rank - 1]->from_above;
      p.src[0] = (const uint32_t*)src_down; p.dst[0] = (uint32_t*)dst_down; p.words[0] = bytes_down / 4;
    }
    if (rank < world - 1) {
      p.flag[1] = &peer_box[rank + 1]->from_below;
      p.src[1] = (const uint32_t*)src_up; p.dst[1] = (uint32_t*)dst_up; p.words[1] = bytes_up / 4;
    }
    const size_t most = p.words[0] > p.words[1] ? p.words[0] : p.words[1];
    int grid = (int)((most + 1023) / 1024);
    if (grid < 1) grid = 1;
    if (grid > 64) grid = 64;
    cfdk::k_peer_push<<<grid, 256, 0, stream>>>(p);
    ++launches;
    CFD_CUDA(cudaGetLastError());
    return CFD_OK;
  }
  int peer_reduce(void* data, int n, int op) {
    cfdk::k_peer_reduce<<<1, 32, 0, stream>>>(peer_all(), (unsigned long long*)data, n, op, ++rseq, nullptr);
    ++launches;
    CFD_CUDA(cudaGetLastError());
    return CFD_OK;
  }

  cfdk::SweepPeer<R> sweep_peer(int out) const {
    cfdk::SweepPeer<R> sp;
    memset(&sp, 0, sizeof sp);
    sp.rank = rank;
    sp.world = 1;
    if (!peer_ready) return sp;
    sp.world = world;
    sp.down_out = peer_pp_down[out];
    sp.up_out = peer_pp_up[out];
    sp.down_flag = rank > 0 ? &peer_mailbox[rank - 1]->halo_flag[1] : nullptr;
    sp.up_flag = rank < world - 1 ? &peer_mailbox[rank + 1]->halo_flag[0] : nullptr;
    sp.mine = mailbox;
    for (int r = 0; r < world; ++r) sp.all[r] = peer_mailbox[r];
    sp.tickets = tickets;
    sp.stamp_base = solve_counter * 256ull;
    sp.trace = peer_trace;
    return sp;
  }

  // 2-D tensor maps (cuTensorMapEncodeTiled through the runtime's driver entry point; libcuda is not linked)
  int make_tensor_map(CUtensorMap* map, R* first_row, int box_cols) {
    typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                 const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                 CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static EncodeFn encode = nullptr;
    if (!encode) {
      void* fn = nullptr;
      cudaDriverEntryPointQueryResult q;
      CFD_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
      if (!fn || q != cudaDriverEntryPointSuccess) return fail(CFD_ERR_CUDA, "cuTensorMapEncodeTiled not available");
      encode = (EncodeFn)fn;
    }
    const cuuint64_t dims[2] = {(cuuint64_t)nx, (cuuint64_t)rows_alloc};
    const cuuint64_t strides[1] = {(cuuint64_t)nx * sizeof(R)};
    const cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)cfdk::kChunkRows};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = encode(map, sizeof(R) == 8 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT64 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2,
                              (void*)first_row, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                              CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(CFD_ERR_CUDA, "cuTensorMapEncodeTiled failed with code " + std::to_string((int)r));
    return CFD_OK;
  }

  // divisors of the Jacobi update (src/model.rs:740-746) and, for fp64, their reciprocals refined by the
  // same instruction sequence the compiler's division uses (cfdk::div_fast); launch geometry of the sweep
  int init_sweep_constants() {
    // every loop-invariant divisor with its hoisted reciprocal and dividend window, built on the device
    // (cfdk::make_divg) and mirrored on the host for the kernels that take them by value
    {
      int rc0;
      if (!d_divs && (rc0 = dalloc(&d_divs, (size_t)1))) return rc0;
      const R denom = R(2.0) / (dx * dx) + R(2.0) / (dy * dy);  // :746
      cfdk::k_step_divisors<R><<<1, 32, 0, stream>>>(d_divs, dx, dy, dt, denom);
      CFD_CUDA(cudaMemcpyAsync(&h_divs, d_divs, sizeof h_divs, cudaMemcpyDeviceToHost, stream));
      CFD_CUDA(cudaStreamSynchronize(stream));
      div_dx_sq = h_divs.dx_sq;   // :740
      div_dy_sq = h_divs.dy_sq;   // :742
      div_denom = h_divs.denom;   // :746
    }
    {
      using Ring = cfdk::SweepChunkRing<R>;
      int rc2;
      // the maps describe the LOCAL allocation: its row 0 is global row ja - kHalo
      if ((rc2 = make_tensor_map(&tmap_pp[0], pp[0].row(ja - kHalo), Ring::kPCols))) return rc2;
      if ((rc2 = make_tensor_map(&tmap_pp[1], pp[1].row(ja - kHalo), Ring::kPCols))) return rc2;
      if ((rc2 = make_tensor_map(&tmap_rhs, rhs.row(ja - kHalo), cfdk::kStripCols))) return rc2;
      CFD_CUDA(cudaFuncSetAttribute(cfdk::k_jacobi_sweep5<R>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)sizeof(Ring)));
#ifdef CFD_WITH_AB_SWEEPS
      if ((rc2 = make_tensor_map(&tmap_rhs_halo, rhs.row(ja - kHalo), Ring::kPCols))) return rc2;
      CFD_CUDA(cudaFuncSetAttribute(cfdk::k_jacobi_sweep_t2<R>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)sizeof(cfdk::SweepT2Ring<R>)));
      CFD_CUDA(cudaFuncSetAttribute(cfdk::k_jacobi_sweep4<R>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)sizeof(Ring)));
      CFD_CUDA(cudaFuncSetAttribute(cfdk::k_jacobi_sweep6<R>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)sizeof(Ring)));
      int per_sm6 = 1;
      CFD_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm6, cfdk::k_jacobi_sweep6<R>, cfdk::kSweepWarps * 32, sizeof(Ring)));
      cudaDeviceProp prop6;
      CFD_CUDA(cudaGetDeviceProperties(&prop6, device));
      sweep6_resident_blocks = prop6.multiProcessorCount * (per_sm6 < 1 ? 1 : per_sm6);
#endif
    }
    // one thread per column pair, 128 threads per block; pick the rows per block so that the grid is a
    // whole number of waves of (SM count x resident blocks per SM)
    cudaDeviceProp prop;
    CFD_CUDA(cudaGetDeviceProperties(&prop, device));
    const int sms = prop.multiProcessorCount;
    const int kSweepThreads = cfdk::kSweepWarps * 32;
    const int bx = (nx / 2 + kSweepThreads - 1) / kSweepThreads;
    int per_sm = 4;
#ifdef CFD_WITH_AB_SWEEPS
    if (opt.flags & CFD_FLAG_REGISTER_SWEEP)
      CFD_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, cfdk::k_jacobi_sweep2<R>, 128, 0));
    else if (opt.flags & CFD_FLAG_BULK_SWEEP)
      CFD_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, cfdk::k_jacobi_sweep3<R>, kSweepThreads, 0));
    else if (opt.flags & CFD_FLAG_SWEEP4)
      CFD_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, cfdk::k_jacobi_sweep4<R>, kSweepThreads,
                                                             sizeof(cfdk::SweepChunkRing<R>)));
    else
#endif
      CFD_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, cfdk::k_jacobi_sweep5<R>, kSweepThreads,
                                                             sizeof(cfdk::SweepChunkRing<R>)));
    if (per_sm < 1) per_sm = 1;
    const int resident = sms * per_sm;
    const int rows = sweep_row_end() - sweep_row_begin();
    int rpb;
    if (opt.flags & (CFD_FLAG_REGISTER_SWEEP | CFD_FLAG_BULK_SWEEP)) {
      const int bxw = (opt.flags & CFD_FLAG_REGISTER_SWEEP) ? (nx / 2 + 127) / 128 : bx;  // sweep2 keeps 128-thread blocks
      int gy = (resident + bxw - 1) / bxw;  // one wave
      if (gy > rows) gy = rows;
      if (gy < 1) gy = 1;
      rpb = (rows + gy - 1) / gy;
      if (rpb < 8 && rows >= 8) rpb = 8;
    } else {
      // Tile height of the tensor-TMA sweeps.  SMs progress at visibly different rates (profiles/
      // r1_notes.md: sm__cycles_active min 95k / max 194k at equal work), so the grid is cut into several
      // waves of short tiles that the block scheduler hands out dynamically; 22 rows (+2 halo rows = six
      // 4-row TMA boxes exactly) measured best at 4096^2 (tools/tune_sweep.py).  Small grids get shorter
      // tiles so that every SM still has work.
      const long target_blocks = 4L * resident;
      long want = ((long)rows * bx) / target_blocks;
      if (want > 22) want = 22;
      if (want < 6) want = 6;
      rpb = (int)(((want - 2) / 4) * 4 + 2);  // rows + 2 halo rows = whole number of 4-row boxes
    }
    if (const char* e = getenv("CFD_FUSED_CORRECTOR")) fuse_corr_div = atoi(e) != 0;  // A/B hooks
    if (const char* e = getenv("CFD_CORR_ROWS")) { const int v = atoi(e); corr_rows = (v == 8 || v == 4 || v == 1) ? v : cfdk::kCorrRows; }
    auto tile_hook = [](const char* name, int* rows_out, int* threads_out) {
      int r = 0, t = 0;
      const char* e = getenv(name);
      if (e && sscanf(e, "%dx%d", &r, &t) == 2 && (r == 2 || r == 4) && (t == 128 || t == 256)) { *rows_out = r; *threads_out = t; }
    };
    tile_hook("CFD_DIR_TILE", &dir_rows, &dir_threads);
    tile_hook("CFD_UPD_TILE", &upd_rows, &upd_threads);
    if (const char* e = getenv("CFD_PRED_ROWS")) { const int v = atoi(e); if (v >= 2 && v <= 1024) pred_rows = v; }
    if (const char* e = getenv("CFD_INIT_ROWS")) { const int v = atoi(e); if (v >= 2 && v <= 1024) init_rows = v; }
    if (const char* e = getenv("CFD_DIV_ROWS")) div_rows = atoi(e) == 8 ? 8 : cfdk::kDivRows;
    if (const char* e = getenv("CFD_CORR_THREADS")) corr_threads = atoi(e) == 256 ? 256 : cfdk::kCorrThreads;
    if (corr_rows == 8 || corr_rows == 1) corr_threads = 256;  // forms that exist at one width only
    if (const char* e = getenv("CFD_MG_FINISH_LAUNCH")) mg_finish_launch = atoi(e) != 0;
    if (const char* e = getenv("CFD_SWEEP_ROWS")) {  // tuning hook (tools/tune_sweep.py)
      const int v = atoi(e);
      if (v >= 2) rpb = v;
    }
    if (const char* e = getenv("CFD_T2_ROWS")) {
      const int v = atoi(e);
      if (v >= 4) t2_rows_per_block = v;
    } else {
      const long target_blocks = 4L * resident;
      long want = ((long)rows * bx) / target_blocks;
      t2_rows_per_block = want >= 20 ? 20 : (want >= 12 ? 12 : (want >= 8 ? 8 : 4));
    }
    sweep_rows_per_block = rpb;
    return CFD_OK;
  }

  // the divergence divides by dt (:1436): its hoisted reciprocal comes from the device (make_divg) whenever dt changed
  int refresh_dt_divisor(R dt_sub) {
    if (memcmp(&dt_sub, &h_divs.dt.y, sizeof(R)) == 0) return CFD_OK;
    cfdk::k_step_divisors<R><<<1, 32, 0, stream>>>(d_divs, dx, dy, dt_sub, R(2.0) / (dx * dx) + R(2.0) / (dy * dy));
    CFD_CUDA(cudaMemcpyAsync(&h_divs, d_divs, sizeof h_divs, cudaMemcpyDeviceToHost, stream));
    CFD_CUDA(cudaStreamSynchronize(stream));
    return CFD_OK;
  }

  // interior pressure rows this rank sweeps / owned row ranges of the staggered fields
  int sweep_row_begin() const { return ja > 1 ? ja : 1; }
  int sweep_row_end() const { return jb < ny - 1 ? jb : ny - 1; }
  int v_row_end() const { return owns_top ? jb + 1 : jb; }  // v rows [ja, v_row_end())

  cfdk::StepScalars<R> scalars(R dt_sub) const {
    cfdk::StepScalars<R> s;
    s.dx = dx; s.dy = dy; s.dt = dt_sub; s.nu = nu; s.nx = nx; s.ny = ny;
    return s;
  }

  // ---- strip plumbing: NCCL send/recv of whole rows with the two neighbours, max-allreduce of scalars ----
  ncclDataType_t nccl_real() const { return sizeof(R) == 8 ? ncclFloat64 : ncclFloat32; }

  // refresh `down` halo rows below row `a` and `up` halo rows above row `b` of a field whose owned rows are
  // [a, b): the lower neighbour owns [.., a), the upper one [b, ..)
  int exchange_rows(const Field<R>& f, int a, int b, int down, int up, int send_down, int send_up,
                    cudaStream_t on = nullptr, unsigned long long* max_data = nullptr) {
    if (world == 1) return CFD_OK;
    const size_t rl = f.rowlen;
    if (const PeerBuf* pb = on ? nullptr : peer_find(f.base)) {
      // the neighbours lay the field out exactly like this rank does (falloc): virtual origin = base + front -
      // (first owned row - kHalo) * row length, rows in GLOBAL numbering; what this rank sends lands on the same row there
      auto origin = [&](int r) -> R* {
        return (R*)pb->map[r] + 256 / sizeof(R) - (long)(strip_row_start(ny, world, r) - kHalo) * (long)rl;
      };
      R* dst_down = rank > 0 ? origin(rank - 1) + (long)a * (long)rl : nullptr;
      R* dst_up = rank < world - 1 ? origin(rank + 1) + (long)(b - send_up) * (long)rl : nullptr;
      return peer_push(f.row(a), dst_down, (size_t)send_down * rl * sizeof(R), f.row(b - send_up), dst_up,
                       (size_t)send_up * rl * sizeof(R), max_data, max_data ? 1 : 0);
    }
    cudaStream_t stream = on ? on : this->stream;
    NcclGroupGuard group;
    CFD_NCCL(group.begin());
    if (rank > 0) {
      if (send_down > 0) CFD_NCCL(nccl_api().Send(f.row(a), (size_t)send_down * rl, nccl_real(), rank - 1, comm, stream));
      if (down > 0) CFD_NCCL(nccl_api().Recv(f.row(a - down), (size_t)down * rl, nccl_real(), rank - 1, comm, stream));
    }
    if (rank < world - 1) {
      if (send_up > 0) CFD_NCCL(nccl_api().Send(f.row(b - send_up), (size_t)send_up * rl, nccl_real(), rank + 1, comm, stream));
      if (up > 0) CFD_NCCL(nccl_api().Recv(f.row(b), (size_t)up * rl, nccl_real(), rank + 1, comm, stream));
    }
    CFD_NCCL(group.end());
    return CFD_OK;
  }
  // symmetric halo of depth d on a field with owned rows [a, b)
  int exchange_halo(const Field<R>& f, int a, int b, int d, cudaStream_t on = nullptr) {
    return exchange_rows(f, a, b, d, d, d, d, on);
  }
  // one row from above only (v* row jb feeds the divergence of row jb-1, src/model.rs:1430)
  int fetch_row_above(const Field<R>& f, int a, int b) { return exchange_rows(f, a, b, 0, 1, 1, 0); }

  int allreduce_max_u64(unsigned long long* d, size_t n, cudaStream_t on = nullptr) {
    if (world == 1) return CFD_OK;
    if (peer2_ready && !on && n <= 4) return peer_reduce(d, (int)n, 0);
    cudaStream_t stream = on ? on : this->stream;
    CFD_NCCL(nccl_api().AllReduce(d, d, n, ncclUint64, ncclMax, comm, stream));
    return CFD_OK;
  }

  // device scalars -> mapped pinned host memory by a one-warp kernel (not by the copy engine: see k_publish_words)
  void publish(const void* dev, void* host_mapped, size_t bytes) {
    cfdk::k_publish_words<<<1, 32, 0, stream>>>((const unsigned*)dev, (volatile unsigned*)host_mapped, (int)(bytes / 4));
  }

  // CUDA-event pair of the next pressure solve of this step (begin = [k], end = [k + 1])
  int next_solve_events(size_t* k) {
    if (ev_sweep_used + 2 > ev_sweep.size()) {
      const size_t old_n = ev_sweep.size();
      ev_sweep.resize(ev_sweep_used + 16, nullptr);
      for (size_t q = old_n; q < ev_sweep.size(); ++q) CFD_CUDA(cudaEventCreate(&ev_sweep[q]));
    }
    *k = ev_sweep_used;
    ev_sweep_used += 2;
    return CFD_OK;
  }

  // ---- one pressure solve: recompute_divergence (:1406-1440) + jacobi_pressure (:734-824) ----
  // `elided` (may be null): set when the solve converged before its first iteration from p' = 0, i.e. the correction
  // is identically zero and nothing was written to p' (pp_zero_pending) — the caller skips the corrector.
  int pressure_solve(R dt_sub, const Field<R>& us, const Field<R>& vs, int call_index, R* residual_out, bool* elided = nullptr) {
    const int iters = opt.consts.jacobi_iterations;
    int rc;
    if (elided) *elided = false;
    size_t ev = 0;
    if ((rc = next_solve_events(&ev))) return rc;
    if ((rc = fetch_row_above(vs, ja, v_row_end()))) return rc;
    const bool mgcg = pressure_solver == CFD_SOLVER_MGCG;
    const bool first_solve = call_index == 0;
    if (mgcg && (rc = mgcg_begin(first_solve))) return rc;
    // a cold-start MGCG solve that is expected to need no iteration: the divergence kernel sums rho.rho = rhs.rhs itself
    const bool cold = mgcg && !(call_index == 0 && opt.consts.mg_warm_start != 0);
    const bool decide_early = cold && elided != nullptr && mg_pred[first_solve ? 0 : 1] == 0;
    const bool have_rhs = corr_div_ready;  // the corrector that produced us / vs also left their divergence and its rhs^2 partials
    corr_div_ready = false;
    if (have_rhs) {
      if (!decide_early) return fail(CFD_ERR_UNSUPPORTED, "fused corrector + divergence without an early decision");
      const dim3 grd((nx + corr_threads - 1) / corr_threads, (ny + corr_rows - 1) / corr_rows);  // k_corrector_div's grid
      cfdk::k_mg_reduce<R><<<1, 1024, 0, stream>>>(mg_fine(dt_sub), mg_scalars, mg_partials, (int)(grd.x * grd.y), 0);
      ++launches;
    } else {
      // MGCG: a step's first solve also needs ||rhs||^2 over the unknowns (reference of the relative stopping rule and of
      // the reported ||r|| / ||rhs||): rr_mode 4, or 5 when that solve starts cold (then rho = rhs and the sum is rho.rho too)
      const int rr_mode = mgcg && first_solve ? (cold ? 5 : 4) : 0;
      dim3 blk(256), grd((nx + 255) / 256, (jb - ja + div_rows - 1) / div_rows);
      if (decide_early || rr_mode != 0) {
        const cfdk::MgFine<R> c = mg_fine(dt_sub);
        if (div_rows == 8)
          cfdk::k_divergence<R, true, 8><<<grd, blk, 0, stream>>>(scalars(dt_sub), us.v, vs.v, rhs.v, ja, jb, err_slots, iters, tickets,
                                                                  h_divs.dx, h_divs.dy, h_divs.dt, c, mg_scalars, mg_partials,
                                                                  mg_ticket, rr_mode);
        else
        cfdk::k_divergence<R, true><<<grd, blk, 0, stream>>>(scalars(dt_sub), us.v, vs.v, rhs.v, ja, jb, err_slots, iters, tickets,
                                                             h_divs.dx, h_divs.dy, h_divs.dt, c, mg_scalars, mg_partials,
                                                             mg_ticket, rr_mode);
        cfdk::k_mg_reduce<R><<<1, 1024, 0, stream>>>(c, mg_scalars, mg_partials, (int)(grd.x * grd.y), rr_mode);
        ++launches;
        if (rr_mode != 0 && (rc = mg_finish_strips(c, rr_mode))) return rc;  // strips: the ranks' sums -> bb
      } else {
        if (div_rows == 8)
          cfdk::k_divergence<R, false, 8><<<grd, blk, 0, stream>>>(scalars(dt_sub), us.v, vs.v, rhs.v, ja, jb, err_slots, iters, tickets,
                                                                   h_divs.dx, h_divs.dy, h_divs.dt, cfdk::MgFine<R>{}, nullptr,
                                                                   nullptr, nullptr, 0);
        else
        cfdk::k_divergence<R, false><<<grd, blk, 0, stream>>>(scalars(dt_sub), us.v, vs.v, rhs.v, ja, jb, err_slots, iters, tickets,
                                                              h_divs.dx, h_divs.dy, h_divs.dt, cfdk::MgFine<R>{}, nullptr,
                                                              nullptr, nullptr, 0);
      }
      ++launches;
    }
    if (pressure_solver == CFD_SOLVER_CG) return cg_solve(dt_sub, call_index, residual_out, ev);
    if (mgcg) return mgcg_solve(dt_sub, call_index, residual_out, decide_early, elided, ev);
    if ((rc = materialize_pp_zero())) return rc;  // Jacobi warm-starts from p' (src/model.rs:734-824)
    if ((rc = resolve_mg_rotation())) return rc;
    cfdk::JacobiConsts<R> c;
    c.dx_sq = dx * dx;                                   // :740
    c.dy_sq = dy * dy;                                   // :742
    c.denom = R(2.0) / (dx * dx) + R(2.0) / (dy * dy);   // :746
    c.omega = R(opt.consts.jacobi_omega);
    c.one_minus_omega = R(1.0) - c.omega;                // :745
    c.tol = R(opt.consts.pressure_tolerance);
    c.nx = nx; c.ny = ny; c.cavity = scenario == CFD_SCENARIO_CAVITY;
    c.row_begin = sweep_row_begin(); c.row_end = sweep_row_end();
    const int rows = c.row_end - c.row_begin;
    CFD_CUDA(cudaEventRecord(ev_sweep[ev], stream));
    cfdk::JacobiConsts2<R> c2;
    c2.dx_sq = div_dx_sq; c2.dy_sq = div_dy_sq; c2.denom = div_denom;
    c2.omega = c.omega; c2.one_minus_omega = c.one_minus_omega; c2.tol = c.tol;
    c2.nx = nx; c2.ny = ny; c2.cavity = c.cavity; c2.rows_per_block = sweep_rows_per_block;
    c2.row_begin = c.row_begin; c2.row_end = c.row_end; c2.row_shift = ja - kHalo;
    c2.check_lag = 1;
    c2.fix_pass = -1;
    const dim3 blk1(256), grd1((nx - 2 + 255) / 256, (rows + kJacobiRows - 1) / kJacobiRows);
    const int kSweepThreads = cfdk::kSweepWarps * 32;
    const dim3 blk2(kSweepThreads), grd2((nx / 2 + kSweepThreads - 1) / kSweepThreads, (rows + sweep_rows_per_block - 1) / sweep_rows_per_block);
    const size_t ring_bytes = sizeof(cfdk::SweepChunkRing<R>);
    const bool tuned_default = !(opt.flags & (CFD_FLAG_BASELINE_SWEEP | CFD_FLAG_REGISTER_SWEEP | CFD_FLAG_BULK_SWEEP | CFD_FLAG_SWEEP4));
    const bool use6 = tuned_default && (opt.flags & CFD_FLAG_PERSISTENT_SWEEP);  // persistent warp-queue kernel (A/B)
#ifdef CFD_WITH_AB_SWEEPS
    const int n_units6 = ((nx + cfdk::kStripCols - 1) / cfdk::kStripCols) * ((rows + cfdk::kUnitRows - 1) / cfdk::kUnitRows);
    int grid6 = (n_units6 + cfdk::kSweepWarps - 1) / cfdk::kSweepWarps;
    if (grid6 > sweep6_resident_blocks) grid6 = sweep6_resident_blocks;
    if (grid6 < 1) grid6 = 1;
#endif
    const bool use_t2 = world == 1 && tuned_default && (opt.flags & CFD_FLAG_TEMPORAL) && (iters % 2 == 0) &&
                        rows >= 4;
    bool persisted = false;
    if (world == 1 && tuned_default && !use6 && !(opt.flags & (CFD_FLAG_TEMPORAL | CFD_FLAG_NO_GRAPH))) {
      // ---- small grids: the whole solve in ONE cooperative launch, p' / p'new / rhs of a block's rows in shared memory
      int sms = 0, max_smem = 0, coop = 0;
      CFD_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
      CFD_CUDA(cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, device));
      CFD_CUDA(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, device));
      // as many rows per block as shared memory holds (fewer blocks make the grid barrier cheaper; a block's 1024 threads
      // update its rows in a few passes), but at least one block per ~8 rows so that small grids still use many SMs
      // (r2y, 800 x 264, k_jacobi_persist2: 2 rows per block 5.2 us per sweep, 4: 6.0, 8: 8.1, 16: 9.1 — a block's rows cost
      // more than its share of the barrier: use every SM)
      int threads = 1024, target_rows = 2;
      if (const char* e = getenv("CFD_PERSIST_THREADS")) threads = atoi(e);   // tuning hooks
      if (const char* e = getenv("CFD_PERSIST_ROWS")) target_rows = atoi(e);
      if (const char* e = getenv("CFD_PERSIST_FORM")) { if (atoi(e) == 1 && !getenv("CFD_PERSIST_ROWS")) target_rows = 8; }
      int blocks = (rows + target_rows - 1) / target_rows;
      if (blocks > sms) blocks = sms;
      if (blocks < 1) blocks = 1;
      int rb = (rows + blocks - 1) / blocks;
      while (rb > 1 && ((size_t)2 * (rb + 2) + rb) * (size_t)nx * sizeof(R) + 1024 > (size_t)max_smem && (rows + rb - 2) / (rb - 1) <= sms) --rb;
      const size_t smem = ((size_t)2 * (rb + 2) + rb) * (size_t)nx * sizeof(R);
      if (coop && rows >= 1 && smem + 1024 <= (size_t)max_smem) {
        if (!persist_ready) {
          CFD_CUDA(cudaFuncSetAttribute(cfdk::k_jacobi_persist<R>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem - 1024));
          persist_ready = true;
        }
        cfdk::PersistArgs<R> pa;
        pa.c = c2; pa.pp0 = pp[0].v; pa.pp1 = pp[1].v; pa.rhs = rhs.v;
        pa.ipp = ipp; pa.iters = iters; pa.rows_per_block = rb; pa.err_slots = err_slots; pa.out = h_jres;
        // CFD_PERSIST_FORM=1: the first form (grid.sync once per sweep, A/B); default: barrier off the critical path
        static const bool form1 = getenv("CFD_PERSIST_FORM") != nullptr && atoi(getenv("CFD_PERSIST_FORM")) == 1;
        if (form1) {
          void* args[] = {&pa};
          CFD_CUDA(cudaLaunchCooperativeKernel((const void*)cfdk::k_jacobi_persist<R>, dim3((rows + rb - 1) / rb), dim3(threads), args, smem, stream));
        } else {
          if (!persist_barrier && (rc = dalloc(&persist_barrier, (size_t)2))) return rc;
          if (!persist2_ready) {
            CFD_CUDA(cudaFuncSetAttribute(cfdk::k_jacobi_persist2<R>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem - 1024));
            persist2_ready = true;
          }
          cfdk::PersistArgs2<R> p2;
          p2.a = pa;
          p2.barrier = persist_barrier + (persist_launches & 1);
          p2.barrier_next = persist_barrier + ((persist_launches + 1) & 1);
          ++persist_launches;
          const int pairs = nx / 2;
          p2.tc = pairs < threads ? pairs : threads;
          p2.groups = threads / p2.tc;
          if (p2.groups > rb) p2.groups = rb;
          if (p2.groups < 1) p2.groups = 1;
          void* args[] = {&p2};
          CFD_CUDA(cudaLaunchCooperativeKernel((const void*)cfdk::k_jacobi_persist2<R>, dim3((rows + rb - 1) / rb), dim3(p2.tc * p2.groups), args, smem, stream));
        }
        ++launches;
        persisted = true;
      }
    }
    if (persisted) {
#ifdef CFD_WITH_AB_SWEEPS
    } else if (use_t2) {
      // ---- temporal blocking: two sweeps per pass over HBM (k_jacobi_sweep_t2), each pass followed by the
      // conditional fix-up that restores the reference's stopping point when the FIRST sweep of a pass converged
      cfdk::JacobiConsts2<R> ct = c2, cf = c2;
      ct.rows_per_block = t2_rows_per_block;
      const dim3 grd_t(grd2.x, (rows + t2_rows_per_block - 1) / t2_rows_per_block);
      const size_t t2_bytes = sizeof(cfdk::SweepT2Ring<R>);
      for (int pass = 0; pass < iters / 2; ++pass) {
        const int in = (ipp + pass) & 1, out = in ^ 1, s = 2 * pass;
        cfdk::k_jacobi_sweep_t2<R><<<grd_t, blk2, t2_bytes, stream>>>(ct, tmap_pp[in], tmap_rhs_halo, pp[out].v, err_slots, s);
        cf.fix_pass = pass;
        cfdk::k_jacobi_sweep5<R><<<grd2, blk2, ring_bytes, stream>>>(cf, tmap_pp[in], tmap_rhs, pp[out].v, err_slots, s,
                                                                      cfdk::SweepPeer<R>{});
        launches += 2;
      }
#endif
    } else if (world > 1 && tuned_default && peer_ready) {
      // ---- strips over peer memory: the sweep stores its edge rows into the neighbours' halos and publishes its
      // max|dp'| to every rank's mailbox itself; convergence is checked two sweeps late (check_lag 2).
      c2.check_lag = 2;
      ++solve_counter;
      for (int s = 0; s < iters; ++s) {
        const int in = (ipp + s) & 1, out = in ^ 1;
#ifdef CFD_WITH_AB_SWEEPS
        if (use6)
          cfdk::k_jacobi_sweep6<R><<<grid6, blk2, ring_bytes, stream>>>(c2, tmap_pp[in], tmap_rhs, pp[out].v, err_slots, s,
                                                                         sweep_peer(out), tickets + 1060 + s);
        else
#endif
          cfdk::k_jacobi_sweep5<R><<<grd2, blk2, ring_bytes, stream>>>(c2, tmap_pp[in], tmap_rhs, pp[out].v, err_slots, s,
                                                                        sweep_peer(out));
        ++launches;
      }
      cfdk::k_jacobi_finalize_peer<R><<<1, 32, 0, stream>>>(mailbox, world, solve_counter * 256ull, iters, c.tol, h_jres);
      ++launches;
    } else {
      for (int s = 0; s < iters; ++s) {
        const int in = (ipp + s) & 1, out = in ^ 1;
        if (opt.flags & CFD_FLAG_BASELINE_SWEEP)
          cfdk::k_jacobi_sweep<R, kJacobiRows><<<grd1, blk1, 0, stream>>>(c, pp[in].v, rhs.v, pp[out].v, err_slots, s);
#ifdef CFD_WITH_AB_SWEEPS
        else if (opt.flags & CFD_FLAG_REGISTER_SWEEP)
          cfdk::k_jacobi_sweep2<R><<<dim3((nx / 2 + 127) / 128, grd2.y), dim3(128), 0, stream>>>(c2, pp[in].v, rhs.v, pp[out].v,
                                                                                               err_slots, s);  // fixed 128-thread blocks
        else if (opt.flags & CFD_FLAG_BULK_SWEEP)
          cfdk::k_jacobi_sweep3<R><<<grd2, blk2, 0, stream>>>(c2, pp[in].v, rhs.v, pp[out].v, err_slots, s);
        else if (opt.flags & CFD_FLAG_SWEEP4)
          cfdk::k_jacobi_sweep4<R><<<grd2, blk2, ring_bytes, stream>>>(c2, tmap_pp[in], tmap_rhs, pp[out].v, err_slots, s);
        else if (use6)
          cfdk::k_jacobi_sweep6<R><<<grid6, blk2, ring_bytes, stream>>>(c2, tmap_pp[in], tmap_rhs, pp[out].v, err_slots, s,
                                                                         cfdk::SweepPeer<R>{}, tickets + 1060 + s);
#endif
        else
          cfdk::k_jacobi_sweep5<R><<<grd2, blk2, ring_bytes, stream>>>(c2, tmap_pp[in], tmap_rhs, pp[out].v, err_slots, s,
                                                                        cfdk::SweepPeer<R>{});
        ++launches;
        if (world > 1) {
          // strips, simple form: the next sweep needs the neighbours' new boundary rows and the GLOBAL max|dp'|
          // of this one.  A rank that skipped the sweep (converged) still takes part; what it exchanges is
          // never consumed.
          if (peer_find(pp[out].base)) {  // halo rows and the max in ONE launch over peer memory
            if ((rc = exchange_rows(pp[out], ja, jb, 1, 1, 1, 1, nullptr, err_slots + s))) return rc;
          } else {
            if ((rc = exchange_halo(pp[out], ja, jb, 1))) return rc;
            if ((rc = allreduce_max_u64(err_slots + s, 1))) return rc;
          }
        }
      }
    }
    if (!persisted && !(world > 1 && tuned_default && peer_ready)) {
      cfdk::k_jacobi_finalize<R><<<1, 32, 0, stream>>>(err_slots, iters, c.tol, h_jres);
      ++launches;
    }
    CFD_CUDA(cudaEventRecord(ev_sweep[ev + 1], stream));
    CFD_CUDA(cudaGetLastError());
    CFD_CUDA(cudaStreamSynchronize(stream));
    const int ran = h_jres->sweeps;
    const int flips = use_t2 ? (ran + 1) / 2 : ran;  // buffer swaps performed: one per pass / per sweep
    if (peer_trace && ran == iters) {  // diagnostics: GPU-side timeline of this solve
      std::vector<unsigned long long> t(768);
      cudaMemcpy(t.data(), peer_trace, 768 * 8, cudaMemcpyDeviceToHost);
      for (int s2 = 0; s2 < iters; ++s2) {
        dbg_work += (t[256 + s2] - t[s2]) / 1e3;
        dbg_wait += (t[512 + s2] - t[256 + s2]) / 1e3;
        if (s2 + 1 < iters) dbg_gap += ((double)t[s2 + 1] - (double)t[512 + s2]) / 1e3;
        ++dbg_n;
      }
      dbg_span += (t[512 + iters - 1] - t[0]) / 1e3;
      ++dbg_solves;
    }
    ipp = (ipp + flips) & 1;
    last_S += (uint64_t)ran;
    last_K += 1;
    *residual_out = (R)h_jres->last_error;
    return CFD_OK;
  }

  // EXTENSION, Mode C: conjugate gradients on the Jacobi iteration's own discrete problem (cfd_kernels.cuh).
  // Strips: one halo row of the search direction per iteration, the two dot products sum-allreduced over NCCL.
  int cg_reduce(const cfdk::CgConsts<R>& c, int n, int mode) {
    if (world == 1) {
      cfdk::k_cg_reduce<R><<<1, 1024, 0, stream>>>(c, cg_scalars, cg_partials, n, mode, 0);
      ++launches;
      return CFD_OK;
    }
    cfdk::k_cg_reduce<R><<<1, 1024, 0, stream>>>(c, cg_scalars, cg_partials, n, mode, 1);
    if (peer2_ready) {
      int rc;
      if ((rc = peer_reduce(&cg_scalars->local_sum, 1, 1))) return rc;
    } else {
      CFD_NCCL(nccl_api().AllReduce(&cg_scalars->local_sum, &cg_scalars->local_sum, 1, ncclFloat64, ncclSum, comm, stream));
    }
    cfdk::k_cg_reduce<R><<<1, 32, 0, stream>>>(c, cg_scalars, cg_partials, n, mode, 2);
    launches += 2;
    return CFD_OK;
  }

  int cg_solve(R dt_sub, int call_index, R* residual_out, size_t ev) {
    int rc;
    if ((rc = resolve_mg_rotation())) return rc;
    pp_zero_pending = false;  // k_cg_init writes p' in full
    const int int_lo = sweep_row_begin(), int_hi = sweep_row_end();
    const dim3 blk(cfdk::kCgThreads);
    const dim3 g_all((nx + cfdk::kCgThreads - 1) / cfdk::kCgThreads, jb - ja);
    const dim3 g_int((nx - 2 + cfdk::kCgThreads - 1) / cfdk::kCgThreads, int_hi - int_lo);
    const int n_all = (int)(g_all.x * g_all.y), n_int = (int)(g_int.x * g_int.y);
    if (!cg_r.base) {
      if ((rc = falloc(&cg_r, (size_t)nx))) return rc;
      if ((rc = falloc(&cg_d, (size_t)nx))) return rc;
      if ((rc = dalloc(&cg_partials, (size_t)n_all))) return rc;
      if ((rc = dalloc(&cg_scalars, (size_t)1))) return rc;
      CFD_CUDA(cudaHostAlloc((void**)&h_cg, sizeof(cfdk::CgScalars), cudaHostAllocMapped));
      peer_register(cg_d.base);
      if ((rc = peer_sync())) return rc;
    }
    cfdk::CgConsts<R> c;
    c.dx_sq = dx * dx; c.dy_sq = dy * dy; c.dt = dt_sub; c.tol = R(opt.consts.cg_tolerance);
    c.n_unknowns = R((size_t)(nx - 2) * (size_t)(ny - 2));
    c.nx = nx; c.ny = ny; c.cavity = scenario == CFD_SCENARIO_CAVITY;
    c.own_lo = ja; c.own_hi = jb; c.int_lo = int_lo;
    const Field<R>& xf = pp[ipp];
    R* x = xf.v;
    R* q = pp[ipp ^ 1].v;
    cfdk::CgScalars init;
    memset(&init, 0, sizeof init);
    init.max_iterations = opt.consts.cg_max_iterations;
    init.relative = opt.consts.cg_relative;
    init.bb = call_index == 0 ? -1.0 : step_bb;  // first solve of the step: bb <- r.r of the cold start (k_cg_reduce)
    *h_cg = init;
    CFD_CUDA(cudaEventRecord(ev_sweep[ev], stream));
    CFD_CUDA(cudaMemcpyAsync(cg_scalars, h_cg, sizeof init, cudaMemcpyHostToDevice, stream));
    cfdk::k_cg_init<R><<<g_all, blk, 0, stream>>>(c, rhs.v, x, cg_r.v, cg_d.v, cg_partials);
    ++launches;
    if ((rc = cg_reduce(c, n_all, 0))) return rc;
    const int batch = 32;
    for (;;) {
      publish(cg_scalars, h_cg, sizeof init);
      CFD_CUDA(cudaStreamSynchronize(stream));
      if (h_cg->done) break;
      for (int it = 0; it < batch; ++it) {
        if ((rc = exchange_halo(cg_d, ja, jb, 1))) return rc;
        cfdk::k_cg_apply<R><<<g_int, blk, 0, stream>>>(c, cg_scalars, cg_d.v, q, cg_partials);
        if ((rc = cg_reduce(c, n_int, 1))) return rc;
        cfdk::k_cg_update<R><<<g_int, blk, 0, stream>>>(c, cg_scalars, cg_d.v, q, x, cg_r.v, cg_partials);
        if ((rc = cg_reduce(c, n_int, 2))) return rc;
        cfdk::k_cg_direction<R><<<g_int, blk, 0, stream>>>(c, cg_scalars, cg_r.v, cg_d.v);
        launches += 3;
      }
      CFD_CUDA(cudaGetLastError());
    }
    const int n_edge = (nx > ny ? nx : ny);
    cfdk::k_cg_fill_boundary<R><<<(n_edge + 255) / 256, 256, 0, stream>>>(nx, ny, c.cavity, x, ja, jb);
    ++launches;
    if ((rc = exchange_halo(xf, ja, jb, 1))) return rc;  // the corrector reads p'[j-1] (src/model.rs:1380)
    CFD_CUDA(cudaEventRecord(ev_sweep[ev + 1], stream));
    CFD_CUDA(cudaGetLastError());
    if (call_index == 0) note_first_solve(h_cg->bb, h_cg->rel, h_cg->iterations, dt_sub);
    last_S += (uint64_t)h_cg->iterations;
    last_K += 1;
    *residual_out = (R)h_cg->measure;
    return CFD_OK;
  }

  // EXTENSION, Mode C fast path: CG preconditioned by one multigrid V-cycle (cfd_mg.cuh).  Level geometry is
  // computed on the host in R precision with the same expressions as the oracle (mg_build_levels).
  int mg_setup() {
    const bool cavity = scenario == CFD_SCENARIO_CAVITY;
    const R dx_sq = dx * dx, dy_sq = dy * dy;
    std::vector<R> wx((size_t)nx - 2, R(1)), hy((size_t)ny - 2, R(1));
    int rc;
    for (;;) {
      MgLevelHost L;
      L.mx = (int)wx.size(); L.my = (int)hy.size();
      const size_t mx = wx.size(), my = hy.size();
      std::vector<R> hw(3 * mx + 3 * my, R(0));
      R *WE = hw.data(), *WW = WE + mx, *CYW = WW + mx, *WN = CYW + mx, *WS = WN + my, *CXH = WS + my;
      for (size_t i = 0; i < mx; ++i) {
        if (i + 1 < mx) WE[i] = R(1) / (R(0.5) * (wx[i] + wx[i + 1]));
        else if (!cavity) WE[i] = R(1) / (R(0.5) * wx[i] + R(0.5));
        if (i > 0) WW[i] = R(1) / (R(0.5) * (wx[i - 1] + wx[i]));
        CYW[i] = wx[i] / dy_sq;
      }
      for (size_t j = 0; j < my; ++j) {
        if (j + 1 < my) WN[j] = R(1) / (R(0.5) * (hy[j] + hy[j + 1]));
        if (j > 0) WS[j] = R(1) / (R(0.5) * (hy[j - 1] + hy[j]));
        CXH[j] = hy[j] / dx_sq;
      }
      if ((rc = dalloc(&L.weights, hw.size()))) return rc;
      CFD_CUDA(cudaMemcpyAsync(L.weights, hw.data(), hw.size() * sizeof(R), cudaMemcpyHostToDevice, stream));
      CFD_CUDA(cudaStreamSynchronize(stream));  // hw goes out of scope
      L.dev.mx = L.mx; L.dev.my = L.my;
      L.dev.WE = L.weights; L.dev.WW = L.weights + mx; L.dev.CYW = L.weights + 2 * mx;
      L.dev.WN = L.weights + 3 * mx; L.dev.WS = L.weights + 3 * mx + my; L.dev.CXH = L.weights + 3 * mx + 2 * my;
      L.dev.col_class = nullptr; L.dev.row_class = nullptr; L.dev.diag_table = nullptr;
      if (!mg.empty() && !(mx == 1 && my == 1)) {
        // classes of columns by (WE + WW, CYW) and of rows by (WN + WS, CXH): the diagonal of a cell depends on them only
        auto classify = [](const R* a, const R* b, size_t n, std::vector<unsigned char>& cls, std::vector<std::pair<R, R>>& keys) {
          cls.resize(n);
          for (size_t k = 0; k < n; ++k) {
            const std::pair<R, R> key(a[k], b[k]);
            size_t c = 0;
            while (c < keys.size() && memcmp(&keys[c], &key, sizeof key) != 0) ++c;
            if (c == keys.size()) keys.push_back(key);
            if (keys.size() > (size_t)cfdk::kMgClasses) return false;
            cls[k] = (unsigned char)c;
          }
          return true;
        };
        std::vector<R> a(mx), b(my);
        for (size_t i = 0; i < mx; ++i) a[i] = WE[i] + WW[i];
        for (size_t j = 0; j < my; ++j) b[j] = WN[j] + WS[j];
        std::vector<unsigned char> ccls, rcls;
        std::vector<std::pair<R, R>> ckeys, rkeys;
        if (classify(a.data(), CYW, mx, ccls, ckeys) && classify(b.data(), CXH, my, rcls, rkeys)) {
          std::vector<R> diag((size_t)cfdk::kMgClasses * cfdk::kMgClasses, R(0));
          for (size_t r = 0; r < rkeys.size(); ++r)
            for (size_t c = 0; c < ckeys.size(); ++c)  // CXH[J] * (WE[I] + WW[I]) + CYW[I] * (WN[J] + WS[J]), like mgc_sweep_cell
              diag[r * cfdk::kMgClasses + c] = rkeys[r].second * ckeys[c].first + ckeys[c].second * rkeys[r].first;
          unsigned char* d_cls = nullptr;
          R* d_diag = nullptr;
          if ((rc = dalloc(&d_cls, mx + my))) return rc;
          if ((rc = dalloc(&d_diag, diag.size()))) return rc;
          if ((rc = dalloc(&L.diag_table, diag.size() + 1))) return rc;  // + the common dividend window (cfd_mg_legs.cuh)
          CFD_CUDA(cudaMemcpyAsync(d_cls, ccls.data(), mx, cudaMemcpyHostToDevice, stream));
          CFD_CUDA(cudaMemcpyAsync(d_cls + mx, rcls.data(), my, cudaMemcpyHostToDevice, stream));
          CFD_CUDA(cudaMemcpyAsync(d_diag, diag.data(), diag.size() * sizeof(R), cudaMemcpyHostToDevice, stream));
          cfdk::k_mgc_diag_table<R><<<1, 64, 0, stream>>>(d_diag, L.diag_table, (int)diag.size());
          CFD_CUDA(cudaStreamSynchronize(stream));
          CFD_CUDA(cudaFree(d_diag));
          L.classes = d_cls;
          L.dev.col_class = d_cls; L.dev.row_class = d_cls + mx; L.dev.diag_table = L.diag_table;
        }
      }
      if (!mg.empty()) {
        const size_t n = (mx + 2) * (my + 2);
        if ((rc = dalloc(&L.e, n))) return rc;
        if ((rc = dalloc(&L.rho, n))) return rc;
        if ((rc = dalloc(&L.tmp, n))) return rc;
      }
      mg.push_back(L);
      if (mx == 1 && my == 1) break;
      auto pair_up = [](const std::vector<R>& w) {
        std::vector<R> o((w.size() + 1) / 2, R(0));
        for (size_t k = 0; k < o.size(); ++k) o[k] = w[2 * k] + (2 * k + 1 < w.size() ? w[2 * k + 1] : R(0));
        return o;
      };
      wx = pair_up(wx);
      hy = pair_up(hy);
    }
    if ((rc = falloc(&mg_rho, (size_t)nx))) return rc;
    for (auto& f : mg_b)
      if ((rc = falloc(&f, (size_t)nx))) return rc;
    using Ring = cfdk::SweepChunkRing<R>;
    for (int k = 0; k < 3; ++k) {
      if ((rc = falloc(&mg_hist[k], (size_t)nx))) return rc;
      if ((rc = make_tensor_map(&tmap_hist[k], mg_hist[k].row(ja - kHalo), Ring::kPCols))) return rc;  // may become a p' buffer
    }
    for (int k = 0; k < 3; ++k)
      if ((rc = make_tensor_map(&tmap_mg_b[k], mg_b[k].row(ja - kHalo), Ring::kPCols))) return rc;
    CFD_CUDA(cudaFuncSetAttribute(cfdk::k_jacobi_sweep5<R, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)sizeof(Ring)));
    if ((rc = make_tensor_map(&tmap_mg_rho, mg_rho.row(ja - kHalo), cfdk::kStripCols))) return rc;
    if ((rc = make_tensor_map(&tmap_mg_rho_halo, mg_rho.row(ja - kHalo), Ring::kPCols))) return rc;
    CFD_CUDA(cudaFuncSetAttribute(cfdk::k_jacobi_sweep5<R, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Ring)));
    if ((rc = dalloc(&mg_scalars, (size_t)1))) return rc;
    if ((rc = dalloc(&mg_err, (size_t)kMaxSweepSlots))) return rc;
    if ((rc = dalloc(&mg_ticket, (size_t)4))) return rc;
    {
      // one partial per block of the largest grid that ends in a dot product: the 4-row vector tiles, or the sweep
      const size_t n_vec = (size_t)((nx + 255) / 256) * (size_t)((ny + 1) / 2 + 1);  // down to 2-row x 128-thread tiles
      const size_t n_sweep = (size_t)((nx / 2 + cfdk::kSweepWarps * 32 - 1) / (cfdk::kSweepWarps * 32)) * (size_t)((ny - 2 + sweep_rows_per_block - 1) / sweep_rows_per_block);
      const size_t n_leg = (size_t)((nx - 2 + 55) / 56) * (size_t)((ny - 2 + 15) / 16 + 1);  // at least the leg kernels' grids
      const size_t n_div = (size_t)((nx + 127) / 128) * (size_t)(ny + 1);  // k_divergence, every form of k_corrector_div
      size_t n_max = n_vec > n_sweep ? n_vec : n_sweep;
      if (n_leg > n_max) n_max = n_leg;
      if (n_div > n_max) n_max = n_div;
      if ((rc = dalloc(&mg_partials, n_max))) return rc;
    }
    // the bottom of the V-cycle (every level from the first that fits 64 x 64) runs in one single-block launch
    mg_bottom_level = (int)mg.size() - 1;
    for (int l = 1; l < (int)mg.size(); ++l)
      if (mg[(size_t)l].mx <= 64 && mg[(size_t)l].my <= 64) { mg_bottom_level = l; break; }
    if ((int)mg.size() - mg_bottom_level > cfdk::kMgBottomMax) return fail(CFD_ERR_UNSUPPORTED, "multigrid hierarchy too deep");
    mg_ld = world > 1 ? (mg_bottom_level - 1 < 3 ? mg_bottom_level - 1 : 3) : -1;
    memset(&mg_bottom, 0, sizeof mg_bottom);
    mg_bottom.n = (int)mg.size() - mg_bottom_level;
    mg_bottom.nu = mg_smoothing();
    mg_bottom.omega = R(opt.consts.mg_omega);
    size_t bottom_bytes = 0;
    for (int k = 0; k < mg_bottom.n; ++k) {
      const MgLevelHost& L = mg[(size_t)(mg_bottom_level + k)];
      mg_bottom.lv[k].dev = L.dev;
      mg_bottom.lv[k].e = L.e; mg_bottom.lv[k].rho = L.rho; mg_bottom.lv[k].tmp = L.tmp;
      bottom_bytes += 3 * (size_t)(L.mx + 2) * (size_t)(L.my + 2) * sizeof(R);
    }
    // CFD_MG_BOTTOM_SMEM=1 (A/B): the bottom levels' fields in shared memory when they fit (k_mg_bottom<kSmem>).  Measured
    // (r2v): 64.0 us against 66.0 us per launch — the single block is bound by instruction issue (1024 threads x ~300
    // instructions per phase on ONE SM), not by the L2 round trips, so the default stays the global-memory form.
    static const bool bottom_smem = getenv("CFD_MG_BOTTOM_SMEM") != nullptr;
    mg_bottom_smem = 0;
    if (mg_bottom_level >= 1 && bottom_bytes <= (size_t)200 * 1024 && bottom_smem) {
      mg_bottom_smem = bottom_bytes;
      CFD_CUDA(cudaFuncSetAttribute(cfdk::k_mg_bottom<R, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bottom_bytes));
    }
    // each level's descending / ascending leg as one launch (cfd_mg_legs.cuh): V(nu,nu) with nu = 2, 3, 4
    static const bool no_legs = getenv("CFD_MG_NO_LEGS") != nullptr;
    const int nu_legs = mg_smoothing();
    mg_legs = nu_legs >= 2 && nu_legs <= 4 && !(opt.flags & CFD_FLAG_MG_UNFUSED) && mg.size() > 1 && !no_legs;
    for (int l = 1; l <= mg_ld && mg_legs; ++l) {  // strips: the legs exchange leg_rho_halo() rows of a level's rho
      if (!mg[(size_t)l].dev.diag_table) mg_legs = false;
      for (int r = 0; r < world; ++r)
        if (lvl_hi(l, r) - lvl_lo(l, r) < leg_rho_halo() + 1) mg_legs = false;
    }
    if (world > 1 && (nu_legs > kHalo || jb - ja < 2 * kHalo)) mg_legs = false;  // level 0: nu halo rows of rho and x_nu
    if (mg_legs) {
      if (nu_legs == 2) rc = leg_attributes<2>();
      else if (nu_legs == 3) rc = leg_attributes<3>();
      else rc = leg_attributes<4>();
      if (rc) return rc;
    }
    CFD_CUDA(cudaHostAlloc((void**)&h_mg, sizeof(cfdk::MgScalars), cudaHostAllocMapped));
    CFD_CUDA(cudaStreamSynchronize(stream));
    // strips over peer memory: everything the V-cycle exchanges
    peer_register(mg_rho.base);
    for (auto& f : mg_b) peer_register(f.base);
    for (auto& f : mg_hist) peer_register(f.base);
    for (int l = 1; l <= mg_ld; ++l) { peer_register(mg[(size_t)l].e); peer_register(mg[(size_t)l].tmp); peer_register(mg[(size_t)l].rho); }
    if (world > 1 && mg_ld + 1 < (int)mg.size()) peer_register(mg[(size_t)mg_ld + 1].rho, true);
    if ((rc = peer_sync())) return rc;
    return CFD_OK;
  }

  int mg_smoothing() const { return opt.consts.mg_smoothing < 1 ? 1 : opt.consts.mg_smoothing; }

  // ---- strips (world > 1): level 0 lives in row strips like every other field; coarse levels are allocated in full
  // on every rank.  Levels 1..mg_ld are computed in strips too (each rank its own rows, one halo row exchanged after
  // every sweep); level mg_ld + 1 is gathered and everything below runs replicated on every rank (identical values,
  // no further communication).  The aligned partition (strip_row_start) makes the cells of a strip pair up within
  // the strip on levels 0..3.
  int mg_ld = -1;  // last coarse level computed in strips (-1: none / single domain)
  // owned rows [lo, hi) of level l, in unknown-row numbering of that level, for rank r
  int lvl_lo(int l, int r) const {
    const int a = strip_row_start(ny, world, r);
    return ((a > 1 ? a : 1) - 1) >> l;
  }
  int lvl_hi(int l, int r) const {
    if (r == world - 1) return mg[(size_t)l].my;
    return (strip_row_start(ny, world, r + 1) - 1) >> l;
  }
  bool lvl_dist(int l) const { return world > 1 && l <= mg_ld; }

  // one halo row each way of a coarse-level field (row pitch mx + 2; unknown row J is array row J + 1)
  int exchange_level(R* f, int l, int d = 1) {
    const size_t pitch = (size_t)mg[(size_t)l].mx + 2;
    const int lo = lvl_lo(l, rank), hi = lvl_hi(l, rank);
    const size_t n = (size_t)d * pitch;  // d rows each way (every strip owns at least d rows of the level: mg_setup checks)
    if (const PeerBuf* pb = peer_find(f)) {
      // coarse arrays are allocated in full on every rank: same offsets everywhere.  Down: my first d owned rows (array rows
      // lo + 1 ...) are the lower rank's upper halo rows; up: my last d owned rows (... hi) are the upper rank's lower halo rows
      R* dst_down = rank > 0 ? (R*)pb->map[rank - 1] + (size_t)(lo + 1) * pitch : nullptr;
      R* dst_up = rank < world - 1 ? (R*)pb->map[rank + 1] + (size_t)(hi - d + 1) * pitch : nullptr;
      return peer_push(f + (size_t)(lo + 1) * pitch, dst_down, n * sizeof(R), f + (size_t)(hi - d + 1) * pitch, dst_up, n * sizeof(R));
    }
    NcclGroupGuard group;
    CFD_NCCL(group.begin());
    if (rank > 0) {
      CFD_NCCL(nccl_api().Send(f + (size_t)(lo + 1) * pitch, n, nccl_real(), rank - 1, comm, stream));
      CFD_NCCL(nccl_api().Recv(f + (size_t)(lo - d + 1) * pitch, n, nccl_real(), rank - 1, comm, stream));
    }
    if (rank < world - 1) {
      CFD_NCCL(nccl_api().Send(f + (size_t)(hi - d + 1) * pitch, n, nccl_real(), rank + 1, comm, stream));
      CFD_NCCL(nccl_api().Recv(f + (size_t)(hi + 1) * pitch, n, nccl_real(), rank + 1, comm, stream));
    }
    CFD_NCCL(group.end());
    return CFD_OK;
  }
  // every rank receives every rank's rows of a level-l field (one in-place broadcast per owner, grouped)
  int gather_level(R* f, int l) {
    const size_t pitch = (size_t)mg[(size_t)l].mx + 2;
    if (const PeerBuf* pb = peer_find(f)) {
      if (!pb->want_all) return fail(CFD_ERR_UNSUPPORTED, "gather_level: buffer is not mapped on every rank");
      const int lo = lvl_lo(l, rank), hi = lvl_hi(l, rank);
      cfdk::PeerGather g;
      memset(&g, 0, sizeof g);
      const size_t off = (size_t)(lo + 1) * pitch;
      g.src = (const uint32_t*)(f + off);
      for (int r = 0; r < world; ++r) g.dst[r] = r == rank ? nullptr : (uint32_t*)((R*)pb->map[r] + off);
      g.words = hi > lo ? (size_t)(hi - lo) * pitch * sizeof(R) / 4 : 0;
      g.all = peer_all();
      g.ticket = peer_ticket + 1;
      g.seq = ++gseq;
      int grid = (int)((g.words + 2047) / 2048);
      if (grid < 1) grid = 1;
      if (grid > 128) grid = 128;
      cfdk::k_peer_gather<<<grid, 256, 0, stream>>>(g, nullptr);
      ++launches;
      CFD_CUDA(cudaGetLastError());
      return CFD_OK;
    }
    NcclGroupGuard group;
    CFD_NCCL(group.begin());
    for (int r = 0; r < world; ++r) {
      const int lo = lvl_lo(l, r), hi = lvl_hi(l, r);
      if (hi <= lo) continue;
      R* rows = f + (size_t)(lo + 1) * pitch;
      CFD_NCCL(nccl_api().Broadcast(rows, rows, (size_t)(hi - lo) * pitch, nccl_real(), r, comm, stream));
    }
    CFD_NCCL(group.end());
    return CFD_OK;
  }
  // strips: finish a dot product whose rank-local sum sits in mg_scalars->local_sum
  int mg_finish_strips(const cfdk::MgFine<R>& c, int mode) {
    int rc;
    if ((rc = mg_reduce_strips())) return rc;
    return mg_advance_strips(c, mode);
  }
  // the two halves of mg_finish_strips: the sum over the ranks (NCCL transport: an 8-byte allreduce, which may sit in one
  // grouped launch with a halo exchange, see nccl_batch) and the advance of the CG scalars (peer transport: one kernel does both)
  int mg_reduce_strips() {
    if (world == 1 || peer2_ready) return CFD_OK;
    CFD_NCCL(nccl_api().AllReduce(&mg_scalars->local_sum, &mg_scalars->local_sum, 1, ncclFloat64, ncclSum, comm, stream));
    return CFD_OK;
  }
  int mg_advance_strips(const cfdk::MgFine<R>& c, int mode) {
    if (world == 1) return CFD_OK;
    if (peer2_ready) cfdk::k_mg_advance_peer<R><<<1, 32, 0, stream>>>(peer_all(), c, mg_scalars, mode, ++rseq);
    else cfdk::k_mg_advance<R><<<1, 32, 0, stream>>>(c, mg_scalars, mode);
    ++launches;
    CFD_CUDA(cudaGetLastError());
    return CFD_OK;
  }
  // NCCL transport: the exchanges enqueued while the returned guard is open form ONE grouped launch (halo rows of two fields,
  // or halo rows together with the allreduce of a dot product): one latency on the critical path instead of two.  Peer
  // transport: nothing to batch (every exchange is its own small kernel; the guard stays closed).
  int nccl_batch(NcclGroupGuard* g) {
    if (world > 1 && !peer2_ready && !peer_ready) CFD_NCCL(g->begin());
    return CFD_OK;
  }
  int nccl_batch_end(NcclGroupGuard* g) {
    if (g->open) CFD_NCCL(g->end());
    return CFD_OK;
  }
  // strips with the leg kernels: the descending leg of level 0 reads nu halo rows of rho
  int exchange_rho_for_legs() { return mg_legs ? exchange_halo(mg_rho, ja, jb, mg_smoothing()) : CFD_OK; }

  // ---- the V-cycle with one launch per leg (cfd_mg_legs.cuh) ----
  template <int NU>
  int leg_attributes() {
    using T = LegTile<NU>;
    CFD_CUDA(cudaFuncSetAttribute(cfdk::k_mg0_down3<R, NU>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(cfdk::Leg3Smem<R>)));
    CFD_CUDA(cudaFuncSetAttribute(cfdk::k_mg0_up3<R, NU>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(cfdk::Leg3Smem<R>)));
    CFD_CUDA(cudaFuncSetAttribute(cfdk::k_mgc_down3<R, NU>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(cfdk::Leg3SmemC<R>)));
    CFD_CUDA(cudaFuncSetAttribute(cfdk::k_mgc_up3<R, NU>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(cfdk::Leg3SmemC<R>)));
    // (the default carve-out left room for 3 blocks of 37 KB per SM: ask for the largest shared-memory partition, the legs
    // do not use L1 for anything that matters)
    CFD_CUDA(cudaFuncSetAttribute(cfdk::k_mg0_down3<R, NU>, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared));
    CFD_CUDA(cudaFuncSetAttribute(cfdk::k_mg0_up3<R, NU>, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared));
    CFD_CUDA(cudaFuncSetAttribute(cfdk::k_mgc_down3<R, NU>, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared));
    CFD_CUDA(cudaFuncSetAttribute(cfdk::k_mgc_up3<R, NU>, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared));
    CFD_CUDA(cudaFuncSetAttribute(cfdk::k_mg0_down2<R, NU>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(cfdk::Leg0Smem<R, NU>)));
    CFD_CUDA(cudaFuncSetAttribute(cfdk::k_mg0_up2<R, NU>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(cfdk::Leg0Smem<R, NU>)));
    CFD_CUDA(cudaFuncSetAttribute(cfdk::k_mg0_down<R, T::TX, T::TY, NU>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(cfdk::LegSmem<R, T::TX, T::TY, NU>)));
    CFD_CUDA(cudaFuncSetAttribute(cfdk::k_mg0_up<R, T::TX, T::TY, NU>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(cfdk::LegSmem<R, T::TX, T::TY, NU>)));
    CFD_CUDA(cudaFuncSetAttribute(cfdk::k_mgc_down<R, T::TX, T::TY, NU>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(cfdk::LegSmemC<R, T::TX, T::TY, NU>)));
    CFD_CUDA(cudaFuncSetAttribute(cfdk::k_mgc_up<R, T::TX, T::TY, NU>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(cfdk::LegSmemC<R, T::TX, T::TY, NU>)));
    CFD_CUDA(cudaFuncSetAttribute(cfdk::k_mgc_down<R, kLegSX, kLegSY, NU>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(cfdk::LegSmemC<R, kLegSX, kLegSY, NU>)));
    CFD_CUDA(cudaFuncSetAttribute(cfdk::k_mgc_up<R, kLegSX, kLegSY, NU>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(cfdk::LegSmemC<R, kLegSX, kLegSY, NU>)));
    return CFD_OK;
  }

  // level l >= 1 (a level with a diagonal table).  Strips: x_nu is also recomputed on NU rows beyond the owned ones (from
  // the 2 NU - 1 halo rows of rho exchanged by the level above), so the ascending leg needs only the parents' halo row
  template <int NU>
  int leg_coarse_vcycle(int l) {
    using T = LegTile<NU>;
    MgLevelHost& L = mg[(size_t)l];
    MgLevelHost& C = mg[(size_t)l + 1];
    const R omega = R(opt.consts.mg_omega);
    const bool dist = lvl_dist(l);
    const int lo = dist ? lvl_lo(l, rank) : 0, hi = dist ? lvl_hi(l, rank) : L.my;
    const bool big = (long)L.mx * L.my >= 500000L;
    const int tx = big ? T::TX : kLegSX, ty = big ? T::TY : kLegSY;
    // (the tiles pair rows 2 J, 2 J + 1 for the restriction: the first row of the launch must be even)
    const int x_lo = dist ? (lo >= NU ? (lo - NU) & ~1 : 0) : 0, x_hi = dist ? (hi + NU <= L.my ? hi + NU : L.my) : L.my;
    const int c_lo = dist ? lvl_lo(l + 1, rank) : 0, c_hi = dist ? lvl_hi(l + 1, rank) : C.my;
    const dim3 g_dn((L.mx + tx - 1) / tx, (x_hi - x_lo + ty - 1) / ty), g_up((L.mx + tx - 1) / tx, (hi - lo + ty - 1) / ty);
    int rc;
    const bool tiled = leg_form() == 3;
    using G3 = cfdk::Leg3<NU>;
    const dim3 g_dn3((L.mx + G3::TX - 1) / G3::TX, (x_hi - x_lo + G3::TY - 1) / G3::TY), g_up3((L.mx + G3::TX - 1) / G3::TX, (hi - lo + G3::TY - 1) / G3::TY);
    if (tiled) cfdk::k_mgc_down3<R, NU><<<g_dn3, cfdk::kLegThreads, sizeof(cfdk::Leg3SmemC<R>), stream>>>(L.dev, L.rho, L.e, C.mx, C.rho, omega, x_lo, x_hi, c_lo, c_hi, mg_scalars);
    else if (big) cfdk::k_mgc_down<R, T::TX, T::TY, NU><<<g_dn, cfdk::kLegThreads, sizeof(cfdk::LegSmemC<R, T::TX, T::TY, NU>), stream>>>(L.dev, L.rho, L.e, C.mx, C.rho, omega, x_lo, x_hi, c_lo, c_hi, mg_scalars);
    else cfdk::k_mgc_down<R, kLegSX, kLegSY, NU><<<g_dn, cfdk::kLegThreads, sizeof(cfdk::LegSmemC<R, kLegSX, kLegSY, NU>), stream>>>(L.dev, L.rho, L.e, C.mx, C.rho, omega, x_lo, x_hi, c_lo, c_hi, mg_scalars);
    ++launches;
    if (dist) {
      if (lvl_dist(l + 1)) { if ((rc = exchange_level(C.rho, l + 1, leg_rho_halo()))) return rc; }
      else if ((rc = gather_level(C.rho, l + 1))) return rc;
    }
    if ((rc = mg_coarse_vcycle(l + 1))) return rc;
    if (tiled) cfdk::k_mgc_up3<R, NU><<<g_up3, cfdk::kLegThreads, sizeof(cfdk::Leg3SmemC<R>), stream>>>(L.dev, L.e, L.rho, C.mx, C.cur, L.tmp, omega, lo, hi, mg_scalars);
    else if (big) cfdk::k_mgc_up<R, T::TX, T::TY, NU><<<g_up, cfdk::kLegThreads, sizeof(cfdk::LegSmemC<R, T::TX, T::TY, NU>), stream>>>(L.dev, L.e, L.rho, C.mx, C.cur, L.tmp, omega, lo, hi, mg_scalars);
    else cfdk::k_mgc_up<R, kLegSX, kLegSY, NU><<<g_up, cfdk::kLegThreads, sizeof(cfdk::LegSmemC<R, kLegSX, kLegSY, NU>), stream>>>(L.dev, L.e, L.rho, C.mx, C.cur, L.tmp, omega, lo, hi, mg_scalars);
    ++launches;
    L.cur = L.tmp;
    // the finer level's ascending leg reads the parents of its NU halo rows: (NU + 1) / 2 halo rows of this correction
    if (dist && (rc = exchange_level(L.tmp, l, (NU + 1) / 2))) return rc;
    return CFD_OK;
  }

  // level 0: z <- V-cycle(rho); *zc / *zo = current / other smoothing buffer (mg_b indices), z ends up in *zc.  Strips: the
  // descending leg recomputes x_k on the rows beyond the owned ones it still needs (NU halo rows of rho); the ascending leg
  // reads NU halo rows of x_nu and one of the parents' correction
  template <int NU>
  int leg_fine_vcycle(const cfdk::MgFine<R>& c, const cfdk::JacobiConsts2<R>& c2, int* zc, int* zo) {
    using T = LegTile<NU>;
    MgLevelHost& C = mg[1];
    const int rows = c.row_hi - c.row_lo;
    const dim3 g_leg((nx - 2 + T::TX - 1) / T::TX, (rows + T::TY - 1) / T::TY);
    const size_t leg_bytes = sizeof(cfdk::LegSmem<R, T::TX, T::TY, NU>);
    int rc;
    // register-tiled kernels (k_mg0_down3 / _up3); CFD_MG_LEGS=2: column strips, 1: flat-indexed (A/B, cross-check)
    const int form = leg_form();
    using G = cfdk::Leg0<NU>;
    using G3 = cfdk::Leg3<NU>;
    const dim3 g_leg2((nx - 2 + G::TX - 1) / G::TX, (rows + G::TY - 1) / G::TY);
    const dim3 g_leg3((nx - 2 + G3::TX - 1) / G3::TX, (rows + G3::TY - 1) / G3::TY);
    const size_t leg2_bytes = sizeof(cfdk::Leg0Smem<R, NU>), leg3_bytes = sizeof(cfdk::Leg3Smem<R>);
    const int n_partials = form == 1 ? (int)(g_leg.x * g_leg.y) : form == 2 ? (int)(g_leg2.x * g_leg2.y) : (int)(g_leg3.x * g_leg3.y);
    // (strips: the NU halo rows of rho were exchanged together with the reduction that followed the kernel that wrote rho)
    if (form == 1) cfdk::k_mg0_down<R, T::TX, T::TY, NU><<<g_leg, cfdk::kLegThreads, leg_bytes, stream>>>(c, c2, mg_rho.v, mg_b[*zo].v, C.mx, C.rho, mg_scalars);
    else if (form == 2) cfdk::k_mg0_down2<R, NU><<<g_leg2, cfdk::kLegThreads, leg2_bytes, stream>>>(c, c2, mg_rho.v, mg_b[*zo].v, C.mx, C.rho, mg_scalars);
    else cfdk::k_mg0_down3<R, NU><<<g_leg3, cfdk::kLegThreads, leg3_bytes, stream>>>(c, c2, mg_rho.v, mg_b[*zo].v, C.mx, C.rho, mg_scalars);
    ++launches;
    std::swap(*zc, *zo);
    if (world > 1) {
      NcclGroupGuard batch;  // x_nu's halo rows and the level-1 rho in one grouped launch
      if ((rc = nccl_batch(&batch))) return rc;
      if ((rc = exchange_halo(mg_b[*zc], ja, jb, NU))) return rc;
      if (lvl_dist(1)) { if ((rc = exchange_level(C.rho, 1, leg_rho_halo()))) return rc; }
      else if ((rc = gather_level(C.rho, 1))) return rc;
      if ((rc = nccl_batch_end(&batch))) return rc;
    }
    if ((rc = mg_coarse_vcycle(1))) return rc;
    auto prof_mark = [&]() {  // CUDA-event pair around the ascending leg (bench.py's roofline kernel)
      if (!prof_smoother) return;
      if (ev_prof_used + 1 > ev_prof.size()) {
        const size_t old_n = ev_prof.size();
        ev_prof.resize(old_n + 64, nullptr);
        for (size_t k = old_n; k < ev_prof.size(); ++k) cudaEventCreate(&ev_prof[k]);
      }
      cudaEventRecord(ev_prof[ev_prof_used++], stream);
    };
    prof_mark();
    if (form == 1) cfdk::k_mg0_up<R, T::TX, T::TY, NU><<<g_leg, cfdk::kLegThreads, leg_bytes, stream>>>(c, c2, mg_b[*zc].v, mg_rho.v, C.mx, C.cur, mg_b[*zo].v, mg_partials, mg_scalars);
    else if (form == 2) cfdk::k_mg0_up2<R, NU><<<g_leg2, cfdk::kLegThreads, leg2_bytes, stream>>>(c, c2, mg_b[*zc].v, mg_rho.v, C.mx, C.cur, mg_b[*zo].v, mg_partials, mg_scalars);
    else cfdk::k_mg0_up3<R, NU><<<g_leg3, cfdk::kLegThreads, leg3_bytes, stream>>>(c, c2, mg_b[*zc].v, mg_rho.v, C.mx, C.cur, mg_b[*zo].v, mg_partials, mg_scalars);
    prof_mark();
    cfdk::k_mg_reduce<R><<<1, 1024, 0, stream>>>(c, mg_scalars, mg_partials, n_partials, 1);
    launches += 2;
    std::swap(*zc, *zo);
    {
      NcclGroupGuard batch;  // strips: the neighbours' edge rows of z and the sum of rho.z in one grouped launch
      if ((rc = nccl_batch(&batch))) return rc;
      if ((rc = exchange_halo(mg_b[*zc], ja, jb, 1))) return rc;
      if ((rc = mg_reduce_strips())) return rc;
      if ((rc = nccl_batch_end(&batch))) return rc;
    }
    if ((rc = mg_advance_strips(c, 1))) return rc;
    CFD_CUDA(cudaGetLastError());
    return CFD_OK;
  }

  // correction of level l >= 1 from its rho (result in mg[l].cur)
  int mg_coarse_vcycle(int l) {
    MgLevelHost& L = mg[(size_t)l];
    const R omega = R(opt.consts.mg_omega);
    const int nu_s = mg_smoothing();
    const bool dist = lvl_dist(l);
    const int lo = dist ? lvl_lo(l, rank) : 0, hi = dist ? lvl_hi(l, rank) : L.my;
    const dim3 blk(cfdk::kMgThreads), grd((L.mx + cfdk::kMgThreads - 1) / cfdk::kMgThreads, hi - lo);
    R *a = L.e, *b = L.tmp;
    int rc;
    if (l == mg_bottom_level && !(opt.flags & CFD_FLAG_MG_NO_BOTTOM_KERNEL) && !(L.mx == 1 && L.my == 1)) {
      if (mg_bottom_smem > 0) cfdk::k_mg_bottom<R, true><<<1, cfdk::kMgBottomThreads, mg_bottom_smem, stream>>>(mg_bottom, mg_scalars);
      else cfdk::k_mg_bottom<R, false><<<1, cfdk::kMgBottomThreads, 0, stream>>>(mg_bottom, mg_scalars);
      ++launches;
      L.cur = L.e;
      return CFD_OK;
    }
    if (L.mx == 1 && L.my == 1) {  // exact
      cfdk::k_mgc_sweep<R><<<grd, blk, 0, stream>>>(L.dev, a, L.rho, b, R(1), 1, 0, mg_scalars);
      ++launches;
      L.cur = b;
      return CFD_OK;
    }
    if (mg_legs && L.dev.diag_table != nullptr) {
      if (nu_s == 2) return leg_coarse_vcycle<2>(l);
      if (nu_s == 3) return leg_coarse_vcycle<3>(l);
      return leg_coarse_vcycle<4>(l);
    }
    // (the 4-row tiles leave the small levels with too few blocks: there the one-cell-per-thread kernel is faster)
    const bool tab = L.dev.diag_table != nullptr && !(opt.flags & CFD_FLAG_MG_UNFUSED) && (long)L.mx * L.my >= 500000L;
    const dim3 grd_t((L.mx + cfdk::kMgThreads - 1) / cfdk::kMgThreads, (hi - lo + cfdk::kMgcRows - 1) / cfdk::kMgcRows);
    for (int s = 0; s < nu_s; ++s) {
      if (tab) cfdk::k_mgc_sweep_tab<R><<<grd_t, blk, 0, stream>>>(L.dev, a, L.rho, b, omega, s == 0 ? 1 : 0, lo, hi, mg_scalars);
      else cfdk::k_mgc_sweep<R><<<grd, blk, 0, stream>>>(L.dev, a, L.rho, b, omega, s == 0 ? 1 : 0, lo, mg_scalars);
      ++launches;
      std::swap(a, b);
      if (dist && (rc = exchange_level(a, l))) return rc;
    }
    MgLevelHost& C = mg[(size_t)l + 1];
    const int c_lo = dist ? lvl_lo(l + 1, rank) : 0, c_hi = dist ? lvl_hi(l + 1, rank) : C.my;
    const dim3 grd_c((C.mx + cfdk::kMgThreads - 1) / cfdk::kMgThreads, c_hi - c_lo);
    cfdk::k_mgc_restrict<R><<<grd_c, blk, 0, stream>>>(L.dev, a, L.rho, C.mx, c_lo, C.rho, mg_scalars);
    ++launches;
    if (dist && !lvl_dist(l + 1) && (rc = gather_level(C.rho, l + 1))) return rc;
    if ((rc = mg_coarse_vcycle(l + 1))) return rc;
    {
      // strips: also correct the neighbours' edge rows (halo), from the parent's halo rows
      const int p_lo = dist ? (lo > 0 ? lo - 1 : 0) : 0, p_hi = dist ? (hi < L.my ? hi + 1 : L.my) : L.my;
      const dim3 grd_p((L.mx + cfdk::kMgThreads - 1) / cfdk::kMgThreads, p_hi - p_lo);
      cfdk::k_mgc_prolong<R><<<grd_p, blk, 0, stream>>>(L.mx, a, C.mx, C.cur, p_lo, mg_scalars);
      ++launches;
    }
    for (int s = 0; s < nu_s; ++s) {
      if (tab) cfdk::k_mgc_sweep_tab<R><<<grd_t, blk, 0, stream>>>(L.dev, a, L.rho, b, omega, 0, lo, hi, mg_scalars);
      else cfdk::k_mgc_sweep<R><<<grd, blk, 0, stream>>>(L.dev, a, L.rho, b, omega, 0, lo, mg_scalars);
      ++launches;
      std::swap(a, b);
      if (dist && (rc = exchange_level(a, l))) return rc;
    }
    L.cur = a;
    return CFD_OK;
  }

  cfdk::MgFine<R> mg_fine(R dt_sub) const {
    cfdk::MgFine<R> c;
    c.dx_sq = dx * dx; c.dy_sq = dy * dy; c.dt = dt_sub; c.tol = R(opt.consts.cg_tolerance);
    c.ddx_sq = div_dx_sq; c.ddy_sq = div_dy_sq;
    c.n_unknowns = R((size_t)(nx - 2) * (size_t)(ny - 2));
    c.nx = nx; c.ny = ny; c.cavity = scenario == CFD_SCENARIO_CAVITY;
    c.row_lo = sweep_row_begin(); c.row_hi = sweep_row_end();
    c.init_lo = ja; c.init_hi = jb;
    c.defer = world > 1 ? 1 : 0;
    return c;
  }

  // z <- V-cycle(rho), smoothing between the two mg_b buffers that do not hold d; the last smoothing sweep also
  // accumulates rho.z (-> beta).  Returns the index of the buffer holding z.
  int mg_precondition(const cfdk::MgFine<R>& c, int* z_index) {
    const int nu_s = mg_smoothing();
    cfdk::JacobiConsts2<R> c2;
    c2.dx_sq = div_dx_sq; c2.dy_sq = div_dy_sq; c2.denom = div_denom;
    c2.omega = R(opt.consts.mg_omega);
    c2.one_minus_omega = R(1.0) - c2.omega;
    c2.tol = R(0);
    c2.nx = nx; c2.ny = ny; c2.cavity = c.cavity; c2.rows_per_block = sweep_rows_per_block;
    c2.row_begin = c.row_lo; c2.row_end = c.row_hi; c2.row_shift = ja - kHalo;
    c2.check_lag = 1;
    c2.fix_pass = -1;
    const int rows = c.row_hi - c.row_lo;
    const int kSweepThreads = cfdk::kSweepWarps * 32;
    const dim3 blk2(kSweepThreads), grd2((nx / 2 + kSweepThreads - 1) / kSweepThreads, (rows + sweep_rows_per_block - 1) / sweep_rows_per_block);
    const size_t ring_bytes = sizeof(cfdk::SweepChunkRing<R>);
    const int za = (mg_id + 1) % 3, zb = (mg_id + 2) % 3;
    int zc = za, zo = zb;  // current / other smoothing buffer
    int rc;
    cfdk::SweepDot<R> dot;
    dot.c = c; dot.sc = mg_scalars; dot.partials = mg_partials; dot.ticket = mg_ticket;
    auto smooth = [&](bool with_dot) -> int {
      if (prof_smoother) {
        if (ev_prof_used + 2 > ev_prof.size()) {
          ev_prof.resize(ev_prof_used + 64, nullptr);
          for (size_t k = ev_prof_used; k < ev_prof.size(); ++k) cudaEventCreate(&ev_prof[k]);
        }
        cudaEventRecord(ev_prof[ev_prof_used], stream);
      }
      if (with_dot)
        cfdk::k_jacobi_sweep5<R, true><<<grd2, blk2, ring_bytes, stream>>>(c2, tmap_mg_b[zc], tmap_mg_rho, mg_b[zo].v, mg_err, 0,
                                                                           cfdk::SweepPeer<R>{}, dot);
      else
        cfdk::k_jacobi_sweep5<R><<<grd2, blk2, ring_bytes, stream>>>(c2, tmap_mg_b[zc], tmap_mg_rho, mg_b[zo].v, mg_err, 0,
                                                                      cfdk::SweepPeer<R>{}, dot);
      if (prof_smoother) {
        cudaEventRecord(ev_prof[ev_prof_used + 1], stream);
        ev_prof_used += 2;
      }
      ++launches;
      if (with_dot) {  // rho.z from the sweep's per-block partials (-> beta)
        cfdk::k_mg_reduce<R><<<1, 1024, 0, stream>>>(c, mg_scalars, mg_partials, (int)(grd2.x * grd2.y), 1);
        ++launches;
      }
      std::swap(zc, zo);
      return exchange_halo(mg_b[zc], ja, jb, 1);  // strips: the neighbours' new edge rows (no-op on one GPU)
    };
    // V(nu, nu) with nu >= 2: the first two pre-smoothing sweeps are one pass over rho, and the prolongation is folded
    // into the first post-smoothing sweep (k_mg_fused_sweep; bit-identical to the separate kernels, which
    // CFD_FLAG_MG_UNFUSED keeps for the cross-check)
    if (mg_legs) {
      if (nu_s == 2) rc = leg_fine_vcycle<2>(c, c2, &zc, &zo);
      else if (nu_s == 3) rc = leg_fine_vcycle<3>(c, c2, &zc, &zo);
      else rc = leg_fine_vcycle<4>(c, c2, &zc, &zo);
      if (rc) return rc;
      *z_index = zc;
      return CFD_OK;
    }
    const bool fused = nu_s >= 2 && !(opt.flags & CFD_FLAG_MG_UNFUSED);
    const dim3 g_fs((nx + 2 * cfdk::kMgThreads - 1) / (2 * cfdk::kMgThreads), (rows + cfdk::kFsRows - 1) / cfdk::kFsRows);
    if (fused) {
      if ((rc = exchange_halo(mg_rho, ja, jb, 1))) return rc;  // strips: the stencil of the second sweep reaches into rho's halo
      // the tensor-TMA sweep kernel in its kFirst form (input formed from the staged rho); CFD_FUSED0_SIMPLE=1: the
      // register-tile kernel k_mg_fused_sweep<0> instead (A/B)
      static const bool simple0 = getenv("CFD_FUSED0_SIMPLE") != nullptr;
      if (simple0)
        cfdk::k_mg_fused_sweep<R, 0><<<g_fs, cfdk::kMgThreads, 0, stream>>>(c, c2, mg_rho.v, nullptr, 0, nullptr, mg_b[zo].v, mg_scalars);
      else
        cfdk::k_jacobi_sweep5<R, false, true><<<grd2, blk2, ring_bytes, stream>>>(c2, tmap_mg_rho_halo, tmap_mg_rho, mg_b[zo].v, mg_err, 0,
                                                                                 cfdk::SweepPeer<R>{}, dot);
      ++launches;
      std::swap(zc, zo);
      if ((rc = exchange_halo(mg_b[zc], ja, jb, 1))) return rc;
    } else {
      // first sweep from z = 0: pointwise (k_mg_first_sweep) instead of a stencil sweep over a zero field
      const dim3 g_vec((nx + 2 * cfdk::kMgThreads - 1) / (2 * cfdk::kMgThreads), (rows + cfdk::kMgRows - 1) / cfdk::kMgRows);
      cfdk::k_mg_first_sweep<R><<<g_vec, cfdk::kMgThreads, 0, stream>>>(c, c2.omega, c2.one_minus_omega, div_denom,
                                                                        mg_rho.v, mg_b[zc].v, mg_scalars);
      ++launches;
      if ((rc = exchange_halo(mg_b[zc], ja, jb, 1))) return rc;
    }
    for (int s = fused ? 2 : 1; s < nu_s; ++s)
      if ((rc = smooth(false))) return rc;
    bool prolong_fused = false;
    if (mg.size() > 1) {
      MgLevelHost& C = mg[1];
      const int c_lo = world > 1 ? lvl_lo(1, rank) : 0, c_hi = world > 1 ? lvl_hi(1, rank) : C.my;
      const dim3 blk(cfdk::kMgThreads), grd_c((C.mx + cfdk::kMgThreads - 1) / cfdk::kMgThreads, c_hi - c_lo);
      cfdk::k_mg_fine_restrict<R><<<grd_c, blk, 0, stream>>>(c, mg_b[zc].v, mg_rho.v, C.mx, c_lo, C.rho, mg_scalars);
      ++launches;
      if (world > 1 && !lvl_dist(1) && (rc = gather_level(C.rho, 1))) return rc;
      if ((rc = mg_coarse_vcycle(1))) return rc;
      if (fused) {
        // z + correction is formed on the fly by the first post-smoothing sweep (halo rows: from the parent's halo rows)
        cfdk::k_mg_fused_sweep<R, 1><<<g_fs, cfdk::kMgThreads, 0, stream>>>(c, c2, mg_rho.v, mg_b[zc].v, C.mx, C.cur, mg_b[zo].v,
                                                                            mg_scalars);
        ++launches;
        std::swap(zc, zo);
        if ((rc = exchange_halo(mg_b[zc], ja, jb, 1))) return rc;
        prolong_fused = true;
      } else {
        // strips: also correct the neighbours' edge rows (halo) of z, from the parent's halo rows
        const int p_lo = world > 1 && c.row_lo > 1 ? c.row_lo - 1 : c.row_lo;
        const int p_hi = world > 1 && c.row_hi < ny - 1 ? c.row_hi + 1 : c.row_hi;
        const dim3 g_int((nx - 2 + cfdk::kMgThreads - 1) / cfdk::kMgThreads, p_hi - p_lo);
        cfdk::k_mg_fine_prolong<R><<<g_int, blk, 0, stream>>>(c, mg_b[zc].v, C.mx, C.cur, p_lo, mg_scalars);
        ++launches;
      }
    }
    for (int s = prolong_fused ? 1 : 0; s < nu_s; ++s)
      if ((rc = smooth(s == nu_s - 1))) return rc;
    if ((rc = mg_finish_strips(c, 1))) return rc;
    CFD_CUDA(cudaGetLastError());
    *z_index = zc;
    return CFD_OK;
  }

  // ---- carried start-vector state (see the members' comment) ----
  cfdk::MgStart<R> mg_start(bool warm) const {
    cfdk::MgStart<R> g;
    g.a = g.b = g.c = nullptr;
    g.mode = 0;
    if (!warm) return g;
    if (mg_guess_explicit) { g.a = mg_guess.v; g.mode = 1; return g; }
    g.a = mg_hist[0].v; g.b = mg_hist[1].v; g.c = mg_hist[2].v;
    g.mode = opt.consts.mg_warm_start;
    return g;
  }
  void rotate_hist_names() {  // (h0, h1, h2) <- (h2's buffer, h0, h1)
    std::swap(mg_hist[1], mg_hist[2]); std::swap(tmap_hist[1], tmap_hist[2]);
    std::swap(mg_hist[0], mg_hist[1]); std::swap(tmap_hist[0], tmap_hist[1]);
  }
  // the hot path: p' is about to be overwritten by an MGCG solve -> the old p' buffer becomes h0, the retired h2 buffer p'
  void rotate_mg_by_swap() {
    if (!mg_rotate_pending) return;
    std::swap(pp[ipp], mg_hist[2]);
    std::swap(tmap_pp[ipp], tmap_hist[2]);
    rotate_hist_names();
    mg_rotate_pending = false;
  }
  // everything else (state read-back / overwrite, other solvers, peer-mapped p' buffers): same result by a copy
  int resolve_mg_rotation() {
    if (!mg_rotate_pending) return CFD_OK;
    const size_t rl = (size_t)nx, rows = (size_t)(jb - ja) + 2;  // owned rows and one halo row each side
    CFD_CUDA(cudaMemcpyAsync(mg_hist[2].row(ja - 1), pp[ipp].row(ja - 1), rows * rl * sizeof(R), cudaMemcpyDeviceToDevice, stream));
    rotate_hist_names();
    mg_rotate_pending = false;
    return CFD_OK;
  }
  int materialize_pp_zero() {
    if (!pp_zero_pending) return CFD_OK;
    int rc;
    if ((rc = resolve_mg_rotation())) return rc;  // pp[ipp] may still hold the last first-solve's result
    CFD_CUDA(cudaMemsetAsync(pp[ipp].row(ja - 1), 0, ((size_t)(jb - ja) + 2) * (size_t)nx * sizeof(R), stream));
    pp_zero_pending = false;
    return CFD_OK;
  }
  // the complete star fields, should anyone ask for them while they alias the current ones
  int materialize_star() {
    if (!star_alias) return CFD_OK;
    dim3 blk(256), grd((nx + 1 + 255) / 256, v_row_end() - ja);
    cfdk::k_star_materialize<R><<<grd, blk, 0, stream>>>(nx, ny, solid.v, ubuf[iu].v, vbuf[iu].v, ubuf[ius].v, vbuf[ius].v, ja, jb,
                                                         v_row_end());
    CFD_CUDA(cudaGetLastError());
    star_alias = false;
    return CFD_OK;
  }

  // set-up shared by every MGCG solve, before the divergence kernel (which may already accumulate rho.rho)
  void note_first_solve(double bb, double rel, int iterations, R dt_sub) {
    step_bb = bb;
    last_p_rel = rel;
    const R n_unknowns = R((size_t)(nx - 2) * (size_t)(ny - 2));
    last_rhs_rms = (double)(dt_sub * (R)std::sqrt((double)((R)bb / n_unknowns)));
    last_first_solve_iterations = (uint64_t)iterations;
  }

  int mgcg_begin(bool first_solve) {
    int rc;
    if (mg.empty() && (rc = mg_setup())) return rc;
    cfdk::MgScalars init;
    memset(&init, 0, sizeof init);
    init.max_iterations = opt.consts.cg_max_iterations;
    init.relative = opt.consts.cg_relative;
    init.bb = first_solve ? 0.0 : step_bb;  // a first solve's own ||rhs||^2 comes out of its divergence kernel
    *h_mg = init;
    CFD_CUDA(cudaMemcpyAsync(mg_scalars, h_mg, sizeof init, cudaMemcpyHostToDevice, stream));
    return CFD_OK;
  }

  // the finish of a dot product whose kernel left one partial per block (mg_finish_launch): sum in index order, advance the
  // CG scalars (strips: leave the rank's sum for the exchange) — one single-block launch instead of a ticket tail per block
  unsigned* dot_ticket() const { return mg_finish_launch ? nullptr : mg_ticket; }
  void dot_finish(const cfdk::MgFine<R>& c, const dim3& grid, int mode) {
    if (!mg_finish_launch) return;
    cfdk::k_mg_reduce<R><<<1, 1024, 0, stream>>>(c, mg_scalars, mg_partials, (int)(grid.x * grid.y), mode);
    ++launches;
  }

  int mgcg_solve(R dt_sub, int call_index, R* residual_out, bool decided_early, bool* elided, size_t ev) {
    int rc;
    const cfdk::MgFine<R> c = mg_fine(dt_sub);
    const dim3 blk(cfdk::kMgThreads);
    const unsigned gx = (unsigned)((nx + 2 * cfdk::kMgThreads - 1) / (2 * cfdk::kMgThreads));
    const int rows = c.row_hi - c.row_lo;
    const dim3 g_all(gx, (jb - ja + init_rows - 1) / init_rows);                    // every owned row (init)
    // owned rows of unknowns in tiles of dir_rows / upd_rows rows x 2 * threads columns
    const dim3 b_dir(dir_threads), g_dir((unsigned)((nx + 2 * dir_threads - 1) / (2 * dir_threads)), (rows + dir_rows - 1) / dir_rows);
    const dim3 b_upd(upd_threads), g_upd((unsigned)((nx + 2 * upd_threads - 1) / (2 * upd_threads)), (rows + upd_rows - 1) / upd_rows);
    const bool first_solve = call_index == 0;
    const bool warm = first_solve && opt.consts.mg_warm_start != 0;
    int& pred = mg_pred[first_solve ? 0 : 1];
    CFD_CUDA(cudaEventRecord(ev_sweep[ev], stream));
    auto read_scalars = [&]() -> int {
      publish(mg_scalars, h_mg, sizeof(cfdk::MgScalars));
      CFD_CUDA(cudaStreamSynchronize(stream));
      return CFD_OK;
    };
    if (decided_early) {
      // rho.rho came out of the divergence kernel (rho = rhs for a cold start)
      if (world > 1 && (rc = mg_finish_strips(c, 0))) return rc;
      if ((rc = read_scalars())) return rc;
      if (h_mg->done) {  // converged with p' = 0: no set-up pass, no corrector
        CFD_CUDA(cudaEventRecord(ev_sweep[ev + 1], stream));
        pp_zero_pending = true;
        pred = 0;
        last_K += 1;
        *residual_out = (R)h_mg->measure;
        *elided = true;
        return CFD_OK;
      }
    }
    // p' is overwritten from here on: the last first-solve's result moves into the history by a buffer swap
    // (peer-mapped p' buffers, Mode R's fused strips, must stay where they are: copy instead)
    if (peer_ready) { if ((rc = resolve_mg_rotation())) return rc; }
    else rotate_mg_by_swap();
    pp_zero_pending = false;
    const Field<R>& xf = pp[ipp];
    R* x = xf.v;
    R* w = pp[ipp ^ 1].v;
    // first solve of a step: start from the extrapolated history (mg_warm_start); the stencil of the start vector needs
    // the neighbours' edge rows, which every p' carries since the exchange at the end of its solve
    if (warm && mg_guess_explicit && (rc = exchange_halo(mg_guess, ja, jb, 1))) return rc;
    cfdk::k_mg_init<R><<<g_all, blk, 0, stream>>>(c, mg_scalars, rhs.v, mg_start(warm), x, mg_rho.v, mg_partials, dot_ticket(),
                                                  init_rows);
    launches += 1;
    dot_finish(c, g_all, 0);
    if (warm) mg_guess_explicit = false;
    {
      NcclGroupGuard nb;  // strips: rho.rho and (legs) the halo rows of rho the first descending leg reads
      if ((rc = nccl_batch(&nb))) return rc;
      if ((rc = mg_reduce_strips())) return rc;
      if ((rc = exchange_rho_for_legs())) return rc;
      if ((rc = nccl_batch_end(&nb))) return rc;
    }
    if ((rc = mg_advance_strips(c, 0))) return rc;
    // CG iterations are enqueued in batches of the count this solve took last time; every kernel of an iteration is a
    // no-op once the device-side `done` flag is up, so the host synchronises once per batch, not once per iteration
    int batch = pred > 0 ? pred : 1;
    if (!decided_early && pred == 0) {  // expected to converge at once: look before enqueuing a whole V-cycle
      if ((rc = read_scalars())) return rc;
      if (h_mg->done) batch = 0;
    }
    while (batch > 0) {
      for (int it = 0; it < batch; ++it) {
        int zi = 0;
        if ((rc = mg_precondition(c, &zi))) return rc;
        // rho.z (-> beta) came out of the V-cycle's last sweep; d_new goes to the smoothing buffer that is free now
        const int dn = 3 - mg_id - zi;
#define CFD_DIR_APPLY(ROWS, THREADS)                                                                                        \
  cfdk::k_mg_dir_apply<R, ROWS, THREADS><<<g_dir, b_dir, 0, stream>>>(c, mg_scalars, mg_b[zi].v, mg_b[mg_id].v, mg_b[dn].v, w, \
                                                                      mg_partials, dot_ticket())
        if (dir_rows == 2 && dir_threads == 128) CFD_DIR_APPLY(2, 128);
        else if (dir_rows == 2) CFD_DIR_APPLY(2, 256);
        else if (dir_threads == 128) CFD_DIR_APPLY(4, 128);
        else CFD_DIR_APPLY(4, 256);
#undef CFD_DIR_APPLY
        dot_finish(c, g_dir, 2);
        mg_id = dn;
        mg_last_z = zi;
        {
          NcclGroupGuard nb;  // strips: d.w and d's edge rows for the next L d
          if ((rc = nccl_batch(&nb))) return rc;
          if ((rc = mg_reduce_strips())) return rc;
          if ((rc = exchange_halo(mg_b[mg_id], ja, jb, 1))) return rc;
          if ((rc = nccl_batch_end(&nb))) return rc;
        }
        if ((rc = mg_advance_strips(c, 2))) return rc;
#define CFD_UPDATE(ROWS, THREADS)                                                                                      \
  cfdk::k_mg_update<R, ROWS, THREADS><<<g_upd, b_upd, 0, stream>>>(c, mg_scalars, mg_b[mg_id].v, w, x, mg_rho.v, mg_partials, \
                                                                   dot_ticket())
        if (upd_rows == 2 && upd_threads == 128) CFD_UPDATE(2, 128);
        else if (upd_rows == 2) CFD_UPDATE(2, 256);
        else if (upd_threads == 128) CFD_UPDATE(4, 128);
        else CFD_UPDATE(4, 256);
#undef CFD_UPDATE
        launches += 2;
        dot_finish(c, g_upd, 3);
        {
          NcclGroupGuard nb;  // strips: rho.rho and (legs) the new rho's halo rows for the next descending leg
          if ((rc = nccl_batch(&nb))) return rc;
          if ((rc = mg_reduce_strips())) return rc;
          if ((rc = exchange_rho_for_legs())) return rc;
          if ((rc = nccl_batch_end(&nb))) return rc;
        }
        if ((rc = mg_advance_strips(c, 3))) return rc;
      }
      CFD_CUDA(cudaGetLastError());
      if ((rc = read_scalars())) return rc;
      batch = h_mg->done ? 0 : 1;
    }
    pred = h_mg->iterations;
    const int n_edge = (nx > ny ? nx : ny);
    cfdk::k_cg_fill_boundary<R><<<(n_edge + 255) / 256, 256, 0, stream>>>(nx, ny, c.cavity, x, ja, jb);
    ++launches;
    if ((rc = exchange_halo(xf, ja, jb, 1))) return rc;  // the corrector reads p'[j-1] (src/model.rs:1380)
    if (first_solve) mg_rotate_pending = true;           // this p' is the newest entry of the start-vector history
    if (first_solve) note_first_solve(h_mg->bb, h_mg->rel, h_mg->iterations, dt_sub);
    CFD_CUDA(cudaEventRecord(ev_sweep[ev + 1], stream));
    CFD_CUDA(cudaGetLastError());
    last_S += (uint64_t)h_mg->iterations;
    last_K += 1;
    *residual_out = (R)h_mg->measure;
    return CFD_OK;
  }

  // with_div: single domain, MGCG — the divergence of the new u, v (the rhs of the re-correction round that follows) and
  // the partials of its rhs^2 come out of the same pass (k_corrector_div); the next pressure_solve skips its divergence
  int corrector(R dt_sub, const Field<R>& us, const Field<R>& vs, const Field<R>& uk, const Field<R>& vk,
                const Field<R>& uo, const Field<R>& vo, bool with_div = false) {
    if (with_div) {
      const int ct = corr_threads, cr = corr_rows;
      dim3 blk(ct), grd((nx + ct - 1) / ct, (ny + cr - 1) / cr);
#define CFD_CORR_DIV(ROWS, THREADS)                                                                                          \
  cfdk::k_corrector_div<R, ROWS, THREADS><<<grd, blk, 0, stream>>>(scalars(dt_sub), us.v, vs.v, uk.v, vk.v, pp[ipp].v, uo.v, \
                                                                   vo.v, p.v, rhs.v, h_divs.dx, h_divs.dy, h_divs.dt, mg_partials)
      if (cr == 8) CFD_CORR_DIV(8, 256);
      else if (cr == 1) CFD_CORR_DIV(1, 256);
      else if (cr == 4 && ct == 256) CFD_CORR_DIV(4, 256);
      else if (cr == 4) CFD_CORR_DIV(4, 128);
      else if (ct == 256) CFD_CORR_DIV(2, 256);
      else CFD_CORR_DIV(cfdk::kCorrRows, cfdk::kCorrThreads);
#undef CFD_CORR_DIV
      ++launches;
      corr_div_ready = true;
      CFD_CUDA(cudaGetLastError());
      return CFD_OK;
    }
    dim3 blk(256), grd((nx + 1 + 255) / 256, v_row_end() - ja);
    cfdk::k_corrector<R><<<grd, blk, 0, stream>>>(scalars(dt_sub), us.v, vs.v, uk.v, vk.v, pp[ipp].v, uo.v, vo.v, p.v,
                                                  ja, jb, v_row_end(), h_divs.dx, h_divs.dy);
    ++launches;
    CFD_CUDA(cudaGetLastError());
    return CFD_OK;
  }

  // Model::update, src/model.rs:304-379 (with piso_step :529-730 inlined)
  int update() override {
    if (!ready) return fail(CFD_ERR_INVALID_ARGUMENT, "model not initialised");
    CFD_CUDA(cudaSetDevice(device));
    const auto t0 = std::chrono::steady_clock::now();
    launches = 0;
    ev_prof_used = 0;
    corr_div_ready = false;  // (an update() that failed half-way must not leave it set)
    CFD_CUDA(cudaEventRecord(ev_step0, stream));
    if (simulation_step < (uint64_t)opt.consts.ramp_up_steps) {  // :311-316
      current_inlet_velocity = (R(simulation_step) / R(opt.consts.ramp_up_steps)) * target_inlet_velocity;
    } else {
      current_inlet_velocity = target_inlet_velocity;
    }
    int rc0;
    const R dt_sub = dt / R(substep_count);  // :317
    last_piso_substeps = substep_count;
    if ((rc0 = refresh_dt_divisor(dt_sub))) return rc0;
    last_K = 0;
    last_S = 0;
    int rc;
    // u_old <- u, v_old <- v (:307-308): with one piso_step per update (the reference's substep_count, :267) the current
    // buffers simply stay untouched until the step ends; several sub-steps (adaptive_substeps) need a real copy
    const uint64_t n_sub = substep_count;
    if (n_sub > 1) {
      if (!uold_buf.base && ((rc = falloc(&uold_buf, (size_t)nx + 1)) || (rc = falloc(&vold_buf, (size_t)nx)))) return rc;
      CFD_CUDA(cudaMemcpyAsync(uold_buf.row(ja), ubuf[iu].row(ja), own_u() * sizeof(R), cudaMemcpyDeviceToDevice, stream));
      CFD_CUDA(cudaMemcpyAsync(vold_buf.row(ja), vbuf[iu].row(ja), own_v() * sizeof(R), cudaMemcpyDeviceToDevice, stream));
    }
    old_in_copy = n_sub > 1;
    ev_sweep_used = 0;
    for (uint64_t sub = 0; sub < n_sub; ++sub) {  // :322-329, piso_step (:529-730) inlined
    const int X = iu, Y = ius, Z = ifree;
    // strips: the predictor stencils reach two rows into the neighbours (second order)
    if ((rc = exchange_halo(ubuf[X], ja, jb, kPredHalo))) return rc;
    if ((rc = exchange_halo(vbuf[X], ja, v_row_end(), kPredHalo))) return rc;
    // ---- predictor (:538-670): reads u, v; writes the interior of u_star, v_star (the rest is carried state)
    {
      const auto s = scalars(dt_sub);
      const int ju_lo = ja > 1 ? ja : 1, ju_hi = jb < ny - 1 ? jb : ny - 1;      // u rows 1..ny-2
      const int jv_lo = ja > 1 ? ja : 1, jv_hi = v_row_end() < ny ? v_row_end() : ny;  // v rows 1..ny-1
      cfdk::PredDivs<R> pd;
      pd.dx = h_divs.dx; pd.dy = h_divs.dy; pd.dx_sq = h_divs.dx_sq; pd.dy_sq = h_divs.dy_sq;
      if (velocity_scheme != CFD_SCHEME_FIRST_ORDER) {
        dim3 blk(256);
        dim3 gu((nx + 255) / 256, ju_hi - ju_lo), gv((nx - 1 + 255) / 256, jv_hi - jv_lo);
        if (velocity_scheme == CFD_SCHEME_QUICK) {  // extension: the JS twin's QUICK face values
          cfdk::k_predict_u<R, 2><<<gu, blk, 0, stream>>>(s, pd, ubuf[X].v, vbuf[X].v, mask_u.v, ubuf[Y].v, ju_lo, ju_hi);
          cfdk::k_predict_v<R, 2><<<gv, blk, 0, stream>>>(s, pd, ubuf[X].v, vbuf[X].v, mask_v.v, vbuf[Y].v, jv_lo, jv_hi);
        } else {
          cfdk::k_predict_u<R, 1><<<gu, blk, 0, stream>>>(s, pd, ubuf[X].v, vbuf[X].v, mask_u.v, ubuf[Y].v, ju_lo, ju_hi);
          cfdk::k_predict_v<R, 1><<<gv, blk, 0, stream>>>(s, pd, ubuf[X].v, vbuf[X].v, mask_v.v, vbuf[Y].v, jv_lo, jv_hi);
        }
        launches += 2;
      } else {
        // first order: both equations in one pass over u and v
        const int j_end = ju_hi > jv_hi ? ju_hi : jv_hi;
        dim3 blk(128), grd((nx + 127) / 128, (j_end - ju_lo + pred_rows - 1) / pred_rows);
        cfdk::k_predict_first<R><<<grd, blk, 0, stream>>>(s, pd, ubuf[X].v, vbuf[X].v, mask_u.v, mask_v.v, ubuf[Y].v, vbuf[Y].v,
                                                          ju_lo, ju_hi, jv_hi, pred_rows);
        launches += 1;
      }
      CFD_CUDA(cudaGetLastError());
      star_alias = false;  // interior written by the predictor, the carried entries were saved when the alias was made
    }
    // ---- first pressure solve + corrector (:676-693): new u, v go to the free buffer, keeping X as u_old
    R residual = 0;
    if ((rc = pressure_solve(dt_sub, ubuf[Y], vbuf[Y], 0, &residual))) return rc;
    last_pressure_residual = residual;
    // the re-correction round that follows starts with the divergence of the corrected fields: when that solve is expected
    // to be over before its first iteration (mg_pred), the corrector computes it on the way
    const bool with_div = fuse_corr_div && world == 1 && pressure_solver == CFD_SOLVER_MGCG && opt.consts.outer_rounds > 0 &&
                          mg_pred[1] == 0;
    if ((rc = corrector(dt_sub, ubuf[Y], vbuf[Y], ubuf[X], vbuf[X], ubuf[Z], vbuf[Z], with_div))) return rc;
    int cur = Z, star = Y;
    // ---- outer re-correction loop (:696-724): `star <- current` is a role swap
    bool alias = false;
    for (int it = 0; it < opt.consts.outer_rounds; ++it) {
      { const int t = star; star = cur; cur = t; }  // star now aliases the latest u, v; `cur` is overwritten in full
      bool elided = false;
      if ((rc = pressure_solve(dt_sub, ubuf[star], vbuf[star], it + 1, &residual, &elided))) return rc;
      last_pressure_residual = residual;
      if (elided) {
        // p' == 0: the corrector would copy star to cur unchanged (u* - dt * 0, p + 0).  Keep the roles instead: the
        // latest fields stay current, the star fields logically equal them (star_alias)
        { const int t = star; star = cur; cur = t; }
        alias = true;
      } else {
        if ((rc = corrector(dt_sub, ubuf[star], vbuf[star], ubuf[star], vbuf[star], ubuf[cur], vbuf[cur]))) return rc;
        alias = false;
      }
      if (last_pressure_residual < R(opt.consts.outer_tolerance)) break;  // :721
    }
    if (alias) {
      // the entries of the star fields that stay observable or that the boundary conditions are about to
      // change in the current fields: copy them now, before :728
      const int n = (nx > ny ? nx : ny) + 2;
      cfdk::k_star_save_edges<R><<<(n + 255) / 256, 256, 0, stream>>>(nx, ny, ubuf[cur].v, vbuf[cur].v, ubuf[star].v, vbuf[star].v,
                                                                      ja, jb, v_row_end(), owns_bottom ? 1 : 0, owns_top ? 1 : 0);
      ++launches;
      const int sj0 = solid_j0 > ja ? solid_j0 : ja, sj1 = solid_j1 < jb ? solid_j1 : jb;
      if (solid_i1 > solid_i0 && sj1 > sj0) {
        dim3 blk(256), grd((solid_i1 - solid_i0 + 255) / 256, sj1 - sj0);
        cfdk::k_star_save_solids<R><<<grd, blk, 0, stream>>>(nx, solid.v, ubuf[cur].v, vbuf[cur].v, ubuf[star].v, vbuf[star].v,
                                                             solid_i0, solid_i1, sj0, sj1);
        ++launches;
      }
    }
    star_alias = alias;
    // ---- boundary conditions (:728 -> :827-875)
    {
      cfdk::BcScalars<R> b;
      b.dy = dy; b.ly = ly; b.inlet = current_inlet_velocity; b.nx = nx; b.ny = ny;
      b.parabolic = inlet_profile == CFD_INLET_PARABOLIC;
      b.cavity = scenario == CFD_SCENARIO_CAVITY;
      const int n = (nx > ny ? nx : ny) + 1;
      cfdk::k_bc_edges<R><<<(n + 255) / 256, 256, 0, stream>>>(b, ubuf[cur].v, vbuf[cur].v, ja, jb, owns_bottom ? 1 : 0,
                                                              owns_top ? 1 : 0);
      ++launches;
      const int sj0 = solid_j0 > ja ? solid_j0 : ja, sj1 = solid_j1 < jb ? solid_j1 : jb;
      if (solid_i1 > solid_i0 && sj1 > sj0) {  // :869-874 over the obstacle's bounding box
        dim3 blk(256), grd((solid_i1 - solid_i0 + 255) / 256, sj1 - sj0);
        cfdk::k_bc_solids<R><<<grd, blk, 0, stream>>>(nx, solid.v, ubuf[cur].v, vbuf[cur].v, solid_i0, solid_i1, sj0, sj1);
        ++launches;
      }
    }
    iu = cur; ius = star; ifree = X;
    }  // sub-steps
    // ---- residuals and CFL maxima (:333-348, :878-881) over the owned rows, then over the ranks
    CFD_CUDA(cudaMemsetAsync(step_slots, 0, 4 * sizeof(unsigned long long), stream));
    {
      const size_t nu_own = (size_t)(jb - ja) * (nx + 1), nv_own = (size_t)(v_row_end() - ja) * nx;
      const Field<R>& uo = old_in_copy ? uold_buf : ubuf[ifree];
      const Field<R>& vo = old_in_copy ? vold_buf : vbuf[ifree];
      cfdk::k_step_maxima<R><<<148 * 8, 256, 0, stream>>>(ubuf[iu].row(ja), uo.row(ja), nu_own, vbuf[iu].row(ja), vo.row(ja), nv_own,
                                                         step_slots);
      ++launches;
    }
    if ((rc = allreduce_max_u64(step_slots, 4))) return rc;
    h_step[4] = 0ull;
    publish(step_slots, h_step, 4 * sizeof(unsigned long long));
    if (peer2_ready) publish(&box->error, h_step + 4, sizeof(unsigned long long));
    CFD_CUDA(cudaEventRecord(ev_step1, stream));
    CFD_CUDA(cudaGetLastError());
    CFD_CUDA(cudaStreamSynchronize(stream));
    if (h_step[4] != 0ull)
      return fail(CFD_ERR_PEER_TIMEOUT, "strips over peer memory: a neighbour did not answer within 20 s (exchange / reduction " +
                                            std::to_string((unsigned long long)h_step[4]) + "); the model's state is undefined");
    last_u_residual = (R)cfdk::bits_nonneg(h_step[0]);
    last_v_residual = (R)cfdk::bits_nonneg(h_step[1]);
    simulation_step += 1;    // :350
    if (opt.consts.adaptive_substeps) {
      // EXTENSION: the reference's sub-step adaptation, commented out at src/model.rs:352-363, made live
      const R error_norm = last_pressure_residual;
      const R tolerance = R(1e-3);
      if (error_norm > tolerance) {
        const R factor = error_norm / tolerance;
        const R grown = std::ceil(R(substep_count) * factor);
        substep_count = (uint64_t)(grown < R(20.0) ? grown : R(20.0));
      } else if (error_norm < tolerance / R(2.0) && substep_count > 1) {
        substep_count = (uint64_t)std::floor(R(substep_count) / R(2.0));
        if (substep_count < 1) substep_count = 1;
      }
    }
    simulation_time += dt;   // :365
    // compute_automatic_time_step (:878-889) + the (dead) growth limiter (:368-377)
    {
      const R max_u = (R)cfdk::bits_nonneg(h_step[2]), max_v = (R)cfdk::bits_nonneg(h_step[3]);
      const R max_vel = max_u > max_v ? max_u : max_v;
      R new_dt;
      if (max_vel == R(0)) {
        new_dt = dt;
      } else {
        const R cfl = R(opt.consts.cfl);
        const R dt_cfl = cfl * (dx < dy ? dx : dy) / max_vel;
        new_dt = dt_cfl < dt ? dt_cfl : dt;
      }
      const R previous_dt = dt;
      const R max_increase_factor = R(1.1);
      if (new_dt > previous_dt) {
        const R lim = previous_dt * max_increase_factor;
        dt = new_dt < lim ? new_dt : lim;
      } else {
        dt = new_dt;
      }
    }
    float ms = 0;
    CFD_CUDA(cudaEventElapsedTime(&ms, ev_step0, ev_step1));
    last_step_ms = ms;
    last_sweep_ms = 0;
    for (size_t k = 0; k + 1 < ev_sweep_used; k += 2) {
      CFD_CUDA(cudaEventElapsedTime(&ms, ev_sweep[k], ev_sweep[k + 1]));
      last_sweep_ms += ms;
    }
    last_launches = launches;
    last_prof_ms = 0;
    last_prof_launches = ev_prof_used / 2;
    for (size_t k = 0; k + 1 < ev_prof_used; k += 2) {
      CFD_CUDA(cudaEventElapsedTime(&ms, ev_prof[k], ev_prof[k + 1]));
      last_prof_ms += ms;
    }
    last_step_seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();  // :378
    return CFD_OK;
  }

  // Model::set_parameters, src/model.rs:1250-1257
  int set_params(const cfd_params& prm) override {
    // the same enum checks as cfd_model_create_ex (`scenario` is not a parameter of set_parameters and is ignored);
    // dt and viscosity are taken as they are, like the reference's set_parameters (:1250-1257) — a NaN dt "just shows
    // NaNs" there too
    if (prm.pressure_solver < CFD_SOLVER_JACOBI || prm.pressure_solver > CFD_SOLVER_MGCG || prm.velocity_scheme < 0 ||
        prm.velocity_scheme > CFD_SCHEME_QUICK || prm.inlet_profile < 0 || prm.inlet_profile > CFD_INLET_PARABOLIC)
      return fail(CFD_ERR_INVALID_ARGUMENT, "set_parameters: enum value out of range");
    nu = R(prm.viscosity);
    dt = R(prm.dt);
    target_inlet_velocity = R(prm.target_inlet_velocity);
    velocity_scheme = prm.velocity_scheme;
    pressure_solver = prm.pressure_solver;
    inlet_profile = prm.inlet_profile;
    return CFD_OK;
  }

  int ensure_staging(size_t bytes) {
    if (bytes > staging_bytes) {
      cudaFree(staging);
      staging = nullptr; staging_bytes = 0;
      CFD_CUDA(cudaMalloc(&staging, bytes));
      staging_bytes = bytes;
    }
    if (bytes > h_staging_bytes) {
      if (h_staging) cudaFreeHost(h_staging);
      h_staging = nullptr; h_staging_bytes = 0;
      CFD_CUDA(cudaHostAlloc(&h_staging, bytes, cudaHostAllocDefault));
      h_staging_bytes = bytes;
    }
    return CFD_OK;
  }

  // owned parts of the three snapshot fields (whole fields when world == 1)
  size_t own_p() const { return (size_t)(jb - ja) * nx; }
  size_t own_u() const { return (size_t)(jb - ja) * (nx + 1); }
  size_t own_v() const { return (size_t)(v_row_end() - ja) * nx; }

  // device -> host copy of `bytes` on the model's stream.  Pinned destinations (cudaHostAlloc / cudaHostRegister,
  // e.g. cfd_host_alloc) are written by the copy engine directly; pageable ones go through two pinned bounce
  // buffers so that the PCIe transfer of chunk k+1 overlaps the host memcpy of chunk k.
  int copy_to_host(void* dst, const void* src_dev, size_t bytes) {
    cudaPointerAttributes attr;
    memset(&attr, 0, sizeof attr);
    const cudaError_t qe = cudaPointerGetAttributes(&attr, dst);
    if (qe != cudaSuccess) (void)cudaGetLastError();
    if (qe == cudaSuccess && attr.type == cudaMemoryTypeHost) {
      CFD_CUDA(cudaMemcpyAsync(dst, src_dev, bytes, cudaMemcpyDeviceToHost, stream));
      return CFD_OK;
    }
    constexpr size_t kChunk = 8u << 20;
    if (!bounce[0]) {
      for (int k = 0; k < 2; ++k) {
        CFD_CUDA(cudaHostAlloc(&bounce[k], kChunk, cudaHostAllocDefault));
        CFD_CUDA(cudaEventCreateWithFlags(&bounce_ev[k], cudaEventDisableTiming));
      }
    }
    const size_t n_chunks = (bytes + kChunk - 1) / kChunk;
    auto issue = [&](size_t k) -> int {
      const size_t off = k * kChunk, sz = bytes - off < kChunk ? bytes - off : kChunk;
      CFD_CUDA(cudaMemcpyAsync(bounce[k & 1], (const char*)src_dev + off, sz, cudaMemcpyDeviceToHost, stream));
      CFD_CUDA(cudaEventRecord(bounce_ev[k & 1], stream));
      return CFD_OK;
    };
    int rc;
    if (n_chunks && (rc = issue(0))) return rc;
    for (size_t k = 0; k < n_chunks; ++k) {
      if (k + 1 < n_chunks && (rc = issue(k + 1))) return rc;
      CFD_CUDA(cudaEventSynchronize(bounce_ev[k & 1]));
      const size_t off = k * kChunk, sz = bytes - off < kChunk ? bytes - off : kChunk;
      memcpy((char*)dst + off, bounce[k & 1], sz);
    }
    return CFD_OK;
  }

  // Model::get_snapshot, src/model.rs:1259-1267: p, u, v narrowed to f32 on the device, then D2H
  int get_snapshot(float* hp, float* hu, float* hv, float* hdt) override {
    CFD_CUDA(cudaSetDevice(device));
    const size_t np_ = own_p(), nu_ = own_u(), nv_ = own_v(), total = np_ + nu_ + nv_;
    if (total * sizeof(float) > staging_bytes) {
      cudaFree(staging);
      staging = nullptr; staging_bytes = 0;
      CFD_CUDA(cudaMalloc(&staging, total * sizeof(float)));
      staging_bytes = total * sizeof(float);
    }
    float* d = (float*)staging;
    const int grid_sz = 148 * 8;
    if (hp) cfdk::k_to_f32<R><<<grid_sz, 256, 0, stream>>>(p.row(ja), d, np_);
    if (hu) cfdk::k_to_f32<R><<<grid_sz, 256, 0, stream>>>(ubuf[iu].row(ja), d + np_, nu_);
    if (hv) cfdk::k_to_f32<R><<<grid_sz, 256, 0, stream>>>(vbuf[iu].row(ja), d + np_ + nu_, nv_);
    CFD_CUDA(cudaGetLastError());
    int rc;
    if (hp && (rc = copy_to_host(hp, d, np_ * sizeof(float)))) return rc;
    if (hu && (rc = copy_to_host(hu, d + np_, nu_ * sizeof(float)))) return rc;
    if (hv && (rc = copy_to_host(hv, d + np_ + nu_, nv_ * sizeof(float)))) return rc;
    CFD_CUDA(cudaStreamSynchronize(stream));
    if (hdt) *hdt = (float)dt;
    return CFD_OK;
  }

  // get_snapshot in two halves (see the header): begin = narrow on the model's stream + copy on the copy stream
  int snapshot_begin(float* hp, float* hu, float* hv) override {
    CFD_CUDA(cudaSetDevice(device));
    if (snap_in_flight >= 2) return fail(CFD_ERR_INVALID_ARGUMENT, "snapshot_begin: two snapshots are already in flight");
    for (float* h : {hp, hu, hv}) {
      if (!h) continue;
      cudaPointerAttributes attr;
      memset(&attr, 0, sizeof attr);
      if (cudaPointerGetAttributes(&attr, h) != cudaSuccess || attr.type != cudaMemoryTypeHost) {
        (void)cudaGetLastError();
        return fail(CFD_ERR_INVALID_ARGUMENT, "snapshot_begin: destinations must be page-locked (cfd_host_alloc)");
      }
    }
    const size_t np_ = own_p(), nu_ = own_u(), nv_ = own_v(), total = np_ + nu_ + nv_;
    if (!copy_stream) {
      CFD_CUDA(cudaStreamCreateWithFlags(&copy_stream, cudaStreamNonBlocking));
      for (int k = 0; k < 2; ++k) {
        CFD_CUDA(cudaEventCreateWithFlags(&snap_ready[k], cudaEventDisableTiming));
        CFD_CUDA(cudaEventCreateWithFlags(&snap_done[k], cudaEventDisableTiming));
      }
    }
    if (total > snap_stage_floats) {
      if (snap_in_flight) return fail(CFD_ERR_INVALID_ARGUMENT, "snapshot_begin: staging buffers are busy");
      for (int k = 0; k < 2; ++k) {
        cudaFree(snap_stage[k]);
        snap_stage[k] = nullptr;
        CFD_CUDA(cudaMalloc((void**)&snap_stage[k], total * sizeof(float)));
      }
      snap_stage_floats = total;
    }
    const int slot = snap_head;
    float* d = snap_stage[slot];
    const int grid_sz = 148 * 8;
    if (hp) cfdk::k_to_f32<R><<<grid_sz, 256, 0, stream>>>(p.row(ja), d, np_);
    if (hu) cfdk::k_to_f32<R><<<grid_sz, 256, 0, stream>>>(ubuf[iu].row(ja), d + np_, nu_);
    if (hv) cfdk::k_to_f32<R><<<grid_sz, 256, 0, stream>>>(vbuf[iu].row(ja), d + np_ + nu_, nv_);
    CFD_CUDA(cudaGetLastError());
    CFD_CUDA(cudaEventRecord(snap_ready[slot], stream));
    CFD_CUDA(cudaStreamWaitEvent(copy_stream, snap_ready[slot], 0));
    if (hp) CFD_CUDA(cudaMemcpyAsync(hp, d, np_ * sizeof(float), cudaMemcpyDeviceToHost, copy_stream));
    if (hu) CFD_CUDA(cudaMemcpyAsync(hu, d + np_, nu_ * sizeof(float), cudaMemcpyDeviceToHost, copy_stream));
    if (hv) CFD_CUDA(cudaMemcpyAsync(hv, d + np_ + nu_, nv_ * sizeof(float), cudaMemcpyDeviceToHost, copy_stream));
    CFD_CUDA(cudaEventRecord(snap_done[slot], copy_stream));
    snap_dt[slot] = (float)dt;
    snap_head ^= 1;
    snap_in_flight += 1;
    return CFD_OK;
  }
  int snapshot_end(float* hdt) override {
    if (snap_in_flight == 0) return fail(CFD_ERR_INVALID_ARGUMENT, "snapshot_end: no snapshot in flight");
    CFD_CUDA(cudaSetDevice(device));
    const int slot = (snap_head + 2 - snap_in_flight) & 1;  // the oldest one
    CFD_CUDA(cudaEventSynchronize(snap_done[slot]));
    if (hdt) *hdt = snap_dt[slot];
    snap_in_flight -= 1;
    return CFD_OK;
  }

  // The UI's colour map (src/app.rs:235-404) on the device: nx x ny RGBA pixels into `rgba` (host memory)
  int render(int mode, unsigned char* rgba, float* min_out, float* max_out) override {
    if (world > 1) return fail(CFD_ERR_UNSUPPORTED, "render: single domain only in this version");
    if (mode < 0 || mode > 2) return fail(CFD_ERR_INVALID_ARGUMENT, "render: mode must be 0 (pressure), 1 (velocity) or 2 (vorticity)");
    CFD_CUDA(cudaSetDevice(device));
    const size_t bytes = n_p * 4 + 16;
    if (bytes > staging_bytes) {
      cudaFree(staging);
      staging = nullptr; staging_bytes = 0;
      CFD_CUDA(cudaMalloc(&staging, bytes));
      staging_bytes = bytes;
    }
    unsigned* slots = (unsigned*)((char*)staging + n_p * 4);
    const unsigned init[2] = {0xff800000u /* key of +inf */, 0x007fffffu /* key of -inf */};
    CFD_CUDA(cudaMemcpyAsync(slots, init, sizeof init, cudaMemcpyHostToDevice, stream));
    cfdk::RenderGeom g;
    g.nx = nx; g.ny = ny; g.mode = mode; g.has_obstacle = grid.has_obstacle != 0;
    g.dx = grid.dx; g.dy = grid.dy; g.cx = grid.center_x; g.cy = grid.center_y; g.radius = grid.radius;
    const dim3 blk(256), grd((nx + 255) / 256, ny);
    cfdk::k_render_minmax<R><<<grd, blk, 0, stream>>>(g, p.v, ubuf[iu].v, vbuf[iu].v, slots);
    cfdk::k_render_pixels<R><<<grd, blk, 0, stream>>>(g, p.v, ubuf[iu].v, vbuf[iu].v, slots, (uchar4*)staging);
    CFD_CUDA(cudaGetLastError());
    unsigned h_slots[2];
    CFD_CUDA(cudaMemcpyAsync(h_slots, slots, sizeof h_slots, cudaMemcpyDeviceToHost, stream));
    int rc;
    if (rgba && (rc = copy_to_host(rgba, staging, n_p * 4))) return rc;
    CFD_CUDA(cudaStreamSynchronize(stream));
    if (min_out) *min_out = cfdk::f32_from_order_key(h_slots[0]);
    if (max_out) *max_out = cfdk::f32_from_order_key(h_slots[1]);
    return CFD_OK;
  }

  // Model::get_residuals, src/model.rs:1269-1280
  int get_residuals(cfd_residuals* out) override {
    out->simulation_step = simulation_step;
    out->simulation_time = (float)simulation_time;
    out->dt = (float)dt;
    out->p = (float)last_pressure_residual;
    out->u = (float)last_u_residual;
    out->v = (float)last_v_residual;
    out->step_seconds = last_step_seconds;
    out->piso_substeps = last_piso_substeps;
    out->jacobi_calls = last_K;
    out->sweeps = last_S;
    out->simulation_time_f64 = (double)simulation_time;
    out->dt_f64 = (double)dt;
    out->p_f64 = (double)last_pressure_residual;
    out->u_f64 = (double)last_u_residual;
    out->v_f64 = (double)last_v_residual;
    const bool mode_c = pressure_solver != CFD_SOLVER_JACOBI;
    out->p_rel_f64 = mode_c ? last_p_rel : 0.0;
    out->rhs_rms_f64 = mode_c ? last_rhs_rms : 0.0;
    out->first_solve_iterations = mode_c ? last_first_solve_iterations : 0;
    return CFD_OK;
  }

  // first owned entry and owned length of a real field
  R* real_field(int field, size_t* n, bool for_write = false) {
    // logical state -> physical buffers first (all rare paths: parity harness / restart)
    const bool star = field == CFD_FIELD_U_STAR || field == CFD_FIELD_V_STAR;
    const bool current = field == CFD_FIELD_U || field == CFD_FIELD_V;
    if ((star || (for_write && current)) && materialize_star() != CFD_OK) { *n = 0; return nullptr; }
    if (field == CFD_FIELD_P_PRIME && materialize_pp_zero() != CFD_OK) { *n = 0; return nullptr; }
    if (field == CFD_FIELD_P_PRIME && for_write && resolve_mg_rotation() != CFD_OK) { *n = 0; return nullptr; }
    if (for_write && (field == CFD_FIELD_MG_LAST || field == CFD_FIELD_MG_LAST2) && !mg_guess_explicit) {
      // the start vector is independent state in the oracle: pin its current value before the history changes
      size_t ng;
      if (!real_field(CFD_FIELD_MG_GUESS, &ng)) { *n = 0; return nullptr; }
      mg_guess_explicit = true;
    }
    switch (field) {
      case CFD_FIELD_P: *n = own_p(); return p.row(ja);
      case CFD_FIELD_U: *n = own_u(); return ubuf[iu].row(ja);
      case CFD_FIELD_V: *n = own_v(); return vbuf[iu].row(ja);
      case CFD_FIELD_U_STAR: *n = own_u(); return ubuf[ius].row(ja);
      case CFD_FIELD_V_STAR: *n = own_v(); return vbuf[ius].row(ja);
      case CFD_FIELD_RHS: *n = own_p(); return rhs.row(ja);
      case CFD_FIELD_P_PRIME: *n = own_p(); return pp[ipp].row(ja);
      // after a step the free buffer still holds the fields the step started from (u_old, v_old)
      case CFD_FIELD_U_OLD: *n = own_u(); return old_in_copy ? uold_buf.row(ja) : ubuf[ifree].row(ja);
      case CFD_FIELD_V_OLD: *n = own_v(); return old_in_copy ? vold_buf.row(ja) : vbuf[ifree].row(ja);
      case CFD_FIELD_MG_GUESS:
      case CFD_FIELD_MG_LAST:
      case CFD_FIELD_MG_LAST2: {
        if (mg.empty() && mg_setup() != CFD_OK) { *n = 0; return nullptr; }
        if (resolve_mg_rotation() != CFD_OK) { *n = 0; return nullptr; }
        *n = own_p();
        if (field == CFD_FIELD_MG_LAST) return mg_hist[0].row(ja);
        if (field == CFD_FIELD_MG_LAST2) return mg_hist[1].row(ja);
        // the start vector is derived state: write it out (unless it was set explicitly)
        if (!mg_guess.base && falloc(&mg_guess, (size_t)nx) != CFD_OK) { *n = 0; return nullptr; }
        if (!mg_guess_explicit) {
          cfdk::MgStart<R> g = mg_start(opt.consts.mg_warm_start != 0);
          if (g.mode == 0) { g.a = mg_hist[0].v; g.mode = 1; }  // cold starts: the oracle still records the last p'
          // from one halo row below to one above the owned rows (an explicit start vector needs its halo)
          const size_t off = (size_t)((long)ja * (long)nx);
          cfdk::MgStart<R> go = g;
          go.a = g.a + off; go.b = g.b ? g.b + off : nullptr; go.c = g.c ? g.c + off : nullptr;
          cfdk::k_mg_start_materialize<R><<<148 * 8, cfdk::kMgThreads, 0, stream>>>(go, mg_guess.row(ja), own_p());
        }
        return mg_guess.row(ja);
      }
      case CFD_FIELD_MG_Z:  // inspection only: z of the last V-cycle
        if (mg.empty() || mg_last_z < 0) { *n = 0; return nullptr; }
        *n = own_p();
        return mg_b[mg_last_z].row(ja);
      default: *n = 0; return nullptr;
    }
  }

  int field_len(int field, uint64_t* len) override {
    if (field == CFD_FIELD_MASK_U) { *len = own_u(); return CFD_OK; }
    if (field == CFD_FIELD_MASK_V) { *len = own_v(); return CFD_OK; }
    size_t n;
    if (!real_field(field, &n)) return fail(CFD_ERR_INVALID_ARGUMENT, "unknown field id");
    *len = n;
    return CFD_OK;
  }

  int get_field_f64(int field, double* out, uint64_t len) override {
    CFD_CUDA(cudaSetDevice(device));
    uint64_t n64;
    int rc;
    if ((rc = field_len(field, &n64))) return rc;
    if (len != n64) return fail(CFD_ERR_INVALID_ARGUMENT, "get_field_f64: wrong length");
    if ((rc = ensure_staging(n64 * sizeof(double)))) return rc;
    double* d = (double*)staging;
    if (field == CFD_FIELD_MASK_U || field == CFD_FIELD_MASK_V) {
      cfdk::k_u8_to_f64<<<148 * 8, 256, 0, stream>>>(field == CFD_FIELD_MASK_U ? mask_u.row(ja) : mask_v.row(ja), d,
                                                     (size_t)n64);
    } else {
      size_t n;
      R* src = real_field(field, &n);
      cfdk::k_to_f64<R><<<148 * 8, 256, 0, stream>>>(src, d, n);
    }
    CFD_CUDA(cudaGetLastError());
    CFD_CUDA(cudaMemcpyAsync(h_staging, d, n64 * sizeof(double), cudaMemcpyDeviceToHost, stream));
    CFD_CUDA(cudaStreamSynchronize(stream));
    memcpy(out, h_staging, n64 * sizeof(double));
    return CFD_OK;
  }

  int set_field_f64(int field, const double* in, uint64_t len) override {
    CFD_CUDA(cudaSetDevice(device));
    size_t n;
    R* dst = real_field(field, &n, true);
    if (!dst) return fail(CFD_ERR_INVALID_ARGUMENT, "set_field_f64: field is not writable");
    if (len != n) return fail(CFD_ERR_INVALID_ARGUMENT, "set_field_f64: wrong length");
    int rc;
    if ((rc = ensure_staging(n * sizeof(double)))) return rc;
    memcpy(h_staging, in, n * sizeof(double));
    CFD_CUDA(cudaMemcpyAsync(staging, h_staging, n * sizeof(double), cudaMemcpyHostToDevice, stream));
    cfdk::k_from_f64<R><<<148 * 8, 256, 0, stream>>>((const double*)staging, dst, n);
    CFD_CUDA(cudaGetLastError());
    CFD_CUDA(cudaStreamSynchronize(stream));
    // strips: p' halos are state too (they are refreshed only after a sweep); so are those of the start-vector history
    if (field == CFD_FIELD_P_PRIME && (rc = exchange_halo(pp[ipp], ja, jb, 1))) return rc;
    if (field == CFD_FIELD_MG_LAST && (rc = exchange_halo(mg_hist[0], ja, jb, 1))) return rc;
    if (field == CFD_FIELD_MG_LAST2 && (rc = exchange_halo(mg_hist[1], ja, jb, 1))) return rc;
    if (field == CFD_FIELD_MG_GUESS) mg_guess_explicit = true;
    return CFD_OK;
  }

  int rows(uint64_t* j0, uint64_t* j1) override {
    *j0 = (uint64_t)ja;
    *j1 = (uint64_t)jb;
    return CFD_OK;
  }

  // ---- tracer particles (index.html:1472-1543) ----
  cfdk::TracerGeom tracer_geom() const {
    cfdk::TracerGeom g;
    g.nx = nx; g.ny = ny; g.dx = (double)dx; g.dy = (double)dy; g.lx = (double)lx; g.ly = (double)ly;
    return g;
  }
  int tracers_reserve(size_t want) {
    if (want <= tr_capacity) return CFD_OK;
    size_t cap = tr_capacity ? tr_capacity : 1024;
    while (cap < want) cap *= 2;
    if (cap > 0x7fffffffull) return fail(CFD_ERR_INVALID_ARGUMENT, "too many tracers");
    double2* fresh[2] = {nullptr, nullptr};
    for (int k = 0; k < 2; ++k) CFD_CUDA(cudaMalloc((void**)&fresh[k], cap * sizeof(double2)));
    if (tr_n) CFD_CUDA(cudaMemcpyAsync(fresh[0], tr_pos[tr_cur], (size_t)tr_n * sizeof(double2), cudaMemcpyDeviceToDevice, stream));
    CFD_CUDA(cudaStreamSynchronize(stream));
    cudaFree(tr_pos[0]); cudaFree(tr_pos[1]); cudaFree(tr_keep);
    tr_pos[0] = fresh[0]; tr_pos[1] = fresh[1]; tr_cur = 0;
    tr_keep = nullptr;
    CFD_CUDA(cudaMalloc((void**)&tr_keep, cap));
    if (!tr_count_dev) {
      CFD_CUDA(cudaMalloc((void**)&tr_count_dev, sizeof(unsigned)));
      CFD_CUDA(cudaHostAlloc((void**)&h_tr_count, sizeof(unsigned), cudaHostAllocDefault));
    }
    tr_capacity = cap;
    return CFD_OK;
  }
  int tracers_inject() override {
    if (world > 1) return fail(CFD_ERR_UNSUPPORTED, "tracers: single domain only in this version");
    CFD_CUDA(cudaSetDevice(device));
    int rc;
    if ((rc = tracers_reserve((size_t)tr_n + (size_t)ny))) return rc;
    cfdk::k_tracers_inject<<<(ny + 255) / 256, 256, 0, stream>>>(tracer_geom(), tr_pos[tr_cur], tr_n);
    CFD_CUDA(cudaGetLastError());
    tr_n += (unsigned)ny;
    return CFD_OK;
  }
  int tracers_update(double dt_tr) override {
    if (world > 1) return fail(CFD_ERR_UNSUPPORTED, "tracers: single domain only in this version");
    if (tr_n == 0) return CFD_OK;
    CFD_CUDA(cudaSetDevice(device));
    cfdk::k_tracers_advect<R><<<(tr_n + 255) / 256, 256, 0, stream>>>(tracer_geom(), ubuf[iu].v, vbuf[iu].v, dt_tr, tr_pos[tr_cur],
                                                                     tr_keep, tr_n);
    cfdk::k_tracers_compact<<<1, 1024, 0, stream>>>(tr_pos[tr_cur], tr_keep, tr_n, tr_pos[tr_cur ^ 1], tr_count_dev);
    CFD_CUDA(cudaGetLastError());
    CFD_CUDA(cudaMemcpyAsync(h_tr_count, tr_count_dev, sizeof(unsigned), cudaMemcpyDeviceToHost, stream));
    CFD_CUDA(cudaStreamSynchronize(stream));
    tr_n = *h_tr_count;
    tr_cur ^= 1;
    return CFD_OK;
  }
  int tracers_get(double* xy, uint64_t capacity, uint64_t* n) override {
    if (n) *n = tr_n;
    const uint64_t take = capacity < tr_n ? capacity : tr_n;
    if (xy && take) {
      CFD_CUDA(cudaSetDevice(device));
      CFD_CUDA(cudaMemcpyAsync(xy, tr_pos[tr_cur], take * sizeof(double2), cudaMemcpyDeviceToHost, stream));
      CFD_CUDA(cudaStreamSynchronize(stream));
    }
    return CFD_OK;
  }
  int tracers_clear() override {
    tr_n = 0;
    return CFD_OK;
  }

  int profile_smoother(int enable) override {
    prof_smoother = enable != 0;
    return CFD_OK;
  }
  int last_smoother_timing(double* ms, uint64_t* n) override {
    if (ms) *ms = last_prof_ms;
    if (n) *n = last_prof_launches;
    return CFD_OK;
  }

  int last_timing(double* step_ms, double* sweep_ms, uint64_t* n_launches) override {
    if (step_ms) *step_ms = last_step_ms;
    if (sweep_ms) *sweep_ms = last_sweep_ms;
    if (n_launches) *n_launches = last_launches;
    return CFD_OK;
  }
};

}  // namespace

struct cfd_model {
  std::unique_ptr<ModelBase> impl;
};

extern "C" {

void cfd_solver_consts_default(cfd_solver_consts* out) {
  if (out) consts_default(out);
}

void cfd_options_default(cfd_options* out) {
  if (!out) return;
  memset(out, 0, sizeof *out);
  out->precision = 64;
  out->device = -1;
  out->rank = 0;
  out->world_size = 1;
  out->nccl_unique_id = nullptr;
  out->flags = 0;
  consts_default(&out->consts);
}

int cfd_model_create_ex(const cfd_grid* grid, const cfd_params* params, const cfd_options* opts, cfd_model** out) {
  if (!grid || !params || !out) return fail(CFD_ERR_INVALID_ARGUMENT, "null argument");
  *out = nullptr;
  cfd_options o;
  if (opts) o = *opts; else cfd_options_default(&o);
  // the reference's 8-lane chunking panics unless nx % 8 is 0 (or 1); this build takes 0 only.
  if (grid->nx % 8 != 0 || grid->nx < 16 || grid->ny < 4)
    return fail(CFD_ERR_INVALID_ARGUMENT, "grid: need nx % 8 == 0, nx >= 16, ny >= 4 (the reference panics otherwise)");
  if (grid->nx > (1u << 30) || grid->ny > (1u << 30) || (grid->nx + 1) * (grid->ny + 1) > 0x7fffffffull)
    return fail(CFD_ERR_INVALID_ARGUMENT, "grid too large (at most 2^31 - 1 entries per field)");
  if (!(grid->dx > 0.0f) || !(grid->dy > 0.0f)) return fail(CFD_ERR_INVALID_ARGUMENT, "grid: dx, dy must be positive");
  if (o.precision != 64 && o.precision != 32) return fail(CFD_ERR_INVALID_ARGUMENT, "precision must be 64 or 32");
#ifndef CFD_WITH_AB_SWEEPS
  if (o.flags & (CFD_FLAG_REGISTER_SWEEP | CFD_FLAG_BULK_SWEEP | CFD_FLAG_SWEEP4 | CFD_FLAG_TEMPORAL | CFD_FLAG_PERSISTENT_SWEEP))
    return fail(CFD_ERR_UNSUPPORTED, "the A/B sweep kernels (register prefetch, bulk copy, one-row, temporal, persistent queue) are "
                                     "compiled into libcfd_b200_ab.so only (csrc/Makefile: make ab; load it with CFD_B200_LIB)");
#endif
  if (o.world_size < 1 || o.rank < 0 || o.rank >= o.world_size) return fail(CFD_ERR_INVALID_ARGUMENT, "rank / world_size out of range");
  if (o.world_size > 1 && !o.nccl_unique_id) return fail(CFD_ERR_INVALID_ARGUMENT, "world_size > 1 needs nccl_unique_id");
  if (o.world_size > 1 && (grid->ny - 2) / (uint64_t)o.world_size < 17) return fail(CFD_ERR_INVALID_ARGUMENT, "strips need at least 17 unknown rows per rank");
  if (o.consts.jacobi_iterations < 1 || o.consts.jacobi_iterations > kMaxSweepSlots)
    return fail(CFD_ERR_INVALID_ARGUMENT, "jacobi_iterations must be in 1..256");
  if (o.consts.outer_rounds < 0 || o.consts.outer_rounds > 1000) return fail(CFD_ERR_INVALID_ARGUMENT, "outer_rounds out of range");
  if (o.consts.mg_warm_start < 0 || o.consts.mg_warm_start > 3 || o.consts.mg_smoothing < 1 || o.consts.mg_smoothing > 16 || !(o.consts.mg_omega > 0.0) || !(o.consts.mg_omega <= 1.0))
    return fail(CFD_ERR_INVALID_ARGUMENT, "mg_smoothing must be in 1..16 and mg_omega in (0, 1]");
  if (params->velocity_scheme < 0 || params->velocity_scheme > CFD_SCHEME_QUICK || params->inlet_profile < 0 ||
      params->inlet_profile > 1 || params->scenario < 0 || params->scenario > 1 || params->pressure_solver < 0 ||
      params->pressure_solver > CFD_SOLVER_MGCG)
    return fail(CFD_ERR_INVALID_ARGUMENT, "params: enum value out of range");
  std::unique_ptr<cfd_model> m(new cfd_model());
  if (o.precision == 32) m->impl.reset(new ModelImpl<float>(*grid, *params, o));
  else m->impl.reset(new ModelImpl<double>(*grid, *params, o));
  const int rc = m->impl->init();
  if (rc) return rc;
  *out = m.release();
  return CFD_OK;
}

int cfd_model_create(const cfd_grid* grid, const cfd_params* params, cfd_model** out) {
  return cfd_model_create_ex(grid, params, nullptr, out);
}

void cfd_model_destroy(cfd_model* m) { delete m; }

#define CFD_CHECK_MODEL(m) \
  if (!(m) || !(m)->impl) return fail(CFD_ERR_INVALID_ARGUMENT, "null model")

int cfd_model_update(cfd_model* m) {
  CFD_CHECK_MODEL(m);
  return m->impl->update();
}

int cfd_model_update_n(cfd_model* m, uint64_t n) {
  CFD_CHECK_MODEL(m);
  for (uint64_t k = 0; k < n; ++k) {
    const int rc = m->impl->update();
    if (rc) return rc;
  }
  return CFD_OK;
}

int cfd_model_set_params(cfd_model* m, const cfd_params* params) {
  CFD_CHECK_MODEL(m);
  if (!params) return fail(CFD_ERR_INVALID_ARGUMENT, "null params");
  return m->impl->set_params(*params);
}

int cfd_model_get_snapshot(cfd_model* m, float* p, float* u, float* v, float* dt) {
  CFD_CHECK_MODEL(m);
  return m->impl->get_snapshot(p, u, v, dt);
}

int cfd_model_snapshot_begin(cfd_model* m, float* p, float* u, float* v) {
  CFD_CHECK_MODEL(m);
  return m->impl->snapshot_begin(p, u, v);
}

int cfd_model_snapshot_end(cfd_model* m, float* dt) {
  CFD_CHECK_MODEL(m);
  return m->impl->snapshot_end(dt);
}

int cfd_model_render_rgba(cfd_model* m, int32_t mode, uint8_t* rgba, float* min_out, float* max_out) {
  CFD_CHECK_MODEL(m);
  return m->impl->render(mode, rgba, min_out, max_out);
}

int cfd_model_get_residuals(cfd_model* m, cfd_residuals* out) {
  CFD_CHECK_MODEL(m);
  if (!out) return fail(CFD_ERR_INVALID_ARGUMENT, "null out");
  return m->impl->get_residuals(out);
}

int cfd_model_field_len(cfd_model* m, int32_t field, uint64_t* len) {
  CFD_CHECK_MODEL(m);
  if (!len) return fail(CFD_ERR_INVALID_ARGUMENT, "null len");
  return m->impl->field_len(field, len);
}

int cfd_model_get_field_f64(cfd_model* m, int32_t field, double* out, uint64_t len) {
  CFD_CHECK_MODEL(m);
  if (!out) return fail(CFD_ERR_INVALID_ARGUMENT, "null out");
  return m->impl->get_field_f64(field, out, len);
}

int cfd_model_set_field_f64(cfd_model* m, int32_t field, const double* in, uint64_t len) {
  CFD_CHECK_MODEL(m);
  if (!in) return fail(CFD_ERR_INVALID_ARGUMENT, "null in");
  return m->impl->set_field_f64(field, in, len);
}

int cfd_model_rows(cfd_model* m, uint64_t* j0, uint64_t* j1) {
  CFD_CHECK_MODEL(m);
  if (!j0 || !j1) return fail(CFD_ERR_INVALID_ARGUMENT, "null out");
  return m->impl->rows(j0, j1);
}

int cfd_strip_rows(uint64_t ny, int32_t world_size, int32_t rank, uint64_t* j0, uint64_t* j1) {
  if (!j0 || !j1 || world_size < 1 || rank < 0 || rank >= world_size || ny < 4 || ny > (1u << 30))
    return fail(CFD_ERR_INVALID_ARGUMENT, "cfd_strip_rows: bad argument");
  *j0 = (uint64_t)strip_row_start((int)ny, world_size, rank);
  *j1 = (uint64_t)strip_row_start((int)ny, world_size, rank + 1);
  return CFD_OK;
}

int cfd_model_last_timing(cfd_model* m, double* step_ms, double* sweep_ms, uint64_t* kernel_launches) {
  CFD_CHECK_MODEL(m);
  return m->impl->last_timing(step_ms, sweep_ms, kernel_launches);
}

int cfd_model_tracers_inject(cfd_model* m) {
  CFD_CHECK_MODEL(m);
  return m->impl->tracers_inject();
}

int cfd_model_tracers_update(cfd_model* m, double dt) {
  CFD_CHECK_MODEL(m);
  return m->impl->tracers_update(dt);
}

int cfd_model_tracers_count(cfd_model* m, uint64_t* n) {
  CFD_CHECK_MODEL(m);
  if (!n) return fail(CFD_ERR_INVALID_ARGUMENT, "null out");
  return m->impl->tracers_get(nullptr, 0, n);
}

int cfd_model_tracers_get(cfd_model* m, double* xy, uint64_t capacity, uint64_t* n) {
  CFD_CHECK_MODEL(m);
  if (!xy && capacity) return fail(CFD_ERR_INVALID_ARGUMENT, "null out");
  return m->impl->tracers_get(xy, capacity, n);
}

int cfd_model_tracers_clear(cfd_model* m) {
  CFD_CHECK_MODEL(m);
  return m->impl->tracers_clear();
}

int cfd_model_profile_smoother(cfd_model* m, int32_t enable) {
  CFD_CHECK_MODEL(m);
  return m->impl->profile_smoother(enable);
}

int cfd_model_last_smoother_timing(cfd_model* m, double* ms, uint64_t* launches) {
  CFD_CHECK_MODEL(m);
  return m->impl->last_smoother_timing(ms, launches);
}

int cfd_nccl_unique_id(void* out128) {
  if (!out128) return fail(CFD_ERR_INVALID_ARGUMENT, "null out");
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
  CFD_NCCL_READY();
  ncclUniqueId id;
  CFD_NCCL(nccl_api().GetUniqueId(&id));
  memcpy(out128, &id, sizeof id);
  return CFD_OK;
}

int cfd_selftest_division(double divisor, uint64_t samples, uint64_t seed, int32_t mode, uint64_t* mismatches,
                           uint64_t* fast_path_taken) {
  if (!mismatches || !(divisor > 0.0) || mode < 0 || mode > 3) return fail(CFD_ERR_INVALID_ARGUMENT, "selftest_division: bad argument");
  unsigned long long* d = nullptr;
  CFD_CUDA(cudaMalloc((void**)&d, 2 * sizeof(unsigned long long)));
  CFD_CUDA(cudaMemset(d, 0, 2 * sizeof(unsigned long long)));
  const int blocks = 148 * 4, threads = 256;
  const unsigned long long per_thread = (samples + (uint64_t)blocks * threads - 1) / ((uint64_t)blocks * threads);
  cfdk::k_selftest_division<<<blocks, threads>>>(divisor, per_thread, seed, mode, d, d + 1);
  CFD_CUDA(cudaGetLastError());
  unsigned long long h[2];
  CFD_CUDA(cudaMemcpy(h, d, sizeof h, cudaMemcpyDeviceToHost));
  CFD_CUDA(cudaFree(d));
  *mismatches = h[0];
  if (fast_path_taken) *fast_path_taken = h[1];
  return CFD_OK;
}

int cfd_host_alloc(uint64_t bytes, void** out) {
  if (!out) return fail(CFD_ERR_INVALID_ARGUMENT, "null out");
  *out = nullptr;
  CFD_CUDA(cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocDefault));
  return CFD_OK;
}

void cfd_host_free(void* ptr) {
  if (ptr) cudaFreeHost(ptr);
}

const char* cfd_last_error(void) { return g_last_error.c_str(); }

int cfd_abi_version(void) { return CFD_ABI_VERSION; }

}  // extern "C"
