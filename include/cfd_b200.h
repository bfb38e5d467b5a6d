/*
 * cfd_b200.h — C ABI of the B200-native replacement for cfd-demo's solver hot path.
 *
 * The reference has no FFI today; its boundary is the Rust pub API of `src/model.rs` as consumed by
 * `src/app.rs`.  Every entry point below names the reference item it replaces (paths relative to the
 * reference repo).  The Rust shim in `cfd_demo_b200/rust/` and the C++ mirror in `cfd_demo_b200/host/`
 * bind exactly these symbols; see INTEGRATION.md.
 *
 * Conventions: all pointers are HOST memory; return 0 (CFD_OK) on success, non-zero on error with a
 * thread-local message from cfd_last_error(); a model is owned by one thread at a time (the reference
 * moves `Model` into one solver thread, src/model.rs:1282-1287); distinct models are independent.
 * API scalars stay f32 where the reference's are f32 (Grid / SimulationParams / Residuals / SimSnapshot);
 * the shipped arithmetic is fp64 (precision 64) and promotes them on entry.
 */
#ifndef CFD_B200_H
#define CFD_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CFD_ABI_VERSION 3

/* ---- status codes ---------------------------------------------------------------------------- */
#define CFD_OK 0
#define CFD_ERR_INVALID_ARGUMENT 1 /* reference behaviour: panic (slice index / unwrap)            */
#define CFD_ERR_CUDA 2
#define CFD_ERR_UNSUPPORTED 3
#define CFD_ERR_NCCL 4
#define CFD_ERR_PEER_TIMEOUT 5 /* strips over peer memory: a neighbour did not answer within 20 s (it died or returned an error) */

/* ---- enums (values are ABI) -------------------------------------------------------------------- */
/* VelocityScheme, src/model.rs:142-146 */
#define CFD_SCHEME_FIRST_ORDER 0
#define CFD_SCHEME_SECOND_ORDER 1
/* extension: the JS twin's QUICK face values (index.html:471-549 u, :643-723 v) inside the Rust
 * model's predictor (same loops, flux velocities, Laplacian, masks as SecondOrder: src/model.rs:538-670) */
#define CFD_SCHEME_QUICK 2
/* InletProfile, src/model.rs:155-159 */
#define CFD_INLET_UNIFORM 0
#define CFD_INLET_PARABOLIC 1
/* PressureSolver, src/model.rs:149-152 (Jacobi is the reference's only variant; CG is an extension) */
#define CFD_SOLVER_JACOBI 0
#define CFD_SOLVER_CG 1
/* extension: conjugate gradients preconditioned by one geometric-multigrid V-cycle (Mode C);
 * the damped-Jacobi sweep of the reference (src/model.rs:748-815) is its smoother on the finest level */
#define CFD_SOLVER_MGCG 2
/* Scenario: the reference hard-codes the channel (src/model.rs:807-815,827-875); cavity is an extension */
#define CFD_SCENARIO_CHANNEL 0
#define CFD_SCENARIO_CAVITY 1

/* field ids for cfd_model_get_field_f64 / cfd_model_field_len (Model fields, src/model.rs:181-205) */
#define CFD_FIELD_P 0
#define CFD_FIELD_U 1
#define CFD_FIELD_V 2
#define CFD_FIELD_U_STAR 3
#define CFD_FIELD_V_STAR 4
#define CFD_FIELD_RHS 5
#define CFD_FIELD_P_PRIME 6
#define CFD_FIELD_U_OLD 7
#define CFD_FIELD_V_OLD 8
#define CFD_FIELD_MASK_U 9  /* 0/1 as double */
#define CFD_FIELD_MASK_V 10 /* 0/1 as double */
#define CFD_FIELD_MG_GUESS 11 /* MGCG extension: start vector of the next step's first solve (carried state) */
#define CFD_FIELD_MG_LAST 12  /* MGCG extension: p' the last first-solve ended with (carried state, mg_warm_start 2) */
#define CFD_FIELD_MG_LAST2 13 /* MGCG extension: the one before that (carried state, mg_warm_start 3) */
#define CFD_FIELD_MG_Z 14     /* MGCG extension, read-only inspection: the preconditioned residual z the solver's last V-cycle produced */
#define CFD_FIELD_COUNT 15

/* ---- PODs ---------------------------------------------------------------------------------------- */
/* Grid + Option<Cylinder>, src/model.rs:121-139 */
typedef struct cfd_grid {
  uint64_t nx, ny;     /* pressure cells */
  float lx, ly, dx, dy;
  int32_t has_obstacle; /* Option<Cylinder>: 0 = None */
  float center_x, center_y, radius;
} cfd_grid;

/* SimulationParams, src/model.rs:13-21 (defaults :44-55) + the `scenario` extension (0 = reference) */
typedef struct cfd_params {
  float dt, viscosity, target_inlet_velocity;
  int32_t velocity_scheme, inlet_profile, pressure_solver, scenario;
} cfd_params;

/* Solver literals of the reference, made data so that the extensions are additive.
 * Defaults (cfd_solver_consts_default) are the reference's: src/model.rs:269 (ramp 100),
 * :735-737 (omega .75, tol 1e-4, 50 sweeps), :696,:721 (20 outer rounds, 1e-4), :885 (CFL .2). */
typedef struct cfd_solver_consts {
  int32_t ramp_up_steps;
  int32_t jacobi_iterations;
  int32_t outer_rounds;
  int32_t cg_max_iterations; /* extension (CFD_SOLVER_CG) */
  double jacobi_omega;
  double pressure_tolerance;
  double outer_tolerance;
  double cfl;
  double cg_tolerance;       /* extension (CG, MGCG): stop when dt * rms(Poisson residual) <= this */
  double mg_omega;           /* extension (MGCG): damping of the Jacobi smoother, default 0.8 */
  int32_t mg_smoothing;      /* extension (MGCG): pre- and post-smoothing sweeps per level, default 2 (bench.py: 3; 2..4 run with one launch per leg of the V-cycle) */
  int32_t mg_warm_start;     /* extension (MGCG), start vector of the FIRST solve of a step: 1 = the p' the first solve
                              * of the previous step ended with — like the reference's Jacobi, which never resets p'
                              * (src/model.rs:734-824); 2 = linear extrapolation in time from the last two,
                              * 2 p'_n - p'_(n-1) (the JS twin's "extrapolated initial guess", index.html:262-270);
                              * 3 (default) = quadratic, 3 p'_n - 3 p'_(n-1) + p'_(n-2);
                              * 0 and every re-correction solve = start from p' = 0 */
  int32_t cg_relative;       /* extension (CG, MGCG) stopping rule: 0 (default) = dt * rms(r) <= cg_tolerance, the rms
                              * divergence the correction leaves in the velocity; 1 = ||r||_2 <= cg_tolerance * ||rhs||_2 with
                              * rhs the right-hand side of the step's FIRST solve (the relative L2 norm SURVEY 8d states for
                              * the 4096^2 cavity; re-correction solves measure against the same reference) */
  int32_t adaptive_substeps; /* extension: 0 (default) = substep_count stays 1 like the reference, whose
                              * adaptation is commented out (src/model.rs:352-363); 1 = that commented-out rule is live:
                              * after a step, error = last_pressure_residual; error > 1e-3 -> substeps = min(ceil(substeps *
                              * error / 1e-3), 20); error < 1e-3 / 2 and substeps > 1 -> substeps = max(floor(substeps / 2), 1) */
} cfd_solver_consts;

typedef struct cfd_options {
  int32_t precision;   /* 64 (shipped path) or 32 (the reference's own f32 arithmetic) */
  int32_t device;      /* CUDA ordinal; -1 = current device */
  int32_t rank;        /* strip decomposition over rows; world_size 1 = whole grid on one GPU */
  int32_t world_size;
  const void* nccl_unique_id; /* 128-byte ncclUniqueId, same on every rank, when world_size > 1 */
  uint32_t flags;      /* CFD_FLAG_* */
  cfd_solver_consts consts;
} cfd_options;

#define CFD_FLAG_NO_GRAPH 1u      /* Mode R: one launch per Jacobi sweep also on small grids, instead of the single cooperative launch per
                                   * solve (k_jacobi_persist) that grids whose rows fit into shared memory get by default (A/B, cross-check) */
#define CFD_FLAG_BASELINE_SWEEP 2u /* simple one-column-per-thread Jacobi kernel, compiler divisions (cross-check) */
#define CFD_FLAG_REGISTER_SWEEP 4u /* register-prefetch Jacobi kernel instead of the TMA-staged one (A/B) */
#define CFD_FLAG_SWEEP4 16u        /* one-row-per-step tensor-TMA Jacobi kernel instead of the row-pair one (A/B) */
#define CFD_FLAG_NCCL_EXCHANGE 32u /* Mode R strips: NCCL send/recv + allreduce after every sweep (the default since the peer path became opt-in) */
#define CFD_FLAG_PEER_EXCHANGE 512u /* Mode R strips: the sweep kernel stores its edge rows into the neighbours' halos and publishes its max over
                                      * NVLink peer memory (CUDA IPC), no NCCL in the sweep loop: 0.78 instead of 0.69 weak-scaling efficiency at 2 GPUs,
                                      * but it hung at start-up in 2 of 8 two-GPU test runs (records reused across solves; fixed by double-buffering them,
                                      * validated by a single run so far), so it stays opt-in until the fix has been repeated enough */
#define CFD_FLAG_TEMPORAL 64u      /* two sweeps per HBM pass (k_jacobi_sweep_t2) instead of one per launch (A/B; slower, issue-bound) */
#define CFD_FLAG_PERSISTENT_SWEEP 128u /* persistent warp-queue kernel (k_jacobi_sweep6) instead of one block per tile (A/B; slower) */
#define CFD_FLAG_MG_NO_BOTTOM_KERNEL 256u /* MGCG: one launch per operation on every level instead of the single-block bottom kernel (A/B, cross-check) */
#define CFD_FLAG_MG_UNFUSED 1024u  /* MGCG: separate first-sweep / sweep and prolongation / sweep kernels instead of the fused passes (A/B, cross-check) */
#define CFD_FLAG_PEER_STRIPS 2048u /* strips: halo rows, gathers and scalar reductions as single small kernels over NVLink peer memory
                                    * (CUDA IPC; cfd_peer.cuh) instead of NCCL send/recv / broadcast / allreduce — every solver mode;
                                    * waits are bounded (CFD_ERR_PEER_TIMEOUT).  Environment override: CFD_PEER_STRIPS=0/1 */
#define CFD_FLAG_BULK_SWEEP 8u     /* row-by-row cp.async.bulk Jacobi kernel instead of the tensor-TMA one (A/B) */

/* Residuals, src/model.rs:23-32.  f32 members mirror the reference; the trailing members are
 * additions (solver counters for the roofline accounting, full-precision copies for parity tests). */
typedef struct cfd_residuals {
  uint64_t simulation_step;
  float simulation_time, dt, p, u, v;
  double step_seconds;        /* Residuals::step_time */
  uint64_t piso_substeps;
  uint64_t jacobi_calls;      /* K: pressure solves in the last step (2..21) */
  uint64_t sweeps;            /* S: Jacobi sweeps (or CG iterations) in the last step */
  double simulation_time_f64, dt_f64, p_f64, u_f64, v_f64;
  /* Mode C (CG, MGCG), last step: the first solve's final ||r||_2 / ||rhs||_2 and its dt * rms(rhs) (0 in Mode R) */
  double p_rel_f64, rhs_rms_f64;
  uint64_t first_solve_iterations; /* CG iterations of the step's first solve (the re-correction solves take the rest) */
} cfd_residuals;

typedef struct cfd_model cfd_model; /* opaque: owns device buffers, streams, graphs, (optional) NCCL comm */

/* ---- lifecycle ------------------------------------------------------------------------------------ */
void cfd_solver_consts_default(cfd_solver_consts* out);
void cfd_options_default(cfd_options* out);

/* Model::new(grid, &params), src/model.rs:219-299.  fp64, current device, reference constants.
 * Requires nx % 8 == 0, nx >= 16, ny >= 4 (the reference panics for nx % 8 not in {0,1}: SURVEY N1). */
int cfd_model_create(const cfd_grid* grid, const cfd_params* params, cfd_model** out);
int cfd_model_create_ex(const cfd_grid* grid, const cfd_params* params, const cfd_options* opts,
                        cfd_model** out);
/* Drop for Model (the reference's solver thread ends by panic, src/model.rs:1319; the shim frees here). */
void cfd_model_destroy(cfd_model* m);

/* ---- stepping -------------------------------------------------------------------------------------- */
/* Model::update(&mut self), src/model.rs:304-379: exactly one timestep, synchronous. */
int cfd_model_update(cfd_model* m);
/* Extension: n timesteps in one FFI call (saves the per-call binding overhead; in this version it is exactly n calls of
 * cfd_model_update, convergence scalars and dt control still pass through the host between steps). */
int cfd_model_update_n(cfd_model* m, uint64_t n);
/* Model::set_parameters(&mut self, &params), src/model.rs:1250-1257 (`scenario` is ignored here, as
 * the reference has no such parameter to change). */
int cfd_model_set_params(cfd_model* m, const cfd_params* params);

/* ---- state read-back -------------------------------------------------------------------------------- */
/* Model::get_snapshot(&self) -> SimSnapshot{p,u,v: Vec<f32>, dt}, src/model.rs:1259-1267.
 * Caller-allocated, reference layout: p nx*ny, u (nx+1)*ny, v nx*(ny+1); any pointer may be NULL.
 * With world_size > 1 each rank receives its own rows only (see cfd_model_rows). */
int cfd_model_get_snapshot(cfd_model* m, float* p, float* u, float* v, float* dt);
/* The same in two halves, so that the device->host copy of step n's snapshot overlaps the computation of step n+1 — the
 * shape of the reference's own protocol, where the UI posts Command::GetSnapshot and picks the result up on a later frame
 * while the solver thread keeps stepping (src/model.rs:100-102, :1300-1306; src/app.rs:95-104, :470).
 * begin: narrows the CURRENT p, u, v to f32 into a device staging buffer (on the model's stream: the fields may change right
 *        after) and starts the copy into p / u / v on a second stream; returns at once.  The destinations must be page-locked
 *        (cfd_host_alloc) and stay valid until `end`; any of them may be NULL.  At most two snapshots may be in flight.
 * end:   waits for the OLDEST snapshot in flight; *dt = the dt that was current at its `begin`. */
int cfd_model_snapshot_begin(cfd_model* m, float* p, float* u, float* v);
int cfd_model_snapshot_end(cfd_model* m, float* dt);
/* Page-locked host memory for snapshot buffers (no reference counterpart: `SimSnapshot` owns plain Vec<f32>).
 * cfd_model_get_snapshot detects pinned destinations and lets the copy engine write them directly; pageable
 * destinations are served through pinned bounce buffers (chunked, PCIe transfer overlapped with the memcpy). */
int cfd_host_alloc(uint64_t bytes, void** out);
void cfd_host_free(void* ptr);
/* The UI's colour map of a snapshot (src/app.rs:235-404: pressure :238-279, velocity magnitude :281-330, vorticity
 * :332-398; min/max normalisation :247-249, red-blue ramp, grey cylinder overlay :262-268) computed on the device from
 * the f32-narrowed fields: nx*ny RGBA pixels (egui::Color32::from_rgb), row j of the image = row j of the grid, into
 * caller-allocated host memory (pinned or pageable).  mode: 0 pressure, 1 velocity, 2 vorticity.  min_out / max_out
 * (may be NULL) receive the normalisation range BEFORE the `max = min + 1` adjustment.  Single domain only. */
#define CFD_RENDER_PRESSURE 0
#define CFD_RENDER_VELOCITY 1
#define CFD_RENDER_VORTICITY 2
int cfd_model_render_rgba(cfd_model* m, int32_t mode, uint8_t* rgba, float* min_out, float* max_out);
/* Model::get_residuals(&self), src/model.rs:1269-1280. */
int cfd_model_get_residuals(cfd_model* m, cfd_residuals* out);
/* Parity harness: any field (CFD_FIELD_*) widened to double, reference layout. */
int cfd_model_field_len(cfd_model* m, int32_t field, uint64_t* len);
int cfd_model_get_field_f64(cfd_model* m, int32_t field, double* out, uint64_t len);
/* Parity harness / restart: overwrite a field from host doubles (reference layout). */
int cfd_model_set_field_f64(cfd_model* m, int32_t field, const double* in, uint64_t len);
/* Rows [j0, j1) of pressure cells owned by this rank (whole grid when world_size == 1). */
int cfd_model_rows(cfd_model* m, uint64_t* j0, uint64_t* j1);
/* The strip partition itself (pure host arithmetic, no device needed): rows [j0, j1) of rank `rank` of `world_size`
 * for a grid of ny rows.  Interior boundaries sit at 1 + (a multiple of 16), so that the unknown rows of a strip pair
 * up within the strip on the first four multigrid levels (MGCG on strips). */
int cfd_strip_rows(uint64_t ny, int32_t world_size, int32_t rank, uint64_t* j0, uint64_t* j1);

/* ---- tracer particles (the JS twin, index.html:1472-1543; no Rust counterpart) ------- */
/* Tracers live on the device next to the fields, in injection order.  Single domain only (world_size 1).
 * inject: initTracers / injectTracers (:1475-1483, :1537-1543) — appends one tracer per cell row on the inlet, at
 *         (0, (j + 1/2) dy); the JS animation loop calls it at start and every 100 timesteps (:1233-1237).
 * update: updateTracers(dt) (:1485-1497) — x += dt * u(x, y), y += dt * v(x, y) with the bilinearly interpolated
 *         cell-centred velocity of the CURRENT fields (getVelocityAt, :1499-1526); tracers that leave [0, lx] x [0, ly] are
 *         dropped, the others keep their order.  The JS passes the solver's dt after each timestep.
 * get:    positions as (x, y) pairs of doubles, at most `capacity` tracers; *n = how many exist. */
int cfd_model_tracers_inject(cfd_model* m);
int cfd_model_tracers_update(cfd_model* m, double dt);
int cfd_model_tracers_count(cfd_model* m, uint64_t* n);
int cfd_model_tracers_get(cfd_model* m, double* xy, uint64_t capacity, uint64_t* n);
int cfd_model_tracers_clear(cfd_model* m);

/* ---- measurement hooks (bench.py / profiles; no reference counterpart) ------------------------------- */
/* Device time in ms of the last cfd_model_update / update_n, and of the Jacobi sweeps inside it,
 * both from CUDA events on the model's own stream. */
int cfd_model_last_timing(cfd_model* m, double* step_ms, double* sweep_ms, uint64_t* kernel_launches);

/* MGCG: bracket every launch of the fine-level smoother (k_jacobi_sweep5) with CUDA events on the model's stream;
 * cfd_model_last_smoother_timing returns their summed duration and count for the last cfd_model_update. */
int cfd_model_profile_smoother(cfd_model* m, int32_t enable);
int cfd_model_last_smoother_timing(cfd_model* m, double* ms, uint64_t* launches);

/* Self-test of the hot kernels' exact division by a loop-invariant divisor (cfdk::div_c): draws `samples`
 * dividends (mode 0 random bit patterns, 1 moderate magnitudes, 2 near representable quotients, 3 near
 * rounding midpoints) and counts those whose result differs bitwise from the compiler's `x / divisor`. */
int cfd_selftest_division(double divisor, uint64_t samples, uint64_t seed, int32_t mode, uint64_t* mismatches,
                          uint64_t* fast_path_taken);

/* ---- misc ------------------------------------------------------------------------------------------- */
/* Fills 128 bytes with a fresh ncclUniqueId (call on rank 0, broadcast to the others). */
int cfd_nccl_unique_id(void* out128);
const char* cfd_last_error(void);
int cfd_abi_version(void);

#ifdef __cplusplus
}
#endif
#endif /* CFD_B200_H */
