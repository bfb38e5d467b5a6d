//! Drop-in replacement for the reference's `src/model.rs`: the same pub types and `Model` API, with the fields
//! resident on a B200 behind the C ABI of `include/cfd_b200.h`.  `app.rs` compiles against it unchanged
//! (it only uses `Grid`, `Cylinder`, `SimulationParams`, the three enums, `Model::{new, run}`,
//! `SimulationControlHandle::*`, `SimSnapshot`, `Residuals`).
//!
//! SOURCE ONLY — never compiled: there is no Rust toolchain in this repository's build environment
//!.  The tested boundary is the C ABI these `extern "C"` items bind.
use std::{
    os::raw::{c_char, c_int},
    sync::mpsc,
    thread,
    time::{Duration, Instant},
};

// ---- the C ABI (include/cfd_b200.h) -------------------------------------------------------------------
#[repr(C)]
struct CfdGrid { nx: u64, ny: u64, lx: f32, ly: f32, dx: f32, dy: f32, has_obstacle: i32, center_x: f32, center_y: f32, radius: f32 }
#[repr(C)]
struct CfdParams { dt: f32, viscosity: f32, target_inlet_velocity: f32, velocity_scheme: i32, inlet_profile: i32, pressure_solver: i32, scenario: i32 }
#[repr(C)]
#[derive(Default)]
struct CfdResiduals {
    simulation_step: u64, simulation_time: f32, dt: f32, p: f32, u: f32, v: f32, step_seconds: f64,
    piso_substeps: u64, jacobi_calls: u64, sweeps: u64,
    simulation_time_f64: f64, dt_f64: f64, p_f64: f64, u_f64: f64, v_f64: f64,
    p_rel_f64: f64, rhs_rms_f64: f64, first_solve_iterations: u64, // ABI 3
}
/// `cfd_solver_consts` (ABI 3): the reference's solver literals plus the extensions' knobs.
#[repr(C)]
#[derive(Clone, Copy)]
pub struct SolverConsts {
    pub ramp_up_steps: i32, pub jacobi_iterations: i32, pub outer_rounds: i32, pub cg_max_iterations: i32,
    pub jacobi_omega: f64, pub pressure_tolerance: f64, pub outer_tolerance: f64, pub cfl: f64, pub cg_tolerance: f64,
    pub mg_omega: f64, pub mg_smoothing: i32, pub mg_warm_start: i32, pub cg_relative: i32, pub adaptive_substeps: i32,
}
#[repr(C)]
struct CfdOptions {
    precision: i32, device: i32, rank: i32, world_size: i32, nccl_unique_id: *const std::ffi::c_void, flags: u32,
    consts: SolverConsts,
}
#[repr(C)]
struct CfdModel { _private: [u8; 0] }

extern "C" {
    fn cfd_model_create(grid: *const CfdGrid, params: *const CfdParams, out: *mut *mut CfdModel) -> c_int;
    fn cfd_options_default(out: *mut CfdOptions);
    fn cfd_model_create_ex(grid: *const CfdGrid, params: *const CfdParams, opts: *const CfdOptions, out: *mut *mut CfdModel) -> c_int;
    fn cfd_model_destroy(m: *mut CfdModel);
    fn cfd_model_update(m: *mut CfdModel) -> c_int;
    fn cfd_model_set_params(m: *mut CfdModel, params: *const CfdParams) -> c_int;
    fn cfd_model_get_snapshot(m: *mut CfdModel, p: *mut f32, u: *mut f32, v: *mut f32, dt: *mut f32) -> c_int;
    fn cfd_model_get_residuals(m: *mut CfdModel, out: *mut CfdResiduals) -> c_int;
    /// extension: the UI's colour map (src/app.rs:235-404) computed on the device; mode 0 pressure, 1 velocity, 2 vorticity
    fn cfd_model_render_rgba(m: *mut CfdModel, mode: i32, rgba: *mut u8, min_out: *mut f32, max_out: *mut f32) -> c_int;
    fn cfd_last_error() -> *const c_char;
}

fn check(rc: c_int) {
    if rc != 0 {
        let msg = unsafe { std::ffi::CStr::from_ptr(cfd_last_error()) }.to_string_lossy().into_owned();
        panic!("cfd_b200: {msg}"); // the reference panics on every error path too (unwrap / slice index)
    }
}

// ---- the reference's pub types, unchanged (src/model.rs:13-63, 121-159) --------------------------------
#[derive(Clone)]
pub struct SimulationParams {
    pub dt: f32,
    pub viscosity: f32,
    pub target_inlet_velocity: f32,
    pub velocity_scheme: VelocityScheme,
    pub inlet_profile: InletProfile,
    pub pressure_solver: PressureSolver,
    /// Extension (the reference hard-codes the channel, src/model.rs:807-815, :827-875); `Default` keeps it.
    pub scenario: Scenario,
}
pub struct Residuals {
    pub simulation_step: usize, pub simulation_time: f32, pub dt: f32, pub p: f32, pub u: f32, pub v: f32,
    pub step_time: Duration, pub piso_substeps: usize,
}
#[derive(Clone)]
pub struct SimSnapshot { pub p: Vec<f32>, pub u: Vec<f32>, pub v: Vec<f32>, pub dt: f32, pub paused: bool }
impl Default for SimulationParams {
    fn default() -> Self {
        Self { dt: 0.005, viscosity: 0.000001, target_inlet_velocity: 1.0, velocity_scheme: VelocityScheme::FirstOrder,
               inlet_profile: InletProfile::Uniform, pressure_solver: PressureSolver::Jacobi, scenario: Scenario::Channel }
    }
}
pub enum Command { Stop, GetSnapshot, SetParams(SimulationParams), Pause, Resume }
#[derive(Clone)]
pub struct Grid { pub nx: usize, pub ny: usize, pub lx: f32, pub ly: f32, pub dx: f32, pub dy: f32, pub obstacle: Option<Cylinder> }
#[derive(Clone)]
pub struct Cylinder { pub center_x: f32, pub center_y: f32, pub radius: f32 }
#[derive(Debug, Clone, Copy, PartialEq, Eq)]
pub enum VelocityScheme { FirstOrder, SecondOrder, /** extension: the JS twin's QUICK face values */ Quick }
/// `Jacobi` is the reference's only variant (src/model.rs:149-152); `Cg` and `Mgcg` are the converged solvers the
/// reference's doc comment asks for ("Jacobi, SOR or multigrid", :526) — app.rs's combo box lists whatever is here.
#[derive(Debug, Clone, Copy, PartialEq, Eq)]
pub enum PressureSolver { Jacobi, Cg, Mgcg }
#[derive(Debug, Clone, Copy, PartialEq, Eq)]
pub enum Scenario { Channel, Cavity }
#[derive(Debug, Clone, Copy, PartialEq, Eq)]
pub enum InletProfile { Uniform, Parabolic }

fn params_to_c(p: &SimulationParams) -> CfdParams {
    CfdParams {
        dt: p.dt, viscosity: p.viscosity, target_inlet_velocity: p.target_inlet_velocity,
        velocity_scheme: match p.velocity_scheme { VelocityScheme::FirstOrder => 0, VelocityScheme::SecondOrder => 1, VelocityScheme::Quick => 2 },
        inlet_profile: match p.inlet_profile { InletProfile::Uniform => 0, InletProfile::Parabolic => 1 },
        pressure_solver: match p.pressure_solver { PressureSolver::Jacobi => 0, PressureSolver::Cg => 1, PressureSolver::Mgcg => 2 },
        scenario: match p.scenario { Scenario::Channel => 0, Scenario::Cavity => 1 },
    }
}

fn grid_to_c(grid: &Grid) -> CfdGrid {
    CfdGrid {
        nx: grid.nx as u64, ny: grid.ny as u64, lx: grid.lx, ly: grid.ly, dx: grid.dx, dy: grid.dy,
        has_obstacle: grid.obstacle.is_some() as i32,
        center_x: grid.obstacle.as_ref().map_or(0.0, |c| c.center_x),
        center_y: grid.obstacle.as_ref().map_or(0.0, |c| c.center_y),
        radius: grid.obstacle.as_ref().map_or(0.0, |c| c.radius),
    }
}
impl Default for SolverConsts {
    fn default() -> Self {
        let mut o = std::mem::MaybeUninit::<CfdOptions>::uninit();
        unsafe { cfd_options_default(o.as_mut_ptr()); o.assume_init().consts }
    }
}

// ---- Model: an owning handle; all fields live in HBM ----------------------------------------------------
pub struct Model { pub grid: Grid, handle: *mut CfdModel }
unsafe impl Send for Model {} // moved into exactly one solver thread, like the reference (src/model.rs:1287)

impl Model {
    pub fn new(grid: Grid, params: &SimulationParams) -> Self {
        let g = grid_to_c(&grid);
        let mut handle = std::ptr::null_mut();
        check(unsafe { cfd_model_create(&g, &params_to_c(params), &mut handle) });
        Self { grid, handle }
    }
    /// Extension: like `new`, with the solver constants spelled out (e.g. `cg_relative = 1` for the benchmarked
    /// relative stopping rule of the converged solvers; `SolverConsts::default()` = the reference's literals).
    pub fn with_consts(grid: Grid, params: &SimulationParams, consts: SolverConsts) -> Self {
        let g = grid_to_c(&grid);
        let mut o = std::mem::MaybeUninit::<CfdOptions>::uninit();
        let mut handle = std::ptr::null_mut();
        unsafe {
            cfd_options_default(o.as_mut_ptr());
            let mut o = o.assume_init();
            o.consts = consts;
            check(cfd_model_create_ex(&g, &params_to_c(params), &o, &mut handle));
        }
        Self { grid, handle }
    }
    pub fn update(&mut self) { check(unsafe { cfd_model_update(self.handle) }) }
    pub fn set_parameters(&mut self, params: &SimulationParams) {
        check(unsafe { cfd_model_set_params(self.handle, &params_to_c(params)) })
    }
    pub fn get_snapshot(&self) -> SimSnapshot {
        let (nx, ny) = (self.grid.nx, self.grid.ny);
        let (mut p, mut u, mut v) = (vec![0.0f32; nx * ny], vec![0.0f32; (nx + 1) * ny], vec![0.0f32; nx * (ny + 1)]);
        let mut dt = 0.0f32;
        check(unsafe { cfd_model_get_snapshot(self.handle, p.as_mut_ptr(), u.as_mut_ptr(), v.as_mut_ptr(), &mut dt) });
        SimSnapshot { p, u, v, dt, paused: false }
    }
    /// Extension (not in the reference): the finished `nx x ny` RGBA image of app.rs:235-404 for
    /// `egui::ColorImage::from_rgba_unmultiplied([nx, ny], &rgba)` — one third of the bytes of a snapshot and no
    /// per-pixel work on the UI thread.  `mode`: 0 pressure, 1 velocity magnitude, 2 vorticity.
    pub fn render_rgba(&self, mode: i32) -> Vec<u8> {
        let mut rgba = vec![0u8; self.grid.nx * self.grid.ny * 4];
        check(unsafe { cfd_model_render_rgba(self.handle, mode, rgba.as_mut_ptr(), std::ptr::null_mut(), std::ptr::null_mut()) });
        rgba
    }
    pub fn get_residuals(&self) -> Residuals {
        let mut r = CfdResiduals::default();
        check(unsafe { cfd_model_get_residuals(self.handle, &mut r) });
        Residuals { simulation_step: r.simulation_step as usize, simulation_time: r.simulation_time, dt: r.dt, p: r.p,
                    u: r.u, v: r.v, step_time: Duration::from_secs_f64(r.step_seconds), piso_substeps: r.piso_substeps as usize }
    }
    /// `Model::run(self)` of the reference (src/model.rs:1282-1332): the model moves into one solver thread that
    /// serves the command queue between timesteps.  Written here as a small state machine (`SolverLoop`) rather than
    /// one closure; the observable protocol is the reference's: at most one snapshot per pass over the queue,
    /// residuals after every step, ~60 Hz idling while paused.
    pub fn run(self) -> SimulationControlHandle {
        let (commands, inbox) = mpsc::channel();
        let (snapshots_out, snapshots) = mpsc::channel();
        let (residuals_out, residuals) = mpsc::channel();
        thread::spawn(move || SolverLoop { model: self, inbox, snapshots_out, residuals_out, paused: false }.serve());
        SimulationControlHandle { commands, snapshots, residuals }
    }
}

struct SolverLoop {
    model: Model,
    inbox: mpsc::Receiver<Command>,
    snapshots_out: mpsc::Sender<SimSnapshot>,
    residuals_out: mpsc::Sender<Residuals>,
    paused: bool,
}
impl SolverLoop {
    /// One pass over the pending commands; returns false when asked to stop.
    fn drain_commands(&mut self) -> bool {
        let mut snapshot_pending = false;
        while let Ok(command) = self.inbox.try_recv() {
            match command {
                Command::Stop => return false,
                Command::Pause => self.paused = true,
                Command::Resume => self.paused = false,
                Command::SetParams(p) => self.model.set_parameters(&p),
                Command::GetSnapshot => snapshot_pending = true, // coalesced: one device read-back per pass
            }
        }
        if snapshot_pending {
            let snapshot = SimSnapshot { paused: self.paused, ..self.model.get_snapshot() };
            // a closed channel means the UI dropped its handle: end the thread (the reference panics here, :1304)
            return self.snapshots_out.send(snapshot).is_ok();
        }
        true
    }
    fn serve(mut self) {
        while self.drain_commands() {
            if self.paused {
                thread::sleep(Duration::from_millis(16));
                continue;
            }
            let _t0 = Instant::now();
            self.model.update();
            if self.residuals_out.send(self.model.get_residuals()).is_err() {
                break; // handle dropped (the reference's thread dies by `unwrap()` at :1319); Drop frees the GPU
            }
        }
    }
}
impl Drop for Model {
    fn drop(&mut self) { unsafe { cfd_model_destroy(self.handle) } }
}

// ---- SimulationControlHandle: the pub API of the reference's handle (src/model.rs:65-117) over three channels ---
pub struct SimulationControlHandle {
    commands: mpsc::Sender<Command>,
    snapshots: mpsc::Receiver<SimSnapshot>,
    residuals: mpsc::Receiver<Residuals>,
}
impl SimulationControlHandle {
    fn post(&self, command: Command) {
        self.commands.send(command).expect("cfd_b200: the solver thread is gone");
    }
    pub fn stop(&self) { self.post(Command::Stop) }
    pub fn pause(&self) { self.post(Command::Pause) }
    pub fn resume(&self) { self.post(Command::Resume) }
    pub fn request_snapshot(&self) { self.post(Command::GetSnapshot) }
    pub fn set_params(&self, params: SimulationParams) { self.post(Command::SetParams(params)) }
    /// newest snapshot the solver thread has produced since the last call, if any
    pub fn get_last_available_snapshot(&self) -> Option<SimSnapshot> { self.snapshots.try_iter().last() }
    /// every `Residuals` record produced since the last call, oldest first
    pub fn get_new_log_messages(&self) -> Vec<Residuals> { self.residuals.try_iter().collect() }
}
