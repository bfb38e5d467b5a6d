"""ctypes mirror of include/cfd_b200.h (PODs and constants only; no library is loaded here)."""
import ctypes as C

CFD_ABI_VERSION = 3

CFD_OK = 0
CFD_ERR_INVALID_ARGUMENT = 1
CFD_ERR_CUDA = 2
CFD_ERR_UNSUPPORTED = 3
CFD_ERR_NCCL = 4
CFD_ERR_PEER_TIMEOUT = 5

SCHEME_FIRST_ORDER, SCHEME_SECOND_ORDER, SCHEME_QUICK = 0, 1, 2
INLET_UNIFORM, INLET_PARABOLIC = 0, 1
SOLVER_JACOBI, SOLVER_CG, SOLVER_MGCG = 0, 1, 2
SCENARIO_CHANNEL, SCENARIO_CAVITY = 0, 1

FIELD_P, FIELD_U, FIELD_V, FIELD_U_STAR, FIELD_V_STAR, FIELD_RHS, FIELD_P_PRIME = range(7)
FIELD_U_OLD, FIELD_V_OLD, FIELD_MASK_U, FIELD_MASK_V, FIELD_MG_GUESS, FIELD_MG_LAST, FIELD_MG_LAST2 = 7, 8, 9, 10, 11, 12, 13
FIELD_MG_Z = 14  # read-only inspection: z of the last V-cycle
FIELD_NAMES = {
    FIELD_P: "p", FIELD_U: "u", FIELD_V: "v", FIELD_U_STAR: "u_star", FIELD_V_STAR: "v_star",
    FIELD_RHS: "rhs", FIELD_P_PRIME: "p_prime", FIELD_U_OLD: "u_old", FIELD_V_OLD: "v_old",
    FIELD_MASK_U: "mask_u", FIELD_MASK_V: "mask_v", FIELD_MG_GUESS: "mg_guess", FIELD_MG_LAST: "mg_last", FIELD_MG_LAST2: "mg_last2", FIELD_MG_Z: "mg_z",
}

FLAG_NO_GRAPH = 1
FLAG_BASELINE_SWEEP = 2
FLAG_REGISTER_SWEEP = 4
FLAG_BULK_SWEEP = 8
FLAG_SWEEP4 = 16
FLAG_NCCL_EXCHANGE = 32
FLAG_TEMPORAL = 64
FLAG_PERSISTENT_SWEEP = 128
FLAG_MG_NO_BOTTOM_KERNEL = 256
FLAG_PEER_EXCHANGE = 512
FLAG_MG_UNFUSED = 1024
FLAG_PEER_STRIPS = 2048


class CfdGrid(C.Structure):
    _fields_ = [("nx", C.c_uint64), ("ny", C.c_uint64),
                ("lx", C.c_float), ("ly", C.c_float), ("dx", C.c_float), ("dy", C.c_float),
                ("has_obstacle", C.c_int32),
                ("center_x", C.c_float), ("center_y", C.c_float), ("radius", C.c_float)]


class CfdParams(C.Structure):
    _fields_ = [("dt", C.c_float), ("viscosity", C.c_float), ("target_inlet_velocity", C.c_float),
                ("velocity_scheme", C.c_int32), ("inlet_profile", C.c_int32),
                ("pressure_solver", C.c_int32), ("scenario", C.c_int32)]


class CfdSolverConsts(C.Structure):
    _fields_ = [("ramp_up_steps", C.c_int32), ("jacobi_iterations", C.c_int32),
                ("outer_rounds", C.c_int32), ("cg_max_iterations", C.c_int32),
                ("jacobi_omega", C.c_double), ("pressure_tolerance", C.c_double),
                ("outer_tolerance", C.c_double), ("cfl", C.c_double), ("cg_tolerance", C.c_double),
                ("mg_omega", C.c_double), ("mg_smoothing", C.c_int32), ("mg_warm_start", C.c_int32),
                ("cg_relative", C.c_int32), ("adaptive_substeps", C.c_int32)]


class CfdOptions(C.Structure):
    _fields_ = [("precision", C.c_int32), ("device", C.c_int32), ("rank", C.c_int32),
                ("world_size", C.c_int32), ("nccl_unique_id", C.c_void_p), ("flags", C.c_uint32),
                ("consts", CfdSolverConsts)]


class CfdResiduals(C.Structure):
    _fields_ = [("simulation_step", C.c_uint64),
                ("simulation_time", C.c_float), ("dt", C.c_float), ("p", C.c_float),
                ("u", C.c_float), ("v", C.c_float),
                ("step_seconds", C.c_double), ("piso_substeps", C.c_uint64),
                ("jacobi_calls", C.c_uint64), ("sweeps", C.c_uint64),
                ("simulation_time_f64", C.c_double), ("dt_f64", C.c_double), ("p_f64", C.c_double),
                ("u_f64", C.c_double), ("v_f64", C.c_double),
                ("p_rel_f64", C.c_double), ("rhs_rms_f64", C.c_double), ("first_solve_iterations", C.c_uint64)]

    def as_dict(self):
        return {name: getattr(self, name) for name, _ in self._fields_}
