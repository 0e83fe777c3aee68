"""ctypes binding of libsph_b200.so (include/sph_b200.h).

There is no CPU path: importing this module never needs a GPU (so symbol/ABI checks run anywhere),
but `create()` fails loudly when the CUDA library or a CUDA device is missing.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# SPH_B200_LIB: developer knob to A/B a variant build of the same library (never a different backend)
LIB_PATH = os.environ.get("SPH_B200_LIB") or os.path.join(_HERE, "_build", "libsph_b200.so")

SOLVER_IDS = {"wcsph": 0, "pcisph": 1, "iisph": 2, "dfsph": 3, "pbf": 4}

# enum SphField
F_FLUID_POS, F_FLUID_VEL, F_BOUNDARY_POS, F_RIGID_POS, F_RIGID_VEL, F_RIGID_FORCE, F_FLUID_ACC, F_RIGID_VERTICES, F_FLUID_GID = range(9)
(F_RHO, F_ALPHA, F_RHO_DERIVATIVE, F_RHO_ADV, F_VEL_ADV, F_CELL1D, F_NEIGHBOR_COUNT,
 F_BOUNDARY_NEIGHBOR_COUNT, F_PRESSURE, F_FORCE_A, F_FORCE_B, F_SCALAR_A, F_SCALAR_B, F_SCALAR_C,
 F_VEC_A, F_VEC_B, F_VEC_C, F_PAYLOAD_1, F_PAYLOAD_3, F_POS_RHO) = range(16, 36)
F_CELL_START, F_SORTED_INDEX, F_BOUNDARY_CELL_START, F_BOUNDARY_SORTED_INDEX = range(64, 68)

# enum SphPhase
PH_BUILD_GRID = 0
PH_DF_INITIALIZE, PH_DF_DIVERGENCE, PH_DF_EXT_FORCE_VEL_ADV, PH_DF_DENSITY, PH_DF_POSITION = range(10, 15)
PH_WC_PRESSURE, PH_WC_KINEMATIC = 20, 21
PH_PC_EXT_FORCE, PH_PC_ITERATION, PH_PC_INTEGRATION = 30, 31, 32
PH_II_PREDICT_ADVECTION, PH_II_PRESSURE_SOLVE, PH_II_INTEGRATION = 40, 41, 42
PH_PBF_PREDICT, PH_PBF_LAMBDA, PH_PBF_DELTA_POS, PH_PBF_UPDATE_POS = 50, 51, 52, 53
PH_WRITEBACK = 90
# the same steps one sweep at a time (selfcheck.py)
PH_BUILD_LISTS = 100
PH_DF_WARM_START, PH_DF_DRHO_FIRST, PH_DF_DIV_VEL, PH_DF_DIV_DRHO, PH_DF_DEN_RHO, PH_DF_DEN_VEL = range(110, 116)
PH_WC_EOS, PH_WC_FORCE = 120, 121
PH_PC_PREDICT, PH_PC_RHO_FIRST, PH_PC_PRESS_FORCE, PH_PC_RHO = range(130, 134)
PH_II_ADVECT, PH_II_AII, PH_II_SOLVE_BEGIN, PH_II_DIJ, PH_II_UPDATE = range(140, 145)

# kernel classes of sph_profile_end (csrc/sph_internal.h)
KERNEL_CLASSES = ["grid", "lists", "df_warm_start", "df_drho", "df_div_iter", "df_ext_force", "df_rho_adv",
                  "df_vel_adv_iter", "df_position", "ctl", "wc_force", "wc_kinematic", "pc_ext", "pc_predict",
                  "pc_rho", "pc_force", "pc_integrate", "ii_adv", "ii_aii", "ii_dij", "ii_update", "ii_integrate",
                  "rigid", "other", "mg_exchange", "mg_begin_step", "mg_wait"]


# SphStats.error_flags (csrc/sph_common.cuh): conditions under which the run has left the reference's physics
ERROR_BITS = {1: "a particle left the grid and was clamped into it (PS:393-395)",
              2: "a fluid neighbour list overflowed max_neighbors (the reference has no cap): raise solver.max_neighbors",
              4: "a boundary neighbour list overflowed its capacity",
              8: "the DFSPH density loop hit the library's 1000-pass cap (the reference's loop has none, DF:225)",
              16: "a non-finite value appeared in the solver state",
              32: "a peer rank stopped answering a halo exchange (multi-GPU)",
              64: "an index check of the bounds-checked build failed (internal error)"}


def decode_error_flags(flags):
    return [msg for bit, msg in ERROR_BITS.items() if flags & bit]


class SphCamera(ctypes.Structure):
    _fields_ = [("pos", ctypes.c_double * 3), ("look_at", ctypes.c_double * 3), ("up", ctypes.c_double * 3),
                ("fov_y_deg", ctypes.c_double), ("light_pos", ctypes.c_double * 3), ("ambient", ctypes.c_double),
                ("background", ctypes.c_uint8 * 4)]


RENDER_FLUID, RENDER_RIGID = 1, 2


class SphLattice(ctypes.Structure):
    _fields_ = [("particle_radius", ctypes.c_double), ("start_pos", ctypes.c_double * 3),
                ("water_size", ctypes.c_double * 3), ("box_min", ctypes.c_double * 3), ("box_max", ctypes.c_double * 3)]


class SphConfig(ctypes.Structure):
    _fields_ = [
        ("box_min", ctypes.c_double * 3),
        ("box_max", ctypes.c_double * 3),
        ("particle_radius", ctypes.c_double),
        ("gravity", ctypes.c_double),
        ("delta_time", ctypes.c_double),
        ("boundary_handle", ctypes.c_int32),
        ("fs_couple", ctypes.c_int32),
        ("solver", ctypes.c_int32),
        ("n_fluid", ctypes.c_int32),
        ("n_boundary", ctypes.c_int32),
        ("n_rigid", ctypes.c_int32),
        ("active_rigid", ctypes.c_int32),
        ("grid_num", ctypes.c_int32 * 3),
        ("max_neighbors", ctypes.c_int32),
        ("max_boundary_neighbors", ctypes.c_int32),
        ("strict", ctypes.c_int32),
        ("n_ghost_capacity", ctypes.c_int32),
        ("rigid_rho", ctypes.c_double),
        ("use_graph", ctypes.c_int32),
        ("reserved", ctypes.c_int32),
    ]


class SphStats(ctypes.Structure):
    _fields_ = [
        ("delta_time", ctypes.c_float),
        ("ps_delta_time", ctypes.c_float),
        ("simulate_cnt", ctypes.c_int32),
        ("error_flags", ctypes.c_int32),
        ("div_iters", ctypes.c_int32),
        ("div_first_err", ctypes.c_float),
        ("div_err", ctypes.c_float),
        ("den_iters", ctypes.c_int32),
        ("den_err", ctypes.c_float),
        ("pc_iters", ctypes.c_int32),
        ("pc_err", ctypes.c_float),
        ("ii_iters", ctypes.c_int32),
        ("ii_residual", ctypes.c_float),
        ("max_neighbors_seen", ctypes.c_int32),
        ("max_boundary_neighbors_seen", ctypes.c_int32),
        ("pc_delta", ctypes.c_float),
        ("pc_max_index", ctypes.c_int32),
        ("kernel_launches", ctypes.c_int32),
        ("div_active", ctypes.c_int32),
        ("den_active", ctypes.c_int32),
        ("loop_active", ctypes.c_int32),
    ]


class SphRigidInfo(ctypes.Structure):
    _fields_ = [("centroid", ctypes.c_float * 3), ("inertia", ctypes.c_float * 9), ("inertia_inv", ctypes.c_float * 9),
                ("vel", ctypes.c_float * 3), ("omega", ctypes.c_float * 3), ("alpha", ctypes.c_float * 3),
                ("acc", ctypes.c_float * 3), ("attitude", ctypes.c_float * 3), ("force_sum", ctypes.c_float * 3),
                ("torque", ctypes.c_float * 3), ("mass", ctypes.c_float), ("delta_time", ctypes.c_float),
                ("collision_cnt", ctypes.c_int32), ("simulate_cnt", ctypes.c_int32),
                ("max_surface_vel", ctypes.c_float), ("reserved", ctypes.c_int32)]


# every symbol include/sph_b200.h declares: (name, restype, argtypes)
_vp, _i, _f = ctypes.c_void_p, ctypes.c_int, ctypes.c_float
_fp = ctypes.POINTER(ctypes.c_float)
PROTOTYPES = [
    ("sph_create", _i, [ctypes.POINTER(SphConfig), _i, ctypes.POINTER(_vp)]),
    ("sph_destroy", _i, [_vp]),
    ("sph_last_error", ctypes.c_char_p, [_vp]),
    ("sph_abi_version", _i, []),
    ("sph_bind", _i, [_vp, _i, _vp, ctypes.c_size_t]),
    ("sph_init_boundary", _i, [_vp, _vp]),
    ("sph_init_rigid", _i, [_vp, _vp]),
    ("sph_rigid_set_state", _i, [_vp, ctypes.POINTER(SphRigidInfo)]),
    ("sph_pcisph_precompute", _i, [_vp, _vp]),
    ("sph_pcisph_delta", _i, [_vp, _i, _vp]),
    ("sph_pcisph_set_delta", _i, [_vp, ctypes.c_float, _i, _vp]),
    ("sph_step", _i, [_vp, _i, _vp]),
    ("sph_phase", _i, [_vp, _i, _vp]),
    ("sph_rigid_step", _i, [_vp, _vp]),
    ("sph_rigid_state", _i, [_vp, _vp]),
    ("sph_set_delta_time", _i, [_vp, _f, _vp]),
    ("sph_fetch", _i, [_vp, _i, _vp, ctypes.c_size_t, _vp]),
    ("sph_init_fluid_lattice", _i, [ctypes.POINTER(SphLattice), ctypes.c_longlong, _vp, ctypes.c_size_t, _vp, _i, _vp]),
    ("sph_init_boundary_shell", _i, [ctypes.POINTER(SphLattice), ctypes.c_size_t, _vp, _i, _vp]),
    ("sph_visualize", _i, [_vp, _i, _vp, _i, ctypes.c_size_t, _vp]),
    ("sph_upload_state", _i, [_vp, _vp, _vp, _vp]),
    ("sph_upload_state_xyz", _i, [_vp, _vp, _vp, _vp]),
    ("sph_download_state_xyz", _i, [_vp, _vp, _vp, _vp]),
    ("sph_download_state", _i, [_vp, _vp, _vp, _vp]),
    ("sph_read_stats", _i, [_vp, ctypes.POINTER(SphStats)]),
    ("sph_copy_work_state", _i, [_vp, _vp, _vp]),
    ("sph_render", _i, [_vp, ctypes.POINTER(SphCamera), _i, _i, _i, _vp, _i, _vp, _i, _vp, _vp, _vp]),
    ("sph_profile_begin", _i, [_vp]),
    ("sph_profile_end", _i, [_vp, _fp, ctypes.POINTER(ctypes.c_int32), _i]),
    ("sph_comm_unique_id", _i, [ctypes.c_char_p]),
    ("sph_comm_init", _i, [_vp, ctypes.c_char_p, _i, _i, _i, _i]),
    ("sph_comm_info", _i, [_vp, ctypes.POINTER(ctypes.c_int32)]),
    ("sph_set_counts", _i, [_vp, _i, _i]),
]

_lib = None


class SphError(RuntimeError):
    pass


def load():
    """Load the CUDA library.  No fallback: a missing build is an error."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise SphError(
                "libsph_b200.so is not built (%s). Run `python -c 'import __graft_entry__ as g; g.build()'` "
                "or `python cfd_taichi_b200/build.py`. There is no CPU fallback." % LIB_PATH)
        L = ctypes.CDLL(LIB_PATH)
        for name, res, args in PROTOTYPES:
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(rc, handle=None):
    if rc != 0:
        msg = load().sph_last_error(handle)
        raise SphError("sph_b200 error %d: %s" % (rc, msg.decode() if msg else "?"))
