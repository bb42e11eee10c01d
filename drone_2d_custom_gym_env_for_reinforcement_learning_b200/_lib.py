"""ctypes binding of the C ABI declared in include/drone2d.h.

There is no CPU fallback: if the CUDA extension cannot be loaded this module
raises and nothing in the package works."""
import ctypes as C
import os

from . import _build

OBS_DIM = 27
ACT_DIM = 2
MAX_WPS = 16
MAX_OBS_SHARED = 64
MAX_OBS_ENV = 32
STATE_DOUBLES = 36
INFO_DOUBLES = 16
STATS_DOUBLES = 16
ABI_VERSION = 1
MODE_TEST, MODE_CURRICULUM = 0, 1

EXPORTS = [
    "d2d_abi_version", "d2d_last_error", "d2d_default_config", "d2d_state_bytes", "d2d_create", "d2d_destroy",
    "d2d_set_scenario", "d2d_set_tables", "d2d_get_tables", "d2d_reset", "d2d_step", "d2d_step_host",
    "d2d_step_host_ex",
    "d2d_set_taps", "d2d_set_info_done_only", "d2d_get_state", "d2d_set_state", "d2d_episode_stats", "d2d_episode_stats_dev",
    "d2d_debug_closest_u", "d2d_debug_role_cycles", "d2d_debug_skip_physics", "d2d_launch_count", "d2d_set_stage", "d2d_gae",
    "d2d_policy_weights_bytes", "d2d_policy_pack", "d2d_policy_act", "d2d_rollout",
    "d2d_policy_act_ex", "d2d_policy_value", "d2d_rollout_collect", "d2d_set_flight_path_ring",
]


class Config(C.Structure):
    """struct d2d_config (include/drone2d.h)."""
    _fields_ = [(n, C.c_int32) for n in (
        "abi_version", "mode", "stage", "n_steps", "n_wps", "spawn_corner_lo", "spawn_corner_hi",
        "random_path_spawn", "use_Lambda", "iterations")] + [(n, C.c_double) for n in (
            "screen_x", "screen_y", "path_segment_length", "lookahead",
            "danger_range", "danger_angle", "abs_inv_CA_min_rew", "PA_band_edge", "PA_scale",
            "PP_vel_scale", "PP_rew_max", "PP_rew_min", "rew_collision", "reach_end_radius", "rew_reach_end",
            "AA_angle", "AA_band", "rew_AA", "stage_spawn_chance",
            "spawn_xmin", "spawn_xmax", "spawn_ymin", "spawn_ymax",
            "mass_f", "mass_m", "width", "height", "gravity_x", "gravity_y", "dt", "force_scale")]


class D2DError(RuntimeError):
    pass


_lib = None


def lib_path():
    return _build.LIB


def load():
    """Load (building first if the sources are newer) libdrone2d.so."""
    global _lib
    if _lib is not None:
        return _lib
    path = os.environ.get("D2D_LIB") or _build.build()     # D2D_LIB: load a specific build (kernel tuning)
    if not os.path.exists(path):
        raise D2DError("CUDA extension %s is missing" % path)
    L = C.CDLL(path)
    vp, i64, u64, i32 = C.c_void_p, C.c_int64, C.c_uint64, C.c_int
    dp = C.POINTER(C.c_double)
    L.d2d_abi_version.restype = i32
    L.d2d_last_error.restype = C.c_char_p
    L.d2d_default_config.argtypes = [C.POINTER(Config)]
    L.d2d_state_bytes.argtypes = [C.POINTER(Config), i64, C.POINTER(C.c_size_t)]
    L.d2d_create.argtypes = [C.POINTER(Config), i64, i64, u64, i32, vp, C.c_size_t, C.POINTER(vp)]
    L.d2d_destroy.argtypes = [vp]
    L.d2d_set_scenario.argtypes = [vp, i32, dp, dp, dp, dp, i32, dp, vp]
    L.d2d_set_tables.argtypes = [vp, dp, dp, dp, dp, C.POINTER(C.c_int32), dp]
    L.d2d_get_tables.argtypes = [vp, dp, dp, dp, dp, C.POINTER(C.c_int32), dp]
    L.d2d_reset.argtypes = [vp, vp, vp, vp]
    L.d2d_step.argtypes = [vp, vp, vp, vp, vp, vp, i32, vp]
    L.d2d_step_host.argtypes = [vp, vp, vp, vp, vp, i32, vp]
    L.d2d_step_host_ex.argtypes = [vp, vp, vp, vp, vp, vp, i32, vp]
    L.d2d_set_taps.argtypes = [vp, vp, vp]
    L.d2d_set_info_done_only.argtypes = [vp, i32]
    L.d2d_set_flight_path_ring.argtypes = [vp, vp, i32]
    L.d2d_get_state.argtypes = [vp, dp]
    L.d2d_set_state.argtypes = [vp, dp]
    L.d2d_episode_stats.argtypes = [vp, dp, i32]
    L.d2d_episode_stats_dev.argtypes = [vp, C.POINTER(vp)]
    L.d2d_debug_closest_u.argtypes = [vp, vp, vp, vp, vp]
    L.d2d_debug_role_cycles.argtypes = [vp, C.POINTER(C.c_uint64), i32]
    L.d2d_debug_skip_physics.argtypes = [vp, i32]
    L.d2d_set_stage.argtypes = [vp, i32, C.c_double]
    L.d2d_gae.argtypes = [vp, vp, vp, vp, vp, i32, i64, C.c_float, C.c_float, vp, vp, vp]
    L.d2d_policy_pack.argtypes = [vp] * 8
    L.d2d_policy_act.argtypes = [vp, vp, vp, vp, i32, u64, vp]
    L.d2d_rollout.argtypes = [vp, vp, vp, vp, vp, vp, i32, i32, u64, vp]
    L.d2d_policy_act_ex.argtypes = [vp, vp, vp, vp, vp, vp, i32, u64, vp]
    L.d2d_policy_value.argtypes = [vp, vp, vp, vp, i64, vp]
    L.d2d_rollout_collect.argtypes = [vp, vp, vp, i32] + [vp] * 9 + [C.c_float, C.c_float, u64, vp]
    L.d2d_launch_count.argtypes = [vp]
    L.d2d_launch_count.restype = i64
    for name in EXPORTS:
        if name not in ("d2d_last_error", "d2d_launch_count"):
            getattr(L, name).restype = i32
    if L.d2d_abi_version() != ABI_VERSION:
        raise D2DError("ABI version mismatch: library %d, binding %d" % (L.d2d_abi_version(), ABI_VERSION))
    _lib = L
    return L


def check(rc):
    if rc != 0:
        raise D2DError("libdrone2d: error %d: %s" % (rc, load().d2d_last_error().decode()))


def default_config():
    c = Config()
    check(load().d2d_default_config(C.byref(c)))
    return c
