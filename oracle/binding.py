"""ctypes binding of oracle/d2d_oracle.c plus the numpy restatement of the
reference's table builders.  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).
"""
import ctypes as C
import json
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
GOLDEN = os.path.join(REPO, "tests", "golden")
LIB_PATH = os.path.join(HERE, "libd2d_oracle.so")

OBS_DIM = 27
MAX_WPS = 16
MAX_OBS = 64


class Body(C.Structure):
    _fields_ = [(n, C.c_double) for n in ("px", "py", "vx", "vy", "a", "w")]


class State(C.Structure):
    _fields_ = [("body", Body * 3), ("jx", C.c_double * 6), ("jy", C.c_double * 6),
                ("collided", C.c_int), ("la_latched", C.c_int), ("steps", C.c_int), ("pad_", C.c_int),
                ("path_error", C.c_double), ("total_reward", C.c_double)]


class Path(C.Structure):
    _fields_ = [("n_wps", C.c_int), ("pad_", C.c_int),
                ("wx", C.c_double * MAX_WPS), ("wy", C.c_double * MAX_WPS), ("us", C.c_double * MAX_WPS),
                ("xp", (C.c_double * 3) * MAX_WPS), ("yp", (C.c_double * 3) * MAX_WPS)]


class Obstacles(C.Structure):
    _fields_ = [("n", C.c_int), ("pad_", C.c_int),
                ("x", C.c_double * MAX_OBS), ("y", C.c_double * MAX_OBS), ("r", C.c_double * MAX_OBS)]


class Config(C.Structure):
    _fields_ = [(n, C.c_double) for n in (
        "screen_x", "screen_y", "lookahead", "danger_range", "danger_angle", "abs_inv_CA_min_rew",
        "PA_band_edge", "PA_scale", "PP_vel_scale", "PP_rew_max", "PP_rew_min",
        "rew_collision", "reach_end_radius", "rew_reach_end", "AA_angle", "AA_band", "rew_AA")] + [
        ("use_Lambda", C.c_int), ("n_steps", C.c_int)] + [(n, C.c_double) for n in (
            "mass_f", "mass_m", "width", "height", "gravity_x", "gravity_y", "dt", "force_scale")] + [
        ("iterations", C.c_int), ("pad_", C.c_int)]


class StepOut(C.Structure):
    _fields_ = [("obs", C.c_double * OBS_DIM), ("reward", C.c_double)] + [(n, C.c_double) for n in (
        "r_CA", "r_PA", "r_PP", "r_coll", "r_end", "r_AA", "dist_closest_obs", "ape", "u_star",
        "closest_x", "closest_y", "la_x", "la_y")] + [
        ("nfev", C.c_int), ("near_idx", C.c_int * 3),
        ("end_collision", C.c_int), ("end_reached", C.c_int), ("end_time", C.c_int), ("end_alpha", C.c_int),
        ("done", C.c_int), ("pad_", C.c_int)]


_lib = None


def build(force=False):
    """Compile libd2d_oracle.so with the committed Makefile (gcc, strict IEEE)."""
    src = [os.path.join(HERE, f) for f in ("d2d_oracle.c", "d2d_oracle.h", "Makefile")]
    stale = (not os.path.exists(LIB_PATH)) or any(os.path.getmtime(s) > os.path.getmtime(LIB_PATH) for s in src)
    if force or stale:
        subprocess.check_call(["make", "-C", HERE, "-B", "libd2d_oracle.so"], stdout=subprocess.DEVNULL)
    return LIB_PATH


def load(path):
    """dlopen an oracle build and declare its prototypes (tests also load variant builds, e.g.
    -DD2O_SPACE_DAMPING=0.9, to show that the genuine trajectories tell them apart)."""
    L = C.CDLL(path)
    L.d2o_default_config.argtypes = [C.POINTER(Config)]
    L.d2o_moment_for_box.restype = C.c_double
    L.d2o_moment_for_box.argtypes = [C.c_double] * 3
    L.d2o_spawn.argtypes = [C.POINTER(State), C.c_double, C.c_double, C.c_double, C.POINTER(Config)]
    L.d2o_space_step.restype = C.c_int
    L.d2o_space_step.argtypes = [C.POINTER(State), C.POINTER(Obstacles), C.POINTER(Config), C.c_double, C.c_double]
    L.d2o_path_eval.argtypes = [C.POINTER(Path), C.c_double, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    L.d2o_path_gradient.argtypes = L.d2o_path_eval.argtypes
    L.d2o_closest_u.restype = C.c_double
    L.d2o_closest_u.argtypes = [C.POINTER(Path), C.c_double, C.c_double, C.POINTER(C.c_int), C.POINTER(C.c_double)]
    L.d2o_observe.argtypes = [C.POINTER(State), C.POINTER(Path), C.POINTER(Obstacles), C.POINTER(Config), C.POINTER(StepOut)]
    L.d2o_env_step.argtypes = [C.POINTER(State), C.POINTER(Path), C.POINTER(Obstacles), C.POINTER(Config),
                               C.c_double, C.c_double, C.POINTER(StepOut)]
    L.d2o_env_step_batch.argtypes = [C.c_int, C.POINTER(State), C.POINTER(Path), C.c_int, C.POINTER(Obstacles),
                                     C.c_int, C.POINTER(Config), C.POINTER(C.c_float), C.POINTER(StepOut)]
    L.d2o_philox4x32_10.argtypes = [C.POINTER(C.c_uint), C.POINTER(C.c_uint), C.POINTER(C.c_uint)]
    L.d2o_after_physics.argtypes = [C.POINTER(State), C.POINTER(Path), C.POINTER(Obstacles), C.POINTER(Config),
                                    C.POINTER(StepOut)]
    if hasattr(L, "d2o_frame_velocity_rollout"):
        dp = C.POINTER(C.c_double)
        L.d2o_frame_velocity_rollout.restype = None
        L.d2o_frame_velocity_rollout.argtypes = [C.POINTER(State), dp, C.c_int, C.POINTER(Config), dp, dp, C.c_double]
    L.d2o_frame_touches.restype = C.c_int
    L.d2o_frame_touches.argtypes = [C.POINTER(Body), C.POINTER(Obstacles), C.POINTER(Config)]
    return L


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = load(LIB_PATH)
    return _lib


def default_config(**overrides):
    c = Config()
    lib().d2o_default_config(C.byref(c))
    for k, v in overrides.items():
        if not hasattr(c, k):
            raise KeyError(k)
        setattr(c, k, v)
    return c


def qpmi_tables(wps):
    """QPMI2D.__init__ (predef_path.py:10-50): chord-length parameters and, for
    every interior waypoint, the quadratic through its two neighbours obtained
    the way the reference does it (np.linalg.inv(U) @ coords)."""
    wps = np.asarray(wps, dtype=np.float64)
    seg = np.cumsum(np.sqrt(np.sum(np.diff(wps, axis=0) ** 2, axis=1)))
    us = np.array([0, *seg[:]])
    xp, yp = [], []
    for n in range(1, len(wps) - 1):
        U = np.vstack([np.hstack([us[n - 1] ** 2, us[n - 1], 1]),
                       np.hstack([us[n] ** 2, us[n], 1]),
                       np.hstack([us[n + 1] ** 2, us[n + 1], 1])])
        xp.append(np.linalg.inv(U).dot(np.array([wps[n - 1][0], wps[n][0], wps[n + 1][0]])))
        yp.append(np.linalg.inv(U).dot(np.array([wps[n - 1][1], wps[n][1], wps[n + 1][1]])))
    return us, np.array(xp), np.array(yp)


def make_path(wps, us=None, xp=None, yp=None):
    wps = np.asarray(wps, dtype=np.float64)
    if us is None:
        us, xp, yp = qpmi_tables(wps)
    n = len(wps)
    assert 3 <= n <= MAX_WPS
    p = Path()
    p.n_wps = n
    for i in range(n):
        p.wx[i], p.wy[i], p.us[i] = wps[i][0], wps[i][1], us[i]
    for i in range(n - 2):
        for k in range(3):
            p.xp[i][k] = xp[i][k]
            p.yp[i][k] = yp[i][k]
    return p


def make_obstacles(circles):
    o = Obstacles()
    o.n = len(circles)
    assert o.n <= MAX_OBS
    for i, (x, y, r) in enumerate(circles):
        o.x[i], o.y[i], o.r[i] = x, y, r
    return o


_scen_cache = None

# Test-mode spawn rectangles (env:221-311): x ~ U(50, xmax), y ~ U(ymin, ymax), W=H=1300.
SCENARIO_SPAWN = {
    "perpendicular": (50.0, 1300 / 2 - 100, 50.0, 1300 - 100.0),
    "parallel": (50.0, 1300 / 2 - 300, 150.0, 1300 - 300.0),
    "S_parallel": (50.0, 1300 / 2 - 300, 150.0, 1300 - 300.0),
    "corridor": (50.0, 1300 / 2 - 400, 150.0, 1300 - 300.0),
    "S_corridor": (50.0, 1300 / 2 - 450, 150.0, 1300 - 300.0),
    "large": (50.0, 1300 / 2 - 1300 / 4 - 50, 150.0, 1300 - 300.0),
    "impossible": (50.0, 1300 / 2, 150.0, 1300 - 300.0),
}


def load_scenario(name):
    """Scenario tables exactly as the reference's test_scenarios.py produced them
    (tests/golden/scenarios.json)."""
    global _scen_cache
    if _scen_cache is None:
        with open(os.path.join(GOLDEN, "scenarios.json")) as f:
            _scen_cache = json.load(f)
    r = _scen_cache[name]
    return make_path(r["wps"], r["us"], r["x_params"], r["y_params"]), make_obstacles(r["obstacles"])


def spawn(x, y, angle, cfg):
    s = State()
    lib().d2o_spawn(C.byref(s), x, y, angle, C.byref(cfg))
    return s


def env_step(state, path, obstacles, cfg, a_left, a_right):
    out = StepOut()
    lib().d2o_env_step(C.byref(state), C.byref(path), C.byref(obstacles), C.byref(cfg),
                       float(np.float32(a_left)), float(np.float32(a_right)), C.byref(out))
    return out


def env_step_batch(states, paths, obstacles, cfg, actions, L=None):
    """states: (State * n); paths / obstacles: one struct (shared) or an array of n."""
    n = len(states)
    actions = np.ascontiguousarray(actions, dtype=np.float32).reshape(n, 2)
    outs = (StepOut * n)()
    p_arr = isinstance(paths, C.Array)
    o_arr = isinstance(obstacles, C.Array)
    (L or lib()).d2o_env_step_batch(
        n, states,
        paths if p_arr else C.pointer(paths), 1 if p_arr else 0,
        obstacles if o_arr else C.pointer(obstacles), 1 if o_arr else 0,
        C.byref(cfg), actions.ctypes.data_as(C.POINTER(C.c_float)), outs)
    return outs


def closest_u(path, px, py):
    nfev = C.c_int()
    fval = C.c_double()
    u = lib().d2o_closest_u(C.byref(path), px, py, C.byref(nfev), C.byref(fval))
    return u, nfev.value, fval.value


def path_eval(path, u):
    x, y = C.c_double(), C.c_double()
    lib().d2o_path_eval(C.byref(path), u, C.byref(x), C.byref(y))
    return x.value, y.value


def path_gradient(path, u):
    x, y = C.c_double(), C.c_double()
    lib().d2o_path_gradient(C.byref(path), u, C.byref(x), C.byref(y))
    return x.value, y.value


def philox4x32_10(ctr, key):
    c = (C.c_uint * 4)(*ctr)
    k = (C.c_uint * 2)(*key)
    o = (C.c_uint * 4)()
    lib().d2o_philox4x32_10(c, k, o)
    return list(o)
