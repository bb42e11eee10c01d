"""Python port of the reference environment (TEST INFRASTRUCTURE ONLY).

`RefPortEnv` mirrors `Drone2dEnv` (drone_2d_env.py:22-162, 182-392, 394-615,
908-912) call for call: Python glue, numpy/`random` global RNG draws in the
reference's order, `scipy.optimize.fminbound` for the closest point (called
twice per step exactly like env:730,737), a Python loop over obstacles, and a C
library for the rigid-body step (the reference calls Chipmunk through pymunk's
cffi layer; here the restated step of d2d_oracle.c is called through ctypes).
It is what `bench.py --impl reference` and the `cpu_baseline` leg time, and what
the oracle's C step is cross-checked against.

`reset_tables` / `sample_*` restate the reset distributions (env:199-372,
predef_path.py:307-363, obstacles.py:58-89) on an explicit draw source so the
same code can consume numpy's global RNG (baseline), recorded draws (golden
tests) or the Philox stream the CUDA auto-reset is specified on.
"""
import ctypes as C
import math
import random

import numpy as np

from . import binding as B

STAGES = ("stage_1", "stage_2", "stage_3", "stage_4", "stage_5")
TEST_SCENARIOS = tuple(B.SCENARIO_SPAWN)


# --------------------------------------------------------------------------- #
# draw sources                                                                #
# --------------------------------------------------------------------------- #
class NumpyDraws:
    """The reference's RNGs: Python `random` for corner / spawn pose, numpy's
    global generator for waypoints and obstacles."""

    def randint(self, a, b):
        return random.randint(a, b)

    def py_uniform(self, a, b):
        return random.uniform(a, b)

    def uniform(self, a, b):
        return np.random.uniform(a, b)

    def normal(self, mu, sd):
        return np.random.normal(mu, sd)

    def binomial1(self, p):
        return np.random.binomial(1, p)


class RecordedDraws:
    """Replays the raw values logged by tests/golden/make_golden.py."""

    def __init__(self, log):
        self.log = list(log)
        self.i = 0

    def _next(self, kind, a, b):
        k, ra, rb, v = self.log[self.i]
        self.i += 1
        assert k == kind and ra == a and rb == b, (k, ra, rb, kind, a, b)
        return v

    def uniform(self, a, b):
        return self._next("uniform", a, b)

    def normal(self, mu, sd):
        return self._next("normal", mu, sd)


# --------------------------------------------------------------------------- #
# path (predef_path.py) in Python -- used by the samplers and RefPortEnv      #
# --------------------------------------------------------------------------- #
class PortPath:
    """QPMI2D (predef_path.py:9-266)."""

    def __init__(self, wps):
        self.waypoints = np.asarray(wps, dtype=np.float64)
        self.us, self.x_params, self.y_params = B.qpmi_tables(self.waypoints)
        self.length = self.us[-1]

    def get_u_index(self, u):
        n = 0
        while n < len(self.us) - 1:
            if u <= self.us[n + 1]:
                break
            n += 1
        return n

    def __call__(self, u):
        us, xp, yp = self.us, self.x_params, self.y_params
        if u >= us[0] and u <= us[1]:
            x = xp[0][0] * u ** 2 + xp[0][1] * u + xp[0][2]
            y = yp[0][0] * u ** 2 + yp[0][1] * u + yp[0][2]
        elif (u >= us[-2] - 0.001 and u <= us[-1]) or self.get_u_index(u) == len(us) - 1:
            x = xp[-1][0] * u ** 2 + xp[-1][1] * u + xp[-1][2]
            y = yp[-1][0] * u ** 2 + yp[-1][1] * u + yp[-1][2]
        else:
            n = self.get_u_index(u)
            mu_r = (u - us[n]) / (us[n + 1] - us[n])
            mu_f = (us[n + 1] - u) / (us[n + 1] - us[n])
            x1 = xp[n - 1][0] * u ** 2 + xp[n - 1][1] * u + xp[n - 1][2]
            y1 = yp[n - 1][0] * u ** 2 + yp[n - 1][1] * u + yp[n - 1][2]
            x2 = xp[n][0] * u ** 2 + xp[n][1] * u + xp[n][2]
            y2 = yp[n][0] * u ** 2 + yp[n][1] * u + yp[n][2]
            x = mu_r * x2 + mu_f * x1
            y = mu_r * y2 + mu_f * y1
        return np.array([x, y])

    def calculate_gradient(self, u):
        us, xp, yp = self.us, self.x_params, self.y_params
        if u >= us[0] and u <= us[1]:
            dx = xp[0][0] * u * 2 + xp[0][1]
            dy = yp[0][0] * u * 2 + yp[0][1]
        elif u >= us[-2]:
            dx = xp[-1][0] * u * 2 + xp[-1][1]
            dy = yp[-1][0] * u * 2 + yp[-1][1]
        else:
            n = self.get_u_index(u)
            mu_r = (u - us[n]) / (us[n + 1] - us[n])
            mu_f = (us[n + 1] - u) / (us[n + 1] - us[n])
            dx1 = xp[n - 1][0] * u * 2 + xp[n - 1][1]
            dy1 = yp[n - 1][0] * u * 2 + yp[n - 1][1]
            dx2 = xp[n][0] * u * 2 + xp[n][1]
            dy2 = yp[n][0] * u * 2 + yp[n][1]
            dx = mu_r * dx2 + mu_f * dx1
            dy = mu_r * dy2 + mu_f * dy1
        return np.array([dx, dy])

    def get_direction_angle(self, u):
        dx, dy = self.calculate_gradient(u)
        return np.arctan2(dy, dx)

    def get_closest_u(self, position):
        from scipy.optimize import fminbound
        return fminbound(lambda u: np.linalg.norm(self(u) - position), full_output=0,
                         x1=0.0 - 10.0, x2=self.length + 10.0, xtol=1e-6, maxfun=500)

    def get_closest_position(self, position):
        return self(self.get_closest_u(position))

    def get_lookahead_point(self, position, lookahead_distance):
        u = self.get_closest_u(position)
        u_la = self.length if u + lookahead_distance > self.length else u + lookahead_distance
        return self(u_la)

    def to_struct(self):
        return B.make_path(self.waypoints, self.us, self.x_params, self.y_params)


# --------------------------------------------------------------------------- #
# reset samplers                                                              #
# --------------------------------------------------------------------------- #
def sample_waypoints(n_wps, distance, scen, W, H, draws):
    """generate_random_waypoints_2d (predef_path.py:307-363)."""
    if scen == "DL":
        x1, y1, lo, hi = draws.uniform(100, 180), draws.uniform(100, 180), 0, np.pi / 2
    elif scen == "UL":
        x1, y1, lo, hi = draws.uniform(100, 180), draws.uniform(H - 180, H - 100), 0, -np.pi / 2
    elif scen == "DR":
        x1, y1, lo, hi = draws.uniform(W - 180, W - 100), draws.uniform(100, 180), np.pi / 2, np.pi
    elif scen == "UR":
        x1, y1, lo, hi = draws.uniform(W - 180, W - 100), draws.uniform(H - 180, H - 100), -np.pi / 2, -np.pi
    else:
        raise ValueError(scen)
    wps = [np.array([x1, y1])]
    for i in range(n_wps - 1):
        az = draws.uniform(lo, hi)
        wps.append(np.array([wps[i][0] + distance * np.cos(az), wps[i][1] + distance * np.sin(az)]))
    return np.array(wps)


def sample_obstacles(n, path, mean, std, on_path, draws):
    """generate_obstacles_around_path (obstacles.py:58-89) -> list of (x, y, r)."""
    out = []
    L = path.length
    while len(out) < n:
        u_obs = draws.uniform(0.20 * L, 0.90 * L)
        ang = path.get_direction_angle(u_obs)
        dist = draws.normal(mean, std)
        on = path(u_obs)
        pos = on + dist * np.array([np.cos(ang - np.pi / 2), np.sin(ang - np.pi / 2)])
        size = draws.uniform(10, 50)
        if np.linalg.norm(pos - on) > size + 10 and not on_path:
            out.append((float(pos[0]), float(pos[1]), float(size)))
        elif on_path:
            out.append((float(on[0]), float(on[1]), float(size)))
    return out


def reset_tables(mode, scenario, cfg_kwargs, draws, stage_spawn_chance=None):
    """env:199-372.  Returns (PortPath|scenario struct, obstacles list, spawn (x, y, angle))."""
    W, H = cfg_kwargs.get("screensize_x", 1300), cfg_kwargs.get("screensize_y", 1300)
    if mode == "test":
        xmin, xmax, ymin, ymax = B.SCENARIO_SPAWN[scenario]
        x1 = draws.py_uniform(xmin, xmax)
        y1 = draws.py_uniform(ymin, ymax)
        ang = draws.py_uniform(-np.pi / 4, np.pi / 4)
        return None, None, (x1, y1, ang)
    corners = cfg_kwargs.get("spawn_corners", (1, 4))
    if cfg_kwargs.get("random_path_spawn", True):
        scen = {1: "DL", 2: "DR", 3: "UL", 4: "UR"}[draws.randint(corners[0], corners[1])]
    else:
        scen = "DR"
    wps = sample_waypoints(cfg_kwargs.get("n_wps", 12), cfg_kwargs.get("path_segment_length", 100), scen, W, H, draws)
    path = PortPath(wps)
    x1, y1 = wps[0][0], wps[0][1]
    ang = draws.py_uniform(-np.pi / 4, np.pi / 4)
    obstacles = []
    if scenario == "stage_1":
        pass
    elif scenario == "stage_2":
        x1 = draws.py_uniform(100, W - 100)
        y1 = draws.py_uniform(100, H - 100)
    elif scenario == "stage_3":
        p = 0.6 if stage_spawn_chance is None else stage_spawn_chance
        if draws.binomial1(p) == 1:
            obstacles = sample_obstacles(1, path, 0, 100, False, draws)
    elif scenario == "stage_4":
        p = 1 if stage_spawn_chance is None else stage_spawn_chance
        if draws.binomial1(p) == 1:
            obstacles = sample_obstacles(1, path, 0, 0, True, draws)
    elif scenario == "stage_5":
        n_obs = draws.normal(1, 4)
        if n_obs < 0 and n_obs > -3:
            n_obs = 1
        if n_obs < -3:
            n_obs = 0
        if n_obs != 0:
            obstacles = sample_obstacles(n_obs, path, 0, 100, False, draws)
            obstacles.append(sample_obstacles(1, path, 0, 0, True, draws)[0])
    else:
        raise ValueError(scenario)
    return path, obstacles, (x1, y1, ang)


# --------------------------------------------------------------------------- #
# single-env port of Drone2dEnv                                               #
# --------------------------------------------------------------------------- #
def _m1to1(v, lo, hi):
    return 2.0 * (v - lo) / (hi - lo) - 1


def _invm1to1(v, lo, hi):
    return (v + 1) * (hi - lo) / 2.0 + lo


def _ssa(a):
    return ((a + np.pi) % (2 * np.pi)) - np.pi


_KW2CFG = ["lookahead", "danger_range", "danger_angle", "abs_inv_CA_min_rew", "PA_band_edge", "PA_scale",
           "PP_vel_scale", "PP_rew_max", "PP_rew_min", "rew_collision", "reach_end_radius", "rew_reach_end",
           "AA_angle", "AA_band", "rew_AA", "n_steps"]


def config_from_kwargs(kw):
    over = {k: kw[k] for k in _KW2CFG if k in kw}
    if "screensize_x" in kw:
        over["screen_x"] = kw["screensize_x"]
    if "screensize_y" in kw:
        over["screen_y"] = kw["screensize_y"]
    if "use_Lambda" in kw:
        over["use_Lambda"] = 1 if kw["use_Lambda"] is True else 0
    return B.default_config(**over)


class RefPortEnv:
    """One environment, reference call structure (see module docstring)."""

    def __init__(self, **kwargs):
        self.kwargs = kwargs
        self.mode = kwargs.get("mode", "curriculum")
        self.scenario = kwargs.get("scenario", "stage_1")
        self.cfg = config_from_kwargs(kwargs)
        self.lookahead = self.cfg.lookahead
        self.W, self.H = self.cfg.screen_x, self.cfg.screen_y
        self._lib = B.lib()
        self.draws = NumpyDraws()
        self._init_world()

    # env:182-392
    def _init_world(self):
        path, obstacles, (x, y, ang) = reset_tables(self.mode, self.scenario, self.kwargs, self.draws)
        if self.mode == "test":
            rec_path, rec_obs = B.load_scenario(self.scenario)
            wps = [(rec_path.wx[i], rec_path.wy[i]) for i in range(rec_path.n_wps)]
            path = PortPath(wps)
            obstacles = [(rec_obs.x[i], rec_obs.y[i], rec_obs.r[i]) for i in range(rec_obs.n)]
        self.predef_path = path
        self.wps = path.waypoints
        self.obstacles = obstacles
        self.ob_struct = B.make_obstacles(obstacles)
        self.state = B.spawn(x, y, ang, self.cfg)
        self.k_obs = min(3, len(obstacles))
        self.LA_in_last_wp = False
        self.done = False
        self.current_time_step = 0
        self.path_error = 0
        self.total_reward = 0
        self.info = {}

    def reset(self):
        self._init_world()
        return self.get_observation()

    # env:617-629, 948-961
    def get_obstacle_distances(self, k):
        f = self.state.body[0]
        d = {}
        for i, (ox, oy, r) in enumerate(self.obstacles):
            best = float("inf")
            for vx, vy in ((50.0, -5.0), (50.0, 5.0), (-50.0, 5.0), (-50.0, -5.0)):
                dist = math.sqrt((vx + f.px - ox) ** 2 + (vy + f.py - oy) ** 2) - r
                if dist < best:
                    best = dist
            d[i] = best
        s = {kk: v for kk, v in sorted(d.items(), key=lambda item: item[1])}
        return dict(list(s.items())[0:k])

    # env:631-773
    def get_observation(self):
        f = self.state.body[0]
        W, H = self.W, self.H
        x, y, alpha = f.px, f.py, f.a
        xt, yt = self.wps[-1][0], self.wps[-1][1]
        o = [_m1to1(f.vx, -1330, 1330), _m1to1(f.vy, -1330, 1330), np.clip(f.w / 11.7, -1, 1), alpha / np.pi,
             _m1to1(xt - x, 0, W), _m1to1(yt - y, 0, H), _m1to1(x, 0, W), _m1to1(y, 0, H)]
        slots = [1, 0, 0, 1, 0, 0, 1, 0, 0]
        if self.obstacles != []:
            near = self.get_obstacle_distances(self.k_obs)
            for j, (idx, dist) in enumerate(near.items()):
                ang = np.arctan2(y - self.obstacles[idx][1], x - self.obstacles[idx][0])
                ang = _ssa(ang - alpha - np.pi)
                slots[3 * j:3 * j + 3] = [_m1to1(dist, 0, np.sqrt(W ** 2 + H ** 2)), np.sin(ang), np.cos(ang)]
        o += slots
        va = _ssa(np.arctan2(f.vy, f.vx) - alpha)
        o += [np.sin(va), np.cos(va)]
        closest_point = self.predef_path.get_closest_position([x, y])
        o += [_m1to1(closest_point[0], 0, W), _m1to1(closest_point[1], 0, H)]
        la = self.predef_path.get_lookahead_point([x, y], self.lookahead)
        if np.abs(la[0] - xt) < 10 and np.abs(la[1] - yt) < 10:
            self.LA_in_last_wp = True
        if self.LA_in_last_wp:
            la = self.wps[-1]
        o += [_m1to1(la[0], 0, W), _m1to1(la[1], 0, H)]
        R = np.array([[np.cos(alpha), -np.sin(alpha)], [np.sin(alpha), np.cos(alpha)]])
        for pt in (la, closest_point):
            b = np.matmul(R, pt - np.array([x, y]))
            ang = _ssa(np.arctan2(b[1], b[0]) - alpha)
            o += [np.sin(ang), np.cos(ang)]
        return np.array(o)

    # env:394-615
    def step(self, action):
        c = self.cfg
        self._lib.d2o_space_step(C.byref(self.state), C.byref(self.ob_struct), C.byref(c),
                                 float(action[0]), float(action[1]))
        self.current_time_step += 1
        obs = self.get_observation()
        W, H = self.W, self.H
        vx, vy = _invm1to1(obs[0], -1330, 1330), _invm1to1(obs[1], -1330, 1330)
        alpha = obs[3] * np.pi
        tdx, tdy = _invm1to1(obs[4], 0, W), _invm1to1(obs[5], 0, H)
        px, py = _invm1to1(obs[6], 0, W), _invm1to1(obs[7], 0, H)
        vel_angle = (np.arctan2(obs[17] * np.pi, obs[18] * np.pi) + 2 * np.pi) % (2 * np.pi)
        cpx, cpy = _invm1to1(obs[19], 0, W), _invm1to1(obs[20], 0, H)
        la_angle = (np.arctan2(obs[23], obs[24]) + 2 * np.pi) % (2 * np.pi)
        lam_pa = lam_ca = 1
        r_ca = 0
        d1 = np.inf
        if self.obstacles != []:
            d1 = _invm1to1(obs[8], 0, np.sqrt(W ** 2 + H ** 2))
            oa = (np.arctan2(obs[9], obs[10]) + 2 * np.pi) % (2 * np.pi)
            ad = abs(np.rad2deg((oa - vel_angle + np.pi) % (2 * np.pi) - np.pi))
            DR, DA, k = c.danger_range, c.danger_angle, c.abs_inv_CA_min_rew
            if d1 < DR and c.use_Lambda:
                lam_pa = max((d1 / DR) / 2, 0.10)
                lam_ca = 1 - lam_pa
            if d1 < DR:
                rr = min(-(((DR + k * DR) / (d1 + k * DR)) - 1), 0)
                ar = min(-(((DA + k * DA) / (ad + k * DA)) - 1), 0)
                r_ca = rr + ar
        dist = np.linalg.norm(np.array([cpx, cpy]) - np.array([px, py]))
        r_pa = -(2 * (np.clip(dist, 0, c.PA_band_edge) / c.PA_band_edge) - 1) * c.PA_scale
        vel = np.sqrt(vx ** 2 + vy ** 2)
        diff = abs((la_angle - vel_angle + np.pi) % (2 * np.pi) - np.pi)
        r_pp = np.clip(np.cos(diff) * vel * c.PP_vel_scale, c.PP_rew_min, c.PP_rew_max)
        r_coll, e1 = (c.rew_collision, True) if self.state.collided else (0, False)
        r_end, e2 = (c.rew_reach_end, True) if (abs(tdx) < c.reach_end_radius and abs(tdy) < c.reach_end_radius) else (0, False)
        r_aa, e5 = 0, False
        if alpha > c.AA_band:
            r_aa = -np.sin(alpha)
        if alpha < -c.AA_band:
            r_aa = np.sin(alpha)
        if abs(alpha) >= c.AA_angle:
            r_aa, e5 = c.rew_AA, True
        e4 = self.current_time_step == c.n_steps
        reward = r_aa + r_pa * lam_pa + r_pp + r_coll + r_ca * lam_ca + r_end
        self.path_error += dist
        self.total_reward += reward
        self.info = {"reward": reward, "collision_avoidance_reward": r_ca * lam_ca, "path_adherence": r_pa * lam_pa,
                     "path_progression": r_pp, "collision_reward": r_coll, "reach_end_reward": r_end,
                     "agressive_alpha_reward": r_aa, "env_steps": self.current_time_step, "dist_closest_obs": d1}
        if e1 or e2 or e4 or e5:
            self.done = True
            # env:595-610: the later branches overwrite n_collisions with 0
            self.info.update(n_collisions=int(e1 and not (e2 or e4 or e5)), n_successful_runs=int(e2),
                             n_failed_runs=int(e1 or e4 or e5),
                             APE=self.path_error / self.current_time_step, total_reward=self.total_reward)
        return obs, reward, self.done, self.info
