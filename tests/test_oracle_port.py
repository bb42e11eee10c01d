"""Cross-checks inside the oracle (no GPU): the Python port of the reference env
(scipy fminbound, numpy, Python obstacle loop) against the C restatement, the
reset samplers against the reference's recorded draws, and physical sanity of
the restated Chipmunk step (which no golden vector can pin, see oracle/__init__)."""
import ctypes as C
import json
import math
import os
import random

import numpy as np
import pytest

import oracle
from oracle import ref_port
from conftest import GOLDEN


def _load(name):
    with open(os.path.join(GOLDEN, name)) as f:
        return json.load(f)


def test_waypoint_sampler_matches_reference_draws():
    for rec in _load("waypoint_gen.json"):
        wps = ref_port.sample_waypoints(rec["n_wps"], rec["distance"], rec["scen"], 1300, 1300,
                                        ref_port.RecordedDraws(rec["draws"]))
        assert np.array_equal(wps, np.array(rec["wps"]))


def test_obstacle_sampler_matches_reference_draws():
    for rec in _load("obstacle_gen.json"):
        path = ref_port.PortPath(rec["path"]["wps"])
        off = ref_port.sample_obstacles(3, path, 0, 100, False, ref_port.RecordedDraws(rec["off_draws"]))
        on = ref_port.sample_obstacles(1, path, 0, 0, True, ref_port.RecordedDraws(rec["on_draws"]))
        np.testing.assert_allclose(np.array(off), np.array(rec["off"]), rtol=0, atol=1e-12)
        np.testing.assert_allclose(np.array(on), np.array(rec["on"]), rtol=0, atol=1e-12)


def test_moments_and_topology():
    lib = oracle.lib()
    assert abs(lib.d2o_moment_for_box(0.2, 100.0, 10.0) - 0.2 * (100 ** 2 + 10 ** 2) / 12) < 1e-12
    assert abs(lib.d2o_moment_for_box(0.4, 20.0, 20.0) - 0.4 * (20 ** 2 + 20 ** 2) / 12) < 1e-12
    cfg = oracle.default_config()
    s = oracle.spawn(300.0, 400.0, 0.3, cfg)
    # motors sit R=40 along the body x axis (Drone.py:11,37,51)
    assert abs(math.hypot(s.body[2].px - 300, s.body[2].py - 400) - 40) < 1e-12
    assert abs(s.body[1].px + s.body[2].px - 600) < 1e-12


def _free_step(s, cfg, a_l, a_r, n=1):
    ob = oracle.make_obstacles([])
    for _ in range(n):
        oracle.lib().d2o_space_step(C.byref(s), C.byref(ob), C.byref(cfg), a_l, a_r)


def test_physics_centre_of_mass_and_rigidity():
    """Internal pivot impulses cancel: the centre of mass of the 3 bodies obeys
    m*a = m*g + thrust exactly (semi-implicit Euler), whatever the solver does."""
    cfg = oracle.default_config()
    s = oracle.spawn(500.0, 600.0, 0.2, cfg)
    m = np.array([0.2, 0.4, 0.4])

    def com_v():
        return np.array([sum(m[i] * s.body[i].vx for i in range(3)), sum(m[i] * s.body[i].vy for i in range(3))]) / m.sum()

    rng = np.random.RandomState(0)
    for _ in range(50):
        a = rng.uniform(-1, 1, 2)
        ang = s.body[0].a
        v0 = com_v()
        _free_step(s, cfg, a[0], a[1])
        F = ((a[0] / 2 + 0.5) + (a[1] / 2 + 0.5)) * 1000
        dv = (np.array([0, -1000.0]) + F * np.array([-math.sin(ang), math.cos(ang)]) / m.sum()) / 60
        np.testing.assert_allclose(com_v() - v0, dv, rtol=0, atol=1e-9)
        # the 3 pivots per motor keep it glued to the frame: anchor error stays tiny
        for b, sign in ((1, -1), (2, 1)):
            ex = s.body[0].px + sign * 40 * math.cos(s.body[0].a) - s.body[b].px
            ey = s.body[0].py + sign * 40 * math.sin(s.body[0].a) - s.body[b].py
            assert math.hypot(ex, ey) < 0.5
            assert abs(s.body[b].a - s.body[0].a) < 0.05


def test_physics_hover_and_torque_sign():
    cfg = oracle.default_config()
    s = oracle.spawn(500.0, 600.0, 0.0, cfg)
    _free_step(s, cfg, 0.0, 0.0, n=30)          # 500 N + 500 N = weight of 1 kg at g = 1000
    assert abs(s.body[0].py - 600.0) < 1e-6 and abs(s.body[0].vy) < 1e-6
    s = oracle.spawn(500.0, 600.0, 0.0, cfg)
    _free_step(s, cfg, -1.0, 1.0, n=5)          # right motor only -> counter-clockwise
    assert s.body[0].w > 0 and s.body[0].a > 0
    s = oracle.spawn(500.0, 600.0, 0.0, cfg)
    _free_step(s, cfg, -1.0, -1.0, n=60)        # free fall for 1 s (positions move first): y = y0 - g*t*(t-dt)/2
    assert abs(s.body[0].vy + 1000.0) < 1e-6
    assert abs(s.body[0].py - (600.0 - 1000.0 * (1.0 - 1 / 60) / 2)) < 1e-6


def test_collision_flag_is_rotated_box_vs_circle():
    cfg = oracle.default_config()
    lib = oracle.lib()
    # frame box is 100 x 10 (Drone.py:16): tip at +-50 along the body axis
    for ang, cx, cy, r, want in [(0.0, 560.0, 600.0, 10.0, 1), (0.0, 560.0, 600.0, 9.9, 0),
                                 (0.0, 500.0, 615.0, 10.0, 1), (0.0, 500.0, 615.1, 10.0, 0),
                                 (math.pi / 2, 500.0, 660.0, 10.0, 1), (math.pi / 2, 560.0, 600.0, 10.0, 0),
                                 (0.0, 500.0, 600.0, 1.0, 1)]:
        s = oracle.spawn(500.0, 600.0, ang, cfg)
        ob = oracle.make_obstacles([(cx, cy, r)])
        hit = lib.d2o_space_step(C.byref(s), C.byref(ob), C.byref(cfg), 0.0, 0.0)
        assert hit == want and s.collided == want, (ang, cx, cy, r)


@pytest.mark.parametrize("scenario", ["corridor", "S_corridor", "large"])
def test_python_port_agrees_with_c_restatement(scenario):
    """Two independent restatements of env:394-773 (Python + scipy.fminbound vs.
    plain C with its own Brent) fed the same states and actions."""
    random.seed(5)
    np.random.seed(5)
    kw = dict(mode="test", scenario=scenario)
    env = ref_port.RefPortEnv(**kw)
    path, obst = oracle.load_scenario(scenario)
    cfg = env.cfg
    st = oracle.State.from_buffer_copy(env.state)
    rng = np.random.RandomState(1)
    for t in range(40):
        a = rng.uniform(-1, 1, 2).astype(np.float32)
        obs, rew, done, info = env.step(a)
        out = oracle.env_step(st, path, obst, cfg, a[0], a[1])
        o = np.array(out.obs[:])
        # u* from two Brent implementations: identical evaluation sequence (see test_oracle_golden)
        np.testing.assert_allclose(o, obs, rtol=0, atol=5e-9)
        assert abs(out.reward - rew) < 1e-7
        assert bool(out.done) == bool(done)
        if done:
            break


def test_stage5_obstacle_count_distribution():
    """env:362-372: P(no obstacle) = P(N(1,4) < -3) = 15.9 %, mean count 3.5."""
    np.random.seed(0)
    random.seed(0)
    counts = []
    for _ in range(400):
        _, obstacles, _ = ref_port.reset_tables("curriculum", "stage_5", {}, ref_port.NumpyDraws())
        counts.append(len(obstacles))
    counts = np.array(counts)
    assert 0.10 < np.mean(counts == 0) < 0.22
    assert not np.any(counts == 1)
    assert 3.0 < counts.mean() < 4.1
