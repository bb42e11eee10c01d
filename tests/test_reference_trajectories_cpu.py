"""The oracle against GENUINE reference outputs: the trajectories, rewards, APEs and episode lengths the
reference's own evaluation runs saved under Tests/<scenario>/test_4/ (pymunk/Chipmunk2D + scipy + SB3;
fixtures cut by tests/golden/make_reference_trajectories.py).

What each test pins, and how hard:
  * APE                 path tables, closest-u (bounded Brent), the accumulator    -- per episode, 1e-9
  * termination         reach-end box, time limit, which flag fired, on which step -- exact
  * total reward        all six reward terms (env:457-572)                         -- per episode, ~0.3 %
  * collision           the frame-box-vs-circle begin condition (env:17-19)        -- statistical (the
                        frame angle is not recorded and has to be estimated from the thrust direction)
  * rigid-body step     g, dt, masses, force_scale, joint solver iterations, (absence of) damping, through
                        the distribution of the implied thrust acceleration        -- statistical
"""
import json
import os
import subprocess

import numpy as np
import pytest

from oracle import binding as B
from oracle import closed_loop
from oracle import flight_replay as FR

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
SCENARIOS = ["perpendicular", "parallel", "S_parallel", "corridor", "S_corridor", "impossible", "large"]


@pytest.fixture(scope="module")
def genuine():
    z = np.load(os.path.join(GOLDEN, "reference_trajectories.npz"))
    stats = json.load(open(os.path.join(GOLDEN, "reference_trajectory_stats.json")))
    eps = []
    for scen in SCENARIOS:
        off = 0
        for k, n in enumerate(z[scen + "__len"]):
            eps.append(dict(scenario=scen, xy=z[scen + "__xy"][off:off + n], total_reward=float(z[scen + "__total_reward"][k]),
                            ape=float(z[scen + "__ape"][k]), collision=int(z[scen + "__collision"][k])))
            off += n
    return eps, stats


@pytest.fixture(scope="module")
def replayed(genuine):
    eps, stats = genuine
    cfg = B.default_config(**stats["run_config"])
    return [FR.replay_rewards(e["scenario"], e["xy"], e["collision"], cfg=cfg, angles=FR.estimate_angle(e["xy"])) for e in eps]


def test_fixture_shape(genuine):
    eps, stats = genuine
    assert len(eps) == 70 and sum(e["collision"] for e in eps) >= 25
    assert set(stats["scenarios"]) == set(SCENARIOS)


def test_ape_matches_the_reference_per_episode(genuine, replayed):
    """info['APE'] (env:589-590) of every genuine episode, recomputed by the oracle from the genuine positions:
    path tables + fminbound restatement + closest point + running mean."""
    eps, _ = genuine
    for e, r in zip(eps, replayed):
        assert abs(r["ape"] - e["ape"]) < 1e-9, (e["scenario"], r["ape"], e["ape"])


def test_termination_step_and_reason(genuine, replayed):
    """The oracle's done flag fires on exactly the step the reference ended the episode on, not earlier, and for
    a reason consistent with the reference's n_collisions record."""
    eps, _ = genuine
    for e, r in zip(eps, replayed):
        assert r["done_at"] == len(e["xy"]), (e["scenario"], r["done_at"], len(e["xy"]))
        coll, reached, timeout, alpha = r["end"]
        assert bool(coll) == bool(e["collision"])
        if not e["collision"]:
            assert reached or timeout or alpha
            assert bool(timeout) == (len(e["xy"]) == 1100) or reached


def test_total_reward_matches_the_reference_per_episode(genuine, replayed):
    """info['total_reward'] (env:592).  Velocity is exact ((p[t+1]-p[t])/dt is what Chipmunk integrated);
    the only approximations are the last step's velocity and the angle entering r_AA."""
    eps, _ = genuine
    err = np.array([r["total_reward"] - e["total_reward"] for e, r in zip(eps, replayed)])
    assert np.abs(err).max() < 6.0, err
    assert np.median(np.abs(err)) < 0.5, np.median(np.abs(err))
    scale = np.mean([abs(e["total_reward"]) for e in eps])
    assert np.median(np.abs(err)) / scale < 1e-3


def test_shipped_reward_settings_are_not_the_ones_of_those_runs(genuine):
    """Documents the identification in make_reference_trajectories.py: with rl_config.py's shipped values the
    same replay is off by hundreds, i.e. the per-episode agreement above is not a loose test."""
    eps, _ = genuine
    cfg = B.default_config()
    err = [FR.replay_rewards(e["scenario"], e["xy"], e["collision"], cfg=cfg, angles=FR.estimate_angle(e["xy"]))["total_reward"]
           - e["total_reward"] for e in eps[:10]]
    assert np.median(np.abs(err)) > 50


def _touch_counts(eps, cfg):
    last = before = steps = 0
    lib = B.lib()
    import ctypes as C
    for e in eps:
        _, obst = B.load_scenario(e["scenario"])
        ang = FR.estimate_angle(e["xy"])
        f = B.Body()
        for t in range(len(e["xy"])):
            f.px, f.py, f.a = e["xy"][t, 0], e["xy"][t, 1], ang[t]
            hit = lib.d2o_frame_touches(C.byref(f), C.byref(obst), C.byref(cfg))
            if t == len(e["xy"]) - 1:
                last += int(bool(hit) and bool(e["collision"]))
            else:
                before += int(hit)
                steps += 1
    return last, before, steps


def test_collision_condition_on_genuine_poses(genuine):
    """Frame box (width x height/2, Drone.py:16) against the obstacle circles at the genuine positions with the
    estimated angle: touches on the last step of (nearly) every collision episode and (nearly) never earlier.
    A box of the full nominal height, or one that includes the motors, contradicts the genuine episodes."""
    eps, _ = genuine
    n_coll = sum(e["collision"] for e in eps)
    last, before, steps = _touch_counts(eps, B.default_config())
    assert last >= 0.85 * n_coll, (last, n_coll)
    assert before <= 3e-4 * steps, (before, steps)
    last2, before2, _ = _touch_counts(eps, B.default_config(height=80.0))     # 4x thicker box
    assert before2 > 10 * max(before, 1), (before, before2)
    last3, before3, _ = _touch_counts(eps, B.default_config(width=50.0))      # half as wide
    assert last3 < last - 3, (last, last3)


@pytest.fixture(scope="module")
def corridor_flights():
    """Closed loop of the shipped agent on the oracle for a few solver variants (corridor, 100 episodes)."""
    out = {}
    for name, kw in (("iter10", {}), ("iter5", {"iterations": 5}), ("iter20", {"iterations": 20})):
        res = closed_loop.evaluate("test", "corridor", 100, seed=3, cfg_kwargs=None, cfg_overrides=kw, record_paths=True)
        out[name] = res["flight_paths"]
    return out


def test_implied_thrust_distribution_matches_genuine_pymunk(genuine, corridor_flights):
    """Second differences of the frame position: the genuine pymunk runs and the restated cpSpaceStep give the
    same distribution; 5 or 20 solver iterations instead of pymunk's default 10 do not."""
    _, stats = genuine
    ref = stats["scenarios"]["corridor"]["implied_thrust"]
    pct = ref["percentiles"]
    i50, i75, i95, i99 = (pct.index(p) for p in (50, 75, 95, 99))

    def check(s):
        a, b = np.array(s["thrust_acc"]), np.array(ref["thrust_acc"])
        return (abs(a[i50] - b[i50]) < 25 and abs(a[i75] - b[i75]) < 25 and abs(a[i95] - b[i95]) < 12 and
                abs(a[i99] - b[i99]) < 8 and abs(s["max"] - ref["max"]) < 40 and
                abs(s["frac_above_2000"] - ref["frac_above_2000"]) < 0.004 and abs(s["frac_below_5"] - ref["frac_below_5"]) < 0.003)

    s10 = FR.implied_thrust_stats(corridor_flights["iter10"])
    assert check(s10), (s10, ref)
    for other in ("iter5", "iter20"):
        s = FR.implied_thrust_stats(corridor_flights[other])
        assert not check(s), (other, s)
    # thrust can never exceed 2*force_scale/total mass = 2000 by more than the joint transients
    assert ref["thrust_acc"][i99] < 2050 and s10["thrust_acc"][i99] < 2050


def test_no_damping_in_the_genuine_runs(genuine, corridor_flights, tmp_path):
    """env:376-380 sets `body.damping = 0.9`, a plain Python attribute Chipmunk never reads.  In near-free-fall
    samples of the genuine trajectories the horizontal acceleration does not depend on the horizontal velocity (Space.damping = 0.9
    would give a slope of -0.105 /s); the oracle agrees, and an oracle rebuilt WITH damping does not."""
    _, stats = genuine
    for scen in SCENARIOS:
        it = stats["scenarios"][scen]["implied_thrust"]
        if it["n_low_thrust"] >= 500:
            assert abs(it["horizontal_drag_coefficient"]) < 0.01, (scen, it)
    s10 = FR.implied_thrust_stats(corridor_flights["iter10"])
    assert abs(s10["horizontal_drag_coefficient"]) < 0.01
    so = str(tmp_path / "liboracle_damped.so")
    here = os.path.dirname(B.LIB_PATH)
    subprocess.check_call(["gcc", "-O2", "-fPIC", "-std=c99", "-ffp-contract=off", "-shared", "-DD2O_SPACE_DAMPING=0.9",
                           "-o", so, os.path.join(here, "d2d_oracle.c"), "-lm"])
    res = closed_loop.evaluate("test", "corridor", 100, seed=3, record_paths=True, lib_path=so)
    sd = FR.implied_thrust_stats(res["flight_paths"])
    assert abs(sd["horizontal_drag_coefficient"] - (1 - 0.9 ** (1 / 60)) * 60) < 0.01, sd
