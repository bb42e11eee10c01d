"""Step-level pin of the restated rigid-body step (cpSpaceStep for this world, oracle/d2d_oracle.c:d2o_space_step)
on GENUINE pymunk output: the frame positions the reference's own evaluation runs recorded after every space.step
(Tests/<scenario>/test_4/flight_paths; fixtures cut by tests/golden/make_reference_trajectories.py).

oracle/inverse_dynamics.py recovers, from the recorded positions alone, the motor commands of every step.  If the
restated step is the step pymunk took, commands exist INSIDE the action Box that reproduce every recorded velocity
to rounding noise, and a large share of them sits exactly on the edge of the Box (SB3 clips the Gaussian policy's
samples, main.py:262).  Restatements that differ in the solver's iteration count (9 or 11 instead of pymunk's 10),
warm starting, the order of the joints, the order of position and velocity integration, damping, a mass or gravity
leave residuals four to six orders of magnitude larger.  tests/golden/make_inverse_dynamics.py ran this over all
700 genuine episodes (inverse_dynamics_stats.json) and stored the recovered actions of the 70 committed ones.
"""
import ctypes as C
import json
import os
import subprocess

import numpy as np
import pytest

from oracle import binding as B
from oracle import inverse_dynamics as ID

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
SCENARIOS = ["perpendicular", "parallel", "S_parallel", "corridor", "S_corridor", "impossible", "large"]
HEAD = 90           # steps inverted per episode inside the test (the generator did 150 / whole episodes)


def episodes():
    z = np.load(os.path.join(GOLDEN, "reference_trajectories.npz"))
    out = {}
    for scen in SCENARIOS:
        off = 0
        for k, n in enumerate(z[scen + "__len"]):
            out[(scen, k)] = z[scen + "__xy"][off:off + n]
            off += n
    return out


@pytest.fixture(scope="module")
def genuine():
    return episodes()


@pytest.fixture(scope="module")
def recovered():
    z = np.load(os.path.join(GOLDEN, "inverse_dynamics_actions.npz"))
    return {(s, k): (float(z["%s__%d__alpha0" % (s, k)]), z["%s__%d__actions" % (s, k)]) for s in SCENARIOS for k in range(10)}


def variant_lib(tmp_path, flag):
    so = str(tmp_path / ("liboracle_%s.so" % "".join(ch for ch in flag if ch.isalnum())))
    subprocess.check_call(["gcc", "-O2", "-fPIC", "-std=c99", "-ffp-contract=off", "-shared", flag, "-o", so,
                           os.path.join(os.path.dirname(B.LIB_PATH), "d2d_oracle.c"), "-lm"])
    return B.load(so)


def test_restated_step_reproduces_genuine_velocities(genuine):
    """Live inversion of the first steps of three genuine episodes: every recorded frame velocity is reproduced to
    ~1e-7 px/s (velocities are of order 1e2..1e3) by commands inside the Box, many of them exactly +-1."""
    for key in (("corridor", 0), ("large", 3), ("S_parallel", 7)):
        rec = ID.Episode(genuine[key]).recover_actions(max_steps=HEAD)
        s = ID.summarize(rec, tail=24)
        assert s["residual_median"] < 1e-6, (key, s)
        assert s["residual_max"] < 1e-3, (key, s)
        assert s["frac_raw_outside_box"] == 0.0
        assert s["frac_actions_at_box_edge"] > 0.05, (key, s)     # 10 .. 50 % of the commands are clipped ones
        assert rec["residual"][0] < 1e-5                # first step: known spawn state, angle from the first thrust


def test_variants_of_the_step_are_rejected(genuine, tmp_path):
    """The same inversion under deliberately different restatements: the residual floor is >= 1000x higher."""
    xy = genuine[("corridor", 0)]
    base = ID.summarize(ID.Episode(xy).recover_actions(max_steps=HEAD), tail=24)["residual_median"]
    table = {}
    for name, over, flag in (("9 iterations", {"iterations": 9}, None), ("11 iterations", {"iterations": 11}, None),
                             ("no warm start", {}, "-DD2O_NO_WARM_START"),
                             ("joint order permuted", {}, "-DD2O_JOINT_ORDER_PERMUTED"),
                             ("positions after velocities", {}, "-DD2O_POSITIONS_AFTER")):
        lib = variant_lib(tmp_path, flag) if flag else None
        s = ID.summarize(ID.Episode(xy, cfg=B.default_config(**over), lib=lib).recover_actions(max_steps=HEAD), tail=24)
        table[name] = s["residual_median"]
        assert s["residual_median"] > 1e-3 and s["residual_median"] > 1000 * base, (name, s, base)
    print("residual medians:", base, table)


def test_recovered_actions_fly_the_genuine_trajectories(genuine, recovered):
    """All 70 committed episodes, whole length: the stored commands, applied open loop from the spawn pose by the
    oracle's step, reproduce the recorded positions; they lie in the Box and a large share is exactly +-1."""
    lib, cfg, none = B.lib(), B.default_config(), B.make_obstacles([])
    worst_pos, edge = 0.0, []
    for key, xy in genuine.items():
        alpha0, acts = recovered[key]
        assert acts.shape == (len(xy) - 1, 2) and np.abs(acts).max() <= 1.0
        s = B.spawn(xy[0, 0], xy[0, 1], alpha0, cfg)
        err = np.zeros(len(acts))
        for k in range(len(acts)):
            lib.d2o_space_step(C.byref(s), C.byref(none), C.byref(cfg), acts[k, 0], acts[k, 1])
            if k + 1 < len(acts):
                # position after step k+1 is xy[k+1] = xy[k] + v_k * dt: compare the velocity now, the position next
                err[k] = max(abs(s.body[0].px - xy[k, 0]), abs(s.body[0].py - xy[k, 1]))
        head = err[:max(len(acts) - 24, 1)]
        worst_pos = max(worst_pos, head.max())
        edge.append((np.abs(acts[:-24]) == 1.0).mean())
    assert worst_pos < 1e-3, worst_pos                   # px, over up to 1100 open-loop steps
    assert np.median(edge) > 0.15, np.median(edge)


def test_committed_statistics_over_all_700_episodes():
    st = json.load(open(os.path.join(GOLDEN, "inverse_dynamics_stats.json")))["variants"]
    r = st["restated"]
    assert r["episodes"] == 700 and r["episodes_with_undetermined_spawn_angle"] == 0
    worst = r["episode_residual_median"]["max"]
    assert worst < 1e-5, r                                    # the WORST of the 700 episodes' median residuals (1.6e-6)
    assert r["episode_residual_median"]["median"] < 1e-7, r   # the typical episode (1.4e-8)
    assert st["restated_full"]["episodes"] == 70 and st["restated_full"]["episode_residual_median"]["max"] < 1e-5
    for name, v in st.items():
        if name.startswith("restated"):
            continue
        # the BEST episode of every variant is at least 10x worse than the WORST episode of the restated step, the typical
        # one at least 100x (the mildest variant is 0.1 % more gravity: 19x / 360x; 9 or 11 iterations: 260x / 4700x)
        assert v["episode_residual_median"]["min"] > 10 * worst, (name, v, r)
        assert v["episode_residual_median"]["median"] > 100 * worst, (name, v, r)
