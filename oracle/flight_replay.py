"""Replay of GENUINE reference trajectories through the oracle (TEST INFRASTRUCTURE ONLY).

The reference's evaluation runs (main.py:242-327) saved, per episode, the frame position after
every space.step (env:410-415,984-986: `(x, screen_height - y)`), the episode's total reward, its
length, its APE and whether it ended in a collision.  Those files are outputs of the real
pymunk/Chipmunk2D + scipy stack, which cannot be imported here, so they are the only genuine
samples of the rigid-body step this repository can be checked against.  Two uses:

* `implied_thrust_stats`: Chipmunk moves a body by `p += v*dt` with the velocity left by the
  previous step's solver (cpBodyUpdatePosition), so second differences of the frame position
  give the frame's acceleration; minus gravity this is what thrust and the six pivot joints did
  to it.  Its distribution depends on g, dt, the masses, force_scale, the joint solver's
  iteration count and on whether any damping is applied, and so discriminates restatements.
* `replay_rewards`: positions and first differences (= the observed velocity, exactly) of a
  genuine episode are pushed through the oracle's after-physics half of step() (observation,
  closest-u, nearest obstacles, reward terms, termination) and the summed reward / APE are
  compared with the numbers the reference itself wrote for that episode.
"""
import ctypes as C

import numpy as np

from . import binding as B


def to_world(flight_path, screen_height=1300.0):
    a = np.asarray(flight_path, dtype=np.float64)
    return np.column_stack([a[:, 0], screen_height - a[:, 1]])


def implied_thrust(xy, dt=1.0 / 60, g=(0.0, -1000.0)):
    """(T-2, 2) acceleration of the frame body not explained by gravity, and the velocity before it."""
    xy = np.asarray(xy, dtype=np.float64)
    acc = (xy[2:] - 2 * xy[1:-1] + xy[:-2]) / dt ** 2
    vel = (xy[1:-1] - xy[:-2]) / dt
    return acc - np.asarray(g), vel


PCTS = (1, 5, 25, 50, 75, 95, 99, 99.9)


def implied_thrust_stats(paths, dt=1.0 / 60, g=(0.0, -1000.0)):
    thr, vel = [], []
    for xy in paths:
        if len(xy) >= 5:
            t, v = implied_thrust(xy, dt, g)
            thr.append(t)
            vel.append(v)
    thr, vel = np.concatenate(thr), np.concatenate(vel)
    mag = np.hypot(thr[:, 0], thr[:, 1])
    low = mag < 40.0                      # both propellers (almost) off: near free fall
    # Space.damping = d multiplies every velocity by d**dt per step, i.e. an acceleration of
    # -(1 - d**dt)/dt * v.  Thrust points (nearly) up, so the horizontal component of what is left in
    # near-free-fall samples isolates it: least-squares slope of thr_x on v_x, sign flipped.
    vx, tx = vel[low, 0], thr[low, 0]
    drag = -float(np.sum(vx * tx) / max(np.sum(vx * vx), 1e-30))
    return {
        "n_samples": int(len(mag)),
        "percentiles": [float(p) for p in PCTS],
        "thrust_acc": [float(v) for v in np.percentile(mag, PCTS)],
        "max": float(mag.max()),
        "frac_above_2000": float((mag > 2000).mean()),
        "frac_below_5": float((mag < 5).mean()),
        "n_low_thrust": int(low.sum()),
        # [1/s]; Space.damping = d would give (1 - d**dt)/dt = 0.105 for d = 0.9
        "horizontal_drag_coefficient": drag,
    }


def estimate_angle(xy, dt=1.0 / 60, g=(0.0, -1000.0), min_thrust=150.0):
    """Frame angle at each recorded pose, from the direction of the implied thrust.

    With P_s the pose after step s (= xy[s-1]) and V_s = (P_{s+1}-P_s)/dt the velocity the solver
    left after step s, step s+1 applies the propeller forces along the frame's local +y AT THE
    ANGLE OF P_s (env:403-404) and then integrates, so (V_{s+1}-V_s)/dt - g, which is
    implied_thrust(xy)[s-1], points along (-sin A_s, cos A_s).  The three bodies move almost
    rigidly, the frame centre being their centre of mass.  Where the thrust is too weak to give a
    direction, and for the last two poses (no later samples), the angle is extrapolated linearly."""
    T = len(xy)
    ang = np.full(T, np.nan)
    if T >= 3:
        thr, _ = implied_thrust(xy, dt, g)
        ok = np.hypot(thr[:, 0], thr[:, 1]) > min_thrust
        ang[:T - 2][ok] = np.arctan2(-thr[ok, 0], thr[ok, 1])
    known = np.flatnonzero(~np.isnan(ang))
    if len(known) == 0:
        return np.zeros(T)
    if len(known) == 1:
        return np.full(T, ang[known[0]])
    out = np.interp(np.arange(T), known, ang[known])
    k1, k2 = known[-2], known[-1]
    slope = (ang[k2] - ang[k1]) / (k2 - k1)
    tail = np.arange(k2 + 1, T)
    out[tail] = ang[k2] + slope * (tail - k2)
    return out


def replay_rewards(scenario, xy, ended_in_collision, cfg=None, angles=None):
    """Sum of the oracle's reward over one genuine episode.  Velocity after step t is
    (xy[t+1]-xy[t])/dt exactly (cpBodyUpdatePosition); for the last step it is extrapolated.
    The collision flag comes from the oracle's own frame-vs-circle test when `angles` are given
    (else from the reference's record on the last step)."""
    cfg = cfg or B.default_config()
    path, obst = B.load_scenario(scenario)
    lib = B.lib()
    xy = np.asarray(xy, dtype=np.float64)
    T = len(xy)
    dt = cfg.dt
    vel = np.empty_like(xy)
    vel[:-1] = (xy[1:] - xy[:-1]) / dt
    vel[-1] = 2 * vel[-2] - vel[-3] if T >= 3 else 0.0
    ang = np.zeros(T) if angles is None else np.asarray(angles, dtype=np.float64)
    s = B.State()
    out = B.StepOut()
    total = 0.0
    terms = np.zeros(6)
    touch = np.zeros(T, bool)
    done_at = None
    for t in range(T):
        f = s.body[0]
        f.px, f.py, f.vx, f.vy, f.a, f.w = xy[t, 0], xy[t, 1], vel[t, 0], vel[t, 1], ang[t], 0.0
        touch[t] = bool(lib.d2o_frame_touches(C.byref(f), C.byref(obst), C.byref(cfg)))
        s.steps = t + 1
        s.collided = int(bool(ended_in_collision) and t == T - 1)
        lib.d2o_after_physics(C.byref(s), C.byref(path), C.byref(obst), C.byref(cfg), C.byref(out))
        total += out.reward
        terms += (out.r_CA, out.r_PA, out.r_PP, out.r_coll, out.r_end, out.r_AA)
        if out.done and done_at is None:
            done_at = t + 1
    return {"total_reward": total, "ape": out.ape, "terms": terms, "done_at": done_at, "touch": touch,
            "end": (out.end_collision, out.end_reached, out.end_time, out.end_alpha)}
