"""Inverse dynamics on GENUINE reference trajectories (TEST INFRASTRUCTURE ONLY).

The reference's evaluation runs saved the frame position after every `space.step` (Tests/<scenario>/test_4/
flight_paths, env:410-415,984-986) but not the actions.  The actions can be recovered: Chipmunk integrates
`p += v*dt` with the velocity the previous step's solver left (cpBodyUpdatePosition), so first differences of the
recorded positions ARE the frame velocity after each step, and that velocity is a function of the hidden state and
of the two motor commands of the step.  At the start of an episode the hidden state is known exactly (Drone.__init__,
Drone.py:9-95: three bodies at rest, cached impulses zero) up to the spawn angle, which the direction of the first
thrust gives.  Hence, IF the restated rigid-body step (oracle/d2d_oracle.c:d2o_space_step) is the step pymunk took,
there is an action sequence inside the action Box [-1,1]^2, on the float32 grid (SB3 `predict` returns clipped
float32 actions, main.py:262), that reproduces every recorded velocity to rounding error; a restatement that differs
(solver iterations, warm start, joint order, integration order, masses ...) leaves a residual and recovers actions
that miss the Box and the grid.

Numerics.  The map actions -> frame velocities is a sampled system of relative degree three in the differential
thrust (torque -> angular rate -> angle -> thrust direction); its sampling zeros lie outside the unit circle, so a
step-by-step (causal) inversion is unstable, and so is any scheme that demands an exactly vanishing residual (the
smallest error in a committed action would have to be compensated by an exponentially growing alternating sequence).
The inverse is therefore taken non-causally and inside the action Box: box-constrained Gauss-Newton (active set,
truncated SVD) over a sliding window of W steps that starts from an exactly known state.  Only the OLDEST action of
the window is committed, after W - 1 later steps have pinned it; values that come out on the edge of the Box are
set to exactly +-1 (SB3 clips).  The window grows one step at a time from the spawn (continuation), so Newton stays on
the branch of the physical solution.  What is reported is the residual max|v_model - v_recorded| of every committed
step under the committed actions, and how many recovered actions sit exactly on the edge of the Box.
"""
import ctypes as C

import numpy as np

from . import binding as B


def snap32(a):
    """Nearest point of the action space: clip to the Box, round to float32."""
    return np.clip(a, -1.0, 1.0).astype(np.float32).astype(np.float64)


class Episode:
    def __init__(self, xy, cfg=None, lib=None):
        self.cfg = cfg or B.default_config()
        self.lib = lib or B.lib()
        self.xy = np.ascontiguousarray(xy, dtype=np.float64)
        self.dt = self.cfg.dt
        self.n = len(self.xy) - 1                             # steps whose outcome (a velocity) was recorded
        # vobs[j]: frame velocity after step j (0-based); xy[j] is the position after step j, xy[0] the spawn point
        self.vobs = (self.xy[1:] - self.xy[:-1]) / self.dt
        g = np.array([self.cfg.gravity_x, self.cfg.gravity_y])
        vprev = np.vstack([[0.0, 0.0], self.vobs[:-1]])
        self.acc = (self.vobs - vprev) / self.dt - g          # what thrust and joints did to the frame in step j
        self.mag = np.hypot(self.acc[:, 0], self.acc[:, 1])
        self.theta_obs = np.arctan2(-self.acc[:, 0], self.acc[:, 1])      # ~ frame angle at the start of step j

    # ---- forward model (C) ----
    def rollout(self, state, actions, jac=False, h=1e-6):
        a = np.ascontiguousarray(actions, dtype=np.float64).reshape(-1, 2)
        m = len(a)
        v = np.empty(2 * m)
        J = np.empty((2 * m, 2 * m)) if jac else None
        dp = C.POINTER(C.c_double)
        self.lib.d2o_frame_velocity_rollout(C.byref(state), a.ctypes.data_as(dp), m, C.byref(self.cfg), v.ctypes.data_as(dp),
                                            J.ctypes.data_as(dp) if jac else None, h)
        return (v, J) if jac else v

    def advance(self, state, action):
        none = B.make_obstacles([])
        self.lib.d2o_space_step(C.byref(state), C.byref(none), C.byref(self.cfg), float(action[0]), float(action[1]))
        return np.array([state.body[0].vx, state.body[0].vy])

    def rigid_guess(self, k):
        """Starting value for the action of step k from the rigid-body picture (thrust magnitude -> sum, second
        difference of the thrust direction -> difference)."""
        cfg = self.cfg
        M = cfg.mass_f + 2 * cfg.mass_m
        s = np.clip((self.mag[k] * M - 2 * 0.5 * cfg.force_scale) / (0.5 * cfg.force_scale), -2, 2)
        d = 0.0
        if k + 2 < self.n and min(self.mag[k], self.mag[k + 1], self.mag[k + 2]) > 150:
            R = cfg.width / 2 - cfg.height / 2
            i_tot = (self.lib.d2o_moment_for_box(cfg.mass_f, cfg.width, cfg.height / 2) +
                     2 * (self.lib.d2o_moment_for_box(cfg.mass_m, cfg.height, cfg.height) + cfg.mass_m * R * R))
            th = np.unwrap(self.theta_obs[k:k + 3])
            d = float(np.clip((th[2] - 2 * th[1] + th[0]) / self.dt ** 2 * i_tot / (R * 0.5 * cfg.force_scale), -2, 2))
        return np.clip([(s - d) / 2, (s + d) / 2], -1, 1)

    # ---- the inversion ----
    def recover_actions(self, W=40, max_steps=None, sv_cut=1e-3, snap=True):
        """Returns dict(actions [m,2] committed (snapped), raw [m,2] as estimated, residual [m] = max |v_model - v_obs|
        of each committed step under the committed actions, alpha0)."""
        n = self.n if max_steps is None else min(self.n, max_steps)
        # Spawn angle = direction of the first thrust.  With both motor commands at -1 the thrust is exactly zero and so
        # is the torque: the drone falls without turning until the first step with thrust, whose direction is the angle.
        lit = np.flatnonzero(self.mag > 1.0)
        j0 = int(lit[0]) if len(lit) else 0
        alpha0 = float(self.theta_obs[j0])
        self.alpha0_exact = bool(j0 == 0 or self.mag[:j0].max() < 1e-6)
        state = B.spawn(self.xy[0, 0], self.xy[0, 1], alpha0, self.cfg)   # exact state at the start of the window
        k0 = 0                                                          # first step of the window
        x = np.zeros((0, 2))                                            # current estimate of the window's actions
        committed, raw, resid = [], [], []

        def solve(x, k0):
            """Box-constrained Gauss-Newton (active set): entries sitting on an edge of the Box with the gradient
            pointing outwards are held there; the others move by a truncated-SVD step, projected onto the Box."""
            w = len(x)
            target = self.vobs[k0:k0 + w].ravel()
            x = np.clip(x, -1.0, 1.0)
            for _ in range(10):
                v, J = self.rollout(state, x, jac=True)
                r = v - target
                xf = x.ravel()
                g = J.T @ r
                held = ((xf >= 1.0) & (g < 0)) | ((xf <= -1.0) & (g > 0))
                free = np.flatnonzero(~held)
                if len(free) == 0:
                    break
                U, S, Vt = np.linalg.svd(J[:, free], full_matrices=False)
                keep = S > sv_cut
                d = -(Vt[keep].T * (1.0 / S[keep])) @ (U[:, keep].T @ r)
                c0, sc = r @ r, 1.0
                while True:
                    xn = xf.copy()
                    xn[free] = np.clip(xf[free] + sc * d, -1.0, 1.0)
                    rn = self.rollout(state, xn) - target
                    if rn @ rn < c0 or sc < 1e-3:
                        break
                    sc /= 2
                if rn @ rn >= c0:
                    break
                moved = np.abs(xn - xf).max()
                x = xn.reshape(w, 2)
                if moved < 1e-13:
                    break
            return x

        for m in range(1, n + 1):                                       # m = number of steps covered so far
            x = np.vstack([x, self.rigid_guess(m - 1)])
            if len(x) >= 3:
                x = solve(x, k0)
            while len(x) > W or (m == n and len(x) > 0):
                a_raw = x[0].copy()
                # values on the edge of the Box are exact (SB3 clips); interior ones are committed as estimated
                # (rounding them to float32 here would inject half an ulp, far more than the estimation error)
                a = np.where(np.abs(a_raw) >= 1.0 - 1e-7, np.sign(a_raw), a_raw) if snap else a_raw
                v = self.advance(state, a)
                committed.append(a); raw.append(a_raw); resid.append(float(np.abs(v - self.vobs[k0]).max()))
                x = x[1:]; k0 += 1
        return {"actions": np.array(committed), "raw": np.array(raw), "residual": np.array(resid), "alpha0": alpha0,
                "alpha0_exact": self.alpha0_exact}


def summarize(rec, tail=24):
    """Statistics over the committed steps that had at least `tail` later steps to pin them."""
    A, raw, res = rec["actions"][:-tail], rec["raw"][:-tail], rec["residual"][:-tail]
    if len(A) == 0:
        return None
    a32 = snap32(raw)
    spacing = np.spacing(np.abs(a32).astype(np.float32)).astype(np.float64)
    rel = np.abs(raw - a32) / spacing
    inner = (np.abs(raw) > 0.25) & (np.abs(raw) < 1.0 - 1e-6)
    return {
        "steps": int(len(A)),
        "residual_median": float(np.median(res)), "residual_p99": float(np.percentile(res, 99)), "residual_max": float(res.max()),
        "frac_residual_below_1e-8": float((res < 1e-8).mean()),
        "frac_actions_at_box_edge": float((np.abs(A) == 1.0).mean()),
        "frac_raw_within_1e-6_of_edge": float((np.abs(np.abs(raw) - 1.0) < 1e-6).mean()),
        "frac_raw_outside_box": float((np.abs(raw) > 1.0 + 1e-6).mean()),
        # distance of the un-snapped estimate to the float32 grid in units of the grid spacing (uniform: median 0.25)
        "grid_offset_median": float(np.median(rel[inner])) if inner.any() else None,
        "grid_offset_p90": float(np.percentile(rel[inner], 90)) if inner.any() else None,
    }
