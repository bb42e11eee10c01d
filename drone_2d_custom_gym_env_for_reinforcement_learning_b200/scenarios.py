"""Host-side table builders for the shared test scenarios.

The step kernel consumes a path as (us, per-interior-waypoint quadratic rows)
and obstacles as circles.  `qpmi_tables` derives the path tables from waypoints
(QPMI2D.__init__, predef_path.py:10-50) and `scenario_tables` lays out the seven
deterministic evaluation scenarios of the reference (test_scenarios.py:4-246).
These run once per environment construction, on the host, in float64 numpy.
"""
import numpy as np

from .rl_config import TEST_SCENARIOS


def qpmi_tables(wps):
    """us[n] = cumulative chord length; row r = quadratic through waypoints
    r, r+1, r+2 parameterised by us (solved like the reference: inverse of the
    3x3 Vandermonde matrix times the coordinates, so the tables are bit-identical
    to predef_path.py:28-50)."""
    wps = np.asarray(wps, dtype=np.float64)
    n = len(wps)
    if n < 4:
        raise ValueError("a path needs at least 4 waypoints")
    us = np.concatenate(([0.0], np.cumsum(np.sqrt(np.sum(np.diff(wps, axis=0) ** 2, axis=1)))))
    xp = np.empty((n - 2, 3))
    yp = np.empty((n - 2, 3))
    for r in range(n - 2):
        u = us[r:r + 3]
        vand_inv = np.linalg.inv(np.stack([u ** 2, u, np.ones(3)], axis=1))
        xp[r] = vand_inv.dot(wps[r:r + 3, 0])
        yp[r] = vand_inv.dot(wps[r:r + 3, 1])
    return us, xp, yp


class PathTables:
    """Evaluates the blended piecewise-quadratic path on the host (only used to
    place scenario obstacles and by tests; the hot path evaluates on the GPU)."""

    def __init__(self, wps):
        self.wps = np.asarray(wps, dtype=np.float64)
        self.us, self.xp, self.yp = qpmi_tables(self.wps)
        self.length = float(self.us[-1])

    def _segment(self, u):
        return int(np.searchsorted(self.us[1:], u, side="left"))   # first n with u <= us[n+1]

    def _poly(self, r, u):
        return np.array([self.xp[r][0] * u ** 2 + self.xp[r][1] * u + self.xp[r][2],
                         self.yp[r][0] * u ** 2 + self.yp[r][1] * u + self.yp[r][2]])

    def _dpoly(self, r, u):
        return np.array([self.xp[r][0] * u * 2 + self.xp[r][1], self.yp[r][0] * u * 2 + self.yp[r][1]])

    def point(self, u):
        us = self.us
        if us[0] <= u <= us[1]:
            return self._poly(0, u)
        if u >= us[-2] - 0.001:
            return self._poly(-1, u)
        n = self._segment(u)
        w_next = (u - us[n]) / (us[n + 1] - us[n])
        w_prev = (us[n + 1] - u) / (us[n + 1] - us[n])
        return w_next * self._poly(n, u) + w_prev * self._poly(n - 1, u)

    def tangent_angle(self, u):
        us = self.us
        if us[0] <= u <= us[1]:
            d = self._dpoly(0, u)
        elif u >= us[-2]:
            d = self._dpoly(-1, u)
        else:
            n = self._segment(u)
            w_next = (u - us[n]) / (us[n + 1] - us[n])
            w_prev = (us[n + 1] - u) / (us[n + 1] - us[n])
            d = w_next * self._dpoly(n, u) + w_prev * self._dpoly(n - 1, u)
        return float(np.arctan2(d[1], d[0]))


def _polyline(start, steps):
    """Waypoints obtained by chaining (length, azimuth) steps from `start`."""
    pts = [np.array(start, dtype=np.float64)]
    for length, az in steps:
        pts.append(np.array([pts[-1][0] + length * np.cos(az), pts[-1][1] + length * np.sin(az)]))
    return np.array(pts)


def scenario_waypoints(scenario, W, H, y_offset=0):
    """generate_scen_waypoints_2d (test_scenarios.py:87-167)."""
    if scenario in ("perpendicular", "parallel", "impossible", "corridor"):
        n, seg = 10, 100
        return _polyline((W / 2 - seg * (n - 1) / 2, H / 2 + y_offset), [(seg, 0)] * (n - 1))
    if scenario in ("S_parallel", "S_corridor"):
        n, seg, x0 = (6, 300, W / 10) if scenario == "S_parallel" else (7, 200, W / 7)
        zigzag = [(seg, -np.pi / 4 if k % 2 == 0 else np.pi / 4) for k in range(n - 1)]
        return _polyline((x0, H / 2 + y_offset), zigzag)
    if scenario == "large":
        n = int(W / 100)
        obs_rad, margin, lead = W / 5, 80, W / 10
        arc_seg = (2 * np.pi * (obs_rad + margin) / 2) / (n - 3)
        pts = [np.array([W / 2 - obs_rad - margin - lead, H / 2 - margin])]
        pts.append(np.array([pts[0][0] + lead, pts[0][1]]))
        for k in range(1, n - 1):
            az = np.pi / 2 - (k - 1) * np.pi / (n - 3)
            pts.append(np.array([pts[k][0] + arc_seg * np.cos(az), pts[k][1] + arc_seg * np.sin(az)]))
        pts.append(np.array([pts[-1][0] + lead, pts[-1][1]]))
        return np.array(pts)
    raise ValueError("unknown scenario %r" % (scenario,))


def _circles_along(path, us_at, radius):
    return [(*path.point(u), radius) for u in us_at]


def scenario_tables(scenario, W=1300, H=1300):
    """create_test_scenario (test_scenarios.py:169-246) -> (waypoints, us, xparams,
    yparams, obstacles[n][3]) with obstacles as (x, y, radius)."""
    if scenario not in TEST_SCENARIOS:
        raise ValueError("unknown scenario %r (expected one of %s)" % (scenario, ", ".join(TEST_SCENARIOS)))
    wps = scenario_waypoints(scenario, W, H)
    path = PathTables(wps)
    L = path.length
    circles = []
    if scenario == "perpendicular":
        n, r = 6, 20
        mid = path.point(L / 2)
        normal = path.tangent_angle(L / 2) - np.pi / 2
        direction = np.array([np.cos(normal), np.sin(normal)])
        circles = [(*(mid + (n * r - r - k * r * 2) * direction), r) for k in range(n)]
    elif scenario == "parallel":
        n, r = 6, 30
        first = (L - n * r * 2) / 2 - r
        circles = _circles_along(path, [first + k * r * 2 for k in range(1, n + 1)], r)
    elif scenario == "S_parallel":
        n, r = 20, 15
        first = (L - n * r * 2) / 2
        circles = _circles_along(path, [first + k * r * 2 for k in range(1, n + 1)], r)
    elif scenario in ("corridor", "S_corridor"):
        slots, off = (10, 100) if scenario == "corridor" else (30, 150)
        free = 100
        for sign in (+1, -1):
            wall = PathTables(scenario_waypoints(scenario, W, H, sign * off))
            r = (wall.length - free * 2) / (slots * 2)
            circles += _circles_along(wall, [k * r * 2 + free for k in range(1, slots)], r)
    elif scenario == "impossible":
        n, ring = 20, 100
        r = 2 * np.pi * ring / (n * 2)
        centre = path.point(L)
        heading = path.tangent_angle(L)
        step = 2 * np.pi / n
        circles = [(*(centre + ring * np.array([np.cos(heading - k * step), np.sin(heading - k * step)])), r)
                   for k in range(1, n + 1)]
    elif scenario == "large":
        circles = [(W / 2, H / 2, W / 5)]
    obstacles = np.array(circles, dtype=np.float64).reshape(-1, 3)
    return path.wps, path.us, path.xp, path.yp, obstacles


def spawn_rectangle(scenario, W=1300, H=1300):
    """Uniform spawn rectangle (xmin, xmax, ymin, ymax) of a test scenario (env:221-311)."""
    xmax = {"perpendicular": W / 2 - 100, "parallel": W / 2 - 300, "S_parallel": W / 2 - 300,
            "corridor": W / 2 - 400, "S_corridor": W / 2 - 450, "large": W / 2 - W / 4 - 50,
            "impossible": W / 2}[scenario]
    if scenario == "perpendicular":
        return 50.0, xmax, 50.0, H - 100.0
    return 50.0, xmax, 150.0, H - 300.0
