#!/usr/bin/env python
"""Generate the committed golden fixtures under tests/golden/.

Run ONCE in the authoring container (needs /root/reference, read-only):

    python tests/golden/make_golden.py

It imports the reference's own pure-numpy modules (predef_path.py,
test_scenarios.py, obstacles.py, transformations.py) with pymunk / pygame /
matplotlib stubbed (see _ref_import.py), runs them on seeded inputs and stores
inputs + outputs.  The fixtures pin:

  scenarios.json      test_scenarios.create_test_scenario (test_scenarios.py:169-246)
  path_kats.json      QPMI2D.__call__/calculate_gradient/get_direction_angle/
                      get_closest_u/get_lookahead_point (predef_path.py:88-266)
                      on the 7 scenario paths and 8 random curriculum paths
  waypoint_gen.json   generate_random_waypoints_2d (predef_path.py:307-363): raw
                      np.random draws (recorded by wrapping np.random.uniform)
                      and the waypoints they produce
  obstacle_gen.json   generate_obstacles_around_path (obstacles.py:58-89): raw
                      draws and the accepted circles
  transforms.json     ssa / R_w_b (transformations.py:6-11)
  closed_loop.json    the reference's own evaluation artefacts
                      Tests/<setting>/test_k/results.txt (statistical fixtures)
  agent_19_policy.npz actor weights of ppo_agents/PFCA_see_3_obs_19_90.zip (agent_19_value.npz: its critic)
                      (data, not source; needed on the GPU box where the
                      reference tree does not exist)

The physics step (pymunk / Chipmunk2D) is NOT importable here, so no fixture
pins it: see oracle/README.md ("parity unpinned" for the rigid-body step).
"""
import io
import json
import os
import sys
import zipfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import _ref_import  # noqa: E402

pp, tr, ob, ts = _ref_import.load()
import pymunk  # the stub  # noqa: E402
from scipy.optimize import fminbound  # noqa: E402

W = H = 1300
SCENARIOS = ["perpendicular", "parallel", "S_parallel", "corridor", "S_corridor", "impossible", "large"]


def fl(x):
    return float(x)


def arr(a):
    return np.asarray(a, dtype=np.float64).tolist()


def dump(name, obj):
    with open(os.path.join(HERE, name), "w") as f:
        json.dump(obj, f, indent=None, separators=(",", ":"))
    print("wrote", name, os.path.getsize(os.path.join(HERE, name)), "bytes")


def path_record(path):
    return {
        "wps": arr(path.waypoints),
        "us": arr(path.us),
        "x_params": arr(path.x_params),
        "y_params": arr(path.y_params),
        "length": fl(path.length),
    }


def path_kats(path, rng, n_pos=24):
    L = path.length
    us = path.us
    grid = list(np.linspace(-10.0, L + 10.0, 41))
    grid += [0.0, -5.0, -1e-9, fl(us[1]), fl(us[1]) + 1e-9, fl(us[-2]) - 0.001, fl(us[-2]) - 0.0011,
             fl(us[-2]), L, L + 5.0, 0.5 * (fl(us[1]) + fl(us[2])), fl(us[2])]
    evals = []
    for u in grid:
        u = float(u)
        p = path(u)
        g = path.calculate_gradient(u)
        evals.append({"u": u, "p": arr(p), "grad": arr(g), "dir": fl(path.get_direction_angle(u))})
    closest = []
    for _ in range(n_pos):
        pos = [fl(rng.uniform(0, W)), fl(rng.uniform(0, H))]
        # identical call to predef_path.py:246-247 but with full_output to record nfev
        xopt, fval, ierr, nfev = fminbound(lambda u: np.linalg.norm(path(u) - pos), full_output=1,
                                           x1=-10.0, x2=L + 10.0, xtol=1e-6, maxfun=500)
        u_ref = path.get_closest_u(pos)
        assert float(u_ref) == float(xopt)
        cp = path.get_closest_position(pos)
        la = path.get_lookahead_point(pos, 220)
        closest.append({"pos": pos, "u": fl(xopt), "fval": fl(fval), "nfev": int(nfev),
                        "closest": arr(cp), "lookahead": arr(la)})
    return {"evals": evals, "closest": closest}


def gen_scenarios():
    out = {}
    for s in SCENARIOS:
        wps, path, obst = ts.create_test_scenario(pymunk.Space(), s, W, H)
        rec = path_record(path)
        rec["obstacles"] = [[fl(o.x_pos), fl(o.y_pos), fl(o.radius)] for o in obst]
        out[s] = rec
    dump("scenarios.json", out)


class DrawRecorder:
    """Wrap np.random.uniform/normal so the raw draws made by reference code are logged."""

    def __enter__(self):
        self.log = []
        self._u, self._n = np.random.uniform, np.random.normal

        def uniform(lo=0.0, hi=1.0, size=None):
            v = self._u(lo, hi, size)
            self.log.append(["uniform", fl(lo), fl(hi), fl(v)])
            return v

        def normal(mu=0.0, sd=1.0, size=None):
            v = self._n(mu, sd, size)
            self.log.append(["normal", fl(mu), fl(sd), fl(v)])
            return v

        np.random.uniform, np.random.normal = uniform, normal
        return self

    def __exit__(self, *a):
        np.random.uniform, np.random.normal = self._u, self._n


def gen_paths_and_generators():
    rng = np.random.RandomState(1234)
    kats = {}
    for s in SCENARIOS:
        _, path, _ = ts.create_test_scenario(pymunk.Space(), s, W, H)
        kats[s] = {"path": path_record(path), **path_kats(path, rng)}

    wp_gen = []
    ob_gen = []
    corners = ["DL", "DR", "UL", "UR"]
    for k in range(8):
        np.random.seed(100 + k)
        scen = corners[k % 4]
        with DrawRecorder() as rec:
            wps = pp.generate_random_waypoints_2d(12, 100, scen, screen_x=W, screen_y=H)
        path = pp.QPMI2D(wps)
        wp_gen.append({"scen": scen, "n_wps": 12, "distance": 100, "draws": rec.log, "wps": arr(wps)})
        kats["random_%d_%s" % (k, scen)] = {"path": path_record(path), **path_kats(path, rng, n_pos=16)}
        # obstacle sampler on this path: 3 off-path (mean 0, std 100) then 1 on-path (stage_5 recipe, env:370-372)
        with DrawRecorder() as rec:
            off = ob.generate_obstacles_around_path(3, pymunk.Space(), path, 0, 100)
        with DrawRecorder() as rec2:
            on = ob.generate_obstacles_around_path(1, pymunk.Space(), path, 0, 0, onPath=True)
        ob_gen.append({
            "path": path_record(path),
            "off_draws": rec.log, "off": [[fl(o.x_pos), fl(o.y_pos), fl(o.radius)] for o in off],
            "on_draws": rec2.log, "on": [[fl(o.x_pos), fl(o.y_pos), fl(o.radius)] for o in on],
        })
    dump("path_kats.json", kats)
    dump("waypoint_gen.json", wp_gen)
    dump("obstacle_gen.json", ob_gen)


def gen_transforms():
    rng = np.random.RandomState(7)
    angles = list(rng.uniform(-20, 20, 32)) + [0.0, np.pi, -np.pi, 3 * np.pi, -3 * np.pi, 2 * np.pi, 1e-300]
    out = {"ssa": [[fl(a), fl(tr.ssa(a))] for a in angles],
           "R_w_b": [[fl(a), arr(tr.R_w_b(a))] for a in angles[:8]]}
    dump("transforms.json", out)


def gen_closed_loop():
    """Parse the reference's evaluation artefacts (main.py:305-327 format)."""
    root = "/root/reference/Tests"
    settings = {s: "%s/%s/test_4" % (root, s) for s in SCENARIOS}
    settings.update({"stage_%d" % k: "%s/stage_%d/test_1" % (root, k) for k in range(1, 6)})
    out = {}
    for name, d in settings.items():
        txt = [f for f in os.listdir(d) if f.endswith("results.txt")][0]
        rec = {}
        for line in open(os.path.join(d, txt)):
            k, _, v = line.partition(":")
            rec[k.strip()] = v.strip()
        out[name] = {
            "source": os.path.relpath(os.path.join(d, txt), "/root/reference"),
            "successes": int(rec["Successes"]), "fails": int(rec["Fails"]), "collisions": int(rec["Collisions"]),
            "mean_ape": float(rec["Average APE"]), "mean_flight_time": float(rec["Average flight time"]),
            "agent": rec["Agent path"],
            "n_episodes": int(len(np.load(os.path.join(d, "apes.npy")))),
        }
    dump("closed_loop.json", out)


def gen_policy():
    import torch
    z = zipfile.ZipFile("/root/reference/ppo_agents/PFCA_see_3_obs_19_90.zip")
    sd = torch.load(io.BytesIO(z.read("policy.pth")), map_location="cpu", weights_only=True)
    keep = {k.replace(".", "__"): v.numpy() for k, v in sd.items()
            if k.startswith(("mlp_extractor.policy_net", "action_net", "log_std"))}
    np.savez(os.path.join(HERE, "agent_19_policy.npz"), **keep)
    print("wrote agent_19_policy.npz", {k: v.shape for k, v in keep.items()})
    # the critic of the same checkpoint (value tower of the SB3 MlpPolicy): used by the rollout-collection tests
    val = {k.replace(".", "__"): v.numpy() for k, v in sd.items() if k.startswith(("mlp_extractor.value_net", "value_net"))}
    np.savez(os.path.join(HERE, "agent_19_value.npz"), **val)
    print("wrote agent_19_value.npz", {k: v.shape for k, v in val.items()})


if __name__ == "__main__":
    gen_scenarios()
    gen_paths_and_generators()
    gen_transforms()
    gen_closed_loop()
    gen_policy()
