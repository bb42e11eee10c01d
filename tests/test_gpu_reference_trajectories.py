"""The CUDA path against GENUINE reference outputs (see tests/test_reference_trajectories_cpu.py for the
fixtures): closest-u on the positions the real drone flew through must reproduce the APE the reference
wrote, and the closed loop of the shipped agent on the GPU must show the same implied-thrust distribution
as the genuine pymunk trajectories."""
import importlib
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN

pytestmark = pytest.mark.gpu
PKG = "drone_2d_custom_gym_env_for_reinforcement_learning_b200"
SCENARIOS = ["perpendicular", "parallel", "S_parallel", "corridor", "S_corridor", "impossible", "large"]


@pytest.mark.parametrize("scenario", SCENARIOS)
def test_ape_of_genuine_episodes_from_the_kernels_closest_u(scenario):
    from oracle import binding as B
    pkg = importlib.import_module(PKG)
    z = np.load(os.path.join(GOLDEN, "reference_trajectories.npz"))
    xy, lens, apes = z[scenario + "__xy"], z[scenario + "__len"], z[scenario + "__ape"]
    env = pkg.Drone2dVecEnv(len(xy), seed=1, mode="test", scenario=scenario)
    env.reset()
    u, nfev = env.closest_u(xy)
    env.close()
    path, _ = B.load_scenario(scenario)
    # the oracle's own search on the same points: bit-identical parameter and evaluation count
    uo = np.array([B.closest_u(path, x, y)[:2] for x, y in xy])
    assert np.array_equal(u, uo[:, 0]) and np.array_equal(nfev, uo[:, 1].astype(nfev.dtype))
    pts = np.array([B.path_eval(path, v) for v in u])
    dist = np.sqrt((pts[:, 1] - xy[:, 1]) ** 2 + (pts[:, 0] - xy[:, 0]) ** 2)
    off = 0
    for n, ape in zip(lens, apes):
        assert abs(dist[off:off + n].mean() - ape) < 1e-9, (scenario, dist[off:off + n].mean(), ape)
        off += n


def test_implied_thrust_of_the_gpu_closed_loop_matches_genuine_pymunk():
    import torch
    from oracle import flight_replay as FR
    pkg = importlib.import_module(PKG)
    policy = importlib.import_module(PKG + ".policy")
    ref = json.load(open(os.path.join(GOLDEN, "reference_trajectory_stats.json")))["scenarios"]["corridor"]["implied_thrust"]
    n = 1024
    env = pkg.Drone2dVecEnv(n, seed=11, mode="test", scenario="corridor")
    actor = policy.MlpActor.load(os.path.join(GOLDEN, "agent_19_policy.npz"), env.device)
    g = torch.Generator(device=env.device)
    g.manual_seed(5)
    obs = env.reset_device()
    alive = np.ones(n, bool)
    flights = [[] for _ in range(n)]
    for _ in range(env.cfg.n_steps):
        obs, _, done = env.step_device(actor.predict(obs, False, g), auto_reset=False, terminal_obs=False)
        st = env.get_state()
        d = done.cpu().numpy() != 0
        for i in np.flatnonzero(alive):
            flights[i].append((st[i, 0], st[i, 1]))
        alive &= ~d
        if not alive.any():
            break
    env.close()
    s = FR.implied_thrust_stats([np.array(f) for f in flights])
    pct = ref["percentiles"]
    a, b = np.array(s["thrust_acc"]), np.array(ref["thrust_acc"])
    for p, tol in ((50, 25), (75, 25), (95, 12), (99, 8)):
        k = pct.index(p)
        assert abs(a[k] - b[k]) < tol, (p, a[k], b[k])
    assert abs(s["frac_above_2000"] - ref["frac_above_2000"]) < 0.004
    assert abs(s["frac_below_5"] - ref["frac_below_5"]) < 0.003
    assert abs(s["horizontal_drag_coefficient"]) < 0.01
