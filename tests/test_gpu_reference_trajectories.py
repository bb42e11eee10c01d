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


@pytest.mark.parametrize("scenario", SCENARIOS)
def test_gpu_replay_of_genuine_episodes_reproduces_reward_ape_and_termination(scenario):
    """The product path itself on genuine data: the step kernel in replay mode (d2d_debug_skip_physics: everything
    after space.step) is fed the reference's recorded positions, their first differences as velocity and the
    thrust-direction angle estimate, exactly like oracle/flight_replay.py does for the oracle.  Per episode it must
    give the APE the reference wrote (1e-9), end on the reference's last step for the same reason, and sum up the
    reference's total reward (same bound as the CPU test); it must also agree with the oracle's replay of the same
    inputs to rounding.  Episodes in which the angle estimate makes the frame box touch at the wrong step (the
    oracle's own collision test decides which) are skipped: the kernel would, rightly, end them elsewhere."""
    import torch
    from oracle import binding as B
    from oracle import flight_replay as FR
    pkg = importlib.import_module(PKG)
    z = np.load(os.path.join(GOLDEN, "reference_trajectories.npz"))
    run_cfg = json.load(open(os.path.join(GOLDEN, "reference_trajectory_stats.json")))["run_config"]
    lens = z[scenario + "__len"]
    xy_all, off, eps = z[scenario + "__xy"], 0, []
    ocfg = B.default_config(**run_cfg)
    for k, n in enumerate(lens):
        xy = xy_all[off:off + n]
        off += n
        ang = FR.estimate_angle(xy)
        col = int(z[scenario + "__collision"][k])
        ref = FR.replay_rewards(scenario, xy, col, cfg=ocfg, angles=ang)
        if ref["touch"][:-1].any() or bool(ref["touch"][-1]) != bool(col):
            continue
        vel = np.empty_like(xy)
        vel[:-1] = (xy[1:] - xy[:-1]) * 60.0
        vel[-1] = 2 * vel[-2] - vel[-3]
        eps.append(dict(xy=xy, vel=vel, ang=ang, col=col, total=float(z[scenario + "__total_reward"][k]),
                        ape=float(z[scenario + "__ape"][k]), oracle=ref))
    assert len(eps) >= 6, "too few usable episodes in the fixture"
    n = len(eps)
    env = pkg.Drone2dVecEnv(n, seed=0, info_mode="full", mode="test", scenario=scenario, **run_cfg)
    env.reset_device()
    env.debug_skip_physics(True)
    zero = torch.zeros((n, 2), device=env.device)
    st = env.get_state()
    st[:, 18:36] = 0.0                                       # impulses, flags, counters, accumulators: a fresh episode
    st[:, 31] = 0.0
    total, ape, done_at, bits = np.zeros(n), np.zeros(n), [None] * n, [0] * n
    for t in range(max(len(e["xy"]) for e in eps)):
        for i, e in enumerate(eps):
            if t < len(e["xy"]):
                st[i, 0:6] = (e["xy"][t, 0], e["xy"][t, 1], e["vel"][t, 0], e["vel"][t, 1], e["ang"][t], 0.0)
        env.set_state(st)
        env.step_device(zero, auto_reset=False, terminal_obs=False)
        info = env.info.cpu().numpy()
        st = env.get_state()
        for i, e in enumerate(eps):
            if t < len(e["xy"]):
                total[i], ape[i] = info[i, 10], info[i, 9]
                if info[i, 11] != 0 and done_at[i] is None:
                    done_at[i], bits[i] = t + 1, int(info[i, 11])
    env.close()
    for i, e in enumerate(eps):
        T = len(e["xy"])
        assert done_at[i] == T, (scenario, i, done_at[i], T)
        assert bool(bits[i] & 1) == bool(e["col"])
        assert abs(ape[i] - e["ape"]) < 1e-9, (ape[i], e["ape"])
        assert abs(total[i] - e["total"]) < 6.0, (total[i], e["total"])
        assert abs(total[i] - e["oracle"]["total_reward"]) < 1e-6 * T + 1e-4, (total[i], e["oracle"]["total_reward"])
