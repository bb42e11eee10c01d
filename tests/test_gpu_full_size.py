"""Size-independent properties at BASELINE.json's full sizes (65 536 envs per GPU), where the oracle
is too slow to follow: determinism, conservation of the episode bookkeeping, sane ranges, and
independence from how the batch is split."""
import importlib

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
PKG = "drone_2d_custom_gym_env_for_reinforcement_learning_b200"
N = 65536


def _run(pkg, n, steps, seed, base=0, scenario="stage_5", mode="curriculum"):
    import torch
    env = pkg.Drone2dVecEnv(n, seed=seed, env_id_base=base, info_mode="none", mode=mode, scenario=scenario)
    env.reset_device()
    dones = torch.zeros((), dtype=torch.int64, device=env.device)
    rsum = torch.zeros((), dtype=torch.float64, device=env.device)
    for _ in range(steps):
        obs, rew, done = env.step_device(None)          # in-kernel Philox actions keyed by (seed, global env id, step)
        dones += done.sum()
        rsum += rew.double().sum()
        assert bool(torch.isfinite(obs).all()) and bool(torch.isfinite(rew).all())
    st = env.get_state()
    stats = env.episode_stats()
    out = dict(state=st, stats=stats, dones=int(dones), rsum=float(rsum), obs=obs.cpu().numpy(), launches=env.launch_count)
    env.close()
    return out


def test_full_batch_determinism_and_bookkeeping():
    pkg = importlib.import_module(PKG)
    a = _run(pkg, N, 160, seed=1)
    b = _run(pkg, N, 160, seed=1)
    assert np.array_equal(a["state"], b["state"])                      # bit-reproducible
    for k, v in a["stats"].items():
        # counts are exact; the float sums are accumulated with atomics (order-dependent rounding)
        assert v == b["stats"][k] or abs(v - b["stats"][k]) <= 1e-11 * abs(v), k
    s = a["stats"]
    assert s["episodes"] == a["dones"] == a["state"][:, 35].sum()      # every done = one episode = one reset
    assert s["n_successful_runs"] + s["n_failed_runs"] >= s["episodes"] > N // 4
    assert abs(s["env_steps"] + a["state"][:, 32].sum() - 160 * N) < 0.5   # finished + running steps = all steps
    assert s["inlane_resets"] <= 1e-4 * s["episodes"]               # (almost) every auto-reset used its prefetched episode
    assert a["launches"] == 162                                        # reset + one launch per step + the get_state gather
    # reward bookkeeping: finished episodes' returns + running returns = sum of all rewards (float32 outputs)
    total = s["total_reward"] + a["state"][:, 34].sum()
    assert abs(total - a["rsum"]) <= 1e-5 * abs(total) + 1.0
    st = a["state"]
    assert np.all(np.abs(st[:, 4]) < np.pi / 2 + 0.3)                  # |alpha| ends episodes at pi/2 (one step of slack)
    assert np.all(st[:, 32] <= 1100)
    obs = a["obs"]
    assert np.all(np.abs(obs[:, [9, 10, 12, 13, 15, 16, 17, 18, 23, 24, 25, 26]]) <= 1.0 + 1e-6)   # sines / cosines


def test_full_batch_equals_two_half_batches():
    pkg = importlib.import_module(PKG)
    full = _run(pkg, N, 60, seed=2)
    lo = _run(pkg, N // 2, 60, seed=2, base=0)
    hi = _run(pkg, N // 2, 60, seed=2, base=N // 2)
    assert np.array_equal(full["state"], np.concatenate([lo["state"], hi["state"]]))
    assert full["stats"]["episodes"] == lo["stats"]["episodes"] + hi["stats"]["episodes"]


def test_shared_table_scenario_full_batch():
    pkg = importlib.import_module(PKG)
    a = _run(pkg, N, 80, seed=3, scenario="S_corridor", mode="test")
    assert a["stats"]["episodes"] == a["dones"] > 0
    assert a["stats"]["n_collisions"] > 0                              # 58 obstacles around the path: random flight hits them


def test_long_run_stays_finite_and_consistent():
    """3000 steps at full size (about 1.5 M episodes): no NaN/Inf, bookkeeping still closes, prefetch never misses."""
    pkg = importlib.import_module(PKG)
    a = _run(pkg, N, 3000, seed=7)
    s = a["stats"]
    assert s["episodes"] == a["dones"] == a["state"][:, 35].sum() > 10 * N
    assert abs(s["env_steps"] + a["state"][:, 32].sum() - 3000 * N) < 0.5
    assert s["inlane_resets"] <= 1e-4 * s["episodes"]             # episodes shorter than the 2-step prefetch pipeline are rare
    assert np.isfinite(a["state"]).all()


def test_inlane_fallback_is_used_when_episodes_are_shorter_than_the_prefetch_pipeline():
    import torch
    pkg = importlib.import_module(PKG)
    env = pkg.Drone2dVecEnv(4096, seed=1, info_mode="none", mode="curriculum", scenario="stage_5", n_steps=1)
    env.reset_device()
    for _ in range(20):
        env.step_device(None)
    s = env.episode_stats()
    # step 1 consumes the episodes staged by reset(); after that a staged episode is only in time when the refill
    # blocks happen to read the env's episode counter after the lane has advanced it (a benign race: the content of
    # an episode is a pure function of (env, episode) either way) -- the large majority must take the in-lane path
    print("in-lane resets: %d of %d episodes" % (s["inlane_resets"], s["episodes"]))
    assert s["episodes"] == 20 * 4096 and 14 * 4096 <= s["inlane_resets"] <= 19 * 4096
    env.close()


def test_full_size_strided_oracle_check():
    """Parity against the oracle AT the full batch size: 65 536 stage_5 envs in steady state, every 64th env (1024
    of them) is followed by the CPU oracle for single steps from the device's own state and tables."""
    import ctypes as C
    import torch
    import oracle
    from parity_util import blob_to_states, oracle_config_from, tables_to_structs, TOL_STATE_STEP, TOL_OBS_CORE, TOL_U
    pkg = importlib.import_module(PKG)
    env = pkg.Drone2dVecEnv(N, seed=9, info_mode="full", mode="curriculum", scenario="stage_5")
    cfg = oracle_config_from(env)
    env.reset_device()
    for _ in range(150):
        env.step_device(None)
    idx = np.arange(0, N, 64)
    rng = np.random.RandomState(3)
    checked = flags = nearest = 0
    worst = dict(state=0.0, obs=0.0, u=0.0, reward=0.0)
    for rnd in range(3):
        blob = env.get_state()
        tabs = env.get_tables()
        states = blob_to_states(blob[idx])
        paths, obsts = tables_to_structs(tabs, env.cfg.n_wps, idx)
        actions = rng.uniform(-1, 1, (N, 2)).astype(np.float32)
        env.step_device(torch.from_numpy(actions).to(env.device))
        g_obs, g_info, g_done = env.obs64[idx].cpu().numpy(), env.info[idx].cpu().numpy(), env.dones[idx].cpu().numpy()
        after = env.get_state()[idx]
        for k, i in enumerate(idx):
            out = oracle.env_step(states[k], paths[i], obsts[i], cfg, actions[i, 0], actions[i, 1])
            checked += 1
            bits = out.end_collision | 2 * out.end_reached | 4 * out.end_time | 8 * out.end_alpha
            flags += int(bool(out.done) != bool(g_done[k]) or bits != int(g_info[k, 11]))
            near = [(int(g_info[k, 14]) >> (8 * j)) & 255 for j in range(3)]
            nearest += int([-1 if v == 255 else v for v in near] != list(out.near_idx[:]))
            worst["obs"] = max(worst["obs"], np.abs(g_obs[k, :19] - np.array(out.obs[:19])).max())
            worst["u"] = max(worst["u"], abs(g_info[k, 12] - out.u_star))
            worst["reward"] = max(worst["reward"], abs(g_info[k, 0] - out.reward))
            if not out.done:
                s = states[k]
                ob = np.array([[b.px, b.py, b.vx, b.vy, b.a, b.w] for b in s.body]).ravel()
                worst["state"] = max(worst["state"], np.abs(after[k, :18] - ob).max())
    env.close()
    assert checked == 3 * len(idx) and flags == 0 and nearest == 0, (checked, flags, nearest)
    assert worst["state"] <= TOL_STATE_STEP and worst["obs"] <= TOL_OBS_CORE and worst["u"] <= TOL_U, worst
    assert worst["reward"] <= 1e-6 + 10 * worst["u"], worst
