"""The reference-facing Python surface on the GPU: SB3 VecEnv contract (main.py:88-101,190) and the
single-env gym methods (env:33,394,908), info keys of env:575-613."""
import ctypes as C
import importlib

import numpy as np
import pytest

import oracle
from parity_util import blob_to_states, oracle_config_from

pytestmark = pytest.mark.gpu
PKG = "drone_2d_custom_gym_env_for_reinforcement_learning_b200"


@pytest.fixture(scope="module")
def pkg():
    return importlib.import_module(PKG)


def test_vec_env_contract(pkg):
    n = 300
    env = pkg.Drone2dVecEnv(n, seed=3, mode="curriculum", scenario="stage_5")
    assert env.num_envs == n and env.observation_space.shape == (27,) and env.action_space.shape == (2,)
    assert env.observation_space.dtype == np.float32 and float(env.action_space.low.min()) == -1.0
    obs = env.reset()
    assert isinstance(obs, np.ndarray) and obs.shape == (n, 27) and obs.dtype == np.float32
    rng = np.random.RandomState(0)
    seen_done = 0
    for t in range(150):
        obs, rew, done, infos = env.step(rng.uniform(-1, 1, (n, 2)).astype(np.float32))
        assert obs.shape == (n, 27) and rew.shape == (n,) and done.shape == (n,) and done.dtype == bool
        assert isinstance(infos, list) and len(infos) == n
        for i in np.flatnonzero(done):
            inf = infos[i]
            seen_done += 1
            assert inf["terminal_observation"].shape == (27,)
            for k in ("reward", "collision_avoidance_reward", "path_adherence", "path_progression", "collision_reward",
                      "reach_end_reward", "agressive_alpha_reward", "env_steps", "dist_closest_obs", "n_collisions",
                      "n_successful_runs", "n_failed_runs", "APE", "flight_path", "total_reward"):
                assert k in inf, k
            assert inf["n_successful_runs"] + inf["n_failed_runs"] >= 1
            assert inf["episode"]["l"] == inf["env_steps"] >= 1
            # auto-reset: the returned row is the first observation of a new episode (drone at rest)
            assert abs(obs[i, 0]) < 1e-6 and abs(obs[i, 1]) < 1e-6 and not np.array_equal(obs[i], inf["terminal_observation"])
    assert seen_done > 50
    st = env.episode_stats()
    assert st["episodes"] == seen_done and st["n_successful_runs"] + st["n_failed_runs"] >= seen_done
    assert env.seed()[:2] == [3, 4] and env.env_is_wrapped(object) == [False] * n
    assert env.get_attr("num_envs", [0, 1]) == [n, n]
    env.step_async(np.zeros((n, 2), np.float32))
    env.step_wait()
    with pytest.raises(RuntimeError):
        env.step_wait()
    env.close()
    env.close()


def test_full_info_mode_matches_episode_mode(pkg):
    a = pkg.Drone2dVecEnv(64, seed=5, info_mode="full", mode="test", scenario="corridor")
    b = pkg.Drone2dVecEnv(64, seed=5, info_mode="episode", mode="test", scenario="corridor")
    a.reset(); b.reset()
    rng = np.random.RandomState(1)
    for _ in range(40):
        act = rng.uniform(-1, 1, (64, 2)).astype(np.float32)
        oa, ra, da, ia = a.step(act)
        ob, rb, db, ib = b.step(act)
        assert np.array_equal(oa, ob) and np.array_equal(ra, rb) and np.array_equal(da, db)
        assert all(abs(ia[i]["reward"] - ra[i]) < 1e-4 * max(1, abs(ra[i])) for i in range(64))
        assert all(ia[i]["env_steps"] >= 1 for i in range(64))
    a.close(); b.close()


def test_single_env_gym_surface_and_parity(pkg):
    """Drone2dEnv(**kwargs): float64 observations like the reference (env:765), old gym 4-tuple, no auto-reset."""
    cfg = dict(pkg.rl_config.env_test_config)          # mode='test', scenario='large', render flags accepted
    env = pkg.Drone2dEnv(seed=11, **cfg)
    obs = env.reset()
    assert obs.shape == (27,) and obs.dtype == np.float64
    vec = env._vec
    ocfg = oracle_config_from(vec)
    state = blob_to_states(vec.get_state())[0]
    path, obst = oracle.load_scenario("large")
    rng = np.random.RandomState(2)
    for t in range(60):
        a = rng.uniform(-1, 1, 2).astype(np.float32)
        obs, rew, done, info = env.step(a)
        out = oracle.env_step(state, path, obst, ocfg, a[0], a[1])
        assert isinstance(rew, float) and isinstance(done, bool) and isinstance(info, dict)
        np.testing.assert_allclose(obs[:19], np.array(out.obs[:19]), rtol=0, atol=1e-9)
        np.testing.assert_allclose(obs[19:23], np.array(out.obs[19:23]), rtol=0, atol=1e-7)
        assert abs(rew - out.reward) < 1e-4 and done == bool(out.done)
        assert info["env_steps"] == t + 1
        if done:
            assert "APE" in info and "total_reward" in info
            break
    env.close()


def test_step_host_abi_matches_device_path(pkg):
    """d2d_step_host (host buffers, pageable and page-locked) == d2d_step on device tensors."""
    import torch
    n = 1000
    a = pkg.Drone2dVecEnv(n, seed=8, info_mode="none", mode="curriculum", scenario="stage_3")
    b = pkg.Drone2dVecEnv(n, seed=8, info_mode="none", mode="curriculum", scenario="stage_3")
    a.reset_device(); b.reset_device()
    rng = np.random.RandomState(3)
    obs_h = np.empty((n, 27), np.float32); rew_h = np.empty(n, np.float32); done_h = np.empty(n, np.uint8)
    obs_p = torch.empty((n, 27)).pin_memory()
    st = C.c_void_p(torch.cuda.current_stream(a.device).cuda_stream)
    for t in range(50):
        act = rng.uniform(-1, 1, (n, 2)).astype(np.float32)
        dst = obs_h if t % 2 == 0 else obs_p.numpy()
        rc = a._L.d2d_step_host(a._h, act.ctypes.data, dst.ctypes.data, rew_h.ctypes.data, done_h.ctypes.data, 1, st)
        assert rc == 0
        o, r, d = b.step_device(torch.from_numpy(act).to(b.device))
        assert np.array_equal(dst, o.cpu().numpy()) and np.array_equal(rew_h, r.cpu().numpy())
        assert np.array_equal(done_h, d.cpu().numpy())
    a.close(); b.close()


def test_error_paths(pkg):
    import torch
    env = pkg.Drone2dVecEnv(16, mode="test", scenario="corridor")
    env.reset_device()
    with pytest.raises(ValueError):
        env.step_device(torch.zeros((15, 2), device=env.device))
    with pytest.raises(ValueError):
        env.set_state(np.zeros((16, 35)))
    with pytest.raises(pkg.D2DError):
        env.get_tables()                               # per-env tables exist only in curriculum mode
    with pytest.raises(pkg.D2DError):
        env.set_stage(3)
    assert env._L.d2d_step(env._h, None, None, None, None, None, 1, None) == -1
    env.close()


def test_host_step_zero_copy_views_alternate(pkg):
    """copy_outputs=False: step() hands out views of two page-locked buffer sets that the kernel writes directly;
    the arrays of step t survive step t + 1 and are reused by step t + 2.  Values equal the copying path."""
    a = pkg.Drone2dVecEnv(512, seed=3, info_mode="none", copy_outputs=False, mode="curriculum", scenario="stage_3")
    b = pkg.Drone2dVecEnv(512, seed=3, info_mode="none", mode="curriculum", scenario="stage_3")
    a.reset(); b.reset()
    rng = np.random.RandomState(1)
    kept = []
    for t in range(6):
        act = rng.uniform(-1, 1, (512, 2)).astype(np.float32)
        oa, ra, da, ia = a.step(act)
        ob, rb, db, ib = b.step(act)
        assert np.array_equal(oa, ob) and np.array_equal(ra, rb) and np.array_equal(da, db)
        assert not oa.flags.owndata and ob.flags.owndata
        kept.append((oa, ob.copy()))
        if t >= 1:
            assert np.array_equal(kept[t - 1][0], kept[t - 1][1])          # the previous step's view is still intact
        if t >= 2:
            assert np.shares_memory(kept[t - 2][0], oa) and not np.shares_memory(kept[t - 1][0], oa)
    # terminal observations of finished envs arrive through the same mapped buffers
    for _ in range(400):
        act = rng.uniform(-1, 1, (512, 2)).astype(np.float32)
        oa, ra, da, ia = a.step(act)
        ob, rb, db, ib = b.step(act)
        for i in np.flatnonzero(da):
            assert np.array_equal(ia[i]["terminal_observation"], ib[i]["terminal_observation"])
            assert not np.array_equal(ia[i]["terminal_observation"], oa[i])   # obs already belongs to the next episode
    a.close(); b.close()


def test_flight_path_in_infos_when_render_path_is_set():
    """env:410-415,612: with render_path the finished episode's frame positions (x, screen_height - y), one per
    step, come back as info["flight_path"] (read at main.py:276).  Device ring -> gathered for finishing envs."""
    pkg = importlib.import_module(PKG)
    n, T = 96, 25
    env = pkg.Drone2dVecEnv(n, seed=5, mode="test", scenario="corridor", render_path=True, n_steps=T)
    assert env.path_ring is not None and env.path_ring.shape[0] == T
    env.reset()
    rng = np.random.RandomState(0)
    track = [[] for _ in range(n)]
    seen = 0
    for t in range(2 * T + 3):
        pre_steps = env.get_state()[:, 32].copy()
        obs, rew, dones, infos = env.step(rng.uniform(-1, 1, (n, 2)).astype(np.float32))
        st = env.get_state()
        for i in range(n):
            if dones[i]:
                fp = infos[i]["flight_path"]
                assert len(fp) == infos[i]["env_steps"] == int(pre_steps[i]) + 1
                ref = np.array(track[i], dtype=np.float64).reshape(-1, 2)
                got = np.array(fp)
                assert np.abs(got[:-1] - ref).max() < 1e-3 if len(ref) else True      # float32 ring vs float64 state
                # the last point is the terminal position: consistent with the terminal observation (obs[6:8])
                tobs = infos[i]["terminal_observation"]
                assert abs((tobs[6] + 1) * 650.0 - got[-1, 0]) < 0.05 and abs(1300.0 - (tobs[7] + 1) * 650.0 - got[-1, 1]) < 0.05
                track[i] = []
                seen += 1
            else:
                track[i].append((st[i, 0], 1300.0 - st[i, 1]))
    assert seen >= 2 * n
    env.close()
    # off by default (training configuration): no ring, info key present and empty
    env = pkg.Drone2dVecEnv(8, seed=5, mode="test", scenario="corridor", n_steps=3)
    env.reset()
    for _ in range(3):
        _, _, dones, infos = env.step(np.zeros((8, 2), np.float32))
    assert dones.all() and infos[0]["flight_path"] == [] and env.path_ring is None
    env.close()


def test_seed_rekeys_the_reset_stream():
    """SB3 VecEnv.seed(seed): the next reset draws from the new stream; the same seed gives the same episodes."""
    pkg = importlib.import_module(PKG)
    env = pkg.Drone2dVecEnv(64, seed=1, mode="curriculum", scenario="stage_5")
    a = env.reset().copy()
    assert env.seed(7) == [7 + i for i in range(64)]
    b = env.reset().copy()
    env.seed(1)
    c = env.reset().copy()
    assert not np.array_equal(a, b) and np.array_equal(a, c)
    obs, rew, dones, infos = env.step(np.zeros((64, 2), np.float32))       # the rebuilt handle steps and fills infos
    assert obs.shape == (64, 27) and np.isfinite(obs).all()
    env.close()


def test_flight_path_ring_shorter_than_the_episode_keeps_the_last_positions():
    """A ring of K < T positions: the info carries the last K positions of the episode, oldest first."""
    pkg = importlib.import_module(PKG)
    n, T, K = 32, 40, 16
    env = pkg.Drone2dVecEnv(n, seed=9, mode="test", scenario="corridor", n_steps=T, info_mode="episode")
    env.enable_flight_paths(ring=K)
    env.reset()
    track = [[] for _ in range(n)]
    checked = 0
    for t in range(T + 1):
        obs, rew, dones, infos = env.step(np.zeros((n, 2), np.float32))
        st = env.get_state()
        for i in range(n):
            if dones[i]:
                fp = np.array(infos[i]["flight_path"])
                steps = infos[i]["env_steps"]
                want = np.array(track[i], dtype=np.float64).reshape(-1, 2)[-(min(K, steps) - 1):] if min(K, steps) > 1 else np.zeros((0, 2))
                assert len(fp) == min(K, steps)
                assert np.abs(fp[:-1] - want).max() < 1e-3          # all but the terminal position (state was reset)
                track[i] = []
                checked += 1
            else:
                track[i].append((st[i, 0], 1300.0 - st[i, 1]))
    assert checked >= n
    env.close()
