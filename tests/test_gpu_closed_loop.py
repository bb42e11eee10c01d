"""Closed loop on the GPU: the shipped agent_19 actor (torch, on device) drives the batched
environment; outcomes must reproduce the reference's evaluation artefacts and agree with the
CPU oracle's closed loop (BASELINE config 4 and SURVEY Appendix B.3)."""
import importlib
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN
from test_closed_loop_cpu import FIX, check_against_fixture

pytestmark = pytest.mark.gpu
PKG = "drone_2d_custom_gym_env_for_reinforcement_learning_b200"
SETTINGS = [("test", s) for s in ("corridor", "S_corridor", "parallel", "S_parallel", "perpendicular", "large", "impossible")] + \
           [("curriculum", "stage_%d" % k) for k in range(1, 6)]


@pytest.mark.parametrize("mode,scenario", SETTINGS)
def test_agent19_outcomes_match_reference_fixtures(mode, scenario):
    import torch
    pkg = importlib.import_module(PKG)
    policy = importlib.import_module(PKG + ".policy")
    n = 4096
    env = pkg.Drone2dVecEnv(n, seed=5, mode=mode, scenario=scenario)
    actor = policy.MlpActor.load(os.path.join(GOLDEN, "agent_19_policy.npz"), env.device)
    res = policy.rollout_one_episode_each(env, actor, seed=1)
    env.close()
    assert res["finished"].all()          # n_steps = 1100 ends every episode
    check_against_fixture(res, FIX[scenario], n, scenario)


def test_gpu_and_oracle_closed_loops_agree():
    """Same spec, two implementations, independent random streams: rates agree within sampling noise."""
    from oracle import closed_loop
    pkg = importlib.import_module(PKG)
    policy = importlib.import_module(PKG + ".policy")
    n_gpu, n_cpu = 8192, 400
    env = pkg.Drone2dVecEnv(n_gpu, seed=9, mode="test", scenario="corridor")
    actor = policy.MlpActor.load(os.path.join(GOLDEN, "agent_19_policy.npz"), env.device)
    g = policy.rollout_one_episode_each(env, actor, seed=2)
    env.close()
    c = closed_loop.evaluate("test", "corridor", n_cpu, seed=4)
    for key in ("n_successful_runs", "n_collisions"):
        pg, pc = g[key].mean(), c[key].mean()
        sigma = np.sqrt(max(pc * (1 - pc), 0.02) / n_cpu + max(pg * (1 - pg), 0.02) / n_gpu)
        assert abs(pg - pc) <= 4 * sigma, (key, pg, pc, sigma)
    assert abs(g["env_steps"].mean() - c["env_steps"].mean()) <= 0.08 * c["env_steps"].mean()


def test_run_test_writes_reference_artifacts(tmp_path):
    """The reference's test mode end to end: 100 episodes, files in the Tests/agent_<nr>/test_<k>/<scenario>/ layout."""
    ev = importlib.import_module(PKG + ".evaluate")
    d, res = ev.run_test("ppo_agents/PFCA_see_3_obs_19_90.zip", "corridor", run_n_times=100, out_root=str(tmp_path / "Tests"),
                         seed=3, weights_path=os.path.join(GOLDEN, "agent_19_policy.npz"))
    assert d.endswith(os.path.join("agent_19", "test_0", "corridor"))
    t = np.load(os.path.join(d, "time_spent.npy"))
    fp = json.load(open(os.path.join(d, "flight_paths")))
    assert t.shape == (100,) and len(fp) == 100
    assert all(len(fp[i]) == t[i] for i in range(100))                 # one [x, H-y] per step until done
    assert all(0 <= p[0] <= 1400 for p in fp[0])
    txt = open(os.path.join(d, "corridor_19_results.txt")).read()
    assert "Successes: %d" % int(res["n_successful_runs"].sum()) in txt
    check_against_fixture(res, FIX["corridor"], 100, "corridor/run_test")


def test_rollout_collector_and_gae_kernel():
    """(f3) device-resident rollout: GAE kernel == SB3's recursion, buffers consistent, curriculum hook works."""
    import torch
    pkg = importlib.import_module(PKG)
    ro = importlib.import_module(PKG + ".rollout")
    env = pkg.Drone2dVecEnv(2048, seed=2, info_mode="none", mode="curriculum", scenario="stage_1")
    pol = ro.MlpActorCritic.random_init(env.device, seed=1)
    col = ro.RolloutCollector(env, pol, n_steps=96, seed=3).collect()
    assert col.obs.shape == (96, 2048, 27) and col.num_timesteps == 96 * 2048
    assert int(col.episode_starts[0].sum()) == 2048 and int(col.episode_starts[1:].sum()) > 0   # random policy: episodes end
    a_ref, r_ref = ro.gae_reference(col.rewards.cpu().numpy(), col.values.cpu().numpy(), col.episode_starts.cpu().numpy(),
                                    pol.value(col._last_obs).cpu().numpy(), col._last_starts.cpu().numpy())
    np.testing.assert_allclose(col.advantages.cpu().numpy(), a_ref, rtol=2e-4, atol=2e-3)
    np.testing.assert_allclose(col.returns.cpu().numpy(), r_ref, rtol=2e-4, atol=2e-3)
    # curriculum scheduler: switch to stage_5 by training progress; new episodes get obstacles
    env.set_stage(sim_num=2_500_000)
    assert env.cfg.stage == 5
    for _ in range(300):
        env.step_device(None)
    assert env.get_tables()["n_obs"].max() >= 2
    env.close()


def test_fused_policy_kernel_matches_torch_actor():
    """d2d_policy_act (hand-written float32 MLP kernel) vs the torch float32 actor on real observations."""
    import torch
    pkg = importlib.import_module(PKG)
    policy = importlib.import_module(PKG + ".policy")
    n = 5000                                                   # partial last block
    env = pkg.Drone2dVecEnv(n, seed=4, info_mode="none", mode="test", scenario="S_corridor")
    obs = env.reset_device()
    w = os.path.join(GOLDEN, "agent_19_policy.npz")
    ref, fused = policy.MlpActor.load(w, env.device), policy.FusedActor.load(w, env)
    for _ in range(30):
        a_ref = ref.predict(obs, deterministic=True)
        a_fused = fused.predict(obs, deterministic=True)
        assert float((a_ref - a_fused).abs().max()) <= 2e-5
        obs, _, _ = env.step_device(a_fused)
    # stochastic: noise = std * N(0,1) from the Philox stream, different every call, clipped to the Box
    mu = ref.mean(obs)
    a1 = fused.predict(obs).clone()
    a2 = fused.predict(obs).clone()
    assert float((a1 - a2).abs().max()) > 0.1 and float(a1.abs().max()) <= 1.0
    # where mean +- std lies inside the Box, clipping cannot affect the event |z| < 1: P = 68.3 %
    safe = ((mu.abs() + ref.std) < 1.0)
    z = (a1 - mu) / ref.std
    frac = float(((z.abs() < 1.0) & safe).sum()) / float(safe.sum())
    assert int(safe.sum()) > 1000 and abs(frac - 0.6827) < 0.03
    assert abs(float((a1 > mu)[safe].float().mean()) - 0.5) < 0.03
    env.close()


def test_closed_loop_with_fused_policy_kernel():
    pkg = importlib.import_module(PKG)
    policy = importlib.import_module(PKG + ".policy")
    env = pkg.Drone2dVecEnv(4096, seed=6, mode="test", scenario="large")
    actor = policy.FusedActor.load(os.path.join(GOLDEN, "agent_19_policy.npz"), env)
    res = policy.rollout_one_episode_each(env, actor, seed=1)
    env.close()
    check_against_fixture(res, FIX["large"], 4096, "large/fused")
