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


def _agent19_actor_critic(ro, device):
    import torch
    a, v = np.load(os.path.join(GOLDEN, "agent_19_policy.npz")), np.load(os.path.join(GOLDEN, "agent_19_value.npz"))
    actor = {k.replace("__", "."): torch.from_numpy(a[k]) for k in a.files}
    critic = {k.replace("__", "."): torch.from_numpy(v[k]) for k in v.files}
    return ro.MlpActorCritic(actor, critic, device)


def test_rollout_collector_and_gae_kernel():
    """(f3) hand-written rollout collection (d2d_rollout_collect): the kernels write observations, unclipped actions,
    log-probabilities, rewards, episode starts and values into the [T, N] tensors; compared with the torch float32
    statement of the same SB3 policy (MlpActorCritic) and with SB3's GAE recursion.  Curriculum hook works."""
    import math
    import torch
    pkg = importlib.import_module(PKG)
    ro = importlib.import_module(PKG + ".rollout")
    T, n = 96, 2048
    env = pkg.Drone2dVecEnv(n, seed=2, info_mode="none", mode="curriculum", scenario="stage_3")
    pol = _agent19_actor_critic(ro, env.device)
    col = ro.RolloutCollector(env, pol, n_steps=T, seed=3)
    l0 = env.launch_count
    col.collect()
    assert env.launch_count - l0 == 2 * T + 1              # T x (policy, step) + the critic; GAE is handle-free
    assert col.obs.shape == (T, n, 27) and col.num_timesteps == T * n
    assert int(col.episode_starts[0].sum()) == n and int(col.episode_starts[1:].sum()) > 0
    obs, raw = col.obs_buf, col.actions
    # actor: the stored sample is mean + std * eps with eps ~ N(0, 1); its log-probability under the torch policy
    mu = pol.mean(obs[:T].reshape(-1, 27)).reshape(T, n, 2)
    eps = (raw - mu) / pol.std
    assert abs(float(eps.mean())) < 0.01 and abs(float(eps.std()) - 1.0) < 0.01
    logp_ref = (-0.5 * eps * eps - pol.log_std - 0.5 * math.log(2 * math.pi)).sum(-1)
    assert float((col.log_probs - logp_ref).abs().max()) < 2e-3      # eps carries the 3e-6 error of the mean divided by std
    # the env was stepped with the clipped sample: a second handle replays the same actions from the same seed
    env2 = pkg.Drone2dVecEnv(n, seed=2, info_mode="none", mode="curriculum", scenario="stage_3")
    o2 = env2.reset_device()
    assert torch.equal(o2, obs[0])
    for t in range(8):
        o2, r2, d2 = env2.step_device(raw[t].clamp(-1.0, 1.0), terminal_obs=False)
        assert torch.equal(o2, obs[t + 1]) and torch.equal(r2, col.rewards[t]) and torch.equal(d2, col.dones[t + 1])
    env2.close()
    # critic: one launch over all T + 1 observations vs the torch value tower
    v_ref = pol.value(obs.reshape(-1, 27)).reshape(T + 1, n)
    assert float((col.values_buf - v_ref).abs().max()) <= 2e-5 * max(1.0, float(v_ref.abs().max()))
    a_ref, r_ref = ro.gae_reference(col.rewards.cpu().numpy(), col.values.cpu().numpy(), col.episode_starts.cpu().numpy(),
                                    col.last_values.cpu().numpy(), col.last_dones.cpu().numpy())
    np.testing.assert_allclose(col.advantages.cpu().numpy(), a_ref, rtol=2e-4, atol=2e-3)
    np.testing.assert_allclose(col.returns.cpu().numpy(), r_ref, rtol=2e-4, atol=2e-3)
    # a second rollout continues where the first one stopped
    last = col.last_obs.clone()
    col.collect()
    assert torch.equal(col.obs_buf[0], last) and col.num_timesteps == 2 * T * n
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
