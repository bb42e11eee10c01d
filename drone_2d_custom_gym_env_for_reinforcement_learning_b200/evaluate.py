"""Test-mode evaluation with the reference's on-disk artefact layout (main.py:242-327).

`run_test` plays `run_n_times` episodes of a scenario with a shipped PPO agent (one
environment per episode, all in one batch on the GPU) and `write_test_artifacts`
writes exactly what the reference's `mode="test"` block writes:

    <out_root>/agent_<nr>/test_<k>/<scenario>/flight_paths          JSON, per episode [[x, H - y], ...] (env:984-986)
    <out_root>/agent_<nr>/test_<k>/<scenario>/collisions.npy        int32   (run_n_times,)
    <out_root>/agent_<nr>/test_<k>/<scenario>/rewards.npy           float64
    <out_root>/agent_<nr>/test_<k>/<scenario>/apes.npy              float64
    <out_root>/agent_<nr>/test_<k>/<scenario>/time_spent.npy        int32
    <out_root>/agent_<nr>/test_<k>/<scenario>/<scenario>_<nr>_results.txt

so results can be diffed file-for-file against the reference's `Tests/**` tree.  GIFs
and PNG plots (pygame / matplotlib) are out of scope.
"""
import json
import os

import numpy as np


def agent_number(agent_path):
    """main.py:291: `agent_path.split('_')[-2].split('.')[0]`."""
    return agent_path.split('_')[-2].split('.')[0]


def _test_dir(out_root, agent, scenario):
    """main.py:297-302: reuse the last test_<k> unless it already holds this scenario."""
    base = os.path.join(out_root, agent)
    os.makedirs(base, exist_ok=True)
    existing = sorted(d for d in os.listdir(base) if d.startswith("test_"))
    if existing:
        last = "test_%d" % (len(existing) - 1)
        if os.path.isdir(os.path.join(base, last)) and scenario not in os.listdir(os.path.join(base, last)):
            return os.path.join(base, last)
    return os.path.join(base, "test_%d" % len(existing))


def results_text(successes, fails, collision_sum, apes, time_spent, agent_path):
    """The eight lines of main.py:318-326 (same labels, same `str()` formatting)."""
    return ('Successes: ' + str(successes) + '\n' + 'Fails: ' + str(fails) + '\n' + 'Collisions: ' + str(collision_sum) + '\n' +
            'Success rate: ' + str(successes / (successes + fails)) + '\n' +
            'Collision rate: ' + str(collision_sum / (successes + fails)) + '\n' +
            'Average APE: ' + str(np.mean(apes)) + '\n' + 'Average flight time: ' + str(np.mean(time_spent)) + '\n' +
            'Agent path: ' + agent_path + '\n')


def write_test_artifacts(res, flight_paths, agent_path, scenario, out_root="Tests"):
    """`res`: dict of per-episode arrays (policy.rollout_one_episode_each); returns the directory written."""
    nr = agent_number(agent_path)
    tdir = os.path.join(_test_dir(out_root, "agent_" + nr, scenario), scenario)
    os.makedirs(tdir, exist_ok=True)
    with open(os.path.join(tdir, "flight_paths"), "w") as f:
        json.dump(flight_paths, f)
    collisions = np.asarray(res["n_collisions"]).astype(np.int32)
    time_spent = np.asarray(res["env_steps"]).astype(np.int32)
    apes = np.asarray(res["APE"], dtype=np.float64)
    rewards = np.asarray(res["total_reward"], dtype=np.float64)
    np.save(os.path.join(tdir, "collisions.npy"), collisions)
    np.save(os.path.join(tdir, "rewards.npy"), rewards)
    np.save(os.path.join(tdir, "apes.npy"), apes)
    np.save(os.path.join(tdir, "time_spent.npy"), time_spent)
    successes = int(np.sum(np.asarray(res["n_successful_runs"]) == 1))
    fails = int(np.sum(np.asarray(res["n_failed_runs"]) == 1))
    with open(os.path.join(tdir, "%s_%s_results.txt" % (scenario, nr)), "w") as f:
        f.write(results_text(successes, fails, int(np.sum(collisions)), apes, time_spent, agent_path))
    return tdir


def run_test(agent_path, scenario, run_n_times=100, out_root="Tests", device=None, seed=0, deterministic=False,
             weights_path=None, **env_kwargs):
    """The reference's test mode for one scenario.  `weights_path` lets the actor come from another file
    (e.g. the .npz fixture) while `agent_path` names the agent in the artefacts."""
    import torch
    from .policy import MlpActor
    from .vec_env import Drone2dVecEnv
    mode = "curriculum" if str(scenario).startswith("stage_") else "test"
    kw = dict(env_kwargs)
    kw.setdefault("mode", mode)
    kw["scenario"] = scenario
    env = Drone2dVecEnv(run_n_times, device=device, seed=seed, **kw)
    actor = MlpActor.load(weights_path or agent_path, env.device)
    g = torch.Generator(device=env.device)
    g.manual_seed(seed)
    W, H = env.cfg.screen_x, env.cfg.screen_y
    obs = env.reset_device()
    n = run_n_times
    alive = torch.ones(n, dtype=torch.bool, device=env.device)
    out = torch.zeros((n, 6), dtype=torch.float64, device=env.device)
    track = []                                   # per step: (alive mask, x, H - y) of the frame (env:984-986)
    for _ in range(env.cfg.n_steps):
        act = actor.predict(obs, deterministic, g)
        obs, _, done = env.step_device(act, auto_reset=False, terminal_obs=False)
        x = (obs[:, 6].double() + 1.0) * W / 2.0
        y = (obs[:, 7].double() + 1.0) * H / 2.0
        track.append((alive.clone(), x, H - y))
        newly = alive & (done != 0)
        if newly.any():
            info = env.info_rows(newly)
            bits = info[:, 11].long()
            out[newly] = torch.stack([((bits & 2) != 0).double(), ((bits & 13) != 0).double(),
                                      (((bits & 1) != 0) & ((bits & 14) == 0)).double(), info[:, 9], info[:, 7],
                                      info[:, 10]], 1)
            alive &= ~newly
        if not bool(alive.any()):
            break
    env.close()
    r = out.cpu().numpy()
    res = {"n_successful_runs": r[:, 0], "n_failed_runs": r[:, 1], "n_collisions": r[:, 2], "APE": r[:, 3],
           "env_steps": r[:, 4], "total_reward": r[:, 5]}
    masks = torch.stack([t[0] for t in track]).cpu().numpy()
    xs = torch.stack([t[1] for t in track]).cpu().numpy()
    ys = torch.stack([t[2] for t in track]).cpu().numpy()
    flight_paths = [[[float(xs[t, i]), float(ys[t, i])] for t in range(len(track)) if masks[t, i]] for i in range(n)]
    tdir = write_test_artifacts(res, flight_paths, agent_path, scenario, out_root)
    return tdir, res
