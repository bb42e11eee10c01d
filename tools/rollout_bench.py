"""(f3) rate of the hand-written PPO rollout collection (RolloutCollector -> d2d_rollout_collect): T x (tcgen05 policy
kernel, step kernel) + one critic launch over the T+1 stored observations + GAE, nothing else on the stream."""
import importlib
import json
import os
import sys

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
PKG = "drone_2d_custom_gym_env_for_reinforcement_learning_b200"
pkg = importlib.import_module(PKG)
ro = importlib.import_module(PKG + ".rollout")
G = os.path.join(REPO, "tests", "golden")
a, v = np.load(os.path.join(G, "agent_19_policy.npz")), np.load(os.path.join(G, "agent_19_value.npz"))
E, T = int(sys.argv[1]) if len(sys.argv) > 1 else 65536, 64
for mode, scen in (("curriculum", "stage_5"), ("test", "large")):
    env = pkg.Drone2dVecEnv(E, seed=0, info_mode="none", mode=mode, scenario=scen)
    pol = ro.MlpActorCritic({k.replace("__", "."): torch.from_numpy(a[k]) for k in a.files},
                            {k.replace("__", "."): torch.from_numpy(v[k]) for k in v.files}, env.device)
    col = ro.RolloutCollector(env, pol, n_steps=T, seed=1)
    for _ in range(5):
        col.collect()
    torch.cuda.synchronize()
    l0 = env.launch_count
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    R = 6
    for _ in range(R):
        col.collect()
    ev1.record()
    torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1)
    st = env.episode_stats()
    print(json.dumps({"scenario": scen, "envs": E, "T": T, "us_per_env_step_batch": round(ms * 1000 / (R * T), 2),
                      "env_steps_per_s": round(E * R * T / ms * 1000), "launches_per_rollout": (env.launch_count - l0) / R,
                      "episodes": st["episodes"], "mean_adv": float(col.advantages.mean()), "mean_value": float(col.values.mean())}), flush=True)
    env.close()
