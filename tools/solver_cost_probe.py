"""How much of the step's critical path is the rigid-body solver?  Times the stage_5 / 65 536-env step with 1, 5, 10
(pymunk's default) and 20 Gauss-Seidel iterations: the slope is the solver's share of the launch time, i.e. the most
that moving it off the critical path could buy.  usage: python tools/solver_cost_probe.py"""
import importlib
import json
import os
import sys

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
pkg = importlib.import_module("drone_2d_custom_gym_env_for_reinforcement_learning_b200")
E, R = 65536, 256
g = torch.Generator(device="cuda")
g.manual_seed(1)
ring = (torch.rand((R, E, 2), device="cuda", generator=g) * 2 - 1).contiguous()
for it in (10, 1, 5, 20):
    env = pkg.Drone2dVecEnv(E, seed=0, info_mode="none", mode="curriculum", scenario="stage_5", solver_iterations=it)
    env.reset_device()
    for k in range(400):
        env.step_device(ring[k % R], terminal_obs=False)
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for k in range(1000):
        env.step_device(ring[k % R], terminal_obs=False)
    ev1.record()
    torch.cuda.synchronize()
    us = ev0.elapsed_time(ev1)
    st = env.episode_stats()
    env.role_cycles()
    for k in range(64):
        env.step_device(ring[k % R], terminal_obs=False)
    rc = env.role_cycles()
    print(json.dumps({"solver_iterations": it, "us_per_step": round(us, 2), "mean_episode_len": round(st["env_steps"] / max(st["episodes"], 1), 1),
                      "kcyc(max,mean)": {k: (v["max"] // 1000, v["mean"] // 1000) for k, v in rc.items() if v["warps"]}}), flush=True)
    env.close()
