"""Kernel-tuning helper: time the step kernel early (no episode has ended yet) and in steady state.
usage: D2D_LIB=/path/to/variant.so python tools/quick_bench.py [envs] [scenario]"""
import importlib
import json
import os
import sys

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
pkg = importlib.import_module("drone_2d_custom_gym_env_for_reinforcement_learning_b200")

E = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
scen = sys.argv[2] if len(sys.argv) > 2 else "stage_5"
mode = "curriculum" if scen.startswith("stage") else "test"
env = pkg.Drone2dVecEnv(E, seed=0, info_mode="none", mode=mode, scenario=scen)
env.reset_device()
g = torch.Generator(device=env.device)
g.manual_seed(1)
R = 256
ring = (torch.rand((R, E, 2), device=env.device, generator=g) * 2 - 1).contiguous()


def timed(n, k0):
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    ev0.record()
    for k in range(n):
        env.step_device(ring[(k0 + k) % R], terminal_obs=False)
    ev1.record()
    torch.cuda.synchronize()
    return ev0.elapsed_time(ev1) / n * 1000.0


for k in range(3):
    env.step_device(ring[k], terminal_obs=False)
env.role_cycles()
early = timed(10, 3)
rc_early = env.role_cycles()
for k in range(400):
    env.step_device(ring[k % R], terminal_obs=False)
env.role_cycles()
steady = timed(400, 0)
rc_steady = env.role_cycles()
st = env.episode_stats()
print(json.dumps({"lib": os.path.basename(os.environ.get("D2D_LIB", "default")), "envs": E, "scenario": scen,
                  "early_us": round(early, 1), "steady_us": round(steady, 1),
                  "steady_Msteps_s": round(E / steady, 1), "mean_ep_len": round(st["env_steps"] / max(st["episodes"], 1), 1),
                  "episodes": int(st["episodes"]), "inlane_resets": int(st["inlane_resets"]),
                  "kcyc_early(max,mean)": {k: (v["max"] // 1000, v["mean"] // 1000) for k, v in rc_early.items()},
                  "kcyc_steady(max,mean)": {k: (v["max"] // 1000, v["mean"] // 1000) for k, v in rc_steady.items()}}))
