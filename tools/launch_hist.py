"""Distribution of single-launch durations of the step kernel in steady state (CUDA events around every launch):
how much of the mean step time comes from the rare launches that contain an in-lane reset?"""
import importlib, json, os, sys
import numpy as np, torch
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, REPO)
pkg = importlib.import_module("drone_2d_custom_gym_env_for_reinforcement_learning_b200")
E = 65536
env = pkg.Drone2dVecEnv(E, seed=0, info_mode="none", mode="curriculum", scenario="stage_5")
env.reset_device()
g = torch.Generator(device=env.device); g.manual_seed(1)
ring = (torch.rand((64, E, 2), device=env.device, generator=g) * 2 - 1).contiguous()
for k in range(400): env.step_device(ring[k % 64], terminal_obs=False)
s0 = env.episode_stats()
N = 2000
evs = [torch.cuda.Event(enable_timing=True) for _ in range(N + 1)]
torch.cuda.synchronize()
evs[0].record()
for k in range(N):
    env.step_device(ring[k % 64], terminal_obs=False)
    evs[k + 1].record()
torch.cuda.synchronize()
dt = np.array([evs[k].elapsed_time(evs[k + 1]) * 1000.0 for k in range(N)])
s1 = env.episode_stats()
med = float(np.median(dt))
slow = dt > 1.25 * med
print(json.dumps({"launches": N, "mean_us": round(float(dt.mean()), 2), "median_us": round(med, 2),
                  "p10_us": round(float(np.percentile(dt, 10)), 2), "p90_us": round(float(np.percentile(dt, 90)), 2),
                  "p99_us": round(float(np.percentile(dt, 99)), 2), "max_us": round(float(dt.max()), 2),
                  "slow_launches(>1.25 median)": int(slow.sum()), "mean_excess_us_from_slow": round(float((dt[slow] - med).sum() / N), 3),
                  "episodes": int(s1["episodes"] - s0["episodes"]), "inlane_resets": int(s1["inlane_resets"] - s0["inlane_resets"])}))
