"""Per-launch critical-path probe: which role's slowest warp bounds each step kernel?"""
import importlib, os, sys
import numpy as np, torch
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, REPO)
pkg = importlib.import_module("drone_2d_custom_gym_env_for_reinforcement_learning_b200")
E = 65536
env = pkg.Drone2dVecEnv(E, seed=0, info_mode="none", mode="curriculum", scenario="stage_5")
env.reset_device()
g = torch.Generator(device=env.device); g.manual_seed(1)
ring = (torch.rand((16, E, 2), device=env.device, generator=g) * 2 - 1).contiguous()
for k in range(300): env.step_device(ring[k % 16], terminal_obs=False)
env.role_cycles()
rows = []
for k in range(60):
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(); env.step_device(ring[k % 16], terminal_obs=False); ev1.record(); torch.cuda.synchronize()
    rc = env.role_cycles()
    rows.append((ev0.elapsed_time(ev1) * 1000, rc["main_total"]["max"] / 1e3, rc["main_total"]["mean"] / 1e3,
                 rc["refill_sample"]["max"] / 1e3, rc["refill_observe"]["max"] / 1e3, rc["observe"]["max"] / 1e3))
a = np.array(rows)
print("per-launch medians: kernel_us %.1f | main max %.0fk mean %.0fk | sample max %.0fk | observe-role max %.0fk | main observe-phase max %.0fk"
      % tuple(np.median(a, 0)))
print("fraction of launches where the limiter is: main %.2f sample %.2f observe-role %.2f" % tuple(
    np.mean(np.argmax(a[:, [1, 3, 4]], 1) == j) for j in range(3)))
