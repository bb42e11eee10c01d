"""Per-phase warp clock spans of the step kernel (d2d_debug_role_cycles): max and mean over warps,
median over launches.  Shows how much of the launch time is tail (slowest warp) and in which phase."""
import importlib, os, sys
import numpy as np, torch
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, REPO)
pkg = importlib.import_module("drone_2d_custom_gym_env_for_reinforcement_learning_b200")
E = 65536
kw = dict(mode="curriculum", scenario=sys.argv[1] if len(sys.argv) > 1 else "stage_5")
if not kw["scenario"].startswith("stage_"):
    kw["mode"] = "test"
env = pkg.Drone2dVecEnv(E, seed=0, info_mode="none", **kw)
env.reset_device()
g = torch.Generator(device=env.device); g.manual_seed(1)
ring = (torch.rand((16, E, 2), device=env.device, generator=g) * 2 - 1).contiguous()
for k in range(300): env.step_device(ring[k % 16], terminal_obs=False)
env.role_cycles()
names = ("main_total", "physics", "obstacle_pass", "observe", "reward_swap", "refill_sample", "refill_observe")
rows, us = [], []
for k in range(60):
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(); env.step_device(ring[k % 16], terminal_obs=False); ev1.record(); torch.cuda.synchronize()
    rc = env.role_cycles()
    us.append(ev0.elapsed_time(ev1) * 1000)
    rows.append([v for nm in names for v in (rc[nm]["max"] / 1e3, rc[nm]["mean"] / 1e3)])
a = np.median(np.array(rows), 0)
print("%s: kernel %.1f us (%.0fk cycles at 1.965 GHz)" % (kw["scenario"], np.median(us), np.median(us) * 1.965))
for j, nm in enumerate(names):
    print("  %-15s max %6.1fk  mean %6.1fk" % (nm, a[2 * j], a[2 * j + 1]))
