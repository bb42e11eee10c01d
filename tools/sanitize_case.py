"""Tiny workload for compute-sanitizer: both kernel templates, resets, fallback path, taps."""
import importlib, os, sys
import torch
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, REPO)
pkg = importlib.import_module("drone_2d_custom_gym_env_for_reinforcement_learning_b200")
for kw in (dict(mode="curriculum", scenario="stage_5", n_steps=3), dict(mode="curriculum", scenario="stage_5"),
           dict(mode="test", scenario="S_corridor"), dict(mode="test", scenario="large", n_steps=2)):
    env = pkg.Drone2dVecEnv(200, seed=1, **kw)       # 200: partial last block
    env.reset_device()
    for k in range(12):
        env.step_device(None)
    env.closest_u(torch.rand(200, 2).numpy() * 1300)
    st = env.get_state(); env.set_state(st)
    if kw["mode"] == "curriculum":
        tb = env.get_tables(); env.set_tables(tb["wps"], tb["us"], tb["xparams"], tb["yparams"], tb["n_obs"], tb["obstacles"])
    print(kw, env.episode_stats()["episodes"])
    env.close()
print("sanitize case done")
