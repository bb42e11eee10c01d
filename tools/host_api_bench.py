"""SB3-facing host API cost: Drone2dVecEnv.step(np.ndarray) at full batch."""
import importlib, os, sys, time
import numpy as np
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, REPO)
pkg = importlib.import_module("drone_2d_custom_gym_env_for_reinforcement_learning_b200")
for mode, copy in (("episode", True), ("none", True), ("none", False)):
    env = pkg.Drone2dVecEnv(65536, seed=0, info_mode=mode, copy_outputs=copy, mode="curriculum", scenario="stage_5")
    env.reset()
    rng = np.random.RandomState(0)
    acts = rng.uniform(-1, 1, (8, 65536, 2)).astype(np.float32)
    for k in range(20): env.step(acts[k % 8])
    t0 = time.perf_counter()
    K = 100
    for k in range(K): env.step(acts[k % 8])
    dt = (time.perf_counter() - t0) / K
    print("info_mode=%s copy_outputs=%s: %.2f ms per VecEnv.step -> %.1f M env-steps/s" % (mode, copy, dt * 1e3, 65536 / dt / 1e6))
    env.close()
