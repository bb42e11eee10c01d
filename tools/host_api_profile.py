"""Where does Drone2dVecEnv.step (the SB3-facing call) spend its time at 65 536 envs?  cProfile of 40 steps."""
import cProfile
import importlib
import os
import pstats
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
pkg = importlib.import_module("drone_2d_custom_gym_env_for_reinforcement_learning_b200")
E = 65536
env = pkg.Drone2dVecEnv(E, seed=0, info_mode="episode", mode="curriculum", scenario="stage_5")
env.reset()
for _ in range(300):
    env.step_device(None, terminal_obs=False)
acts = np.random.RandomState(0).uniform(-1, 1, (4, E, 2)).astype(np.float32)
for k in range(5):
    env.step(acts[k % 4])
pr = cProfile.Profile()
pr.enable()
for k in range(40):
    env.step(acts[k % 4])
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(14)
