"""First-contact script for the GPU box: prints parity statistics for a few configurations."""
import importlib
import json
import os
import sys
import time

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "tests"))
pkg = importlib.import_module("drone_2d_custom_gym_env_for_reinforcement_learning_b200")
from parity_util import run_parity  # noqa: E402

for cfg in [dict(mode="test", scenario="corridor", n_envs=256, n_rounds=64),
            dict(mode="test", scenario="S_corridor", n_envs=256, n_rounds=64),
            dict(mode="test", scenario="large", n_envs=256, n_rounds=64),
            dict(mode="curriculum", scenario="stage_1", n_envs=256, n_rounds=48),
            dict(mode="curriculum", scenario="stage_5", n_envs=256, n_rounds=48)]:
    t = time.time()
    res = run_parity(pkg, seed=1, **cfg)
    print(json.dumps({"cfg": cfg, "res": res, "s": round(time.time() - t, 1)}), flush=True)
