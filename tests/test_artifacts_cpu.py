"""The artefact writer reproduces the reference's test-mode files (main.py:289-327)."""
import importlib
import json
import os

import numpy as np

PKG = "drone_2d_custom_gym_env_for_reinforcement_learning_b200"
ev = importlib.import_module(PKG + ".evaluate")


def test_results_file_layout_and_format(tmp_path):
    n = 100
    rng = np.random.RandomState(0)
    succ = (rng.rand(n) < 0.88).astype(float)
    res = {"n_successful_runs": succ, "n_failed_runs": 1 - succ, "n_collisions": 1 - succ,
           "APE": rng.uniform(5, 300, n), "env_steps": rng.randint(100, 1100, n).astype(float),
           "total_reward": rng.uniform(-100, 1500, n)}
    paths = [[[1.0, 2.0], [3.0, 4.0]] for _ in range(n)]
    agent = "ppo_agents/PFCA_see_3_obs_19_90.zip"
    d = ev.write_test_artifacts(res, paths, agent, "corridor", out_root=str(tmp_path / "Tests"))
    assert d.endswith(os.path.join("Tests", "agent_19", "test_0", "corridor"))      # main.py:291-302
    assert np.load(os.path.join(d, "collisions.npy")).dtype == np.int32
    assert np.load(os.path.join(d, "time_spent.npy")).dtype == np.int32
    assert np.load(os.path.join(d, "apes.npy")).dtype == np.float64 and np.load(os.path.join(d, "rewards.npy")).shape == (n,)
    assert json.load(open(os.path.join(d, "flight_paths")))[0] == [[1.0, 2.0], [3.0, 4.0]]
    lines = open(os.path.join(d, "corridor_19_results.txt")).read().splitlines()
    assert [l.split(":")[0] for l in lines] == ["Successes", "Fails", "Collisions", "Success rate", "Collision rate",
                                                "Average APE", "Average flight time", "Agent path"]
    assert lines[0] == "Successes: %d" % succ.sum() and lines[7] == "Agent path: " + agent
    assert lines[3] == "Success rate: " + str(succ.sum() / n)
    # a second scenario goes into the same test_0, the same scenario again opens test_1
    d2 = ev.write_test_artifacts(res, paths, agent, "large", out_root=str(tmp_path / "Tests"))
    d3 = ev.write_test_artifacts(res, paths, agent, "corridor", out_root=str(tmp_path / "Tests"))
    assert os.path.join("test_0", "large") in d2 and os.path.join("test_1", "corridor") in d3
