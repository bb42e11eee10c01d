"""Statistical pin of the oracle's restated rigid-body step: the reference's own
evaluation artefacts (Tests/<setting>/test_k/results.txt, 100 pymunk episodes
each with the shipped agent_19 policy; committed as tests/golden/closed_loop.json)
must be reproduced by the oracle closed loop within sampling noise."""
import json
import math
import os

import numpy as np
import pytest

from oracle import closed_loop
from conftest import GOLDEN

FIX = json.load(open(os.path.join(GOLDEN, "closed_loop.json")))
N = 120


def check_against_fixture(res, fix, n_ours, what=""):
    n_ref = fix["n_episodes"]
    for key, ref in (("n_successful_runs", fix["successes"] / n_ref), ("n_collisions", fix["collisions"] / n_ref)):
        ours = float(np.mean(res[key]))
        p = (ours * n_ours + ref * n_ref) / (n_ours + n_ref)
        sigma = math.sqrt(max(p * (1 - p), 0.02) * (1.0 / n_ours + 1.0 / n_ref))
        assert abs(ours - ref) <= 3.5 * sigma + 0.02, (what, key, ours, ref, sigma)
    # flight time and path error: heavy-tailed per-episode quantities, compare means loosely
    assert abs(np.mean(res["env_steps"]) - fix["mean_flight_time"]) <= 0.2 * fix["mean_flight_time"], what
    assert abs(np.mean(res["APE"]) - fix["mean_ape"]) <= 0.35 * fix["mean_ape"] + 3.0, what


@pytest.mark.parametrize("mode,scenario", [("test", "corridor"), ("test", "large"), ("test", "S_parallel"),
                                           ("curriculum", "stage_1"), ("curriculum", "stage_5")])
def test_oracle_closed_loop_reproduces_reference_fixtures(mode, scenario):
    res = closed_loop.evaluate(mode, scenario, N, seed=3)
    check_against_fixture(res, FIX[scenario], N, scenario)
