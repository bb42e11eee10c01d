#!/usr/bin/env python
"""Step-level pin of the restated rigid-body step on the reference's GENUINE pymunk trajectories.

Run ONCE in the authoring container (needs /root/reference, read-only; ~15 min on 8 cores):

    python tests/golden/make_inverse_dynamics.py

For all 700 genuine episodes (Tests/<scenario>/test_4/flight_paths, 7 scenarios x 100) the actions are recovered by
oracle/inverse_dynamics.py over the first HEAD steps, under the restated step and under a table of deliberately
different restatements; for the 70 episodes that are committed as fixtures (reference_trajectories.npz) the actions of
the WHOLE episode are recovered and stored, so that tests can push them through the oracle and through the product
kernel and compare with the recorded positions.

  inverse_dynamics_stats.json     per variant: distribution over episodes of the per-episode median / max residual,
                                  fraction of recovered actions exactly on the edge of the action Box
  inverse_dynamics_actions.npz    per committed episode: alpha0 and the recovered actions [steps-1, 2] (float64)
"""
import json
import multiprocessing as mp
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import binding as B  # noqa: E402
from oracle import flight_replay as FR  # noqa: E402
from oracle import inverse_dynamics as ID  # noqa: E402

ROOT = "/root/reference/Tests"
SCENARIOS = ["perpendicular", "parallel", "S_parallel", "corridor", "S_corridor", "impossible", "large"]
HEAD = 150
# (name, d2o_config overrides, compile flag of the oracle)
VARIANTS = [
    ("restated", {}, None),
    ("iterations_9", {"iterations": 9}, None),
    ("iterations_11", {"iterations": 11}, None),
    ("no_warm_start", {}, "-DD2O_NO_WARM_START"),
    ("joint_order_permuted", {}, "-DD2O_JOINT_ORDER_PERMUTED"),
    ("positions_after_velocities", {}, "-DD2O_POSITIONS_AFTER"),
    ("space_damping_0.9", {}, "-DD2O_SPACE_DAMPING=0.9"),
    ("motor_mass_0.41", {"mass_m": 0.41}, None),
    ("gravity_-1001", {"gravity_y": -1001.0}, None),
]
_TMP = tempfile.mkdtemp(prefix="d2o_variants_")


def variant_lib(flag):
    if flag is None:
        return B.lib()
    so = os.path.join(_TMP, "lib_%s.so" % "".join(ch for ch in flag if ch.isalnum()))
    if not os.path.exists(so):
        subprocess.check_call(["gcc", "-O2", "-fPIC", "-std=c99", "-ffp-contract=off", "-shared", flag, "-o", so,
                               os.path.join(os.path.dirname(B.LIB_PATH), "d2d_oracle.c"), "-lm"])
    return B.load(so)


def load_paths(scen):
    fp = json.load(open("%s/%s/test_4/flight_paths" % (ROOT, scen)))
    return [FR.to_world(p) for p in fp]


def job(args):
    scen, idx, vname, over, flag, steps = args
    xy = load_paths(scen)[idx]
    ep = ID.Episode(xy, cfg=B.default_config(**over), lib=variant_lib(flag))
    rec = ep.recover_actions(max_steps=steps)
    s = ID.summarize(rec, tail=24)
    s["alpha0_exact"] = rec["alpha0_exact"]
    return scen, idx, vname, s, (rec if steps is None else None)


def dist(v):
    v = np.asarray(v, dtype=np.float64)
    return {"min": float(v.min()), "median": float(np.median(v)), "p90": float(np.percentile(v, 90)), "max": float(v.max())}


def main():
    z = np.load(os.path.join(HERE, "reference_trajectories.npz"))
    all_paths = {s: load_paths(s) for s in SCENARIOS}
    # which of the 100 episodes per scenario are the committed ones: match by length and first position
    committed = {}
    for scen in SCENARIOS:
        off, ids = 0, []
        for n in z[scen + "__len"]:
            xy = z[scen + "__xy"][off:off + n]
            off += n
            ids.append(next(i for i, p in enumerate(all_paths[scen]) if len(p) == n and np.array_equal(p[0], xy[0])))
        committed[scen] = ids
    jobs = []
    for vname, over, flag in VARIANTS:
        variant_lib(flag)                                    # build before forking
        eps = range(100) if vname == "restated" else range(0, 100, 5)
        for scen in SCENARIOS:
            jobs += [(scen, i, vname, over, flag, HEAD) for i in eps]
    jobs += [(scen, i, "restated_full", {}, None, None) for scen in SCENARIOS for i in committed[scen]]
    with mp.get_context("fork").Pool(len(os.sched_getaffinity(0))) as pool:
        results = pool.map(job, jobs, chunksize=4)
    stats = {"source": "Tests/<scenario>/test_4/flight_paths (genuine pymunk + SB3 agent_19 runs of the reference)",
             "method": "oracle/inverse_dynamics.py, window 40, first %d steps of every episode (tail of 24 dropped)" % HEAD,
             "variants": {}}
    for vname, _, _ in VARIANTS + [("restated_full", None, None)]:
        rs = [r[3] for r in results if r[2] == vname and r[3] is not None]
        stats["variants"][vname] = {
            "episodes": len(rs), "steps": int(sum(r["steps"] for r in rs)),
            "episode_residual_median": dist([r["residual_median"] for r in rs]),
            "episode_residual_max": dist([r["residual_max"] for r in rs]),
            "frac_actions_at_box_edge": dist([r["frac_actions_at_box_edge"] for r in rs]),
            "episodes_with_median_residual_below_1e-4": int(sum(r["residual_median"] < 1e-4 for r in rs)),
            "episodes_with_undetermined_spawn_angle": int(sum(not r["alpha0_exact"] for r in rs)),
            "per_episode_residual_median": [float("%.3e" % r["residual_median"]) for r in rs],
        }
    arrays = {}
    for scen, idx, vname, s, rec in results:
        if vname == "restated_full":
            k = committed[scen].index(idx)
            arrays["%s__%d__actions" % (scen, k)] = rec["actions"]
            arrays["%s__%d__alpha0" % (scen, k)] = np.float64(rec["alpha0"])
    np.savez_compressed(os.path.join(HERE, "inverse_dynamics_actions.npz"), **arrays)
    with open(os.path.join(HERE, "inverse_dynamics_stats.json"), "w") as f:
        json.dump(stats, f, indent=1)
    for k, v in stats["variants"].items():
        print("%-28s episodes %3d  per-episode median residual: median %.2e max %.2e | per-episode max residual: median %.2e max %.2e | at edge %.3f"
              % (k, v["episodes"], v["episode_residual_median"]["median"], v["episode_residual_median"]["max"],
                 v["episode_residual_max"]["median"], v["episode_residual_max"]["max"], v["frac_actions_at_box_edge"]["median"]))


if __name__ == "__main__":
    main()
