"""agent_19 on all 12 settings, 4096 GPU episodes each, next to the reference's fixtures -> profiles/<tag>_closed_loop.md"""
import importlib, json, os, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, REPO)
PKG = "drone_2d_custom_gym_env_for_reinforcement_learning_b200"
pkg = importlib.import_module(PKG); policy = importlib.import_module(PKG + ".policy")
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
fix = json.load(open(os.path.join(REPO, "tests", "golden", "closed_loop.json")))
rows = ["# %s — closed loop with the shipped agent_19 actor (GPU, 4096 episodes per setting)" % tag, "",
        "Reference columns: the reference's own 100-episode pymunk artefacts (Tests/<setting>/test_k/results.txt).", "",
        "| setting | success % (ours / ref) | collision % (ours / ref) | mean APE px (ours / ref) | mean flight time (ours / ref) |",
        "|---|---|---|---|---|"]
for name, f in fix.items():
    mode = "curriculum" if name.startswith("stage") else "test"
    env = pkg.Drone2dVecEnv(4096, seed=11, mode=mode, scenario=name)
    actor = policy.MlpActor.load(os.path.join(REPO, "tests", "golden", "agent_19_policy.npz"), env.device)
    r = policy.rollout_one_episode_each(env, actor, seed=5)
    env.close()
    rows.append("| %s | %.1f / %d | %.1f / %d | %.1f / %.1f | %.1f / %.1f |" % (
        name, 100 * r["n_successful_runs"].mean(), f["successes"], 100 * r["n_collisions"].mean(), f["collisions"],
        r["APE"].mean(), f["mean_ape"], r["env_steps"].mean(), f["mean_flight_time"]))
    print(rows[-1], flush=True)
os.makedirs(os.path.join(REPO, "gpurun_out"), exist_ok=True)
open(os.path.join(REPO, "gpurun_out", "%s_closed_loop.md" % tag), "w").write("\n".join(rows) + "\n")
