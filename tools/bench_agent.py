"""BASELINE config 4: `large` scenario, agent_19 policy driving 65536 envs end to end on the device."""
import importlib, json, os, sys
import torch
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, REPO)
PKG = "drone_2d_custom_gym_env_for_reinforcement_learning_b200"
pkg = importlib.import_module(PKG); policy = importlib.import_module(PKG + ".policy"); stats = importlib.import_module(PKG + ".stats")
E = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
scen = sys.argv[2] if len(sys.argv) > 2 else "large"
W = os.path.join(REPO, "tests", "golden", "agent_19_policy.npz")
out = {"envs": E, "scenario": scen}
for name in ("fused_kernel_rollout", "torch_actor"):
    env = pkg.Drone2dVecEnv(E, seed=0, info_mode="none", mode="curriculum" if scen.startswith("stage") else "test", scenario=scen)
    obs = env.reset_device()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    K = 1500
    if name == "fused_kernel_rollout":
        actor = policy.FusedActor.load(W, env)
        actor.rollout(300); torch.cuda.synchronize(); env.episode_stats()
        ev0.record(); actor.rollout(K); ev1.record()
    else:
        actor = policy.MlpActor.load(W, env.device)
        for _ in range(300): obs, _, _ = env.step_device(actor.predict(obs), terminal_obs=False)
        torch.cuda.synchronize(); env.episode_stats()
        ev0.record()
        for _ in range(K): obs, _, _ = env.step_device(actor.predict(obs), terminal_obs=False)
        ev1.record()
    torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1)
    s = stats.summarize(list(env.episode_stats().values()))
    out[name] = {"us_per_step": round(1000 * ms / K, 1), "env_steps_per_s": round(E * K / ms * 1000), "episodes": s["episodes"],
                 "success_rate": round(s["Success rate"], 3), "collision_rate": round(s["Collision rate"], 3),
                 "mean_flight_time": round(s["Average flight time"], 1)}
    env.close()
print(json.dumps(out))
