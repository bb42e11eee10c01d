"""Policy kernel (d2d_policy_act, tcgen05) against the torch float32 actor, and its launch time."""
import importlib, json, os, sys
import torch
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, REPO)
PKG = "drone_2d_custom_gym_env_for_reinforcement_learning_b200"
pkg = importlib.import_module(PKG); policy = importlib.import_module(PKG + ".policy")
E = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
W = os.path.join(REPO, "tests", "golden", "agent_19_policy.npz")
env = pkg.Drone2dVecEnv(E, seed=0, info_mode="none", mode="test", scenario="large")
obs = env.reset_device()
ref, fused = policy.MlpActor.load(W, env.device), policy.FusedActor.load(W, env)
out = {"envs": E}
for k in range(3):
    a_ref = ref.predict(obs, deterministic=True)
    a_fused = fused.predict(obs, deterministic=True)
    torch.cuda.synchronize()
    out["max_abs_err_%d" % k] = float((a_ref - a_fused).abs().max())
    obs, _, _ = env.step_device(a_ref)
    obs = obs.clone()
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for _ in range(20): fused.predict(obs)
torch.cuda.synchronize()
ev0.record()
for _ in range(200): fused.predict(obs)
ev1.record(); torch.cuda.synchronize()
out["policy_us"] = round(ev0.elapsed_time(ev1) * 1000 / 200, 2)
a = fused.predict(obs).clone(); mu = ref.mean(obs)
z = (a - mu) / ref.std
safe = (mu.abs() + 2 * ref.std) < 1.0
out["noise_mean"] = float(z[safe].mean()); out["noise_std"] = float(z[safe].std()); out["n_safe"] = int(safe.sum())
print(json.dumps(out))
