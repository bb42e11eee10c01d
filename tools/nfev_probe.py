"""How divergent is the closest-point search inside a warp?  (decides whether lane compaction pays)"""
import importlib, os, sys
import numpy as np, torch
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, REPO)
pkg = importlib.import_module("drone_2d_custom_gym_env_for_reinforcement_learning_b200")
E = 65536
env = pkg.Drone2dVecEnv(E, seed=0, info_mode="full", mode="curriculum", scenario="stage_5")
env.reset_device()
g = torch.Generator(device=env.device); g.manual_seed(1)
ring = (torch.rand((256, E, 2), device=env.device, generator=g) * 2 - 1).contiguous()
for k in range(400): env.step_device(ring[k % 256], terminal_obs=False)
nf = env.info[:, 13].cpu().numpy().astype(int)
w = nf.reshape(-1, 32)
print("lane nfev: mean %.1f median %d p90 %d p99 %d max %d" % (nf.mean(), np.median(nf), np.percentile(nf, 90), np.percentile(nf, 99), nf.max()))
print("per-warp max: mean %.1f ; per-warp mean: %.1f ; efficiency %.2f" % (w.max(1).mean(), w.mean(1).mean(), w.mean() / w.max(1).mean()))
b = nf.reshape(-1, 128)
K = 6
tot = 0
for blk in b:
    for r in range(0, 60, K):
        act = (blk > r).sum()
        tot += -(-act // 32) * K
print("block compaction (128 lanes, rounds of %d): warp-iterations per block %.1f vs now %.1f" % (K, tot / len(b), w.max(1).reshape(-1, 4).sum(1).mean()))
print("hist:", np.bincount(np.minimum(nf, 50))[5:51].tolist())
