"""Step-kernel experiment: what would lanes sorted by search length buy, WITHOUT the scattered-access penalty of an
indirection?  A steady-state batch is read back (state + tables + per-env evaluation counts), physically re-ordered
(identity = control, sorted by nfev descending, sorted by n_obs then nfev) and written into fresh handles; a few steps
without auto-reset are then timed with the per-phase warp clock spans.
usage: python tools/sort_probe.py [envs] [scenario]"""
import importlib
import json
import os
import sys

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
pkg = importlib.import_module("drone_2d_custom_gym_env_for_reinforcement_learning_b200")

E = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
scen = sys.argv[2] if len(sys.argv) > 2 else "stage_5"
mode = "curriculum" if scen.startswith("stage") else "test"
dev = torch.device("cuda")
g = torch.Generator(device=dev)
g.manual_seed(1)
R = 64
ring = (torch.rand((R, E, 2), device=dev, generator=g) * 2 - 1).contiguous()

src = pkg.Drone2dVecEnv(E, seed=0, info_mode="full", mode=mode, scenario=scen)
src.reset_device()
for k in range(500):
    src.step_device(ring[k % R], terminal_obs=False)
nfev = src.info[:, 13].cpu().numpy().astype(int)
blob = src.get_state()
tabs = src.get_tables() if mode == "curriculum" else None
nobs = tabs["n_obs"] if tabs is not None else np.zeros(E, int)
print(json.dumps({"nfev_mean": float(nfev.mean()), "nfev_p50": int(np.median(nfev)), "frac_ge_40": float((nfev >= 40).mean()),
                  "hist": np.bincount(np.minimum(nfev, 50) // 5).tolist()}), flush=True)
src.close()


def warp_stats(order):
    n = nfev[order][: E // 32 * 32].reshape(-1, 32)
    return {"mean_warp_max_nfev": float(n.max(1).mean()), "mean_nfev": float(n.mean())}


def run(name, order):
    env = pkg.Drone2dVecEnv(E, seed=0, info_mode="none", mode=mode, scenario=scen)
    env.reset_device()
    if tabs is not None:
        env.set_tables(tabs["wps"][order], tabs["us"][order], tabs["xparams"][order], tabs["yparams"][order],
                       tabs["n_obs"][order], tabs["obstacles"][order])
    env.set_state(blob[order])
    idx = torch.from_numpy(order).to(dev)
    acts = ring[:, idx].contiguous()
    env.step_device(acts[0], auto_reset=False, terminal_obs=False)       # warm caches / first touch
    env.role_cycles()
    us = []
    for k in range(1, 9):
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        env.step_device(acts[k], auto_reset=False, terminal_obs=False)
        ev1.record()
        torch.cuda.synchronize()
        us.append(ev0.elapsed_time(ev1) * 1000)
    rc = env.role_cycles()
    chk = float(env.rewards.double().sum().item())
    print(json.dumps({"order": name, "us_per_step": [round(u, 1) for u in us], **warp_stats(order), "reward_sum": round(chk, 6),
                      "kcyc(max,mean)": {k: (v["max"] // 1000, v["mean"] // 1000) for k, v in rc.items() if v["warps"]}}), flush=True)
    env.close()


ident = np.arange(E)
run("identity", ident)
run("nfev_desc", np.argsort(-nfev, kind="stable"))
run("nobs_then_nfev", np.lexsort((-nfev, -nobs)))
run("random", np.random.RandomState(0).permutation(E))


def blockwise(keyfn, B=128):
    out = []
    for b in range(0, E, B):
        idx = np.arange(b, min(b + B, E))
        out.append(idx[keyfn(idx)])
    return np.concatenate(out)


run("block128_nfev_desc", blockwise(lambda idx: np.argsort(-nfev[idx], kind="stable")))
run("block128_nobs_then_nfev", blockwise(lambda idx: np.lexsort((-nfev[idx], -nobs[idx]))))
run("block128_nfev_then_nobs", blockwise(lambda idx: np.lexsort((-nobs[idx], -nfev[idx]))))
run("global_nfev_then_nobs", np.lexsort((-nobs, -nfev)))
