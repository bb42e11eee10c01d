"""Closed-loop evaluation on the CPU oracle (TEST INFRASTRUCTURE ONLY): the
reference's test loop (main.py:242-327) with the shipped agent_19 actor, one
episode per environment, reset distributions drawn with numpy/`random` exactly
as the reference does (oracle/ref_port.py), transitions by the C restatement."""
import ctypes as C
import os
import random

import numpy as np

from . import binding as B
from . import ref_port


def load_actor(npz_path=None):
    z = np.load(npz_path or os.path.join(B.GOLDEN, "agent_19_policy.npz"))
    return {k: z[k].astype(np.float64) for k in z.files}


def act(actor, obs, rng, deterministic=False):
    """SB3 MlpPolicy actor on float32-rounded observations, Gaussian sampling, clip to [-1, 1]."""
    x = obs.astype(np.float32).astype(np.float64)
    h = np.tanh(x @ actor["mlp_extractor__policy_net__0__weight"].T + actor["mlp_extractor__policy_net__0__bias"])
    h = np.tanh(h @ actor["mlp_extractor__policy_net__2__weight"].T + actor["mlp_extractor__policy_net__2__bias"])
    mu = h @ actor["action_net__weight"].T + actor["action_net__bias"]
    if not deterministic:
        mu = mu + np.exp(actor["log_std"]) * rng.standard_normal(mu.shape)
    return np.clip(mu, -1, 1).astype(np.float32)


def evaluate(mode, scenario, n_episodes, seed=0, deterministic=False, cfg_kwargs=None, cfg_overrides=None,
             record_paths=False, lib_path=None):
    """`cfg_overrides` sets raw d2o_config fields (e.g. iterations), `lib_path` runs the transitions on
    another build of the oracle, `record_paths` adds the frame position after every step
    (world coordinates) as res["flight_paths"]; all three serve tests/test_reference_trajectories_cpu.py."""
    random.seed(seed)
    np.random.seed(seed)
    rng = np.random.RandomState(seed + 77)
    actor = load_actor()
    kw = dict(cfg_kwargs or {})
    cfg = ref_port.config_from_kwargs(kw)
    for k, v in (cfg_overrides or {}).items():
        if not hasattr(cfg, k):
            raise KeyError(k)
        setattr(cfg, k, v)
    draws = ref_port.NumpyDraws()
    n = n_episodes
    states = (B.State * n)()
    shared = (mode == "test")
    if shared:
        path, obst = B.load_scenario(scenario)
        paths, obsts = path, obst
    else:
        paths, obsts = (B.Path * n)(), (B.Obstacles * n)()
    for i in range(n):
        pth, ob, (x, y, a) = ref_port.reset_tables(mode, scenario, kw, draws)
        states[i] = B.spawn(x, y, a, cfg)
        if not shared:
            paths[i] = pth.to_struct()
            obsts[i] = B.make_obstacles(ob)
    lib = B.load(lib_path) if lib_path else B.lib()
    flights = [[] for _ in range(n)]
    obs = np.zeros((n, 27))
    out0 = B.StepOut()
    for i in range(n):
        lib.d2o_observe(C.byref(states[i]), C.byref(paths if shared else paths[i]), C.byref(obsts if shared else obsts[i]),
                        C.byref(cfg), C.byref(out0))
        obs[i] = out0.obs[:]
    alive = np.ones(n, bool)
    res = {k: np.zeros(n) for k in ("n_successful_runs", "n_failed_runs", "n_collisions", "APE", "env_steps", "total_reward")}
    for _ in range(cfg.n_steps):
        actions = act(actor, obs, rng, deterministic)
        outs = B.env_step_batch(states, paths, obsts, cfg, actions, L=lib)
        for i in range(n):
            o = outs[i]
            obs[i] = o.obs[:]
            if record_paths and alive[i]:
                flights[i].append((states[i].body[0].px, states[i].body[0].py))
            if alive[i] and o.done:
                alive[i] = False
                res["n_successful_runs"][i] = o.end_reached
                res["n_failed_runs"][i] = int(o.end_collision or o.end_time or o.end_alpha)
                res["n_collisions"][i] = int(o.end_collision and not (o.end_reached or o.end_time or o.end_alpha))
                res["APE"][i] = o.ape
                res["env_steps"][i] = states[i].steps
                res["total_reward"][i] = states[i].total_reward
        if not alive.any():
            break
    if record_paths:
        res["flight_paths"] = [np.array(f) for f in flights]
    return res
