"""Shared driver of the GPU-vs-oracle parity tests (and of __graft_entry__.smoke).

Both sides start every episode from the IDENTICAL float64 state (read back from
the device after its reset, or injected with set_state) and receive identical
float32 actions; they then evolve independently (no re-synchronisation until an
episode ends), so the reported errors are genuine open-loop divergences over
the horizon.
"""
import ctypes as C

import numpy as np

import oracle

# stated tolerances (see DESIGN.md "Parity")
TOL_STATE_STEP = 1e-9        # body pose / velocity after ONE step from an identical state
TOL_STATE_HORIZON = 1e-6     # ... after a 64-step open-loop horizon
TOL_OBS_CORE = 1e-9          # obs[0:19]: no dependence on the Brent search
TOL_U = 5e-5                 # closest-u: Brent on the flat-bottomed distance objective is only
#                              determined to ~sqrt(eps * dist) ~ 1e-5 (see oracle/d2d_oracle.c:norm2)
TOL_REWARD = 1e-6


def blob_to_states(blob):
    n = blob.shape[0]
    states = (oracle.State * n)()
    for i in range(n):
        b = blob[i]
        s = states[i]
        for k in range(3):
            body = s.body[k]
            body.px, body.py, body.vx, body.vy, body.a, body.w = b[6 * k:6 * k + 6]
        for j in range(6):
            s.jx[j], s.jy[j] = b[18 + 2 * j], b[19 + 2 * j]
        s.collided, s.la_latched, s.steps = int(b[30]), int(b[31]), int(b[32])
        s.path_error, s.total_reward = b[33], b[34]
    return states


def states_to_blob(states, episode=None):
    n = len(states)
    blob = np.zeros((n, 36))
    for i in range(n):
        s = states[i]
        for k in range(3):
            body = s.body[k]
            blob[i, 6 * k:6 * k + 6] = [body.px, body.py, body.vx, body.vy, body.a, body.w]
        for j in range(6):
            blob[i, 18 + 2 * j], blob[i, 19 + 2 * j] = s.jx[j], s.jy[j]
        blob[i, 30:35] = [s.collided, s.la_latched, s.steps, s.path_error, s.total_reward]
        blob[i, 35] = 0 if episode is None else episode[i]
    return blob


def tables_to_structs(tables, n_wps, idx=None):
    """Per-env oracle Path / Obstacles arrays from Drone2dVecEnv.get_tables()."""
    n = len(tables["n_obs"])
    idx = range(n) if idx is None else idx
    paths, obsts = {}, {}
    for i in idx:
        paths[i] = oracle.make_path(tables["wps"][i, :n_wps], tables["us"][i, :n_wps],
                                    tables["xparams"][i, :n_wps - 2], tables["yparams"][i, :n_wps - 2])
        k = int(tables["n_obs"][i])
        obsts[i] = oracle.make_obstacles([tuple(tables["obstacles"][i, j]) for j in range(k)])
    return paths, obsts


def oracle_config_from(env):
    c = env.cfg
    return oracle.default_config(
        screen_x=c.screen_x, screen_y=c.screen_y, lookahead=c.lookahead, danger_range=c.danger_range,
        danger_angle=c.danger_angle, abs_inv_CA_min_rew=c.abs_inv_CA_min_rew, PA_band_edge=c.PA_band_edge,
        PA_scale=c.PA_scale, PP_vel_scale=c.PP_vel_scale, PP_rew_max=c.PP_rew_max, PP_rew_min=c.PP_rew_min,
        rew_collision=c.rew_collision, reach_end_radius=c.reach_end_radius, rew_reach_end=c.rew_reach_end,
        AA_angle=c.AA_angle, AA_band=c.AA_band, rew_AA=c.rew_AA, use_Lambda=c.use_Lambda, n_steps=c.n_steps)


def run_parity(pkg, mode, scenario, n_envs, n_rounds, seed=0, action_scale=1.0, inject_spawn=None, **kwargs):
    """Step `n_envs` environments for `n_steps` on the GPU and in the oracle; returns error statistics."""
    import torch
    env = pkg.Drone2dVecEnv(n_envs, seed=seed, info_mode="full", mode=mode, scenario=scenario, **kwargs)
    cfg = oracle_config_from(env)
    per_env = (mode == "curriculum")
    env.reset_device()
    if inject_spawn is not None:
        # states produced on the oracle side (reference spawn distribution) and injected into the device
        st0 = (oracle.State * n_envs)()
        for i, (x, y, a) in enumerate(inject_spawn):
            st0[i] = oracle.spawn(x, y, a, cfg)
        env.set_state(states_to_blob(st0))
    blob = env.get_state()
    states = blob_to_states(blob)
    n_wps = env.cfg.n_wps if per_env else len(env.scenario["wps"])
    if per_env:
        paths, obsts = tables_to_structs(env.get_tables(), n_wps)
    else:
        sc = env.scenario
        path = oracle.make_path(sc["wps"], sc["us"], sc["xparams"], sc["yparams"])
        obst = oracle.make_obstacles([tuple(o) for o in sc["obstacles"]])
    rng = np.random.RandomState(seed + 1)
    age = np.zeros(n_envs, dtype=int)                   # steps since both sides were last identical
    R = dict(max_state_err=0.0, max_state_err_1step=0.0, max_obs_core_err=0.0, max_obs_path_err=0.0,
             max_u_err=0.0, max_reward_err=0.0, flag_mismatch=0, nearest_mismatch=0, nfev_mismatch=0,
             u_gt_1e9=0, env_steps=0, episodes=0, max_f32_obs_err=0.0, max_ape_err=0.0, max_reset_obs_err=0.0,
             reset_obs_checked=0)
    for t in range(n_rounds):
        actions = (rng.uniform(-1, 1, (n_envs, 2)) * action_scale).astype(np.float32)
        env.step_device(torch.from_numpy(actions).to(env.device))
        g_obs64 = env.obs64.cpu().numpy()
        g_info = env.info.cpu().numpy()
        g_rew = env.rewards.cpu().numpy()
        g_done = env.dones.cpu().numpy().astype(bool)
        g_term = env.terminal_obs.cpu().numpy()
        g_obs32 = env.obs.cpu().numpy()
        blob = env.get_state()
        outs = []
        for i in range(n_envs):
            p, o = (paths[i], obsts[i]) if per_env else (path, obst)
            outs.append(oracle.env_step(states[i], p, o, cfg, actions[i, 0], actions[i, 1]))
        age += 1
        resync = []
        for i in range(n_envs):
            out = outs[i]
            R["env_steps"] += 1
            o_obs = np.array(out.obs[:])
            if bool(out.done) != bool(g_done[i]) or int(g_info[i, 11]) != (
                    out.end_collision | 2 * out.end_reached | 4 * out.end_time | 8 * out.end_alpha):
                R["flag_mismatch"] += 1
                resync.append(i)
                continue
            near = [(int(g_info[i, 14]) >> (8 * j)) & 255 for j in range(3)]
            near = [-1 if v == 255 else v for v in near]
            if near != list(out.near_idx[:]):
                R["nearest_mismatch"] += 1
            du = abs(g_info[i, 12] - out.u_star)
            R["max_u_err"] = max(R["max_u_err"], du)
            R["u_gt_1e9"] += du > 1e-9
            R["nfev_mismatch"] += int(g_info[i, 13]) != out.nfev
            R["max_obs_core_err"] = max(R["max_obs_core_err"], np.abs(g_obs64[i, :19] - o_obs[:19]).max())
            R["max_obs_path_err"] = max(R["max_obs_path_err"], np.abs(g_obs64[i, 19:23] - o_obs[19:23]).max())
            R["max_reward_err"] = max(R["max_reward_err"], abs(g_info[i, 0] - out.reward))
            R["max_ape_err"] = max(R["max_ape_err"], abs(g_info[i, 9] - out.ape))
            assert abs(float(g_rew[i]) - np.float32(g_info[i, 0])) <= 1e-6 * max(1.0, abs(g_info[i, 0]))
            f32 = g_term[i] if g_done[i] else g_obs32[i]
            R["max_f32_obs_err"] = max(R["max_f32_obs_err"], np.abs(f32 - g_obs64[i].astype(np.float32)).max())
            if g_done[i]:
                R["episodes"] += 1
                resync.append(i)
                continue
            s = states[i]
            ob = np.array([[b.px, b.py, b.vx, b.vy, b.a, b.w] for b in s.body]).ravel()
            err = np.abs(blob[i, :18] - ob).max()
            R["max_state_err"] = max(R["max_state_err"], err)
            if age[i] == 1:
                R["max_state_err_1step"] = max(R["max_state_err_1step"], err)
        if resync:
            # finished (auto-reset on the device) or diverged: restart from the device's new state
            new_states = blob_to_states(blob[resync])
            if per_env:
                tb = env.get_tables()
                np_, no_ = tables_to_structs(tb, n_wps, resync)
            for k, i in enumerate(resync):
                states[i] = new_states[k]
                age[i] = 0
                if per_env:
                    paths[i], obsts[i] = np_[i], no_[i]
                if g_done[i]:
                    # the observation returned at a done step is the first observation of the NEW episode
                    # (env.reset() -> get_observation(), env:908-912): check it against the oracle
                    out0 = oracle.StepOut()
                    pth, obt = (paths[i], obsts[i]) if per_env else (path, obst)
                    oracle.lib().d2o_observe(C.byref(states[i]), C.byref(pth), C.byref(obt), C.byref(cfg), C.byref(out0))
                    ref0 = np.array(out0.obs[:])
                    dist = max(np.hypot(out0.closest_x - states[i].body[0].px, out0.closest_y - states[i].body[0].py), 1e-3)
                    tol_ang = np.zeros(27)
                    tol_ang[25:27] = 2 * TOL_U / dist      # bearing of the closest point: ill-conditioned near the path
                    tol_ang[23:25] = 2 * TOL_U / 100.0
                    err = np.abs(g_obs32[i].astype(np.float64) - ref0) - tol_ang
                    R["max_reset_obs_err"] = max(R["max_reset_obs_err"], float(err.max()))
                    R["reset_obs_checked"] += 1
    R["launches"] = env.launch_count
    env.close()
    return {k: (float(v) if isinstance(v, (np.floating, float)) else int(v)) for k, v in R.items()}
