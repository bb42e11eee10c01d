"""GPU twin of tests/test_inverse_dynamics_cpu.py: the PRODUCT kernel's rigid-body step against genuine pymunk
output.  The motor commands recovered from the reference's recorded trajectories (tests/golden/
inverse_dynamics_actions.npz, see oracle/inverse_dynamics.py) are fed to `d2d_step` through the C ABI:

  * one step from every hidden state of six genuine episodes (states along the oracle's replay, float32 actions as
    the gym boundary carries them): the kernel's new frame velocity equals the velocity pymunk produced
    (first difference of the recorded positions) to ~1e-6 px/s -- and a kernel configured with 9 or 11 solver
    iterations instead of pymunk's 10 misses it by four orders of magnitude more;
  * open loop from the spawn pose, whole episodes: the kernel flies the 70 genuine trajectories to a few hundredths
    of a pixel (float32 commands).
"""
import importlib
import os

import numpy as np
import pytest

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
SCENARIOS = ["perpendicular", "parallel", "S_parallel", "corridor", "S_corridor", "impossible", "large"]
PKG = "drone_2d_custom_gym_env_for_reinforcement_learning_b200"
pytestmark = pytest.mark.gpu


def _load():
    z = np.load(os.path.join(GOLDEN, "reference_trajectories.npz"))
    a = np.load(os.path.join(GOLDEN, "inverse_dynamics_actions.npz"))
    eps = []
    for scen in SCENARIOS:
        off = 0
        for k, n in enumerate(z[scen + "__len"]):
            eps.append(dict(scenario=scen, xy=z[scen + "__xy"][off:off + n], alpha0=float(a["%s__%d__alpha0" % (scen, k)]),
                            actions=a["%s__%d__actions" % (scen, k)]))
            off += n
    return eps


def _hidden_states(ep):
    """Oracle replay of one episode under the recovered (float32-rounded) actions: state before every step."""
    import ctypes as C
    import oracle
    from parity_util import states_to_blob
    cfg, none, lib = oracle.default_config(), oracle.make_obstacles([]), oracle.lib()
    acts = ep["actions"]                                 # float64, as recovered: the hidden states follow them exactly
    n = len(acts) - 24                                   # the last actions of an episode are not pinned by data
    s = oracle.spawn(ep["xy"][0, 0], ep["xy"][0, 1], ep["alpha0"], cfg)
    states = (oracle.State * n)()
    for k in range(n):
        states[k] = s
        lib.d2o_space_step(C.byref(s), C.byref(none), C.byref(cfg), float(acts[k, 0]), float(acts[k, 1]))
    vobs = (ep["xy"][1:] - ep["xy"][:-1]) * 60.0
    # the C ABI carries float32 actions (like the gym boundary): rounding costs <= 3e-8 * 8 px/s per step
    return states_to_blob(states), acts[:n].astype(np.float32), vobs[:n]


def _one_step_residual(pkg, blob, acts, vobs, **kwargs):
    import torch
    env = pkg.Drone2dVecEnv(len(blob), seed=0, info_mode="none", mode="test", scenario="corridor", **kwargs)
    env.reset_device()
    env.set_state(blob)
    env.step_device(torch.from_numpy(acts).to(env.device), auto_reset=False, terminal_obs=False)
    out = env.get_state()
    env.close()
    return np.abs(out[:, 2:4] - vobs).max(1)


def test_kernel_step_reproduces_genuine_pymunk_velocities():
    pkg = importlib.import_module(PKG)
    eps = _load()
    parts = [_hidden_states(eps[i]) for i in (0, 13, 27, 30, 44, 58)]
    blob, acts, vobs = (np.concatenate([p[j] for p in parts]) for j in range(3))
    assert len(blob) > 1500
    res = _one_step_residual(pkg, blob, acts, vobs)
    assert np.median(res) < 2e-6, np.median(res)
    assert np.percentile(res, 99) < 1e-3, np.percentile(res, 99)
    for it in (9, 11):
        other = _one_step_residual(pkg, blob, acts, vobs, solver_iterations=it)
        assert np.median(other) > 1000 * max(np.median(res), 1e-7), (it, np.median(other), np.median(res))


def test_kernel_flies_the_genuine_trajectories_open_loop():
    import oracle
    import torch
    from parity_util import states_to_blob
    pkg = importlib.import_module(PKG)
    eps = _load()
    n, T = len(eps), max(len(e["xy"]) for e in eps)
    cfg = oracle.default_config()
    st0 = (oracle.State * n)()
    acts = np.zeros((T, n, 2), dtype=np.float32)
    lens = np.array([len(e["xy"]) for e in eps])
    for i, e in enumerate(eps):
        st0[i] = oracle.spawn(e["xy"][0, 0], e["xy"][0, 1], e["alpha0"], cfg)
        acts[:len(e["actions"]), i] = e["actions"].astype(np.float32)
    env = pkg.Drone2dVecEnv(n, seed=0, info_mode="none", mode="test", scenario="corridor", n_steps=10 ** 6)
    env.reset_device()
    env.set_state(states_to_blob(st0))
    a_dev = torch.from_numpy(acts).to(env.device)
    worst = np.zeros(n)
    for k in range(T):
        env.step_device(a_dev[k], auto_reset=False, terminal_obs=False)
        pos = env.get_state()[:, 0:2]                    # frame position after step k = recorded xy[k]
        live = np.flatnonzero(k < lens - 24)
        if len(live):
            ref = np.stack([eps[i]["xy"][k] for i in live])
            worst[live] = np.maximum(worst[live], np.abs(pos[live] - ref).max(1))
    env.close()
    # px, after up to ~1100 open-loop steps.  The recovered commands enter as float32 (C ABI); their rounding (3e-8) is
    # integrated twice through the attitude, which costs up to a few hundredths of a pixel over a whole episode (the
    # oracle, fed the unrounded float64 commands, stays within 1e-3 px: tests/test_inverse_dynamics_cpu.py)
    assert worst.max() < 0.5, worst.max()
    assert np.median(worst) < 0.05, np.median(worst)
