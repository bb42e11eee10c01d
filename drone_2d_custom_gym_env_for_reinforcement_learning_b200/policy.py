"""Torch-only loader and forward pass for the reference's shipped SB3 PPO
checkpoints (ppo_agents/*.zip, main.py:251 `PPO.load`, main.py:262
`model.predict`).  stable-baselines3 is not needed: the actor of an SB3 2.1
`MlpPolicy` is  obs -> Linear(27,64) -> tanh -> Linear(64,64) -> tanh ->
Linear(64,2) = mean;  action = clip(mean + exp(log_std) * eps, -1, 1)  with
eps ~ N(0,1) (stochastic `predict`, the reference's default) or eps = 0.

The weights stay on the device, so policy and environment exchange
observations / actions without leaving HBM (`rollout`).
"""
import io
import zipfile

import numpy as np
import torch

_KEYS = ("mlp_extractor.policy_net.0.weight", "mlp_extractor.policy_net.0.bias",
         "mlp_extractor.policy_net.2.weight", "mlp_extractor.policy_net.2.bias",
         "action_net.weight", "action_net.bias", "log_std")


def load_actor_weights(path):
    """Actor tensors from an SB3 .zip (policy.pth inside) or from the .npz the golden fixture uses."""
    if str(path).endswith(".npz"):
        z = np.load(path)
        return {k: torch.from_numpy(z[k.replace(".", "__")]) for k in _KEYS}
    with zipfile.ZipFile(path) as z:
        sd = torch.load(io.BytesIO(z.read("policy.pth")), map_location="cpu", weights_only=True)
    missing = [k for k in _KEYS if k not in sd]
    if missing:
        raise ValueError("not an SB3 MlpPolicy actor checkpoint, missing %s" % missing)
    if tuple(sd[_KEYS[0]].shape)[1] != 27:
        raise ValueError("checkpoint expects observations of size %d, the environment emits 27 "
                         "(legacy agent)" % sd[_KEYS[0]].shape[1])
    return {k: sd[k] for k in _KEYS}


class MlpActor:
    def __init__(self, weights, device):
        self.device = torch.device(device)
        w = {k: v.to(self.device, torch.float32).contiguous() for k, v in weights.items()}
        self.w0, self.b0 = w[_KEYS[0]], w[_KEYS[1]]
        self.w1, self.b1 = w[_KEYS[2]], w[_KEYS[3]]
        self.w2, self.b2 = w[_KEYS[4]], w[_KEYS[5]]
        self.std = w["log_std"].exp()

    @classmethod
    def load(cls, path, device):
        return cls(load_actor_weights(path), device)

    def mean(self, obs):
        h = torch.tanh(torch.addmm(self.b0, obs, self.w0.t()))
        h = torch.tanh(torch.addmm(self.b1, h, self.w1.t()))
        return torch.addmm(self.b2, h, self.w2.t())

    def predict(self, obs, deterministic=False, generator=None):
        """SB3 `predict`: Gaussian sample around the mean (unless deterministic), clipped to the action Box."""
        mu = self.mean(obs)
        if not deterministic:
            mu = mu + self.std * torch.randn(mu.shape, device=mu.device, dtype=mu.dtype, generator=generator)
        return mu.clamp_(-1.0, 1.0)


class FusedActor:
    """The same actor through the C ABI (`d2d_policy_act`): one hand-written kernel per call, Gaussian
    noise from the environment handle's Philox stream -- no torch kernels between two env steps."""

    def __init__(self, weights, env):
        import ctypes as C
        from . import _lib
        self.env, self._L, self._C, self._lib = env, _lib.load(), C, _lib
        w = [weights[k].detach().to(torch.float32).contiguous().cpu() for k in _KEYS]
        nbytes = self._L.d2d_policy_weights_bytes()
        host = torch.zeros(nbytes, dtype=torch.uint8)
        _lib.check(self._L.d2d_policy_pack(w[0].data_ptr(), w[1].data_ptr(), w[2].data_ptr(), w[3].data_ptr(),
                                           w[4].data_ptr(), w[5].data_ptr(), w[6].data_ptr(), host.data_ptr()))
        self.packed = host.to(env.device)
        self.actions = torch.empty((env.num_envs, 2), dtype=torch.float32, device=env.device)
        self.step = 0

    @classmethod
    def load(cls, path, env):
        return cls(load_actor_weights(path), env)

    def predict(self, obs, deterministic=False, generator=None):
        st = self._C.c_void_p(torch.cuda.current_stream(self.env.device).cuda_stream)
        self._lib.check(self._L.d2d_policy_act(self.env._h, self.packed.data_ptr(), obs.data_ptr(), self.actions.data_ptr(),
                                               0 if deterministic else 1, self.step, st))
        self.step += 1
        return self.actions

    def rollout(self, n_steps, deterministic=False):
        """n_steps closed-loop transitions enqueued from C (`d2d_rollout`): 2 launches per step, no host work."""
        env = self.env
        st = self._C.c_void_p(torch.cuda.current_stream(env.device).cuda_stream)
        self._lib.check(self._L.d2d_rollout(env._h, self.packed.data_ptr(), env.obs.data_ptr(), self.actions.data_ptr(),
                                            env.rewards.data_ptr(), env.dones.data_ptr(), int(n_steps),
                                            0 if deterministic else 1, self.step, st))
        self.step += int(n_steps)
        return env.obs


def rollout_one_episode_each(env, actor, deterministic=False, seed=0, max_steps=None):
    """Every env plays exactly ONE episode (the reference's test loop, main.py:258-287, one run per
    env): returns per-episode arrays success / fail / collision / APE / steps / total_reward."""
    g = torch.Generator(device=env.device)
    g.manual_seed(seed)
    if env.info is None:
        env.enable_taps()
    obs = env.reset_device()
    n = env.num_envs
    alive = torch.ones(n, dtype=torch.bool, device=env.device)
    out = torch.zeros((n, 6), dtype=torch.float64, device=env.device)
    max_steps = max_steps or env.cfg.n_steps
    for _ in range(max_steps):
        act = actor.predict(obs, deterministic, g)
        obs, _, done = env.step_device(act, auto_reset=False, terminal_obs=False)
        newly = alive & (done != 0)
        if newly.any():
            info = env.info_rows(newly)
            bits = info[:, 11].long()
            rec = torch.stack([((bits & 2) != 0).double(), ((bits & 13) != 0).double(),
                               (((bits & 1) != 0) & ((bits & 14) == 0)).double(), info[:, 9], info[:, 7], info[:, 10]], 1)
            out[newly] = rec
            alive &= ~newly
        if not bool(alive.any()):
            break
    r = out.cpu().numpy()
    return {"n_successful_runs": r[:, 0], "n_failed_runs": r[:, 1], "n_collisions": r[:, 2], "APE": r[:, 3],
            "env_steps": r[:, 4], "total_reward": r[:, 5], "finished": (~alive).cpu().numpy()}
