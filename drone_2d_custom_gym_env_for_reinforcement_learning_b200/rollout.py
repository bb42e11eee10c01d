"""Device-resident PPO rollout collection (the caller of `step` in the reference's training
loop, main.py:190-208 -> SB3 `PPO.collect_rollouts`): observations, actions, rewards, episode
starts, values and log-probabilities are written into [T, N] tensors on the GPU, advantages and
returns come from one `d2d_gae` launch.  Nothing crosses PCIe during a rollout; the learner
(SB3's PPO update) is out of scope.
"""
import ctypes as C
import io
import math
import zipfile

import numpy as np
import torch

from . import _lib
from .policy import MlpActor, load_actor_weights

_VKEYS = ("mlp_extractor.value_net.0.weight", "mlp_extractor.value_net.0.bias",
          "mlp_extractor.value_net.2.weight", "mlp_extractor.value_net.2.bias", "value_net.weight", "value_net.bias")


class MlpActorCritic(MlpActor):
    """SB3 MlpPolicy with separate 64-64 tanh towers for actor and critic (the shipped checkpoints)."""

    def __init__(self, actor_weights, value_weights, device):
        super().__init__(actor_weights, device)
        v = {k: t.to(self.device, torch.float32).contiguous() for k, t in value_weights.items()}
        self.v0, self.vb0, self.v1, self.vb1, self.v2, self.vb2 = (v[k] for k in _VKEYS)
        self.log_std = self.std.log()

    @classmethod
    def load(cls, path, device):
        with zipfile.ZipFile(path) as z:
            sd = torch.load(io.BytesIO(z.read("policy.pth")), map_location="cpu", weights_only=True)
        return cls(load_actor_weights(path), {k: sd[k] for k in _VKEYS}, device)

    @classmethod
    def random_init(cls, device, seed=0):
        g = torch.Generator().manual_seed(seed)
        mk = lambda *s: torch.randn(*s, generator=g) * 0.1  # noqa: E731
        a = {"mlp_extractor.policy_net.0.weight": mk(64, 27), "mlp_extractor.policy_net.0.bias": torch.zeros(64),
             "mlp_extractor.policy_net.2.weight": mk(64, 64), "mlp_extractor.policy_net.2.bias": torch.zeros(64),
             "action_net.weight": mk(2, 64) * 0.1, "action_net.bias": torch.zeros(2), "log_std": torch.zeros(2)}
        v = dict(zip(_VKEYS, (mk(64, 27), torch.zeros(64), mk(64, 64), torch.zeros(64), mk(1, 64), torch.zeros(1))))
        return cls(a, v, device)

    def value(self, obs):
        h = torch.tanh(torch.addmm(self.vb0, obs, self.v0.t()))
        h = torch.tanh(torch.addmm(self.vb1, h, self.v1.t()))
        return torch.addmm(self.vb2, h, self.v2.t()).squeeze(-1)

    def act(self, obs, generator=None):
        """SB3 `policy.forward`: unclipped Gaussian sample, its log-probability, the state value."""
        mu = self.mean(obs)
        eps = torch.randn(mu.shape, device=mu.device, dtype=mu.dtype, generator=generator)
        a = mu + self.std * eps
        logp = (-0.5 * eps * eps - self.log_std - 0.5 * math.log(2 * math.pi)).sum(-1)
        return a, logp, self.value(obs)


def gae(rewards, values, episode_starts, last_values, last_dones, gamma=0.99, gae_lambda=0.95):
    """Advantages and returns of a [T, N] rollout through the C ABI (`d2d_gae`)."""
    T, N = rewards.shape
    adv = torch.empty_like(rewards)
    ret = torch.empty_like(rewards)
    L = _lib.load()
    st = C.c_void_p(torch.cuda.current_stream(rewards.device).cuda_stream)
    _lib.check(L.d2d_gae(rewards.data_ptr(), values.data_ptr(), episode_starts.data_ptr(), last_values.data_ptr(),
                         last_dones.data_ptr(), T, N, gamma, gae_lambda, adv.data_ptr(), ret.data_ptr(), st))
    return adv, ret


def gae_reference(rewards, values, episode_starts, last_values, last_dones, gamma=0.99, gae_lambda=0.95):
    """numpy restatement of SB3's RolloutBuffer.compute_returns_and_advantage (tests)."""
    T = rewards.shape[0]
    adv = np.zeros_like(rewards)
    last = 0
    for t in reversed(range(T)):
        if t == T - 1:
            nnt, nv = 1.0 - last_dones.astype(np.float32), last_values
        else:
            nnt, nv = 1.0 - episode_starts[t + 1].astype(np.float32), values[t + 1]
        delta = rewards[t] + gamma * nv * nnt - values[t]
        last = delta + gamma * gae_lambda * nnt * last
        adv[t] = last
    return adv, adv + values


class RolloutCollector:
    """SB3 `collect_rollouts` on the device: T steps of N environments per call."""

    def __init__(self, env, policy, n_steps=128, gamma=0.99, gae_lambda=0.95, seed=0):
        self.env, self.policy, self.T = env, policy, int(n_steps)
        self.gamma, self.lam = gamma, gae_lambda
        n, dev = env.num_envs, env.device
        self.obs = torch.empty((self.T, n, 27), dtype=torch.float32, device=dev)
        self.actions = torch.empty((self.T, n, 2), dtype=torch.float32, device=dev)
        self.rewards = torch.empty((self.T, n), dtype=torch.float32, device=dev)
        self.values = torch.empty((self.T, n), dtype=torch.float32, device=dev)
        self.log_probs = torch.empty((self.T, n), dtype=torch.float32, device=dev)
        self.episode_starts = torch.empty((self.T, n), dtype=torch.uint8, device=dev)
        self.gen = torch.Generator(device=dev)
        self.gen.manual_seed(seed)
        self._last_obs = env.reset_device().clone()
        self._last_starts = torch.ones(n, dtype=torch.uint8, device=dev)
        self.num_timesteps = 0

    def collect(self):
        env, pol = self.env, self.policy
        for t in range(self.T):
            a, logp, v = pol.act(self._last_obs, self.gen)
            self.obs[t].copy_(self._last_obs)
            self.actions[t].copy_(a)
            self.values[t].copy_(v)
            self.log_probs[t].copy_(logp)
            self.episode_starts[t].copy_(self._last_starts)
            obs, rew, done = env.step_device(a.clamp(-1.0, 1.0), terminal_obs=False)     # SB3 clips for the env only
            self.rewards[t].copy_(rew)
            self._last_obs.copy_(obs)
            self._last_starts.copy_(done)
        self.num_timesteps += self.T * env.num_envs
        last_values = pol.value(self._last_obs)
        self.advantages, self.returns = gae(self.rewards, self.values, self.episode_starts, last_values,
                                            self._last_starts, self.gamma, self.lam)
        return self
