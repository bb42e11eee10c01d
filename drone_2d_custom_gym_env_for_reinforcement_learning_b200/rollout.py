"""Device-resident PPO rollout collection (the caller of `step` in the reference's training
loop, main.py:190-208 -> SB3 `PPO.collect_rollouts`): observations, actions, rewards, episode
starts, values and log-probabilities are written into [T, N] tensors on the GPU by the hand-written
kernels themselves (`RolloutCollector` -> `d2d_rollout_collect`), advantages and returns come from one
`d2d_gae` launch.  Nothing crosses PCIe during a rollout; the learner (SB3's PPO update) is out of scope.
`MlpActorCritic` is the plain torch float32 statement of the same policy (the reference the tests compare with).
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


def pack_tower(w0, b0, w1, b1, w2, b2, log_std, device):
    """`d2d_policy_pack` of one 27 -> 64 -> 64 -> (<= 2) tanh tower: the shared-memory image the tcgen05 kernel reads."""
    L = _lib.load()
    rows = [t.detach().to(torch.float32).cpu().contiguous() for t in (w0, b0, w1, b1)]
    w2p, b2p = torch.zeros((2, 64), dtype=torch.float32), torch.zeros(2, dtype=torch.float32)
    w2p[:w2.shape[0]] = w2.detach().float().cpu()
    b2p[:b2.shape[0]] = b2.detach().float().cpu()
    ls = torch.zeros(2, dtype=torch.float32) if log_std is None else log_std.detach().float().cpu().contiguous()
    host = torch.zeros(L.d2d_policy_weights_bytes(), dtype=torch.uint8)
    _lib.check(L.d2d_policy_pack(rows[0].data_ptr(), rows[1].data_ptr(), rows[2].data_ptr(), rows[3].data_ptr(), w2p.data_ptr(),
                                 b2p.data_ptr(), ls.data_ptr(), host.data_ptr()))
    return host.to(device)


class RolloutCollector:
    """SB3 `collect_rollouts` + `compute_returns_and_advantage` on the device, T steps of N environments per call,
    through ONE C call (`d2d_rollout_collect`): per step the tcgen05 policy kernel (sample, unclipped action,
    log-probability) and the step kernel, which write straight into the [T, N] tensors below; then one launch of the
    same tcgen05 kernel evaluates the critic on all T + 1 stored observations, and one GAE launch.  No torch kernel and
    no host work between two steps.  `policy` (MlpActorCritic) only supplies the weights."""

    def __init__(self, env, policy, n_steps=128, gamma=0.99, gae_lambda=0.95, seed=0):
        self.env, self.policy, self.T = env, policy, int(n_steps)
        self.gamma, self.lam = float(gamma), float(gae_lambda)
        n, dev, T = env.num_envs, env.device, self.T
        self._L = _lib.load()
        self.obs_buf = torch.empty((T + 1, n, 27), dtype=torch.float32, device=dev)
        self.actions = torch.empty((T, n, 2), dtype=torch.float32, device=dev)        # unclipped, as SB3 stores them
        self.log_probs = torch.empty((T, n), dtype=torch.float32, device=dev)
        self.rewards = torch.empty((T, n), dtype=torch.float32, device=dev)
        self.dones = torch.empty((T + 1, n), dtype=torch.uint8, device=dev)
        self.values_buf = torch.empty((T + 1, n), dtype=torch.float32, device=dev)
        self.advantages = torch.empty((T, n), dtype=torch.float32, device=dev)
        self.returns = torch.empty((T, n), dtype=torch.float32, device=dev)
        self._scratch = torch.empty((n, 2), dtype=torch.float32, device=dev)
        p = policy
        self.actor_packed = pack_tower(p.w0, p.b0, p.w1, p.b1, p.w2, p.b2, p.log_std, dev)
        self.critic_packed = pack_tower(p.v0, p.vb0, p.v1, p.vb1, p.v2, p.vb2, None, dev)
        self.obs_buf[0].copy_(env.reset_device())
        self.dones[0].fill_(1)
        self.policy_step = int(seed) << 20
        self.num_timesteps = 0

    # SB3 RolloutBuffer views
    obs = property(lambda self: self.obs_buf[:self.T])
    episode_starts = property(lambda self: self.dones[:self.T])
    values = property(lambda self: self.values_buf[:self.T])
    last_values = property(lambda self: self.values_buf[self.T])
    last_dones = property(lambda self: self.dones[self.T])
    last_obs = property(lambda self: self.obs_buf[self.T])

    def collect(self):
        env, T = self.env, self.T
        if self.num_timesteps:
            self.obs_buf[0].copy_(self.obs_buf[T])                 # device-to-device copies between two rollouts
            self.dones[0].copy_(self.dones[T])
        st = C.c_void_p(torch.cuda.current_stream(env.device).cuda_stream)
        _lib.check(self._L.d2d_rollout_collect(
            env._h, self.actor_packed.data_ptr(), self.critic_packed.data_ptr(), T, self.obs_buf.data_ptr(),
            self.actions.data_ptr(), self.log_probs.data_ptr(), self.rewards.data_ptr(), self.dones.data_ptr(),
            self.values_buf.data_ptr(), self.advantages.data_ptr(), self.returns.data_ptr(), self._scratch.data_ptr(),
            self.gamma, self.lam, self.policy_step, st))
        self.policy_step += T
        self.num_timesteps += T * env.num_envs
        env._host_stepped = False
        return self
