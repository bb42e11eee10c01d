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
            info = env.info[newly]
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
