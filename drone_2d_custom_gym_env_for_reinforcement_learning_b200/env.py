"""Single-environment adapter with the reference's gym.Env surface
(`Drone2dEnv(**kwargs)`, reset() -> obs, step(action) -> (obs, reward, done,
info); drone_2d_env.py:33,394,908).  It drives a 1-env Drone2dVecEnv, so every
transition is still the CUDA kernel; observations come from the float64 tap and
therefore have the reference's dtype (float64, env:765)."""
import numpy as np

from .vec_env import Drone2dVecEnv, INFO_KEYS


class Drone2dEnv:
    metadata = {"render_modes": []}

    def __init__(self, device=None, seed=0, **kwargs):
        self.kwargs = kwargs
        self._vec = Drone2dVecEnv(1, device=device, seed=seed, info_mode="full", **kwargs)
        self.observation_space = self._vec.observation_space
        self.action_space = self._vec.action_space
        self.done = False
        self.info = {}

    def reset(self):
        """env:908-912: everything is re-initialised; returns the first observation."""
        self._vec.reset_device()
        self.done = False
        self.info = {k: 0 for k in INFO_KEYS}
        return self._vec.obs64.cpu().numpy()[0].copy()

    def step(self, action):
        a = np.asarray(action, dtype=np.float32).reshape(1, 2)
        import torch
        self._vec.step_device(torch.from_numpy(a).to(self._vec.device), auto_reset=False)
        obs = self._vec.obs64.cpu().numpy()[0].copy()
        row = self._vec.info.cpu().numpy()[0]
        done = bool(self._vec.dones.cpu().numpy()[0])
        info = self._vec._info_row(row, done)
        if info["dist_closest_obs"] == float("inf"):
            info["dist_closest_obs"] = np.inf
        self.done = self.done or done
        self.info = info
        return obs, float(row[0]), self.done, info

    def render(self, mode="human", close=False):
        return None

    def close(self):
        self._vec.close()
