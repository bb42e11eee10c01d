"""Single-environment adapter with the reference's gym.Env surface
(`Drone2dEnv(**kwargs)`, reset() -> obs, step(action) -> (obs, reward, done,
info); drone_2d_env.py:33,394,908).  It drives a 1-env Drone2dVecEnv, so every
transition is still the CUDA kernel; observations come from the float64 tap and
therefore have the reference's dtype (float64, env:765)."""
import numpy as np

from .vec_env import Drone2dVecEnv, INFO_KEYS


class Drone2dEnv:
    metadata = {"render_modes": ["rgb_array"]}

    def __init__(self, device=None, seed=0, **kwargs):
        self.kwargs = kwargs
        self._vec = Drone2dVecEnv(1, device=device, seed=seed, info_mode="full", **kwargs)
        self.observation_space = self._vec.observation_space
        self.action_space = self._vec.action_space
        self.done = False
        self.info = {}
        self._flight_path = []

    def reset(self):
        """env:908-912: everything is re-initialised; returns the first observation."""
        self._vec.reset_device()
        self.done = False
        self.info = {k: 0 for k in INFO_KEYS}
        self._flight_path = []
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
        if self.kwargs.get("render_path", False):
            # env:410-415,984-986: (x, screen_height - y) of the frame after every step, handed out at the end
            W, H = self._vec.cfg.screen_x, self._vec.cfg.screen_y
            self._flight_path.append(((obs[6] + 1.0) * W / 2.0, H - (obs[7] + 1.0) * H / 2.0))
            if done:
                info["flight_path"] = self._flight_path
        self.done = self.done or done
        self.info = info
        return obs, float(row[0]), self.done, info

    def render(self, mode="human", close=False):
        """env:775-776: nothing unless `render_sim` was requested.  With it, the frame the reference would
        put on the pygame window is returned as an RGB array instead (render.py; no window is opened)."""
        if not self.kwargs.get("render_sim", False):
            return None
        if self.kwargs.get("render_path", False):
            return self._vec.render(flight_path=self._flight_path)
        return self._vec.render()

    def close(self):
        self._vec.close()
