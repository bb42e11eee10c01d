"""B200-native batched Drone2dEnv step.

Public surface:
  Drone2dVecEnv   SB3-VecEnv-shaped, device-resident batch of environments
  Drone2dEnv      single-env adapter with the reference's gym.Env methods
  rl_config       the reference's configuration dictionaries
  scenarios       host-side table builders for the shared test scenarios
"""
from . import rl_config, scenarios  # noqa: F401
from ._build import build  # noqa: F401
from ._lib import D2DError, load as load_library  # noqa: F401
from .env import Drone2dEnv  # noqa: F401
from .vec_env import Drone2dVecEnv, config_from_kwargs  # noqa: F401

__all__ = ["Drone2dVecEnv", "Drone2dEnv", "rl_config", "scenarios", "build", "load_library", "D2DError",
           "config_from_kwargs"]
