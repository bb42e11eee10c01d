"""Configuration surface of the reference (its rl_config.py:5-79 exports `rl_config`,
`env_train_config` and `env_test_config`; scripts do `from rl_config import *`).

The three dictionaries are rebuilt here from one annotated table so that every key of
the reference keeps its name and default, grouped by what the batched environment does
with it.  Keys marked "dead" are accepted for compatibility and ignored, exactly as the
reference ignores them (drone_2d_env.py:917-946 is never called; rendering is out of scope).
"""
import math

# (key, default, role)                                               reference line
_ENV_KEYS = (
    # --- rendering (GUI, out of scope: accepted and ignored) ---
    ("render_sim", False, "dead"),                                   # rl_config.py:11
    ("render_path", False, "dead"),
    ("render_shade", False, "dead"),
    ("render_text", False, "dead"),
    ("shade_distance", 75, "dead"),
    # --- episode ---
    ("n_steps", 1100, "termination: time limit"),                    # :16
    ("n_fall_steps", 5, "dead"),
    ("change_target", False, "dead"),
    ("initial_throw", True, "dead"),
    # --- curriculum path sampling ---
    ("random_path_spawn", True, "reset: random start corner"),       # :20
    ("path_segment_length", 100, "reset: waypoint spacing"),
    ("n_wps", 12, "reset: waypoints per path"),
    ("screensize_x", 1300, "world width (normalisation)"),
    ("screensize_y", 1300, "world height (normalisation)"),
    ("lookahead", 220, "observation: look-ahead distance along the path"),
    ("spawn_corners", (1, 4), "reset: corner range DL, DR, UL, UR = 1..4"),
    # --- reward shaping ---
    ("danger_range", 150, "reward: collision avoidance range"),      # :27
    ("danger_angle", 20, "reward: collision avoidance angle [deg]"),
    ("abs_inv_CA_min_rew", 1 / 8, "reward: inverse of the CA floor"),
    ("PA_band_edge", 40, "reward: path adherence band"),
    ("PA_scale", 2, "reward: path adherence scale"),
    ("PP_vel_scale", 0.08, "reward: path progression velocity scale"),
    ("PP_rew_max", 2.5, "reward: path progression clip"),
    ("PP_rew_min", -1, "reward: path progression clip"),
    ("rew_collision", -50, "reward + termination: collision"),
    ("reach_end_radius", 20, "termination: goal square half-size"),
    ("rew_reach_end", 30, "reward: goal"),
    ("AA_angle", math.pi / 2, "termination: tilt limit"),
    ("AA_band", math.pi / 4, "reward: tilt penalty band"),
    ("rew_AA", -1, "reward: tilt limit"),
    ("use_Lambda", True, "reward: blend PA / CA near obstacles"),
    # --- what to simulate ---
    ("mode", "test", "'curriculum' (per-env random paths) or 'test' (shared scenario)"),   # :42
    ("scenario", "large", "test scenario name, 'stage_1'..'stage_5', or 'None'"),
)

TEST_SCENARIOS = ("perpendicular", "parallel", "S_parallel", "corridor", "S_corridor", "large", "impossible")
STAGES = tuple("stage_%d" % k for k in range(1, 6))

rl_config = {"total_timesteps": 9000000, "ent_coef": 0.01}          # rl_config.py:5-8 (SB3 PPO settings)

env_train_config = {key: default for key, default, _ in _ENV_KEYS}

# rl_config.py:63-71: the "visualisation" variant the reference's test mode uses
env_test_config = dict(env_train_config, render_sim=True, render_path=True, render_shade=True, render_text=True,
                       initial_throw=False, n_fall_steps=0)


def describe(key):
    """What the batched environment does with a configuration key."""
    for k, _, role in _ENV_KEYS:
        if k == key:
            return role
    raise KeyError(key)


def stage_from_timesteps(sim_num):
    """(stage, obstacle spawn chance) for a training progress of `sim_num` environment steps.

    The reference derives `sim_num` from the newest `logs/rl_model_<n>_steps.zip` each time an
    environment is rebuilt (drone_2d_env.py:79-85) and switches at 0.7 M / 1.0 M / 1.6 M / 2.0 M
    steps, ramping the obstacle probability 0.2 -> 0.6 in stage 3 and 0.6 -> 1 in stage 4
    (drone_2d_env.py:326-372).  Here the trainer passes the progress explicitly."""
    bounds = ((700000, 1, None), (1000000, 2, None), (1600000, 3, (1000000, 0.2, 0.6)), (2000000, 4, (1600000, 0.6, 1.0)))
    for upper, stage, ramp in bounds:
        if sim_num < upper:
            if ramp is None:
                return stage, 0.0
            start, lo, hi = ramp
            return stage, (sim_num - start) * (hi - lo) / (upper - start) + lo
    return 5, 1.0
