"""Configuration surface of the reference (rl_config.py:5-79): the same three
dictionaries with the same keys and defaults, so code written against the
reference's `from rl_config import *` keeps working."""
import numpy as np

rl_config = {
    'total_timesteps': 9000000,
    'ent_coef': 0.01,
}

env_train_config = {
    'render_sim': False,
    'render_path': False,
    'render_shade': False,
    'render_text': False,
    'shade_distance': 75,
    'n_steps': 1100,
    'n_fall_steps': 5,            # accepted, dead in the reference (env:917-946 is never called)
    'change_target': False,       # accepted, GUI only
    'initial_throw': True,        # accepted, dead in the reference
    'random_path_spawn': True,
    'path_segment_length': 100,
    'n_wps': 12,
    'screensize_x': 1300,
    'screensize_y': 1300,
    'lookahead': 220,
    'spawn_corners': (1, 4),
    'danger_range': 150,
    'danger_angle': 20,
    'abs_inv_CA_min_rew': 1 / 8,
    'PA_band_edge': 40,
    'PA_scale': 2,
    'PP_vel_scale': 0.08,
    'PP_rew_max': 2.5,
    'PP_rew_min': -1,
    'rew_collision': -50,
    'reach_end_radius': 20,
    'rew_reach_end': 30,
    'AA_angle': np.pi / 2,
    'AA_band': np.pi / 4,
    'rew_AA': -1,
    'use_Lambda': True,
    'mode': 'test',
    'scenario': 'large',
}

env_test_config = env_train_config.copy()
env_test_config['render_sim'] = True
env_test_config['render_path'] = True
env_test_config['render_shade'] = True
env_test_config['render_text'] = True
env_test_config['initial_throw'] = False
env_test_config['n_fall_steps'] = 0

TEST_SCENARIOS = ('perpendicular', 'parallel', 'S_parallel', 'corridor', 'S_corridor', 'large', 'impossible')
STAGES = ('stage_1', 'stage_2', 'stage_3', 'stage_4', 'stage_5')

# Curriculum thresholds of env:326-372 (the reference reads the training progress
# from checkpoint file names, env:79-85; here the caller passes `sim_num`).
_CURRICULUM = ((700000, 1), (1000000, 2), (1600000, 3), (2000000, 4))


def stage_from_timesteps(sim_num):
    """(stage, obstacle spawn chance) for a training progress of `sim_num` steps."""
    if sim_num < 700000:
        return 1, 0.0
    if sim_num < 1000000:
        return 2, 0.0
    if sim_num < 1600000:
        return 3, (sim_num - 1000000) * (0.6 - 0.2) / (1600000 - 1000000) + 0.2
    if sim_num < 2000000:
        return 4, (sim_num - 1600000) * (1 - 0.6) / (2000000 - 1600000) + 0.6
    return 5, 1.0
