"""Renderer adapter on live environments."""
import importlib
import os

import numpy as np
import pytest

from conftest import GOLDEN

pytestmark = pytest.mark.gpu
PKG = "drone_2d_custom_gym_env_for_reinforcement_learning_b200"


def _px(img, x, y, H=1300.0, scale=1.0):
    return tuple(int(v) for v in img[int(round((H - y) * scale)), int(round(x * scale))])


def test_vec_env_render_draws_the_live_state():
    pkg = importlib.import_module(PKG)
    R = importlib.import_module(PKG + ".render")
    for kw in (dict(mode="test", scenario="large"), dict(mode="curriculum", scenario="stage_5")):
        env = pkg.Drone2dVecEnv(8, seed=3, info_mode="full", **kw)
        env.reset()
        import torch
        for _ in range(5):
            env.step_device(torch.full((8, 2), 0.2, device=env.device))
        for i in (0, 7):
            img = env.render(index=i)
            st = env.get_state()[i]
            assert img.shape == (1300, 1300, 3)
            assert _px(img, st[0], st[1]) == R.FRAME_COLOR
            tabs = env.scenario["obstacles"] if kw["mode"] == "test" else env.get_tables()["obstacles"][i][:env.get_tables()["n_obs"][i]]
            for ox, oy, r in tabs:
                if 0 <= ox < 1300 and 0 <= oy < 1300 and np.hypot(ox - st[0], oy - st[1]) > r + 80:
                    assert _px(img, ox, oy) in (R.OBSTACLE_COLOR, R.NEAREST_OBS_COLOR, R.LOOKAHEAD_COLOR, R.PATH_COLOR,
                                                R.CLOSEST_COLOR, R.TARGET_COLOR)
        small = env.get_images(indices=[0, 1], pixels=260)
        assert len(small) == 2 and small[0].shape == (260, 260, 3)
        env.close()


def test_record_episode_writes_an_animation(tmp_path):
    pkg = importlib.import_module(PKG)
    R = importlib.import_module(PKG + ".render")
    policy = importlib.import_module(PKG + ".policy")
    env = pkg.Drone2dVecEnv(4, seed=2, info_mode="full", mode="test", scenario="corridor")
    actor = policy.MlpActor.load(os.path.join(GOLDEN, "agent_19_policy.npz"), env.device)
    out = str(tmp_path / "corridor.apng")
    frames, flight = R.record_episode(env, actor, index=1, path=out, pixels=325, every=8, max_steps=240, seed=1)
    env.close()
    assert len(flight) >= 100 and len(frames) == (len(flight) + 7) // 8
    assert os.path.getsize(out) > 1000 and R.read_png(out).shape == (325, 325, 3)
    # the drone moved along the corridor between the first and the last frame
    assert flight[-1][0] > flight[0][0] + 100
    assert not np.array_equal(frames[0], frames[-1])


def test_single_env_records_the_flight_path_like_the_reference():
    pkg = importlib.import_module(PKG)
    env = pkg.Drone2dEnv(seed=4, mode="test", scenario="perpendicular", render_sim=True, render_path=True, n_steps=60)
    env.reset()
    done, n = False, 0
    while not done:
        obs, r, done, info = env.step(np.array([0.1, 0.1], np.float32))
        n += 1
    fp = info["flight_path"]                                   # env:612: handed out with the terminal info
    assert len(fp) == n == info["env_steps"]
    st = env._vec.get_state()[0]
    assert abs(fp[-1][0] - st[0]) < 1e-9 and abs(fp[-1][1] - (1300 - st[1])) < 1e-9
    img = env.render()
    assert img is not None and img.shape == (1300, 1300, 3)
    env.close()
    quiet = pkg.Drone2dEnv(seed=4, mode="test", scenario="perpendicular")
    quiet.reset()
    assert quiet.render() is None                              # env:776: render_sim False -> nothing
    quiet.close()
