"""Renderer adapter, host side: raster primitives, scene composition on a hand-made state,
PNG / APNG writers.  No GPU needed: draw_scene only takes arrays."""
import importlib
import struct
import types
import zlib

import numpy as np

PKG = "drone_2d_custom_gym_env_for_reinforcement_learning_b200"
R = importlib.import_module(PKG + ".render")
scen = importlib.import_module(PKG + ".scenarios")


def _cfg():
    return types.SimpleNamespace(screen_x=1300.0, screen_y=1300.0, width=100.0, height=20.0, force_scale=1000.0)


def test_canvas_primitives_and_y_flip():
    cv = R.Canvas(100, 100)
    cv.disc((20, 80), 5, (255, 0, 0))                       # world y = 80 -> screen row 20
    assert tuple(cv.img[20, 20]) == (255, 0, 0) and tuple(cv.img[80, 20]) == R.BACKGROUND
    assert abs(int((cv.img[:, :, 1] == 0).sum()) - np.pi * 25) < 12
    cv.segment((0, 50), (99, 50), (0, 255, 0), 4)
    col = cv.img[:, 50]
    assert int((col[:, 1] == 255).sum() - (col[:, 0] == 243).sum() * 0) >= 4 and tuple(cv.img[50, 50]) == (0, 255, 0)
    cv.polygon([(60, 10), (90, 10), (90, 30), (60, 30)], (1, 2, 3))
    assert tuple(cv.img[80, 75]) == (1, 2, 3) and int((cv.img == (1, 2, 3)).all(-1).sum()) in range(560, 660)
    half = R.Canvas(1300, 1300, pixels=650)
    assert half.img.shape == (650, 650, 3) and half.scale == 0.5


def test_scene_composition_colours_and_order():
    cfg = _cfg()
    wps, us, xp, yp, obstacles = scen.scenario_tables("corridor", 1300, 1300)
    path = scen.PathTables(wps)
    pts = [tuple(path.point(u)) for u in np.linspace(0, path.length, 200)]
    state = np.zeros(36)
    x, y, a = 300.0, 700.0, 0.3
    c, s = np.cos(a), np.sin(a)
    state[0:6] = (x, y, 60.0, -20.0, a, 0.0)
    state[6:12] = (x - 40 * c, y - 40 * s, 0, 0, a, 0)
    state[12:18] = (x + 40 * c, y + 40 * s, 0, 0, a, 0)
    obs = np.zeros(27)
    closest, la = path.point(150.0), path.point(370.0)
    obs[19], obs[20] = 2 * closest[0] / 1300 - 1, 2 * closest[1] / 1300 - 1
    obs[21], obs[22] = 2 * la[0] / 1300 - 1, 2 * la[1] / 1300 - 1
    img = R.draw_scene(cfg, wps, obstacles, pts, state, obs, action=np.array([1.0, -1.0], np.float32),
                       flight_path=[(100.0 + k, 1300 - 650.0) for k in range(50)], spawn_rect=(50, 150, 200, 1000))
    assert img.shape == (1300, 1300, 3) and img.dtype == np.uint8

    def px(wx, wy):
        return tuple(int(v) for v in img[int(round(1300 - wy)), int(round(wx))])

    assert px(5, 5) == R.BACKGROUND
    assert px(x, y) == R.FRAME_COLOR                                   # the frame box is drawn over the vectors
    assert px(x - 40 * c, y - 40 * s) in (R.MOTOR_COLOR, R.FORCE_SCALE_COLOR, R.FORCE_COLOR)
    assert px(x - 40 * c + 8 * c, y - 40 * s + 8 * s - 0.0) == R.MOTOR_COLOR
    ox, oy, orad = obstacles[0]
    assert px(ox, oy + orad - 3) == R.OBSTACLE_COLOR
    assert px(*la) == R.LOOKAHEAD_COLOR
    assert px(wps[-1][0], wps[-1][1]) == R.TARGET_COLOR                # target dot over the black end waypoint
    # full-thrust left motor: red vector of the full scale length; idle right motor: only the grey scale bar
    lx, ly = x - 40 * c - s * 40, y - 40 * s + c * 40
    rx, ry = x + 40 * c - s * 40, y + 40 * s + c * 40
    assert px(lx, ly) == R.FORCE_COLOR and px(rx, ry) == R.FORCE_SCALE_COLOR
    assert px(120.0, 650.0) in (R.FLIGHT_PATH_COLOR, R.PATH_COLOR)


def test_png_and_apng_files(tmp_path):
    rng = np.random.RandomState(0)
    img = rng.randint(0, 255, (37, 53, 3)).astype(np.uint8)
    p = str(tmp_path / "a.png")
    R.write_png(p, img)
    assert np.array_equal(R.read_png(p), img)
    frames = [img, img[::-1].copy(), img[:, ::-1].copy()]
    q = str(tmp_path / "a.apng")
    R.write_apng(q, frames, fps=30)
    assert np.array_equal(R.read_png(q), img)                          # a plain PNG reader sees the first frame
    b = open(q, "rb").read()
    tags, off = [], 8
    while off < len(b):
        n, tag = struct.unpack(">I", b[off:off + 4])[0], b[off + 4:off + 8]
        assert zlib.crc32(tag + b[off + 8:off + 8 + n]) & 0xFFFFFFFF == struct.unpack(">I", b[off + 8 + n:off + 12 + n])[0]
        tags.append(tag)
        off += 12 + n
    assert tags == [b"IHDR", b"acTL", b"fcTL", b"IDAT", b"fcTL", b"fdAT", b"fcTL", b"fdAT", b"IEND"]
