"""Renderer adapter: one environment's state is copied to the host and drawn with the
reference's scene composition (`Drone2dEnv.render`, drone_2d_env.py:775-906) -- same elements, same
colours, same z-order -- into a numpy RGB array.  pygame is not a dependency: a few numpy raster
primitives stand in for `pygame.draw.*` / `space.debug_draw`, and PNG / animated-PNG files are written
with zlib (the reference writes GIFs through imageio, main.py:266-295).

Not on the hot path: nothing here runs unless a caller asks for a frame.  What is NOT drawn: the reward
text overlay (needs a font), the three bearing arcs and the drone "shade" sprites (img/shade.png).
"""
import struct
import zlib

import numpy as np

from .scenarios import PathTables, spawn_rectangle

BACKGROUND = (243, 243, 243)          # env:779
PATH_COLOR = (0, 0, 0)                # env:823-830
CLOSEST_COLOR = (0, 0, 255)           # env:841
LOOKAHEAD_COLOR = (0, 150, 150)       # env:847-849
VELOCITY_COLOR = (0, 0, 0)            # env:857-858 (draw_red_velocity False)
NEAREST_OBS_COLOR = (0, 255, 0)       # env:863-864
FRAME_COLOR = (66, 135, 245)          # Drone.py:25
MOTOR_COLOR = (33, 93, 191)           # Drone.py:42,56
OBSTACLE_COLOR = (188, 72, 72)        # obstacles.py:60
FORCE_SCALE_COLOR = (179, 179, 179)   # env:881
FORCE_COLOR = (255, 0, 0)             # env:884
TARGET_COLOR = (255, 0, 0)            # env:893
FLIGHT_PATH_COLOR = (16, 19, 97)      # env:900


class Canvas:
    """RGB raster with the reference's screen convention: world (x, y) is drawn at pixel
    (x, screen_height - y) (env:821,986), optionally down-scaled."""

    def __init__(self, world_w, world_h, pixels=None, background=BACKGROUND):
        self.scale = 1.0 if pixels is None else float(pixels) / float(world_w)
        self.w = int(round(world_w * self.scale))
        self.h = int(round(world_h * self.scale))
        self.world_h = float(world_h)
        self.img = np.empty((self.h, self.w, 3), dtype=np.uint8)
        self.img[:] = background

    def to_px(self, p):
        return np.array([p[0] * self.scale, (self.world_h - p[1]) * self.scale], dtype=np.float64)

    def _window(self, lo, hi):
        x0, y0 = max(int(np.floor(lo[0])), 0), max(int(np.floor(lo[1])), 0)
        x1, y1 = min(int(np.ceil(hi[0])) + 1, self.w), min(int(np.ceil(hi[1])) + 1, self.h)
        if x0 >= x1 or y0 >= y1:
            return None
        ys, xs = np.mgrid[y0:y1, x0:x1]
        return (slice(y0, y1), slice(x0, x1)), xs + 0.5, ys + 0.5

    def disc(self, centre, radius, color):
        c, r = self.to_px(centre), max(radius * self.scale, 0.75)
        win = self._window(c - r, c + r)
        if win is None:
            return
        sl, xs, ys = win
        self.img[sl][(xs - c[0]) ** 2 + (ys - c[1]) ** 2 <= r * r] = color

    def segment(self, a, b, color, width=1.0):
        a, b = self.to_px(a), self.to_px(b)
        half = max(width * self.scale, 1.0) / 2.0
        win = self._window(np.minimum(a, b) - half, np.maximum(a, b) + half)
        if win is None:
            return
        sl, xs, ys = win
        d = b - a
        L2 = float(d @ d)
        t = np.zeros_like(xs) if L2 == 0.0 else np.clip(((xs - a[0]) * d[0] + (ys - a[1]) * d[1]) / L2, 0.0, 1.0)
        dist2 = (xs - (a[0] + t * d[0])) ** 2 + (ys - (a[1] + t * d[1])) ** 2
        self.img[sl][dist2 <= half * half] = color

    def polyline(self, pts, color, width=1.0):
        for a, b in zip(pts[:-1], pts[1:]):
            self.segment(a, b, color, width)

    def polygon(self, pts, color):
        """Filled convex polygon (world coordinates, either winding)."""
        P = np.array([self.to_px(p) for p in pts])
        win = self._window(P.min(0), P.max(0))
        if win is None:
            return
        sl, xs, ys = win
        pos = np.ones(xs.shape, bool)
        neg = np.ones(xs.shape, bool)
        for a, b in zip(P, np.roll(P, -1, axis=0)):
            cross = (b[0] - a[0]) * (ys - a[1]) - (b[1] - a[1]) * (xs - a[0])
            pos &= cross >= 0
            neg &= cross <= 0
        self.img[sl][pos | neg] = color

    def rect_outline(self, x, y_top_screen, w, h, color, width=2.0):
        """pygame.Rect semantics: (x, y) is the top-left corner in SCREEN coordinates (env:833)."""
        y1 = self.world_h - y_top_screen
        y0 = y1 - h
        self.polyline([(x, y0), (x + w, y0), (x + w, y1), (x, y1), (x, y0)], color, width)


def _box(body, half_w, half_h):
    x, y, a = body[0], body[1], body[4]
    c, s = np.cos(a), np.sin(a)
    return [(x + c * lx - s * ly, y + s * lx + c * ly)
            for lx, ly in ((-half_w, -half_h), (half_w, -half_h), (half_w, half_h), (-half_w, half_h))]


def _local_to_world(body, lx, ly):
    c, s = np.cos(body[4]), np.sin(body[4])
    return (body[0] + c * lx - s * ly, body[1] + s * lx + c * ly)


def draw_scene(cfg, wps, obstacles, path_points, state_row, obs_row, action=None, flight_path=None,
               spawn_rect=None, pixels=None):
    """Compose one frame.  `state_row`: the 36-double state blob of the env (include/drone2d.h),
    `obs_row`: its 27 observations, `action`: the last action (for the thrust vectors, env:878-891),
    `flight_path`: [(x, screen_height - y), ...] as the reference records it (env:984-986)."""
    W, H = float(cfg.screen_x), float(cfg.screen_y)
    cv = Canvas(W, H, pixels)
    frame, left, right = state_row[0:6], state_row[6:12], state_row[12:18]
    pos = (frame[0], frame[1])

    cv.polyline(path_points, PATH_COLOR, 1.0)                                         # env:821-823
    cv.disc(wps[0], 5, PATH_COLOR)
    cv.disc(wps[-1], 5, PATH_COLOR)
    if spawn_rect is not None:
        cv.rect_outline(*spawn_rect, PATH_COLOR, 2.0)                                 # env:831-833

    closest = ((obs_row[19] + 1) * W / 2, (obs_row[20] + 1) * H / 2)                   # obs 19-22 are m1to1 of the points
    lookahead = ((obs_row[21] + 1) * W / 2, (obs_row[22] + 1) * H / 2)
    cv.disc(closest, 5, CLOSEST_COLOR)                                                # env:841
    cv.segment(pos, lookahead, LOOKAHEAD_COLOR, 4.0)                                  # env:847
    cv.disc(lookahead, 5, LOOKAHEAD_COLOR)
    cv.segment(pos, (pos[0] + frame[2], pos[1] + frame[3]), VELOCITY_COLOR, 4.0)      # env:857
    if len(obstacles):                                                                # env:861-864
        corners = np.array([(pos[0] + sx * cfg.width / 2, pos[1] + sy * cfg.height / 4) for sx in (-1, 1) for sy in (-1, 1)])
        ob = np.asarray(obstacles, dtype=np.float64)
        d = np.sqrt(((corners[None, :, :] - ob[:, None, :2]) ** 2).sum(-1)).min(1) - ob[:, 2]   # env:948-961
        k = int(np.argmin(d))
        cv.segment(pos, (ob[k, 0], ob[k, 1]), NEAREST_OBS_COLOR, 4.0)

    for x, y, r in obstacles:                                                         # space.debug_draw, env:875
        cv.disc((x, y), r, OBSTACLE_COLOR)
    cv.polygon(_box(frame, cfg.width / 2, cfg.height / 4), FRAME_COLOR)                # Drone.py:16 (width, height/2)
    cv.polygon(_box(left, cfg.height / 2, cfg.height / 2), MOTOR_COLOR)                # Drone.py:33,47 (height, height)
    cv.polygon(_box(right, cfg.height / 2, cfg.height / 2), MOTOR_COLOR)

    R = cfg.width / 2 - cfg.height / 2                                                 # drone_radius (Drone.py:11)
    vec = 0.05                                                                        # env:878
    for side, k in ((-R, 0), (R, 1)):
        root = _local_to_world(frame, side, 0.0)
        cv.segment(root, _local_to_world(frame, side, cfg.force_scale * vec), FORCE_SCALE_COLOR, 4.0)
        if action is not None:
            f = (float(action[k]) / 2 + 0.5) * cfg.force_scale                        # env:400-401
            cv.segment(root, _local_to_world(frame, side, f * vec), FORCE_COLOR, 4.0)
    cv.disc(wps[-1], 5, TARGET_COLOR)                                                 # env:893 (target = last waypoint)
    if flight_path is not None and len(flight_path) > 2:                              # env:899-900
        cv.polyline([(p[0], H - p[1]) for p in flight_path], FLIGHT_PATH_COLOR, 1.0)
    return cv.img


class FrameRenderer:
    """Pulls what `draw_scene` needs out of a Drone2dVecEnv (state blob, tables, observation)."""

    def __init__(self, env, pixels=None, path_samples=400):
        self.env = env
        self.pixels = pixels
        self.path_samples = int(path_samples)
        self._path_cache = {}

    def _tables(self, index):
        env = self.env
        if getattr(env, "scenario", None) is not None:
            sc = env.scenario
            return sc["wps"], sc["obstacles"], "shared"
        t = env.get_tables()
        n_wps = int(env.cfg.n_wps)
        return t["wps"][index, :n_wps], t["obstacles"][index, :int(t["n_obs"][index])], None

    def frame(self, index=0, action=None, flight_path=None):
        env = self.env
        wps, obstacles, key = self._tables(index)
        cache_key = key or (index, wps.tobytes())
        if cache_key not in self._path_cache:
            if key is None:
                self._path_cache.clear()
            path = PathTables(wps)
            self._path_cache[cache_key] = [tuple(path.point(u)) for u in np.linspace(0.0, path.length, self.path_samples)]
        state = env.get_state()[index]
        obs = env.current_obs_row(index)
        rect = None
        if key == "shared" and env.kwargs.get("scenario") not in (None, "None"):
            # env:227-231: pygame.Rect(xmin, ymin, xmax - xmin, ymax) in SCREEN coordinates (not flipped, and the
            # height is ymax rather than ymax - ymin) -- reproduced as drawn
            xmin, xmax, ymin, ymax = spawn_rectangle(env.kwargs["scenario"], env.cfg.screen_x, env.cfg.screen_y)
            rect = (xmin, ymin, xmax - xmin, ymax)
        return draw_scene(env.cfg, wps, obstacles, self._path_cache[cache_key], state, obs, action=action,
                          flight_path=flight_path, spawn_rect=rect, pixels=self.pixels)


# ---------------------------------------------------------------------------------- files
def _chunk(tag, data):
    return struct.pack(">I", len(data)) + tag + data + struct.pack(">I", zlib.crc32(tag + data) & 0xFFFFFFFF)


def _idat_payload(rgb):
    h, w, _ = rgb.shape
    raw = np.empty((h, 1 + 3 * w), dtype=np.uint8)
    raw[:, 0] = 0                                   # filter type 0 on every scanline
    raw[:, 1:] = rgb.reshape(h, 3 * w)
    return zlib.compress(raw.tobytes(), 6)


def write_png(path, rgb):
    rgb = np.ascontiguousarray(rgb, dtype=np.uint8)
    h, w, _ = rgb.shape
    with open(path, "wb") as f:
        f.write(b"\x89PNG\r\n\x1a\n" + _chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, 2, 0, 0, 0)) +
                _chunk(b"IDAT", _idat_payload(rgb)) + _chunk(b"IEND", b""))


def write_apng(path, frames, fps=30):
    """Animated PNG (every browser plays it): the counterpart of the reference's imageio GIF (main.py:295)."""
    frames = [np.ascontiguousarray(f, dtype=np.uint8) for f in frames]
    h, w, _ = frames[0].shape
    out = [b"\x89PNG\r\n\x1a\n", _chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, 2, 0, 0, 0)),
           _chunk(b"acTL", struct.pack(">II", len(frames), 0))]
    seq = 0
    for k, fr in enumerate(frames):
        out.append(_chunk(b"fcTL", struct.pack(">IIIIIHHBB", seq, w, h, 0, 0, 1, int(fps), 0, 0)))
        seq += 1
        data = _idat_payload(fr)
        if k == 0:
            out.append(_chunk(b"IDAT", data))
        else:
            out.append(_chunk(b"fdAT", struct.pack(">I", seq) + data))
            seq += 1
    out.append(_chunk(b"IEND", b""))
    with open(path, "wb") as f:
        f.write(b"".join(out))


def read_png(path):
    """Decoder for the files written above (8-bit RGB, filter 0, no interlace); first frame only."""
    b = open(path, "rb").read()
    assert b[:8] == b"\x89PNG\r\n\x1a\n"
    off, w, h, data = 8, None, None, b""
    while off < len(b):
        n, tag = struct.unpack(">I", b[off:off + 4])[0], b[off + 4:off + 8]
        body = b[off + 8:off + 8 + n]
        if zlib.crc32(tag + body) & 0xFFFFFFFF != struct.unpack(">I", b[off + 8 + n:off + 12 + n])[0]:
            raise ValueError("bad CRC in chunk %r" % tag)
        if tag == b"IHDR":
            w, h = struct.unpack(">II", body[:8])
        elif tag == b"IDAT":
            data += body
        off += 12 + n
    raw = np.frombuffer(zlib.decompress(data), dtype=np.uint8).reshape(h, 1 + 3 * w)
    return raw[:, 1:].reshape(h, w, 3).copy()


def record_episode(env, actor, index=0, path=None, pixels=650, every=2, max_steps=None, seed=0, deterministic=False):
    """The reference's GIF mode (main.py:255-295) for env `index`: run the actor until that env is done,
    grab every `every`-th frame (the reference keeps every 2nd, at 30 fps), optionally write an APNG."""
    import torch
    g = torch.Generator(device=env.device)
    g.manual_seed(seed)
    rend = FrameRenderer(env, pixels=pixels)
    H = env.cfg.screen_y
    obs = env.reset_device()
    frames, flight = [], []
    for t in range(max_steps or env.cfg.n_steps):
        act = actor.predict(obs, deterministic, g)
        obs, _, done = env.step_device(act, auto_reset=False, terminal_obs=False)
        st = env.get_state()[index]
        flight.append((st[0], H - st[1]))
        if t % every == 0:
            frames.append(rend.frame(index, action=act[index].cpu().numpy(), flight_path=flight))
        if bool(done[index]):
            break
    if path is not None:
        write_apng(path, frames, fps=30)
    return frames, flight
