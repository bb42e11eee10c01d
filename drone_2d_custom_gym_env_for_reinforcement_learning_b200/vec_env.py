"""Drone2dVecEnv -- the batched, device-resident replacement of the reference's
`Drone2dEnv` (drone_2d_env.py:22-162, 394-615, 908-912) behind the SB3 VecEnv
contract the reference consumes it through (`SubprocVecEnv`, main.py:88-101,190).

PyTorch is plumbing only: it owns the device memory (one persistent state
tensor + the I/O tensors) and the stream; every transition is ONE launch of the
hand-written sm_100a kernel in csrc/drone2d.cu through the C ABI of
include/drone2d.h.  There is no CPU fallback.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib
from .rl_config import STAGES, TEST_SCENARIOS, env_train_config, stage_from_timesteps
from .scenarios import scenario_tables, spawn_rectangle
from .spaces import Box

INFO_KEYS = ("reward", "collision_avoidance_reward", "path_adherence", "path_progression", "collision_reward",
             "reach_end_reward", "agressive_alpha_reward", "env_steps", "dist_closest_obs")
STATS_KEYS = ("episodes", "n_successful_runs", "n_failed_runs", "n_collisions", "env_steps", "APE", "total_reward",
              "reward", "collision_reward", "collision_avoidance_reward", "path_adherence", "path_progression",
              "reach_end_reward", "agressive_alpha_reward", "steps_taken", "inlane_resets")
_EMPTY_INFO = {}


class _DeviceArray:
    """Zero-copy view of library-owned device memory for torch.as_tensor."""

    def __init__(self, ptr, shape, typestr):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False),
                                         "version": 2, "strides": None}


def config_from_kwargs(kwargs):
    """Map the reference's kwargs (rl_config.py:10-44, consumed at env:34-66) onto
    struct d2d_config.  Unknown keys raise TypeError; the reference would raise
    KeyError for missing ones -- here missing keys take the rl_config defaults."""
    kw = dict(env_train_config)
    extra = {"sim_num": 0, "solver_iterations": None}      # solver_iterations: pymunk Space.iterations (default 10)
    for k, v in kwargs.items():
        if k in kw:
            kw[k] = v
        elif k in extra:
            extra[k] = v
        else:
            raise TypeError("unexpected keyword argument %r" % (k,))
    mode, scenario = kw["mode"], kw["scenario"]
    c = _lib.default_config()
    if mode == "curriculum":
        c.mode = _lib.MODE_CURRICULUM
        if scenario in STAGES:
            c.stage = int(scenario[-1])
            c.stage_spawn_chance = {3: 0.6, 4: 1.0}.get(c.stage, 0.0)          # env:340-341,354-355
        elif scenario in ("None", None):
            c.stage, c.stage_spawn_chance = stage_from_timesteps(int(extra["sim_num"]))   # env:79-85,326-372
        else:
            raise ValueError("mode='curriculum' needs scenario in %s or 'None', got %r" % (STAGES, scenario))
    elif mode == "test":
        if scenario not in TEST_SCENARIOS:
            # the reference crashes at env:316 (no waypoints) for stage_* in test mode
            raise ValueError("mode='test' needs scenario in %s, got %r" % (TEST_SCENARIOS, scenario))
        c.mode = _lib.MODE_TEST
        c.spawn_xmin, c.spawn_xmax, c.spawn_ymin, c.spawn_ymax = spawn_rectangle(
            scenario, kw["screensize_x"], kw["screensize_y"])
    else:
        raise ValueError("mode must be 'curriculum' or 'test', got %r" % (mode,))
    if extra["solver_iterations"] is not None:
        c.iterations = int(extra["solver_iterations"])
    c.n_steps = int(kw["n_steps"])
    c.n_wps = int(kw["n_wps"])
    c.spawn_corner_lo, c.spawn_corner_hi = int(kw["spawn_corners"][0]), int(kw["spawn_corners"][1])
    c.random_path_spawn = 1 if kw["random_path_spawn"] is True else 0
    c.use_Lambda = 1 if kw["use_Lambda"] is True else 0                       # `is True`, env:496
    c.screen_x, c.screen_y = float(kw["screensize_x"]), float(kw["screensize_y"])
    c.path_segment_length = float(kw["path_segment_length"])
    for k in ("lookahead", "danger_range", "danger_angle", "abs_inv_CA_min_rew", "PA_band_edge", "PA_scale",
              "PP_vel_scale", "PP_rew_max", "PP_rew_min", "rew_collision", "reach_end_radius", "rew_reach_end",
              "AA_angle", "AA_band", "rew_AA"):
        setattr(c, k, float(kw[k]))
    return c, kw


class Drone2dVecEnv:
    """N independent Drone2dEnv instances stepped by one kernel launch.

    SB3 VecEnv surface: num_envs, observation_space, action_space, reset(),
    step_async()/step_wait()/step(), close(), seed(), get_attr/set_attr/
    env_method/env_is_wrapped.  `step` auto-resets finished envs and puts the
    last observation of the finished episode in infos[i]["terminal_observation"].

    Device-resident surface (no host copies): reset_device(), step_device().

    info_mode: "episode" (default) -> infos carry the reference's keys only for
    envs that finished this step (all other entries are one shared empty dict: do
    not mutate it; the kernel writes info rows of finishing envs only, no float64
    observation tap); "full" -> every env, every key of env:575-587, float64
    observation tap `obs64` (slow for large N); "none" -> no taps at all (benchmarks).
    """

    metadata = {"render_modes": ["rgb_array"]}

    def __init__(self, num_envs, device=None, seed=0, env_id_base=0, info_mode="episode", copy_outputs=True, **kwargs):
        if not torch.cuda.is_available():
            raise RuntimeError("Drone2dVecEnv needs a CUDA device: there is no CPU fallback")
        if info_mode not in ("episode", "full", "none"):
            raise ValueError("info_mode must be 'episode', 'full' or 'none'")
        self._L = _lib.load()
        self.num_envs = int(num_envs)
        self.device = torch.device(device if device is not None else "cuda:%d" % torch.cuda.current_device())
        self.cfg, self.kwargs = config_from_kwargs(kwargs)
        self.info_mode = info_mode
        self.seed_value = int(seed)
        self.env_id_base = int(env_id_base)
        n = self.num_envs
        self.observation_space = Box(low=np.full(27, -1, dtype=np.float32), high=np.full(27, 1, dtype=np.float32),
                                     dtype=np.float32)             # env:160-162 (values are NOT clipped)
        self.action_space = Box(low=np.array([-1, -1], dtype=np.float32), high=np.array([1, 1], dtype=np.float32),
                                dtype=np.float32)                  # env:155-157
        nbytes = C.c_size_t()
        _lib.check(self._L.d2d_state_bytes(C.byref(self.cfg), n, C.byref(nbytes)))
        dev = self.device
        self.state = torch.zeros(nbytes.value, dtype=torch.uint8, device=dev)     # the persistent SoA state
        self.obs = torch.zeros((n, 27), dtype=torch.float32, device=dev)
        self.rewards = torch.zeros(n, dtype=torch.float32, device=dev)
        self.dones = torch.zeros(n, dtype=torch.uint8, device=dev)
        self.terminal_obs = torch.zeros((n, 27), dtype=torch.float32, device=dev)
        self.actions = torch.zeros((n, 2), dtype=torch.float32, device=dev)
        self._h = C.c_void_p()
        _lib.check(self._L.d2d_create(C.byref(self.cfg), n, self.env_id_base, self.seed_value & (2 ** 64 - 1),
                                      dev.index if dev.index is not None else torch.cuda.current_device(),
                                      self.state.data_ptr(), nbytes.value, C.byref(self._h)))
        self.obs64 = self.info = self._np_info = None
        if info_mode == "full":
            self.enable_taps()
        elif info_mode == "episode":
            self.enable_taps(done_only=True)
        if self.cfg.mode == _lib.MODE_TEST:
            self.set_scenario(*scenario_tables(self.kwargs["scenario"], self.cfg.screen_x, self.cfg.screen_y))
        # env:410-415: with render_path the reference appends the frame position after every step and hands the list out as
        # info["flight_path"] when the episode ends.  Here: a device ring of the last `ring` positions per env (whole
        # episodes when n_steps fits; bounded to ~1 GB), gathered for the envs that finish.
        self.path_ring = None
        if self.kwargs.get("render_path") is True and info_mode != "none":
            self.enable_flight_paths()
        # Page-locked host buffers of the host-facing step(): the kernel reads the actions from them and writes
        # obs / reward / done / terminal observations into them directly (d2d_step_host_ex, mapped memory).
        # Two output sets alternate, so the arrays of step t stay intact until step t + 2 (see copy_outputs).
        self.copy_outputs = bool(copy_outputs)
        self._h_act = torch.zeros((n, 2), dtype=torch.float32).pin_memory()
        self._h_out = [(torch.zeros((n, 27), dtype=torch.float32).pin_memory(), torch.zeros(n, dtype=torch.float32).pin_memory(),
                        torch.zeros(n, dtype=torch.uint8).pin_memory()) for _ in range(2)]
        self._h_term = torch.zeros((n, 27), dtype=torch.float32).pin_memory()
        self._np_act = self._h_act.numpy()
        self._np_out = [tuple(t.numpy() for t in trio) for trio in self._h_out]
        self._np_term = self._h_term.numpy()
        self._flip = 0
        self._empty_infos = [_EMPTY_INFO] * n
        self._host_stepped = False               # the newest observations are in the host buffers, not in self.obs
        self._pending = None
        self._closed = False
        self._renderer = None
        self._last_actions = None

    # ------------------------------------------------------------------ helpers
    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    @staticmethod
    def _dp(a):
        return a.ctypes.data_as(C.POINTER(C.c_double))

    def enable_taps(self, done_only=False):
        """Allocate and register the float64 taps (include/drone2d.h: d2d_set_taps).  Default: observation and
        info rows of every env on every step; done_only: only the info tap, and an env's row is written only on
        the step that ends its episode (all that step() needs in info_mode="episode": 22 MB less per step)."""
        n = self.num_envs
        self.obs64 = None if done_only else torch.zeros((n, 27), dtype=torch.float64, device=self.device)
        if done_only:
            # only the rows of finishing envs are written: they go straight into page-locked host memory (mapped into
            # the kernel's address space), so building the info dicts needs no device gather and no extra copy
            self.info = torch.zeros((n, _lib.INFO_DOUBLES), dtype=torch.float64).pin_memory()
            self._np_info = self.info.numpy()
        else:
            self.info = torch.zeros((n, _lib.INFO_DOUBLES), dtype=torch.float64, device=self.device)
            self._np_info = None
        _lib.check(self._L.d2d_set_taps(self._h, self.obs64.data_ptr() if self.obs64 is not None else None, self.info.data_ptr()))
        _lib.check(self._L.d2d_set_info_done_only(self._h, 1 if done_only else 0))

    def enable_flight_paths(self, ring=None):
        np_ = (self.num_envs + 31) // 32 * 32
        if ring is None:
            ring = max(1, min(int(self.cfg.n_steps), (1 << 30) // (8 * np_)))
        self.path_ring = torch.zeros((int(ring), np_, 2), dtype=torch.float32, device=self.device)
        _lib.check(self._L.d2d_set_flight_path_ring(self._h, self.path_ring.data_ptr(), int(ring)))

    def flight_paths_of(self, indices, steps):
        """[(x, screen_height - y), ...] of the episodes that just ended in envs `indices` (`steps`: their lengths)."""
        if self.path_ring is None or len(indices) == 0:
            return [[] for _ in indices]
        K = self.path_ring.shape[0]
        cols = self.path_ring[:, torch.as_tensor(np.asarray(indices), device=self.device)].cpu().numpy()     # [K, m, 2]
        out = []
        for j, T in enumerate(steps):
            T = int(T)
            if T <= K:
                rows = cols[:T, j]
            else:                                      # only the last K positions survive, oldest first
                rows = np.roll(cols[:, j], -(T % K), axis=0)
            out.append([(float(x), float(y)) for x, y in rows])
        return out

    def info_rows(self, mask):
        """Info-tap rows of the envs selected by a boolean device mask, as a float64 device tensor (wherever the tap
        lives: device memory, or page-locked host memory in info_mode="episode")."""
        if self.info.is_cuda:
            return self.info[mask]
        torch.cuda.current_stream(self.device).synchronize()      # the kernel's stores into host memory have landed
        return self.info[mask.cpu()].to(self.device)

    def disable_taps(self):
        self.obs64 = self.info = self._np_info = None
        _lib.check(self._L.d2d_set_taps(self._h, None, None))

    def set_scenario(self, wps, us, xparams, yparams, obstacles):
        """Upload one shared path + obstacle set (test mode)."""
        wps, us = np.ascontiguousarray(wps, np.float64), np.ascontiguousarray(us, np.float64)
        xp, yp = np.ascontiguousarray(xparams, np.float64), np.ascontiguousarray(yparams, np.float64)
        ob = np.ascontiguousarray(obstacles, np.float64).reshape(-1, 3)
        _lib.check(self._L.d2d_set_scenario(self._h, len(wps), self._dp(wps), self._dp(us), self._dp(xp), self._dp(yp),
                                            len(ob), self._dp(ob) if len(ob) else None, self._stream()))
        self.scenario = {"wps": wps, "us": us, "xparams": xp, "yparams": yp, "obstacles": ob}

    # ------------------------------------------------------- device-resident API
    def reset_device(self, mask=None):
        """Re-spawn all envs (or those with mask != 0) on the device; returns the obs tensor [N, 27] (cuda)."""
        m = None
        if mask is not None:
            m = mask.to(device=self.device, dtype=torch.uint8).contiguous()
        _lib.check(self._L.d2d_reset(self._h, m.data_ptr() if m is not None else None, self.obs.data_ptr(),
                                     self._stream()))
        self._host_stepped = False
        return self.obs

    def step_device(self, actions=None, auto_reset=True, terminal_obs=True):
        """One transition for all envs, nothing leaves the device.  `actions`: cuda
        float32 [N, 2] (None -> in-kernel uniform random actions).  Returns the
        persistent (obs, rewards, dones) tensors, overwritten by the next call."""
        a = None
        if actions is not None:
            a = actions
            if a.dtype != torch.float32 or not a.is_contiguous() or a.device != self.device:
                a = a.to(device=self.device, dtype=torch.float32).contiguous()
            if a.shape != (self.num_envs, 2):
                raise ValueError("actions must have shape (%d, 2)" % self.num_envs)
        _lib.check(self._L.d2d_step(self._h, a.data_ptr() if a is not None else None, self.obs.data_ptr(),
                                    self.rewards.data_ptr(), self.dones.data_ptr(),
                                    self.terminal_obs.data_ptr() if terminal_obs else None,
                                    1 if auto_reset else 0, self._stream()))
        self._last_actions = a
        self._host_stepped = False
        return self.obs, self.rewards, self.dones

    # ------------------------------------------------------------- SB3 VecEnv API
    def reset(self):
        self.reset_device()
        return self.obs.cpu().numpy()

    def step_async(self, actions):
        self._pending = np.asarray(actions, dtype=np.float32).reshape(self.num_envs, 2)

    def step_wait(self):
        if self._pending is None:
            raise RuntimeError("step_wait() without step_async()")
        actions, self._pending = self._pending, None
        np.copyto(self._np_act, actions)
        self._flip ^= 1
        obs_t, rew_t, done_t = self._h_out[self._flip]
        _lib.check(self._L.d2d_step_host_ex(self._h, self._h_act.data_ptr(), obs_t.data_ptr(), rew_t.data_ptr(), done_t.data_ptr(),
                                            self._h_term.data_ptr(), 1, self._stream()))
        self._host_stepped = True
        obs, rew, done = self._np_out[self._flip]
        dones = done.view(np.bool_)
        if self.copy_outputs:
            obs, rew, dones = obs.copy(), rew.copy(), dones.copy()
        return obs, rew, dones, self._build_infos(dones)

    def step(self, actions):
        self.step_async(actions)
        return self.step_wait()

    def _info_row(self, row, done):
        d = {k: float(row[j]) for j, k in enumerate(INFO_KEYS)}
        d["env_steps"] = int(row[7])
        if done:
            d.update(self._done_keys(row))
        return d

    @staticmethod
    def _done_keys(r):
        bits = int(r[11])
        return {"n_collisions": int(bool(bits & 1) and not (bits & (2 | 4 | 8))),   # env:595-610 overwrite order
                "n_successful_runs": int(bool(bits & 2)), "n_failed_runs": int(bool(bits & (1 | 4 | 8))),
                "APE": float(r[9]), "total_reward": float(r[10]),
                "flight_path": [],                 # env:612; positions are not recorded on the device
                "episode": {"r": float(r[10]), "l": int(r[7])}}

    _DONE_KEYS = ("reward", "collision_avoidance_reward", "path_adherence", "path_progression", "collision_reward",
                  "reach_end_reward", "agressive_alpha_reward", "dist_closest_obs", "APE", "total_reward", "env_steps",
                  "n_collisions", "n_successful_runs", "n_failed_runs")

    def _done_infos(self, rows):
        """Info dicts of the envs that finished this step (rows: their info-tap rows [m, 16]).  The integer keys are
        derived for all rows at once; each dict is then one dict(zip(...)) over Python scalars."""
        bits = rows[:, 11].astype(np.int64)
        steps = rows[:, 7].astype(np.int64).tolist()
        cols = [rows[:, j].tolist() for j in (0, 1, 2, 3, 4, 5, 6, 8, 9, 10)]
        cols.append(steps)
        cols.append((((bits & 1) != 0) & ((bits & 14) == 0)).astype(np.int64).tolist())      # env:595-610 overwrite order
        cols.append(((bits & 2) != 0).astype(np.int64).tolist())
        cols.append(((bits & 13) != 0).astype(np.int64).tolist())
        keys = self._DONE_KEYS
        out = []
        for vals in zip(*cols):
            d = dict(zip(keys, vals))
            d["flight_path"] = []
            d["episode"] = {"r": vals[9], "l": vals[10]}
            out.append(d)
        return out

    def _build_infos(self, dones):
        n = self.num_envs
        idx = np.flatnonzero(dones)
        if self.info_mode == "full":
            rows = self.info.cpu().numpy()
            infos = [self._info_row(rows[i], dones[i]) for i in range(n)]
        else:
            infos = list(self._empty_infos)
            if len(idx) and self.info is not None:
                if self._np_info is not None and self._host_stepped:
                    rows = self._np_info[idx]            # written by the kernel into page-locked host memory
                else:
                    rows = self.info_rows(torch.from_numpy(dones).to(self.device)).cpu().numpy()
                done_infos = self._done_infos(rows)
                if self.path_ring is not None:
                    for d, fp in zip(done_infos, self.flight_paths_of(idx, rows[:, 7])):
                        d["flight_path"] = fp
                for i, d in zip(idx.tolist(), done_infos):
                    infos[i] = d
            else:
                for i in idx:
                    infos[i] = {}
        if len(idx):
            term = self._np_term[idx]            # rows written by the kernel for the envs that finished in this step
            for k, i in enumerate(idx):
                infos[i]["terminal_observation"] = term[k]
        return infos

    def close(self):
        if not self._closed and self._h:
            self._L.d2d_destroy(self._h)
            self._h = C.c_void_p()
            self._closed = True

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass

    def seed(self, seed=None):
        """SB3 VecEnv.seed: re-key the counter-based reset stream (takes effect at the next reset(); the handle is
        rebuilt on the same state tensor, taps / scenario / stage / flight-path ring are registered again) and return
        the per-env seeds (global env i draws from stream (seed, env_id_base + i, episode))."""
        if seed is not None and int(seed) != self.seed_value:
            self.seed_value = int(seed)
            torch.cuda.synchronize(self.device)
            self._L.d2d_destroy(self._h)
            self._h = C.c_void_p()
            dev = self.device
            _lib.check(self._L.d2d_create(C.byref(self.cfg), self.num_envs, self.env_id_base, self.seed_value & (2 ** 64 - 1),
                                          dev.index if dev.index is not None else torch.cuda.current_device(),
                                          self.state.data_ptr(), self.state.numel(), C.byref(self._h)))
            if self.info is not None:
                _lib.check(self._L.d2d_set_taps(self._h, self.obs64.data_ptr() if self.obs64 is not None else None, self.info.data_ptr()))
                _lib.check(self._L.d2d_set_info_done_only(self._h, 1 if self.info_mode == "episode" else 0))
            if self.cfg.mode == _lib.MODE_TEST:
                sc = self.scenario
                self.set_scenario(sc["wps"], sc["us"], sc["xparams"], sc["yparams"], sc["obstacles"])
            else:
                _lib.check(self._L.d2d_set_stage(self._h, int(self.cfg.stage), float(self.cfg.stage_spawn_chance)))
            if self.path_ring is not None:
                _lib.check(self._L.d2d_set_flight_path_ring(self._h, self.path_ring.data_ptr(), int(self.path_ring.shape[0])))
        return [self.seed_value + self.env_id_base + i for i in range(self.num_envs)]

    def get_attr(self, attr_name, indices=None):
        idx = range(self.num_envs) if indices is None else ([indices] if isinstance(indices, int) else indices)
        return [getattr(self, attr_name) for _ in idx]

    def set_attr(self, attr_name, value, indices=None):
        setattr(self, attr_name, value)

    def env_method(self, method_name, *args, indices=None, **kwargs):
        idx = range(self.num_envs) if indices is None else ([indices] if isinstance(indices, int) else indices)
        return [getattr(self, method_name)(*args, **kwargs) for _ in idx]

    def env_is_wrapped(self, wrapper_class, indices=None):
        idx = range(self.num_envs) if indices is None else ([indices] if isinstance(indices, int) else indices)
        return [False for _ in idx]

    def render(self, mode="rgb_array", index=0, pixels=None, flight_path=None):
        """Renderer adapter (render.py): uint8 [h, w, 3] frame of env `index`, composed like the reference's
        Drone2dEnv.render (env:775-906).  Copies that env's state to the host; not for use inside a hot loop."""
        from .render import FrameRenderer
        if self._renderer is None or self._renderer.pixels != pixels:
            self._renderer = FrameRenderer(self, pixels=pixels)
        if self._host_stepped:
            act = self._np_act[index].copy()
        else:
            act = None if self._last_actions is None else self._last_actions[index].cpu().numpy()
        return self._renderer.frame(index, action=act, flight_path=flight_path)

    def current_obs_row(self, index):
        """float64 copy of env `index`'s newest observation, wherever the last step left it."""
        if self._host_stepped:
            return self._np_out[self._flip][0][index].astype(np.float64)
        # the float64 tap keeps the TERMINAL observation of an env that finished in the last step (what the parity
        # tests compare); its current observation (first of the new episode) exists in float32 only
        use64 = self.obs64 is not None and not bool(self.dones[index].item())
        return (self.obs64 if use64 else self.obs)[index].double().cpu().numpy()

    def get_images(self, indices=None, pixels=325):
        """SB3 VecEnv.get_images: one frame per env (bounded to the first 16 unless `indices` is given)."""
        idx = range(min(self.num_envs, 16)) if indices is None else indices
        return [self.render(index=i, pixels=pixels) for i in idx]

    # ------------------------------------------------- state / tables / statistics
    def get_state(self):
        """float64 [N, 36] blob, layout documented in include/drone2d.h (d2d_get_state)."""
        out = np.empty((self.num_envs, _lib.STATE_DOUBLES), dtype=np.float64)
        _lib.check(self._L.d2d_get_state(self._h, self._dp(out)))
        return out

    def set_state(self, blob):
        blob = np.ascontiguousarray(blob, dtype=np.float64)
        if blob.shape != (self.num_envs, _lib.STATE_DOUBLES):
            raise ValueError("state blob must have shape (%d, %d)" % (self.num_envs, _lib.STATE_DOUBLES))
        _lib.check(self._L.d2d_set_state(self._h, self._dp(blob)))

    def get_tables(self):
        """Per-env path/obstacle tables (curriculum mode) as generated by the in-kernel reset."""
        n = self.num_envs
        t = {"wps": np.empty((n, _lib.MAX_WPS, 2)), "us": np.empty((n, _lib.MAX_WPS)),
             "xparams": np.empty((n, _lib.MAX_WPS, 3)), "yparams": np.empty((n, _lib.MAX_WPS, 3)),
             "n_obs": np.empty(n, dtype=np.int32), "obstacles": np.empty((n, _lib.MAX_OBS_ENV, 3))}
        _lib.check(self._L.d2d_get_tables(self._h, self._dp(t["wps"]), self._dp(t["us"]), self._dp(t["xparams"]),
                                          self._dp(t["yparams"]), t["n_obs"].ctypes.data_as(C.POINTER(C.c_int32)),
                                          self._dp(t["obstacles"])))
        return t

    def set_tables(self, wps, us, xparams, yparams, n_obs, obstacles):
        n = self.num_envs
        arrs = [np.ascontiguousarray(a, np.float64) for a in (wps, us, xparams, yparams)]
        n_obs = np.ascontiguousarray(n_obs, np.int32)
        obstacles = np.ascontiguousarray(obstacles, np.float64)
        want = [(n, _lib.MAX_WPS, 2), (n, _lib.MAX_WPS), (n, _lib.MAX_WPS, 3), (n, _lib.MAX_WPS, 3)]
        for a, w in zip(arrs, want):
            if a.shape != w:
                raise ValueError("table shape %s != %s" % (a.shape, w))
        if obstacles.shape != (n, _lib.MAX_OBS_ENV, 3) or n_obs.shape != (n,):
            raise ValueError("obstacle table shapes")
        _lib.check(self._L.d2d_set_tables(self._h, *[self._dp(a) for a in arrs],
                                          n_obs.ctypes.data_as(C.POINTER(C.c_int32)), self._dp(obstacles)))

    def closest_u(self, positions):
        """QPMI2D.get_closest_u on arbitrary positions [N, 2] (env i uses its own path)."""
        pos = torch.as_tensor(np.ascontiguousarray(positions, np.float64), device=self.device)
        u = torch.empty(self.num_envs, dtype=torch.float64, device=self.device)
        nfev = torch.empty(self.num_envs, dtype=torch.int32, device=self.device)
        _lib.check(self._L.d2d_debug_closest_u(self._h, pos.data_ptr(), u.data_ptr(), nfev.data_ptr(), self._stream()))
        return u.cpu().numpy(), nfev.cpu().numpy()

    def episode_stats_tensor(self):
        """Zero-copy float64[16] device tensor of the episode-statistics accumulator
        (the vector that is all-reduced over ranks once per rollout)."""
        ptr = C.c_void_p()
        _lib.check(self._L.d2d_episode_stats_dev(self._h, C.byref(ptr)))
        return torch.as_tensor(_DeviceArray(ptr.value, (_lib.STATS_DOUBLES,), "<f8"), device=self.device)

    def episode_stats(self, reset=True):
        out = np.zeros(_lib.STATS_DOUBLES, dtype=np.float64)
        _lib.check(self._L.d2d_episode_stats(self._h, self._dp(out), 1 if reset else 0))
        return dict(zip(STATS_KEYS, out.tolist()))

    def set_stage(self, stage=None, spawn_chance=None, sim_num=None):
        """Curriculum scheduler: set the stage explicitly or from the training progress `sim_num`
        (thresholds and obstacle-chance ramps of env:326-372)."""
        if sim_num is not None:
            stage, spawn_chance = stage_from_timesteps(int(sim_num))
        if spawn_chance is None:
            spawn_chance = {3: 0.6, 4: 1.0}.get(int(stage), 0.0)
        _lib.check(self._L.d2d_set_stage(self._h, int(stage), float(spawn_chance)))
        self.cfg.stage, self.cfg.stage_spawn_chance = int(stage), float(spawn_chance)

    def debug_skip_physics(self, enable=True):
        """Replay mode (d2d_debug_skip_physics): step() runs everything after space.step on the injected state."""
        _lib.check(self._L.d2d_debug_skip_physics(self._h, 1 if enable else 0))

    def role_cycles(self, reset=True):
        """Profiling hook (d2d_debug_role_cycles): longest per-warp clock span per role / phase."""
        out = (C.c_uint64 * 24)()
        _lib.check(self._L.d2d_debug_role_cycles(self._h, out, 1 if reset else 0))
        names = ("main_total", "refill_sample", "refill_observe", "physics", "obstacle_pass", "observe", "reward_swap")
        return {nm: {"max": int(out[r]), "mean": int(out[8 + r]) // max(int(out[16 + r]), 1), "warps": int(out[16 + r])}
                for r, nm in enumerate(names)}

    @property
    def launch_count(self):
        return int(self._L.d2d_launch_count(self._h))
