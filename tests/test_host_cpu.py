"""Host-side logic and the C-ABI surface, without a GPU: the library must load and
export every symbol include/drone2d.h declares; configuration mapping and the
scenario table builders must reproduce the reference bit for bit."""
import ctypes as C
import importlib
import json
import os
import re

import numpy as np
import pytest

from conftest import GOLDEN, REPO

PKG = "drone_2d_custom_gym_env_for_reinforcement_learning_b200"
pkg = importlib.import_module(PKG)
_lib = importlib.import_module(PKG + "._lib")


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(REPO, "include", "drone2d.h")).read()
    declared = sorted(set(re.findall(r"\b(d2d_[a-z0-9_]+)\s*\(", hdr)))
    assert len(declared) >= 19
    L = _lib.load()
    for name in declared:
        assert hasattr(L, name), name
    assert sorted(_lib.EXPORTS) == declared
    assert L.d2d_abi_version() == 1


def test_library_is_built_for_sm_100a_only():
    import subprocess
    out = subprocess.run(["cuobjdump", "--list-elf", _lib.lib_path()], capture_output=True, text=True).stdout
    assert "sm_100a" in out
    assert not re.search(r"sm_(?!100a)\d+", out), out


def test_default_config_matches_rl_config():
    c = _lib.default_config()
    ref = pkg.rl_config.env_train_config
    for k in ("lookahead", "danger_range", "danger_angle", "abs_inv_CA_min_rew", "PA_band_edge", "PA_scale",
              "PP_vel_scale", "PP_rew_max", "PP_rew_min", "rew_collision", "reach_end_radius", "rew_reach_end",
              "AA_angle", "AA_band", "rew_AA", "n_steps", "n_wps", "path_segment_length"):
        assert getattr(c, k) == ref[k], k
    assert (c.screen_x, c.screen_y) == (1300, 1300)
    assert (c.mass_f, c.mass_m, c.width, c.height) == (0.2, 0.4, 100, 20)       # Drone(x, y, a, 20, 100, 0.2, 0.4, 0.4)
    assert (c.gravity_x, c.gravity_y, c.dt, c.iterations) == (0, -1000, 1.0 / 60, 10)


def test_reference_config_keys_are_all_accepted():
    ref_keys = {'render_sim', 'render_path', 'render_shade', 'render_text', 'shade_distance', 'n_steps', 'n_fall_steps',
                'change_target', 'initial_throw', 'random_path_spawn', 'path_segment_length', 'n_wps', 'screensize_x',
                'screensize_y', 'lookahead', 'spawn_corners', 'danger_range', 'danger_angle', 'abs_inv_CA_min_rew',
                'PA_band_edge', 'PA_scale', 'PP_vel_scale', 'PP_rew_max', 'PP_rew_min', 'rew_collision',
                'reach_end_radius', 'rew_reach_end', 'AA_angle', 'AA_band', 'rew_AA', 'use_Lambda', 'mode', 'scenario'}
    assert set(pkg.rl_config.env_train_config) == ref_keys            # rl_config.py:11-43
    c, kw = pkg.config_from_kwargs(dict(pkg.rl_config.env_test_config))
    assert c.mode == _lib.MODE_TEST and kw["scenario"] == "large"
    assert (c.spawn_xmin, c.spawn_xmax, c.spawn_ymin, c.spawn_ymax) == (50.0, 1300 / 2 - 1300 / 4 - 50, 150.0, 1000.0)


def test_config_mapping_modes_and_errors():
    c, _ = pkg.config_from_kwargs(dict(mode="curriculum", scenario="stage_3"))
    assert (c.mode, c.stage, c.stage_spawn_chance) == (_lib.MODE_CURRICULUM, 3, 0.6)
    c, _ = pkg.config_from_kwargs(dict(mode="curriculum", scenario="stage_4"))
    assert (c.stage, c.stage_spawn_chance) == (4, 1.0)
    c, _ = pkg.config_from_kwargs(dict(mode="curriculum", scenario="None", sim_num=1300000))
    assert c.stage == 3 and abs(c.stage_spawn_chance - 0.4) < 1e-12    # env:343 ramp 0.2 -> 0.6
    c, _ = pkg.config_from_kwargs(dict(mode="curriculum", scenario="None", sim_num=2500000))
    assert c.stage == 5
    c, _ = pkg.config_from_kwargs(dict(mode="curriculum", scenario="stage_1", use_Lambda=1))
    assert c.use_Lambda == 0                                           # `self.use_Lambda is True` (env:496)
    with pytest.raises(ValueError):
        pkg.config_from_kwargs(dict(mode="test", scenario="stage_1"))  # the reference crashes at env:316
    with pytest.raises(ValueError):
        pkg.config_from_kwargs(dict(mode="flight", scenario="large"))
    with pytest.raises(TypeError):
        pkg.config_from_kwargs(dict(mode="test", scenario="large", not_a_key=1))


def test_abi_argument_validation_without_gpu():
    L = _lib.load()
    c = _lib.default_config()
    n = C.c_size_t()
    assert L.d2d_state_bytes(C.byref(c), 65536, C.byref(n)) == 0 and n.value > 65536 * 272
    assert L.d2d_state_bytes(C.byref(c), 0, C.byref(n)) == -1
    c.abi_version = 99
    assert L.d2d_state_bytes(C.byref(c), 16, C.byref(n)) == -1
    assert b"abi_version" in L.d2d_last_error()
    c = _lib.default_config()
    c.stage = 7
    assert L.d2d_state_bytes(C.byref(c), 16, C.byref(n)) == -1
    c = _lib.default_config()
    c.n_wps = 40
    assert L.d2d_state_bytes(C.byref(c), 16, C.byref(n)) == -1
    assert L.d2d_step(None, None, None, None, None, None, 1, None) == -1
    assert L.d2d_reset(None, None, None, None) == -1
    assert L.d2d_launch_count(None) == 0


def test_scenario_tables_bit_exact_with_reference():
    """scenarios.scenario_tables vs test_scenarios.create_test_scenario (golden from the reference's own code)."""
    with open(os.path.join(GOLDEN, "scenarios.json")) as f:
        gold = json.load(f)
    assert sorted(gold) == sorted(pkg.rl_config.TEST_SCENARIOS)
    for name, rec in gold.items():
        wps, us, xp, yp, ob = pkg.scenarios.scenario_tables(name)
        assert np.array_equal(wps, np.array(rec["wps"])), name
        assert np.array_equal(us, np.array(rec["us"])), name
        assert np.array_equal(xp, np.array(rec["x_params"])), name
        assert np.array_equal(yp, np.array(rec["y_params"])), name
        assert np.array_equal(ob, np.array(rec["obstacles"]).reshape(-1, 3)), name


def test_host_path_tables_match_reference_kats():
    with open(os.path.join(GOLDEN, "path_kats.json")) as f:
        kats = json.load(f)
    for name, rec in kats.items():
        path = pkg.scenarios.PathTables(rec["path"]["wps"])
        assert np.array_equal(path.us, np.array(rec["path"]["us"]))
        for e in rec["evals"]:
            if e["u"] < 0:
                continue      # host PathTables is only used on [0, L] (obstacle placement)
            np.testing.assert_allclose(path.point(e["u"]), e["p"], rtol=0, atol=1e-9)
            assert abs(path.tangent_angle(e["u"]) - e["dir"]) < 1e-12


def test_spawn_rectangles():
    S = pkg.scenarios.spawn_rectangle
    assert S("perpendicular") == (50.0, 550.0, 50.0, 1200.0)          # env:223-226
    assert S("corridor") == (50.0, 250.0, 150.0, 1000.0)              # env:262-265
    assert S("S_corridor") == (50.0, 200.0, 150.0, 1000.0)            # env:275-278
    assert S("impossible") == (50.0, 650.0, 150.0, 1000.0)            # env:301-304


def test_vec_env_fails_loudly_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        pkg.Drone2dVecEnv(4, mode="test", scenario="corridor")


def test_policy_weight_image_layout_and_split():
    """d2d_policy_pack (host code of the library): the image the tcgen05 kernel copies into shared memory.
    Every operand element sits at its UMMA K-major core-matrix position (8 rows x 8 halfs), hi = fp16(w) and
    hi + lo reproduces the float32 weight to 22 bits, the first layer's bias is column 27, columns 28..31 are zero;
    the first-layer weights come last (the kernel overwrites them with the hidden activations)."""
    import ctypes as C
    z = np.load(os.path.join(GOLDEN, "agent_19_policy.npz"))
    keys = ("mlp_extractor__policy_net__0__weight", "mlp_extractor__policy_net__0__bias", "mlp_extractor__policy_net__2__weight",
            "mlp_extractor__policy_net__2__bias", "action_net__weight", "action_net__bias", "log_std")
    w = [np.ascontiguousarray(z[k], dtype=np.float32) for k in keys]
    L = _lib.load()
    n = L.d2d_policy_weights_bytes()
    small = 4 * (64 + 128 + 2 + 2 + 2 + 26)                       # b1, w2, b2, std, log_std, padding to 128 bytes
    assert n == 2 * 2 * 64 * 64 + small + 2 * 2 * 64 * 32 and n % 16 == 0 and (2 * 2 * 64 * 64 + small) % 128 == 0
    img = np.zeros(n, dtype=np.uint8)
    fp = lambda a: a.ctypes.data_as(C.c_void_p)
    _lib.check(L.d2d_policy_pack(*[fp(a) for a in w], fp(img)))

    def idx(r, k, R):                                   # d2d_umma_idx (halfs)
        return ((k >> 3) * (R >> 3) + (r >> 3)) * 64 + (r & 7) * 8 + (k & 7)

    w1_hi, w1_lo = img[0:8192].view(np.float16), img[8192:16384].view(np.float16)
    tail = img[16384:16384 + small].view(np.float32)
    w0 = img[16384 + small:]
    w0_hi, w0_lo = w0[0:4096].view(np.float16), w0[4096:8192].view(np.float16)
    full0 = np.zeros((64, 32), np.float32)
    full0[:, :27], full0[:, 27] = w[0], w[1]
    for j in range(64):
        for k in range(32):
            o = idx(j, k, 64)
            assert w0_hi[o] == np.float16(full0[j, k])
            assert abs(float(w0_hi[o]) + float(w0_lo[o]) - float(full0[j, k])) <= abs(float(full0[j, k])) * 2.0 ** -21 + 6e-8
        for k in range(64):
            o = idx(j, k, 64)
            assert w1_hi[o] == np.float16(w[2][j, k])
            assert abs(float(w1_hi[o]) + float(w1_lo[o]) - float(w[2][j, k])) <= abs(float(w[2][j, k])) * 2.0 ** -21 + 6e-8
    assert sorted(set(idx(j, k, 64) for j in range(64) for k in range(64))) == list(range(4096))   # a permutation
    assert np.array_equal(tail[:64], w[3]) and np.array_equal(tail[64:192], w[4].ravel())
    assert np.array_equal(tail[192:194], w[5]) and np.allclose(tail[194:196], np.exp(w[6]), rtol=1e-6)
    assert np.array_equal(tail[196:198], w[6]) and not tail[198:].any()
