/* include/drone2d.h -- C ABI of the B200-native batched Drone2dEnv step.
 *
 * The reference has no native interface: the path sits behind the Python
 * gym.Env methods of Drone2dEnv (reference drone_2d_custom_gym_env/
 * drone_2d_env.py, "env" below).  Each entry point names the reference
 * interface it replaces.  All functions return 0 on success or a negative
 * D2D_E* code; d2d_last_error() gives the message.  No exceptions cross the
 * ABI.  All "dev" pointers are device pointers owned by the caller (PyTorch
 * tensors); the library allocates nothing per step.  Launches are asynchronous
 * on the caller-supplied cudaStream_t (passed as void*); no host sync inside
 * unless stated.  A handle is bound to one device and is not thread-safe.
 */
#ifndef DRONE2D_H
#define DRONE2D_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define D2D_ABI_VERSION 1
#define D2D_OBS_DIM 27          /* env:160-162,765-773 */
#define D2D_ACT_DIM 2           /* env:155-157 */
#define D2D_MAX_WPS 16          /* waypoints per path (reference uses 6..14) */
#define D2D_MAX_OBS_SHARED 64   /* circles of a shared test scenario (max 58, test_scenarios.py:218-233) */
#define D2D_MAX_OBS_ENV 32      /* per-env circles in curriculum stages (env:364-372 is unbounded:
                                   ceil(N(1,4))+1; P(count > 32) ~ 3e-14, clamped) */
#define D2D_STATE_DOUBLES 36    /* doubles per env in the d2d_get_state/d2d_set_state blob */
#define D2D_INFO_DOUBLES 16     /* doubles per env in the per-step info blob */
#define D2D_STATS_DOUBLES 16    /* doubles in the episode-statistics vector */

#define D2D_OK 0
#define D2D_EINVAL (-1)   /* bad argument */
#define D2D_ECUDA (-2)    /* CUDA runtime error */
#define D2D_ESTATE (-3)   /* call order / missing tables */

enum { D2D_MODE_TEST = 0, D2D_MODE_CURRICULUM = 1 };

/* rl_config.py:10-44: every key that reaches the arithmetic of the path, plus
 * the Drone.py / init_pymunk constants (env:185,233,406).  Render keys and the
 * dead keys (n_fall_steps, initial_throw, change_target; env:917-946 is never
 * called) are accepted by the Python layer and do not reach the ABI. */
typedef struct d2d_config {
    int32_t abi_version;            /* D2D_ABI_VERSION */
    int32_t mode;                   /* D2D_MODE_*  ('mode' kwarg) */
    int32_t stage;                  /* 1..5 ('scenario' = stage_k) in curriculum mode */
    int32_t n_steps;                /* 'n_steps' */
    int32_t n_wps;                  /* 'n_wps' (curriculum); test mode: set by d2d_set_scenario */
    int32_t spawn_corner_lo, spawn_corner_hi; /* 'spawn_corners' */
    int32_t random_path_spawn;      /* 'random_path_spawn' */
    int32_t use_Lambda;             /* 'use_Lambda' */
    int32_t iterations;             /* pymunk Space.iterations (default 10) */
    double screen_x, screen_y;      /* 'screensize_x/y' */
    double path_segment_length;     /* 'path_segment_length' */
    double lookahead;
    double danger_range, danger_angle, abs_inv_CA_min_rew;
    double PA_band_edge, PA_scale;
    double PP_vel_scale, PP_rew_max, PP_rew_min;
    double rew_collision, reach_end_radius, rew_reach_end;
    double AA_angle, AA_band, rew_AA;
    double stage_spawn_chance;      /* stage_3: 0.6, stage_4: 1.0 (env:340-358); explicit instead of
                                       the checkpoint-filename glob of env:79-85 */
    double spawn_xmin, spawn_xmax, spawn_ymin, spawn_ymax; /* test mode spawn rectangle (env:221-311) */
    double mass_f, mass_m, width, height;                    /* Drone(x,y,a,20,100,0.2,0.4,0.4) */
    double gravity_x, gravity_y, dt, force_scale;            /* env:185,406,150 */
} d2d_config;

typedef struct d2d_env d2d_env;

/* Fill *cfg with rl_config.py:10-44 defaults (mode = curriculum, stage_1). */
int d2d_default_config(d2d_config *cfg);

/* Bytes of device memory the persistent SoA state + per-env tables of n_envs
 * environments need.  The caller allocates it (256-byte aligned) and keeps it
 * alive for the lifetime of the handle. */
int d2d_state_bytes(const d2d_config *cfg, int64_t n_envs, size_t *bytes);

/* Replaces Drone2dEnv.__init__(**kwargs) (env:33-162) for n_envs environments.
 * env_id_base: global index of env 0 (rank offset) -- the reset RNG is keyed by
 * (seed, global env id, episode), so results do not depend on the sharding. */
int d2d_create(const d2d_config *cfg, int64_t n_envs, int64_t env_id_base, uint64_t seed,
               int device, void *state_dev, size_t state_bytes, d2d_env **out);
int d2d_destroy(d2d_env *env);

/* Replaces test_scenarios.create_test_scenario (test_scenarios.py:169-246) as the
 * table source in test mode: one path + obstacle set shared by all envs.  HOST
 * pointers: wps [n_wps][2], us [n_wps], xparams/yparams [n_wps-2][3] (QPMI2D
 * tables, predef_path.py:20-50), obstacles [n_obs][3] = (x, y, r). */
int d2d_set_scenario(d2d_env *env, int n_wps, const double *wps, const double *us,
                     const double *xparams, const double *yparams,
                     int n_obs, const double *obstacles, void *stream);

/* Curriculum mode, parity/checkpoint use: overwrite the per-env tables from
 * HOST arrays wps [n][D2D_MAX_WPS][2], us [n][D2D_MAX_WPS],
 * xparams/yparams [n][D2D_MAX_WPS][3], n_obs [n], obstacles [n][D2D_MAX_OBS_ENV][3].
 * d2d_get_tables reads them back (what the in-kernel reset generated).  Both sync.
 * d2d_set_tables also re-derives the path points of the closest-point search's opening
 * evaluations (the `hint` table of the state block) from the new rows. */
int d2d_set_tables(d2d_env *env, const double *wps, const double *us, const double *xparams,
                   const double *yparams, const int32_t *n_obs, const double *obstacles);
int d2d_get_tables(d2d_env *env, double *wps, double *us, double *xparams,
                   double *yparams, int32_t *n_obs, double *obstacles);

/* Replaces Drone2dEnv.reset() (env:908-912) for all envs (mask_dev == NULL) or
 * the envs with mask_dev[i] != 0: fused re-spawn (env:182-392) from the
 * counter-based RNG + first observation.  obs_dev: float32 [n][27]. */
int d2d_reset(d2d_env *env, const uint8_t *mask_dev, float *obs_dev, void *stream);

/* Replaces Drone2dEnv.step(action) (env:394-615) + the auto-reset SB3's VecEnv
 * performs on done.  ONE kernel launch.
 *   actions_dev      float32 [n][2]  (NULL: uniform random actions in [-1,1) from the
 *                                     in-kernel Philox stream keyed by the step counter)
 *   obs_dev          float32 [n][27] observation AFTER auto-reset (SB3 convention)
 *   reward_dev       float32 [n]
 *   done_dev         uint8   [n]
 *   terminal_obs_dev float32 [n][27] or NULL: rows of finished envs receive the
 *                                     last observation of the finished episode
 *                                     (SB3 infos[i]["terminal_observation"])
 *   auto_reset       0: finished envs are left in their terminal state (gym.Env semantics) */
int d2d_step(d2d_env *env, const float *actions_dev, float *obs_dev, float *reward_dev,
             uint8_t *done_dev, float *terminal_obs_dev, int auto_reset, void *stream);

/* Same transition through HOST buffers (the end-to-end boundary), synchronous on `stream`.
 * Page-locked caller buffers (cudaHostAlloc / cudaHostRegister / torch pin_memory) are mapped
 * into the device's address space and used by the kernel directly: it reads the actions from
 * them and writes obs/reward/done into them over PCIe, no staging copy.  Pageable buffers go
 * through pinned staging buffers owned by the handle (H2D copy, launch, three D2H copies). */
int d2d_step_host(d2d_env *env, const float *actions_host, float *obs_host, float *reward_host,
                  uint8_t *done_host, int auto_reset, void *stream);
/* The same with the SB3 `terminal_observation` rows: terminal_obs_host (page-locked float32
 * [n][27], or NULL) receives, for every env that finished in this step, the last observation of
 * the finished episode; other rows are left untouched. */
int d2d_step_host_ex(d2d_env *env, const float *actions_host, float *obs_host, float *reward_host,
                     uint8_t *done_host, float *terminal_obs_host, int auto_reset, void *stream);

/* Optional float64 taps of the LAST d2d_step/d2d_reset, for parity tests and the
 * info dict (env:575-613).  Pass device pointers (or NULL to disable):
 *   obs64_dev  double [n][27]               observation before the f32 cast (terminal obs for finished envs)
 *   info_dev   double [n][D2D_INFO_DOUBLES]: reward, collision_avoidance_reward,
 *              path_adherence, path_progression, collision_reward, reach_end_reward,
 *              agressive_alpha_reward, env_steps, dist_closest_obs, APE, total_reward,
 *              end-condition bits (1 collision | 2 reached | 4 time | 8 alpha),
 *              u* (closest path parameter), Brent evaluations, nearest-obstacle
 *              indices packed (i0 + 256*i1 + 65536*i2, 255 = none), episode number */
int d2d_set_taps(d2d_env *env, double *obs64_dev, double *info_dev);
/* With enable != 0 the info tap receives an env's row only on the step that ends its episode (the
 * rows SB3 reads: infos of finished envs); 0 (default): every env, every step. */
int d2d_set_info_done_only(d2d_env *env, int enable);

/* Optional recording of the flight path (env:410-415 `add_postion_to_flight_path`, read back as
 * info["flight_path"] at env:612 and main.py:276): ring_dev float32 [ring_len][np][2] with np = n rounded up to a
 * multiple of 32; after every step the kernel stores (x, screen_y - y) of the frame at row (step counter - 1) %
 * ring_len, so the rows 0..T-1 of a column hold the T positions of the episode that just ended (its last ring_len
 * positions if T > ring_len) until the next step.  NULL switches the recording off (default). */
int d2d_set_flight_path_ring(d2d_env *env, float *ring_dev, int ring_len);

/* State blob (HOST, float64 [n][D2D_STATE_DOUBLES], sync): per env
 *   [0..17]  frame, left motor, right motor: px, py, vx, vy, angle, omega
 *   [18..29] cached pivot impulses jx,jy of left_1..3, right_1..3
 *   [30] collided  [31] look-ahead latched  [32] step counter
 *   [33] path-error sum  [34] reward sum  [35] episode number */
int d2d_get_state(d2d_env *env, double *state_host);
int d2d_set_state(d2d_env *env, const double *state_host);

/* Episode statistics accumulated on the device since the last call with
 * reset_after != 0 (HOST, double [D2D_STATS_DOUBLES], sync):
 *   episodes, successes, failures, collisions, sum env_steps, sum APE,
 *   sum total_reward, then the sums of the LAST-step info terms that
 *   tensorboardlogger.py:80-88 averages: reward, collision_reward,
 *   collision_avoidance_reward, path_adherence, path_progression,
 *   reach_end_reward, agressive_alpha_reward; [14] = steps taken, [15] = auto-resets that could
 *   not use a prefetched episode and sampled in the lane (diagnostic; ~5 per million episodes). */
int d2d_episode_stats(d2d_env *env, double *stats_host, int reset_after);
/* Device pointer of the same vector (for an NCCL all-reduce over ranks). */
int d2d_episode_stats_dev(d2d_env *env, double **stats_dev);

/* Debug tap of the closest-point search (QPMI2D.get_closest_u, predef_path.py:226-248)
 * on arbitrary positions: pos_dev double [n][2] -> u_dev double [n], nfev_dev int32 [n].
 * Env i uses its own path tables. */
int d2d_debug_closest_u(d2d_env *env, const double *pos_dev, double *u_dev, int32_t *nfev_dev, void *stream);

/* Policy in the loop (the caller of step in the reference's test mode: `PPO.load` main.py:251,
 * `model.predict` main.py:262): the SB3 MlpPolicy actor 27 -> 64 -> 64 -> 2 (tanh) evaluated on
 * the device in ONE launch, so observations and actions never leave HBM/L2 between steps.
 * The two hidden layers run on the tensor cores (tcgen05.mma kind::tf32, accumulators in TMEM,
 * operands split hi + lo and three products accumulated, so the result agrees with a float32
 * evaluation to ~3e-6); see csrc/drone2d_policy.cuh.
 *   d2d_policy_weights_bytes  size of the packed weight block (an image of the kernel's shared memory)
 *   d2d_policy_pack           HOST: pack `policy.pth` tensors (row-major float32: w0 [64][27], b0 [64],
 *                             w1 [64][64], b1 [64], w2 [2][64], b2 [2], log_std [2]) into the block
 *   d2d_policy_act            actions[n][2] = clip(mean(obs) + std * eps, -1, 1); eps ~ N(0,1) from the
 *                             Philox stream keyed by (seed, global env id, policy_step) if stochastic,
 *                             else 0 (deterministic `predict`).  packed_weights_dev: device copy of the block. */
int d2d_policy_weights_bytes(void);
int d2d_policy_pack(const float *w0, const float *b0, const float *w1, const float *b1, const float *w2,
                    const float *b2, const float *log_std, void *packed_host);
int d2d_policy_act(d2d_env *env, const void *packed_weights_dev, const float *obs_dev, float *actions_dev,
                   int stochastic, uint64_t policy_step, void *stream);

/* The same evaluation with everything a PPO rollout buffer stores (SB3 `policy.forward` inside
 * `collect_rollouts`, the loop main.py:190-208 runs): actions_dev receives the clipped action the env is stepped
 * with, actions_raw_dev (or NULL) the unclipped sample mean + std * eps, log_prob_dev (or NULL) its log-probability
 * under the diagonal Gaussian. */
int d2d_policy_act_ex(d2d_env *env, const void *packed_weights_dev, const float *obs_dev, float *actions_dev,
                      float *actions_raw_dev, float *log_prob_dev, int stochastic, uint64_t policy_step, void *stream);

/* The critic of the same MlpPolicy (value tower 27 -> 64 -> 64 -> 1) on n_rows observations (float32 [n_rows][27],
 * e.g. a whole rollout [T + 1][n]) in one launch of the tcgen05 kernel.  packed_value_weights_dev: d2d_policy_pack
 * of the tower with w2 = [value_net.weight; 0], b2 = [value_net.bias, 0], log_std = 0. */
int d2d_policy_value(d2d_env *env, const void *packed_value_weights_dev, const float *obs_dev, float *values_dev,
                     int64_t n_rows, void *stream);

/* Closed-loop rollout without the host in the loop: n_steps x (d2d_policy_act, d2d_step with
 * auto-reset) enqueued back to back on `stream`; obs_dev must hold the current observations on
 * entry and holds the final ones on return.  Episode statistics accumulate in the handle. */
int d2d_rollout(d2d_env *env, const void *packed_weights_dev, float *obs_dev, float *actions_dev,
                float *reward_dev, uint8_t *done_dev, int n_steps, int stochastic, uint64_t policy_step0,
                void *stream);

/* PPO rollout collection on the device (SB3 `collect_rollouts` + `compute_returns_and_advantage`, the loop the
 * reference's training mode runs, main.py:190-208): n_steps x (d2d_policy_act_ex, d2d_step with auto-reset), then the
 * critic over all stored observations and GAE, enqueued from C on `stream` -- no host work and no library kernel
 * between two steps; the kernels write straight into the caller's rollout tensors:
 *   obs_buf_dev   float32 [n_steps + 1][n][27]  row 0 = current observations on entry, row t + 1 written by step t
 *   actions_dev   float32 [n_steps][n][2]       unclipped samples        log_probs_dev float32 [n_steps][n]
 *   rewards_dev   float32 [n_steps][n]
 *   dones_dev     uint8   [n_steps + 1][n]      row 0 = episode starts of step 0 on entry, row t + 1 written by step t
 *                                               (rows 0..n_steps-1 are SB3's episode_starts, row n_steps its last dones)
 *   values_dev    float32 [n_steps + 1][n]      row n_steps = bootstrap values
 *   advantages_dev, returns_dev float32 [n_steps][n];  action_scratch_dev float32 [n][2] (clipped actions of the step) */
int d2d_rollout_collect(d2d_env *env, const void *actor_weights_dev, const void *critic_weights_dev, int n_steps,
                        float *obs_buf_dev, float *actions_dev, float *log_probs_dev, float *rewards_dev, uint8_t *dones_dev,
                        float *values_dev, float *advantages_dev, float *returns_dev, float *action_scratch_dev, float gamma,
                        float gae_lambda, uint64_t policy_step0, void *stream);

/* Curriculum scheduler hook: the reference derives the stage from the newest checkpoint file
 * name every time an env is rebuilt (env:79-85, 326-372); here the trainer sets it explicitly.
 * Takes effect for episodes sampled after the call (up to two already prefetched episodes per
 * env keep the stage they were drawn with). */
int d2d_set_stage(d2d_env *env, int stage, double spawn_chance);

/* Generalised advantage estimation of a device-resident rollout (the step after `step` in the
 * PPO loop the reference runs through SB3, main.py:190-208; SB3 RolloutBuffer semantics):
 * rewards/values/advantages/returns float32 [n_steps][n_envs], episode_starts uint8
 * [n_steps][n_envs], last_values float32 [n_envs], last_dones uint8 [n_envs].  One launch. */
int d2d_gae(const float *rewards_dev, const float *values_dev, const uint8_t *episode_starts_dev,
            const float *last_values_dev, const uint8_t *last_dones_dev, int n_steps, int64_t n_envs,
            float gamma, float gae_lambda, float *advantages_dev, float *returns_dev, void *stream);

/* Profiling hook: per-warp clock64() spans since the last reset, for role/phase r =
 * [0] a main lane's whole step, [1] the episode-sampling refill role, [2] the first-observation
 * refill role, [3] thrust + rigid-body solve, [4] obstacle pass, [5] observation incl. the
 * closest-point search, [6] reward / termination / swap; [7] unused.
 * HOST uint64 [24], sync: out[r] = max, out[8 + r] = sum, out[16 + r] = number of warps.
 * Recording is off until the first call (which returns zeros) so that ordinary runs do not pay
 * for the per-warp atomics. */
int d2d_debug_role_cycles(d2d_env *env, uint64_t *out24, int reset_after);

/* Replay mode for tests: with enable != 0 the step kernel skips thrust + cpSpaceStep and treats the
 * frame pose / velocity found in the state (d2d_set_state) as the outcome of space.step, i.e. it runs
 * env:407-615 only.  tests/test_gpu_reference_trajectories.py pushes the reference's genuine pymunk
 * trajectories through it and compares reward, APE and termination with what the reference wrote. */
int d2d_debug_skip_physics(d2d_env *env, int enable);

/* Number of kernels this library has launched on this handle. */
int64_t d2d_launch_count(const d2d_env *env);

const char *d2d_last_error(void);
int d2d_abi_version(void);

#ifdef __cplusplus
}
#endif
#endif /* DRONE2D_H */
