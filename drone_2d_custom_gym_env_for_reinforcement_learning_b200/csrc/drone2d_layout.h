// drone2d_layout.h -- device-side data layout of the batched Drone2dEnv state.
//
// Everything is structure-of-arrays over the environment index so that the 32
// lanes of a warp (one lane = one environment) touch 32 consecutive 16-byte
// elements: every state access is one fully coalesced 512-byte LDG.128/STG.128
// wavefront pair.  Pairs that are always used together are packed as double2.
//
//   pos  [3][N] double2  (px, py)      frame, left motor, right motor
//   vel  [3][N] double2  (vx, vy)
//   ang  [3][N] double2  (angle, omega)
//   jacc [6][N] double2  (jx, jy)      cached pivot impulses left_1..3, right_1..3
//   acc  [N]    double2  (path-error sum, reward sum)
//   meta [N]    int4     (step counter, flags, episode number, n_obs)
//   last [N]    double4-ish record of the last finished episode (2 x double2)
//
// Prefetched next episode ("staging slot", one per env): the re-spawn of env i
// for episode k+1 is a pure function of (seed, global env id, k+1), so it is
// computed ahead of time by dedicated refill warps of the step kernel and only
// swapped in when episode k ends (no sampling / closest-point search on the
// step's critical path):
//   stg_xy  [N] double2 (x, y)   stg_an [N] double2 (angle, unused)
//   stg_meta[N] int4    (episode the slot holds, n_obs, look-ahead latch, 0)
//   stg_obs [N][27] float        first observation of the staged episode
//   queue [2][3][N] int32 + qcount[2][3]: two-stage refill pipeline.  Envs whose slot was
//   consumed in step t are appended to queue A[t % 3]; the "sample" refill blocks of step
//   t+1 draw the episode (tables + spawn pose) and append to queue B[(t+1) % 3]; the
//   "observe" refill blocks of step t+2 compute its first observation and publish the
//   slot.  Each stage is shorter than the main lanes' step, so prefetching never extends
//   the kernel; a slot that is not ready when needed falls back to in-lane sampling.
//
// Per-env tables (curriculum mode only; rows of waypoint/segment index major so
// that lanes reading the same row coalesce).  Two buffers per env: the live
// episode (buffer 0) and the staged one (buffer 1, copied into buffer 0 when it is swapped in):
//   us   [MAX_WPS][N]        double
//   pa/pb/pc [MAX_WPS][N]    double2 (ax, ay) (bx, by) (cx, cy)   QPMI2D rows
//   obxy [MAX_OBS_ENV][N]    double2 ; obr [MAX_OBS_ENV][N] double
//   wps  [MAX_WPS][N]        double2 (x, y)
// The live block sits right behind the dynamic state (one contiguous "hot" prefix of the
// allocation); the staged block has the same layout `stg_delta` bytes further on.
#pragma once
#include <stddef.h>
#include <stdint.h>

#include "../../include/drone2d.h"

// Launch geometry of every per-env kernel (one lane per environment).
#define D2D_HINTS 4   // opening of the bounded Brent search: x0, x1, x2 (second point better), x2 (second point worse)
#ifndef D2D_BLOCK
#define D2D_BLOCK 128
#endif
#ifndef D2D_MIN_BLOCKS
#define D2D_MIN_BLOCKS 4
#endif

#if defined(__CUDACC__)
#define D2D_HD __host__ __device__
#else
#define D2D_HD
#endif

struct D2dLayout {
    int64_t n;          // environments
    int64_t np;         // padded to a multiple of 32 (row stride)
    size_t off_pos, off_vel, off_ang, off_jacc, off_acc, off_meta, off_last;
    size_t off_stg_xy, off_stg_an, off_stg_meta, off_stg_obs, off_queue, off_qcount;
    size_t off_us, off_wps, off_pa, off_pb, off_pc, off_hint, off_obxy, off_obr;
    size_t bytes;
    size_t hot_bytes;      // leading bytes touched by every step (state + live us/rows)
    size_t stg_delta;      // byte distance from a live per-env table to its staged copy
    int per_env_tables;
};

static inline size_t d2d_align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

static inline void d2d_make_layout(D2dLayout *L, int64_t n, int per_env_tables)
{
    L->n = n;
    L->np = (n + 31) / 32 * 32;
    L->per_env_tables = per_env_tables;
    size_t np = (size_t)L->np, o = 0;
    auto take = [&](size_t bytes) { size_t r = o; o = d2d_align_up(o + bytes, 256); return r; };
    // hot region first (touched by every step): dynamic state, then the LIVE per-env tables with the
    // most re-read arrays leading -- d2d_step marks [0, hot_bytes) as L2-persisting.
    L->off_pos = take(3 * np * 16);
    L->off_vel = take(3 * np * 16);
    L->off_ang = take(3 * np * 16);
    L->off_jacc = take(6 * np * 16);
    L->off_acc = take(np * 16);
    L->off_meta = take(np * 16);
    L->off_us = L->off_wps = L->off_pa = L->off_pb = L->off_pc = L->off_hint = L->off_obxy = L->off_obr = 0;
    L->stg_delta = 0;
    size_t live_begin = o;
    if (per_env_tables) {
        L->off_us = take((size_t)D2D_MAX_WPS * np * 8);
        L->off_pa = take((size_t)D2D_MAX_WPS * np * 16);
        L->off_pb = take((size_t)D2D_MAX_WPS * np * 16);
        L->off_pc = take((size_t)D2D_MAX_WPS * np * 16);
        L->off_hint = take((size_t)D2D_HINTS * np * 16);
    }
    L->hot_bytes = o;
    if (per_env_tables) {
        L->off_obxy = take((size_t)D2D_MAX_OBS_ENV * np * 16);
        L->off_obr = take((size_t)D2D_MAX_OBS_ENV * np * 8);
        L->off_wps = take((size_t)D2D_MAX_WPS * np * 16);
    }
    size_t live_end = o;
    L->off_last = take(2 * np * 16);
    L->off_stg_xy = take(np * 16);
    L->off_stg_an = take(np * 16);
    L->off_stg_meta = take(np * 16);
    L->off_stg_obs = take(np * D2D_OBS_DIM * 4);
    L->off_queue = take(3 * 3 * np * 4);       // A[3][np], B env[3][np], B target episode[3][np]
    L->off_qcount = take(256);
    if (per_env_tables) {
        // staged copy of the live table block, same internal layout, `stg_delta` bytes further on
        size_t stg_begin = take(live_end - live_begin);
        L->stg_delta = stg_begin - live_begin;
    }
    L->bytes = o;
}

// meta.y flag bits
#define D2D_F_COLLIDED 1
#define D2D_F_LATCHED 2
#define D2D_F_DONE 4      // terminal state kept (auto_reset == 0)
#define D2D_F_INIT 8      // spawned at least once (episode counter valid)

// One shared scenario (test mode): path + circles, read-only, staged into shared memory.
struct D2dSharedTables {
    int n_wps, n_obs;
    double us[D2D_MAX_WPS];
    double wx[D2D_MAX_WPS], wy[D2D_MAX_WPS];
    double ax[D2D_MAX_WPS], ay[D2D_MAX_WPS], bx[D2D_MAX_WPS], by[D2D_MAX_WPS], cx[D2D_MAX_WPS], cy[D2D_MAX_WPS];
    double ox[D2D_MAX_OBS_SHARED], oy[D2D_MAX_OBS_SHARED], orad[D2D_MAX_OBS_SHARED];
};

// Constants derived once on the host from d2d_config (Drone.py:9-95 topology).
struct D2dConst {
    // config copy (arithmetic keys)
    int mode, stage, n_steps, n_wps, corner_lo, corner_hi, random_path_spawn, use_Lambda, iterations;
    double W, H, seg_len, lookahead;
    double danger_range, danger_angle, k_ca;
    double PA_band_edge, PA_scale, PP_vel_scale, PP_rew_max, PP_rew_min;
    double rew_collision, reach_end_radius, rew_reach_end, AA_angle, AA_band, rew_AA;
    double spawn_chance, sp_xmin, sp_xmax, sp_ymin, sp_ymax;
    double gx, gy, dt, force_scale;
    // derived
    double R, d;                 // motor arm, joint spacing
    double hw, hh;               // frame box half extents (50, 5)
    double minv_f, iinv_f, minv_m, iinv_m;
    double msum;                 // minv_m + minv_f
    double bias_coef;            // -(1 - pow(0, dt)) / dt
    double diag;                 // sqrt(W^2 + H^2)
    // correctly rounded reciprocals of the constant divisors of the observation / reward (exact
    // division by a constant in 3 FMAs, see div_const): 2660 (velocity range), W, H, diag, 11.7, pi,
    // danger_range, PA_band_edge
    double r_vel, r_W, r_H, r_diag, r_omega, r_pi, r_DR, r_PA;
    uint32_t seed_lo, seed_hi;
    int64_t env_id_base;
};

// Kernel parameter block.
struct D2dParams {
    int64_t n, np;
    double2 *pos, *vel, *ang, *jacc, *acc, *last;
    int4 *meta;
    double *us;
    double2 *wps, *pa, *pb, *pc, *obxy;
    double2 *hint;                  // [D2D_HINTS][np]: path points of the opening evaluations of every closest-point search
    double *obr;
    uint32_t stg_off8, stg_off16;   // element offsets (8- / 16-byte elements) from a live table to its staged copy
    double2 *stg_xy, *stg_an;
    int4 *stg_meta;
    float *stg_obs;
    int *queue, *qcount;
    int refill_blocks;         // leading blocks of the step grid that sample staged episodes (sample role)
    int observe_blocks;        // blocks after them that compute the staged episodes' first observation (observe role)
    int use_hints;             // 1: the search's opening evaluations come from `hint` (0 only for the A/B test: same bits)
    const D2dSharedTables *shared;
    // I/O of one step
    const float *actions;
    float *obs, *reward, *terminal_obs;
    uint8_t *done;
    const uint8_t *mask;
    double *obs64, *info, *stats;
    float2 *path_ring;                 // optional [path_ring_len][np] (x, H - y) of the frame after every step (env:984-986), or NULL
    int path_ring_len;
    unsigned long long *role_cycles;   // [8] max clock64 span per role/phase of the last launches (profiling hook)
    int info_done_only;                // write an env's info row only on the step that ends its episode (d2d_set_info_done_only)
    int skip_physics;                  // debug: take the pose/velocity in the state as the result of space.step (d2d_debug_skip_physics)
    int probe;                         // record role_cycles (off until d2d_debug_role_cycles is first called)
    uint64_t step_index;       // global step counter (keys the random-action stream)
    int auto_reset;
    D2dConst c;
};
