/* oracle/d2d_oracle.h -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * Plain-C, float64, one-environment-at-a-time restatement of the reference's
 * Drone2dEnv.step hot path.  Only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py may use it.
 *
 * PARITY STATUS
 *   pinned   : path evaluation, bounded Brent (closest-u), look-ahead, scenario
 *              tables -- checked against tests/golden/ (vectors produced by the
 *              reference's own predef_path.py / test_scenarios.py, see
 *              tests/golden/make_golden.py).
 *   pinned at step level on GENUINE pymunk output: the rigid-body step.  The reference delegates it to pymunk
 *              (Chipmunk2D 7.0.x, un-vendored, un-pinned; neither importable nor on disk here, nor on the GPU
 *              box -- profiles/r02_box_probe.txt).  d2o_space_step restates Chipmunk's published cpSpaceStep /
 *              cpBodyUpdatePosition / cpBodyUpdateVelocity / cpPivotJoint algorithm for this world (3 bodies,
 *              6 pivots, 10 iterations), anchored on the reference's call sites (Drone.py:9-95,
 *              drone_2d_env.py:184-191,403-406).  It is checked against the frame positions the reference's own
 *              evaluation runs recorded after every space.step (700 episodes, 395 k steps): motor commands inside
 *              the action Box exist under which d2o_space_step reproduces every recorded velocity to ~1e-8 px/s
 *              (oracle/inverse_dynamics.py, tests/test_inverse_dynamics_cpu.py), while 9 or 11 solver
 *              iterations, no warm start, another joint order, another integration order, damping, a 2.5 %
 *              heavier motor or 0.1 % more gravity leave residuals 4-7 orders of magnitude larger.  What this
 *              cannot see: the angle and the motor bodies are not recorded, only the frame position is.
 *   pinned on genuine step() outputs (same fixtures): APE, termination step
 *              and reason, total reward of 70 reference episodes replayed
 *              through d2o_after_physics.
 */
#ifndef D2D_ORACLE_H
#define D2D_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

#define D2O_MAX_WPS 16
#define D2O_MAX_OBS 64
#define D2O_OBS_DIM 27

/* One dynamic body: position, velocity, angle, angular velocity. */
typedef struct {
    double px, py, vx, vy, a, w;
} d2o_body;

/* Everything the reference keeps between two step() calls of one episode. */
typedef struct {
    d2o_body body[3];   /* frame, left motor, right motor (Drone.py:30,44,58)   */
    double jx[6], jy[6];/* cached pivot impulses, left_1..3, right_1..3         */
    int collided;       /* space.collison (env:187, set by the begin callback)  */
    int la_latched;     /* LA_in_last_wp (env:112,739-741)                      */
    int steps;          /* current_time_step (env:139,407)                      */
    int pad_;
    double path_error;  /* env:132,589                                           */
    double total_reward;/* env:124,592                                           */
} d2o_state;

/* QPMI2D tables (predef_path.py:10-50). */
typedef struct {
    int n_wps;
    int pad_;
    double wx[D2O_MAX_WPS], wy[D2O_MAX_WPS];
    double us[D2O_MAX_WPS];
    double xp[D2O_MAX_WPS][3];  /* rows 0..n_wps-3 */
    double yp[D2O_MAX_WPS][3];
} d2o_path;

typedef struct {
    int n;
    int pad_;
    double x[D2O_MAX_OBS], y[D2O_MAX_OBS], r[D2O_MAX_OBS];
} d2o_obstacles;

/* rl_config.py:10-44 (the keys that reach the arithmetic) + Drone.py constants. */
typedef struct {
    double screen_x, screen_y;
    double lookahead;
    double danger_range, danger_angle, abs_inv_CA_min_rew;
    double PA_band_edge, PA_scale;
    double PP_vel_scale, PP_rew_max, PP_rew_min;
    double rew_collision, reach_end_radius, rew_reach_end;
    double AA_angle, AA_band, rew_AA;
    int use_Lambda;
    int n_steps;
    /* drone / space constants */
    double mass_f, mass_m;        /* 0.2, 0.4  (env:233 ...)                     */
    double width, height;         /* 100, 20                                     */
    double gravity_x, gravity_y;  /* 0, -1000 (env:185)                          */
    double dt;                    /* 1/60 (env:406)                              */
    double force_scale;           /* 1000 (env:150)                              */
    int iterations;               /* 10 (pymunk default)                         */
    int pad_;
} d2o_config;

/* Per-step outputs that tests compare (superset of the reference's info dict). */
typedef struct {
    double obs[D2O_OBS_DIM];
    double reward;
    double r_CA, r_PA, r_PP, r_coll, r_end, r_AA; /* info terms, already lambda-weighted (env:575-581) */
    double dist_closest_obs;                       /* env:584-587 (inf if none) */
    double ape;                                    /* env:590 */
    double u_star;                                 /* closest-u of this step */
    double closest_x, closest_y, la_x, la_y;
    int nfev;                                      /* Brent evaluations */
    int near_idx[3];                               /* nearest-k obstacle indices, -1 if absent */
    int end_collision, end_reached, end_time, end_alpha;
    int done;
    int pad_;
} d2o_step_out;

void d2o_default_config(d2o_config *c);
double d2o_moment_for_box(double m, double w, double h);

/* Drone.__init__ (Drone.py:9-95): bodies at rest, impulses cleared. */
void d2o_spawn(d2o_state *s, double x, double y, double angle, const d2o_config *c);

/* env:400-406: thrust on the frame body + cpSpaceStep(dt). Returns 1 if the
 * frame box touches an obstacle at the new pose (also ORed into s->collided). */
int d2o_space_step(d2o_state *s, const d2o_obstacles *ob, const d2o_config *c,
                   double action_left, double action_right);

/* predef_path.py:88-142 / 145-188 / 226-248 */
void d2o_path_eval(const d2o_path *p, double u, double *x, double *y);
void d2o_path_gradient(const d2o_path *p, double u, double *dx, double *dy);
double d2o_closest_u(const d2o_path *p, double px, double py, int *nfev, double *fval);

/* env:617-629 + 948-961: distances (all) and the stable nearest-k. */
int d2o_nearest(const d2o_state *s, const d2o_obstacles *ob, const d2o_config *c,
                int k, int *idx, double *dist);

/* env:631-773 (mutates the look-ahead latch like the reference). */
void d2o_observe(d2o_state *s, const d2o_path *p, const d2o_obstacles *ob,
                 const d2o_config *c, d2o_step_out *o);

/* env:394-615: one full transition (no auto-reset). */
void d2o_env_step(d2o_state *s, const d2o_path *p, const d2o_obstacles *ob,
                  const d2o_config *c, double action_left, double action_right,
                  d2o_step_out *o);

/* env:421-615: the part of step() after space.step() (observation, reward, termination,
 * accumulators) on the state as it stands; d2o_env_step = d2o_space_step + steps++ + this. */
void d2o_after_physics(d2o_state *s, const d2o_path *p, const d2o_obstacles *ob,
                       const d2o_config *c, d2o_step_out *o);

/* The begin-callback condition of env:17-19,190-191 at a given frame pose: 1 iff the frame box
 * (width x height/2, Drone.py:16) touches any obstacle circle. */
int d2o_frame_touches(const d2o_body *frame, const d2o_obstacles *ob, const d2o_config *c);

/* Forward model of the inverse-dynamics discriminator: frame velocity after each of n steps from *s0
 * (unmodified) under actions [n][2]; jac (or NULL) receives the forward-difference Jacobian [2n][2n]. */
void d2o_frame_velocity_rollout(const d2o_state *s0, const double *actions, int n, const d2o_config *c,
                                double *v_out, double *jac, double h);

/* Batch helpers (AoS arrays of n envs; path/obstacles either shared (stride 0)
 * or per env (stride 1)); actions float32 [n,2] like the gym boundary. */
void d2o_env_step_batch(int n, d2o_state *s, const d2o_path *p, int path_stride,
                        const d2o_obstacles *ob, int ob_stride, const d2o_config *c,
                        const float *actions, d2o_step_out *o);

/* Philox4x32-10 (Salmon et al., SC'11), the counter-based generator the
 * auto-reset is specified on; exported for the known-answer test. */
void d2o_philox4x32_10(const unsigned ctr[4], const unsigned key[2], unsigned out[4]);

#ifdef __cplusplus
}
#endif
#endif
