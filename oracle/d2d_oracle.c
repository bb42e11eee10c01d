/* oracle/d2d_oracle.c -- TEST INFRASTRUCTURE ONLY (see d2d_oracle.h for the
 * parity status: path/Brent pinned by tests/golden, rigid-body step UNPINNED).
 *
 * CPU restatement (float64, strict IEEE: build with -ffp-contract=off, no
 * -ffast-math) of the reference hot path.  Citations are file:line relative to
 * /root/reference/drone_2d_custom_gym_env/ ("env" = drone_2d_env.py).  Third
 * party arithmetic that is not in the reference tree is restated from its
 * published algorithm:
 *   - Chipmunk2D 7.0.x (via pymunk 6.x, unpinned): cpSpaceStep.c, cpBody.c,
 *     cpPivotJoint.c, chipmunk_private.h (k_tensor, apply_impulses,
 *     relative_velocity), cpPolyShape.c (cpMomentForPoly), cpCollision.c
 *     (circle-vs-poly test d <= r).
 *   - scipy.optimize.fminbound -> _minimize_scalar_bounded
 *     (scipy/optimize/_optimize.py:2289-2440, scipy 1.18.1).
 *   - Philox4x32-10 (Random123) for the counter-based reset stream.
 */
#include "d2d_oracle.h"

#include <math.h>
#include <string.h>

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

/* pymunk Space.damping: never set by the reference (default 1.0).  env:376-380 assigns a plain Python
 * attribute `body.damping = 0.9`, which Chipmunk never sees; tests/test_reference_trajectories_cpu.py
 * (test_no_damping_in_the_genuine_runs) and tests/test_inverse_dynamics_cpu.py check this against genuine pymunk
 * trajectories by rebuilding the oracle with -DD2O_SPACE_DAMPING=0.9. */
/* Further compile-time variants of the solver, used ONLY by the inverse-dynamics discriminator
 * (tests/test_inverse_dynamics_cpu.py) to show that the genuine pymunk trajectories reject them:
 *   -DD2O_NO_WARM_START          cached impulses are not re-applied at the start of a step
 *   -DD2O_JOINT_ORDER_PERMUTED   joints solved right_1..3 before left_1..3 (insertion order swapped)
 *   -DD2O_POSITIONS_AFTER        positions integrated after the velocity solve (symplectic Euler order) */
#ifndef D2O_SPACE_DAMPING
#define D2O_SPACE_DAMPING 1.0
#endif

/* ------------------------------------------------------------------------- */
/* small helpers                                                             */
/* ------------------------------------------------------------------------- */

/* Python / numpy float modulo (result takes the sign of the divisor). */
static double py_mod(double a, double b)
{
    double m = fmod(a, b);
    if (m != 0.0) {
        if ((b < 0.0) != (m < 0.0)) m += b;
    } else {
        m = copysign(0.0, b);
    }
    return m;
}

/* transformations.py:6-7 */
static double ssa(double angle) { return py_mod(angle + M_PI, 2 * M_PI) - M_PI; }

/* env:972-978 */
static double m1to1(double v, double lo, double hi) { return 2.0 * (v - lo) / (hi - lo) - 1; }
static double invm1to1(double v, double lo, double hi) { return (v + 1) * (hi - lo) / 2.0 + lo; }

static double clipd(double v, double lo, double hi) { return v < lo ? lo : (v > hi ? hi : v); }

/* np.linalg.norm of a 2-vector = sqrt(x.dot(x)).  numpy hands the dot to BLAS
 * ddot, whose x86-64 kernels accumulate with a fused multiply-add: the value is
 * sqrt(fma(dy, dy, dx*dx)).  With this form d2o_closest_u reproduces scipy's
 * fminbound on the golden vectors bit-for-bit (295 of 296 cases; the plain
 * dx*dx+dy*dy form only 237 -- Brent's last iterations compare objective
 * values that differ by rounding noise, so u* is only determined to ~1e-5). */
static double norm2(double dx, double dy) { return sqrt(fma(dy, dy, dx * dx)); }

void d2o_default_config(d2o_config *c)
{
    /* rl_config.py:10-44 */
    memset(c, 0, sizeof *c);
    c->screen_x = 1300; c->screen_y = 1300;
    c->lookahead = 220;
    c->danger_range = 150; c->danger_angle = 20; c->abs_inv_CA_min_rew = 1.0 / 8;
    c->PA_band_edge = 40; c->PA_scale = 2;
    c->PP_vel_scale = 0.08; c->PP_rew_max = 2.5; c->PP_rew_min = -1;
    c->rew_collision = -50; c->reach_end_radius = 20; c->rew_reach_end = 30;
    c->AA_angle = M_PI / 2; c->AA_band = M_PI / 4; c->rew_AA = -1;
    c->use_Lambda = 1; c->n_steps = 1100;
    c->mass_f = 0.2; c->mass_m = 0.4; c->width = 100; c->height = 20;
    c->gravity_x = 0; c->gravity_y = -1000;
    c->dt = 1.0 / 60; c->force_scale = 1000; c->iterations = 10;
}

/* cpMomentForPoly on the 4 box vertices that Poly.create_box produces
 * (Drone.py:16-17,33-34).  w,h are the FULL box sizes. */
double d2o_moment_for_box(double m, double w, double h)
{
    double hw = w / 2, hh = h / 2;
    double vx[4] = { hw, hw, -hw, -hw };
    double vy[4] = { -hh, hh, hh, -hh };
    double sum1 = 0.0, sum2 = 0.0;
    for (int i = 0; i < 4; i++) {
        int j = (i + 1) % 4;
        double a = vx[j] * vy[i] - vy[j] * vx[i];
        double b = (vx[i] * vx[i] + vy[i] * vy[i]) + (vx[i] * vx[j] + vy[i] * vy[j]) +
                   (vx[j] * vx[j] + vy[j] * vy[j]);
        sum1 += a * b;
        sum2 += a;
    }
    return (m * sum1) / (6.0 * sum2);
}

/* ------------------------------------------------------------------------- */
/* drone topology (Drone.py:9-95)                                            */
/* ------------------------------------------------------------------------- */

typedef struct {
    double R, d;             /* drone_radius = w/2 - h/2, joint spacing = h/2 - 3 */
    double minv[3], iinv[3];
    double ax[6], ay[6];     /* anchor on the motor (body a of the joint)         */
    double bx[6], by[6];     /* anchor on the frame (body b of the joint)         */
    int motor[6];            /* 1 = left, 2 = right                               */
} topo_t;

static void make_topo(const d2o_config *c, topo_t *t)
{
    t->R = c->width / 2 - c->height / 2;       /* Drone.py:11 */
    t->d = c->height / 2 - 3;                  /* Drone.py:10 */
    double i_f = d2o_moment_for_box(c->mass_f, c->width, c->height / 2);  /* Drone.py:16-17 */
    double i_m = d2o_moment_for_box(c->mass_m, c->height, c->height);     /* Drone.py:33-34 */
    t->minv[0] = 1.0 / c->mass_f; t->iinv[0] = 1.0 / i_f;
    t->minv[1] = t->minv[2] = 1.0 / c->mass_m;
    t->iinv[1] = t->iinv[2] = 1.0 / i_m;
    /* Drone.py:61-95 */
    double mx[3] = { -t->d, 0.0, t->d };
    for (int k = 0; k < 3; k++) {
        t->motor[k] = 1; t->ax[k] = mx[k]; t->ay[k] = 0; t->bx[k] = -t->R + mx[k]; t->by[k] = 0;
        t->motor[3 + k] = 2; t->ax[3 + k] = mx[k]; t->ay[3 + k] = 0; t->bx[3 + k] = t->R + mx[k]; t->by[3 + k] = 0;
    }
}

void d2o_spawn(d2o_state *s, double x, double y, double angle, const d2o_config *c)
{
    topo_t t; make_topo(c, &t);
    memset(s, 0, sizeof *s);
    /* Drone.py:20-21, 37-38, 51-52 */
    s->body[0].px = x; s->body[0].py = y; s->body[0].a = angle;
    s->body[1].px = cos(angle + M_PI) * t.R + x; s->body[1].py = sin(angle + M_PI) * t.R + y; s->body[1].a = angle;
    s->body[2].px = cos(angle) * t.R + x; s->body[2].py = sin(angle) * t.R + y; s->body[2].a = angle;
}

/* ------------------------------------------------------------------------- */
/* cpSpaceStep specialised to this world                                     */
/* ------------------------------------------------------------------------- */

/* Circle-vs-box narrow phase: Chipmunk collides iff the distance between the
 * circle centre and the (solid, radius-0) polygon is <= r. */
static int box_touches_circle(const d2o_body *f, double rc, double rs, double hw, double hh,
                              double cx, double cy, double r)
{
    double dx = cx - f->px, dy = cy - f->py;
    double qx = rc * dx + rs * dy;       /* into the box frame */
    double qy = -rs * dx + rc * dy;
    double ex = fabs(qx) - hw; if (ex < 0) ex = 0;
    double ey = fabs(qy) - hh; if (ey < 0) ey = 0;
    return sqrt(ex * ex + ey * ey) <= r;
}

int d2o_frame_touches(const d2o_body *frame, const d2o_obstacles *ob, const d2o_config *c)
{
    int hit = 0;
    double rc = cos(frame->a), rs = sin(frame->a);
    double hw = c->width / 2, hh = (c->height / 2) / 2;   /* box size (width, height/2), Drone.py:16 */
    for (int k = 0; k < ob->n; k++)
        if (box_touches_circle(frame, rc, rs, hw, hh, ob->x[k], ob->y[k], ob->r[k])) hit = 1;
    return hit;
}

int d2o_space_step(d2o_state *s, const d2o_obstacles *ob, const d2o_config *c,
                   double action_left, double action_right)
{
    topo_t t; make_topo(c, &t);
    const double dt = c->dt;
    d2o_body *B = s->body;

    /* env:400-404.  numpy 1.26 (training-time) promotes float32-scalar op
     * python-scalar to float64, so the thrust is evaluated in double. */
    double FL = (action_left / 2 + 0.5) * c->force_scale;
    double FR = (action_right / 2 + 0.5) * c->force_scale;

    /* cpBodyApplyForceAtLocalPoint on the frame with its CURRENT transform. */
    double fx = 0, fy = 0, tq = 0;
    {
        double rx = cos(B[0].a), ry = sin(B[0].a);
        double F[2] = { FL, FR };
        double lx[2] = { -t.R, t.R };
        for (int k = 0; k < 2; k++) {
            /* cpTransformVect(transform, (0,F)) */
            double fwx = rx * 0.0 + (-ry) * F[k];
            double fwy = ry * 0.0 + rx * F[k];
            /* cpTransformPoint(transform, (lx,0)) ; cog = 0 */
            double pwx = rx * lx[k] + (-ry) * 0.0 + B[0].px;
            double pwy = ry * lx[k] + rx * 0.0 + B[0].py;
            double cgx = rx * 0.0 + (-ry) * 0.0 + B[0].px;
            double cgy = ry * 0.0 + rx * 0.0 + B[0].py;
            double r_x = pwx - cgx, r_y = pwy - cgy;
            fx = fx + fwx; fy = fy + fwy;
            tq += r_x * fwy - r_y * fwx;
        }
    }

    /* 1. cpBodyUpdatePosition for bodies in insertion order. */
    double rc[3], rs[3];
    for (int i = 0; i < 3; i++) {
#ifndef D2O_POSITIONS_AFTER
        B[i].px = B[i].px + (B[i].vx + 0.0) * dt;
        B[i].py = B[i].py + (B[i].vy + 0.0) * dt;
        B[i].a = B[i].a + (B[i].w + 0.0) * dt;
#endif
        rc[i] = cos(B[i].a); rs[i] = sin(B[i].a);
    }

    /* 2. collision detection at the new poses: only the frame box (type 1)
     * against obstacle circles (type 2) has a handler (env:17-19,190-191);
     * everything is a sensor so nothing is ever solved. */
    int hit = d2o_frame_touches(&B[0], ob, c);
    if (hit) s->collided = 1;

    /* 4. cpPivotJoint preStep for the 6 joints. */
    double r1x[6], r1y[6], r2x[6], r2y[6], k11[6], k12[6], k21[6], k22[6], bsx[6], bsy[6];
    for (int j = 0; j < 6; j++) {
        int m = t.motor[j];
        r1x[j] = rc[m] * t.ax[j] + (-rs[m]) * t.ay[j];
        r1y[j] = rs[m] * t.ax[j] + rc[m] * t.ay[j];
        r2x[j] = rc[0] * t.bx[j] + (-rs[0]) * t.by[j];
        r2y[j] = rs[0] * t.bx[j] + rc[0] * t.by[j];
        /* k_tensor(a = motor, b = frame) */
        double m_sum = t.minv[m] + t.minv[0];
        double a11 = m_sum, a12 = 0.0, a21 = 0.0, a22 = m_sum;
        double ai = t.iinv[m];
        double r1xsq = r1x[j] * r1x[j] * ai, r1ysq = r1y[j] * r1y[j] * ai, r1nxy = -r1x[j] * r1y[j] * ai;
        a11 += r1ysq; a12 += r1nxy; a21 += r1nxy; a22 += r1xsq;
        double bi = t.iinv[0];
        double r2xsq = r2x[j] * r2x[j] * bi, r2ysq = r2y[j] * r2y[j] * bi, r2nxy = -r2x[j] * r2y[j] * bi;
        a11 += r2ysq; a12 += r2nxy; a21 += r2nxy; a22 += r2xsq;
        double det = a11 * a22 - a12 * a21;
        double det_inv = 1.0 / det;
        k11[j] = a22 * det_inv; k12[j] = -a12 * det_inv;
        k21[j] = -a21 * det_inv; k22[j] = a11 * det_inv;
        /* bias: error_bias = 0 -> bias_coef = 1 - pow(0, dt) = 1; maxBias = inf */
        double dx = (B[0].px + r2x[j]) - (B[m].px + r1x[j]);
        double dy = (B[0].py + r2y[j]) - (B[m].py + r1y[j]);
        double coef = -(1.0 - pow(0.0, dt)) / dt;
        bsx[j] = dx * coef; bsy[j] = dy * coef;
    }

    /* 5. cpBodyUpdateVelocity, damping = pow(1, dt) = 1 (env:376-380 sets a
     * plain Python attribute and has no physical effect). */
    {
        double damping = pow(D2O_SPACE_DAMPING, dt);
        double f_x[3] = { fx, 0, 0 }, f_y[3] = { fy, 0, 0 }, tt[3] = { tq, 0, 0 };
        for (int i = 0; i < 3; i++) {
            B[i].vx = B[i].vx * damping + (c->gravity_x + f_x[i] * t.minv[i]) * dt;
            B[i].vy = B[i].vy * damping + (c->gravity_y + f_y[i] * t.minv[i]) * dt;
            B[i].w = B[i].w * damping + tt[i] * t.iinv[i] * dt;
        }
    }

    /* 6. warm start (dt_coef = dt/prev_dt = 1; on the first step every jAcc is 0). */
#ifdef D2O_NO_WARM_START
    for (int j = 0; j < 6; j++) { s->jx[j] = 0; s->jy[j] = 0; }
#endif
    for (int j = 0; j < 6; j++) {
        int m = t.motor[j];
        double jx = s->jx[j] * 1.0, jy = s->jy[j] * 1.0;
        B[m].vx = B[m].vx + (-jx) * t.minv[m]; B[m].vy = B[m].vy + (-jy) * t.minv[m];
        B[m].w += t.iinv[m] * (r1x[j] * (-jy) - r1y[j] * (-jx));
        B[0].vx = B[0].vx + jx * t.minv[0]; B[0].vy = B[0].vy + jy * t.minv[0];
        B[0].w += t.iinv[0] * (r2x[j] * jy - r2y[j] * jx);
    }

    /* 7. sequential impulses. */
    for (int it = 0; it < c->iterations; it++) {
        for (int jj = 0; jj < 6; jj++) {
#ifdef D2O_JOINT_ORDER_PERMUTED
            const int j = (jj + 3) % 6;
#else
            const int j = jj;
#endif
            int m = t.motor[j];
            /* relative_velocity(a, b, r1, r2) */
            double v1x = B[m].vx + (-r1y[j]) * B[m].w, v1y = B[m].vy + r1x[j] * B[m].w;
            double v2x = B[0].vx + (-r2y[j]) * B[0].w, v2y = B[0].vy + r2x[j] * B[0].w;
            double vrx = v2x - v1x, vry = v2y - v1y;
            double ex = bsx[j] - vrx, ey = bsy[j] - vry;
            double jx = ex * k11[j] + ey * k12[j];
            double jy = ex * k21[j] + ey * k22[j];
            double ox = s->jx[j], oy = s->jy[j];
            s->jx[j] = ox + jx; s->jy[j] = oy + jy;      /* clamp to maxForce*dt = inf: no-op */
            jx = s->jx[j] - ox; jy = s->jy[j] - oy;
            B[m].vx = B[m].vx + (-jx) * t.minv[m]; B[m].vy = B[m].vy + (-jy) * t.minv[m];
            B[m].w += t.iinv[m] * (r1x[j] * (-jy) - r1y[j] * (-jx));
            B[0].vx = B[0].vx + jx * t.minv[0]; B[0].vy = B[0].vy + jy * t.minv[0];
            B[0].w += t.iinv[0] * (r2x[j] * jy - r2y[j] * jx);
        }
    }
#ifdef D2O_POSITIONS_AFTER
    for (int i = 0; i < 3; i++) {
        B[i].px = B[i].px + B[i].vx * dt; B[i].py = B[i].py + B[i].vy * dt; B[i].a = B[i].a + B[i].w * dt;
    }
#endif
    return hit;
}

/* ------------------------------------------------------------------------- */
/* QPMI2D (predef_path.py)                                                   */
/* ------------------------------------------------------------------------- */

/* predef_path.py:53-63 */
static int u_index(const d2o_path *p, double u)
{
    int n = 0;
    while (n < p->n_wps - 1) {
        if (u <= p->us[n + 1]) break;
        n++;
    }
    return n;
}

static double poly2(const double q[3], double u) { return q[0] * (u * u) + q[1] * u + q[2]; }

/* predef_path.py:88-142 */
void d2o_path_eval(const d2o_path *p, double u, double *x, double *y)
{
    int N = p->n_wps, rows = N - 2;
    if (u >= p->us[0] && u <= p->us[1]) {
        *x = poly2(p->xp[0], u); *y = poly2(p->yp[0], u);
    } else if ((u >= p->us[N - 2] - 0.001 && u <= p->us[N - 1]) || u_index(p, u) == N - 1) {
        *x = poly2(p->xp[rows - 1], u); *y = poly2(p->yp[rows - 1], u);
    } else {
        int n = u_index(p, u);
        double mu_r = (u - p->us[n]) / (p->us[n + 1] - p->us[n]);
        double mu_f = (p->us[n + 1] - u) / (p->us[n + 1] - p->us[n]);
        int r1 = n - 1; if (r1 < 0) r1 += rows;   /* Python negative index (u < 0) */
        double x1 = poly2(p->xp[r1], u), y1 = poly2(p->yp[r1], u);
        double x2 = poly2(p->xp[n], u), y2 = poly2(p->yp[n], u);
        *x = mu_r * x2 + mu_f * x1;
        *y = mu_r * y2 + mu_f * y1;
    }
}

/* predef_path.py:145-188 */
void d2o_path_gradient(const d2o_path *p, double u, double *dx, double *dy)
{
    int N = p->n_wps, rows = N - 2;
    if (u >= p->us[0] && u <= p->us[1]) {
        *dx = p->xp[0][0] * u * 2 + p->xp[0][1]; *dy = p->yp[0][0] * u * 2 + p->yp[0][1];
    } else if (u >= p->us[N - 2]) {
        *dx = p->xp[rows - 1][0] * u * 2 + p->xp[rows - 1][1];
        *dy = p->yp[rows - 1][0] * u * 2 + p->yp[rows - 1][1];
    } else {
        int n = u_index(p, u);
        double mu_r = (u - p->us[n]) / (p->us[n + 1] - p->us[n]);
        double mu_f = (p->us[n + 1] - u) / (p->us[n + 1] - p->us[n]);
        int r1 = n - 1; if (r1 < 0) r1 += rows;
        double dx1 = p->xp[r1][0] * u * 2 + p->xp[r1][1], dy1 = p->yp[r1][0] * u * 2 + p->yp[r1][1];
        double dx2 = p->xp[n][0] * u * 2 + p->xp[n][1], dy2 = p->yp[n][0] * u * 2 + p->yp[n][1];
        *dx = mu_r * dx2 + mu_f * dx1;
        *dy = mu_r * dy2 + mu_f * dy1;
    }
}

static double path_dist(const d2o_path *p, double u, double px, double py)
{
    double x, y; d2o_path_eval(p, u, &x, &y);
    double dx = x - px, dy = y - py;
    return norm2(dx, dy);                    /* np.linalg.norm, predef_path.py:246 */
}

static double signd(double v) { return (v > 0) - (v < 0); }

/* predef_path.py:226-248 -> scipy _minimize_scalar_bounded(xatol=1e-6, maxiter=500)
 * on [-10, L+10]. */
double d2o_closest_u(const d2o_path *p, double px, double py, int *nfev, double *fval)
{
    const double xatol = 1e-6; const int maxfun = 500;
    double x1 = 0.0 - 10.0, x2 = p->us[p->n_wps - 1] + 10.0;
    double sqrt_eps = sqrt(2.2e-16);
    double golden_mean = 0.5 * (3.0 - sqrt(5.0));
    double a = x1, b = x2;
    double fulc = a + golden_mean * (b - a);
    double nfc = fulc, xf = fulc;
    double rat = 0.0, e = 0.0;
    double x = xf;
    double fx = path_dist(p, x, px, py);
    int num = 1;
    double fu;
    double ffulc = fx, fnfc = fx;
    double xm = 0.5 * (a + b);
    double tol1 = sqrt_eps * fabs(xf) + xatol / 3.0;
    double tol2 = 2.0 * tol1;

    while (fabs(xf - xm) > (tol2 - 0.5 * (b - a))) {
        int golden = 1;
        if (fabs(e) > tol1) {
            golden = 0;
            double r = (xf - nfc) * (fx - ffulc);
            double q = (xf - fulc) * (fx - fnfc);
            double pp = (xf - fulc) * q - (xf - nfc) * r;
            q = 2.0 * (q - r);
            if (q > 0.0) pp = -pp;
            q = fabs(q);
            r = e;
            e = rat;
            if ((fabs(pp) < fabs(0.5 * q * r)) && (pp > q * (a - xf)) && (pp < q * (b - xf))) {
                rat = (pp + 0.0) / q;
                x = xf + rat;
                if (((x - a) < tol2) || ((b - x) < tol2)) {
                    double si = signd(xm - xf) + ((xm - xf) == 0);
                    rat = tol1 * si;
                }
            } else {
                golden = 1;
            }
        }
        if (golden) {
            if (xf >= xm) e = a - xf; else e = b - xf;
            rat = golden_mean * e;
        }
        double si = signd(rat) + (rat == 0);
        x = xf + si * fmax(fabs(rat), tol1);
        fu = path_dist(p, x, px, py);
        num += 1;

        if (fu <= fx) {
            if (x >= xf) a = xf; else b = xf;
            fulc = nfc; ffulc = fnfc;
            nfc = xf; fnfc = fx;
            xf = x; fx = fu;
        } else {
            if (x < xf) a = x; else b = x;
            if ((fu <= fnfc) || (nfc == xf)) {
                fulc = nfc; ffulc = fnfc;
                nfc = x; fnfc = fu;
            } else if ((fu <= ffulc) || (fulc == xf) || (fulc == nfc)) {
                fulc = x; ffulc = fu;
            }
        }
        xm = 0.5 * (a + b);
        tol1 = sqrt_eps * fabs(xf) + xatol / 3.0;
        tol2 = 2.0 * tol1;
        if (num >= maxfun) break;
    }
    if (nfev) *nfev = num;
    if (fval) *fval = fx;
    return xf;
}

/* ------------------------------------------------------------------------- */
/* observation (env:617-773)                                                 */
/* ------------------------------------------------------------------------- */

/* env:948-961: min over the 4 UNROTATED box vertices of |v + pos - centre| - r */
static double corner_distance(const d2o_body *f, double hw, double hh, double cx, double cy, double r)
{
    /* Poly.create_box vertex order: (hw,-hh), (hw,hh), (-hw,hh), (-hw,-hh) */
    double vx[4] = { hw, hw, -hw, -hw }, vy[4] = { -hh, hh, hh, -hh };
    double best = INFINITY;
    for (int i = 0; i < 4; i++) {
        double wx = vx[i] + f->px, wy = vy[i] + f->py;
        double dx = wx - cx, dy = wy - cy;
        double d = sqrt(dx * dx + dy * dy) - r;
        if (d < best) best = d;
    }
    return best;
}

/* env:617-629: stable ascending sort by distance, first k. */
int d2o_nearest(const d2o_state *s, const d2o_obstacles *ob, const d2o_config *c,
                int k, int *idx, double *dist)
{
    double hw = c->width / 2, hh = (c->height / 2) / 2;
    double d[D2O_MAX_OBS]; int order[D2O_MAX_OBS];
    for (int i = 0; i < ob->n; i++) {
        d[i] = corner_distance(&s->body[0], hw, hh, ob->x[i], ob->y[i], ob->r[i]);
        order[i] = i;
    }
    for (int i = 1; i < ob->n; i++) {            /* insertion sort: stable */
        int o = order[i], j = i - 1;
        while (j >= 0 && d[order[j]] > d[o]) { order[j + 1] = order[j]; j--; }
        order[j + 1] = o;
    }
    if (k > ob->n) k = ob->n;
    for (int i = 0; i < k; i++) { idx[i] = order[i]; dist[i] = d[order[i]]; }
    return k;
}

void d2o_observe(d2o_state *s, const d2o_path *p, const d2o_obstacles *ob,
                 const d2o_config *c, d2o_step_out *o)
{
    const d2o_body *f = &s->body[0];
    const double W = c->screen_x, H = c->screen_y;
    double *obs = o->obs;
    double x_t = p->wx[p->n_wps - 1], y_t = p->wy[p->n_wps - 1];   /* env:316-317 */

    obs[0] = m1to1(f->vx, -1330, 1330);               /* env:633-636 */
    obs[1] = m1to1(f->vy, -1330, 1330);
    obs[2] = clipd(f->w / 11.7, -1, 1);               /* env:639-640 */
    obs[3] = f->a / M_PI;                             /* env:643-644 */
    double x = f->px, y = f->py;
    obs[4] = m1to1(x_t - x, 0, W);                    /* env:648-649 */
    obs[5] = m1to1(y_t - y, 0, H);
    obs[6] = m1to1(x, 0, W);                          /* env:655-656 */
    obs[7] = m1to1(y, 0, H);
    double alpha = f->a;

    for (int j = 0; j < 3; j++) { obs[8 + 3 * j] = 1; obs[9 + 3 * j] = 0; obs[10 + 3 * j] = 0; o->near_idx[j] = -1; }
    if (ob->n > 0) {                                  /* env:670-720 */
        int idx[3]; double dist[3];
        int k = d2o_nearest(s, ob, c, 3, idx, dist);
        double diag = sqrt(W * W + H * H);
        for (int j = 0; j < k; j++) {
            double ang = atan2(y - ob->y[idx[j]], x - ob->x[idx[j]]);
            ang = ssa(ang - alpha - M_PI);
            obs[8 + 3 * j] = m1to1(dist[j], 0, diag);
            obs[9 + 3 * j] = sin(ang);
            obs[10 + 3 * j] = cos(ang);
            o->near_idx[j] = idx[j];
        }
    }

    /* env:723-727 */
    double va = ssa(atan2(f->vy, f->vx) - alpha);
    obs[17] = sin(va); obs[18] = cos(va);

    /* env:730-734: closest point (the second, identical fminbound of env:737 is
     * evaluated once here; same arguments -> same result). */
    int nfev = 0;
    double u = d2o_closest_u(p, x, y, &nfev, 0);
    double cx, cy; d2o_path_eval(p, u, &cx, &cy);
    obs[19] = m1to1(cx, 0, W); obs[20] = m1to1(cy, 0, H);
    o->u_star = u; o->nfev = nfev; o->closest_x = cx; o->closest_y = cy;

    /* env:737-749 + predef_path.py:257-266 */
    double L = p->us[p->n_wps - 1];
    double u_la = (u + c->lookahead > L) ? L : u + c->lookahead;
    double lx, ly; d2o_path_eval(p, u_la, &lx, &ly);
    if (fabs(lx - x_t) < 10 && fabs(ly - y_t) < 10) s->la_latched = 1;
    if (s->la_latched) { lx = x_t; ly = y_t; }
    obs[21] = m1to1(lx, 0, W); obs[22] = m1to1(ly, 0, H);
    o->la_x = lx; o->la_y = ly;

    /* env:752-763; R_w_b rotates by +alpha (transformations.py:9-11) */
    double ca = cos(alpha), sa = sin(alpha);
    {
        double dx = lx - x, dy = ly - y;
        double bx = ca * dx + (-sa) * dy, by = sa * dx + ca * dy;
        double ang = ssa(atan2(by, bx) - alpha);
        obs[23] = sin(ang); obs[24] = cos(ang);
    }
    {
        double dx = cx - x, dy = cy - y;
        double bx = ca * dx + (-sa) * dy, by = sa * dx + ca * dy;
        double ang = ssa(atan2(by, bx) - alpha);
        obs[25] = sin(ang); obs[26] = cos(ang);
    }
}

/* ------------------------------------------------------------------------- */
/* step = thrust + physics + observation + reward + termination (env:394-615) */
/* ------------------------------------------------------------------------- */

void d2o_env_step(d2o_state *s, const d2o_path *p, const d2o_obstacles *ob,
                  const d2o_config *c, double action_left, double action_right,
                  d2o_step_out *o)
{
    d2o_space_step(s, ob, c, action_left, action_right);
    s->steps += 1;                                          /* env:407 */
    d2o_after_physics(s, p, ob, c, o);
}

/* env:421-615: everything step() does after space.step(): observation, reward terms,
 * termination flags and the episode accumulators, for the pose/velocity/collision flag/step
 * count already in *s.  Split out so that tests can replay genuine reference trajectories
 * (Tests/<scenario>/flight_paths) through the reward without the restated physics. */
void d2o_after_physics(d2o_state *s, const d2o_path *p, const d2o_obstacles *ob,
                       const d2o_config *c, d2o_step_out *o)
{
    const double W = c->screen_x, H = c->screen_y;
    const double TWO_PI = 2 * M_PI;
    memset(o, 0, sizeof *o);
    d2o_observe(s, p, ob, c, o);
    const double *obs = o->obs;

    /* env:423-445: decode the (float64) observation back */
    double vel_x = invm1to1(obs[0], -1330, 1330), vel_y = invm1to1(obs[1], -1330, 1330);
    double alpha = obs[3] * M_PI;
    double tdx = invm1to1(obs[4], 0, W), tdy = invm1to1(obs[5], 0, H);
    double pos_x = invm1to1(obs[6], 0, W), pos_y = invm1to1(obs[7], 0, H);
    double vel_angle = py_mod(atan2(obs[17] * M_PI, obs[18] * M_PI) + TWO_PI, TWO_PI);
    double cpx = invm1to1(obs[19], 0, W), cpy = invm1to1(obs[20], 0, H);
    double la_angle = py_mod(atan2(obs[23], obs[24]) + TWO_PI, TWO_PI);

    double lambda_PA = 1, lambda_CA = 1;
    double r_CA = 0;
    double d1 = INFINITY;
    if (ob->n > 0) {                                         /* env:462-523 */
        d1 = invm1to1(obs[8], 0, sqrt(W * W + H * H));
        double obs_angle = py_mod(atan2(obs[9], obs[10]) + TWO_PI, TWO_PI);
        double angle_diff = fabs((py_mod(obs_angle - vel_angle + M_PI, TWO_PI) - M_PI) * (180.0 / M_PI));
        double DR = c->danger_range, DA = c->danger_angle, kk = c->abs_inv_CA_min_rew;
        if (d1 < DR && c->use_Lambda) {
            lambda_PA = (d1 / DR) / 2;
            if (lambda_PA < 0.10) lambda_PA = 0.10;
            lambda_CA = 1 - lambda_PA;
        }
        if (d1 < DR) {
            double range_rew = -(((DR + kk * DR) / (d1 + kk * DR)) - 1);
            double angle_rew = -(((DA + kk * DA) / (angle_diff + kk * DA)) - 1);
            if (angle_rew > 0) angle_rew = 0;
            if (range_rew > 0) range_rew = 0;
            r_CA = range_rew + angle_rew;
        }
    }

    /* env:527-530 */
    double ddx = cpx - pos_x, ddy = cpy - pos_y;
    double dist_from_path = norm2(ddx, ddy);                 /* np.linalg.norm, env:529 */
    double r_PA = -(2 * (clipd(dist_from_path, 0, c->PA_band_edge) / c->PA_band_edge) - 1) * c->PA_scale;

    /* env:534-539 */
    double velocity = sqrt(vel_x * vel_x + vel_y * vel_y);
    double scaled_velocity = velocity * c->PP_vel_scale;
    double vel_LA_diff = fabs(py_mod(la_angle - vel_angle + M_PI, TWO_PI) - M_PI);
    double r_PP = clipd(cos(vel_LA_diff) * scaled_velocity, c->PP_rew_min, c->PP_rew_max);

    /* env:543-570 */
    double r_coll = 0; int end1 = 0;
    if (s->collided) { r_coll = c->rew_collision; end1 = 1; }
    double r_end = 0; int end2 = 0;
    if (fabs(tdx) < c->reach_end_radius && fabs(tdy) < c->reach_end_radius) { end2 = 1; r_end = c->rew_reach_end; }
    double r_AA = 0; int end5 = 0;
    if (alpha > c->AA_band) r_AA = -sin(alpha);
    if (alpha < -c->AA_band) r_AA = sin(alpha);
    if (fabs(alpha) >= c->AA_angle) { r_AA = c->rew_AA; end5 = 1; }
    int end4 = (s->steps == c->n_steps);

    /* env:572 */
    double reward = r_AA + r_PA * lambda_PA + r_PP + r_coll + r_CA * lambda_CA + r_end;

    o->reward = reward;
    o->r_CA = r_CA * lambda_CA; o->r_PA = r_PA * lambda_PA; o->r_PP = r_PP;
    o->r_coll = r_coll; o->r_end = r_end; o->r_AA = r_AA;
    o->dist_closest_obs = d1;
    s->path_error += dist_from_path;                          /* env:589-592 */
    o->ape = s->path_error / s->steps;
    s->total_reward += reward;
    o->end_collision = end1; o->end_reached = end2; o->end_time = end4; o->end_alpha = end5;
    o->done = end1 || end2 || end4 || end5;
}

/* Frame velocity after each of n consecutive space steps from *s0 (which is not modified) under the
 * given actions [n][2]: the forward model of the inverse-dynamics discriminator (oracle/inverse_dynamics.py).
 * With jac != NULL also the forward-difference Jacobian d v[2n] / d actions[2n] (row-major [2n][2n], step h). */
void d2o_frame_velocity_rollout(const d2o_state *s0, const double *actions, int n, const d2o_config *c,
                                double *v_out, double *jac, double h)
{
    d2o_obstacles none; none.n = 0;
    d2o_state s = *s0;
    for (int k = 0; k < n; k++) {
        d2o_space_step(&s, &none, c, actions[2 * k], actions[2 * k + 1]);
        v_out[2 * k] = s.body[0].vx; v_out[2 * k + 1] = s.body[0].vy;
    }
    if (!jac) return;
    const int m = 2 * n;
    for (int i = 0; i < m * m; i++) jac[i] = 0.0;
    /* a perturbation of action k only changes steps >= k: restart from the state before step k */
    d2o_state base = *s0;
    for (int k = 0; k < n; k++) {
        for (int q = 0; q < 2; q++) {
            d2o_state t = base;
            for (int j = k; j < n; j++) {
                double aL = actions[2 * j], aR = actions[2 * j + 1];
                if (j == k) { if (q == 0) aL += h; else aR += h; }
                d2o_space_step(&t, &none, c, aL, aR);
                jac[(2 * j) * m + 2 * k + q] = (t.body[0].vx - v_out[2 * j]) / h;
                jac[(2 * j + 1) * m + 2 * k + q] = (t.body[0].vy - v_out[2 * j + 1]) / h;
            }
        }
        d2o_space_step(&base, &none, c, actions[2 * k], actions[2 * k + 1]);
    }
}

void d2o_env_step_batch(int n, d2o_state *s, const d2o_path *p, int path_stride,
                        const d2o_obstacles *ob, int ob_stride, const d2o_config *c,
                        const float *actions, d2o_step_out *o)
{
    for (int i = 0; i < n; i++)
        d2o_env_step(&s[i], &p[(long)i * path_stride], &ob[(long)i * ob_stride], c,
                     (double)actions[2 * i], (double)actions[2 * i + 1], &o[i]);
}

/* ------------------------------------------------------------------------- */
/* Philox4x32-10                                                              */
/* ------------------------------------------------------------------------- */
void d2o_philox4x32_10(const unsigned ctr[4], const unsigned key[2], unsigned out[4])
{
    unsigned c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
    unsigned k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; r++) {
        unsigned long long p0 = (unsigned long long)0xD2511F53u * c0;
        unsigned long long p1 = (unsigned long long)0xCD9E8D57u * c2;
        unsigned n0 = (unsigned)(p1 >> 32) ^ c1 ^ k0;
        unsigned n1 = (unsigned)p1;
        unsigned n2 = (unsigned)(p0 >> 32) ^ c3 ^ k1;
        unsigned n3 = (unsigned)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
