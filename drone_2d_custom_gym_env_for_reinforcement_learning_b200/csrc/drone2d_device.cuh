// drone2d_device.cuh -- device functions of the fused Drone2dEnv step (sm_100a).
//
// One lane = one environment; all arithmetic is float64 (the reference is
// float64 end to end and a float32 solver drifts 1e-1 px in 200 steps).
// File:line citations are relative to the reference's drone_2d_custom_gym_env/
// ("env" = drone_2d_env.py); Chipmunk/scipy behaviour is restated in
// SURVEY.md Appendix A.
//
// Two arithmetic regimes:
//   * "strict" helpers (S::mul/add/...) map to the round-to-nearest intrinsics,
//     which nvcc never contracts into FMAs.  The closest-point search (Brent on
//     a flat-bottomed objective) and the nearest-obstacle ranking are
//     discontinuous functions of their inputs, so they are evaluated with the
//     reference's exact operation order: identical inputs give bit-identical
//     u*, evaluation counts and obstacle indices.
//   * everything else (rigid-body solver, normalisation, reward) is ordinary
//     double arithmetic with FMA contraction; it differs from the reference by
//     rounding only (parity tolerance in tests/).
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>

#include "drone2d_layout.h"

#define D2D_US_STRIDE D2D_BLOCK   // lane-private shared-memory columns: bank-conflict free for any per-lane index
#define D2D_PI 3.14159265358979323846
#define D2D_TWO_PI 6.28318530717958647692

namespace d2d {

struct S {  // strict IEEE double ops, never contracted
    static __device__ __forceinline__ double mul(double a, double b) { return __dmul_rn(a, b); }
    static __device__ __forceinline__ double add(double a, double b) { return __dadd_rn(a, b); }
    static __device__ __forceinline__ double sub(double a, double b) { return __dsub_rn(a, b); }
    static __device__ __forceinline__ double div(double a, double b) { return __ddiv_rn(a, b); }
    static __device__ __forceinline__ double fma(double a, double b, double c) { return __fma_rn(a, b, c); }
    static __device__ __forceinline__ double sqrt(double a) { return __dsqrt_rn(a); }
};

// ---------------------------------------------------------------------------
// scalar helpers
// ---------------------------------------------------------------------------

// Exact fmod(a, b) for b > 0 and |a/b| < 2^20: the remainder of an IEEE fmod is
// always representable, so one FMA recovers it once the quotient is right.
__device__ __forceinline__ double fmod_small(double a, double b)
{
    double q = trunc(a / b);
    double r = __fma_rn(-q, b, a);
    // a/b may have rounded across an integer: repair by one period (exact).
    if (a >= 0.0) {
        if (r < 0.0) r += b;
        else if (r >= b) r -= b;
    } else {
        if (r > 0.0) r -= b;
        else if (r <= -b) r += b;
    }
    return r;
}

// Python / numpy float modulo for a positive divisor (result in [0, b)).
__device__ __forceinline__ double py_mod_pos(double a, double b)
{
    double m = fmod_small(a, b);
    if (m < 0.0) m += b;
    return m;   // m == -0.0 cannot survive: (-0.0 < 0) is false, numpy returns +0.0 via copysign; harmless
}

// transformations.py:6-7
__device__ __forceinline__ double ssa(double a) { return py_mod_pos(a + D2D_PI, D2D_TWO_PI) - D2D_PI; }

// x / d for a constant divisor d with rd = RN(1/d) precomputed on the host: Markstein's sequence
// q = RN(x*rd), rem = x - d*q (exact, FMA), RN(q + rem*rd) = RN(x/d) -- the IEEE quotient the
// reference computes, in 3 FMA-class instructions instead of the ~14 of a double division.
__device__ __forceinline__ double div_const(double x, double d, double rd)
{
    const double q = x * rd;
    const double rem = __fma_rn(-d, q, x);
    return __fma_rn(rem, rd, q);
}

// env:972-978; m1to1 with hi - lo = d and rd = 1/d
__device__ __forceinline__ double m1to1(double v, double lo, double d, double rd) { return div_const(2.0 * (v - lo), d, rd) - 1.0; }
__device__ __forceinline__ double invm1to1(double v, double lo, double hi) { return (v + 1.0) * (hi - lo) / 2.0 + lo; }
__device__ __forceinline__ double clipd(double v, double lo, double hi) { return v < lo ? lo : (v > hi ? hi : v); }
// min / max of finite values (fmin/fmax carry NaN handling that costs ~8 instructions per call)
// |x|, -x and x with a given sign as integer operations on the high word: ptxas materialises
// fabs()/negation of a double as DADD (an FP64-pipe instruction with FP64 latency) whenever the
// result feeds a select; these are single ALU operations.  Bit-identical for every input.
__device__ __forceinline__ double abs_b(double x) { return __hiloint2double(__double2hiint(x) & 0x7fffffff, __double2loint(x)); }
__device__ __forceinline__ double neg_b(double x) { return __hiloint2double(__double2hiint(x) ^ 0x80000000, __double2loint(x)); }
__device__ __forceinline__ double neg_if(double x, bool c) { return __hiloint2double(__double2hiint(x) ^ (c ? 0x80000000 : 0), __double2loint(x)); }

__device__ __forceinline__ double min_f(double a, double b) { return (b < a) ? b : a; }
__device__ __forceinline__ double max_f(double a, double b) { return (b > a) ? b : a; }

// Out-of-line transcendentals: CUDA's double sincos/atan2 expand to ~150 instructions each at
// every call site; the step calls them ~20 times, which made the kernel's hot path several times
// larger than the instruction cache.  One shared copy each (results returned in registers).
__device__ __noinline__ double2 sincos2(double a)
{
    double sn, cs;
    sincos(a, &sn, &cs);
    return make_double2(sn, cs);
}

__device__ __noinline__ double atan2_ool(double y, double x) { return atan2(y, x); }

// sin / cos of a body angle, inline (the rigid-body step needs four per step, at a point where ~60 registers are live:
// every out-of-line call there saves and restores them).  One Cody-Waite reduction by pi/2 and the fdlibm kernels on
// |r| <= pi/4 (error < 1 ulp: rounding-level difference from the reference's libm calls); body angles stay within a
// few radians (|alpha| >= pi/2 ends the episode), anything larger takes the library route.
__device__ __forceinline__ void sincos_body(double a, double &sn, double &cs)
{
    if (fabs(a) > 6.0) { const double2 sc = sincos2(a); sn = sc.x; cs = sc.y; return; }
    const double q = rint(a * 0.63661977236758134308);                        // 2 / pi
    double r = fma(-q, 1.57079632679489655800e+00, a);
    r = fma(-q, 6.12323399573676603587e-17, r);
    const double z = r * r;
    double ps = 1.58969099521155010221e-10;
    ps = fma(ps, z, -2.50507602534068634195e-08); ps = fma(ps, z, 2.75573137070700676789e-06);
    ps = fma(ps, z, -1.98412698298579493134e-04); ps = fma(ps, z, 8.33333333332248946124e-03);
    ps = fma(ps, z, -1.66666666666666324348e-01);
    const double sr = fma(r * z, ps, r);
    double pc = -1.13596475577881948265e-11;
    pc = fma(pc, z, 2.08757232129817482790e-09); pc = fma(pc, z, -2.75573143513906633035e-07);
    pc = fma(pc, z, 2.48015872894767294178e-05); pc = fma(pc, z, -1.38888888888741095749e-03);
    pc = fma(pc, z, 4.16666666666666019037e-02);
    const double cr = fma(z * z, pc, fma(-0.5, z, 1.0));
    const int k = (int)q & 3;
    const double s0 = (k & 1) ? cr : sr, c0 = (k & 1) ? sr : cr;
    sn = (k & 2) ? -s0 : s0;
    cs = ((k + 1) & 2) ? -c0 : c0;
}

// (sin, cos) of ssa((atan2(dy, dx) - alpha) - off) is the bearing pattern of env:695-698, 724-727, 752-763.
// Evaluated without transcendentals: (sin, cos) of atan2(dy, dx) - beta is the unit vector (dx, dy)/r rotated by
// -beta, with (cos beta, sin beta) = sign * (cb, sb) already known (the frame's rotation; sign = -1 folds the "- pi" of
// env:695).  Differs from the atan2 -> ssa -> sincos route of the reference by rounding only (the wrap of ssa does not
// change sin / cos); ~35 instructions instead of ~330.  atan2(0, 0) = 0, i.e. the unit vector (1, 0), for a null vector.
__device__ __forceinline__ double2 bearing_rot(double dy, double dx, double cb, double sb, double sign)
{
    const double r2 = dx * dx + dy * dy;
    double ux = 1.0, uy = 0.0;
    if (r2 > 0.0) {
        const double inv = 1.0 / sqrt(r2);
        ux = dx * inv; uy = dy * inv;
    }
    return make_double2(sign * (uy * cb - ux * sb), sign * (ux * cb + uy * sb));
}

// ---------------------------------------------------------------------------
// Philox4x32-10 counter-based stream (reset sampling, random actions)
// ---------------------------------------------------------------------------
__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                              uint32_t k0, uint32_t k1, uint32_t out[4])
{
#pragma unroll
    for (int r = 0; r < 10; r++) {
        uint32_t h0 = __umulhi(0xD2511F53u, c0), l0 = 0xD2511F53u * c0;
        uint32_t h1 = __umulhi(0xCD9E8D57u, c2), l1 = 0xCD9E8D57u * c2;
        uint32_t n0 = h1 ^ c1 ^ k0, n2 = h0 ^ c3 ^ k1;
        c0 = n0; c1 = l1; c2 = n2; c3 = l0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

#define D2D_STREAM_RESET 0x52534554u   // 'RSET'
#define D2D_STREAM_ACTION 0x4143544Eu  // 'ACTN'

__device__ __forceinline__ double u01_from_bits(uint32_t hi, uint32_t lo)
{
    // 53-bit uniform in [0,1) (numpy's random_double construction)
    return ((double)(hi >> 5) * 67108864.0 + (double)(lo >> 6)) * (1.0 / 9007199254740992.0);
}

// variates from uniforms: ONE definition each, used by the sequential stream and by the random-access (warp-
// cooperative) sampler, so that both round identically
__device__ __forceinline__ double uniform_from(double u, double lo, double hi) { return lo + (hi - lo) * u; }
__device__ __forceinline__ double normal_from(double u1, double u2, double mu, double sd)
{
    // Box-Muller (cosine branch)
    const double z = sqrt(-2.0 * log(1.0 - u1)) * cos(D2D_TWO_PI * u2);
    return mu + sd * z;
}

// Sequential draws d = 0,1,2,... of one (env, episode): block d/2 of the stream.
struct ResetRng {
    uint32_t env, episode, k0, k1, blk;
    double spare;
    bool has;
    __device__ ResetRng(uint32_t env_, uint32_t ep_, uint32_t k0_, uint32_t k1_)
        : env(env_), episode(ep_), k0(k0_), k1(k1_), blk(0), spare(0.0), has(false) {}
    __device__ double u01()
    {
        if (has) { has = false; return spare; }
        uint32_t o[4];
        philox4x32_10(env, episode, blk, D2D_STREAM_RESET, k0, k1, o);
        blk++;
        spare = u01_from_bits(o[2], o[3]);
        has = true;
        return u01_from_bits(o[0], o[1]);
    }
    // index of the next draw, and repositioning (draw d lives in block d/2: first half if d is even)
    __device__ uint32_t position() const { return 2u * blk - (has ? 1u : 0u); }
    __device__ void seek(uint32_t d)
    {
        blk = d >> 1; has = false;
        if (d & 1u) (void)u01();          // consumes the block's first half, leaves the second as the spare
    }
    // draw d without touching the sequential state
    __device__ double u01_at(uint32_t d) const
    {
        uint32_t o[4];
        philox4x32_10(env, episode, d >> 1, D2D_STREAM_RESET, k0, k1, o);
        return (d & 1u) ? u01_from_bits(o[2], o[3]) : u01_from_bits(o[0], o[1]);
    }
    __device__ double uniform(double lo, double hi) { return uniform_from(u01(), lo, hi); }
    // consumes two draws
    __device__ double normal(double mu, double sd)
    {
        const double u1 = u01(), u2 = u01();
        return normal_from(u1, u2, mu, sd);
    }
    __device__ int randint(int a, int b)   // inclusive, like random.randint
    {
        int v = a + (int)floor(u01() * (double)(b - a + 1));
        return v > b ? b : v;
    }
};

// ---------------------------------------------------------------------------
// table views: shared scenario (smem) or per-env rows (global, row-major by index)
// ---------------------------------------------------------------------------
struct SharedView {
    const D2dSharedTables *t;
    __device__ __forceinline__ int n_wps() const { return t->n_wps; }
    __device__ __forceinline__ int n_obs() const { return t->n_obs; }
    __device__ __forceinline__ double us(int k) const { return t->us[k]; }
    __device__ __forceinline__ double2 wp(int k) const { return make_double2(t->wx[k], t->wy[k]); }
    __device__ __forceinline__ double2 pa(int r) const { return make_double2(t->ax[r], t->ay[r]); }
    __device__ __forceinline__ double2 pb(int r) const { return make_double2(t->bx[r], t->by[r]); }
    __device__ __forceinline__ double2 pc(int r) const { return make_double2(t->cx[r], t->cy[r]); }
    __device__ __forceinline__ void obstacle(int k, double &x, double &y, double &r) const
    {
        x = t->ox[k]; y = t->oy[k]; r = t->orad[k];
    }
};

// L2 eviction hints.  The per-step working set (~85 MB at 65 536 envs) cycles through the L2 and is
// re-read from DRAM every step.  Level 1: path rows / us[] -- the loads a search waits on
// (PathCursor::locate) -- are evict_last, obstacle rows (read once per step, front to back) evict_first.
// Level 2 (default): the dynamic state is streamed as well (evict_first loads and stores).
// Measured at 65 536 envs (stage_5): level 0 / 1 / 2 = 75.2 / 75.2 / 74.65 us per step.
#ifndef D2D_L2_HINTS
#define D2D_L2_HINTS 2
#endif
__device__ __forceinline__ uint64_t l2_keep() { uint64_t q; asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(q)); return q; }
__device__ __forceinline__ uint64_t l2_stream() { uint64_t q; asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(q)); return q; }
__device__ __forceinline__ double2 ld_keep(const double2 *a)
{
#if D2D_L2_HINTS
    double2 r; asm volatile("ld.global.L2::cache_hint.v2.f64 {%0,%1}, [%2], %3;" : "=d"(r.x), "=d"(r.y) : "l"(a), "l"(l2_keep())); return r;
#else
    return *a;
#endif
}
__device__ __forceinline__ double2 ld_stream(const double2 *a)
{
#if D2D_L2_HINTS
    double2 r; asm volatile("ld.global.L1::no_allocate.L2::cache_hint.v2.f64 {%0,%1}, [%2], %3;" : "=d"(r.x), "=d"(r.y) : "l"(a), "l"(l2_stream())); return r;
#else
    return *a;
#endif
}
__device__ __forceinline__ double ld_stream(const double *a)
{
#if D2D_L2_HINTS
    double r; asm volatile("ld.global.L1::no_allocate.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(r) : "l"(a), "l"(l2_stream())); return r;
#else
    return *a;
#endif
}
__device__ __forceinline__ void st_stream(double2 *a, double2 v)
{
#if D2D_L2_HINTS >= 2
    asm volatile("st.global.L2::cache_hint.v2.f64 [%0], {%1,%2}, %3;" ::"l"(a), "d"(v.x), "d"(v.y), "l"(l2_stream()) : "memory");
#else
    *a = v;
#endif
}
__device__ __forceinline__ double2 ld_state(const double2 *a)
{
#if D2D_L2_HINTS >= 2
    return ld_stream(a);
#else
    return *a;
#endif
}

struct EnvView {
    const D2dParams *p;
    int64_t i;
    int nobs;
    int par;                // table buffer (0/1) this view reads
    const double *us_s;     // optional shared-memory copy of this lane's us[] (element k at us_s[k * D2D_US_STRIDE])
    // 32-bit element offsets: 2 * 32 rows * np < 2^32 for np < 6.7e7 environments per GPU
    __device__ __forceinline__ uint32_t row8(int k) const { return (par ? p->stg_off8 : 0u) + (uint32_t)k * (uint32_t)p->np + (uint32_t)i; }
    __device__ __forceinline__ uint32_t row16(int k) const { return (par ? p->stg_off16 : 0u) + (uint32_t)k * (uint32_t)p->np + (uint32_t)i; }
    __device__ __forceinline__ int n_wps() const { return p->c.n_wps; }
    __device__ __forceinline__ int n_obs() const { return nobs; }
    __device__ __forceinline__ double us(int k) const { return us_s ? us_s[k * D2D_US_STRIDE] : p->us[row8(k)]; }
    __device__ __forceinline__ double2 wp(int k) const { return p->wps[row16(k)]; }
    __device__ __forceinline__ double2 pa(int r) const { return ld_keep(&p->pa[row16(r)]); }
    __device__ __forceinline__ double2 pb(int r) const { return ld_keep(&p->pb[row16(r)]); }
    __device__ __forceinline__ double2 pc(int r) const { return ld_keep(&p->pc[row16(r)]); }
    __device__ __forceinline__ void obstacle(int k, double &x, double &y, double &r) const
    {
        const double2 xy = ld_stream(&p->obxy[row16(k)]);
        x = xy.x; y = xy.y; r = ld_stream(&p->obr[row8(k)]);
    }
};

// ---------------------------------------------------------------------------
// QPMI2D (predef_path.py:53-188), strict arithmetic
// ---------------------------------------------------------------------------
__device__ __forceinline__ double poly2(double a, double b, double c, double u, double uu)
{
    return S::add(S::add(S::mul(a, uu), S::mul(b, u)), c);   // a*u**2 + b*u + c
}

template <class V>
__device__ __forceinline__ int u_index(const V &v, double u, int N)
{
    // predef_path.py:53-63: first n with u <= us[n+1]; us ascending -> count of us[k] < u
    int n = 0;
    for (int k = 1; k < N; k++) n += (u > v.us(k)) ? 1 : 0;
    return n;
}

// predef_path.py:88-142
template <class V>
__device__ __forceinline__ void path_eval(const V &v, double u, double &x, double &y)
{
    const int N = v.n_wps(), rows = N - 2;
    const double uu = S::mul(u, u);
    int r;
    if (u >= v.us(0) && u <= v.us(1)) {
        r = 0;
    } else if (u >= S::sub(v.us(N - 2), 0.001)) {
        // (us[-2]-0.001 <= u <= us[-1]) or get_u_index(u) == len(us)-1 (i.e. u > us[-1])
        r = rows - 1;
    } else {
        const int n = u_index(v, u, N);
        const double un = v.us(n), un1 = v.us(n + 1);
        const double len = S::sub(un1, un);
        const double mu_r = S::div(S::sub(u, un), len);
        const double mu_f = S::div(S::sub(un1, u), len);
        const int r1 = (n - 1 < 0) ? rows - 1 : n - 1;       // Python negative index for u < 0
        const double2 a1 = v.pa(r1), b1 = v.pb(r1), c1 = v.pc(r1);
        const double2 a2 = v.pa(n), b2 = v.pb(n), c2 = v.pc(n);
        const double x1 = poly2(a1.x, b1.x, c1.x, u, uu), y1 = poly2(a1.y, b1.y, c1.y, u, uu);
        const double x2 = poly2(a2.x, b2.x, c2.x, u, uu), y2 = poly2(a2.y, b2.y, c2.y, u, uu);
        x = S::add(S::mul(mu_r, x2), S::mul(mu_f, x1));
        y = S::add(S::mul(mu_r, y2), S::mul(mu_f, y1));
        return;
    }
    const double2 a = v.pa(r), b = v.pb(r), c = v.pc(r);
    x = poly2(a.x, b.x, c.x, u, uu);
    y = poly2(a.y, b.y, c.y, u, uu);
}

// predef_path.py:145-188 (only used by the obstacle sampler of the reset)
template <class V>
__device__ __forceinline__ void path_gradient(const V &v, double u, double &dx, double &dy)
{
    const int N = v.n_wps(), rows = N - 2;
    int r;
    if (u >= v.us(0) && u <= v.us(1)) {
        r = 0;
    } else if (u >= v.us(N - 2)) {
        r = rows - 1;
    } else {
        const int n = u_index(v, u, N);
        const double un = v.us(n), un1 = v.us(n + 1);
        const double len = S::sub(un1, un);
        const double mu_r = S::div(S::sub(u, un), len);
        const double mu_f = S::div(S::sub(un1, u), len);
        const int r1 = (n - 1 < 0) ? rows - 1 : n - 1;
        const double2 a1 = v.pa(r1), b1 = v.pb(r1), a2 = v.pa(n), b2 = v.pb(n);
        const double dx1 = S::add(S::mul(S::mul(a1.x, u), 2.0), b1.x), dy1 = S::add(S::mul(S::mul(a1.y, u), 2.0), b1.y);
        const double dx2 = S::add(S::mul(S::mul(a2.x, u), 2.0), b2.x), dy2 = S::add(S::mul(S::mul(a2.y, u), 2.0), b2.y);
        dx = S::add(S::mul(mu_r, dx2), S::mul(mu_f, dx1));
        dy = S::add(S::mul(mu_r, dy2), S::mul(mu_f, dy1));
        return;
    }
    const double2 a = v.pa(r), b = v.pb(r);
    dx = S::add(S::mul(S::mul(a.x, u), 2.0), b.x);
    dy = S::add(S::mul(S::mul(a.y, u), 2.0), b.y);
}

// ---------------------------------------------------------------------------
// PathCursor: path_eval with the data of the last visited path REGION cached (two polynomial
// rows in a lane-private shared-memory column, the rest in registers), so that most evaluations
// of the closest-point search need no table access and no segment search.  The arithmetic
// performed for a given u is IDENTICAL to path_eval (same strict operations, same order) --
// only the loads are cached.  Regions (n_wps = N >= 4, us increasing):
//   START     u <= us[1]            rows (last, 0): u < us[0] blends them (Python index -1 quirk),
//                                   us[0] <= u <= us[1] uses row 0 alone (first stretch)
//   INTERIOR  us[n] < u <= us[n+1]  1 <= n <= N-4, blend of rows (n-1, n)
//   END       u > us[N-3]           rows (N-4, N-3): blend below us[N-2] - 0.001, the last row alone
//                                   from there on (last stretch and beyond the end)
// START and END each merge a blended and a single-polynomial piece that share a row: the search
// straddles exactly these junctions (the discontinuity at u = 0, the end of the path) whenever
// the closest point is an end point, which is the slow, 45-evaluation case -- with separate
// pieces every crossing was a cache miss, and in a warp some lane crossed on almost every trip.
// ---------------------------------------------------------------------------
template <class V>
struct PathCursor {
    const V &v;
    int N;
    double vlo_x, vhi_i;         // region validity: vlo_x < u <= vhi_i
    double s_lo, s_hi;           // inside the region, s_lo <= u <= s_hi evaluates row 2 alone
    double lo, un1, len, rlen;   // blending segment: us[n], us[n+1], their difference and its reciprocal
    // the two polynomial rows live in a lane-private shared-memory column
    // (coef[r * D2D_BLOCK], r = 0..5 = a1, b1, c1, a2, b2, c2): 24 registers fewer during the
    // search, 6 conflict-free LDS.128 per evaluation instead.
    // The column is addressed through a 32-bit shared-window address that is passed through a
    // (no-op) warp shuffle: ptxas otherwise rematerialises &s_coef[tid] from S2R on every
    // evaluation (S2R + S2UR + 4 integer ops on the critical path of each search iteration).
    uint32_t caddr;
    double inv_seg;              // (N-1)/L: initial guess of the segment index (corrected exactly)

    template <int R>
    __device__ __forceinline__ double2 ld() const
    {
        double2 r;
        asm volatile("ld.shared.v2.f64 {%0, %1}, [%2+%3];" : "=d"(r.x), "=d"(r.y) : "r"(caddr), "n"(R * D2D_BLOCK * 16));
        return r;
    }
    template <int R>
    __device__ __forceinline__ void st(double2 val)
    {
        asm volatile("st.shared.v2.f64 [%0+%1], {%2, %3};" ::"r"(caddr), "n"(R * D2D_BLOCK * 16), "d"(val.x), "d"(val.y) : "memory");
    }

    __device__ __forceinline__ PathCursor(const V &v_, double2 *coef_) : v(v_), N(v_.n_wps())
    {
        const uint32_t a = (uint32_t)__cvta_generic_to_shared(coef_);
        caddr = __shfl_sync(__activemask(), a, threadIdx.x & 31);
        vlo_x = CUDART_INF; vhi_i = -CUDART_INF;                       // empty: the first use locates
        s_lo = CUDART_INF; s_hi = -CUDART_INF;
        lo = un1 = len = rlen = 0.0;
        inv_seg = (double)(N - 1) / v.us(N - 1);
    }

    __device__ __forceinline__ bool hit(double u) const { return (u > vlo_x) && (u <= vhi_i); }

    __device__ __forceinline__ void load_rows(int r1, int r2)
    {
        st<0>(v.pa(r1)); st<1>(v.pb(r1)); st<2>(v.pc(r1));
        st<3>(v.pa(r2)); st<4>(v.pb(r2)); st<5>(v.pc(r2));
    }

    __device__ __forceinline__ void locate(double u)
    {
        const int rows = N - 2;
        const double us1 = v.us(1), usm3 = v.us(N - 3);
        if (u <= us1) {                                              // START
            lo = v.us(0); un1 = us1;
            vlo_x = -CUDART_INF; vhi_i = us1;
            s_lo = lo; s_hi = us1;
            load_rows(rows - 1, 0);
        } else if (u > usm3) {                                       // END
            lo = usm3; un1 = v.us(N - 2);
            vlo_x = usm3; vhi_i = CUDART_INF;
            s_lo = S::sub(un1, 0.001); s_hi = CUDART_INF;
            load_rows(rows - 2, rows - 1);
        } else {                                                     // INTERIOR
            // get_u_index (predef_path.py:53-63) = number of us[k], k >= 1, below u: start from the
            // uniform-spacing guess and walk to the exact answer (0-1 steps on the reference's paths)
            int n = min(max((int)(u * inv_seg), 1), N - 4);
            while (n < N - 4 && u > v.us(n + 1)) n++;
            while (n > 1 && !(u > v.us(n))) n--;
            lo = v.us(n); un1 = v.us(n + 1);
            vlo_x = lo; vhi_i = un1;
            s_lo = CUDART_INF; s_hi = -CUDART_INF;
            load_rows(n - 1, n);
        }
        len = S::sub(un1, lo);
        rlen = S::div(1.0, len);
    }

    // x / len, correctly rounded, from the cached correctly rounded reciprocal (Markstein:
    // q = RN(x*r), rem = x - len*q exactly (FMA), RN(q + rem*r) = RN(x/len)); bit-identical to
    // the IEEE division the reference performs, at 3 FMA-class instructions instead of ~12.
    __device__ __forceinline__ double div_len(double x) const
    {
        const double q = S::mul(x, rlen);
        const double rem = S::fma(-len, q, x);
        return S::fma(rem, rlen, q);
    }

    __device__ __forceinline__ void eval(double u, double &x, double &y)
    {
        if (!hit(u)) locate(u);
        const double uu = S::mul(u, u);
        const double2 a2 = ld<3>(), b2 = ld<4>(), c2 = ld<5>();
        const double x2 = poly2(a2.x, b2.x, c2.x, u, uu), y2 = poly2(a2.y, b2.y, c2.y, u, uu);
        if (u >= s_lo && u <= s_hi) { x = x2; y = y2; return; }
        const double2 a1 = ld<0>(), b1 = ld<1>(), c1 = ld<2>();
        const double mu_r = div_len(S::sub(u, lo));
        const double mu_f = div_len(S::sub(un1, u));
        const double x1 = poly2(a1.x, b1.x, c1.x, u, uu), y1 = poly2(a1.y, b1.y, c1.y, u, uu);
        x = S::add(S::mul(mu_r, x2), S::mul(mu_f, x1));
        y = S::add(S::mul(mu_r, y2), S::mul(mu_f, y1));
    }
};

// QPMI2D.get_closest_u (predef_path.py:226-248) = scipy fminbound = bounded Brent
// (_minimize_scalar_bounded, scipy/optimize/_optimize.py:2289-2440) on
// [-10, L+10], xatol = 1e-6, maxfun = 500.  Same floating-point operations in the same
// order as the reference, in scipy's own loop shape (first evaluation in front of the loop, update
// rules at its bottom); get_closest_position (predef_path.py:251-255) is one more evaluation at
// the minimiser after the loop (4 selects per trip cheaper than carrying the best point through it).
// Every lane runs its own iteration count (typically 10-47).
// The OPENING of that search does not depend on the query point: the first evaluation is at
// x0 = a + g (b - a), the second (e = 0: a golden-section step) at x1 = x0 + g (b - x0), and the third is a
// golden-section step again -- with three evaluations in hand two of scipy's three points coincide, p = q = 0 and
// the parabolic step is rejected -- from x1 into (x0, b) if the second value was not worse, from x0 into (a, x1) if it
// was.  Four abscissae per PATH, computed here with the search's own operations; their path points are stored with
// the episode's tables (D2dParams::hint), so the three opening evaluations of every search of the episode -- the ones
// that jump across the whole path and each miss the cursor's cached region (a DRAM round trip for the warp) -- need
// no table access at all.
__device__ __forceinline__ void brent_opening(double L, double u[D2D_HINTS])
{
    const double xatol_3 = 1e-6 / 3.0, sqrt_eps = 1.4832396974191326e-08, golden_mean = 0.3819660112501051;
    const double a = -10.0, b = S::add(L, 10.0);
    auto golden_step = [&](double lo, double hi, double xf) {
        const double xm = S::mul(0.5, S::add(lo, hi));
        const double tol1 = S::add(S::mul(sqrt_eps, fabs(xf)), xatol_3);
        const double e = (xf >= xm) ? S::sub(lo, xf) : S::sub(hi, xf);
        const double rat = S::mul(golden_mean, e);
        const double mag = max_f(fabs(rat), tol1);
        return S::add(xf, (rat < 0.0) ? -mag : mag);
    };
    u[0] = S::add(a, S::mul(golden_mean, S::sub(b, a)));
    u[1] = golden_step(a, b, u[0]);
    u[2] = golden_step(u[0], b, u[1]);      // f(x1) <= f(x0): bracket (x0, b), best point x1
    u[3] = golden_step(a, u[1], u[0]);      // f(x1) >  f(x0): bracket (a, x1), best point x0
}

// Path points of the opening, written next to the tables they were evaluated on (reset / staging / d2d_set_tables).
template <class V>
__device__ __forceinline__ void store_hints(const D2dParams &p, const V &v)
{
    double u[D2D_HINTS];
    brent_opening(v.us(v.n_wps() - 1), u);
#pragma unroll
    for (int k = 0; k < D2D_HINTS; k++) {
        double x, y;
        path_eval(v, u[k], x, y);
        p.hint[v.row16(k)] = make_double2(x, y);
    }
}

// `hinted`: the cursor's column holds the D2D_HINTS opening points in its rows 0..3 (they are consumed before the
// cursor's first locate() overwrites the column).
template <class V>
__device__ __forceinline__ double closest_u(PathCursor<V> &pc, double px, double py, int &nfev, double &bx, double &by,
                                            bool hinted = false)
{
    const double xatol_3 = 1e-6 / 3.0;
    const int maxfun = 500;
    const double sqrt_eps = 1.4832396974191326e-08;      // sqrt(2.2e-16)
    const double golden_mean = 0.3819660112501051;        // 0.5*(3 - sqrt(5))
    double a = -10.0, b = S::add(pc.v.us(pc.N - 1), 10.0);
    double fulc = S::add(a, S::mul(golden_mean, S::sub(b, a)));
    double nfc = fulc, xf = fulc;
    double rat = 0.0, e = 0.0;
    double x = xf;
    double fx, ffulc, fnfc;
    {
        double ex, ey;
        if (hinted) { const double2 h = pc.template ld<0>(); ex = h.x; ey = h.y; }
        else pc.eval(x, ex, ey);
        const double dx = S::sub(ex, px), dy = S::sub(ey, py);
        fx = ffulc = fnfc = S::sqrt(S::fma(dy, dy, S::mul(dx, dx)));
    }
    int num = 1;
#pragma unroll 1
    for (;;) {
        const double xm = S::mul(0.5, S::add(a, b));
        const double tol1 = S::add(S::mul(sqrt_eps, fabs(xf)), xatol_3);
        const double tol2 = S::mul(2.0, tol1);
        if (num >= maxfun) break;
        if (!(fabs(S::sub(xf, xm)) > S::sub(tol2, S::mul(0.5, S::sub(b, a))))) break;

        bool golden = true;
        if (fabs(e) > tol1) {                              // try a parabolic step
            double r = S::mul(S::sub(xf, nfc), S::sub(fx, ffulc));
            double q = S::mul(S::sub(xf, fulc), S::sub(fx, fnfc));
            double p = S::sub(S::mul(S::sub(xf, fulc), q), S::mul(S::sub(xf, nfc), r));
            q = S::mul(2.0, S::sub(q, r));
            p = neg_if(p, q > 0.0);
            q = abs_b(q);
            r = e;
            e = rat;
            if ((fabs(p) < fabs(S::mul(S::mul(0.5, q), r))) && (p > S::mul(q, S::sub(a, xf))) &&
                (p < S::mul(q, S::sub(b, xf)))) {
                golden = false;
                rat = S::div(S::add(p, 0.0), q);
                x = S::add(xf, rat);
                if ((S::sub(x, a) < tol2) || (S::sub(b, x) < tol2)) {
                    // si = sign(xm - xf) + ((xm - xf) == 0)
                    rat = neg_if(tol1, S::sub(xm, xf) < 0.0);
                }
            }
        }
        if (golden) {
            e = (xf >= xm) ? S::sub(a, xf) : S::sub(b, xf);
            rat = S::mul(golden_mean, e);
        }
        // x = xf + (sign(rat) + (rat == 0)) * max(|rat|, tol1)
        const double mag = max_f(abs_b(rat), tol1);
        x = S::add(xf, neg_if(mag, rat < 0.0));

        // objective: np.linalg.norm(path(x) - pos) = sqrt(fma(dy, dy, dx*dx))  (BLAS ddot, predef_path.py:246)
        double ex, ey;
        if (hinted && num < 3) {
            // second and third evaluation (see brent_opening): after the second one the best point is x1 > nfc = x0
            // if it was not worse, x0 < nfc = x1 if it was
            const double2 h1 = pc.template ld<1>(), h2 = pc.template ld<2>(), h3 = pc.template ld<3>();
            const double2 h = (num == 1) ? h1 : ((xf > nfc) ? h2 : h3);
            ex = h.x; ey = h.y;
        } else {
            pc.eval(x, ex, ey);
        }
        const double dx = S::sub(ex, px), dy = S::sub(ey, py);
        const double fu = S::sqrt(S::fma(dy, dy, S::mul(dx, dx)));
        num += 1;
        // scipy's update rules (`if fu <= fx: ... else: ...`, _optimize.py:2393-2412) as one
        // branch-free dataflow: the lanes of a warp take different branches on almost every
        // trip, and the branchy form costs ~90 register moves per trip on top of both paths.
        const bool better = (fu <= fx);
        const bool right = (x >= xf);
        const bool c1 = (fu <= fnfc) || (nfc == xf);                          // worse: new point becomes nfc
        const bool c2 = (fu <= ffulc) || (fulc == xf) || (fulc == nfc);       // worse: new point becomes fulc
        const bool shift = better || c1;                                      // (nfc, fnfc) -> (fulc, ffulc)
        const double t = better ? xf : x;                                     // the new bracket end
        const double na = (better == right) ? t : a;
        const double nb = (better != right) ? t : b;
        const double n_fulc = shift ? nfc : (c2 ? x : fulc);
        const double n_ffulc = shift ? fnfc : (c2 ? fu : ffulc);
        const double n_nfc = better ? xf : (c1 ? x : nfc);
        const double n_fnfc = better ? fx : (c1 ? fu : fnfc);
        a = na; b = nb;
        fulc = n_fulc; ffulc = n_ffulc; nfc = n_nfc; fnfc = n_fnfc;
        xf = better ? x : xf; fx = better ? fu : fx;
    }
    nfev = num;
    pc.eval(xf, bx, by);   // the path point of the minimiser: the same operations on the same u as inside the loop
    return xf;
}

// ---------------------------------------------------------------------------
// per-env dynamic state held in registers for the duration of a step
// ---------------------------------------------------------------------------
struct EnvState {
    double px[3], py[3], vx[3], vy[3], an[3], om[3];
    double jx[6], jy[6];
    double path_error, total_reward;
    int steps, flags, episode, n_obs;
};

__device__ __forceinline__ void load_state(const D2dParams &p, int64_t i, EnvState &s)
{
#pragma unroll
    for (int b = 0; b < 3; b++) {
        const double2 P = ld_state(&p.pos[b * p.np + i]), V = ld_state(&p.vel[b * p.np + i]), A = ld_state(&p.ang[b * p.np + i]);
        s.px[b] = P.x; s.py[b] = P.y; s.vx[b] = V.x; s.vy[b] = V.y; s.an[b] = A.x; s.om[b] = A.y;
    }
#pragma unroll
    for (int j = 0; j < 6; j++) {
        const double2 J = ld_state(&p.jacc[j * p.np + i]);
        s.jx[j] = J.x; s.jy[j] = J.y;
    }
    const double2 A = p.acc[i];
    s.path_error = A.x; s.total_reward = A.y;
    const int4 m = p.meta[i];
    s.steps = m.x; s.flags = m.y; s.episode = m.z; s.n_obs = m.w;
}

// Bodies and impulses are final right after the solver: they are stored at once so
// that their ~60 registers are free during the observation / closest-point search.
__device__ __forceinline__ void store_bodies(const D2dParams &p, int64_t i, const EnvState &s)
{
#pragma unroll
    for (int b = 0; b < 3; b++) {
        st_stream(&p.pos[b * p.np + i], make_double2(s.px[b], s.py[b]));
        st_stream(&p.vel[b * p.np + i], make_double2(s.vx[b], s.vy[b]));
        st_stream(&p.ang[b * p.np + i], make_double2(s.an[b], s.om[b]));
    }
#pragma unroll
    for (int j = 0; j < 6; j++) st_stream(&p.jacc[j * p.np + i], make_double2(s.jx[j], s.jy[j]));
}

__device__ __forceinline__ void store_scalars(const D2dParams &p, int64_t i, const EnvState &s)
{
    p.acc[i] = make_double2(s.path_error, s.total_reward);
    p.meta[i] = make_int4(s.steps, s.flags, s.episode, s.n_obs);
}

__device__ __forceinline__ void store_state(const D2dParams &p, int64_t i, const EnvState &s)
{
    store_bodies(p, i, s);
    store_scalars(p, i, s);
}

// Drone.__init__ (Drone.py:9-95): three bodies at rest, impulses cleared.
__device__ __forceinline__ void spawn(EnvState &s, const D2dConst &c, double x, double y, double angle)
{
    const double2 sc1 = sincos2(angle), sc2 = sincos2(angle + D2D_PI);
    const double sn = sc1.x, cs = sc1.y, sn2 = sc2.x, cs2 = sc2.y;
    s.px[0] = x; s.py[0] = y;
    s.px[1] = cs2 * c.R + x; s.py[1] = sn2 * c.R + y;     // Drone.py:37
    s.px[2] = cs * c.R + x; s.py[2] = sn * c.R + y;       // Drone.py:51
#pragma unroll
    for (int b = 0; b < 3; b++) { s.vx[b] = 0.0; s.vy[b] = 0.0; s.an[b] = angle; s.om[b] = 0.0; }
#pragma unroll
    for (int j = 0; j < 6; j++) { s.jx[j] = 0.0; s.jy[j] = 0.0; }
    s.path_error = 0.0; s.total_reward = 0.0;
    s.steps = 0;
    s.flags = D2D_F_INIT;
}

// ---------------------------------------------------------------------------
// thrust + cpSpaceStep (env:400-406).  Returns the frame's
// new rotation (cos, sin) for the collision test / observation.
// ---------------------------------------------------------------------------
// `bias_s` (optional): lane-private shared-memory column (element j at bias_s[j * 32]) for the six joints' bias
// velocities.  ptxas has no registers left for them and otherwise RECOMPUTES them from the positions inside every
// Gauss-Seidel iteration (36 of its 178 FP64 instructions, on the one phase of the kernel that is FP64-pipe bound).
__device__ __forceinline__ void physics_step(EnvState &s, const D2dConst &c, double act_l, double act_r,
                                             double &rc0, double &rs0, double2 *bias_s = nullptr)
{
    const double dt = c.dt;
    // env:400-401 (float64 arithmetic on the float32 action, numpy 1.x promotion)
    const double FL = (act_l / 2.0 + 0.5) * c.force_scale;
    const double FR = (act_r / 2.0 + 0.5) * c.force_scale;

    // cpBodyApplyForceAtLocalPoint on the frame, transform of the CURRENT pose
    double fx, fy, tq;
    {
        double ry, rx;
        sincos_body(s.an[0], ry, rx);
        const double fwxL = -ry * FL, fwyL = rx * FL, fwxR = -ry * FR, fwyR = rx * FR;
        const double rLx = (rx * -c.R + s.px[0]) - s.px[0], rLy = (ry * -c.R + s.py[0]) - s.py[0];
        const double rRx = (rx * c.R + s.px[0]) - s.px[0], rRy = (ry * c.R + s.py[0]) - s.py[0];
        fx = fwxL + fwxR; fy = fwyL + fwyR;
        tq = (rLx * fwyL - rLy * fwxL) + (rRx * fwyR - rRy * fwxR);
    }

    // 1. cpBodyUpdatePosition
    double rc[3], rs[3];
#pragma unroll
    for (int b = 0; b < 3; b++) {
        s.px[b] = s.px[b] + s.vx[b] * dt;
        s.py[b] = s.py[b] + s.vy[b] * dt;
        s.an[b] = s.an[b] + s.om[b] * dt;
        sincos_body(s.an[b], rs[b], rc[b]);
    }
    rc0 = rc[0]; rs0 = rs[0];

    // 4. cpPivotJoint preStep.  Anchors lie on the body x axis, so r1 = Rot(a_motor)*(ax, 0)
    //    = ax*(c, s) and r2 = bx*(c0, s0) with ax in {-d, 0, +d}, bx in {-(R+d), -R, -(R-d), R-d, R, R+d}.
    //    Negating an anchor negates r exactly, so only 2 motor vectors and 3 frame vectors are kept;
    //    joint j uses  r1 = S1[j] * q1[motor],  r2 = S2[j] * q2[Q2[j]]  (signs folded into the FMAs).
    //      j:        0 (L1)  1 (L2)  2 (L3)  3 (R1)  4 (R2)  5 (R3)
    //      r1:       -q1[L]    0     +q1[L]  -q1[R]    0     +q1[R]        q1[m] = d*(c_m, s_m)
    //      r2:       -q2[2]  -q2[1]  -q2[0]  +q2[0]  +q2[1]  +q2[2]        q2 = {R-d, R, R+d}*(c0, s0)
    double q1x[2], q1y[2], q2x[3], q2y[3];
    q1x[0] = rc[1] * c.d; q1y[0] = rs[1] * c.d; q1x[1] = rc[2] * c.d; q1y[1] = rs[2] * c.d;
    q2x[0] = rc[0] * (c.R - c.d); q2y[0] = rs[0] * (c.R - c.d);
    q2x[1] = rc[0] * c.R; q2y[1] = rs[0] * c.R;
    q2x[2] = rc[0] * (c.R + c.d); q2y[2] = rs[0] * (c.R + c.d);
    // Opaque to the optimiser from here on: ptxas otherwise re-derives the lever arms from (cos, sin) and the
    // constants inside every Gauss-Seidel iteration (an FP64 multiply + add each) to save their registers.
#pragma unroll
    for (int k = 0; k < 3; k++) asm volatile("" : "+d"(q2x[k]), "+d"(q2y[k]));
#pragma unroll
    for (int k = 0; k < 2; k++) asm volatile("" : "+d"(q1x[k]), "+d"(q1y[k]));
    double k11[6], k12[6], k22[6], bsx[6], bsy[6];
    // 32-bit shared-window address of the bias column, passed through a (no-op) shuffle so that ptxas keeps it in a
    // register instead of re-deriving it from S2R inside every iteration (as for PathCursor::caddr)
    uint32_t bias_a = 0;
    if (bias_s) bias_a = __shfl_sync(__activemask(), (uint32_t)__cvta_generic_to_shared(bias_s), threadIdx.x & 31);
#pragma unroll
    for (int j = 0; j < 6; j++) {
        const int m = (j < 3) ? 1 : 2;
        const double s1 = (double)((j % 3) - 1);                 // -1, 0, +1
        const double s2 = (j < 3) ? -1.0 : 1.0;
        const int qi = (j < 3) ? 2 - j : j - 3;
        const double r1x = s1 * q1x[m - 1], r1y = s1 * q1y[m - 1];
        const double r2x = s2 * q2x[qi], r2y = s2 * q2y[qi];
        // k_tensor(a = motor, b = frame)
        double a11 = c.msum, a12 = 0.0, a22 = c.msum;
        a11 += r1y * r1y * c.iinv_m; a12 += -r1x * r1y * c.iinv_m; a22 += r1x * r1x * c.iinv_m;
        a11 += r2y * r2y * c.iinv_f; a12 += -r2x * r2y * c.iinv_f; a22 += r2x * r2x * c.iinv_f;
        const double det_inv = 1.0 / (a11 * a22 - a12 * a12);
        k11[j] = a22 * det_inv; k12[j] = -a12 * det_inv; k22[j] = a11 * det_inv;
        // bias = -delta / dt (error_bias = 0 -> bias_coef = 1)
        bsx[j] = ((s.px[0] + r2x) - (s.px[m] + r1x)) * c.bias_coef;
        bsy[j] = ((s.py[0] + r2y) - (s.py[m] + r1y)) * c.bias_coef;
        if (bias_s) asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(bias_a + (uint32_t)(j * 32 * 16)), "d"(bsx[j]), "d"(bsy[j]) : "memory");
    }

    // 5. cpBodyUpdateVelocity (space damping 1; env:376-380 has no physical effect)
    s.vx[0] = s.vx[0] + (c.gx + fx * c.minv_f) * dt;
    s.vy[0] = s.vy[0] + (c.gy + fy * c.minv_f) * dt;
    s.om[0] = s.om[0] + tq * c.iinv_f * dt;
#pragma unroll
    for (int b = 1; b < 3; b++) {
        s.vx[b] = s.vx[b] + c.gx * dt;
        s.vy[b] = s.vy[b] + c.gy * dt;
    }

    // 6. warm start with the cached impulses (dt_coef = 1; all zero on the first step)
#pragma unroll
    for (int j = 0; j < 6; j++) {
        const int m = (j < 3) ? 1 : 2;
        const double s1 = (double)((j % 3) - 1), s2 = (j < 3) ? -1.0 : 1.0;
        const int qi = (j < 3) ? 2 - j : j - 3;
        const double jx = s.jx[j], jy = s.jy[j];
        s.vx[m] -= jx * c.minv_m; s.vy[m] -= jy * c.minv_m;
        if (j % 3 != 1) s.om[m] -= c.iinv_m * (s1 * q1x[m - 1] * jy - s1 * q1y[m - 1] * jx);
        s.vx[0] += jx * c.minv_f; s.vy[0] += jy * c.minv_f;
        s.om[0] += c.iinv_f * (s2 * q2x[qi] * jy - s2 * q2y[qi] * jx);
    }

    // 7. sequential impulses: iterations x 6 joints, Gauss-Seidel order left_1..right_3.
    //    The centre joints (r1 = 0) exert no torque on the motor (adding +-0 is exact).
#pragma unroll 1
    for (int it = 0; it < c.iterations; it++) {
#pragma unroll
        for (int j = 0; j < 6; j++) {
            const int m = (j < 3) ? 1 : 2;
            const double s1 = (double)((j % 3) - 1), s2 = (j < 3) ? -1.0 : 1.0;
            const int qi = (j < 3) ? 2 - j : j - 3;
            const double r2x = s2 * q2x[qi], r2y = s2 * q2y[qi];
            double vax = s.vx[m], vay = s.vy[m];
            if (j % 3 != 1) {
                vax = s.vx[m] - s1 * q1y[m - 1] * s.om[m];
                vay = s.vy[m] + s1 * q1x[m - 1] * s.om[m];
            }
            const double vrx = (s.vx[0] - r2y * s.om[0]) - vax;
            const double vry = (s.vy[0] + r2x * s.om[0]) - vay;
            double bx = bsx[j], by = bsy[j];
            if (bias_s) asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(bx), "=d"(by) : "r"(bias_a + (uint32_t)(j * 32 * 16)) : "memory");
            const double ex = bx - vrx, ey = by - vry;
            double jx = ex * k11[j] + ey * k12[j];
            double jy = ex * k12[j] + ey * k22[j];
            const double ox = s.jx[j], oy = s.jy[j];
            s.jx[j] = ox + jx; s.jy[j] = oy + jy;          // clamp to maxForce*dt = inf is a no-op
            jx = s.jx[j] - ox; jy = s.jy[j] - oy;
            s.vx[m] -= jx * c.minv_m; s.vy[m] -= jy * c.minv_m;
            if (j % 3 != 1) s.om[m] -= c.iinv_m * (s1 * q1x[m - 1] * jy - s1 * q1y[m - 1] * jx);
            s.vx[0] += jx * c.minv_f; s.vy[0] += jy * c.minv_f;
            s.om[0] += c.iinv_f * (r2x * jy - r2y * jx);
        }
    }
}

// ---------------------------------------------------------------------------
// one pass over the obstacles: sensor overlap of the rotated frame box
//  and the reference's nearest-k ranking (env:617-629,
// 948-961: min over the 4 UNROTATED box corners of |corner - centre| - r,
// stable ascending order).
// ---------------------------------------------------------------------------
struct Nearest {
    double d[3];
    int idx[3];
    double ox[3], oy[3];
};

template <class V>
__device__ __forceinline__ bool scan_obstacles(const V &v, const D2dConst &c, double x, double y,
                                               double rc0, double rs0, bool test_collision, Nearest &nr)
{
    bool hit = false;
#pragma unroll
    for (int k = 0; k < 3; k++) { nr.d[k] = CUDART_INF; nr.idx[k] = -1; nr.ox[k] = 0.0; nr.oy[k] = 0.0; }
    const int n = v.n_obs();
    // corners of the unrotated box: (+-hw + x, +-hh + y)   (v1 + body.position, env:955)
    const double cxp = S::add(c.hw, x), cxm = S::add(-c.hw, x);
    const double cyp = S::add(c.hh, y), cym = S::add(-c.hh, y);
    for (int k0 = 0; k0 < n; k0 += 4) {
        // batches of 4: the per-lane table loads of a batch are issued back to back
        double bx[4], by[4], br[4];
#pragma unroll
        for (int j = 0; j < 4; j++) {
            bx[j] = 0.0; by[j] = 0.0; br[j] = 0.0;
            if (k0 + j < n) v.obstacle(k0 + j, bx[j], by[j], br[j]);
        }
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int k = k0 + j;
            if (k >= n) break;
            const double ox = bx[j], oy = by[j], orad = br[j];
            if (test_collision) {
                const double dx = ox - x, dy = oy - y;
                const double qx = rc0 * dx + rs0 * dy, qy = rc0 * dy - rs0 * dx;
                const double ex = max_f(fabs(qx) - c.hw, 0.0), ey = max_f(fabs(qy) - c.hh, 0.0);
                hit = hit || (ex * ex + ey * ey <= orad * orad);      // dist(centre, box) <= r, without the sqrt
            }
            // min over corners of sqrt(dx^2+dy^2) - r: the rounded operations are monotone,
            // so the minimum separates per axis and one sqrt suffices (bit-identical).
            const double dxp = S::sub(cxp, ox), dxm = S::sub(cxm, ox);
            const double dyp = S::sub(cyp, oy), dym = S::sub(cym, oy);
            const double qx2 = min_f(S::mul(dxp, dxp), S::mul(dxm, dxm));
            const double qy2 = min_f(S::mul(dyp, dyp), S::mul(dym, dym));
            const double dist = S::sub(S::sqrt(S::add(qx2, qy2)), orad);
            // stable insertion (strict <: ties keep the lower index first, like sorted()); only
            // (distance, index) travel through the insertion, the centres are re-read afterwards
            if (dist < nr.d[2]) {
                if (dist < nr.d[1]) {
                    nr.d[2] = nr.d[1]; nr.idx[2] = nr.idx[1];
                    if (dist < nr.d[0]) {
                        nr.d[1] = nr.d[0]; nr.idx[1] = nr.idx[0];
                        nr.d[0] = dist; nr.idx[0] = k;
                    } else {
                        nr.d[1] = dist; nr.idx[1] = k;
                    }
                } else {
                    nr.d[2] = dist; nr.idx[2] = k;
                }
            }
        }
    }
#pragma unroll
    for (int j = 0; j < 3; j++) {
        if (nr.idx[j] >= 0) {
            double rr;
            v.obstacle(nr.idx[j], nr.ox[j], nr.oy[j], rr);
        }
    }
    return hit;
}

// ---------------------------------------------------------------------------
// observation (env:631-773).  Writes the 27 values through `put(j, value)` and
// returns the float64 quantities the reward needs.
// ---------------------------------------------------------------------------
struct ObsAux {
    double o0, o1, o3, o4, o5, o6, o7, o8, o9, o10, o17, o18, o19, o20, o23, o24;
    double u_star;
    int nfev;
    int near_packed;
};

template <class V, class Put>
__device__ __forceinline__ void observe(const V &v, const D2dConst &c, EnvState &s, double rc0, double rs0,
                                        const Nearest &nr, Put put, ObsAux &ax, double2 *coef, bool hinted = false,
                                        const double2 *tgt_staged = nullptr)
{
    const double W = c.W, H = c.H;
    const double x = s.px[0], y = s.py[0], alpha = s.an[0];
    const double2 tgt = tgt_staged ? *tgt_staged : v.wp(v.n_wps() - 1);   // env:316-317
    ax.o0 = m1to1(s.vx[0], -1330.0, 2660.0, c.r_vel); put(0, ax.o0);      // env:633-636
    ax.o1 = m1to1(s.vy[0], -1330.0, 2660.0, c.r_vel); put(1, ax.o1);
    put(2, clipd(div_const(s.om[0], 11.7, c.r_omega), -1.0, 1.0));        // env:639-640
    ax.o3 = div_const(alpha, D2D_PI, c.r_pi); put(3, ax.o3);              // env:643-644
    ax.o4 = m1to1(tgt.x - x, 0.0, W, c.r_W); put(4, ax.o4);               // env:648-649
    ax.o5 = m1to1(tgt.y - y, 0.0, H, c.r_H); put(5, ax.o5);
    ax.o6 = m1to1(x, 0.0, W, c.r_W); put(6, ax.o6);                       // env:655-656
    ax.o7 = m1to1(y, 0.0, H, c.r_H); put(7, ax.o7);

    // env:660-720: k nearest obstacles, missing slots = (1, 0, 0)
    ax.o8 = 1.0; ax.o9 = 0.0; ax.o10 = 0.0;
    ax.near_packed = 0;
#pragma unroll
    for (int j = 0; j < 3; j++) {
        double dn = 1.0, sa = 0.0, ca = 0.0;
        if (nr.idx[j] >= 0) {
            const double2 sc = bearing_rot(y - nr.oy[j], x - nr.ox[j], rc0, rs0, -1.0);      // env:693-698, "- alpha - pi"
            dn = m1to1(nr.d[j], 0.0, c.diag, c.r_diag);
            sa = sc.x; ca = sc.y;
        }
        put(8 + 3 * j, dn); put(9 + 3 * j, sa); put(10 + 3 * j, ca);
        if (j == 0) { ax.o8 = dn; ax.o9 = sa; ax.o10 = ca; }
        ax.near_packed |= ((nr.idx[j] < 0 ? 255 : nr.idx[j]) & 255) << (8 * j);
    }

    // env:723-727
    {
        const double2 sc = bearing_rot(s.vy[0], s.vx[0], rc0, rs0, 1.0);
        ax.o17 = sc.x; ax.o18 = sc.y;
        put(17, ax.o17); put(18, ax.o18);
    }

    // env:730-734.  The reference runs the identical fminbound a second time at
    // env:737; same arguments -> same result, evaluated once.
    int nfev;
    PathCursor<V> pc(v, coef);
    double cx, cy;
    const double u = closest_u(pc, x, y, nfev, cx, cy, hinted);
    ax.o19 = m1to1(cx, 0.0, W, c.r_W); ax.o20 = m1to1(cy, 0.0, H, c.r_H);
    put(19, ax.o19); put(20, ax.o20);
    ax.u_star = u; ax.nfev = nfev;

    // env:737-749, predef_path.py:257-266: look-ahead point + permanent latch
    const double L = v.us(v.n_wps() - 1);
    const double u_la = (S::add(u, c.lookahead) > L) ? L : S::add(u, c.lookahead);
    double lx, ly;
    pc.eval(u_la, lx, ly);
    if (fabs(lx - tgt.x) < 10.0 && fabs(ly - tgt.y) < 10.0) s.flags |= D2D_F_LATCHED;
    if (s.flags & D2D_F_LATCHED) { lx = tgt.x; ly = tgt.y; }
    put(21, m1to1(lx, 0.0, W, c.r_W)); put(22, m1to1(ly, 0.0, H, c.r_H));

    // env:752-763: R_w_b rotates the vector by +alpha (transformations.py:9-11) and the frame angle is subtracted
    // again: what is left is the WORLD bearing of the vector itself, i.e. its unit vector (rotation by 0)
    {
        const double2 sc = bearing_rot(ly - y, lx - x, 1.0, 0.0, 1.0);
        ax.o23 = sc.x; ax.o24 = sc.y;
        put(23, ax.o23); put(24, ax.o24);
    }
    {
        const double2 sc = bearing_rot(cy - y, cx - x, 1.0, 0.0, 1.0);
        put(25, sc.x); put(26, sc.y);
    }
}

// ---------------------------------------------------------------------------
// reward + termination (env:423-613)
// ---------------------------------------------------------------------------
struct StepResult {
    double reward, r_CA, r_PA, r_PP, r_coll, r_end, r_AA, d1, ape;
    int end_bits;   // 1 collision | 2 reached | 4 time | 8 alpha
    bool done;
};

// The four end conditions (env:543-570) depend on the pose, the step count and the collision flag only: they
// are known right after the rigid-body step and the obstacle scan, i.e. BEFORE the closest-point search.  The
// step kernel evaluates them there (so that a lane whose episode ends can announce it to the refill queue and
// pull its staged episode towards the L2 while the search runs) and hands the bits to reward_and_done, which
// uses them as they are.  o3, o4, o5 = observation entries 3, 4, 5 (same expressions as observe()).
__device__ __forceinline__ int end_bits_of(const D2dConst &c, const EnvState &s, double o3, double o4, double o5)
{
    const double alpha = o3 * D2D_PI;
    const double tdx = invm1to1(o4, 0.0, c.W), tdy = invm1to1(o5, 0.0, c.H);
    int bits = 0;
    if (s.flags & D2D_F_COLLIDED) bits |= 1;
    if (fabs(tdx) < c.reach_end_radius && fabs(tdy) < c.reach_end_radius) bits |= 2;
    if (fabs(alpha) >= c.AA_angle) bits |= 8;
    if (s.steps == c.n_steps) bits |= 4;
    return bits;
}

__device__ __forceinline__ void reward_and_done(const D2dConst &c, EnvState &s, const ObsAux &o, bool has_obstacles,
                                                double sin_alpha, int bits, StepResult &r)
{
    const double W = c.W, H = c.H;
    // env:423-445: decode the observation back to physical units
    const double vel_x = invm1to1(o.o0, -1330.0, 1330.0), vel_y = invm1to1(o.o1, -1330.0, 1330.0);
    const double alpha = o.o3 * D2D_PI;
    const double pos_x = invm1to1(o.o6, 0.0, W), pos_y = invm1to1(o.o7, 0.0, H);
    const double cpx = invm1to1(o.o19, 0.0, W), cpy = invm1to1(o.o20, 0.0, H);
    // env:435,444-445,482 decode velocity / look-ahead / obstacle angles with arctan2 only to take differences of
    // them (env:484, 537): |((a - b + pi) % 2 pi) - pi| = atan2(|sin(a - b)|, cos(a - b)), and sin / cos of the
    // difference follow from the (sin, cos) pairs of the observation -- same value up to rounding, one atan2 instead of
    // three (and none for the progression reward, which only needs the cosine)

    double lambda_PA = 1.0, lambda_CA = 1.0, r_CA = 0.0, d1 = CUDART_INF;
    if (has_obstacles) {                                          // env:462-523
        d1 = invm1to1(o.o8, 0.0, c.diag);
        const double sin_d = o.o9 * o.o18 - o.o10 * o.o17, cos_d = o.o10 * o.o18 + o.o9 * o.o17;   // obstacle - velocity
        const double DR = c.danger_range, DA = c.danger_angle, kk = c.k_ca;
        if (d1 < DR && c.use_Lambda) {
            lambda_PA = div_const(d1, DR, c.r_DR) / 2.0;
            if (lambda_PA < 0.10) lambda_PA = 0.10;
            lambda_CA = 1.0 - lambda_PA;
        }
        if (d1 < DR) {
            const double angle_diff = atan2_ool(fabs(sin_d), cos_d) * (180.0 / D2D_PI);
            double range_rew = -(((DR + kk * DR) / (d1 + kk * DR)) - 1.0);
            double angle_rew = -(((DA + kk * DA) / (angle_diff + kk * DA)) - 1.0);
            if (angle_rew > 0.0) angle_rew = 0.0;
            if (range_rew > 0.0) range_rew = 0.0;
            r_CA = range_rew + angle_rew;
        }
    }
    // env:527-530
    const double ddx = cpx - pos_x, ddy = cpy - pos_y;
    const double dist_from_path = sqrt(fma(ddy, ddy, ddx * ddx));
    const double r_PA = -(2.0 * div_const(clipd(dist_from_path, 0.0, c.PA_band_edge), c.PA_band_edge, c.r_PA) - 1.0) * c.PA_scale;
    // env:534-539
    const double velocity = sqrt(vel_x * vel_x + vel_y * vel_y);
    const double cos_vel_LA = o.o24 * o.o18 + o.o23 * o.o17;             // cos(look-ahead angle - velocity angle)
    const double r_PP = clipd(cos_vel_LA * (velocity * c.PP_vel_scale), c.PP_rew_min, c.PP_rew_max);
    // env:543-570
    double r_coll = 0.0, r_end = 0.0, r_AA = 0.0;
    if (bits & 1) r_coll = c.rew_collision;
    if (bits & 2) r_end = c.rew_reach_end;
    // env:557-565: np.sin(alpha) of the decoded angle; the sine of the frame angle is already known from the rigid-body
    // step (the decode alpha/pi*pi moves the angle by an ulp at most: rounding-level difference)
    if (alpha > c.AA_band) r_AA = -sin_alpha;
    if (alpha < -c.AA_band) r_AA = sin_alpha;
    if (bits & 8) r_AA = c.rew_AA;
    // env:572
    r.reward = r_AA + r_PA * lambda_PA + r_PP + r_coll + r_CA * lambda_CA + r_end;
    r.r_CA = r_CA * lambda_CA; r.r_PA = r_PA * lambda_PA; r.r_PP = r_PP;
    r.r_coll = r_coll; r.r_end = r_end; r.r_AA = r_AA; r.d1 = d1;
    s.path_error += dist_from_path;                               // env:589-592
    r.ape = s.path_error / (double)s.steps;
    s.total_reward += r.reward;
    r.end_bits = bits;
    r.done = (bits != 0) || (s.flags & D2D_F_DONE);
}

// ---------------------------------------------------------------------------
// fused re-spawn (env:182-392) from the counter-based stream.  Draw order per
// (env, episode) -- mirrored by oracle/philox_reset.py:
//   curriculum: corner, x1, y1, (n_wps-1) azimuths, spawn angle, then per stage
//               stage_2: x, y | stage_3/4: bernoulli, sampler | stage_5: normal(1,4), samplers
//   test:       x, y, spawn angle
//   sampler attempt: u_obs, normal(mean,std) (2 draws), radius
// ---------------------------------------------------------------------------
// QPMI2D.__call__ and calculate_gradient (predef_path.py:88-188) at the same u with ONE segment
// lookup (the obstacle sampler needs both, obstacles.py:65-69).  Note the two reference branches
// differ: the point uses the last row from us[-2] - 0.001 on, the gradient from us[-2] on.
template <class V>
__device__ __forceinline__ void path_point_and_gradient(const V &v, double u, double &x, double &y, double &gx, double &gy)
{
    const int N = v.n_wps(), rows = N - 2;
    const double us0 = v.us(0), us1 = v.us(1), usm2 = v.us(N - 2);
    const double uu = S::mul(u, u);
    const bool first = (u >= us0 && u <= us1);
    const bool p_last = !first && (u >= S::sub(usm2, 0.001));
    const bool g_last = !first && (u >= usm2);
    if (first || (p_last && g_last)) {
        const int r = first ? 0 : rows - 1;
        const double2 a = v.pa(r), b = v.pb(r), c = v.pc(r);
        x = poly2(a.x, b.x, c.x, u, uu); y = poly2(a.y, b.y, c.y, u, uu);
        gx = S::add(S::mul(S::mul(a.x, u), 2.0), b.x); gy = S::add(S::mul(S::mul(a.y, u), 2.0), b.y);
        return;
    }
    int n = min(max((int)(u * ((double)(N - 1) / v.us(N - 1))), 0), N - 2);
    while (n < N - 1 && u > v.us(n + 1)) n++;
    while (n > 0 && !(u > v.us(n))) n--;
    const double un = v.us(n), un1 = v.us(n + 1);
    const double len = S::sub(un1, un);
    const double mu_r = S::div(S::sub(u, un), len), mu_f = S::div(S::sub(un1, u), len);
    const int r1 = (n - 1 < 0) ? rows - 1 : n - 1;
    const double2 a1 = v.pa(r1), b1 = v.pb(r1), c1 = v.pc(r1), a2 = v.pa(n), b2 = v.pb(n), c2 = v.pc(n);
    if (p_last) {                      // u in [us[-2] - 0.001, us[-2]): point from the last row alone
        const double2 a = v.pa(rows - 1), b = v.pb(rows - 1), c = v.pc(rows - 1);
        x = poly2(a.x, b.x, c.x, u, uu); y = poly2(a.y, b.y, c.y, u, uu);
    } else {
        const double x1 = poly2(a1.x, b1.x, c1.x, u, uu), y1 = poly2(a1.y, b1.y, c1.y, u, uu);
        const double x2 = poly2(a2.x, b2.x, c2.x, u, uu), y2 = poly2(a2.y, b2.y, c2.y, u, uu);
        x = S::add(S::mul(mu_r, x2), S::mul(mu_f, x1));
        y = S::add(S::mul(mu_r, y2), S::mul(mu_f, y1));
    }
    const double dx1 = S::add(S::mul(S::mul(a1.x, u), 2.0), b1.x), dy1 = S::add(S::mul(S::mul(a1.y, u), 2.0), b1.y);
    const double dx2 = S::add(S::mul(S::mul(a2.x, u), 2.0), b2.x), dy2 = S::add(S::mul(S::mul(a2.y, u), 2.0), b2.y);
    gx = S::add(S::mul(mu_r, dx2), S::mul(mu_f, dx1));
    gy = S::add(S::mul(mu_r, dy2), S::mul(mu_f, dy1));
}

// generate_obstacles_around_path (obstacles.py:58-89): rejection sampler.  The offset direction
// (cos, sin)(atan2(gy, gx) - pi/2) is (gy, -gx)/|g|, evaluated without the transcendental round trip
// (differs from the reference's value by rounding only; reset parity tolerance 1e-7 px).
// One attempt from its three variates (u ~ U(0.2L, 0.9L), offset ~ N(mean, sd), radius ~ U(10, 50)): the candidate
// circle and whether obstacles.py:71-86 keeps it.  Shared by the serial and the warp-cooperative sampler, so that
// both produce the same obstacles bit for bit.
__device__ __forceinline__ bool obstacle_attempt(EnvView &ev, double u_obs, double dist, double size, bool on_path,
                                                 double &ox, double &oy)
{
    double gx, gy, x, y;
    path_point_and_gradient(ev, u_obs, x, y, gx, gy);
    if (on_path) { ox = x; oy = y; return true; }
    const double inv_g = rsqrt(gx * gx + gy * gy);
    ox = x + dist * (gy * inv_g); oy = y + dist * (-gx * inv_g);
    const double ddx = ox - x, ddy = oy - y;
    return sqrt(ddx * ddx + ddy * ddy) > size + 10.0;
}

__device__ __forceinline__ int sample_obstacles(const D2dParams &p, EnvView &ev, ResetRng &rng, int slot, int count,
                                                double mean, double sd, bool on_path)
{
    const double L = ev.us(ev.n_wps() - 1);
    int made = 0, attempts = 0;
    while (made < count && slot < D2D_MAX_OBS_ENV && attempts < 4096) {   // obstacles.py:63 `while num_obstacles < n`
        attempts++;
        const double u_obs = rng.uniform(0.20 * L, 0.90 * L);
        const double dist = rng.normal(mean, sd);
        const double size = rng.uniform(10.0, 50.0);
        double ox, oy;
        if (obstacle_attempt(ev, u_obs, dist, size, on_path, ox, oy)) {
            p.obxy[ev.row16(slot)] = make_double2(ox, oy);
            p.obr[ev.row8(slot)] = size;
            slot++; made++;
        }
    }
    return slot;
}

// The same sampler run by a whole warp for ONE environment (off-path obstacles): lane l evaluates attempt a0 + l,
// whose four variates are draws draw0 + 4 (a0 + l) .. + 3 of the env's counter-based stream; the accepted attempts
// are kept in attempt order until `count` circles exist, exactly like the serial loop (which stops drawing right
// after the attempt that completes the count).  All arguments are warp-uniform.  Returns the new slot count and,
// through `draws_used`, how far the serial stream has advanced.  The obstacle rejection loop has a data-dependent
// trip count (~1.4 attempts per circle, up to 31 circles) and was the longest chain of the refill roles.
__device__ __noinline__ int sample_obstacles_coop(const D2dParams &p, EnvView &ev, const ResetRng &rng, uint32_t draw0,
                                                     int slot, int count, double mean, double sd, uint32_t &draws_used)
{
    const double L = ev.us(ev.n_wps() - 1);
    const int lane = threadIdx.x & 31;
    int made = 0;
    uint32_t a0 = 0;
    draws_used = 0;
    while (made < count && slot < D2D_MAX_OBS_ENV && a0 < 4096u) {
        const uint32_t d = draw0 + 4u * (a0 + (uint32_t)lane);
        const double u_obs = uniform_from(rng.u01_at(d), 0.20 * L, 0.90 * L);
        const double dist = normal_from(rng.u01_at(d + 1), rng.u01_at(d + 2), mean, sd);
        const double size = uniform_from(rng.u01_at(d + 3), 10.0, 50.0);
        double ox, oy;
        const bool ok = obstacle_attempt(ev, u_obs, dist, size, false, ox, oy);
        const unsigned acc = __ballot_sync(0xffffffffu, ok);
        const int need = min(count - made, D2D_MAX_OBS_ENV - slot);
        const int rank = __popc(acc & ((1u << lane) - 1u));
        if (ok && rank < need) {
            p.obxy[ev.row16(slot + rank)] = make_double2(ox, oy);
            p.obr[ev.row8(slot + rank)] = size;
        }
        const int got = __popc(acc);
        if (got >= need) {
            // the attempt that completed the count: the need-th accepted lane of this batch
            unsigned m = acc;
            for (int k = 1; k < need; k++) m &= m - 1;
            const int last = __ffs(m) - 1;
            slot += need; made += need;
            draws_used = 4u * (a0 + (uint32_t)last + 1u);
            return slot;
        }
        slot += got; made += got;
        a0 += 32u;
        draws_used = 4u * a0;
    }
    return slot;
}

// Samples episode `episode` of env i: writes its per-env tables into table buffer `par`
// (curriculum mode) and returns the spawn pose + obstacle count.  Pure function of
// (seed, global env id, episode): whoever computes it -- a refill warp ahead of time, the
// reset kernel, or the owning lane as a fallback -- gets the same episode.
struct ResetOut {
    double x, y, angle;
    int n_obs;
};

// Off-path obstacles the caller will draw itself, with the whole warp (sample_obstacles_coop), instead of the lane:
// `count` circles starting at draw `draw0` of the episode's stream, then (stage_5) one on-path circle.
struct ObstaclePlan {
    int count;
    uint32_t draw0;
    int then_on_path;
};

template <bool PER_ENV>
__device__ __noinline__ ResetOut reset_env(const D2dParams &p, int64_t i, int episode, int par, double *us_s,
                                           ObstaclePlan *plan = nullptr)
{
    const D2dConst &c = p.c;
    ResetRng rng((uint32_t)(c.env_id_base + i), (uint32_t)episode, c.seed_lo, c.seed_hi);
    if (!PER_ENV) {
        // env:221-311: uniform spawn rectangle of the scenario
        const double x = rng.uniform(c.sp_xmin, c.sp_xmax);
        const double y = rng.uniform(c.sp_ymin, c.sp_ymax);
        const double a = rng.uniform(-D2D_PI / 4, D2D_PI / 4);
        return ResetOut{x, y, a, 0};
    }
    // ---- curriculum: random path (predef_path.py:307-363) ----
    EnvView ev{&p, i, 0, par, us_s};     // us[] of the new path is mirrored in the lane's shared-memory column
    const int N = c.n_wps;
    int corner = 2;                                              // 'DR' when random_path_spawn is False (env:214)
    if (c.random_path_spawn) corner = rng.randint(c.corner_lo, c.corner_hi);   // 1 DL, 2 DR, 3 UL, 4 UR (env:202-211)
    const bool right = (corner == 2 || corner == 4), up = (corner == 3 || corner == 4);
    double wx = right ? rng.uniform(c.W - 180.0, c.W - 100.0) : rng.uniform(100.0, 180.0);
    double wy = up ? rng.uniform(c.H - 180.0, c.H - 100.0) : rng.uniform(100.0, 180.0);
    double az_lo, az_hi;
    if (corner == 1) { az_lo = 0.0; az_hi = D2D_PI / 2; }
    else if (corner == 3) { az_lo = 0.0; az_hi = -D2D_PI / 2; }
    else if (corner == 2) { az_lo = D2D_PI / 2; az_hi = D2D_PI; }
    else { az_lo = -D2D_PI / 2; az_hi = -D2D_PI; }
    const double x0 = wx, y0 = wy;
    // QPMI2D.__init__ (predef_path.py:20-50): chord-length parameter and the
    // quadratic through each interior waypoint and its neighbours (Newton form;
    // the reference inverts the 3x3 Vandermonde matrix -- same polynomial).
    double u_prev2 = 0.0, u_prev = 0.0, u_cur = 0.0;
    double xa = 0.0, ya = 0.0, xb = wx, yb = wy;                  // waypoints k-2, k-1
    p.wps[ev.row16(0)] = make_double2(wx, wy);
    p.us[ev.row8(0)] = 0.0;
    us_s[0] = 0.0;
    for (int k = 1; k < N; k++) {
        const double az = rng.uniform(az_lo, az_hi);
        const double2 sc = sincos2(az);
        const double sn = sc.x, cs = sc.y;
        const double nx = xb + c.seg_len * cs, ny = yb + c.seg_len * sn;
        const double ddx = nx - xb, ddy = ny - yb;
        u_cur = u_prev + sqrt(ddx * ddx + ddy * ddy);
        p.wps[ev.row16(k)] = make_double2(nx, ny);
        p.us[ev.row8(k)] = u_cur;
        us_s[k * D2D_US_STRIDE] = u_cur;
        if (k >= 2) {
            const double u0 = u_prev2, u1 = u_prev, u2 = u_cur;
            const double dx01 = (xb - xa) / (u1 - u0), dx12 = (nx - xb) / (u2 - u1);
            const double dy01 = (yb - ya) / (u1 - u0), dy12 = (ny - yb) / (u2 - u1);
            const double axx = (dx12 - dx01) / (u2 - u0), ayy = (dy12 - dy01) / (u2 - u0);
            const double bxx = dx01 - axx * (u0 + u1), byy = dy01 - ayy * (u0 + u1);
            const double cxx = xa - axx * u0 * u0 - bxx * u0, cyy = ya - ayy * u0 * u0 - byy * u0;
            p.pa[ev.row16(k - 2)] = make_double2(axx, ayy);
            p.pb[ev.row16(k - 2)] = make_double2(bxx, byy);
            p.pc[ev.row16(k - 2)] = make_double2(cxx, cyy);
        }
        xa = xb; ya = yb; xb = nx; yb = ny;
        u_prev2 = u_prev; u_prev = u_cur;
    }
    store_hints(p, ev);                                           // opening points of this path's searches
    double sx = x0, sy = y0;                                      // env:320-321
    const double sa = rng.uniform(-D2D_PI / 4, D2D_PI / 4);       // env:322
    int n_obs = 0;
    if (c.stage == 2) {                                           // env:329-333
        sx = rng.uniform(100.0, c.W - 100.0);
        sy = rng.uniform(100.0, c.H - 100.0);
    } else if (c.stage == 3) {                                    // env:334-347
        if (rng.u01() < c.spawn_chance) {
            if (plan) { plan->count = 1; plan->draw0 = rng.position(); plan->then_on_path = 0; }
            else n_obs = sample_obstacles(p, ev, rng, 0, 1, 0.0, 100.0, false);
        }
    } else if (c.stage == 4) {                                    // env:348-361
        if (rng.u01() < c.spawn_chance) n_obs = sample_obstacles(p, ev, rng, 0, 1, 0.0, 0.0, true);
    } else if (c.stage == 5) {                                    // env:362-372
        double nn = rng.normal(1.0, 4.0);
        if (nn < 0.0 && nn > -3.0) nn = 1.0;
        if (nn < -3.0) nn = 0.0;
        if (nn != 0.0) {
            int want = (int)ceil(nn);                             // `while num_obstacles < n` with float n
            if (want < 0) want = 0;                               // nn == -3 exactly
            if (want > D2D_MAX_OBS_ENV - 1) want = D2D_MAX_OBS_ENV - 1;
            if (plan && want > 0) {
                plan->count = want; plan->draw0 = rng.position(); plan->then_on_path = 1;
            } else {
                n_obs = sample_obstacles(p, ev, rng, 0, want, 0.0, 100.0, false);
                n_obs = sample_obstacles(p, ev, rng, n_obs, 1, 0.0, 0.0, true);
            }
        }
    }
    return ResetOut{sx, sy, sa, n_obs};
}

}  // namespace d2d
