// drone2d.cu -- kernels + C ABI (include/drone2d.h) of the batched Drone2dEnv step.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -shared -Xcompiler -fPIC
#include <cuda_runtime.h>

#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <algorithm>
#include <cstring>
#include <new>
#include <vector>

#include "drone2d_device.cuh"

using namespace d2d;

// ===========================================================================
// kernels
// ===========================================================================

template <bool PER_ENV>
struct ViewOf;
template <>
struct ViewOf<true> {
    typedef EnvView type;
    static __device__ __forceinline__ EnvView make(const D2dParams &p, const D2dSharedTables *, int64_t i, int n_obs, int par,
                                                   const double *us_s = nullptr)
    {
        return EnvView{&p, i, n_obs, par, us_s};
    }
};
template <>
struct ViewOf<false> {
    typedef SharedView type;
    static __device__ __forceinline__ SharedView make(const D2dParams &, const D2dSharedTables *t, int64_t, int, int,
                                                      const double * = nullptr)
    {
        return SharedView{t};
    }
};

// The staging slot of an env is published by a refill lane (stage_observe) and may be consumed by the env's
// main lane in the SAME launch: the publication is a release store of stg_meta, the consumer reads stg_meta
// with acquire semantics and the slot's payload through L2 (never a stale L1 line).
__device__ __forceinline__ void st_release_int4(int4 *ptr, int4 v)
{
    asm volatile("st.release.gpu.global.v4.s32 [%0], {%1, %2, %3, %4};" ::"l"(ptr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ int4 ld_acquire_int4(const int4 *ptr)
{
    int4 v;
    asm volatile("ld.acquire.gpu.global.v4.s32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(ptr) : "memory");
    return v;
}

__device__ __forceinline__ void stage_shared_tables(const D2dParams &p, D2dSharedTables *dst, int tid, int nthreads)
{
    const int *src = reinterpret_cast<const int *>(p.shared);
    int *d = reinterpret_cast<int *>(dst);
    for (int k = tid; k < (int)(sizeof(D2dSharedTables) / sizeof(int)); k += nthreads) d[k] = src[k];
}

// First observation of a freshly spawned drone (at rest at (x, y, angle)) in episode tables
// `par`: Drone2dEnv.reset() -> get_observation() (env:908-912).  Returns the flag bits the
// observation sets (the look-ahead latch can already engage, env:739-741).  One out-of-line
// copy shared by the refill warps, the reset kernel and the in-lane fallback.
template <bool PER_ENV>
__device__ __noinline__ int first_observation(const D2dParams &p, const D2dSharedTables *tab, int64_t i, double x,
                                              double y, double angle, int n_obs, int par, float *row, double *o64,
                                              double *info, int episode, double2 *coef)
{
    EnvState s;
    s.px[0] = x; s.py[0] = y; s.an[0] = angle; s.vx[0] = 0.0; s.vy[0] = 0.0; s.om[0] = 0.0;
    s.flags = 0;
    const double2 sc0 = sincos2(angle);
    const double rs0 = sc0.x, rc0 = sc0.y;
    typename ViewOf<PER_ENV>::type v = ViewOf<PER_ENV>::make(p, tab, i, n_obs, par);
    Nearest nr;
    scan_obstacles(v, p.c, x, y, rc0, rs0, false, nr);
    auto put = [&](int j, double val) {
        if (row) row[j] = (float)val;
        if (o64) o64[j] = val;
    };
    ObsAux ax;
    observe(v, p.c, s, rc0, rs0, nr, put, ax, coef);
    if (info) {
        for (int k = 0; k < D2D_INFO_DOUBLES; k++) info[k] = 0.0;
        info[12] = ax.u_star; info[13] = (double)ax.nfev; info[14] = (double)ax.near_packed; info[15] = (double)episode;
    }
    return s.flags & D2D_F_LATCHED;
}

// Prefetch of episode `meta.z + 1` of env e into its staging slot (tables buffer = the one
// that is not live), in two stages so that neither is longer than a main lane's step:
//   stage A  sample the episode: tables + spawn pose            (stg_meta = {-1, n_obs, 0, target})
//   stage B  first observation of the staged episode, publish   (stg_meta = {target, n_obs, latch, target})
template <bool PER_ENV>
__device__ __forceinline__ void stage_sample(const D2dParams &p, int64_t e, double *us_s)
{
    const int4 m = p.meta[e];
    const int target = m.z + 1;
    const int par = 1;                              // staged episodes always live in table buffer 1
    const ResetOut ro = reset_env<PER_ENV>(p, e, target, par, us_s);
    p.stg_xy[e] = make_double2(ro.x, ro.y);
    p.stg_an[e] = make_double2(ro.angle, 0.0);
    p.stg_meta[e] = make_int4(-1, ro.n_obs, par, target);
}

template <bool PER_ENV>
__device__ __forceinline__ void stage_observe(const D2dParams &p, const D2dSharedTables *tab, int64_t e, int target,
                                              double2 *coef)
{
    const int4 sm = p.stg_meta[e];
    const int par = sm.z;
    // the env moved on (in-lane fallback) or the slot was re-sampled for a later episode: drop the stale task
    if (sm.w != target || sm.x != -1 || p.meta[e].z + 1 != target) return;
    const double2 xy = p.stg_xy[e];
    const int latch = first_observation<PER_ENV>(p, tab, e, xy.x, xy.y, p.stg_an[e].x, sm.y, par,
                                                 p.stg_obs + e * D2D_OBS_DIM, nullptr, nullptr, target, coef);
    st_release_int4(&p.stg_meta[e], make_int4(target, sm.y, latch, target));
}

// ---- sample role of the step kernel: queue A of the previous step -> queue B of this step ----
// Out of line: its registers (and spills) stay out of the main lanes' allocation.
template <bool PER_ENV>
__device__ __noinline__ void refill_sample_role(const D2dParams &p, double *s_us, int first, int total, int q_read, int q_write)
{
    const int tid = threadIdx.x;
    const D2dConst &c = p.c;
    int *qa = p.queue, *qb = p.queue + 3 * p.np, *qb_tgt = p.queue + 6 * p.np;
    int *qa_cnt = p.qcount, *qb_cnt = p.qcount + 3;
    // Each lane draws its task's path and spawn pose; the off-path obstacles of each task (a rejection loop
    // with up to 31 circles, the longest chain of the refill roles) are then drawn by the WHOLE warp, one
    // attempt per lane, with the same result as the lane's own loop (sample_obstacles_coop).
    const int cnt = qa_cnt[q_read];
    const int lane = tid & 31;
    for (int it = 0;; it++) {
        const int t = first + it * total;
        const bool active = t < cnt;
        if (!__any_sync(0xffffffffu, active)) break;
        int e = 0, target = 0;
        ResetOut ro = {0.0, 0.0, 0.0, 0};
        ObstaclePlan plan = {0, 0u, 0};
        if (active) {
            e = qa[(int64_t)q_read * p.np + t];
            target = p.meta[e].z + 1;
            ro = reset_env<PER_ENV>(p, e, target, 1, &s_us[PER_ENV ? tid : 0], PER_ENV ? &plan : nullptr);
        }
        if (PER_ENV) {
            __syncwarp();                                   // the tasks' tables and us[] columns are visible to the warp
            unsigned todo = __ballot_sync(0xffffffffu, active && plan.count > 0);
            while (todo) {
                const int src = __ffs(todo) - 1;
                todo &= todo - 1;
                const int e_s = __shfl_sync(0xffffffffu, e, src), tgt_s = __shfl_sync(0xffffffffu, target, src);
                const int cnt_s = __shfl_sync(0xffffffffu, plan.count, src);
                const uint32_t draw0_s = __shfl_sync(0xffffffffu, plan.draw0, src);
                EnvView ev{&p, (int64_t)e_s, 0, 1, &s_us[(tid & ~31) + src]};
                const ResetRng rng((uint32_t)(c.env_id_base + e_s), (uint32_t)tgt_s, c.seed_lo, c.seed_hi);
                uint32_t used;
                const int n_made = sample_obstacles_coop(p, ev, rng, draw0_s, 0, cnt_s, 0.0, 100.0, used);
                if (lane == src) { ro.n_obs = n_made; plan.draw0 += used; }
            }
            __syncwarp();
            if (active && plan.count > 0 && plan.then_on_path) {
                EnvView ev{&p, (int64_t)e, 0, 1, &s_us[tid]};
                ResetRng rng((uint32_t)(c.env_id_base + e), (uint32_t)target, c.seed_lo, c.seed_hi);
                rng.seek(plan.draw0);
                ro.n_obs = sample_obstacles(p, ev, rng, ro.n_obs, 1, 0.0, 0.0, true);
            }
        }
        if (active) {
            p.stg_xy[e] = make_double2(ro.x, ro.y);
            p.stg_an[e] = make_double2(ro.angle, 0.0);
            p.stg_meta[e] = make_int4(-1, ro.n_obs, 1, target);
            const int slot = atomicAdd(&qb_cnt[q_write], 1);
            qb[(int64_t)q_write * p.np + slot] = e;
            qb_tgt[(int64_t)q_write * p.np + slot] = target;
        }
    }
}

// The fused step: thrust -> cpSpaceStep -> sensor overlap -> nearest-3 ->
// closest/look-ahead point -> 27-float observation -> reward -> termination ->
// episode statistics -> auto-reset.  Grid = [sample blocks | observe blocks | main blocks]:
//   main blocks    one lane per environment; a finished env swaps in its prefetched episode
//                  (spawn pose + first observation + table-buffer flip) and queues a refill;
//   sample blocks  draw the next episode of the envs that consumed their slot one step ago;
//   observe blocks compute the first observation of the episodes drawn one step ago.
// The sampling and the first closest-point search thus run concurrently with the main
// lanes, two steps ahead of any possible use, instead of extending a finished lane.
template <bool PER_ENV>
__global__ void __launch_bounds__(D2D_BLOCK, D2D_MIN_BLOCKS) d2d_step_kernel(const __grid_constant__ D2dParams p)
{
    __shared__ __align__(16) float s_obs[D2D_BLOCK * D2D_OBS_DIM];   // read back as float4 by the coalesced row store
    __shared__ double s_stats[D2D_STATS_DOUBLES];
    __shared__ int s_any_done;
    __shared__ D2dSharedTables s_tab;
    __shared__ double s_us[PER_ENV ? D2D_MAX_WPS * D2D_BLOCK : 1];   // lane-private us[] columns
    __shared__ double2 s_coef[6 * D2D_BLOCK];                        // lane-private polynomial rows (PathCursor)
    __shared__ int s_swap[PER_ENV ? D2D_BLOCK : 1];                  // >= 0: this lane consumed its prefetched episode (= its n_obs)

    const int tid = threadIdx.x;
    const D2dConst &c = p.c;
    const int q_write = (int)(p.step_index % 3), q_read = (int)((p.step_index + 2) % 3), q_clear = (int)((p.step_index + 1) % 3);

    // a dependent kernel launched with programmatic stream serialization (the policy kernel) may start its
    // prologue while this grid runs; it synchronises on our completion before touching our outputs
    asm volatile("griddepcontrol.launch_dependents;");
    if (!PER_ENV) stage_shared_tables(p, &s_tab, tid, D2D_BLOCK);
    if (tid < D2D_STATS_DOUBLES) s_stats[tid] = 0.0;
    if (tid == 0) s_any_done = 0;
    if (PER_ENV) s_swap[tid] = -1;
    asm volatile("griddepcontrol.wait;" ::: "memory");          // from here on the previous kernel's writes are visible
    if (blockIdx.x == 0 && tid == 0) { p.qcount[q_clear] = 0; p.qcount[3 + q_clear] = 0; }
    __syncthreads();

    const long long t_start = clock64();
    int *qa = p.queue, *qb = p.queue + 3 * p.np, *qb_tgt = p.queue + 6 * p.np;
    int *qa_cnt = p.qcount, *qb_cnt = p.qcount + 3;
    if ((int)blockIdx.x < p.refill_blocks + p.observe_blocks) {
        const bool sample_role = (int)blockIdx.x < p.refill_blocks;
        const int total = (sample_role ? p.refill_blocks : p.observe_blocks) * D2D_BLOCK, n_warps = total / 32;
        const int g = (sample_role ? blockIdx.x : blockIdx.x - p.refill_blocks) * D2D_BLOCK + tid;
        const int first = (g & 31) * n_warps + (g >> 5);        // task dealt to (warp, lane): lane-major over warps
        if (sample_role) {
            refill_sample_role<PER_ENV>(p, s_us, first, total, q_read, q_write);
        } else {
            // ---- observe role: queue B of the previous step ----
            const int cnt = qb_cnt[q_read];
            for (int t = first; t < cnt; t += total)
                stage_observe<PER_ENV>(p, &s_tab, (int64_t)qb[(int64_t)q_read * p.np + t],
                                       qb_tgt[(int64_t)q_read * p.np + t], &s_coef[tid]);
        }
        if (p.probe && (tid & 31) == 0 && first < (sample_role ? qa_cnt[q_read] : qb_cnt[q_read])) {
            const unsigned long long dt = (unsigned long long)(clock64() - t_start);
            const int r = sample_role ? 1 : 2;
            atomicMax(&p.role_cycles[r], dt); atomicAdd(&p.role_cycles[8 + r], dt); atomicAdd(&p.role_cycles[16 + r], 1ull);
        }
        return;
    }

    const int64_t blk = (int64_t)blockIdx.x - (p.refill_blocks + p.observe_blocks);
    const int64_t i = blk * D2D_BLOCK + tid;
    const bool valid = i < p.n;
    const unsigned vmask = __ballot_sync(0xffffffffu, valid);       // the lanes of this warp that own an environment
    const int lane = tid & 31;
    bool rewrite_row = false;                                        // this lane's row changed after the early store (episode swap)

    if (valid) {
        EnvState s;
        load_state(p, i, s);
        if (PER_ENV) {
            // Latency hiding for the closest-point search, issued while the solver runs:
            // us[] of the live path goes global -> shared with cp.async (no registers held).
            // (Pulling the polynomial rows into L2 here as well was measured: no gain.)
            const uint32_t base = (uint32_t)i;                  // live tables = buffer 0
            for (int k = 0; k < c.n_wps; k++) {
                const uint32_t dst = (uint32_t)__cvta_generic_to_shared(&s_us[k * D2D_BLOCK + tid]);
#if D2D_L2_HINTS
                asm volatile("cp.async.ca.shared.global.L2::cache_hint [%0], [%1], 8, %2;" ::"r"(dst), "l"(p.us + base + (uint32_t)k * (uint32_t)p.np), "l"(l2_keep()));
#else
                asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(p.us + base + (uint32_t)k * (uint32_t)p.np));
#endif
            }
            // ... and the path points of the search's three opening evaluations (brent_opening) go into rows 0..3 of
            // the lane's cursor column, which the cursor itself only fills at its first locate() (fourth evaluation)
            for (int k = 0; k < (p.use_hints ? D2D_HINTS : 0); k++) {
                const uint32_t dst = (uint32_t)__cvta_generic_to_shared(&s_coef[k * D2D_BLOCK + tid]);
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(p.hint + base + (uint32_t)k * (uint32_t)p.np));
            }
            {   // ... and the target waypoint (row 4 of the column; observe() reads it before the cursor's first locate)
                const uint32_t dst = (uint32_t)__cvta_generic_to_shared(&s_coef[4 * D2D_BLOCK + tid]);
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(p.wps + base + (uint32_t)(c.n_wps - 1) * (uint32_t)p.np));
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
        }

        float2 act;
        if (p.actions) {
            act = reinterpret_cast<const float2 *>(p.actions)[i];
        } else {
            uint32_t o[4];
            philox4x32_10((uint32_t)(c.env_id_base + i), (uint32_t)p.step_index, (uint32_t)(p.step_index >> 32),
                          D2D_STREAM_ACTION, c.seed_lo, c.seed_hi, o);
            act.x = (float)(2.0 * u01_from_bits(o[0], o[1]) - 1.0);
            act.y = (float)(2.0 * u01_from_bits(o[2], o[3]) - 1.0);
        }

        double rc0, rs0;
        if (!p.skip_physics) {
            // the warp's own 32 observation rows (3 456 B, idle until observe()) hold the solver's bias terms (3 072 B)
            physics_step(s, c, (double)act.x, (double)act.y, rc0, rs0,
                         reinterpret_cast<double2 *>(&s_obs[(tid & ~31) * D2D_OBS_DIM]) + (tid & 31));
        } else {
            // replay mode (tests): the injected state already is the outcome of space.step
            const double2 sc = sincos2(s.an[0]);
            rs0 = sc.x; rc0 = sc.y;
        }
        s.steps += 1;                                               // env:407
        store_bodies(p, i, s);
        if (p.path_ring)                                            // env:410-415: add_postion_to_flight_path (x, H - y)
            p.path_ring[(int64_t)((s.steps - 1) % p.path_ring_len) * p.np + i] = make_float2((float)s.px[0], (float)(c.H - s.py[0]));
        const long long t_phys = clock64();

        float *row = &s_obs[tid * D2D_OBS_DIM];
        const int par = 0;                                          // live tables = buffer 0
        if (PER_ENV) asm volatile("cp.async.wait_all;" ::: "memory");      // us[] column of this lane is in shared memory
        typename ViewOf<PER_ENV>::type v = ViewOf<PER_ENV>::make(p, &s_tab, i, s.n_obs, par, &s_us[tid]);
        Nearest nr;
        const bool hit = scan_obstacles(v, c, s.px[0], s.py[0], rc0, rs0, true, nr);
        if (hit) s.flags |= D2D_F_COLLIDED;                         // begin callback, env:17-19

        // end conditions, known before the search (see end_bits_of)
        int end_bits;
        {
            const double2 tgt = PER_ENV ? s_coef[4 * D2D_BLOCK + tid] : v.wp(v.n_wps() - 1);
            end_bits = end_bits_of(c, s, div_const(s.an[0], D2D_PI, c.r_pi), m1to1(tgt.x - s.px[0], 0.0, c.W, c.r_W),
                                   m1to1(tgt.y - s.py[0], 0.0, c.H, c.r_H));
        }
        const bool will_swap = (end_bits != 0 || (s.flags & D2D_F_DONE)) && p.auto_reset;   // = r.done below
        if (will_swap) {
            // ask the next step's refill blocks for the episode after the next one, and start pulling the
            // staged episode (spawn pose, first observation) into the L2: both round trips overlap the search
            qa[(int64_t)q_write * p.np + atomicAdd(&qa_cnt[q_write], 1)] = (int)i;
            asm volatile("prefetch.global.L2 [%0];" ::"l"(&p.stg_meta[i]));
            asm volatile("prefetch.global.L2 [%0];" ::"l"(&p.stg_xy[i]));
            asm volatile("prefetch.global.L2 [%0];" ::"l"(&p.stg_an[i]));
            asm volatile("prefetch.global.L2 [%0];" ::"l"(p.stg_obs + i * D2D_OBS_DIM));
            asm volatile("prefetch.global.L2 [%0];" ::"l"(p.stg_obs + i * D2D_OBS_DIM + D2D_OBS_DIM - 1));
        }
        if (PER_ENV) {
            // ... and its staged tables, by the whole warp (the copy at the end of the kernel reads them)
            unsigned m = __ballot_sync(vmask, will_swap);
            const uint32_t np32 = (uint32_t)p.np;
            while (m) {
                const int src = __ffs(m) - 1;
                m &= m - 1;
                const uint32_t e = (uint32_t)(blk * D2D_BLOCK + (tid & ~31) + src);
#pragma unroll
                for (int r = 0; r < 2; r++) {
                    const int t = lane + 32 * r;
                    if (t < 4 * c.n_wps) {
                        const double2 *a = (t & 3) == 0 ? p.wps : ((t & 3) == 1 ? p.pa : ((t & 3) == 2 ? p.pb : p.pc));
                        asm volatile("prefetch.global.L2 [%0];" ::"l"(a + p.stg_off16 + (uint32_t)(t >> 2) * np32 + e));
                    }
                    if (t < c.n_wps) asm volatile("prefetch.global.L2 [%0];" ::"l"(p.us + p.stg_off8 + (uint32_t)t * np32 + e));
                    else if (t - c.n_wps < D2D_MAX_OBS_ENV) asm volatile("prefetch.global.L2 [%0];" ::"l"(p.obr + p.stg_off8 + (uint32_t)(t - c.n_wps) * np32 + e));
                }
                if (lane < D2D_MAX_OBS_ENV) asm volatile("prefetch.global.L2 [%0];" ::"l"(p.obxy + p.stg_off16 + (uint32_t)lane * np32 + e));
            }
        }

        double *o64 = p.obs64 ? p.obs64 + i * D2D_OBS_DIM : nullptr;
        auto put = [&](int j, double val) {
            row[j] = (float)val;
            if (o64) o64[j] = val;
        };
        ObsAux ax;
        const long long t_scan = clock64();
        observe(v, c, s, rc0, rs0, nr, put, ax, &s_coef[tid], PER_ENV && p.use_hints,
                PER_ENV ? &s_coef[4 * D2D_BLOCK + tid] : nullptr);
        // The warp's 32 observation rows are contiguous in the output ([N][27] float32): store them now, coalesced,
        // instead of at block end -- when the outputs live in mapped host memory (d2d_step_host) the PCIe writes
        // then start while reward / swap and the block's slower warps are still running.
        __syncwarp(vmask);
        {
            const int64_t w0 = blk * D2D_BLOCK + (tid & ~31);
            const float *src = &s_obs[(tid & ~31) * D2D_OBS_DIM];
            float *dst = p.obs + w0 * D2D_OBS_DIM;
            if (vmask == 0xffffffffu) {
                const float4 *s4 = reinterpret_cast<const float4 *>(src);
                float4 *d4 = reinterpret_cast<float4 *>(dst);
                for (int k = lane; k < 32 * D2D_OBS_DIM / 4; k += 32) __stcs(&d4[k], s4[k]);   // streamed: read by the policy, not by the next step
            } else {
                const int nv = __popc(vmask);                      // last warp of the batch: lanes 0..nv-1 own an env
                for (int k = lane; k < nv * D2D_OBS_DIM; k += nv) dst[k] = src[k];
            }
        }
        const long long t_obs = clock64();

        StepResult r;
        reward_and_done(c, s, ax, v.n_obs() > 0, rs0, end_bits, r);
        p.reward[i] = (float)r.reward;
        p.done[i] = r.done ? 1 : 0;
        if (p.info && (r.done || !p.info_done_only)) {
            double *f = p.info + i * D2D_INFO_DOUBLES;
            f[0] = r.reward; f[1] = r.r_CA; f[2] = r.r_PA; f[3] = r.r_PP; f[4] = r.r_coll; f[5] = r.r_end;
            f[6] = r.r_AA; f[7] = (double)s.steps; f[8] = r.d1; f[9] = r.ape; f[10] = s.total_reward;
            f[11] = (double)r.end_bits; f[12] = ax.u_star; f[13] = (double)ax.nfev;
            f[14] = (double)ax.near_packed; f[15] = (double)s.episode;
        }
        if (r.done) {
            if (!(s.flags & D2D_F_DONE)) {
                // episode statistics (env:593-613; main.py:272-282; tensorboardlogger.py:80-88)
                const int b = r.end_bits;
                s_any_done = 1;
                atomicAdd(&s_stats[0], 1.0);
                if (b & 2) atomicAdd(&s_stats[1], 1.0);
                if (b & (1 | 4 | 8)) atomicAdd(&s_stats[2], 1.0);
                if ((b & 1) && !(b & (2 | 4 | 8))) atomicAdd(&s_stats[3], 1.0);   // later branches overwrite n_collisions
                atomicAdd(&s_stats[4], (double)s.steps);
                atomicAdd(&s_stats[5], r.ape);
                atomicAdd(&s_stats[6], s.total_reward);
                atomicAdd(&s_stats[7], r.reward);
                atomicAdd(&s_stats[8], r.r_coll);
                atomicAdd(&s_stats[9], r.r_CA);
                atomicAdd(&s_stats[10], r.r_PA);
                atomicAdd(&s_stats[11], r.r_PP);
                atomicAdd(&s_stats[12], r.r_end);
                atomicAdd(&s_stats[13], r.r_AA);
                p.last[i] = make_double2(s.total_reward, (double)s.steps);
                p.last[p.np + i] = make_double2(r.ape, (double)b);
            }
            if (!p.auto_reset) {
                s.flags |= D2D_F_DONE;
            } else {
                if (p.terminal_obs) {
                    float *t = p.terminal_obs + i * D2D_OBS_DIM;
#pragma unroll
                    for (int j = 0; j < D2D_OBS_DIM; j++) t[j] = row[j];
                }
                const int next_ep = s.episode + 1;
                const int4 sm = ld_acquire_int4(&p.stg_meta[i]);
                double nx, ny, na;
                int n_obs, new_flags;
                if (sm.x == next_ep) {
                    // swap in the prefetched episode: its tables go from the staging buffer (1) into the
                    // live buffer (0) (warp-cooperative copy at the end of the kernel), the spawn pose and
                    // first observation are taken here.  (Copying ~0.8 KB once per episode keeps the
                    // per-step working set dense: with a per-env buffer flip, live and staged rows of
                    // neighbouring envs share cache lines and the table traffic doubles.)
                    const double2 xy = __ldcg(&p.stg_xy[i]);
                    nx = xy.x; ny = xy.y; na = __ldcg(&p.stg_an[i]).x;
                    n_obs = sm.y;
                    new_flags = D2D_F_INIT | (sm.z ? D2D_F_LATCHED : 0);
                    if (PER_ENV) s_swap[tid] = n_obs;                // tables are copied by the whole warp, below
                    const float *sob = p.stg_obs + i * D2D_OBS_DIM;
#pragma unroll
                    for (int j = 0; j < D2D_OBS_DIM; j++) row[j] = __ldcg(&sob[j]);
                    rewrite_row = true;
                } else {
                    // slot not ready (episode ended one step after the previous one, or state was
                    // injected): sample in the lane, into the live table buffer
                    atomicAdd(&s_stats[15], 1.0);
                    const ResetOut ro = reset_env<PER_ENV>(p, i, next_ep, par, &s_us[PER_ENV ? tid : 0]);
                    nx = ro.x; ny = ro.y; na = ro.angle; n_obs = ro.n_obs;
                    new_flags = D2D_F_INIT;
                    new_flags |= first_observation<PER_ENV>(p, &s_tab, i, nx, ny, na, n_obs, par, row, nullptr, nullptr,
                                                            next_ep, &s_coef[tid]);
                    rewrite_row = true;
                }
                spawn(s, c, nx, ny, na);
                s.flags = new_flags; s.episode = next_ep; s.n_obs = n_obs;
                store_bodies(p, i, s);
            }
        }
        store_scalars(p, i, s);
        if (p.probe && (tid & 31) == 0) {
            const long long t_end = clock64();
            const unsigned long long d[7] = {(unsigned long long)(t_end - t_start), 0, 0, (unsigned long long)(t_phys - t_start),
                                             (unsigned long long)(t_scan - t_phys), (unsigned long long)(t_obs - t_scan),
                                             (unsigned long long)(t_end - t_obs)};
#pragma unroll
            for (int r = 0; r < 7; r++) {
                if (r == 1 || r == 2) continue;
                atomicMax(&p.role_cycles[r], d[r]); atomicAdd(&p.role_cycles[8 + r], d[r]); atomicAdd(&p.role_cycles[16 + r], 1ull);
            }
        }
    }
    if (PER_ENV) {
        // Table copy of the envs that consumed their prefetched episode, by the whole warp: one element
        // per lane and round, all loads of an env in flight together (a finished lane copying its ~60-100
        // strided elements alone stretched the slowest warp of the launch by ~15k cycles).
        const int my_nobs = s_swap[tid];
        unsigned m = __ballot_sync(0xffffffffu, my_nobs >= 0);
        const int lane = tid & 31;
        const uint32_t np32 = (uint32_t)p.np;
        while (m) {
            const int src = __ffs(m) - 1;
            m &= m - 1;
            const uint32_t e = (uint32_t)(blk * D2D_BLOCK + (tid & ~31) + src);
            const int nob = __shfl_sync(0xffffffffu, my_nobs, src);
            const int n16 = 4 * c.n_wps;                           // wps, pa, pb, pc rows: element t = 4 * row + array
            auto arr16 = [&](int a) -> double2 * { return a == 0 ? p.wps : (a == 1 ? p.pa : (a == 2 ? p.pb : p.pc)); };
            double2 v16[2], vob;
            double v8[2];
#pragma unroll
            for (int r = 0; r < 2; r++) {
                const int t = lane + 32 * r;
                if (t < n16) v16[r] = arr16(t & 3)[p.stg_off16 + (uint32_t)(t >> 2) * np32 + e];
                const int u = t - c.n_wps;                         // 8-byte elements: us rows, then obstacle radii
                if (t < c.n_wps) v8[r] = p.us[p.stg_off8 + (uint32_t)t * np32 + e];
                else if (u < nob) v8[r] = p.obr[p.stg_off8 + (uint32_t)u * np32 + e];
            }
            if (lane < nob) vob = p.obxy[p.stg_off16 + (uint32_t)lane * np32 + e];
            double2 vh;
            if (lane < D2D_HINTS) vh = p.hint[p.stg_off16 + (uint32_t)lane * np32 + e];
            if (lane < D2D_HINTS) p.hint[(uint32_t)lane * np32 + e] = vh;
#pragma unroll
            for (int r = 0; r < 2; r++) {
                const int t = lane + 32 * r;
                if (t < n16) arr16(t & 3)[(uint32_t)(t >> 2) * np32 + e] = v16[r];
                const int u = t - c.n_wps;
                if (t < c.n_wps) p.us[(uint32_t)t * np32 + e] = v8[r];
                else if (u < nob) p.obr[(uint32_t)u * np32 + e] = v8[r];
            }
            if (lane < nob) p.obxy[(uint32_t)lane * np32 + e] = vob;
        }
    }
    __syncthreads();

    // rows of envs that swapped in a new episode: their first observation replaces the terminal one that was
    // stored early (27 lanes, one coalesced 108-byte store per row)
    {
        unsigned m = __ballot_sync(0xffffffffu, rewrite_row);
        while (m) {
            const int src = __ffs(m) - 1;
            m &= m - 1;
            if (lane < D2D_OBS_DIM)
                p.obs[(blk * D2D_BLOCK + (tid & ~31) + src) * D2D_OBS_DIM + lane] = s_obs[((tid & ~31) + src) * D2D_OBS_DIM + lane];
        }
    }
    if (s_any_done && (tid < 14 || tid == 15)) atomicAdd(&p.stats[tid], s_stats[tid]);
}

// Drone2dEnv.reset() for all / masked envs: re-spawn + first observation, and prefetch of
// the episode after it.
template <bool PER_ENV>
__global__ void __launch_bounds__(D2D_BLOCK) d2d_reset_kernel(const __grid_constant__ D2dParams p)
{
    __shared__ D2dSharedTables s_tab;
    __shared__ double2 s_coef[6 * D2D_BLOCK];
    __shared__ double s_us[PER_ENV ? D2D_MAX_WPS * D2D_BLOCK : 1];
    const int tid = threadIdx.x;
    const int64_t i = (int64_t)blockIdx.x * D2D_BLOCK + tid;
    if (!PER_ENV) stage_shared_tables(p, &s_tab, tid, D2D_BLOCK);
    __syncthreads();
    if (i >= p.n) return;
    if (p.mask && !p.mask[i]) return;
    const D2dConst &c = p.c;
    const int4 m = p.meta[i];
    const int episode = (m.y & D2D_F_INIT) ? m.z + 1 : 0;
    const int par = 0;
    const ResetOut ro = reset_env<PER_ENV>(p, i, episode, par, &s_us[PER_ENV ? tid : 0]);
    const int latch = first_observation<PER_ENV>(p, &s_tab, i, ro.x, ro.y, ro.angle, ro.n_obs, par,
                                                 p.obs ? p.obs + i * D2D_OBS_DIM : nullptr,
                                                 p.obs64 ? p.obs64 + i * D2D_OBS_DIM : nullptr,
                                                 p.info ? p.info + i * D2D_INFO_DOUBLES : nullptr, episode, &s_coef[tid]);
    EnvState s;
    spawn(s, c, ro.x, ro.y, ro.angle);
    s.flags = D2D_F_INIT | latch;
    s.episode = episode; s.n_obs = ro.n_obs;
    store_state(p, i, s);
    stage_sample<PER_ENV>(p, i, &s_us[PER_ENV ? tid : 0]);
    stage_observe<PER_ENV>(p, &s_tab, i, episode + 1, &s_coef[tid]);
}

template <bool PER_ENV>
__global__ void d2d_closest_u_kernel(const __grid_constant__ D2dParams p, const double *pos, double *u_out, int *nfev_out)
{
    __shared__ D2dSharedTables s_tab;
    __shared__ double2 s_coef[6 * D2D_BLOCK];
    const int tid = threadIdx.x;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + tid;
    if (!PER_ENV) stage_shared_tables(p, &s_tab, tid, blockDim.x);
    __syncthreads();
    if (i >= p.n) return;
    typename ViewOf<PER_ENV>::type v = ViewOf<PER_ENV>::make(p, &s_tab, i, 0, 0);
    int nfev;
    PathCursor<typename ViewOf<PER_ENV>::type> pc(v, &s_coef[tid]);
    double bx, by;
    u_out[i] = closest_u(pc, pos[2 * i], pos[2 * i + 1], nfev, bx, by);
    nfev_out[i] = nfev;
}

// AoS blob <-> SoA state (d2d_get_state / d2d_set_state)
__global__ void d2d_pack_state_kernel(const D2dParams p, double *blob, int to_blob)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= p.n) return;
    double *b = blob + i * D2D_STATE_DOUBLES;
    EnvState s;
    if (to_blob) {
        load_state(p, i, s);
        for (int k = 0; k < 3; k++) {
            b[6 * k + 0] = s.px[k]; b[6 * k + 1] = s.py[k]; b[6 * k + 2] = s.vx[k]; b[6 * k + 3] = s.vy[k];
            b[6 * k + 4] = s.an[k]; b[6 * k + 5] = s.om[k];
        }
        for (int j = 0; j < 6; j++) { b[18 + 2 * j] = s.jx[j]; b[19 + 2 * j] = s.jy[j]; }
        b[30] = (s.flags & D2D_F_COLLIDED) ? 1.0 : 0.0;
        b[31] = (s.flags & D2D_F_LATCHED) ? 1.0 : 0.0;
        b[32] = (double)s.steps; b[33] = s.path_error; b[34] = s.total_reward; b[35] = (double)s.episode;
    } else {
        load_state(p, i, s);   // keeps n_obs
        for (int k = 0; k < 3; k++) {
            s.px[k] = b[6 * k + 0]; s.py[k] = b[6 * k + 1]; s.vx[k] = b[6 * k + 2]; s.vy[k] = b[6 * k + 3];
            s.an[k] = b[6 * k + 4]; s.om[k] = b[6 * k + 5];
        }
        for (int j = 0; j < 6; j++) { s.jx[j] = b[18 + 2 * j]; s.jy[j] = b[19 + 2 * j]; }
        s.flags = D2D_F_INIT | (b[30] != 0.0 ? D2D_F_COLLIDED : 0) | (b[31] != 0.0 ? D2D_F_LATCHED : 0);
        s.steps = (int)b[32]; s.path_error = b[33]; s.total_reward = b[34]; s.episode = (int)b[35];
        store_state(p, i, s);
    }
}

// AoS host-layout tables <-> per-env SoA tables (d2d_set_tables / d2d_get_tables)
__global__ void d2d_pack_tables_kernel(const D2dParams p, double *wps, double *us, double *xp, double *yp, int *n_obs,
                                       double *obst, int to_host)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= p.n) return;
    for (int k = 0; k < D2D_MAX_WPS; k++) {
        const int64_t r = (int64_t)k * p.np + i;
        double *w = wps + (i * D2D_MAX_WPS + k) * 2;
        double *x = xp + (i * D2D_MAX_WPS + k) * 3, *y = yp + (i * D2D_MAX_WPS + k) * 3;
        if (to_host) {
            const double2 W = p.wps[r], A = p.pa[r], B = p.pb[r], C = p.pc[r];
            w[0] = W.x; w[1] = W.y; us[i * D2D_MAX_WPS + k] = p.us[r];
            x[0] = A.x; x[1] = B.x; x[2] = C.x; y[0] = A.y; y[1] = B.y; y[2] = C.y;
        } else {
            p.wps[r] = make_double2(w[0], w[1]); p.us[r] = us[i * D2D_MAX_WPS + k];
            p.pa[r] = make_double2(x[0], y[0]); p.pb[r] = make_double2(x[1], y[1]); p.pc[r] = make_double2(x[2], y[2]);
        }
    }
    for (int k = 0; k < D2D_MAX_OBS_ENV; k++) {
        const int64_t r = (int64_t)k * p.np + i;
        double *o = obst + (i * D2D_MAX_OBS_ENV + k) * 3;
        if (to_host) {
            const double2 xy = p.obxy[r];
            o[0] = xy.x; o[1] = xy.y; o[2] = p.obr[r];
        } else {
            p.obxy[r] = make_double2(o[0], o[1]); p.obr[r] = o[2];
        }
    }
    if (to_host) {
        n_obs[i] = p.meta[i].w;
    } else {
        int4 m = p.meta[i];
        m.w = n_obs[i];
        p.meta[i] = m;
        const EnvView ev{&p, i, 0, 0, nullptr};
        store_hints(p, ev);                                      // the injected path's opening points
    }
}

// Generalised advantage estimation over a device-resident rollout [T][N] (SB3
// RolloutBuffer.compute_returns_and_advantage): one lane per environment walks the time axis
// backwards; every access is coalesced over N.  Memory-bound: 5 floats read + 2 written per (t, n).
__global__ void d2d_gae_kernel(const float *__restrict__ rewards, const float *__restrict__ values,
                               const uint8_t *__restrict__ episode_starts, const float *__restrict__ last_values,
                               const uint8_t *__restrict__ last_dones, int T, int64_t N, float gamma, float lam,
                               float *__restrict__ advantages, float *__restrict__ returns)
{
    const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    float next_value = last_values[n];
    float next_non_terminal = 1.0f - (float)(last_dones[n] != 0);
    float gae = 0.0f;
    for (int t = T - 1; t >= 0; t--) {
        const int64_t k = (int64_t)t * N + n;
        const float v = values[k];
        const float delta = rewards[k] + gamma * next_value * next_non_terminal - v;
        gae = delta + gamma * lam * next_non_terminal * gae;
        advantages[k] = gae;
        returns[k] = gae + v;
        next_value = v;
        next_non_terminal = 1.0f - (float)(episode_starts[k] != 0);
    }
}

#include "drone2d_policy.cuh"   // the policy in the loop: tcgen05 (TMEM-accumulated tf32x3) MLP kernel

// ===========================================================================
// host side: handle + C ABI
// ===========================================================================

static thread_local char g_err[512] = "";

static int fail(int code, const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
    return code;
}

#define CUDA_TRY(expr)                                                                            \
    do {                                                                                          \
        cudaError_t e_ = (expr);                                                                  \
        if (e_ != cudaSuccess) return fail(D2D_ECUDA, "%s: %s", #expr, cudaGetErrorString(e_));  \
    } while (0)

struct d2d_env {
    d2d_config cfg;
    D2dLayout lay;
    D2dParams p;                 // template parameter block (I/O pointers filled per call)
    int device;
    bool tables_set;             // test mode: scenario uploaded
    D2dSharedTables *shared_dev;
    double *stats_dev;
    uint64_t step_index;
    int64_t launches;
    int policy_grid_cap;         // persistent grid size of the policy kernel (0 = not queried yet)
    // L2 residency control of the hot prefix of the state (see d2d_create)
    void *l2_base;
    size_t l2_bytes;
    float l2_hit_ratio;
    // host-boundary staging (d2d_step_host)
    float *h_act, *h_obs, *h_rew;
    uint8_t *h_done;
    float *d_act, *d_obs, *d_rew;
    uint8_t *d_done;
};

struct DeviceGuard {
    int prev;
    bool switched;
    explicit DeviceGuard(int dev) : prev(-1), switched(false)
    {
        if (cudaGetDevice(&prev) == cudaSuccess && prev != dev) switched = (cudaSetDevice(dev) == cudaSuccess);
    }
    ~DeviceGuard()
    {
        if (switched) cudaSetDevice(prev);
    }
};

// cpMomentForPoly on the box vertices Poly.create_box produces (Drone.py:16-17,33-34)
static double moment_for_box(double m, double w, double h)
{
    const double hw = w / 2, hh = h / 2;
    const double vx[4] = {hw, hw, -hw, -hw}, vy[4] = {-hh, hh, hh, -hh};
    double sum1 = 0.0, sum2 = 0.0;
    for (int i = 0; i < 4; i++) {
        const int j = (i + 1) % 4;
        const double a = vx[j] * vy[i] - vy[j] * vx[i];
        const double b = (vx[i] * vx[i] + vy[i] * vy[i]) + (vx[i] * vx[j] + vy[i] * vy[j]) + (vx[j] * vx[j] + vy[j] * vy[j]);
        sum1 += a * b;
        sum2 += a;
    }
    return (m * sum1) / (6.0 * sum2);
}

static int check_config(const d2d_config *c)
{
    if (!c) return fail(D2D_EINVAL, "config is NULL");
    if (c->abi_version != D2D_ABI_VERSION) return fail(D2D_EINVAL, "abi_version %d != %d", c->abi_version, D2D_ABI_VERSION);
    if (c->mode != D2D_MODE_TEST && c->mode != D2D_MODE_CURRICULUM) return fail(D2D_EINVAL, "mode must be 0 (test) or 1 (curriculum)");
    if (c->mode == D2D_MODE_CURRICULUM) {
        if (c->stage < 1 || c->stage > 5) return fail(D2D_EINVAL, "stage must be 1..5 in curriculum mode");
        if (c->n_wps < 4 || c->n_wps > D2D_MAX_WPS) return fail(D2D_EINVAL, "n_wps must be in [4, %d]", D2D_MAX_WPS);
        if (c->spawn_corner_lo < 1 || c->spawn_corner_hi > 4 || c->spawn_corner_lo > c->spawn_corner_hi)
            return fail(D2D_EINVAL, "spawn_corners must satisfy 1 <= lo <= hi <= 4");
    }
    if (c->n_steps < 1) return fail(D2D_EINVAL, "n_steps must be >= 1");
    if (c->iterations < 1 || c->iterations > 1000) return fail(D2D_EINVAL, "iterations out of range");
    if (!(c->dt > 0) || !(c->mass_f > 0) || !(c->mass_m > 0) || !(c->width > c->height) || !(c->height > 6))
        return fail(D2D_EINVAL, "drone constants out of range");
    if (!(c->screen_x > 0) || !(c->screen_y > 0)) return fail(D2D_EINVAL, "screen size must be positive");
    return D2D_OK;
}

static void derive_constants(const d2d_config *g, D2dConst *c, uint64_t seed, int64_t env_id_base)
{
    memset(c, 0, sizeof *c);
    c->mode = g->mode; c->stage = g->stage; c->n_steps = g->n_steps; c->n_wps = g->n_wps;
    c->corner_lo = g->spawn_corner_lo; c->corner_hi = g->spawn_corner_hi; c->random_path_spawn = g->random_path_spawn;
    c->use_Lambda = g->use_Lambda; c->iterations = g->iterations;
    c->W = g->screen_x; c->H = g->screen_y; c->seg_len = g->path_segment_length; c->lookahead = g->lookahead;
    c->danger_range = g->danger_range; c->danger_angle = g->danger_angle; c->k_ca = g->abs_inv_CA_min_rew;
    c->PA_band_edge = g->PA_band_edge; c->PA_scale = g->PA_scale; c->PP_vel_scale = g->PP_vel_scale;
    c->PP_rew_max = g->PP_rew_max; c->PP_rew_min = g->PP_rew_min; c->rew_collision = g->rew_collision;
    c->reach_end_radius = g->reach_end_radius; c->rew_reach_end = g->rew_reach_end;
    c->AA_angle = g->AA_angle; c->AA_band = g->AA_band; c->rew_AA = g->rew_AA;
    c->spawn_chance = g->stage_spawn_chance;
    c->sp_xmin = g->spawn_xmin; c->sp_xmax = g->spawn_xmax; c->sp_ymin = g->spawn_ymin; c->sp_ymax = g->spawn_ymax;
    c->gx = g->gravity_x; c->gy = g->gravity_y; c->dt = g->dt; c->force_scale = g->force_scale;
    c->R = g->width / 2 - g->height / 2;                      // Drone.py:11
    c->d = g->height / 2 - 3;                                 // Drone.py:10
    c->hw = g->width / 2; c->hh = (g->height / 2) / 2;        // frame box (width, height/2), Drone.py:16
    c->minv_f = 1.0 / g->mass_f; c->minv_m = 1.0 / g->mass_m;
    c->iinv_f = 1.0 / moment_for_box(g->mass_f, g->width, g->height / 2);
    c->iinv_m = 1.0 / moment_for_box(g->mass_m, g->height, g->height);
    c->msum = c->minv_m + c->minv_f;
    c->bias_coef = -(1.0 - pow(0.0, g->dt)) / g->dt;          // error_bias = 0 (Drone.py:64)
    c->diag = sqrt(g->screen_x * g->screen_x + g->screen_y * g->screen_y);
    c->r_vel = 1.0 / 2660.0; c->r_W = 1.0 / g->screen_x; c->r_H = 1.0 / g->screen_y; c->r_diag = 1.0 / c->diag;
    c->r_omega = 1.0 / 11.7; c->r_pi = 1.0 / M_PI; c->r_DR = 1.0 / g->danger_range; c->r_PA = 1.0 / g->PA_band_edge;
    c->seed_lo = (uint32_t)seed; c->seed_hi = (uint32_t)(seed >> 32);
    c->env_id_base = env_id_base;
}

extern "C" {

int d2d_abi_version(void) { return D2D_ABI_VERSION; }
const char *d2d_last_error(void) { return g_err; }

int d2d_default_config(d2d_config *c)
{
    if (!c) return fail(D2D_EINVAL, "config is NULL");
    memset(c, 0, sizeof *c);
    c->abi_version = D2D_ABI_VERSION;
    c->mode = D2D_MODE_CURRICULUM; c->stage = 1;
    c->n_steps = 1100; c->n_wps = 12; c->spawn_corner_lo = 1; c->spawn_corner_hi = 4; c->random_path_spawn = 1;
    c->use_Lambda = 1; c->iterations = 10;
    c->screen_x = 1300; c->screen_y = 1300; c->path_segment_length = 100; c->lookahead = 220;
    c->danger_range = 150; c->danger_angle = 20; c->abs_inv_CA_min_rew = 1.0 / 8;
    c->PA_band_edge = 40; c->PA_scale = 2; c->PP_vel_scale = 0.08; c->PP_rew_max = 2.5; c->PP_rew_min = -1;
    c->rew_collision = -50; c->reach_end_radius = 20; c->rew_reach_end = 30;
    c->AA_angle = M_PI / 2; c->AA_band = M_PI / 4; c->rew_AA = -1;
    c->stage_spawn_chance = 0.6;
    c->spawn_xmin = 50; c->spawn_xmax = 250; c->spawn_ymin = 150; c->spawn_ymax = 1000;
    c->mass_f = 0.2; c->mass_m = 0.4; c->width = 100; c->height = 20;
    c->gravity_x = 0; c->gravity_y = -1000; c->dt = 1.0 / 60; c->force_scale = 1000;
    return D2D_OK;
}

int d2d_state_bytes(const d2d_config *cfg, int64_t n_envs, size_t *bytes)
{
    int rc = check_config(cfg);
    if (rc) return rc;
    if (n_envs < 1 || !bytes) return fail(D2D_EINVAL, "n_envs must be >= 1 and bytes non-NULL");
    D2dLayout L;
    d2d_make_layout(&L, n_envs, cfg->mode == D2D_MODE_CURRICULUM);
    *bytes = L.bytes;
    return D2D_OK;
}

int d2d_create(const d2d_config *cfg, int64_t n_envs, int64_t env_id_base, uint64_t seed, int device,
               void *state_dev, size_t state_bytes, d2d_env **out)
{
    int rc = check_config(cfg);
    if (rc) return rc;
    if (!out || !state_dev || n_envs < 1) return fail(D2D_EINVAL, "bad arguments to d2d_create");
    if (((uintptr_t)state_dev & 255) != 0) return fail(D2D_EINVAL, "state_dev must be 256-byte aligned");
    D2dLayout L;
    d2d_make_layout(&L, n_envs, cfg->mode == D2D_MODE_CURRICULUM);
    if (state_bytes < L.bytes) return fail(D2D_EINVAL, "state buffer too small: %zu < %zu", state_bytes, L.bytes);
    DeviceGuard guard(device);
    d2d_env *e = new (std::nothrow) d2d_env();
    if (!e) return fail(D2D_EINVAL, "out of host memory");
    memset(e, 0, sizeof *e);
    e->cfg = *cfg; e->lay = L; e->device = device;
    char *base = (char *)state_dev;
    D2dParams &p = e->p;
    p.n = L.n; p.np = L.np;
    p.pos = (double2 *)(base + L.off_pos); p.vel = (double2 *)(base + L.off_vel); p.ang = (double2 *)(base + L.off_ang);
    p.jacc = (double2 *)(base + L.off_jacc); p.acc = (double2 *)(base + L.off_acc); p.meta = (int4 *)(base + L.off_meta);
    p.last = (double2 *)(base + L.off_last);
    p.stg_xy = (double2 *)(base + L.off_stg_xy); p.stg_an = (double2 *)(base + L.off_stg_an);
    p.stg_meta = (int4 *)(base + L.off_stg_meta); p.stg_obs = (float *)(base + L.off_stg_obs);
    p.queue = (int *)(base + L.off_queue); p.qcount = (int *)(base + L.off_qcount);
    // refill capacity per stage: one lane per 16 envs (a larger backlog is handled by striding over
    // the queue).  Tasks are dealt round-robin over the refill WARPS so that each warp carries only
    // a few lanes: the sampling loops have data-dependent trip counts and a warp runs for the
    // maximum over its lanes.  When the main blocks alone fit in one wave, the refill blocks are
    // capped so that the whole grid still does (a second wave would double the step time).
    {
        int want = (int)((L.n + 16 * D2D_BLOCK - 1) / (16 * D2D_BLOCK));
        int occ = 0, sms = 0;
        cudaError_t oe = (cfg->mode == D2D_MODE_CURRICULUM)
                             ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, d2d_step_kernel<true>, D2D_BLOCK, 0)
                             : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, d2d_step_kernel<false>, D2D_BLOCK, 0);
        if (oe == cudaSuccess) oe = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
        if (oe == cudaSuccess && occ > 0 && sms > 0) {
            const int64_t slots = (int64_t)occ * sms, main_blocks = (L.n + D2D_BLOCK - 1) / D2D_BLOCK;
            if (slots > main_blocks + 2) want = (int)std::min<int64_t>(want, (slots - main_blocks) / 2);
        } else {
            cudaGetLastError();
        }
        p.refill_blocks = want < 1 ? 1 : want;
        p.observe_blocks = p.refill_blocks;
        // tuning overrides (tools/refill_split_probe.py)
        if (const char *sb = getenv("D2D_SAMPLE_BLOCKS")) { const int v = atoi(sb); if (v >= 1) p.refill_blocks = v; }
        if (const char *ob = getenv("D2D_OBSERVE_BLOCKS")) { const int v = atoi(ob); if (v >= 1) p.observe_blocks = v; }
        p.use_hints = 1;
        if (const char *uh = getenv("D2D_SEARCH_HINTS")) p.use_hints = atoi(uh) != 0;   // tests: hinted == unhinted, bit for bit
    }
    if (L.per_env_tables) {
        p.us = (double *)(base + L.off_us); p.wps = (double2 *)(base + L.off_wps);
        p.pa = (double2 *)(base + L.off_pa); p.pb = (double2 *)(base + L.off_pb); p.pc = (double2 *)(base + L.off_pc);
        p.hint = (double2 *)(base + L.off_hint);
        p.obxy = (double2 *)(base + L.off_obxy); p.obr = (double *)(base + L.off_obr);
        p.stg_off8 = (uint32_t)(L.stg_delta / 8); p.stg_off16 = (uint32_t)(L.stg_delta / 16);
    }
    derive_constants(cfg, &p.c, seed, env_id_base);
    cudaError_t err = cudaMalloc(&e->shared_dev, sizeof(D2dSharedTables));
    if (err == cudaSuccess) err = cudaMalloc(&e->stats_dev, (D2D_STATS_DOUBLES + 24) * sizeof(double));
    if (err == cudaSuccess) err = cudaMemset(e->stats_dev, 0, (D2D_STATS_DOUBLES + 24) * sizeof(double));
    if (err == cudaSuccess) err = cudaMemset(e->shared_dev, 0, sizeof(D2dSharedTables));
    if (err == cudaSuccess) err = cudaMemset(state_dev, 0, L.bytes);
    if (err != cudaSuccess) {
        cudaFree(e->shared_dev); cudaFree(e->stats_dev);
        delete e;
        return fail(D2D_ECUDA, "d2d_create: %s", cudaGetErrorString(err));
    }
    p.shared = e->shared_dev;
    p.stats = e->stats_dev;
    // Optional (D2D_L2_PERSIST=1): mark the hot prefix (dynamic state + live us/rows, touched once per
    // step) as L2-persisting so that it survives from one step launch to the next.  Measured on B200 at
    // 65 536 envs: 114.4 us/step with the window vs 112.3 us without -- off by default.
    {
        const char *env = getenv("D2D_L2_PERSIST");
        int max_persist = 0, max_window = 0;
        if ((env && env[0] == '1') &&
            cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, device) == cudaSuccess &&
            cudaDeviceGetAttribute(&max_window, cudaDevAttrMaxAccessPolicyWindowSize, device) == cudaSuccess &&
            max_persist > 0 && max_window > 0) {
            const size_t win = std::min<size_t>(L.hot_bytes, (size_t)max_window);
            const size_t persist = std::min<size_t>(win, (size_t)max_persist);
            if (cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, persist) == cudaSuccess) {
                e->l2_base = state_dev; e->l2_bytes = win;
                e->l2_hit_ratio = (float)std::min(1.0, (double)persist / (double)win);
            }
        }
        cudaGetLastError();
    }
    p.role_cycles = (unsigned long long *)(e->stats_dev + D2D_STATS_DOUBLES);
    e->tables_set = (cfg->mode == D2D_MODE_CURRICULUM);
    *out = e;
    return D2D_OK;
}

int d2d_destroy(d2d_env *e)
{
    if (!e) return D2D_OK;
    DeviceGuard guard(e->device);
    cudaFree(e->shared_dev); cudaFree(e->stats_dev);
    cudaFree(e->d_act); cudaFree(e->d_obs); cudaFree(e->d_rew); cudaFree(e->d_done);
    cudaFreeHost(e->h_act); cudaFreeHost(e->h_obs); cudaFreeHost(e->h_rew); cudaFreeHost(e->h_done);
    delete e;
    return D2D_OK;
}

int d2d_set_scenario(d2d_env *e, int n_wps, const double *wps, const double *us, const double *xparams,
                     const double *yparams, int n_obs, const double *obstacles, void *stream)
{
    if (!e) return fail(D2D_EINVAL, "handle is NULL");
    if (e->cfg.mode != D2D_MODE_TEST) return fail(D2D_ESTATE, "d2d_set_scenario needs mode = test");
    if (n_wps < 4 || n_wps > D2D_MAX_WPS) return fail(D2D_EINVAL, "n_wps must be in [4, %d]", D2D_MAX_WPS);
    if (n_obs < 0 || n_obs > D2D_MAX_OBS_SHARED) return fail(D2D_EINVAL, "n_obs must be in [0, %d]", D2D_MAX_OBS_SHARED);
    if (!wps || !us || !xparams || !yparams || (n_obs > 0 && !obstacles)) return fail(D2D_EINVAL, "NULL table pointer");
    for (int k = 1; k < n_wps; k++)
        if (!(us[k] > us[k - 1])) return fail(D2D_EINVAL, "us must be strictly increasing");
    DeviceGuard guard(e->device);
    D2dSharedTables t;
    memset(&t, 0, sizeof t);
    t.n_wps = n_wps; t.n_obs = n_obs;
    for (int k = 0; k < n_wps; k++) { t.us[k] = us[k]; t.wx[k] = wps[2 * k]; t.wy[k] = wps[2 * k + 1]; }
    for (int r = 0; r < n_wps - 2; r++) {
        t.ax[r] = xparams[3 * r]; t.bx[r] = xparams[3 * r + 1]; t.cx[r] = xparams[3 * r + 2];
        t.ay[r] = yparams[3 * r]; t.by[r] = yparams[3 * r + 1]; t.cy[r] = yparams[3 * r + 2];
    }
    for (int k = 0; k < n_obs; k++) { t.ox[k] = obstacles[3 * k]; t.oy[k] = obstacles[3 * k + 1]; t.orad[k] = obstacles[3 * k + 2]; }
    cudaStream_t st = (cudaStream_t)stream;
    CUDA_TRY(cudaMemcpyAsync(e->shared_dev, &t, sizeof t, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaStreamSynchronize(st));   // t lives on this stack frame
    e->p.c.n_wps = n_wps;
    e->tables_set = true;
    return D2D_OK;
}

static int grid_for(int64_t n, int block) { return (int)((n + block - 1) / block); }

static int launch_tables_kernel(d2d_env *e, double *wps, double *us, double *xp, double *yp, int32_t *n_obs,
                                double *obst, int to_host)
{
    const int64_t n = e->lay.n;
    const size_t b_w = n * D2D_MAX_WPS * 2 * 8, b_u = n * D2D_MAX_WPS * 8, b_p = n * D2D_MAX_WPS * 3 * 8;
    const size_t b_n = n * 4, b_o = n * D2D_MAX_OBS_ENV * 3 * 8;
    char *buf = nullptr;
    const size_t total = b_w + b_u + 2 * b_p + b_o + d2d_align_up(b_n, 8);
    CUDA_TRY(cudaDeviceSynchronize());          // steps may be in flight on a non-blocking stream
    CUDA_TRY(cudaMalloc(&buf, total));
    double *dw = (double *)buf, *du = (double *)(buf + b_w), *dx = (double *)(buf + b_w + b_u);
    double *dy = (double *)(buf + b_w + b_u + b_p), *dob = (double *)(buf + b_w + b_u + 2 * b_p);
    int *dn = (int *)(buf + b_w + b_u + 2 * b_p + b_o);
    cudaError_t err = cudaSuccess;
    if (!to_host) {
        err = cudaMemcpy(dw, wps, b_w, cudaMemcpyHostToDevice);
        if (!err) err = cudaMemcpy(du, us, b_u, cudaMemcpyHostToDevice);
        if (!err) err = cudaMemcpy(dx, xp, b_p, cudaMemcpyHostToDevice);
        if (!err) err = cudaMemcpy(dy, yp, b_p, cudaMemcpyHostToDevice);
        if (!err) err = cudaMemcpy(dob, obst, b_o, cudaMemcpyHostToDevice);
        if (!err) err = cudaMemcpy(dn, n_obs, b_n, cudaMemcpyHostToDevice);
    }
    if (!err) {
        d2d_pack_tables_kernel<<<grid_for(n, 128), 128>>>(e->p, dw, du, dx, dy, dn, dob, to_host);
        e->launches++;
        err = cudaGetLastError();
        if (!err) err = cudaDeviceSynchronize();
    }
    if (!err && to_host) {
        err = cudaMemcpy(wps, dw, b_w, cudaMemcpyDeviceToHost);
        if (!err) err = cudaMemcpy(us, du, b_u, cudaMemcpyDeviceToHost);
        if (!err) err = cudaMemcpy(xp, dx, b_p, cudaMemcpyDeviceToHost);
        if (!err) err = cudaMemcpy(yp, dy, b_p, cudaMemcpyDeviceToHost);
        if (!err) err = cudaMemcpy(obst, dob, b_o, cudaMemcpyDeviceToHost);
        if (!err) err = cudaMemcpy(n_obs, dn, b_n, cudaMemcpyDeviceToHost);
    }
    cudaFree(buf);
    if (err) return fail(D2D_ECUDA, "table transfer: %s", cudaGetErrorString(err));
    return D2D_OK;
}

int d2d_set_tables(d2d_env *e, const double *wps, const double *us, const double *xparams, const double *yparams,
                   const int32_t *n_obs, const double *obstacles)
{
    if (!e) return fail(D2D_EINVAL, "handle is NULL");
    if (e->cfg.mode != D2D_MODE_CURRICULUM) return fail(D2D_ESTATE, "per-env tables exist only in curriculum mode");
    if (!wps || !us || !xparams || !yparams || !n_obs || !obstacles) return fail(D2D_EINVAL, "NULL table pointer");
    for (int64_t i = 0; i < e->lay.n; i++) {
        if (n_obs[i] < 0 || n_obs[i] > D2D_MAX_OBS_ENV) return fail(D2D_EINVAL, "n_obs[%lld] out of range", (long long)i);
        for (int k = 1; k < e->cfg.n_wps; k++)
            if (!(us[i * D2D_MAX_WPS + k] > us[i * D2D_MAX_WPS + k - 1]))
                return fail(D2D_EINVAL, "us of env %lld must be strictly increasing", (long long)i);
    }
    DeviceGuard guard(e->device);
    return launch_tables_kernel(e, (double *)wps, (double *)us, (double *)xparams, (double *)yparams, (int32_t *)n_obs,
                                (double *)obstacles, 0);
}

int d2d_get_tables(d2d_env *e, double *wps, double *us, double *xparams, double *yparams, int32_t *n_obs, double *obstacles)
{
    if (!e) return fail(D2D_EINVAL, "handle is NULL");
    if (e->cfg.mode != D2D_MODE_CURRICULUM) return fail(D2D_ESTATE, "per-env tables exist only in curriculum mode");
    if (!wps || !us || !xparams || !yparams || !n_obs || !obstacles) return fail(D2D_EINVAL, "NULL table pointer");
    DeviceGuard guard(e->device);
    return launch_tables_kernel(e, wps, us, xparams, yparams, n_obs, obstacles, 1);
}

int d2d_set_taps(d2d_env *e, double *obs64_dev, double *info_dev)
{
    if (!e) return fail(D2D_EINVAL, "handle is NULL");
    e->p.obs64 = obs64_dev;
    e->p.info = info_dev;
    return D2D_OK;
}

int d2d_reset(d2d_env *e, const uint8_t *mask_dev, float *obs_dev, void *stream)
{
    if (!e) return fail(D2D_EINVAL, "handle is NULL");
    if (!e->tables_set) return fail(D2D_ESTATE, "test mode: call d2d_set_scenario before d2d_reset");
    DeviceGuard guard(e->device);
    D2dParams p = e->p;
    p.mask = mask_dev; p.obs = obs_dev;
    cudaStream_t st = (cudaStream_t)stream;
    const int grid = grid_for(p.n, D2D_BLOCK);
    if (e->cfg.mode == D2D_MODE_CURRICULUM) d2d_reset_kernel<true><<<grid, D2D_BLOCK, 0, st>>>(p);
    else d2d_reset_kernel<false><<<grid, D2D_BLOCK, 0, st>>>(p);
    e->launches++;
    CUDA_TRY(cudaGetLastError());
    return D2D_OK;
}

int d2d_step(d2d_env *e, const float *actions_dev, float *obs_dev, float *reward_dev, uint8_t *done_dev,
             float *terminal_obs_dev, int auto_reset, void *stream)
{
    if (!e) return fail(D2D_EINVAL, "handle is NULL");
    if (!e->tables_set) return fail(D2D_ESTATE, "test mode: call d2d_set_scenario before d2d_step");
    if (!obs_dev || !reward_dev || !done_dev) return fail(D2D_EINVAL, "obs/reward/done must be non-NULL");
    if (((uintptr_t)obs_dev & 15) != 0) return fail(D2D_EINVAL, "obs_dev must be 16-byte aligned (rows are stored as 128-bit words)");
    if (actions_dev && ((uintptr_t)actions_dev & 7) != 0) return fail(D2D_EINVAL, "actions_dev must be 8-byte aligned");
    DeviceGuard guard(e->device);
    D2dParams p = e->p;
    p.actions = actions_dev; p.obs = obs_dev; p.reward = reward_dev; p.done = done_dev;
    p.terminal_obs = terminal_obs_dev; p.auto_reset = auto_reset; p.step_index = e->step_index++;
    cudaStream_t st = (cudaStream_t)stream;
    const int grid = grid_for(p.n, D2D_BLOCK) + p.refill_blocks + p.observe_blocks;
    cudaLaunchConfig_t lc = {};
    lc.gridDim = dim3((unsigned)grid); lc.blockDim = dim3(D2D_BLOCK); lc.dynamicSmemBytes = 0; lc.stream = st;
    cudaLaunchAttribute attr[2];
    // Programmatic dependent launch: the blocks of this step may become resident while the previous kernel of
    // the stream (the previous step, or the policy kernel) drains; they touch global memory only after
    // griddepcontrol.wait.  The previous grid is fully resident or done by then, so nothing can starve.
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    lc.attrs = attr; lc.numAttrs = 1;
    if (e->l2_bytes) {
        attr[1].id = cudaLaunchAttributeAccessPolicyWindow;
        attr[1].val.accessPolicyWindow.base_ptr = e->l2_base;
        attr[1].val.accessPolicyWindow.num_bytes = e->l2_bytes;
        attr[1].val.accessPolicyWindow.hitRatio = e->l2_hit_ratio;
        attr[1].val.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
        attr[1].val.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
        lc.numAttrs = 2;
    }
    if (e->cfg.mode == D2D_MODE_CURRICULUM) CUDA_TRY(cudaLaunchKernelEx(&lc, d2d_step_kernel<true>, p));
    else CUDA_TRY(cudaLaunchKernelEx(&lc, d2d_step_kernel<false>, p));
    e->launches++;
    CUDA_TRY(cudaGetLastError());
    return D2D_OK;
}

static bool is_pinned_host(const void *ptr)
{
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, ptr) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeHost;
}

int d2d_step_host(d2d_env *e, const float *actions_host, float *obs_host, float *reward_host, uint8_t *done_host,
                  int auto_reset, void *stream)
{
    return d2d_step_host_ex(e, actions_host, obs_host, reward_host, done_host, nullptr, auto_reset, stream);
}

int d2d_step_host_ex(d2d_env *e, const float *actions_host, float *obs_host, float *reward_host, uint8_t *done_host,
                     float *terminal_obs_host, int auto_reset, void *stream)
{
    if (!e) return fail(D2D_EINVAL, "handle is NULL");
    if (!obs_host || !reward_host || !done_host) return fail(D2D_EINVAL, "obs/reward/done must be non-NULL");
    if (terminal_obs_host && !is_pinned_host(terminal_obs_host))
        return fail(D2D_EINVAL, "terminal_obs_host must be page-locked (only the rows of finished envs are written)");
    DeviceGuard guard(e->device);
    const int64_t n = e->lay.n;
    const size_t b_act = n * 2 * sizeof(float), b_obs = n * D2D_OBS_DIM * sizeof(float), b_rew = n * sizeof(float);
    cudaStream_t st = (cudaStream_t)stream;
    // Page-locked caller buffers (cudaHostAlloc / cudaHostRegister / torch pin_memory) are used
    // directly; pageable ones go through the handle's pinned staging buffers.
    const bool pin_a = actions_host && is_pinned_host(actions_host);
    const bool pin_o = is_pinned_host(obs_host), pin_r = is_pinned_host(reward_host), pin_d = is_pinned_host(done_host);
    static const bool force_copy = getenv("D2D_HOST_COPY") && getenv("D2D_HOST_COPY")[0] == '1';   // A/B switch (tools/e2e_ab.py)
    if (!force_copy && pin_o && pin_r && pin_d && (!actions_host || pin_a)) {
        // All caller buffers are page-locked: the kernel reads the actions from them and writes its
        // outputs into them directly (mapped host memory).  The PCIe writes of the blocks that finish
        // early overlap the rest of the launch, and three copies + their launch latencies disappear:
        // 2.3e8 -> 2.9e8 env-steps/s end to end at 65 536 envs.
        float *zo = nullptr, *zr = nullptr;
        uint8_t *zd = nullptr;
        float *za = nullptr, *zt = nullptr;
        bool mapped = cudaHostGetDevicePointer((void **)&zo, obs_host, 0) == cudaSuccess &&
                      (!terminal_obs_host || cudaHostGetDevicePointer((void **)&zt, terminal_obs_host, 0) == cudaSuccess) &&
                      cudaHostGetDevicePointer((void **)&zr, reward_host, 0) == cudaSuccess &&
                      cudaHostGetDevicePointer((void **)&zd, done_host, 0) == cudaSuccess &&
                      (!actions_host || cudaHostGetDevicePointer((void **)&za, (void *)actions_host, 0) == cudaSuccess);
        if (mapped) {
            int rc = d2d_step(e, za, zo, zr, zd, zt, auto_reset, stream);
            if (rc) return rc;
            CUDA_TRY(cudaStreamSynchronize(st));
            return D2D_OK;
        }
        (void)cudaGetLastError();                // not mapped (registered without the mapped flag): copy path below
    }
    // copy path: device + page-locked staging buffers of the handle, allocated on first use (all or nothing)
    if (!e->d_obs) {
        cudaError_t err = cudaMalloc(&e->d_act, b_act);
        if (!err) err = cudaMalloc(&e->d_rew, b_rew);
        if (!err) err = cudaMalloc(&e->d_done, n);
        if (!err) err = cudaHostAlloc(&e->h_act, b_act, cudaHostAllocDefault);
        if (!err) err = cudaHostAlloc(&e->h_obs, b_obs, cudaHostAllocDefault);
        if (!err) err = cudaHostAlloc(&e->h_rew, b_rew, cudaHostAllocDefault);
        if (!err) err = cudaHostAlloc(&e->h_done, n, cudaHostAllocDefault);
        if (!err) err = cudaMalloc(&e->d_obs, b_obs);            // last: d_obs != NULL marks the complete set
        if (err) {
            cudaFree(e->d_act); cudaFree(e->d_rew); cudaFree(e->d_done); cudaFree(e->d_obs);
            cudaFreeHost(e->h_act); cudaFreeHost(e->h_obs); cudaFreeHost(e->h_rew); cudaFreeHost(e->h_done);
            e->d_act = e->d_rew = e->d_obs = nullptr; e->d_done = nullptr;
            e->h_act = e->h_obs = e->h_rew = nullptr; e->h_done = nullptr;
            return fail(D2D_ECUDA, "d2d_step_host: staging buffers: %s", cudaGetErrorString(err));
        }
    }
    const float *dact = nullptr;
    if (actions_host) {
        const float *src = actions_host;
        if (!pin_a) { memcpy(e->h_act, actions_host, b_act); src = e->h_act; }
        CUDA_TRY(cudaMemcpyAsync(e->d_act, src, b_act, cudaMemcpyHostToDevice, st));
        dact = e->d_act;
    }
    float *dterm = nullptr;
    if (terminal_obs_host && cudaHostGetDevicePointer((void **)&dterm, terminal_obs_host, 0) != cudaSuccess)
        return fail(D2D_ECUDA, "terminal_obs_host is page-locked but not mapped into the device address space");
    int rc = d2d_step(e, dact, e->d_obs, e->d_rew, e->d_done, dterm, auto_reset, stream);
    if (rc) return rc;
    CUDA_TRY(cudaMemcpyAsync(pin_o ? obs_host : e->h_obs, e->d_obs, b_obs, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(pin_r ? reward_host : e->h_rew, e->d_rew, b_rew, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(pin_d ? done_host : e->h_done, e->d_done, n, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    if (!pin_o) memcpy(obs_host, e->h_obs, b_obs);
    if (!pin_r) memcpy(reward_host, e->h_rew, b_rew);
    if (!pin_d) memcpy(done_host, e->h_done, n);
    return D2D_OK;
}

static int state_transfer(d2d_env *e, double *host, int to_host)
{
    if (!e || !host) return fail(D2D_EINVAL, "NULL argument");
    DeviceGuard guard(e->device);
    const int64_t n = e->lay.n;
    const size_t bytes = n * D2D_STATE_DOUBLES * sizeof(double);
    double *blob = nullptr;
    CUDA_TRY(cudaDeviceSynchronize());          // steps may be in flight on a non-blocking stream
    CUDA_TRY(cudaMalloc(&blob, bytes));
    cudaError_t err = cudaSuccess;
    if (!to_host) err = cudaMemcpy(blob, host, bytes, cudaMemcpyHostToDevice);
    if (!err) {
        d2d_pack_state_kernel<<<grid_for(n, 128), 128>>>(e->p, blob, to_host);
        e->launches++;
        err = cudaGetLastError();
        if (!err) err = cudaDeviceSynchronize();
    }
    if (!err && to_host) err = cudaMemcpy(host, blob, bytes, cudaMemcpyDeviceToHost);
    cudaFree(blob);
    if (err) return fail(D2D_ECUDA, "state transfer: %s", cudaGetErrorString(err));
    return D2D_OK;
}

int d2d_get_state(d2d_env *e, double *state_host) { return state_transfer(e, state_host, 1); }
int d2d_set_state(d2d_env *e, const double *state_host) { return state_transfer(e, (double *)state_host, 0); }

int d2d_episode_stats(d2d_env *e, double *stats_host, int reset_after)
{
    if (!e || !stats_host) return fail(D2D_EINVAL, "NULL argument");
    DeviceGuard guard(e->device);
    CUDA_TRY(cudaDeviceSynchronize());
    CUDA_TRY(cudaMemcpy(stats_host, e->stats_dev, D2D_STATS_DOUBLES * sizeof(double), cudaMemcpyDeviceToHost));
    stats_host[14] = (double)e->step_index * (double)e->lay.n;
    if (reset_after) CUDA_TRY(cudaMemset(e->stats_dev, 0, D2D_STATS_DOUBLES * sizeof(double)));
    return D2D_OK;
}

int d2d_episode_stats_dev(d2d_env *e, double **stats_dev)
{
    if (!e || !stats_dev) return fail(D2D_EINVAL, "NULL argument");
    *stats_dev = e->stats_dev;
    return D2D_OK;
}

int d2d_debug_closest_u(d2d_env *e, const double *pos_dev, double *u_dev, int32_t *nfev_dev, void *stream)
{
    if (!e || !pos_dev || !u_dev || !nfev_dev) return fail(D2D_EINVAL, "NULL argument");
    if (!e->tables_set) return fail(D2D_ESTATE, "tables not set");
    DeviceGuard guard(e->device);
    cudaStream_t st = (cudaStream_t)stream;
    const int grid = grid_for(e->p.n, D2D_BLOCK);
    if (e->cfg.mode == D2D_MODE_CURRICULUM) d2d_closest_u_kernel<true><<<grid, D2D_BLOCK, 0, st>>>(e->p, pos_dev, u_dev, nfev_dev);
    else d2d_closest_u_kernel<false><<<grid, D2D_BLOCK, 0, st>>>(e->p, pos_dev, u_dev, nfev_dev);
    e->launches++;
    CUDA_TRY(cudaGetLastError());
    return D2D_OK;
}

int d2d_policy_weights_bytes(void) { return (int)sizeof(D2dPolicyImage); }

int d2d_policy_pack(const float *w0, const float *b0, const float *w1, const float *b1, const float *w2, const float *b2,
                    const float *log_std, void *packed_host)
{
    if (!w0 || !b0 || !w1 || !b1 || !w2 || !b2 || !log_std || !packed_host) return fail(D2D_EINVAL, "NULL argument");
    D2dPolicyImage *p = (D2dPolicyImage *)packed_host;
    memset(p, 0, sizeof *p);
    // x = hi + lo, hi = fp16(x) (round to nearest even), lo = fp16(x - hi): 22 significant bits together
    auto split = [](float v, __half &hi, __half &lo) {
        hi = __float2half_rn(v);
        lo = __float2half_rn(v - __half2float(hi));
    };
    for (int j = 0; j < D2D_POL_H; j++) {
        for (int k = 0; k < D2D_POL_K0; k++) {           // 27 weights, the bias (matched by x[27] = 1), zero padding
            const float v = k < 27 ? w0[j * 27 + k] : (k == 27 ? b0[j] : 0.0f);
            const int o = d2d_umma_idx(j, k, D2D_POL_H);
            split(v, p->w0_hi[o], p->w0_lo[o]);
        }
        for (int k = 0; k < D2D_POL_K1; k++) {
            const int o = d2d_umma_idx(j, k, D2D_POL_H);
            split(w1[j * D2D_POL_K1 + k], p->w1_hi[o], p->w1_lo[o]);
        }
        p->b1[j] = b1[j];
    }
    memcpy(p->w2, w2, sizeof p->w2);
    p->b2[0] = b2[0]; p->b2[1] = b2[1];
    p->std[0] = expf(log_std[0]); p->std[1] = expf(log_std[1]);
    p->log_std[0] = log_std[0]; p->log_std[1] = log_std[1];
    return D2D_OK;
}

static int launch_policy(d2d_env *e, const void *packed_weights_dev, const float *obs_dev, D2dPolicyOut out, int64_t n_rows,
                         int stochastic, uint64_t policy_step, void *stream)
{
    if (e->policy_grid_cap == 0) {
        // persistent CTAs, four per SM (shared memory: 49 KB each; TMEM: 128 of 512 columns each)
        CUDA_TRY(cudaFuncSetAttribute(d2d_policy_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(D2dPolicySmem)));
        CUDA_TRY(cudaFuncSetAttribute(d2d_policy_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        int sms = 0;
        CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, e->device));
        e->policy_grid_cap = D2D_POL_CTAS_PER_SM * sms;
    }
    const int64_t tiles = (n_rows + D2D_POL_TILE - 1) / D2D_POL_TILE;
    const int grid = (int)(tiles < e->policy_grid_cap ? tiles : e->policy_grid_cap);
    // programmatic dependent launch: the CTAs may become resident and set themselves up (barriers, TMEM, weight
    // image) while the preceding step kernel drains; they read `obs` only after griddepcontrol.wait
    cudaLaunchConfig_t lc = {};
    lc.gridDim = dim3((unsigned)grid); lc.blockDim = dim3(D2D_POL_THREADS);
    lc.dynamicSmemBytes = sizeof(D2dPolicySmem); lc.stream = (cudaStream_t)stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    lc.attrs = attr; lc.numAttrs = 1;
    CUDA_TRY(cudaLaunchKernelEx(&lc, d2d_policy_kernel, (const D2dPolicyImage *)packed_weights_dev, obs_dev, out, n_rows, stochastic,
                                e->p.c.seed_lo, e->p.c.seed_hi, e->p.c.env_id_base, (uint64_t)policy_step));
    e->launches++;
    CUDA_TRY(cudaGetLastError());
    return D2D_OK;
}

int d2d_policy_act(d2d_env *e, const void *packed_weights_dev, const float *obs_dev, float *actions_dev, int stochastic,
                   uint64_t policy_step, void *stream)
{
    return d2d_policy_act_ex(e, packed_weights_dev, obs_dev, actions_dev, nullptr, nullptr, stochastic, policy_step, stream);
}

int d2d_policy_act_ex(d2d_env *e, const void *packed_weights_dev, const float *obs_dev, float *actions_dev,
                      float *actions_raw_dev, float *log_prob_dev, int stochastic, uint64_t policy_step, void *stream)
{
    if (!e || !packed_weights_dev || !obs_dev || !actions_dev) return fail(D2D_EINVAL, "NULL argument");
    DeviceGuard guard(e->device);
    D2dPolicyOut out = {actions_dev, actions_raw_dev, log_prob_dev, nullptr};
    return launch_policy(e, packed_weights_dev, obs_dev, out, e->lay.n, stochastic, policy_step, stream);
}

int d2d_policy_value(d2d_env *e, const void *packed_value_weights_dev, const float *obs_dev, float *values_dev, int64_t n_rows,
                     void *stream)
{
    if (!e || !packed_value_weights_dev || !obs_dev || !values_dev) return fail(D2D_EINVAL, "NULL argument");
    if (n_rows < 1) return fail(D2D_EINVAL, "n_rows must be >= 1");
    DeviceGuard guard(e->device);
    D2dPolicyOut out = {nullptr, nullptr, nullptr, values_dev};
    return launch_policy(e, packed_value_weights_dev, obs_dev, out, n_rows, 0, 0, stream);
}

int d2d_rollout(d2d_env *e, const void *packed_weights_dev, float *obs_dev, float *actions_dev, float *reward_dev,
                uint8_t *done_dev, int n_steps, int stochastic, uint64_t policy_step0, void *stream)
{
    if (!e || !packed_weights_dev || !obs_dev || !actions_dev || !reward_dev || !done_dev) return fail(D2D_EINVAL, "NULL argument");
    if (n_steps < 1) return fail(D2D_EINVAL, "n_steps must be >= 1");
    for (int t = 0; t < n_steps; t++) {
        int rc = d2d_policy_act(e, packed_weights_dev, obs_dev, actions_dev, stochastic, policy_step0 + (uint64_t)t, stream);
        if (rc) return rc;
        rc = d2d_step(e, actions_dev, obs_dev, reward_dev, done_dev, nullptr, 1, stream);
        if (rc) return rc;
    }
    return D2D_OK;
}

int d2d_rollout_collect(d2d_env *e, const void *actor_weights_dev, const void *critic_weights_dev, int n_steps, float *obs_buf_dev,
                        float *actions_dev, float *log_probs_dev, float *rewards_dev, uint8_t *dones_dev, float *values_dev,
                        float *advantages_dev, float *returns_dev, float *action_scratch_dev, float gamma, float gae_lambda,
                        uint64_t policy_step0, void *stream)
{
    if (!e || !actor_weights_dev || !critic_weights_dev || !obs_buf_dev || !actions_dev || !log_probs_dev || !rewards_dev ||
        !dones_dev || !values_dev || !advantages_dev || !returns_dev || !action_scratch_dev)
        return fail(D2D_EINVAL, "NULL argument");
    if (n_steps < 1) return fail(D2D_EINVAL, "n_steps must be >= 1");
    const int64_t n = e->lay.n;
    const int T = n_steps;
    for (int t = 0; t < T; t++) {
        float *obs_t = obs_buf_dev + (int64_t)t * n * D2D_OBS_DIM;
        // SB3 collect_rollouts: sample from the current policy, store the UNclipped action and its log-probability,
        // step the env with the clipped one
        int rc = d2d_policy_act_ex(e, actor_weights_dev, obs_t, action_scratch_dev, actions_dev + (int64_t)t * n * 2,
                                   log_probs_dev + (int64_t)t * n, 1, policy_step0 + (uint64_t)t, stream);
        if (rc) return rc;
        // ... and the step kernel writes observation t+1, reward t and done t (= episode start t+1) straight into the rollout
        rc = d2d_step(e, action_scratch_dev, obs_t + n * D2D_OBS_DIM, rewards_dev + (int64_t)t * n, dones_dev + (int64_t)(t + 1) * n,
                      nullptr, 1, stream);
        if (rc) return rc;
    }
    // the critic over all T + 1 stored observations in ONE launch (the weights do not change during a rollout, so this
    // equals SB3's per-step values); row T gives the bootstrap values
    int rc = d2d_policy_value(e, critic_weights_dev, obs_buf_dev, values_dev, (int64_t)(T + 1) * n, stream);
    if (rc) return rc;
    return d2d_gae(rewards_dev, values_dev, dones_dev, values_dev + (int64_t)T * n, dones_dev + (int64_t)T * n, T, n, gamma,
                   gae_lambda, advantages_dev, returns_dev, stream);
}

int d2d_set_stage(d2d_env *e, int stage, double spawn_chance)
{
    if (!e) return fail(D2D_EINVAL, "handle is NULL");
    if (e->cfg.mode != D2D_MODE_CURRICULUM) return fail(D2D_ESTATE, "stages exist only in curriculum mode");
    if (stage < 1 || stage > 5) return fail(D2D_EINVAL, "stage must be 1..5");
    e->cfg.stage = stage; e->cfg.stage_spawn_chance = spawn_chance;
    e->p.c.stage = stage; e->p.c.spawn_chance = spawn_chance;
    return D2D_OK;
}

int d2d_gae(const float *rewards_dev, const float *values_dev, const uint8_t *episode_starts_dev,
            const float *last_values_dev, const uint8_t *last_dones_dev, int n_steps, int64_t n_envs, float gamma,
            float gae_lambda, float *advantages_dev, float *returns_dev, void *stream)
{
    if (!rewards_dev || !values_dev || !episode_starts_dev || !last_values_dev || !last_dones_dev || !advantages_dev ||
        !returns_dev)
        return fail(D2D_EINVAL, "NULL argument");
    if (n_steps < 1 || n_envs < 1) return fail(D2D_EINVAL, "n_steps and n_envs must be >= 1");
    d2d_gae_kernel<<<grid_for(n_envs, 256), 256, 0, (cudaStream_t)stream>>>(
        rewards_dev, values_dev, episode_starts_dev, last_values_dev, last_dones_dev, n_steps, n_envs, gamma, gae_lambda,
        advantages_dev, returns_dev);
    CUDA_TRY(cudaGetLastError());
    return D2D_OK;
}

int d2d_set_flight_path_ring(d2d_env *e, float *ring_dev, int ring_len)
{
    if (!e) return fail(D2D_EINVAL, "handle is NULL");
    if (ring_dev && ring_len < 1) return fail(D2D_EINVAL, "ring_len must be >= 1");
    if (ring_dev && ((uintptr_t)ring_dev & 7) != 0) return fail(D2D_EINVAL, "ring_dev must be 8-byte aligned");
    e->p.path_ring = (float2 *)ring_dev;
    e->p.path_ring_len = ring_dev ? ring_len : 0;
    return D2D_OK;
}

int d2d_set_info_done_only(d2d_env *e, int enable)
{
    if (!e) return fail(D2D_EINVAL, "handle is NULL");
    e->p.info_done_only = enable ? 1 : 0;
    return D2D_OK;
}

int d2d_debug_skip_physics(d2d_env *e, int enable)
{
    if (!e) return fail(D2D_EINVAL, "handle is NULL");
    e->p.skip_physics = enable ? 1 : 0;
    return D2D_OK;
}

int64_t d2d_launch_count(const d2d_env *e) { return e ? e->launches : 0; }

int d2d_debug_role_cycles(d2d_env *e, uint64_t *out24, int reset_after)
{
    if (!e || !out24) return fail(D2D_EINVAL, "NULL argument");
    DeviceGuard guard(e->device);
    e->p.probe = 1;                              // launches record their clock spans from the first call on
    CUDA_TRY(cudaDeviceSynchronize());
    CUDA_TRY(cudaMemcpy(out24, e->p.role_cycles, 24 * sizeof(uint64_t), cudaMemcpyDeviceToHost));
    if (reset_after) CUDA_TRY(cudaMemset(e->p.role_cycles, 0, 24 * sizeof(uint64_t)));
    return D2D_OK;
}

}  // extern "C"
