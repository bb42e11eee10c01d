// drone2d_policy.cuh -- the policy in the loop on the 5th-generation tensor cores.
//
// The shipped SB3 MlpPolicy actor (main.py:251,262 `PPO.load` / `model.predict`):
//   obs[27] -> tanh(W0 x + b0)[64] -> tanh(W1 h + b1)[64] -> W2 h + b2 = mean[2];
//   action = clip(mean + exp(log_std) * eps, -1, 1), eps ~ N(0,1) from the Philox stream keyed by
//   (seed, global env id, policy step), or eps = 0.
// This is the one GEMM-shaped piece of the path: [N envs x 27] x [27 x 64] and [N x 64] x [64 x 64].
//
// One CTA = one tile of 128 environments (M = 128 = the TMEM lane count), persistent over tiles:
//   * both weight matrices live in shared memory for the CTA's lifetime, already in the tcgen05
//     K-major / no-swizzle core-matrix layout (8 rows x 16 bytes per core matrix), packed on the host;
//   * layer k: one elected thread issues tcgen05.mma.kind::tf32 (M128 N64 K8) over the K extent,
//     accumulating in TMEM (64 float32 columns per layer), tcgen05.commit -> mbarrier;
//   * each thread then owns one environment: tcgen05.ld of its TMEM lane (64 pre-activations),
//     tanh, and either the next layer's A operand written back to shared memory or the 2-wide
//     output layer + Gaussian noise + clip in registers.
// float32 fidelity: every operand is split x = hi + lo with hi exactly representable in tf32
// (round-to-nearest on the 13 dropped bits) and three products hi*hi + hi*lo + lo*hi are accumulated
// ("3xTF32"); the result agrees with a float32 FMA chain to ~1e-6 (tests: 2e-5 against torch).
// The first layer's bias is folded into the GEMM (x[27] = 1, W0[:, 27] = b0).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#define D2D_POL_TILE 128          // environments per CTA tile = UMMA M
#define D2D_POL_H 64              // hidden width = UMMA N
#define D2D_POL_K0 32             // 27 inputs + bias + 4 zero columns (K must be a multiple of 8)
#define D2D_POL_K1 64
#define D2D_STREAM_POLICY 0x504F4C49u   // 'POLI'

// float index of element (r, k) of an [R rows][K] operand in the UMMA K-major no-swizzle layout:
// core matrix = 8 rows x 4 floats (128 contiguous bytes), core matrices ordered row-group-fastest.
__host__ __device__ inline int d2d_umma_idx(int r, int k, int R) { return ((k >> 2) * (R >> 3) + (r >> 3)) * 32 + (r & 7) * 4 + (k & 3); }

// Host-built image, copied verbatim into shared memory.
struct D2dPolicyImage {
    float w0_hi[D2D_POL_H * D2D_POL_K0], w0_lo[D2D_POL_H * D2D_POL_K0];
    float w1_hi[D2D_POL_H * D2D_POL_K1], w1_lo[D2D_POL_H * D2D_POL_K1];
    float b1[D2D_POL_H];
    float w2[2 * D2D_POL_H];
    float b2[2];
    float std[2];
    float log_std[2];
    float pad_[2];
};

__host__ __device__ inline float d2d_tf32_hi(float x)
{
#ifdef __CUDA_ARCH__
    return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xffffe000u);
#else
    uint32_t b;
    memcpy(&b, &x, 4);
    b = (b + 0x1000u) & 0xffffe000u;
    float r;
    memcpy(&r, &b, 4);
    return r;
#endif
}

namespace pol {

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start address, leading (K-chunk)
// and stride (8-row group) byte offsets in 16-byte units, version 1, SWIZZLE_NONE.
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes)
{
    return (uint64_t)((smem_addr >> 4) & 0x3FFFu) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32) | (1ull << 46);
}

// instruction descriptor (cute::UMMA::InstrDescriptor): D = f32, A = B = tf32, both K-major, N = 64, M = 128
__device__ __forceinline__ constexpr uint32_t umma_idesc()
{
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(D2D_POL_H >> 3) << 17) | ((uint32_t)(D2D_POL_TILE >> 4) << 24);
}

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t accumulate)
{
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t"
        "}\n" ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(umma_idesc()), "r"(accumulate), "r"(0u)
        : "memory");
}

__device__ __forceinline__ void umma_commit(uint32_t bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}

// bounded wait: a descriptor mistake must end in a trap, not in a hung GPU
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    const long long t0 = clock64();
    uint32_t done = 0;
    while (!done) {
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t"
            "}\n"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
        if (!done && clock64() - t0 > 2000000000ll) __trap();
    }
}

__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void async_proxy_fence() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// 32 consecutive float32 columns of this thread's TMEM lane
__device__ __forceinline__ void tmem_ld32(uint32_t addr, float *v)
{
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(addr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int j = 0; j < 32; j++) v[j] = __uint_as_float(r[j]);
}

// tanh from the exponential: 1 - 2 / (exp(2|x|) + 1), ex2.approx + rcp.approx, |error| < 3e-7
__device__ __forceinline__ float tanh_fast(float x)
{
    const float ax = fminf(fabsf(x), 15.0f);
    const float e = exp2f(ax * 2.885390081777927f);          // exp(2|x|)
    const float t = 1.0f - __fdividef(2.0f, e + 1.0f);
    return copysignf(t, x);
}

}  // namespace pol

// shared memory of one CTA: weight image, the A operand region (X hi/lo of layer 0, later H hi/lo of
// layer 1: the two never live at the same time; its last KB also carries the partial outputs of the upper
// column half between the last epilogue and the next tile's first), two mbarriers and the TMEM
// base address.  115.5 KB -> two CTAs per SM.
static_assert(sizeof(D2dPolicyImage) % 16 == 0, "bulk copies move multiples of 16 bytes");

struct D2dPolicySmem {
    D2dPolicyImage w;
    float a[2 * D2D_POL_TILE * D2D_POL_K1];       // [hi | lo], each up to 128 x 64
    unsigned long long bar[3];
    uint32_t tmem_base;
};

#define D2D_POL_THREADS 256       // two threads per environment row: columns [0,32) and [32,64) of every layer

// observation elements [16*half, 16*half + 16) of one row (27 real inputs, then the bias column 1, then zeros)
__device__ __forceinline__ void d2d_policy_fetch(const float *__restrict__ obs, int64_t n, int64_t row, int half, float *xr)
{
#pragma unroll
    for (int j = 0; j < 16; j++) {
        const int k = half * 16 + j;
        xr[j] = (k < 27) ? ((row < n) ? __ldg(&obs[row * 27 + k]) : 0.0f) : ((k == 27) ? 1.0f : 0.0f);
    }
}

// What the last stage of the kernel writes per row:
//   actor mode (values == nullptr): actions = clip(mean + std * eps, -1, 1) (what the env is stepped with, SB3 clips for
//       the env only), actions_raw = mean + std * eps and log_prob = sum_d log N(raw_d; mean_d, std_d) if non-null
//       (SB3 `policy.forward`, what the PPO rollout buffer stores, main.py:190-208);
//   value mode (values != nullptr): values = output 0 of the head (the critic tower packed like an actor, row 1 = 0).
struct D2dPolicyOut {
    float *actions, *actions_raw, *log_prob, *values;
};

__global__ void __launch_bounds__(D2D_POL_THREADS, 2) d2d_policy_kernel(const D2dPolicyImage *__restrict__ image,
                                                                        const float *__restrict__ obs, D2dPolicyOut out,
                                                                        int64_t n, int stochastic, uint32_t seed_lo, uint32_t seed_hi,
                                                                        int64_t env_id_base, uint64_t step_index)
{
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    D2dPolicySmem &S = *reinterpret_cast<D2dPolicySmem *>(smem_raw);
    const int tid = threadIdx.x, warp = tid >> 5, row = tid & (D2D_POL_TILE - 1), half = tid >> 7;
    const uint32_t bar0 = pol::smem_u32(&S.bar[0]), bar1 = pol::smem_u32(&S.bar[1]), barw = pol::smem_u32(&S.bar[2]);
    const int64_t n_tiles = (n + D2D_POL_TILE - 1) / D2D_POL_TILE;

    asm volatile("griddepcontrol.launch_dependents;");           // the next step kernel may queue up behind us
    // ---- one-time setup: barriers, TMEM (128 columns: layer-0 and layer-1 accumulators), weights ----
    if (tid == 0) {
        pol::mbar_init(bar0, 1);
        pol::mbar_init(bar1, 1);
        pol::mbar_init(barw, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        // the 49.9 KB weight image comes in as ONE bulk copy of the TMA unit, completing on barw
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(barw), "r"((uint32_t)sizeof(D2dPolicyImage)) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(pol::smem_u32(&S.w)),
                     "l"(image), "r"((uint32_t)sizeof(D2dPolicyImage)), "r"(barw)
                     : "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(pol::smem_u32(&S.tmem_base)), "r"(128u)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    // Programmatic dependent launch: everything above (barriers, TMEM, the weight image) may run while the
    // step kernel that produces `obs` is still finishing; nothing below may.
    asm volatile("griddepcontrol.wait;" ::: "memory");
    float xr[16];
    d2d_policy_fetch(obs, n, (int64_t)blockIdx.x * D2D_POL_TILE + row, half, xr);
    pol::tc_fence_before();
    __syncthreads();
    pol::tc_fence_after();
    pol::mbar_wait(barw, 0);                     // the weight image has landed (async proxy -> async proxy: no proxy fence needed)
    const uint32_t tmem = S.tmem_base;
    const uint32_t tmem_lane = tmem + ((uint32_t)((warp & 3) * 32) << 16);   // a warp reaches the 32 TMEM lanes of its quarter

    float *a_hi = S.a, *a_lo = S.a + D2D_POL_TILE * D2D_POL_K1;
    // written after layer 1 has consumed H, read before the next tile's H is produced; X (the first 16 KB of each half) never reaches it
    float2 *part = reinterpret_cast<float2 *>(a_lo + D2D_POL_TILE * D2D_POL_K1) - D2D_POL_TILE;
    const uint32_t a_hi_s = pol::smem_u32(a_hi);
    const uint32_t w0_hi_s = pol::smem_u32(S.w.w0_hi), w0_lo_s = pol::smem_u32(S.w.w0_lo);
    const uint32_t w1_hi_s = pol::smem_u32(S.w.w1_hi), w1_lo_s = pol::smem_u32(S.w.w1_lo);
    constexpr uint32_t LBO_A = (D2D_POL_TILE / 8) * 128, LBO_B = (D2D_POL_H / 8) * 128, SBO = 128;
    constexpr uint32_t A_LO_OFF = D2D_POL_TILE * D2D_POL_K1 * 4;          // byte offset of the lo half

    uint32_t phase = 0;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, phase ^= 1u) {
        const int64_t base = tile * D2D_POL_TILE;

        // ---- A operand of layer 0: this thread's 16 inputs, split and laid out for the MMA ----
#pragma unroll
        for (int c = 0; c < 4; c++) {
            float4 h, l;
            h.x = d2d_tf32_hi(xr[4 * c]); h.y = d2d_tf32_hi(xr[4 * c + 1]); h.z = d2d_tf32_hi(xr[4 * c + 2]); h.w = d2d_tf32_hi(xr[4 * c + 3]);
            l.x = xr[4 * c] - h.x; l.y = xr[4 * c + 1] - h.y; l.z = xr[4 * c + 2] - h.z; l.w = xr[4 * c + 3] - h.w;
            const int o = d2d_umma_idx(row, half * 16 + 4 * c, D2D_POL_TILE);
            *reinterpret_cast<float4 *>(&a_hi[o]) = h;
            *reinterpret_cast<float4 *>(&a_lo[o]) = l;
        }
        pol::async_proxy_fence();
        pol::tc_fence_before();
        __syncthreads();

        // ---- layer 0: D0[128 x 64] = X[128 x 32] * W0^T, 4 K-steps x 3 split products ----
        if (tid == 0) {
            pol::tc_fence_after();
#pragma unroll
            for (int ks = 0; ks < D2D_POL_K0 / 8; ks++) {
                const uint64_t ah = pol::umma_desc(a_hi_s + ks * 2 * LBO_A, LBO_A, SBO);
                const uint64_t al = pol::umma_desc(a_hi_s + A_LO_OFF + ks * 2 * LBO_A, LBO_A, SBO);
                const uint64_t bh = pol::umma_desc(w0_hi_s + ks * 2 * LBO_B, LBO_B, SBO);
                const uint64_t bl = pol::umma_desc(w0_lo_s + ks * 2 * LBO_B, LBO_B, SBO);
                pol::umma_tf32(tmem, ah, bh, ks > 0);
                pol::umma_tf32(tmem, ah, bl, 1);
                pol::umma_tf32(tmem, al, bh, 1);
            }
            pol::umma_commit(bar0);
        }
        // the next tile's observations travel while the tensor core and the epilogues work
        if (tile + gridDim.x < n_tiles) d2d_policy_fetch(obs, n, (tile + gridDim.x) * D2D_POL_TILE + row, half, xr);
        pol::mbar_wait(bar0, phase);
        pol::tc_fence_after();

        // ---- hidden layer 0 activations -> A operand of layer 1 (this thread's row, its 32 columns) ----
        {
            float v[32];
            pol::tmem_ld32(tmem_lane + half * 32, v);
#pragma unroll
            for (int c = 0; c < 8; c++) {
                float4 h, l;
                const float t0 = pol::tanh_fast(v[4 * c]), t1 = pol::tanh_fast(v[4 * c + 1]);
                const float t2 = pol::tanh_fast(v[4 * c + 2]), t3 = pol::tanh_fast(v[4 * c + 3]);
                h.x = d2d_tf32_hi(t0); h.y = d2d_tf32_hi(t1); h.z = d2d_tf32_hi(t2); h.w = d2d_tf32_hi(t3);
                l.x = t0 - h.x; l.y = t1 - h.y; l.z = t2 - h.z; l.w = t3 - h.w;
                const int o = d2d_umma_idx(row, half * 32 + 4 * c, D2D_POL_TILE);
                *reinterpret_cast<float4 *>(&a_hi[o]) = h;
                *reinterpret_cast<float4 *>(&a_lo[o]) = l;
            }
        }
        pol::async_proxy_fence();
        pol::tc_fence_before();
        __syncthreads();

        // ---- layer 1: D1[128 x 64] = H[128 x 64] * W1^T, 8 K-steps x 3 split products ----
        if (tid == 0) {
            pol::tc_fence_after();
#pragma unroll
            for (int ks = 0; ks < D2D_POL_K1 / 8; ks++) {
                const uint64_t ah = pol::umma_desc(a_hi_s + ks * 2 * LBO_A, LBO_A, SBO);
                const uint64_t al = pol::umma_desc(a_hi_s + A_LO_OFF + ks * 2 * LBO_A, LBO_A, SBO);
                const uint64_t bh = pol::umma_desc(w1_hi_s + ks * 2 * LBO_B, LBO_B, SBO);
                const uint64_t bl = pol::umma_desc(w1_lo_s + ks * 2 * LBO_B, LBO_B, SBO);
                pol::umma_tf32(tmem + D2D_POL_H, ah, bh, ks > 0);
                pol::umma_tf32(tmem + D2D_POL_H, ah, bl, 1);
                pol::umma_tf32(tmem + D2D_POL_H, al, bh, 1);
            }
            pol::umma_commit(bar1);
        }
        pol::mbar_wait(bar1, phase);
        pol::tc_fence_after();

        // ---- output layer over this thread's 32 columns; the two halves of a row meet in shared memory ----
        float mu0 = 0.0f, mu1 = 0.0f;
        {
            float v[32];
            pol::tmem_ld32(tmem_lane + D2D_POL_H + half * 32, v);
#pragma unroll
            for (int j = 0; j < 32; j++) {
                const int jj = half * 32 + j;
                const float h2 = pol::tanh_fast(v[j] + S.w.b1[jj]);
                mu0 = fmaf(S.w.w2[jj], h2, mu0);
                mu1 = fmaf(S.w.w2[D2D_POL_H + jj], h2, mu1);
            }
        }
        if (half) part[row] = make_float2(mu0, mu1);
        pol::tc_fence_before();
        __syncthreads();                         // partial sums visible; TMEM and the A region are free for the next tile
        if (!half && base + row < n) {
            const float2 other = part[row];
            mu0 += other.x + S.w.b2[0];
            mu1 += other.y + S.w.b2[1];
            float e0 = 0.0f, e1 = 0.0f;
            if (stochastic && !out.values) {
                uint32_t o[4];
                philox4x32_10((uint32_t)(env_id_base + base + row), (uint32_t)step_index, (uint32_t)(step_index >> 32),
                              D2D_STREAM_POLICY, seed_lo, seed_hi, o);
                // Box-Muller in float32 (the reference's noise comes from torch's generator: only the law matters)
                const float u1 = ((float)(o[0] >> 8) + 1.0f) * (1.0f / 16777216.0f);       // (0, 1]
                const float u2 = (float)(o[1] >> 8) * (1.0f / 16777216.0f);                // [0, 1)
                const float rr = sqrtf(-2.0f * __logf(u1));
                float sn, cs;
                sincospif(2.0f * u2, &sn, &cs);
                e0 = rr * cs; e1 = rr * sn;
                mu0 = fmaf(S.w.std[0], e0, mu0);
                mu1 = fmaf(S.w.std[1], e1, mu1);
            }
            if (out.values) {
                out.values[base + row] = mu0;                   // value mode: the noise-free head output
            } else {
                if (out.actions_raw) reinterpret_cast<float2 *>(out.actions_raw)[base + row] = make_float2(mu0, mu1);
                if (out.log_prob)                               // -(eps^2)/2 - log(std) - log(2 pi)/2, summed over the two dims
                    out.log_prob[base + row] = -0.5f * (e0 * e0 + e1 * e1) - S.w.log_std[0] - S.w.log_std[1] - 1.8378770664093453f;
                reinterpret_cast<float2 *>(out.actions)[base + row] =
                    make_float2(fminf(fmaxf(mu0, -1.0f), 1.0f), fminf(fmaxf(mu1, -1.0f), 1.0f));
            }
        }
    }

    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(128u) : "memory");
}
