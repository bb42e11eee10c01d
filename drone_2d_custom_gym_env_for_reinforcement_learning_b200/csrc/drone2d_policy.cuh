// drone2d_policy.cuh -- the policy in the loop on the 5th-generation tensor cores.
//
// The shipped SB3 MlpPolicy actor (main.py:251,262 `PPO.load` / `model.predict`):
//   obs[27] -> tanh(W0 x + b0)[64] -> tanh(W1 h + b1)[64] -> W2 h + b2 = mean[2];
//   action = clip(mean + exp(log_std) * eps, -1, 1), eps ~ N(0,1) from the Philox stream keyed by
//   (seed, global env id, policy step), or eps = 0.
// This is the one GEMM-shaped piece of the path: [N envs x 27] x [27 x 64] and [N x 64] x [64 x 64].
//
// One CTA = one tile of 128 environments at a time (M = 128 = the TMEM lane count):
//   * the weight matrices sit in shared memory in the tcgen05 K-major / no-swizzle core-matrix layout
//     (8 rows x 16 bytes per core matrix), packed on the host and brought in by bulk copies of the TMA unit;
//   * layer k: one elected thread issues tcgen05.mma.kind::f16 (M128 N64 K16) over the K extent,
//     accumulating float32 in TMEM (64 columns per layer), tcgen05.commit -> mbarrier;
//   * two threads per environment then tcgen05.ld their 32 columns (16 at a time), apply tanh, and either write
//     the next layer's A operand back to shared memory or finish the 2-wide output layer + Gaussian noise + clip.
// float32 fidelity: every operand is split x = hi + lo with hi = fp16(x), lo = fp16(x - hi) (22 significant bits
// together, like the tf32 split used before) and three products hi*hi + hi*lo + lo*hi are accumulated in float32;
// the result agrees with a float32 FMA chain to ~1e-6 (tests: 2e-5 against torch).  The first layer's bias is folded
// into the GEMM (x[27] = 1, W0[:, 27] = b0).
// Footprint: half-precision operands need half the bytes of tf32 and half the MMAs (K = 16 per instruction); the
// hidden activations H overwrite the region that held W0 and X (both dead by then; W0 is re-fetched per tile, 8 KB),
// so a CTA needs 49 KB of shared memory and 128 TMEM columns: FOUR CTAs per SM, i.e. all 512 tiles of 65 536
// environments are resident at once and the launch lasts one tile's latency chain instead of two.
#pragma once
#include <cstdint>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#define D2D_POL_TILE 128          // environments per CTA tile = UMMA M
#define D2D_POL_H 64              // hidden width = UMMA N
#define D2D_POL_K0 32             // 27 inputs + bias + 4 zero columns (K must be a multiple of 16)
#define D2D_POL_K1 64
#define D2D_STREAM_POLICY 0x504F4C49u   // 'POLI'

// half index of element (r, k) of an [R rows][K] operand in the UMMA K-major no-swizzle layout:
// core matrix = 8 rows x 8 halfs (128 contiguous bytes), core matrices ordered row-group-fastest.
__host__ __device__ inline int d2d_umma_idx(int r, int k, int R) { return ((k >> 3) * (R >> 3) + (r >> 3)) * 64 + (r & 7) * 8 + (k & 7); }

// Host-built image, copied verbatim into shared memory.  The first-layer weights come LAST: together with the X
// operand behind them they form the region the hidden activations overwrite (see D2dPolicySmem).
struct D2dPolicyImage {
    __half w1_hi[D2D_POL_H * D2D_POL_K1], w1_lo[D2D_POL_H * D2D_POL_K1];
    float b1[D2D_POL_H];
    float w2[2 * D2D_POL_H];
    float b2[2];
    float std[2];
    float log_std[2];
    float pad_[26];               // the operand regions below start on a 128-byte (core-matrix) boundary
    __half w0_hi[D2D_POL_H * D2D_POL_K0], w0_lo[D2D_POL_H * D2D_POL_K0];
};

namespace pol {

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start address, leading (K-chunk)
// and stride (8-row group) byte offsets in 16-byte units, version 1, SWIZZLE_NONE.
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes)
{
    return (uint64_t)((smem_addr >> 4) & 0x3FFFu) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32) | (1ull << 46);
}

// instruction descriptor (cute::UMMA::InstrDescriptor): D = f32 (c_format 1), A = B = f16 (format 0), both K-major,
// N = 64, M = 128
__device__ __forceinline__ constexpr uint32_t umma_idesc()
{
    return (1u << 4) | (0u << 7) | (0u << 10) | ((uint32_t)(D2D_POL_H >> 3) << 17) | ((uint32_t)(D2D_POL_TILE >> 4) << 24);
}

__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t accumulate)
{
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t"
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

// 16 consecutive float32 columns of this thread's TMEM lane
__device__ __forceinline__ void tmem_ld16(uint32_t addr, float *v)
{
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(addr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int j = 0; j < 16; j++) v[j] = __uint_as_float(r[j]);
}

// x = hi + lo with hi = fp16(x), lo = fp16(x - hi): 8 consecutive K elements of one operand row = one 16-byte
// core-matrix row of each half
__device__ __forceinline__ void split8(const float *x, uint4 &hi, uint4 &lo)
{
    __half2 h[4], l[4];
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const __half a = __float2half_rn(x[2 * j]), b = __float2half_rn(x[2 * j + 1]);
        h[j] = __halves2half2(a, b);
        l[j] = __halves2half2(__float2half_rn(x[2 * j] - __half2float(a)), __float2half_rn(x[2 * j + 1] - __half2float(b)));
    }
    hi = *reinterpret_cast<uint4 *>(h);
    lo = *reinterpret_cast<uint4 *>(l);
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

// shared memory of one CTA: the weight image, the X operand (hi | lo) and 8 KB more.  The hidden activations
// H (hi | lo, 2 x 16 KB) overwrite [w0_hi .. spare]: W0 and X are dead once layer 0 has been accumulated; the last KB
// of the region carries the partial outputs of the upper column half between the last epilogue and the next tile.
static_assert(sizeof(D2dPolicyImage) % 16 == 0, "bulk copies move multiples of 16 bytes");
static_assert(offsetof(D2dPolicyImage, w0_hi) % 128 == 0, "operand regions start on a core-matrix boundary");

struct D2dPolicySmem {
    D2dPolicyImage w;
    __half x[2 * D2D_POL_TILE * D2D_POL_K0];      // [hi | lo], 2 x 8 KB
    unsigned char spare[8192];
    unsigned long long bar[4];
    uint32_t tmem_base;
};
static_assert(2 * D2D_POL_H * D2D_POL_K0 * 2 + 2 * D2D_POL_TILE * D2D_POL_K0 * 2 + 8192 == 2 * D2D_POL_TILE * D2D_POL_K1 * 2,
              "W0 + X + spare is exactly the H operand (hi | lo)");

#define D2D_POL_THREADS 256       // two threads per environment row: columns [0,32) and [32,64) of every layer
#define D2D_POL_CTAS_PER_SM 4

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

__global__ void __launch_bounds__(D2D_POL_THREADS, D2D_POL_CTAS_PER_SM)
d2d_policy_kernel(const D2dPolicyImage *__restrict__ image, const float *__restrict__ obs, D2dPolicyOut out, int64_t n, int stochastic,
                  uint32_t seed_lo, uint32_t seed_hi, int64_t env_id_base, uint64_t step_index)
{
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    D2dPolicySmem &S = *reinterpret_cast<D2dPolicySmem *>(smem_raw);
    const int tid = threadIdx.x, warp = tid >> 5, row = tid & (D2D_POL_TILE - 1), half = tid >> 7;
    const uint32_t bar0 = pol::smem_u32(&S.bar[0]), bar1 = pol::smem_u32(&S.bar[1]), barw = pol::smem_u32(&S.bar[2]);
    const uint32_t barx = pol::smem_u32(&S.bar[3]);
    const int64_t n_tiles = (n + D2D_POL_TILE - 1) / D2D_POL_TILE;
    constexpr uint32_t W0_BYTES = 2 * D2D_POL_H * D2D_POL_K0 * 2;

    asm volatile("griddepcontrol.launch_dependents;");           // the next step kernel may queue up behind us
    // ---- one-time setup: barriers, TMEM (128 columns: layer-0 and layer-1 accumulators), weights ----
    if (tid == 0) {
        pol::mbar_init(bar0, 1);
        pol::mbar_init(bar1, 1);
        pol::mbar_init(barw, 1);
        pol::mbar_init(barx, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        // the 25 KB weight image comes in as ONE bulk copy of the TMA unit, completing on barw
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
    {
        // The first tile's observations: 128 rows x 27 float32 are 13 824 contiguous bytes -- ONE bulk copy of the TMA unit
        // into the (still unused) X-lo | spare region instead of 16 scalar loads per thread that each touch 32 different
        // sectors; every thread then picks its 16 values out of shared memory (row stride 27 words: conflict-free).
        // A ragged last tile or a base pointer that is not 16-byte aligned takes the scalar path.
        const int64_t base0 = (int64_t)blockIdx.x * D2D_POL_TILE;
        constexpr uint32_t RAW_BYTES = D2D_POL_TILE * 27 * 4;
        static_assert(RAW_BYTES % 16 == 0 && RAW_BYTES <= D2D_POL_TILE * D2D_POL_K0 * 2 + 8192 - 1024, "raw tile fits behind X-hi");
        float *raw = reinterpret_cast<float *>(reinterpret_cast<unsigned char *>(S.x) + D2D_POL_TILE * D2D_POL_K0 * 2);
        const bool bulk = base0 + D2D_POL_TILE <= n && (reinterpret_cast<uintptr_t>(obs) & 15) == 0;   // CTA-uniform
        if (bulk) {
            if (tid == 0) {
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(barx), "r"(RAW_BYTES) : "memory");
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(pol::smem_u32(raw)),
                             "l"(obs + base0 * 27), "r"(RAW_BYTES), "r"(barx)
                             : "memory");
            }
            __syncthreads();                     // barx is initialised and armed before anybody polls it
            pol::mbar_wait(barx, 0);
#pragma unroll
            for (int j = 0; j < 16; j++) {
                const int k = half * 16 + j;
                xr[j] = (k < 27) ? raw[row * 27 + k] : ((k == 27) ? 1.0f : 0.0f);
            }
        } else if (base0 < n) {
            d2d_policy_fetch(obs, n, base0 + row, half, xr);
        }
    }
    pol::tc_fence_before();
    __syncthreads();                             // (bulk path: every thread holds its inputs before X-lo is written over the raw tile)
    pol::tc_fence_after();
    pol::mbar_wait(barw, 0);                     // the weight image has landed (every thread reads b1 / w2 / b2 / std from it)
    const uint32_t tmem = S.tmem_base;
    const uint32_t tmem_lane = tmem + ((uint32_t)((warp & 3) * 32) << 16);   // a warp reaches the 32 TMEM lanes of its quarter

    unsigned char *hreg = reinterpret_cast<unsigned char *>(S.w.w0_hi);      // H operand: hi at +0, lo at +16 KB
    unsigned char *xreg = reinterpret_cast<unsigned char *>(S.x);            // X operand: hi at +0, lo at +8 KB
    // written after layer 1 has consumed H, read before the next tile's H is produced; neither the W0 re-fetch (first
    // 8 KB of the region) nor the next X (the 16 KB behind it) reaches the region's last KB
    float2 *part = reinterpret_cast<float2 *>(hreg + 2 * D2D_POL_TILE * D2D_POL_K1 * 2) - D2D_POL_TILE;
    const uint32_t h_s = pol::smem_u32(hreg), x_s = pol::smem_u32(xreg);
    const uint32_t w0_hi_s = pol::smem_u32(S.w.w0_hi), w0_lo_s = pol::smem_u32(S.w.w0_lo);
    const uint32_t w1_hi_s = pol::smem_u32(S.w.w1_hi), w1_lo_s = pol::smem_u32(S.w.w1_lo);
    constexpr uint32_t LBO_A = (D2D_POL_TILE / 8) * 128, LBO_B = (D2D_POL_H / 8) * 128, SBO = 128;
    constexpr uint32_t X_LO_OFF = D2D_POL_TILE * D2D_POL_K0 * 2, H_LO_OFF = D2D_POL_TILE * D2D_POL_K1 * 2;   // bytes
    // byte offset of this thread's row inside a K chunk (8 halfs = 16 bytes per row, 16 row groups of 128 bytes per chunk)
    const uint32_t row_off = (uint32_t)(row >> 3) * 128 + (uint32_t)(row & 7) * 16;

    uint32_t phase = 0;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, phase ^= 1u) {
        const int64_t base = tile * D2D_POL_TILE;

        // ---- A operand of layer 0: this thread's 16 inputs = K chunks 2*half and 2*half + 1 of its row ----
#pragma unroll
        for (int c = 0; c < 2; c++) {
            uint4 hi, lo;
            pol::split8(&xr[8 * c], hi, lo);
            const uint32_t o = (uint32_t)(2 * half + c) * LBO_A + row_off;
            *reinterpret_cast<uint4 *>(xreg + o) = hi;
            *reinterpret_cast<uint4 *>(xreg + X_LO_OFF + o) = lo;
        }
        pol::async_proxy_fence();
        pol::tc_fence_before();
        __syncthreads();

        // ---- layer 0: D0[128 x 64] = X[128 x 32] * W0^T, 2 K-steps x 3 split products ----
        if (tid == 0) {
            pol::mbar_wait(barw, phase);         // the weight image (first tile) / the re-fetched W0 has landed
            pol::tc_fence_after();
#pragma unroll
            for (int ks = 0; ks < D2D_POL_K0 / 16; ks++) {
                const uint64_t ah = pol::umma_desc(x_s + ks * 2 * LBO_A, LBO_A, SBO);
                const uint64_t al = pol::umma_desc(x_s + X_LO_OFF + ks * 2 * LBO_A, LBO_A, SBO);
                const uint64_t bh = pol::umma_desc(w0_hi_s + ks * 2 * LBO_B, LBO_B, SBO);
                const uint64_t bl = pol::umma_desc(w0_lo_s + ks * 2 * LBO_B, LBO_B, SBO);
                pol::umma_f16(tmem, ah, bh, ks > 0);
                pol::umma_f16(tmem, ah, bl, 1);
                pol::umma_f16(tmem, al, bh, 1);
            }
            pol::umma_commit(bar0);
        }
        // the next tile's observations travel while the tensor core and the epilogues work
        const bool more = tile + gridDim.x < n_tiles;
        if (more) d2d_policy_fetch(obs, n, (tile + gridDim.x) * D2D_POL_TILE + row, half, xr);
        pol::mbar_wait(bar0, phase);
        pol::tc_fence_after();

        // ---- hidden layer 0 activations -> A operand of layer 1 (this thread's row, its 32 columns), over W0 and X ----
#pragma unroll
        for (int q = 0; q < 2; q++) {
            float v[16];
            pol::tmem_ld16(tmem_lane + half * 32 + q * 16, v);
#pragma unroll
            for (int j = 0; j < 16; j++) v[j] = pol::tanh_fast(v[j]);
#pragma unroll
            for (int c = 0; c < 2; c++) {
                uint4 hi, lo;
                pol::split8(&v[8 * c], hi, lo);
                const uint32_t o = (uint32_t)(4 * half + 2 * q + c) * LBO_A + row_off;
                *reinterpret_cast<uint4 *>(hreg + o) = hi;
                *reinterpret_cast<uint4 *>(hreg + H_LO_OFF + o) = lo;
            }
        }
        pol::async_proxy_fence();
        pol::tc_fence_before();
        __syncthreads();

        // ---- layer 1: D1[128 x 64] = H[128 x 64] * W1^T, 4 K-steps x 3 split products ----
        if (tid == 0) {
            pol::tc_fence_after();
#pragma unroll
            for (int ks = 0; ks < D2D_POL_K1 / 16; ks++) {
                const uint64_t ah = pol::umma_desc(h_s + ks * 2 * LBO_A, LBO_A, SBO);
                const uint64_t al = pol::umma_desc(h_s + H_LO_OFF + ks * 2 * LBO_A, LBO_A, SBO);
                const uint64_t bh = pol::umma_desc(w1_hi_s + ks * 2 * LBO_B, LBO_B, SBO);
                const uint64_t bl = pol::umma_desc(w1_lo_s + ks * 2 * LBO_B, LBO_B, SBO);
                pol::umma_f16(tmem + D2D_POL_H, ah, bh, ks > 0);
                pol::umma_f16(tmem + D2D_POL_H, ah, bl, 1);
                pol::umma_f16(tmem + D2D_POL_H, al, bh, 1);
            }
            pol::umma_commit(bar1);
        }
        pol::mbar_wait(bar1, phase);
        pol::tc_fence_after();
        if (tid == 0 && more) {
            // H is dead: bring W0 back for the next tile while the output layer runs (8 KB, completes on barw)
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(barw), "r"(W0_BYTES) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(w0_hi_s),
                         "l"(image->w0_hi), "r"(W0_BYTES), "r"(barw)
                         : "memory");
        }

        // ---- output layer over this thread's 32 columns; the two halves of a row meet in shared memory ----
        float mu0 = 0.0f, mu1 = 0.0f;
#pragma unroll
        for (int q = 0; q < 2; q++) {
            float v[16];
            pol::tmem_ld16(tmem_lane + D2D_POL_H + half * 32 + q * 16, v);
#pragma unroll
            for (int j = 0; j < 16; j++) {
                const int jj = half * 32 + q * 16 + j;
                const float h2 = pol::tanh_fast(v[j] + S.w.b1[jj]);
                mu0 = fmaf(S.w.w2[jj], h2, mu0);
                mu1 = fmaf(S.w.w2[D2D_POL_H + jj], h2, mu1);
            }
        }
        if (half) part[row] = make_float2(mu0, mu1);
        pol::tc_fence_before();
        __syncthreads();                         // partial sums visible; TMEM and the operand regions are free for the next tile
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
        // `part` is read above by threads that then run ahead into the next tile; nobody writes the region's last KB
        // before the next tile's second __syncthreads (H of the next tile), which these readers also pass first
    }

    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(128u) : "memory");
}
