// EXPERIMENTAL (round-2 groundwork; NOT linked into libmogp_b200.so -- only into libmogp_b200_exp.so, which nothing but
// tools/gpu_diag.py i8 loads; not yet run on hardware): fp64 GEMM emulated on the
// int8 tensor pipe of sm_100a (tcgen05.mma kind::i8, int32 accumulators in TMEM) by Ozaki-style slicing.
//
// Why: the O(N^3) stages of the exact-GP step (trailing updates of the Cholesky, L^-1, K^-1 = L^-T L^-1) run on
// DMMA today (37 TFLOP/s measured; there is no fp64 tcgen05.mma).  The dense int8 peak of B200 is 4.5 POP/s; with
// S = 7..8 slices of 7 bits an fp64 product costs 28..36 int8 GEMMs whose int32 accumulation is exact, i.e. ~100
// TFLOP/s fp64-equivalent at full int8 rate.  tools/ozaki_proto.py (numpy, exact integers) shows S = 8, b = 7
// reproducing the fp64 GEMM's own rounding error (1e-15 of the largest entry) on K^-1 = L^-T L^-1 of GP covariances
// with cond(K) up to 1e6, and S = 7 staying below 1e-13 -- far inside the parity tolerances (LML rtol 1e-8).
//
//   A[i, :] = 2^ea[i] * sum_s As[s][i, :] 2^(-b (s+1)),  As[s] int8 signed digits,  likewise B[j, :] with eb[j]
//   C[i, j] = 2^(ea[i] + eb[j]) * sum_d 2^(-b (d+2)) * ( sum_{s+t=d} As[s] Bs[t]^T )[i, j],   d = 0 .. S-1
//
// The inner sums over one anti-diagonal d share a weight, so they accumulate in ONE int32 TMEM accumulator (exact
// while K * 2^(2b-2) * S < 2^31, i.e. K <= 65536 at b = 7, S = 8); the fp64 accumulation over d happens in registers.
//
// This first version is the simplest correct structure, not a fast one: one 128 x 64 tile per CTA, single-buffered
// cp.async operand staging in the no-swizzle K-major canonical layout, one thread issuing the MMAs, all four warps
// in the epilogue.  mogp_i8gemm_selftest compares it with the DMMA GEMM.  Next: TMA + multi-stage pipeline,
// 128 x 256 tiles / cta_group::2, triangular k-clipping, fused slicing of freshly written panels.
#include "common.cuh"
#include <cstdio>

#define I8_BM 128          // tile rows    (UMMA M)
#define I8_BN 64           // tile columns (UMMA N): 64 fp64 accumulators per thread in the epilogue
#define I8_BK 64           // bytes (= int8 elements) of K per stage: two UMMA instructions of K = 32
#define I8_MAX_S 10

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ------------------------------------------------------------------ slicing
// One CTA per row: ex[row] = exponent with |x| 2^-ex < 1/2, digits[s][row][k] = signed digits in [-64, 64] (b = 7).
// src element (row, k) at src[row * rs + k * cs]: the transposes the three GEMM forms need are absorbed here.
__global__ void __launch_bounds__(256) i8_slice_kernel(const double* __restrict__ src, long long rs, long long cs,
                                                       int K, int Kp, int S, int b, int8_t* __restrict__ digits,
                                                       long long slice_stride, int32_t* __restrict__ ex) {
    __shared__ double red[256];
    const int row = blockIdx.x, tid = threadIdx.x;
    const double* x = src + (long long)row * rs;
    double m = 0.0;
    for (int k = tid; k < K; k += 256) m = fmax(m, fabs(x[(long long)k * cs]));
    red[tid] = m;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (tid < o) red[tid] = fmax(red[tid], red[tid + o]);
        __syncthreads();
    }
    m = red[0];
    int e = 0;
    if (m > 0.0) { frexp(m, &e); e += 1; }           // m = f 2^e0 with f in [1/2, 1): |x| 2^-(e0+1) < 1/2
    if (tid == 0) ex[row] = e;
    const double sc = ldexp(1.0, -e), step = ldexp(1.0, b);
    for (int k = tid; k < Kp; k += 256) {
        double r = k < K ? x[(long long)k * cs] * sc : 0.0;
        for (int s = 0; s < S; ++s) {
            r *= step;
            const double d = rint(r);                 // |r| <= 2^(b-1) -> |d| <= 64
            digits[(long long)s * slice_stride + (long long)row * Kp + k] = (int8_t)(int)d;
            r -= d;                                   // exact
        }
    }
}

// ------------------------------------------------------------------ tcgen05 helpers
__device__ __forceinline__ void mbar_init_(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_wait_(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity)
        : "memory");
}
// K-major, no swizzle: core matrix = 8 rows x 16 bytes (128 contiguous bytes); SBO = stride between 8-row groups,
// LBO = stride between the two 16-byte K chunks of one K = 32 instruction (both in 16-byte units); version 1.
__device__ __forceinline__ uint64_t umma_desc_kmajor(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) | (1ull << 46);
}
__device__ __forceinline__ void umma_i8(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// ------------------------------------------------------------------ the tile kernel
// C[M x N] (fp64, ldc) = alpha * A B^T + beta * C with A = (As, ea) [M x Kp], B = (Bs, eb) [N x Kp] sliced as above.
// smem operand tiles: chunk-major, [K chunk of 16 bytes][row][16 bytes]  (see umma_desc_kmajor).
__global__ void __launch_bounds__(128, 1) i8_gemm_kernel(const int8_t* __restrict__ As, const int32_t* __restrict__ ea,
                                                         long long a_slice, const int8_t* __restrict__ Bs,
                                                         const int32_t* __restrict__ eb, long long b_slice, int M, int N,
                                                         int Kp, int S, int bbits, double alpha, double beta,
                                                         double* __restrict__ C, long long ldc) {
    __shared__ __align__(128) int8_t sA[I8_BK / 16][I8_BM][16];     // 8 KB
    __shared__ __align__(128) int8_t sB[I8_BK / 16][I8_BN][16];     // 4 KB
    __shared__ __align__(8) uint64_t mma_bar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    const int m0 = blockIdx.y * I8_BM, n0 = blockIdx.x * I8_BN;

    if (tid == 0) {
        mbar_init_(&mma_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {      // one warp allocates 64 TMEM columns (128 lanes x 64 x 32 bit accumulators)
        __syncwarp();
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "n"(I8_BN) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_d = tmem_base_s;

    // instruction descriptor: D = S32 (2 << 4), A = B = signed 8 bit (1 << 7, 1 << 10), both K-major, N >> 3, M >> 4
    const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(I8_BN >> 3) << 17) | ((uint32_t)(I8_BM >> 4) << 24);

    double acc[I8_BN];                       // this thread's row (lane = tid) of the tile, all 64 columns
#pragma unroll
    for (int j = 0; j < I8_BN; ++j) acc[j] = 0.0;

    uint32_t parity = 0;
    const int nkb = Kp / I8_BK;
    for (int d = 0; d < S; ++d) {
        bool first = true;
        for (int s = 0; s <= d; ++s) {
            const int t = d - s;
            const int8_t* Ag = As + (long long)s * a_slice + (long long)m0 * Kp;
            const int8_t* Bg = Bs + (long long)t * b_slice + (long long)n0 * Kp;
            for (int kb = 0; kb < nkb; ++kb) {
                // stage the operand tiles: 16-byte chunks, consecutive threads take consecutive chunks of one row
                for (int c = tid; c < I8_BM * (I8_BK / 16); c += 128) {
                    const int r = c / (I8_BK / 16), kc = c % (I8_BK / 16);
                    const bool ok = m0 + r < M;
                    const int8_t* src = Ag + (long long)(ok ? r : 0) * Kp + kb * I8_BK + kc * 16;
                    const int sz = ok ? 16 : 0;
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(&sA[kc][r][0])), "l"(src), "r"(sz));
                }
                for (int c = tid; c < I8_BN * (I8_BK / 16); c += 128) {
                    const int r = c / (I8_BK / 16), kc = c % (I8_BK / 16);
                    const bool ok = n0 + r < N;
                    const int8_t* src = Bg + (long long)(ok ? r : 0) * Kp + kb * I8_BK + kc * 16;
                    const int sz = ok ? 16 : 0;
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(&sB[kc][r][0])), "l"(src), "r"(sz));
                }
                asm volatile("cp.async.commit_group;" ::: "memory");
                asm volatile("cp.async.wait_group 0;" ::: "memory");
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic-proxy writes -> visible to the tensor core
                __syncthreads();
                if (tid == 0) {
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
                    for (int ki = 0; ki < I8_BK / 32; ++ki) {
                        const uint64_t da = umma_desc_kmajor(smem_u32(&sA[2 * ki][0][0]), I8_BM * 16, 128);
                        const uint64_t db = umma_desc_kmajor(smem_u32(&sB[2 * ki][0][0]), I8_BN * 16, 128);
                        umma_i8(tmem_d, da, db, idesc, (first && ki == 0) ? 0u : 1u);
                    }
                    umma_commit(&mma_bar);           // arrives when the MMAs above have read smem and written TMEM
                }
                first = false;
                mbar_wait_(&mma_bar, parity);        // single-buffered: the tiles are overwritten next
                parity ^= 1;
            }
        }
        // anti-diagonal d complete: acc += 2^(-b (d + 2)) * (int32 accumulators of this thread's lane)
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const double w = ldexp(1.0, -bbits * (d + 2));
#pragma unroll
        for (int c0 = 0; c0 < I8_BN; c0 += 16) {
            uint32_t v[16];
            const uint32_t taddr = tmem_d + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0;
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                  "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                : "r"(taddr)
                : "memory");
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int j = 0; j < 16; ++j) acc[c0 + j] = fma((double)(int32_t)v[j], w, acc[c0 + j]);
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();                              // every lane has read its accumulators before they are overwritten
    }

    const int row = m0 + tid;
    if (row < M) {
        const int er = ea[row];
#pragma unroll
        for (int j = 0; j < I8_BN; ++j) {
            const int col = n0 + j;
            if (col < N) {
                double v = alpha * ldexp(acc[j], er + eb[col]);
                double* p = C + (long long)row * ldc + col;
                if (beta != 0.0) v += beta * *p;
                *p = v;
            }
        }
    }
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "n"(I8_BN) : "memory");
}

// ------------------------------------------------------------------ second version: warp-specialised, pipelined
// Same arithmetic and shared-memory layout; 128 x 128 tiles, a ring of WS_STAGES operand stages filled by two producer
// warps (cp.async, completion signalled on an mbarrier with cp.async.mbarrier.arrive.noinc), one MMA-issuing thread,
// two TMEM accumulator buffers so that the fp64 epilogue of anti-diagonal d (eight warps, 64 columns of one row per
// thread) overlaps the int8 MMAs of d + 1.  Still no TMA / swizzle / cta_group::2: those come once this runs.
#define WS_BN 128
#define WS_STAGES 4
#define WS_PRODUCERS 64                      // warps 0, 1
#define WS_THREADS 384                       // + warp 2 (MMA, TMEM allocation), warp 3 idle, warps 4..11 epilogue
struct __align__(128) I8WsSmem {
    int8_t a[WS_STAGES][I8_BK / 16][I8_BM][16];      // 4 x 8 KB
    int8_t b[WS_STAGES][I8_BK / 16][WS_BN][16];      // 4 x 8 KB
    uint64_t full[WS_STAGES], empty[WS_STAGES];      // producers -> MMA, MMA (tcgen05.commit) -> producers
    uint64_t acc_full[2], acc_empty[2];              // MMA -> epilogue, epilogue -> MMA
    uint32_t tmem_base;
};
__device__ __forceinline__ void mbar_arrive_(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__global__ void __launch_bounds__(WS_THREADS, 1) i8_gemm_ws_kernel(const int8_t* __restrict__ As, const int32_t* __restrict__ ea,
                                                                    long long a_slice, const int8_t* __restrict__ Bs,
                                                                    const int32_t* __restrict__ eb, long long b_slice, int M,
                                                                    int N, int Kp, int S, int bbits, double alpha, double beta,
                                                                    double* __restrict__ C, long long ldc) {
    extern __shared__ __align__(128) unsigned char i8_smem_raw[];
    I8WsSmem& sm = *reinterpret_cast<I8WsSmem*>(i8_smem_raw);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int m0 = blockIdx.y * I8_BM, n0 = blockIdx.x * WS_BN;
    const int nkb = Kp / I8_BK;

    if (tid == 0) {
        for (int i = 0; i < WS_STAGES; ++i) { mbar_init_(&sm.full[i], WS_PRODUCERS); mbar_init_(&sm.empty[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init_(&sm.acc_full[i], 1); mbar_init_(&sm.acc_empty[i], 256); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {      // 2 x 128 accumulator columns
        __syncwarp();
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sm.tmem_base)), "n"(2 * WS_BN) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem0 = sm.tmem_base;

    if (warp < 2) {
        // ------------------------------------------------------------------ producers
        int stage = 0;
        uint32_t phase = 0;
        for (int d = 0; d < S; ++d)
            for (int s = 0; s <= d; ++s) {
                const int8_t* Ag = As + (long long)s * a_slice + (long long)m0 * Kp;
                const int8_t* Bg = Bs + (long long)(d - s) * b_slice + (long long)n0 * Kp;
                for (int kb = 0; kb < nkb; ++kb) {
                    mbar_wait_(&sm.empty[stage], phase ^ 1);         // slot free (passes immediately on the first lap)
                    // consecutive threads take consecutive rows of one 16-byte K chunk: conflict-free 128-byte core matrices
                    for (int c = tid; c < I8_BM * (I8_BK / 16); c += WS_PRODUCERS) {
                        const int kc = c / I8_BM, r = c % I8_BM;
                        const bool ok = m0 + r < M;
                        const int8_t* src = Ag + (long long)(ok ? r : 0) * Kp + kb * I8_BK + kc * 16;
                        const int sz = ok ? 16 : 0;
                        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(&sm.a[stage][kc][r][0])), "l"(src), "r"(sz));
                    }
                    for (int c = tid; c < WS_BN * (I8_BK / 16); c += WS_PRODUCERS) {
                        const int kc = c / WS_BN, r = c % WS_BN;
                        const bool ok = n0 + r < N;
                        const int8_t* src = Bg + (long long)(ok ? r : 0) * Kp + kb * I8_BK + kc * 16;
                        const int sz = ok ? 16 : 0;
                        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(&sm.b[stage][kc][r][0])), "l"(src), "r"(sz));
                    }
                    // arrives on full[stage] when this thread's copies above have landed
                    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(&sm.full[stage])) : "memory");
                    if (++stage == WS_STAGES) { stage = 0; phase ^= 1; }
                }
            }
    } else if (warp == 2) {
        // ------------------------------------------------------------------ MMA issuer (one thread)
        if (lane == 0) {
            const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(WS_BN >> 3) << 17) | ((uint32_t)(I8_BM >> 4) << 24);
            int stage = 0;
            uint32_t phase = 0, acc_phase[2] = {0, 0};
            for (int d = 0; d < S; ++d) {
                const int buf = d & 1;
                mbar_wait_(&sm.acc_empty[buf], acc_phase[buf] ^ 1);      // epilogue drained this buffer (immediate for d < 2)
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t tmem_d = tmem0 + (uint32_t)(buf * WS_BN);
                bool first = true;
                for (int s = 0; s <= d; ++s)
                    for (int kb = 0; kb < nkb; ++kb) {
                        mbar_wait_(&sm.full[stage], phase);
                        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // cp.async (generic proxy) writes -> tensor core
                        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
                        for (int ki = 0; ki < I8_BK / 32; ++ki) {
                            const uint64_t da = umma_desc_kmajor(smem_u32(&sm.a[stage][2 * ki][0][0]), I8_BM * 16, 128);
                            const uint64_t db = umma_desc_kmajor(smem_u32(&sm.b[stage][2 * ki][0][0]), WS_BN * 16, 128);
                            umma_i8(tmem_d, da, db, idesc, (first && ki == 0) ? 0u : 1u);
                        }
                        first = false;
                        umma_commit(&sm.empty[stage]);                    // frees the slot when these MMAs have read it
                        if (++stage == WS_STAGES) { stage = 0; phase ^= 1; }
                    }
                umma_commit(&sm.acc_full[buf]);                           // anti-diagonal d is complete in TMEM
                acc_phase[buf] ^= 1;
            }
        }
        __syncwarp();
    } else if (warp >= 4) {
        // ------------------------------------------------------------------ epilogue: row (lane quarter) x 64 columns per thread
        const int q = warp & 3, half = (warp - 4) >> 2;
        const int row = m0 + q * 32 + lane;
        double acc[64];
#pragma unroll
        for (int j = 0; j < 64; ++j) acc[j] = 0.0;
        uint32_t full_phase[2] = {0, 0};
        for (int d = 0; d < S; ++d) {
            const int buf = d & 1;
            mbar_wait_(&sm.acc_full[buf], full_phase[buf]);
            full_phase[buf] ^= 1;
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const double w = ldexp(1.0, -bbits * (d + 2));
#pragma unroll
            for (int c0 = 0; c0 < 64; c0 += 16) {
                uint32_t v[16];
                const uint32_t taddr = tmem0 + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * WS_BN + half * 64 + c0);
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                    : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                      "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                    : "r"(taddr)
                    : "memory");
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                for (int j = 0; j < 16; ++j) acc[c0 + j] = fma((double)(int32_t)v[j], w, acc[c0 + j]);
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            mbar_arrive_(&sm.acc_empty[buf]);                             // 256 epilogue threads release the buffer
        }
        if (row < M) {
            const int er = ea[row];
#pragma unroll
            for (int j = 0; j < 64; ++j) {
                const int col = n0 + half * 64 + j;
                if (col < N) {
                    double v = alpha * ldexp(acc[j], er + eb[col]);
                    double* p = C + (long long)row * ldc + col;
                    if (beta != 0.0) v += beta * *p;
                    *p = v;
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 2) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem0), "n"(2 * WS_BN) : "memory");
}

// ------------------------------------------------------------------ host side
// C (M x N, ldc) = alpha * A B^T + beta * C for fp64 A (M x K: element (i,k) at A[i*ars + k*acs]) and B (N x K likewise).
// work: at least S * (M + N) * Kp bytes + 4 * (M + N) bytes, Kp = K rounded up to 64.
static int g_i8_variant = 0;          // 0: simple single-buffered kernel, 1: warp-specialised pipelined kernel
extern "C" void mogp_i8gemm_set_variant(int v) { g_i8_variant = v; }

extern "C" int mogp_dgemm_i8(int M, int N, int K, double alpha, const double* A, long long ars, long long acs,
                             const double* B, long long brs, long long bcs, double beta, double* C, long long ldc, int S,
                             void* work, size_t work_bytes, void* stream) {
    if (S < 1 || S > I8_MAX_S || M < 1 || N < 1 || K < 1) return -1;
    const int b = 7;
    const int Kp = (K + I8_BK - 1) / I8_BK * I8_BK;
    if ((long long)Kp * (1ll << (2 * b - 2)) * S >= (1ll << 31)) return -1;          // int32 exactness
    const size_t need = (size_t)S * ((size_t)M + N) * Kp + 4 * ((size_t)M + N) + 256;
    if (work_bytes < need) return -1;
    cudaStream_t st = (cudaStream_t)stream;
    int8_t* As = (int8_t*)work;
    int8_t* Bs = As + (size_t)S * M * Kp;
    int32_t* ea = (int32_t*)(((uintptr_t)(Bs + (size_t)S * N * Kp) + 127) & ~(uintptr_t)127);
    int32_t* eb = ea + M;
    i8_slice_kernel<<<M, 256, 0, st>>>(A, ars, acs, K, Kp, S, b, As, (long long)M * Kp, ea);
    i8_slice_kernel<<<N, 256, 0, st>>>(B, brs, bcs, K, Kp, S, b, Bs, (long long)N * Kp, eb);
    if (g_i8_variant == 1) {
        static bool attr_done = false;
        if (!attr_done) {
            if (cudaFuncSetAttribute(i8_gemm_ws_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(I8WsSmem)) != cudaSuccess) return -2;
            attr_done = true;
        }
        dim3 grid((N + WS_BN - 1) / WS_BN, (M + I8_BM - 1) / I8_BM);
        i8_gemm_ws_kernel<<<grid, WS_THREADS, sizeof(I8WsSmem), st>>>(As, ea, (long long)M * Kp, Bs, eb, (long long)N * Kp, M, N, Kp, S, b,
                                                                     alpha, beta, C, ldc);
    } else {
        dim3 grid((N + I8_BN - 1) / I8_BN, (M + I8_BM - 1) / I8_BM);
        i8_gemm_kernel<<<grid, 128, 0, st>>>(As, ea, (long long)M * Kp, Bs, eb, (long long)N * Kp, M, N, Kp, S, b, alpha, beta, C, ldc);
    }
    MOGP_COUNT(3);
    return cudaGetLastError() == cudaSuccess ? 0 : -2;
}

// Self-test against the DMMA GEMM: random A (M x K), B (N x K); out[0] = max |C_i8 - C_dmma| / max |C_dmma|,
// out[1] = ms of the int8 path (slicing included), out[2] = ms of the DMMA GEMM.  M, N, K multiples of 64.
__global__ void i8_fill_kernel(double* x, long long n, unsigned seed) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    unsigned h = (unsigned)i * 2654435761u ^ seed;
    h ^= h >> 15; h *= 2246822519u; h ^= h >> 13; h *= 3266489917u; h ^= h >> 16;
    const double u = (double)h / 4294967296.0 - 0.5;
    x[i] = u * exp2((double)((int)(h % 13) - 6));        // a few binades of dynamic range within every row
}
__global__ void i8_maxdiff_kernel(const double* a, const double* b, long long n, double* out /*[2]: max diff, max ref*/) {
    __shared__ double sd[256], sm[256];
    double d = 0.0, m = 0.0;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        d = fmax(d, fabs(a[i] - b[i]));
        m = fmax(m, fabs(b[i]));
    }
    sd[threadIdx.x] = d; sm[threadIdx.x] = m;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) { sd[threadIdx.x] = fmax(sd[threadIdx.x], sd[threadIdx.x + o]); sm[threadIdx.x] = fmax(sm[threadIdx.x], sm[threadIdx.x + o]); }
        __syncthreads();
    }
    if (threadIdx.x == 0) {          // doubles >= 0: their bit patterns order like integers
        atomicMax(reinterpret_cast<unsigned long long*>(out), (unsigned long long)__double_as_longlong(sd[0]));
        atomicMax(reinterpret_cast<unsigned long long*>(out + 1), (unsigned long long)__double_as_longlong(sm[0]));
    }
}
extern "C" int mogp_i8gemm_selftest(int M, int N, int K, int S, double* out_host /*3*/) {
    if (M % 64 || N % 64 || K % 64) return -1;
    double *A = nullptr, *B = nullptr, *C1 = nullptr, *C2 = nullptr, *res = nullptr;
    void* work = nullptr;
    const size_t wb = (size_t)S * ((size_t)M + N) * K + 4 * ((size_t)M + N) + 4096;
    if (cudaMalloc(&A, (size_t)M * K * 8) || cudaMalloc(&B, (size_t)N * K * 8) || cudaMalloc(&C1, (size_t)M * N * 8) ||
        cudaMalloc(&C2, (size_t)M * N * 8) || cudaMalloc(&res, 16) || cudaMalloc(&work, wb))
        return -2;
    i8_fill_kernel<<<(unsigned)(((long long)M * K + 255) / 256), 256>>>(A, (long long)M * K, 17u);
    i8_fill_kernel<<<(unsigned)(((long long)N * K + 255) / 256), 256>>>(B, (long long)N * K, 91u);
    cudaMemset(res, 0, 16);
    cudaEvent_t e0, e1, e2;
    cudaEventCreate(&e0); cudaEventCreate(&e1); cudaEventCreate(&e2);
    int rc = 0;
    for (int rep = 0; rep < 2 && rc == 0; ++rep) {
        cudaEventRecord(e0);
        rc = mogp_dgemm_i8(M, N, K, 1.0, A, K, 1, B, K, 1, 0.0, C1, N, S, work, wb, nullptr);
        cudaEventRecord(e1);
        GemmArgs g{};
        g.A = A; g.lda = K; g.B = B; g.ldb = K; g.C = C2; g.ldc = N;
        g.M = M; g.N = N; g.K = K; g.alpha = 1.0; g.beta = 0.0;
        if (rc == 0 && launch_gemm(0, 1, g, 1, nullptr) != cudaSuccess) rc = -2;
        cudaEventRecord(e2);
    }
    if (rc == 0) {
        i8_maxdiff_kernel<<<256, 256>>>(C1, C2, (long long)M * N, res);
        double h[2];
        if (cudaMemcpy(h, res, 16, cudaMemcpyDeviceToHost) != cudaSuccess) rc = -2;
        else {
            float t1 = 0.f, t2 = 0.f;
            cudaEventElapsedTime(&t1, e0, e1);
            cudaEventElapsedTime(&t2, e1, e2);
            out_host[0] = h[1] > 0.0 ? h[0] / h[1] : -1.0;
            out_host[1] = t1;
            out_host[2] = t2;
        }
    }
    if (cudaDeviceSynchronize() != cudaSuccess) rc = -2;
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaEventDestroy(e2);
    cudaFree(A); cudaFree(B); cudaFree(C1); cudaFree(C2); cudaFree(res); cudaFree(work);
    return rc;
}
