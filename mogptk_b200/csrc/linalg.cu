// Dense fp64 linear algebra for the exact-GP step on sm_100a:
//   * a cp.async-pipelined GEMM on the fp64 tensor pipe (mma.sync m8n8k4 -> SASS DMMA.8x8x4;
//     tcgen05.mma has no .kind::f64) with triangular k-clipping modes,
//   * a 64x64 single-CTA Cholesky leaf that also inverts its block,
//   * blocked right-looking Cholesky (two-level panels), level-batched triangular inverse,
//     K^-1 = L^-T L^-1 with the W = (K^-1 - a a^T)/2 epilogue, triangular mat-vecs.
// Replaces torch.linalg.cholesky / cholesky_solve / solve_triangular and their autograd
// backward at mogptk/gpr/model.py:246,452,470 (reference) -- see include/mogp_b200.h.
#include "common.cuh"
#include <algorithm>
#include <cstdio>
#include <cstdlib>

// ============================================================================ primitives
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc, bool pred) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    int sz = pred ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gsrc), "r"(sz));
}
__device__ __forceinline__ void cp_async8(void* smem_dst, const void* gsrc) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(s), "l"(gsrc));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

// ============================================================================ GEMM
// C[M x N] (+)= alpha * op(A) * op(B), CTA tile BM x BN x 16, WM x WN warps, warp tile
// (BM/WM) x (BN/WN) of m8n8k4 DMMA fragments, STAGES-deep cp.async pipeline.
// TA: A stored [k][m] (m contiguous) instead of [m][k];  TB: B stored [n][k] instead of [k][n].
// Shared-memory row pitches are == 4 (mod 16) doubles, which makes every fragment load
// (lane (g,t) reads [g][t] or [t][g]) bank-conflict free.
template <int BM, int BN, int WM, int WN, int STAGES, int MINB, bool TA, bool TB>
__global__ void __launch_bounds__(WM* WN * 32, MINB) gemm_f64_kernel(GemmArgs g) {
    constexpr int BK = 16;
    constexpr int NTH = WM * WN * 32;
    constexpr int WTM = BM / WM, WTN = BN / WN, MT = WTM / 8, NT = WTN / 8;
    constexpr int LDA_S = TA ? (BM + 4) : (BK + 4);
    constexpr int LDB_S = TB ? (BK + 4) : (BN + 4);
    constexpr int A_STAGE = TA ? BK * LDA_S : BM * LDA_S;
    constexpr int B_STAGE = TB ? BN * LDB_S : BK * LDB_S;
    extern __shared__ __align__(16) double smem[];
    double* As = smem;
    double* Bs = smem + STAGES * A_STAGE;

    // pair mode: the CTA computes two tiles whose clipped k-ranges are complementary (tile t and
    // tile n_tiles-1-t), so every CTA of a triangular-operand GEMM does the same amount of work.
    const int ntn = g.N / BN, ntm = (g.M + BM - 1) / BM;
    for (int rep = 0; rep < 2; ++rep) {
    int tn = blockIdx.x, tm = blockIdx.y;
    if (g.pair == 1) { if (rep == 1) { tn = ntn - 1 - tn; if (tn <= (int)blockIdx.x) break; } }
    else if (g.pair == 2) { if (rep == 1) { tm = ntm - 1 - tm; if (tm <= (int)blockIdx.y) break; } }
    else if (g.pair == 3) {
        // lower-triangular output whose k-range starts at the row tile (K^-1 = Linv^T Linv): row tm has
        // tm+1 tiles of depth K - tm*BM.  Row tm is paired with row ntm-1-tm (same column), which makes the
        // heaviest CTAs K + BM deep instead of leaving single K-deep tiles as the critical path.
        const int mirror = ntm - 1 - tm;
        if (tm > mirror) return;
        if (tm == mirror) { if (rep == 1) break; if (tn > tm) return; }
        else if (tn <= tm) { if (rep == 1) tm = mirror; }
        else { if (rep == 1) break; if (tn > mirror) return; tm = mirror; }
    }
    else if (rep == 1) break;
    if (rep == 1) __syncthreads();
    if (g.lower && (tm + 1) * BM <= tn * BN) return;
    int klo = 0, khi = g.K;
    if (g.klo_mode == 1) klo = tn * BN;
    else if (g.klo_mode == 2) klo = tm * BM;
    if (g.khi_mode == 1) khi = min(g.K, (tm + 1) * BM);
    const long long bz = blockIdx.z;
    const double* __restrict__ A = g.A + bz * g.strideA;
    const double* __restrict__ B = g.B + bz * g.strideB;
    double* C = g.C + bz * g.strideC;
    const int m0 = tm * BM, n0 = tn * BN;
    const int tid = threadIdx.x;
    const int mvalid = g.M - m0;
    const long long lda = g.lda, ldb = g.ldb;

    auto load_stage = [&](int stage, int k0) {
        double* as = As + stage * A_STAGE;
        double* bs = Bs + stage * B_STAGE;
        if (!TA) {
            constexpr int CH = BM * 8;
#pragma unroll
            for (int c = tid; c < CH; c += NTH) {
                int row = c >> 3, cc = c & 7;
                bool ok = row < mvalid;
                const double* src = A + (long long)(m0 + (ok ? row : 0)) * lda + k0 + cc * 2;
                cp_async16(as + row * LDA_S + cc * 2, src, ok);
            }
        } else {
            constexpr int CPR = BM / 2;
            constexpr int CH = BK * CPR;
#pragma unroll
            for (int c = tid; c < CH; c += NTH) {
                int kr = c / CPR, cc = c % CPR;
                bool ok = cc * 2 < mvalid;
                const double* src = A + (long long)(k0 + kr) * lda + m0 + (ok ? cc * 2 : 0);
                cp_async16(as + kr * LDA_S + cc * 2, src, ok);
            }
        }
        if (TB) {
            constexpr int CH = BN * 8;
#pragma unroll
            for (int c = tid; c < CH; c += NTH) {
                int row = c >> 3, cc = c & 7;
                const double* src = B + (long long)(n0 + row) * ldb + k0 + cc * 2;
                cp_async16(bs + row * LDB_S + cc * 2, src, true);
            }
        } else {
            constexpr int CPR = BN / 2;
            constexpr int CH = BK * CPR;
#pragma unroll
            for (int c = tid; c < CH; c += NTH) {
                int kr = c / CPR, cc = c % CPR;
                const double* src = B + (long long)(k0 + kr) * ldb + n0 + cc * 2;
                cp_async16(bs + kr * LDB_S + cc * 2, src, true);
            }
        }
    };

    const int warp = tid >> 5, lane = tid & 31, gq = lane >> 2, tq = lane & 3;
    const int am0 = (warp / WN) * WTM, bn0 = (warp % WN) * WTN;

    double acc[MT][NT][2];
#pragma unroll
    for (int i = 0; i < MT; ++i)
#pragma unroll
        for (int j = 0; j < NT; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

    const int kt0 = klo / BK, kt1 = khi / BK;
#pragma unroll
    for (int s = 0; s < STAGES - 1; ++s) {
        if (kt0 + s < kt1) load_stage(s, (kt0 + s) * BK);
        cp_async_commit();
    }
    for (int kt = kt0; kt < kt1; ++kt) {
        cp_async_wait<STAGES - 2>();
        __syncthreads();
        {
            int nk = kt + STAGES - 1;
            if (nk < kt1) load_stage((nk - kt0) % STAGES, nk * BK);
            cp_async_commit();
        }
        const int stage = (kt - kt0) % STAGES;
        const double* as = As + stage * A_STAGE;
        const double* bs = Bs + stage * B_STAGE;
#pragma unroll
        for (int kk = 0; kk < BK; kk += 4) {
            double a[MT], b[NT];
#pragma unroll
            for (int i = 0; i < MT; ++i)
                a[i] = TA ? as[(kk + tq) * LDA_S + am0 + i * 8 + gq] : as[(am0 + i * 8 + gq) * LDA_S + kk + tq];
#pragma unroll
            for (int j = 0; j < NT; ++j)
                b[j] = TB ? bs[(bn0 + j * 8 + gq) * LDB_S + kk + tq] : bs[(kk + tq) * LDB_S + bn0 + j * 8 + gq];
#pragma unroll
            for (int i = 0; i < MT; ++i)
#pragma unroll
                for (int j = 0; j < NT; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
        }
    }
    cp_async_wait<0>();

    const long long ldc = g.ldc;
    const double beta = ((g.first_touch_row1 > 0 && m0 >= g.first_touch_row1 - 1) ||
                         (g.first_touch_col1 > 0 && n0 >= g.first_touch_col1 - 1)) ? 0.0 : g.beta;   // first touch of this tile
    // The old C values of a fragment row are loaded together BEFORE any of them is overwritten: C is not known to be free of
    // aliases, so a load placed after a store waits for its own round trip -- 16 dependent L2 round trips per thread with the
    // element-by-element form, which is what bounded the rank-64 updates (4 k-iterations of math per tile).
#pragma unroll
    for (int i = 0; i < MT; ++i) {
        const int row = am0 + i * 8 + gq;
        if (row >= mvalid) continue;
        const long long r = m0 + row;
        double2 old[NT];
        if (beta != 0.0) {
#pragma unroll
            for (int j = 0; j < NT; ++j)
                old[j] = *reinterpret_cast<const double2*>(C + r * ldc + n0 + bn0 + j * 8 + 2 * tq);
        }
#pragma unroll
        for (int j = 0; j < NT; ++j) {
            const int col = n0 + bn0 + j * 8 + 2 * tq;
            double2* p = reinterpret_cast<double2*>(C + r * ldc + col);
            double2 v;
            v.x = g.alpha * acc[i][j][0];
            v.y = g.alpha * acc[i][j][1];
            if (beta != 0.0) {
                v.x += beta * old[j].x;
                v.y += beta * old[j].y;
            }
            if (g.epi == 1) {
                const double ai = g.avec[r];
                v.x = 0.5 * (v.x - ai * g.avec[col]);
                v.y = 0.5 * (v.y - ai * g.avec[col + 1]);
            }
            *p = v;
        }
    }
    }   // pair loop
}

extern long long g_mogp_cfg_epoch;
// explicit per-launch priorities (GemmArgs::prio, panel steps): 1 = on
static int g_launch_prio = std::getenv("MOGP_LAUNCH_PRIO") ? std::atoi(std::getenv("MOGP_LAUNCH_PRIO")) : 1;
extern "C" int mogp_set_launch_prio(int on) { g_launch_prio = on; ++g_mogp_cfg_epoch; return 0; }
// Rank-64 updates (K = 64: the trailing updates of the single-level Cholesky sweep and the right-looking updates of the row-wise
// inverse).  With four k-iterations the pipelined kernel above is a chain of four or five dependent memory round trips per
// tile -- load, wait, load, wait, ..., then read C, then write -- and little math in between: 2048^2 x 64 runs at 13 TFLOP/s
// (cuBLAS: 12).  Here a tile issues EVERYTHING it will read at once -- the whole 64-deep A and B panels by cp.async and its
// old C values into registers -- waits once, multiplies, writes: one round trip per tile, four tiles per SM in flight.
template <int BM, int BN, bool TB>
__global__ void __launch_bounds__(128, 4) gemm_f64_k64_kernel(GemmArgs g) {
    constexpr int KD = 64, PA = KD + 4, PB = TB ? KD + 4 : BN + 4;
    constexpr int WTM = BM / 2, WTN = BN / 2, MT = WTM / 8, NT = WTN / 8;
    extern __shared__ __align__(16) double smem[];
    double* As = smem;                         // [BM][PA]
    double* Bs = smem + BM * PA;               // TB: [BN][PB] (n, k)   else [KD][PB] (k, n)
    const int tn = blockIdx.x, tm = blockIdx.y;
    if (g.lower && (tm + 1) * BM <= tn * BN) return;
    const long long bz = blockIdx.z;
    const double* __restrict__ A = g.A + bz * g.strideA;
    const double* __restrict__ B = g.B + bz * g.strideB;
    double* C = g.C + bz * g.strideC;
    const int m0 = tm * BM, n0 = tn * BN, tid = threadIdx.x;
#pragma unroll
    for (int c = tid; c < BM * (KD / 2); c += 128) {
        const int row = c / (KD / 2), cc = c % (KD / 2);
        cp_async16(As + row * PA + cc * 2, A + (long long)(m0 + row) * g.lda + cc * 2, true);
    }
    if (TB) {
#pragma unroll
        for (int c = tid; c < BN * (KD / 2); c += 128) {
            const int row = c / (KD / 2), cc = c % (KD / 2);
            cp_async16(Bs + row * PB + cc * 2, B + (long long)(n0 + row) * g.ldb + cc * 2, true);
        }
    } else {
#pragma unroll
        for (int c = tid; c < KD * (BN / 2); c += 128) {
            const int kr = c / (BN / 2), cc = c % (BN / 2);
            cp_async16(Bs + kr * PB + cc * 2, B + (long long)kr * g.ldb + n0 + cc * 2, true);
        }
    }
    cp_async_commit();
    const int warp = tid >> 5, lane = tid & 31, gq = lane >> 2, tq = lane & 3;
    const int am0 = (warp >> 1) * WTM, bn0 = (warp & 1) * WTN;
    const double beta = ((g.first_touch_row1 > 0 && m0 >= g.first_touch_row1 - 1) ||
                         (g.first_touch_col1 > 0 && n0 >= g.first_touch_col1 - 1)) ? 0.0 : g.beta;
    const long long ldc = g.ldc;
    double2 old[MT][NT];
    if (beta != 0.0) {
#pragma unroll
        for (int i = 0; i < MT; ++i)
#pragma unroll
            for (int j = 0; j < NT; ++j)
                old[i][j] = *reinterpret_cast<const double2*>(C + (long long)(m0 + am0 + i * 8 + gq) * ldc + n0 + bn0 + j * 8 + 2 * tq);
    }
    double acc[MT][NT][2];
#pragma unroll
    for (int i = 0; i < MT; ++i)
#pragma unroll
        for (int j = 0; j < NT; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
    cp_async_wait<0>();
    __syncthreads();
#pragma unroll 4
    for (int kk = 0; kk < KD; kk += 4) {
        double a[MT], b[NT];
#pragma unroll
        for (int i = 0; i < MT; ++i) a[i] = As[(am0 + i * 8 + gq) * PA + kk + tq];
#pragma unroll
        for (int j = 0; j < NT; ++j) b[j] = TB ? Bs[(bn0 + j * 8 + gq) * PB + kk + tq] : Bs[(kk + tq) * PB + bn0 + j * 8 + gq];
#pragma unroll
        for (int i = 0; i < MT; ++i)
#pragma unroll
            for (int j = 0; j < NT; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
    }
#pragma unroll
    for (int i = 0; i < MT; ++i)
#pragma unroll
        for (int j = 0; j < NT; ++j) {
            double2 v = make_double2(g.alpha * acc[i][j][0], g.alpha * acc[i][j][1]);
            if (beta != 0.0) { v.x += beta * old[i][j].x; v.y += beta * old[i][j].y; }
            *reinterpret_cast<double2*>(C + (long long)(m0 + am0 + i * 8 + gq) * ldc + n0 + bn0 + j * 8 + 2 * tq) = v;
        }
}
// Measured (profiles/r02_gemm_k64.txt): stand-alone 4096^2 x 64 21.1 vs 19.8 TFLOP/s, 8192^2 x 64 25.0 vs 23.2 (cuBLAS 18.0 / 19.9), but
// inside the step the four 52 KB tiles per SM crowd out the panel CTAs: cfg2 0.737 -> 0.768 ms, cfg4 2.58 -> 2.68 ms.  Off by default.
static int g_gemm_k64 = std::getenv("MOGP_GEMM_K64") ? std::atoi(std::getenv("MOGP_GEMM_K64")) : 0;
extern "C" int mogp_set_gemm_k64(int on) { g_gemm_k64 = on; ++g_mogp_cfg_epoch; return 0; }
template <int BM, int BN, bool TB>
static cudaError_t launch_gemm_k64(const GemmArgs& g, int batch, cudaStream_t s) {
    constexpr size_t SMEM = (size_t)(BM * 68 + (TB ? BN * 68 : 64 * (BN + 4))) * sizeof(double);
    auto kern = gemm_f64_k64_kernel<BM, BN, TB>;
    static PerDeviceOnce once;
    if (OnceGuard og{once}; og.needed()) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM);
        if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        if (e != cudaSuccess) return e;
    }
    if (g.M <= 0 || g.N <= 0 || batch <= 0) return cudaSuccess;
    dim3 grid(g.N / BN, g.M / BM, batch);
    if (g.prio > 0 && g_launch_prio) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = grid; cfg.blockDim = dim3(128);
        cfg.dynamicSmemBytes = SMEM; cfg.stream = s;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributePriority;
        at[0].val.priority = -(g.prio - 1);
        cfg.attrs = at; cfg.numAttrs = 1;
        MOGP_COUNT(1);
        return cudaLaunchKernelEx(&cfg, kern, g);
    }
    kern<<<grid, 128, SMEM, s>>>(g);
    MOGP_COUNT(1);
    return cudaGetLastError();
}

std::atomic<long long> g_mogp_launches{0};
long long g_mogp_cfg_epoch = 0;      // bumped by the tuning setters: captured step graphs are re-captured
// 0 (default): 64x64 tiles   1: force 128x128 tiles (256 threads)   3: force 128x64 tiles
static int g_gemm_cfg = -1;
static long long g_small_tile_threshold = 1400;
extern "C" void mogp_set_small_tile_threshold(long long t) { g_small_tile_threshold = t; ++g_mogp_cfg_epoch; }

extern "C" void mogp_set_gemm_config(int cfg) { g_gemm_cfg = cfg; ++g_mogp_cfg_epoch; }

template <int BM, int BN, int WM, int WN, int STAGES, int MINB, bool TA, bool TB>
static cudaError_t launch_gemm_cfg(const GemmArgs& g, int batch, cudaStream_t s) {
    constexpr int BK = 16;
    constexpr int LDA_S = TA ? (BM + 4) : (BK + 4);
    constexpr int LDB_S = TB ? (BK + 4) : (BN + 4);
    constexpr int A_STAGE = TA ? BK * LDA_S : BM * LDA_S;
    constexpr int B_STAGE = TB ? BN * LDB_S : BK * LDB_S;
    constexpr size_t SMEM = (size_t)STAGES * (A_STAGE + B_STAGE) * sizeof(double);
    auto kern = gemm_f64_kernel<BM, BN, WM, WN, STAGES, MINB, TA, TB>;
    static PerDeviceOnce once;
    if (OnceGuard og{once}; og.needed()) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM);
        if (e != cudaSuccess) return e;
        // same (maximal) shared-memory carve-out as the Cholesky panel kernel, so that both can be
        // resident on one SM while the look-ahead overlaps them
        e = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        if (e != cudaSuccess) return e;
    }
    if (g.M <= 0 || g.N <= 0 || batch <= 0) return cudaSuccess;
    dim3 grid(g.N / BN, (g.M + BM - 1) / BM, batch);
    if (g.pair == 1) grid.x = (grid.x + 1) / 2;
    if (g.pair == 2) grid.y = (grid.y + 1) / 2;
    if (g.prio > 0 && g_launch_prio) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = grid; cfg.blockDim = dim3(WM * WN * 32);
        cfg.dynamicSmemBytes = SMEM; cfg.stream = s;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributePriority;
        at[0].val.priority = -(g.prio - 1);
        cfg.attrs = at; cfg.numAttrs = 1;
        MOGP_COUNT(1);
        return cudaLaunchKernelEx(&cfg, kern, g);
    }
    kern<<<grid, WM * WN * 32, SMEM, s>>>(g);
    MOGP_COUNT(1);
    return cudaGetLastError();
}

template <bool TA, bool TB>
static cudaError_t launch_gemm_t(const GemmArgs& g, int batch, cudaStream_t s) {
    if (g_gemm_cfg < 0) {
        const char* e = getenv("MOGP_GEMM_CFG");
        g_gemm_cfg = e ? atoi(e) : 0;
    }
    if (g_gemm_cfg == 1 && (g.N % 128) == 0)
        return launch_gemm_cfg<128, 128, 2, 4, 3, 1, TA, TB>(g, batch, s);
    // Measured on B200 (profiles/r01_gemm_sweep.txt): 64x64 tiles at 3 CTAs/SM beat 128x64 (2 CTAs/SM) and
    // 128x128 (1 CTA/SM) at every size we use (e.g. 8192x8192x256: 31.9 vs 27.9 vs 25.9 TFLOP/s), so they
    // are the default; the other shapes stay selectable for experiments (mogp_set_gemm_config).
    if (g_gemm_cfg != 3) {
        // mid-size problems (about one wave of 64x64 tiles or less): 32x64 tiles give twice the CTAs with half
        // the depth each, which shortens the tail that dominates there
        long long tiles = (long long)(g.M / 64) * (g.N / 64) * batch;
        if (g.lower) tiles = tiles / 2 + 1;
        if (g_gemm_cfg == 4 || (g_gemm_cfg != 2 && tiles < g_small_tile_threshold))
            return launch_gemm_cfg<32, 64, 2, 2, 3, 4, TA, TB>(g, batch, s);
        return launch_gemm_cfg<64, 64, 2, 2, 3, 3, TA, TB>(g, batch, s);
    }
    return launch_gemm_cfg<128, 64, 2, 2, 3, 2, TA, TB>(g, batch, s);
}

cudaError_t launch_gemm(int transa, int transb, const GemmArgs& g, int batch, cudaStream_t s) {
    if ((g.N % 64) || (g.K % 16) || (g.M % 64)) return cudaErrorInvalidValue;
    if (g_gemm_k64 && g.K == 64 && transa == 0 && !g.klo_mode && !g.khi_mode && !g.pair && !g.epi)
        return transb ? launch_gemm_k64<32, 64, true>(g, batch, s) : launch_gemm_k64<32, 64, false>(g, batch, s);
    if (transa == 0 && transb == 0) return launch_gemm_t<false, false>(g, batch, s);
    if (transa == 0 && transb == 1) return launch_gemm_t<false, true>(g, batch, s);
    if (transa == 1 && transb == 0) return launch_gemm_t<true, false>(g, batch, s);
    return launch_gemm_t<true, true>(g, batch, s);
}

// 1/sqrt(d) for a positive, normal d without the library routine's special-case branches: hardware
// approximation (MUFU.RSQ64H) refined by two Newton steps (relative error ~1 ulp).  Valid for every
// normal positive double (0.5*d*y*y stays ~0.5); subnormal pivots flush to zero and give inf, which
// only happens for matrices no factorisation would survive.
__device__ __forceinline__ double rsqrt_pos(double d) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
    const double h = 0.5 * d;
    y = y * fma(-h * y, y, 1.5);
    y = y * fma(-h * y, y, 1.5);
    return y;
}

// ============================================================================ Cholesky panel step
// One launch per 64-column step.  Every CTA redundantly factors the 64x64 diagonal block in
// shared memory (8-column sub-panels; the 8x8 pivot block is factored in registers by every
// row-owning thread, so the only block-wide barriers are two per sub-panel) and carries its
// own 64 rows of the panel below through the same sweep (fused triangular solve by
// substitution: no explicit inverse on the critical path).  With has_prev the CTA first applies
// the previous panel's rank-64 update to its 128 x 64 tile on the tensor pipe (look-ahead: the
// rest of that update runs concurrently on a second stream, see potrf_padded).  CTA 0 parks
// L_kk in `Ltmp` (the diagonal block of a scratch matrix) because other CTAs may still be
// reading A_kk.
#define PS 65    // pitch of the row-major staging tile (conflict-free row-per-thread access)
#define PL 132   // pitch of the column-major copy of finished columns (rows contiguous; == 4 mod 16)
#define PZ 68    // pitch of the previous-panel operand tiles (conflict-free DMMA fragment loads)
__global__ void __launch_bounds__(256, 1) potrf_panel_kernel(double* __restrict__ A, long long lda,
                                                             double* __restrict__ Ltmp, long long ldt, int k0, int nrb,
                                                             int has_prev, int32_t* info, long long* dbg) {
    extern __shared__ __align__(16) double sm[];
    double* S = sm;                  // [128][PS] staging: rows 0..63 diagonal block, 64..127 this CTA's rows below
    double* Lc = sm + 128 * PS;      // [64][PL]  finished columns, column-major: Lc[col][row]  (16-byte aligned)
    double* ZZ = Lc;                 // [128][PZ] previous-panel values of the same 128 rows (aliases Lc, used first)
    double* Dsm = Lc + 128 * PZ;     // [8][8]    the updated pivot block of the current sub-panel
    double* Xr = Dsm + 64;           // [128][8]  fragment <-> row-per-thread exchange
    const int tid = threadIdx.x;
    const int b = blockIdx.x;
    const bool has_rows = b < nrb;
    // 8 warps feed the tensor pipe (a lone warp per SM sub-partition only reaches about half the DMMA issue
    // rate); threads 0..127 additionally own one row each for the scalar pivot / substitution phases.
    const int warp = tid >> 5, lane = tid & 31, gq = lane >> 2, tq = lane & 3;
    const int r0 = warp * 16;        // rows of this warp in the tensor-pipe phases
    if (dbg && tid == 64 && b == 0) dbg[0] = clock64();
    const double* Ad = A + (long long)k0 * lda + k0;
    double* Ar = A + (long long)(k0 + 64 + 64 * b) * lda + k0;
    // all loads in flight at once (8-byte cp.async: the padded pitch is not 16-byte aligned)
#pragma unroll
    for (int it = 0; it < 16; ++it) {
        const int idx = tid + it * 256;
        const int r = idx >> 6, c = idx & 63;
        cp_async8(S + r * PS + c, Ad + (long long)r * lda + c);
        if (has_rows) cp_async8(S + (64 + r) * PS + c, Ar + (long long)r * lda + c);
    }
    if (has_prev) {
#pragma unroll
        for (int it = 0; it < 8; ++it) {
            const int idx = tid + it * 256;              // 2048 16-byte chunks per 64 x 64 tile
            const int r = idx >> 5, cc = idx & 31;
            cp_async16(ZZ + r * PZ + cc * 2, Ad + (long long)r * lda - 64 + cc * 2, true);
            if (has_rows) cp_async16(ZZ + (64 + r) * PZ + cc * 2, Ar + (long long)r * lda - 64 + cc * 2, true);
        }
    }
    cp_async_commit();
    cp_async_wait<0>();
    __syncthreads();
    const int nrows = has_rows ? 128 : 64;
    if (has_prev) {
        // S[r][c] -= sum_k ZZ[r][k] * ZZ[c][k]   (rows of this CTA x the 64 rows of the diagonal block)
        if (r0 < nrows) {
            double acc2[2][8][2];
#pragma unroll
            for (int i = 0; i < 2; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    acc2[i][j][0] = S[(r0 + i * 8 + gq) * PS + j * 8 + 2 * tq];
                    acc2[i][j][1] = S[(r0 + i * 8 + gq) * PS + j * 8 + 2 * tq + 1];
                }
#pragma unroll 4
            for (int kk = 0; kk < 64; kk += 4) {
                double a[2], bb[8];
#pragma unroll
                for (int i = 0; i < 2; ++i) a[i] = -ZZ[(r0 + i * 8 + gq) * PZ + kk + tq];
#pragma unroll
                for (int j = 0; j < 8; ++j) bb[j] = ZZ[(j * 8 + gq) * PZ + kk + tq];
#pragma unroll
                for (int i = 0; i < 2; ++i)
#pragma unroll
                    for (int j = 0; j < 8; ++j) dmma884(acc2[i][j][0], acc2[i][j][1], a[i], bb[j]);
            }
#pragma unroll
            for (int i = 0; i < 2; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    S[(r0 + i * 8 + gq) * PS + j * 8 + 2 * tq] = acc2[i][j][0];
                    S[(r0 + i * 8 + gq) * PS + j * 8 + 2 * tq + 1] = acc2[i][j][1];
                }
        }
        __syncthreads();
    }
    const bool active_row = tid < nrows;          // nrows <= 128: threads 128..255 own no row
    if (dbg && tid == 64 && b == 0) dbg[1] = clock64();
    // Left-looking sweep over 8-column sub-panels.  Finished columns live in Lc (column-major).
#pragma unroll 1
    for (int p = 0; p < 8; ++p) {
        const int c0 = p * 8;
        const bool act = active_row && tid >= c0;
        // acc[row][0..8) = A[row][c0..c0+8) - sum_{cp<c0} L[row][cp] * L[c0+.][cp] for the warp's 16 rows, on the
        // tensor pipe (scalar FMAs with broadcast shared-memory operands were LSU-bound: ~70 cycles per column):
        // 2 row tiles x (c0/4) k-steps of DMMA.8x8x4 with A = -L rows and B = pivot rows, both from the
        // column-major copy Lc (pitch == 4 mod 16: conflict-free fragment loads).
        if (dbg && p == 4 && tid == 64 && b == 0) dbg[19] = clock64();
        if (r0 < nrows && r0 + 16 > c0) {
            double cf[2][2];
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                cf[i][0] = S[(r0 + 8 * i + gq) * PS + c0 + 2 * tq];
                cf[i][1] = S[(r0 + 8 * i + gq) * PS + c0 + 2 * tq + 1];
            }
#pragma unroll 4
            for (int k = 0; k < c0; k += 4) {
                const double bfrag = Lc[(k + tq) * PL + c0 + gq];
                const double a0 = -Lc[(k + tq) * PL + r0 + gq], a1 = -Lc[(k + tq) * PL + r0 + 8 + gq];
                dmma884(cf[0][0], cf[0][1], a0, bfrag);
                dmma884(cf[1][0], cf[1][1], a1, bfrag);
            }
            if (dbg && p == 4 && tid == 64 && b == 0) dbg[20] = clock64();
#pragma unroll
            for (int i = 0; i < 2; ++i)
                *reinterpret_cast<double2*>(Xr + (r0 + 8 * i + gq) * 8 + 2 * tq) = make_double2(cf[i][0], cf[i][1]);
        }
        if (dbg && p == 4 && tid == 64 && b == 0) dbg[21] = clock64();
        if (dbg && p == 4 && lane == 0 && b == 0) dbg[32 + warp] = clock64();          // arrival at the exchange barrier
        __syncthreads();
        if (dbg && p == 4 && tid == 64 && b == 0) dbg[22] = clock64();
        double acc[8];
        if (act) {
            const double2* q = reinterpret_cast<const double2*>(Xr + tid * 8);
            const double2 v0 = q[0], v1 = q[1], v2 = q[2], v3 = q[3];
            acc[0] = v0.x; acc[1] = v0.y; acc[2] = v1.x; acc[3] = v1.y;
            acc[4] = v2.x; acc[5] = v2.y; acc[6] = v3.x; acc[7] = v3.y;
            if (tid < c0 + 8) {
#pragma unroll
                for (int c = 0; c < 8; ++c) Dsm[(tid - c0) * 8 + c] = acc[c];
            }
        }
        if (dbg && p == 4 && lane == 0 && b == 0) dbg[40 + warp] = clock64();          // arrival at the pivot-block barrier
        __syncthreads();
        if (dbg && tid == 64 && b == 0) dbg[2 + 2 * p] = clock64();
        if (act) {
            double D[8][8], rinv[8];
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j <= i; ++j) D[i][j] = Dsm[i * 8 + j];
            int badcol = 8;                       // first non-positive pivot of this block (8 = none)
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                // branch-free pivot step (keeps the 8 columns in one basic block so that the compiler can
                // interleave the substitution of finished columns with the pivot chain)
                double d = D[c][c];
                const bool ok = d > 0.0;
                badcol = (!ok && badcol == 8) ? c : badcol;
                d = ok ? d : 1.0;
                const double ri = rsqrt_pos(d);
                rinv[c] = ri;
                D[c][c] = d * ri;
#pragma unroll
                for (int i = c + 1; i < 8; ++i) D[i][c] *= ri;
#pragma unroll
                for (int i = c + 1; i < 8; ++i)
#pragma unroll
                    for (int j = c + 1; j <= i; ++j) D[i][j] = fma(-D[i][c], D[j][c], D[i][j]);
            }
            if (badcol < 8 && tid == c0 && b == 0) atomicCAS(info, 0, k0 + c0 + badcol + 1);
            // Forward substitution of this row against the factored pivot block.  A row of the pivot block
            // itself goes through the same arithmetic (it reproduces its row of the factor bit for bit up to
            // the diagonal) and only masks the entries right of the diagonal: no divergent special case -- an
            // 8-way branch over the pivot rows used to cost their warp ~1000 cycles per sub-panel.
            const int jrow = tid - c0;                // >= 8 for the rows below the pivot block
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const double xc = acc[c] * rinv[c];
                acc[c] = (c <= jrow) ? xc : 0.0;
#pragma unroll
                for (int cc = c + 1; cc < 8; ++cc) acc[cc] = fma(-xc, D[cc][c], acc[cc]);
            }
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                Lc[(c0 + c) * PL + tid] = acc[c];
                S[tid * PS + c0 + c] = acc[c];
            }
        }
        if (dbg && p == 3 && lane == 0 && b == 0) dbg[48 + warp] = clock64();          // arrival at the end-of-sub-panel barrier
        __syncthreads();
        if (dbg && tid == 64 && b == 0) dbg[3 + 2 * p] = clock64();
    }
    if (has_rows)
        for (int idx = tid; idx < 4096; idx += 256) {
            const int r = idx >> 6, c = idx & 63;
            Ar[(long long)r * lda + c] = S[(64 + r) * PS + c];
        }
    if (b == 0) {
        double* Lt = Ltmp + (long long)k0 * ldt + k0;
        for (int idx = tid; idx < 4096; idx += 256) {
            const int r = idx >> 6, c = idx & 63;
            if (c <= r) Lt[(long long)r * ldt + c] = S[r * PS + c];
        }
    }
    if (dbg && tid == 64 && b == 0) dbg[18] = clock64();
}

// ---------------------------------------------------------------------------- warp-specialised panel step
// Same contract as potrf_panel_kernel, different schedule.  Measured on B200: a dependent DFMA chain slows from
// 8.5 to 40 cycles per op when a DMMA stream runs on the same SM sub-partition (they share the fp64 pipe), so
// tensor work cannot simply be overlapped with the pivot chain.  Here warp 0 alone (sub-partition 0) carries the
// scalar chain for all rows of the CTA (one row per lane and 32-row group: one redundant 8x8 pivot factorisation
// per lane instead of one per row), warps 1,2,3,5,6,7 (sub-partitions 1..3, two warps each) do every DMMA, and
// warp 4 -- which would share sub-partition 0 with the chain -- retires after the prologue.  For sub-panel p the
// tensor warps accumulate, while the chain is still busy with sub-panel p-1,
//     E = -[previous panel rows] [previous panel rows of the pivots]^T - L[:, <c0-8] L[piv, <c0-8]^T
// (the previous panel's rank-64 update is applied lazily, 8 columns at a time, instead of up front), then wait
// for the chain, add the last 8 finished columns (two k-steps), add the matrix entries (read straight from
// global memory, issued before the accumulation) and hand the ROWS x 8 block to the chain through Xr.  Finished
// values go from the chain's registers to global memory directly.  Named barriers: 1 = "Xr full" (tensor warps
// arrive, chain waits), 2 = "columns done" (chain arrives, tensor warps wait).
// OT = 8-row tiles of own rows per CTA: 8 (64 rows below the diagonal block per CTA) or 4 (32 rows: the tensor
// work per CTA -- half of which is the redundant update of the diagonal block -- drops from 16 to 12 row tiles,
// which brings it level with the chain; used while twice the CTAs still fit one wave).
#define XP 10    // pitch of the exchange tile (16-byte aligned rows, conflict-free 128-bit row reads)
#define WS_BAR_THREADS 224      // chain warp + six tensor warps
// Programmatic dependent launch: a panel step launched with the programmatic-serialization attribute may start while
// the previous step is still running (its launch latency, ~4 us measured, is what this hides); it must not touch
// anything the previous step produces before pdl_wait() returns.  Both are no-ops for ordinary launches.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
// Kernel spans as seen from inside (all CTAs), one slot per panel step: g_span[2 s] = earliest entry, g_span[2 s + 1] =
// latest warp exit (ns, global timer).  Off unless armed through mogp_panel_spans (a constant-bank flag).
__constant__ int c_span_on = 0;
__device__ unsigned long long g_span[2 * 136];
__device__ __forceinline__ void span_enter(int k0) {
    if (c_span_on && threadIdx.x == 0) atomicMin(&g_span[2 * (k0 >> 6)], global_ns());
}
__device__ __forceinline__ void span_exit(int k0) {
    if (c_span_on && (threadIdx.x & 31) == 0) atomicMax(&g_span[2 * (k0 >> 6) + 1], global_ns());
}
__device__ __forceinline__ void named_bar_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void named_bar_arrive(int id, int n) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(n) : "memory"); }

// In-register Cholesky of the 8x8 pivot block (lower triangle of D in, L out; rinv[c] = 1 / L[c][c]); badcol = first
// non-positive pivot (8 = none; such pivots are replaced by 1 so that everything stays finite).
// Two columns at a time with the closed form of the 2x2 leading block [[a, b], [b, c]]: l11 = sqrt(a), l21 = b / l11,
// l22 = sqrt(det / a) with det = a c - b^2, so rsqrt(a) and rsqrt(det) are independent and the sequential chain is 4
// instead of 8 reciprocal square roots per block.  (det by one FMA: its rounding error eps b^2 / det is no larger than that
// of the column-wise c - l21^2.)
__device__ __forceinline__ void factor_pivot8_pairs(double (&D)[8][8], double (&rinv)[8], int& badcol) {
#pragma unroll
    for (int c = 0; c < 8; c += 2) {
        double a = D[c][c];
        const double b = D[c + 1][c];
        double det = fma(a, D[c + 1][c + 1], -(b * b));
        const bool ok0 = a > 0.0, ok1 = det > 0.0;
        badcol = (!ok0 && badcol == 8) ? c : badcol;
        badcol = (ok0 && !ok1 && badcol == 8) ? c + 1 : badcol;
        a = ok0 ? a : 1.0;
        det = (ok0 && ok1) ? det : 1.0;
        const double r1 = rsqrt_pos(a), rd = rsqrt_pos(det);
        const double l11 = a * r1;
        const double r2 = l11 * rd;                   // 1 / l22 = sqrt(a) / sqrt(det)
        const double l21 = b * r1;
        rinv[c] = r1;
        rinv[c + 1] = r2;
        D[c][c] = l11;
        D[c + 1][c] = l21;
        D[c + 1][c + 1] = (det * rd) * r1;            // sqrt(det) / sqrt(a)
#pragma unroll
        for (int i = c + 2; i < 8; ++i) {
            D[i][c] *= r1;
            D[i][c + 1] = fma(-D[i][c], l21, D[i][c + 1]) * r2;
        }
#pragma unroll
        for (int i = c + 2; i < 8; ++i)
#pragma unroll
            for (int j = c + 2; j <= i; ++j) D[i][j] = fma(-D[i][c + 1], D[j][c + 1], fma(-D[i][c], D[j][c], D[i][j]));
    }
}

// One sub-panel of a tensor warp, with the set of live row tiles fixed at compile time (MASK bit i = slot i live).
// A DMMA under a run-time predicate gets a WARPSYNC in front of it, which serialises it behind the previous
// one (measured: ~65 cycles per DMMA and warp instead of 32, i.e. half the tensor rate); with the predicate
// resolved at compile time the DMMAs of a k-step issue back to back.
//   E = [A entries] - Z Zpiv^T (previous panel, if any) - L[:, 0:k_all) L[piv, 0:k_all)^T,
// columns [0, k_pre) of Lc are already published, columns [k_pre, k_all) only after barrier `wait_id` (skipped when
// wait_id < 0).  The 8-column block goes to Xp, then barrier `arrive_id` is signalled.
template <int NS, int MASK, int PLW>
__device__ __forceinline__ void ws_tensor_subpanel(const double* ZZ, const double* Lc, double* Xp, const int (&mt)[NS],
                                                   const double* const (&rowp)[NS], int c0, int has_prev, int k_pre,
                                                   int k_all, int wait_id, int arrive_id, int bar_threads, int gq, int tq,
                                                   long long* ts) {
    if (ts) ts[0] = clock64();
    double cf[NS][2];
    double2 a0[NS];
#pragma unroll
    for (int i = 0; i < NS; ++i) {
        cf[i][0] = 0.0; cf[i][1] = 0.0;
        a0[i] = make_double2(0.0, 0.0);
        if ((MASK >> i) & 1) a0[i] = *reinterpret_cast<const double2*>(rowp[i] + c0 + 2 * tq);
    }
    if (has_prev) {
#pragma unroll 4
        for (int kk = 0; kk < 64; kk += 4) {
            const double nb = -ZZ[(c0 + gq) * PZ + kk + tq];
#pragma unroll
            for (int i = 0; i < NS; ++i)
                if ((MASK >> i) & 1) dmma884(cf[i][0], cf[i][1], ZZ[(mt[i] * 8 + gq) * PZ + kk + tq], nb);
        }
    }
    if (ts) ts[1] = clock64() + (cf[0][0] == 1.2345e300 ? 1 : 0);      // (data dependence: after the accumulation)
#pragma unroll 2
    for (int k = 0; k < k_pre; k += 4) {
        const double nb = -Lc[(k + tq) * PLW + c0 + gq];
#pragma unroll
        for (int i = 0; i < NS; ++i)
            if ((MASK >> i) & 1) dmma884(cf[i][0], cf[i][1], Lc[(k + tq) * PLW + mt[i] * 8 + gq], nb);
    }
    if (ts) ts[2] = clock64() + (cf[0][0] == 1.2345e300 ? 1 : 0);
    if (wait_id >= 0) {
        named_bar_sync(wait_id, bar_threads);
        if (ts) ts[3] = clock64();
#pragma unroll 2
        for (int k = k_pre; k < k_all; k += 4) {
            const double nb = -Lc[(k + tq) * PLW + c0 + gq];
#pragma unroll
            for (int i = 0; i < NS; ++i)
                if ((MASK >> i) & 1) dmma884(cf[i][0], cf[i][1], Lc[(k + tq) * PLW + mt[i] * 8 + gq], nb);
        }
    }
#pragma unroll
    for (int i = 0; i < NS; ++i)
        if ((MASK >> i) & 1)
            *reinterpret_cast<double2*>(Xp + (mt[i] * 8 + gq) * XP + 2 * tq) =
                make_double2(cf[i][0] + a0[i].x, cf[i][1] + a0[i].y);
    if (ts) ts[4] = clock64() + (cf[0][0] == 1.2345e300 ? 1 : 0);
    __threadfence_block();
    named_bar_arrive(arrive_id, bar_threads);
    if (ts) ts[5] = clock64();
}
#define WS_TENSOR_CASE(M)                                                                                          \
    case M:                                                                                                        \
        ws_tensor_subpanel<NS, (M) & ((1 << NS) - 1), PLW>(ZZ, Lc, Xp, mt, rowp, c0, has_prev, k_pre, k_all, wait_id,  \
                                                           arrive_id, bar_threads, gq, tq, ts);                     \
        break;
template <int NS, int PLW>
__device__ __forceinline__ void ws_tensor_dispatch(int mask, const double* ZZ, const double* Lc, double* Xp,
                                                   const int (&mt)[NS], const double* const (&rowp)[NS], int c0,
                                                   int has_prev, int k_pre, int k_all, int wait_id, int arrive_id,
                                                   int bar_threads, int gq, int tq, long long* ts) {
    switch (mask) {
        WS_TENSOR_CASE(0) WS_TENSOR_CASE(1) WS_TENSOR_CASE(2) WS_TENSOR_CASE(3)
        default:
            if (NS == 3) {
                switch (mask) { WS_TENSOR_CASE(4) WS_TENSOR_CASE(5) WS_TENSOR_CASE(6) WS_TENSOR_CASE(7) default: break; }
            }
            break;
    }
}

// Split hand-over (panel variant 3).  The only thing the next 8x8 pivot block waits for is the final rank-8 update of ITS OWN
// row tile; everything else a sub-panel produces is needed one step later.  So the chain warp publishes the finished columns
// in two parts -- first the 32-row group that holds the next pivot tile (barrier 3), then the rest (barrier 2) -- and the
// tensor warp that owns the next pivot tile finishes that tile alone and hands it over (barrier 4) before it waits for the
// rest, while the chain warp is still substituting the other row groups.  Barrier 1 ("Xr full") keeps its meaning.
// `prio` = this warp's slot that holds the next pivot tile (-1: none).  The finishing DMMAs run under run-time predicates
// (two k-steps per tile: the serialisation in front of predicated DMMAs does not matter here).
template <int NS, int MASK, int PLW>
__device__ __forceinline__ void ws_tensor_subpanel_split(const double* ZZ, const double* Lc, double* Xp, const int (&mt)[NS],
                                                         const double* const (&rowp)[NS], int c0, int has_prev, int k_pre,
                                                         int k_all, bool wait, int prio, int gq, int tq) {
    // two accumulator pairs per tile (even / odd k-steps): the pre-accumulation is a chain of up to 30 dependent DMMAs per tile
    // (26 cycles each) on a tensor pipe that is 20 % busy; two interleaved chains halve its length
    double cf[NS][2], cg[NS][2];
    double2 a0[NS];
#pragma unroll
    for (int i = 0; i < NS; ++i) {
        cf[i][0] = 0.0; cf[i][1] = 0.0; cg[i][0] = 0.0; cg[i][1] = 0.0;
        a0[i] = make_double2(0.0, 0.0);
        if ((MASK >> i) & 1) a0[i] = *reinterpret_cast<const double2*>(rowp[i] + c0 + 2 * tq);
    }
    if (has_prev) {
#pragma unroll 2
        for (int kk = 0; kk < 64; kk += 8) {
            const double nb0 = -ZZ[(c0 + gq) * PZ + kk + tq], nb1 = -ZZ[(c0 + gq) * PZ + kk + 4 + tq];
#pragma unroll
            for (int i = 0; i < NS; ++i)
                if ((MASK >> i) & 1) {
                    dmma884(cf[i][0], cf[i][1], ZZ[(mt[i] * 8 + gq) * PZ + kk + tq], nb0);
                    dmma884(cg[i][0], cg[i][1], ZZ[(mt[i] * 8 + gq) * PZ + kk + 4 + tq], nb1);
                }
        }
    }
    // (k_pre is a multiple of 8: whole finished sub-panels)
#pragma unroll 2
    for (int k = 0; k < k_pre; k += 8) {
        const double nb0 = -Lc[(k + tq) * PLW + c0 + gq], nb1 = -Lc[(k + 4 + tq) * PLW + c0 + gq];
#pragma unroll
        for (int i = 0; i < NS; ++i)
            if ((MASK >> i) & 1) {
                dmma884(cf[i][0], cf[i][1], Lc[(k + tq) * PLW + mt[i] * 8 + gq], nb0);
                dmma884(cg[i][0], cg[i][1], Lc[(k + 4 + tq) * PLW + mt[i] * 8 + gq], nb1);
            }
    }
#pragma unroll
    for (int i = 0; i < NS; ++i) { cf[i][0] += cg[i][0]; cf[i][1] += cg[i][1]; }
    if (wait) {
        if (prio >= 0) {
            named_bar_sync(3, 64);                       // the group of the pivot tile is published
            for (int k = k_pre; k < k_all; k += 4) {
                const double nb = -Lc[(k + tq) * PLW + c0 + gq];
#pragma unroll
                for (int i = 0; i < NS; ++i)
                    if (((MASK >> i) & 1) && i == prio) dmma884(cf[i][0], cf[i][1], Lc[(k + tq) * PLW + mt[i] * 8 + gq], nb);
            }
#pragma unroll
            for (int i = 0; i < NS; ++i)
                if (((MASK >> i) & 1) && i == prio)
                    *reinterpret_cast<double2*>(Xp + (mt[i] * 8 + gq) * XP + 2 * tq) =
                        make_double2(cf[i][0] + a0[i].x, cf[i][1] + a0[i].y);
            __threadfence_block();
            named_bar_arrive(4, 64);                     // the next pivot block is ready
        }
        named_bar_sync(2, WS_BAR_THREADS);               // all finished columns are published
        for (int k = k_pre; k < k_all; k += 4) {
            const double nb = -Lc[(k + tq) * PLW + c0 + gq];
#pragma unroll
            for (int i = 0; i < NS; ++i)
                if (((MASK >> i) & 1) && i != prio) dmma884(cf[i][0], cf[i][1], Lc[(k + tq) * PLW + mt[i] * 8 + gq], nb);
        }
    }
#pragma unroll
    for (int i = 0; i < NS; ++i)
        if (((MASK >> i) & 1) && !(wait && i == prio))
            *reinterpret_cast<double2*>(Xp + (mt[i] * 8 + gq) * XP + 2 * tq) =
                make_double2(cf[i][0] + a0[i].x, cf[i][1] + a0[i].y);
    __threadfence_block();
    named_bar_arrive(1, WS_BAR_THREADS);
}
#define WS_TENSOR_SPLIT_CASE(M)                                                                                      \
    case M:                                                                                                          \
        ws_tensor_subpanel_split<NS, (M) & ((1 << NS) - 1), PLW>(ZZ, Lc, Xp, mt, rowp, c0, has_prev, k_pre, k_all, wait, \
                                                                 prio, gq, tq);                                       \
        break;
template <int NS, int PLW>
__device__ __forceinline__ void ws_tensor_dispatch_split(int mask, const double* ZZ, const double* Lc, double* Xp,
                                                         const int (&mt)[NS], const double* const (&rowp)[NS], int c0,
                                                         int has_prev, int k_pre, int k_all, bool wait, int prio, int gq,
                                                         int tq) {
    switch (mask) {
        WS_TENSOR_SPLIT_CASE(0) WS_TENSOR_SPLIT_CASE(1) WS_TENSOR_SPLIT_CASE(2) WS_TENSOR_SPLIT_CASE(3)
        default:
            if (NS == 3) {
                switch (mask) {
                    WS_TENSOR_SPLIT_CASE(4) WS_TENSOR_SPLIT_CASE(5) WS_TENSOR_SPLIT_CASE(6) WS_TENSOR_SPLIT_CASE(7)
                    default: break;
                }
            }
            break;
    }
}

template <int OT>
struct WsCfg {
    static constexpr int ROWS = 64 + 8 * OT;        // diagonal block + own rows
    static constexpr int NS = OT == 8 ? 3 : 2;      // row-tile slots per tensor warp
    static constexpr int NG = ROWS / 32;            // 32-row groups of the chain warp
    static constexpr int PLW = OT == 8 ? 132 : 100; // pitch of Lc (== 4 mod 16, >= ROWS)
    static constexpr size_t SMEM = (size_t)(ROWS * PZ + 64 * PLW + ROWS * XP) * sizeof(double);
};
// Row tiles (8 rows each): 0..7 = rows of the diagonal block (tile i is finished once p > i), 8.. = the CTA's own
// rows.  Static assignment, balanced per sub-partition over the sweep (warps w and w+4 share a sub-partition).
template <int OT>
__device__ __forceinline__ int ws_tile(int warp, int slot) {
    if (OT == 8) {
        switch (warp) {
            case 1: return slot == 0 ? 8 : slot == 1 ? 9 : 0;
            case 5: return slot == 0 ? 10 : slot == 1 ? 11 : 1;
            case 2: return slot == 0 ? 12 : slot == 1 ? 2 : 6;
            case 6: return slot == 0 ? 13 : slot == 1 ? 4 : -1;
            case 3: return slot == 0 ? 14 : slot == 1 ? 3 : 7;
            default: return slot == 0 ? 15 : slot == 1 ? 5 : -1;
        }
    } else {
        switch (warp) {
            case 1: return slot == 0 ? 8 : 0;
            case 5: return slot == 0 ? 9 : 3;
            case 2: return slot == 0 ? 10 : 1;
            case 6: return slot == 0 ? 4 : 6;
            case 3: return slot == 0 ? 11 : 2;
            default: return slot == 0 ? 5 : 7;
        }
    }
}

template <int OT, bool SPLIT>
__global__ void __launch_bounds__(256, 1) potrf_panel_ws_kernel(double* __restrict__ A, long long lda,
                                                                double* __restrict__ Ltmp, long long ldt, int k0, int nrb,
                                                                int has_prev, int32_t* info, long long* dbg) {
    using Cfg = WsCfg<OT>;
    constexpr int ROWS = Cfg::ROWS, NS = Cfg::NS, NG = Cfg::NG, PLW = Cfg::PLW;
    extern __shared__ __align__(16) double sm[];
    double* ZZ = sm;                 // [ROWS][PZ] previous-panel values of the diagonal-block rows and of this CTA's rows
    double* Lc = sm + ROWS * PZ;     // [64][PLW]  finished columns, column-major: Lc[col][row]
    double* Xr = Lc + 64 * PLW;      // [ROWS][XP] tensor warps -> chain exchange
    const int tid = threadIdx.x;
    const int b = blockIdx.x;
    const bool has_rows = b < nrb;   // nrb counts blocks of 8*OT rows here
    const int warp = tid >> 5, lane = tid & 31, gq = lane >> 2, tq = lane & 3;
    const double* Ad = A + (long long)k0 * lda + k0;
    double* Ar = A + (long long)(k0 + 64 + 8 * OT * b) * lda + k0;
    span_enter(k0);
    pdl_wait();                       // everything below reads what the previous panel step (and the updates before it) wrote
    if (dbg && tid == 0 && b == 0) dbg[0] = clock64();
    if (has_prev) {
#pragma unroll
        for (int it = 0; it < 8; ++it) {
            const int idx = tid + it * 256;              // 2048 16-byte chunks per 64 x 64 tile
            const int r = idx >> 5, cc = idx & 31;
            cp_async16(ZZ + r * PZ + cc * 2, Ad + (long long)r * lda - 64 + cc * 2, true);
            if (has_rows && it < OT) cp_async16(ZZ + (64 + r) * PZ + cc * 2, Ar + (long long)r * lda - 64 + cc * 2, true);
        }
        cp_async_commit();
        cp_async_wait<0>();
    }
    __syncthreads();
    pdl_launch_dependents();          // the next panel step may be scheduled now (it parks in pdl_wait until this grid is done)
    if (dbg && tid == 0 && b == 0) dbg[1] = clock64();
    if (warp == 4) { span_exit(k0); return; }

    if (warp != 0) {
        // ------------------------------------------------------------------ tensor warps
        int mt[NS];
        const double* rowp[NS];
#pragma unroll
        for (int i = 0; i < NS; ++i) {
            mt[i] = ws_tile<OT>(warp, i);
            const int r = (mt[i] < 0 ? 0 : mt[i]) * 8 + gq;
            rowp[i] = (r < 64) ? (Ad + (long long)r * lda) : (Ar + (long long)(r - 64) * lda);
        }
#pragma unroll 1
        for (int p = 0; p < 8; ++p) {
            const int c0 = p * 8;
            int mask = 0;
#pragma unroll
            for (int i = 0; i < NS; ++i)
                if (mt[i] >= 0 && (mt[i] >= 8 ? has_rows : mt[i] >= p)) mask |= 1 << i;
            double* Xp = Xr;
            if (SPLIT) {
                int prio = -1;                           // this warp's slot holding the pivot tile of this sub-panel
#pragma unroll
                for (int i = 0; i < NS; ++i)
                    if (mt[i] == p) prio = i;
                ws_tensor_dispatch_split<NS, PLW>(mask, ZZ, Lc, Xp, mt, rowp, c0, has_prev, c0 - 8, c0, p >= 1, prio, gq, tq);
                continue;
            }
            // columns [0, c0-8) were published before this warp's previous barrier wait; [c0-8, c0) follow barrier 2
            long long* ts = (dbg && warp == 1 && lane == 0 && b == 0 && (p == 0 || p == 3)) ? dbg + (p == 0 ? 40 : 48) : nullptr;
            ws_tensor_dispatch<NS, PLW>(mask, ZZ, Lc, Xp, mt, rowp, c0, has_prev, c0 - 8, c0, p >= 1 ? 2 : -1, 1,
                                        WS_BAR_THREADS, gq, tq, ts);
            if (dbg && warp == 1 && lane == 0 && b == 0) dbg[16 + p] = clock64();
        }
        span_exit(k0);
        return;
    }

    // ---------------------------------------------------------------------- chain warp: lane owns rows lane + 32 g
    double* Lt = Ltmp + (long long)k0 * ldt + k0;
#pragma unroll 1
    for (int p = 0; p < 8; ++p) {
        const int c0 = p * 8;
        if (SPLIT && p >= 1) named_bar_sync(4, 64);      // the pivot tile alone (handed over early by its owner warp)
        else named_bar_sync(1, WS_BAR_THREADS);
        double D[8][8], rinv[8], acc[NG][8];
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j <= i; ++j) D[i][j] = Xr[(c0 + i) * XP + j];
        int badcol = 8;                       // first non-positive pivot of this block (8 = none)
        if (SPLIT) {
            factor_pivot8_pairs(D, rinv, badcol);        // needs the pivot block only: the other rows arrive meanwhile
            if (p >= 1) named_bar_sync(1, WS_BAR_THREADS);
        }
        bool gon[NG];
#pragma unroll
        for (int g = 0; g < NG; ++g) {
            gon[g] = (32 * g + 31 >= c0) && (g < 2 || has_rows);          // warp-uniform
            if (gon[g]) {
                const double2* q = reinterpret_cast<const double2*>(Xr + (lane + 32 * g) * XP);
                const double2 v0 = q[0], v1 = q[1], v2 = q[2], v3 = q[3];
                acc[g][0] = v0.x; acc[g][1] = v0.y; acc[g][2] = v1.x; acc[g][3] = v1.y;
                acc[g][4] = v2.x; acc[g][5] = v2.y; acc[g][6] = v3.x; acc[g][7] = v3.y;
            } else {
#pragma unroll
                for (int c = 0; c < 8; ++c) acc[g][c] = 0.0;
            }
        }
        if (!SPLIT) factor_pivot8_pairs(D, rinv, badcol);
        if (badcol < 8 && lane == 0 && b == 0) atomicCAS(info, 0, k0 + c0 + badcol + 1);
        // SPLIT: the 32-row group that holds the NEXT pivot tile first (published on its own: barrier 3), then the others
        const int gfirst = (SPLIT && p < 7) ? (c0 + 8) >> 5 : -1;
#pragma unroll
        for (int pass = 0; pass < 2; ++pass) {
#pragma unroll
            for (int g = 0; g < NG; ++g) {
                if (!gon[g] || (pass == 0) != (g == gfirst)) continue;
                const int row = lane + 32 * g;
                const int jrow = row - c0;            // 0..7: a row of the pivot block (entries right of the diagonal are masked)
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    const double xc = acc[g][c] * rinv[c];
                    acc[g][c] = (c <= jrow) ? xc : 0.0;
#pragma unroll
                    for (int cc = c + 1; cc < 8; ++cc) acc[g][cc] = fma(-xc, D[cc][c], acc[g][cc]);
                }
                if (jrow >= 0) {
#pragma unroll
                    for (int c = 0; c < 8; ++c) Lc[(c0 + c) * PLW + row] = acc[g][c];
                }
            }
            if (pass == 0 && gfirst >= 0) {
                __threadfence_block();
                named_bar_arrive(3, 64);
            }
        }
        if (dbg && lane == 0 && b == 0) dbg[2 + p] = clock64();
        if (p < 7) {
            __threadfence_block();       // only shared-memory stores are outstanding here: the global ones follow
            named_bar_arrive(2, WS_BAR_THREADS);
        }
        // finished values to global memory, off the critical path (the tensor warps are already released)
#pragma unroll
        for (int g = 0; g < NG; ++g) {
            const int row = lane + 32 * g;
            if (!gon[g] || row < c0) continue;
            double* dst = nullptr;
            if (g >= 2) dst = Ar + (long long)(row - 64) * lda + c0;
            else if (b == 0) dst = Lt + (long long)row * ldt + c0;           // L_kk is parked: other CTAs still read A_kk
            if (dst) {
#pragma unroll
                for (int c = 0; c < 8; c += 2)
                    *reinterpret_cast<double2*>(dst + c) = make_double2(acc[g][c], acc[g][c + 1]);
            }
        }
    }
    if (dbg && lane == 0 && b == 0) dbg[10] = clock64();
    span_exit(k0);
}

// After the sweep, for every diagonal block at once: move L_kk from Ltmp into A, invert it by
// recursive doubling  inv([[A,0],[B,C]]) = [[A^-1,0],[-C^-1 B A^-1, C^-1]]  (s = 1,2,..,32) into
// the diagonal block of Linv (explicit zeros above the diagonal), and emit sum(log diag).
#define LP 65
__global__ void __launch_bounds__(256) diag_finish_kernel(double* __restrict__ A, long long lda,
                                                          const double* __restrict__ Ltmp, long long ldt,
                                                          double* __restrict__ Linv, long long ldi,
                                                          double* __restrict__ logdet_part) {
    extern __shared__ __align__(16) double sm[];
    double* S = sm;
    double* T = sm + 64 * LP;
    double* U = T + 64 * LP;
    const int tid = threadIdx.x, blk = blockIdx.x;
    const double* Lt = Ltmp + (long long)blk * 64 * ldt + (long long)blk * 64;
    double* Ab = A + (long long)blk * 64 * lda + (long long)blk * 64;
    for (int idx = tid; idx < 4096; idx += 256) {
        const int r = idx >> 6, c = idx & 63;
        const double v = (c <= r) ? Lt[(long long)r * ldt + c] : 0.0;
        S[r * LP + c] = v;
        T[r * LP + c] = 0.0;
        if (c <= r) Ab[(long long)r * lda + c] = v;
    }
    __syncthreads();
    if (tid < 64) T[tid * LP + tid] = 1.0 / S[tid * LP + tid];
    __syncthreads();
    for (int s = 1; s < 64; s <<= 1) {
        const int total = 32 * s, ss = s * s;
        for (int e = tid; e < total; e += 256) {
            int p = e / ss, rem = e - p * ss, rr = rem / s, cc = rem - rr * s;
            int o = p * 2 * s;
            double acc = 0.0;
            for (int l = cc; l < s; ++l) acc += S[(o + s + rr) * LP + o + l] * T[(o + l) * LP + o + cc];
            U[(o + s + rr) * LP + o + cc] = acc;
        }
        __syncthreads();
        for (int e = tid; e < total; e += 256) {
            int p = e / ss, rem = e - p * ss, rr = rem / s, cc = rem - rr * s;
            int o = p * 2 * s;
            double acc = 0.0;
            for (int l = 0; l <= rr; ++l) acc += T[(o + s + rr) * LP + o + s + l] * U[(o + s + l) * LP + o + cc];
            T[(o + s + rr) * LP + o + cc] = -acc;
        }
        __syncthreads();
    }
    double* Lb = Linv + (long long)blk * 64 * ldi + (long long)blk * 64;
    for (int idx = tid; idx < 4096; idx += 256) {
        const int r = idx >> 6, c = idx & 63;
        Lb[(long long)r * ldi + c] = T[r * LP + c];
    }
    if (tid < 32) {
        double v = log(S[tid * LP + tid]) + log(S[(tid + 32) * LP + tid + 32]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (tid == 0) logdet_part[blk] = v;
    }
}

// Single-block version for the pipelined inverse (one launch per panel step, right behind it): 64 threads,
// thread j owns column j of X = L_kk^-1 and solves L_kk X = I by forward substitution with the running sums
// r_i = sum_{k<i} L_ik x_kj in registers (fully unrolled: every operand address is a compile-time constant, the
// L entries are warp-uniform shared-memory broadcasts).  No divergence: for k < j the solution entry is an
// explicit zero.  ~2000 DFMA per thread, a dependent chain of 64 x (DMUL + DFMA).
#define DP 66    // pitch of the transposed block (even: 16-byte aligned pairs)
__global__ void __launch_bounds__(64) diag_inv_kernel(double* __restrict__ A, long long lda,
                                                      const double* __restrict__ Ltmp, long long ldt,
                                                      double* __restrict__ Linv, long long ldi,
                                                      double* __restrict__ logdet_part, int blk0) {
    __shared__ __align__(16) double Lt[64 * DP];      // Lt[k][i] = L[i][k]
    __shared__ double dinv[64];
    const int j = threadIdx.x, blk = blk0 + blockIdx.x;
    const double* Ls = Ltmp + (long long)blk * 64 * ldt + (long long)blk * 64;
    double* Ab = A + (long long)blk * 64 * lda + (long long)blk * 64;
    double* Lb = Linv + (long long)blk * 64 * ldi + (long long)blk * 64;
#pragma unroll 8
    for (int r = 0; r < 64; ++r) {
        const double v = (j <= r) ? Ls[(long long)r * ldt + j] : 0.0;
        Lt[j * DP + r] = v;
        if (j <= r) Ab[(long long)r * lda + j] = v;
    }
    __syncthreads();
    {
        const double d = Lt[j * DP + j];
        dinv[j] = 1.0 / d;
        double lg = log(d);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) lg += __shfl_xor_sync(0xffffffffu, lg, o);
        if ((j & 31) == 0) Lt[63 * DP + 64 + (j >> 5)] = lg;     // two spare slots of the last padded row
    }
    __syncthreads();
    if (j == 0) logdet_part[blk] = Lt[63 * DP + 64] + Lt[63 * DP + 65];
    double r[64];
#pragma unroll
    for (int i = 0; i < 64; ++i) r[i] = 0.0;
#pragma unroll
    for (int k = 0; k < 64; ++k) {
        const double dk = dinv[k];
        double x = -r[k] * dk;
        x = (k == j) ? dk : ((k < j) ? 0.0 : x);
        Lb[(long long)k * ldi + j] = x;
#pragma unroll
        for (int i = k + 1; i < 64; ++i) r[i] = fma(Lt[k * DP + i], x, r[i]);
    }
}

// Row-wise pipeline with 64-row groups: the inverse of the diagonal block and the product X[I, 0:r0) = -X_II T[I, 0:r0) in ONE
// launch (one kernel, one dependency and one launch latency less per panel step beside the chain -- every extra launch there
// costs the chain about a microsecond, profiles/r02_early_loss_ab.txt).  CTA c takes columns [64 c, 64 c + 64); every CTA
// inverts L_kk itself (the forward substitution of diag_inv_kernel, ~2 us, while its T tile is on its way by cp.async);
// CTA 0 also moves L_kk into A, writes X_II into Linv and the block's log-determinant.  blk = 0: no columns, one CTA.
#define XP2 68   // pitch of the X_II / T tiles in shared memory (== 4 mod 16: conflict-free DMMA fragment loads)
__global__ void __launch_bounds__(128) xrow_fused_kernel(double* __restrict__ A, long long lda, const double* __restrict__ Ltmp,
                                                         long long ldt, double* __restrict__ Linv, long long ldi,
                                                         double* __restrict__ logdet_part, int blk) {
    extern __shared__ __align__(16) double sm[];
    double* Lt = sm;                    // [64][DP]   Lt[k][i] = L[i][k]
    double* Xs = Lt + 64 * DP;          // [64][XP2]  X_II[m][k]
    double* Ts = Xs + 64 * XP2;         // [64][XP2]  T tile [k][n]
    double* dinv = Ts + 64 * XP2;       // [64] (+ 2 log partials)
    const int tid = threadIdx.x, c = blockIdx.x;
    const long long r0 = (long long)blk * 64;
    const double* Ls = Ltmp + r0 * ldt + r0;
    if (blk > 0) {                      // T[I rows, 64 c ..) on its way (Ltmp holds T below the block diagonal)
        const double* Tg = Ltmp + r0 * ldt + (long long)c * 64;
#pragma unroll
        for (int q = tid; q < 64 * 32; q += 128) {
            const int row = q >> 5, cc = q & 31;
            cp_async16(Ts + row * XP2 + cc * 2, Tg + (long long)row * ldt + cc * 2, true);
        }
        cp_async_commit();
    }
    for (int q = tid; q < 64 * 64; q += 128) {
        const int r = q >> 6, j = q & 63;
        const double v = (j <= r) ? Ls[(long long)r * ldt + j] : 0.0;
        Lt[j * DP + r] = v;
        if (c == 0 && j <= r) A[(r0 + r) * lda + r0 + j] = v;
    }
    __syncthreads();
    if (tid < 64) {
        const int j = tid;
        const double d = Lt[j * DP + j];
        dinv[j] = 1.0 / d;
        double lg = log(d);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) lg += __shfl_xor_sync(0xffffffffu, lg, o);
        if ((j & 31) == 0) dinv[64 + (j >> 5)] = lg;
    }
    __syncthreads();
    if (c == 0 && tid == 0) logdet_part[blk] = dinv[64] + dinv[65];
    if (tid < 64) {
        const int j = tid;
        double* Lb = Linv + r0 * ldi + r0;
        double r[64];
#pragma unroll
        for (int i = 0; i < 64; ++i) r[i] = 0.0;
#pragma unroll
        for (int k = 0; k < 64; ++k) {
            const double dk = dinv[k];
            double x = -r[k] * dk;
            x = (k == j) ? dk : ((k < j) ? 0.0 : x);
            Xs[k * XP2 + j] = x;
            if (c == 0) Lb[(long long)k * ldi + j] = x;
#pragma unroll
            for (int i = k + 1; i < 64; ++i) r[i] = fma(Lt[k * DP + i], x, r[i]);
        }
    }
    if (blk == 0) return;
    cp_async_wait<0>();
    __syncthreads();
    // C[64 x 64] = -X_II T: 4 warps as 2 x 2, warp tile 32 x 32
    const int warp = tid >> 5, lane = tid & 31, gq = lane >> 2, tq = lane & 3;
    const int am0 = (warp >> 1) * 32, bn0 = (warp & 1) * 32;
    double acc[4][4][2];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
#pragma unroll 4
    for (int kk = 0; kk < 64; kk += 4) {
        double a[4], b[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) a[i] = Xs[(am0 + i * 8 + gq) * XP2 + kk + tq];
#pragma unroll
        for (int j = 0; j < 4; ++j) b[j] = Ts[(kk + tq) * XP2 + bn0 + j * 8 + gq];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
    }
    double* Cg = Linv + r0 * ldi + (long long)c * 64;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j)
            *reinterpret_cast<double2*>(Cg + (long long)(am0 + i * 8 + gq) * ldi + bn0 + j * 8 + 2 * tq) =
                make_double2(-acc[i][j][0], -acc[i][j][1]);
}
// Measured (profiles/r02_xrow_fused.txt): correct, and much SLOWER -- the forward substitution repeated by every CTA of the row
// product makes this launch longer than diag_inv_kernel + the GEMM, and the row pipeline falls behind the chain (cfg2 0.754 -> 1.229 ms).
// Off; kept only as the record of the experiment.
static int g_xfuse = std::getenv("MOGP_XFUSE") ? std::atoi(std::getenv("MOGP_XFUSE")) : 0;
extern "C" int mogp_set_xfuse(int on) { g_xfuse = on; ++g_mogp_cfg_epoch; return 0; }
static cudaError_t launch_xrow_fused(double* A, long long lda, const double* Ltmp, long long ldt, double* Linv, long long ldi,
                                     double* logdet_part, int blk, cudaStream_t st) {
    constexpr size_t SMEM = (size_t)(64 * DP + 2 * 64 * XP2 + 72) * sizeof(double);
    static PerDeviceOnce once;
    if (OnceGuard og{once}; og.needed()) {
        cudaError_t e = cudaFuncSetAttribute(xrow_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM);
        if (e != cudaSuccess) return e;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)std::max(1, blk)); cfg.blockDim = dim3(128);
    cfg.dynamicSmemBytes = SMEM; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributePriority;
    at[0].val.priority = -4;
    cfg.attrs = at; cfg.numAttrs = g_launch_prio ? 1 : 0;
    MOGP_COUNT(1);
    return cudaLaunchKernelEx(&cfg, xrow_fused_kernel, A, lda, Ltmp, ldt, Linv, ldi, logdet_part, blk);
}

// on = 1: arm (resets the slots); on = 0: disarm and copy 2 * n values out (ns)
extern "C" int mogp_panel_spans(int on, unsigned long long* out_host, int n) {
    if (on) {
        std::vector<unsigned long long> init(2 * 136);
        for (int i = 0; i < 136; ++i) { init[2 * i] = ~0ull; init[2 * i + 1] = 0ull; }
        if (cudaMemcpyToSymbol(g_span, init.data(), init.size() * 8) != cudaSuccess) return -2;
        const int one = 1;
        return cudaMemcpyToSymbol(c_span_on, &one, 4) == cudaSuccess ? 0 : -2;
    }
    cudaDeviceSynchronize();
    const int zero = 0;
    cudaMemcpyToSymbol(c_span_on, &zero, 4);
    if (n > 136) n = 136;
    return cudaMemcpyFromSymbol(out_host, g_span, (size_t)2 * n * 8) == cudaSuccess ? 0 : -2;
}
// Timeline stamps of one step (diagnostics): slot i of 16 lives in g_span[256 + i]; launched by enqueue_step when armed
// through mogp_set_stamps (which also re-captures the step graphs).
__global__ void stamp_kernel(int slot) { g_span[256 + slot] = global_ns(); }
int g_mogp_stamps = 0;
extern "C" int mogp_set_stamps(int on) { g_mogp_stamps = on; ++g_mogp_cfg_epoch; return 0; }
cudaError_t launch_stamp(int slot, cudaStream_t st) {
    if (!g_mogp_stamps || slot < 0 || slot >= 16) return cudaSuccess;
    stamp_kernel<<<1, 1, 0, st>>>(slot);
    return cudaGetLastError();
}
static long long* g_panel_dbg = nullptr;   // optional phase timestamps of the first panel kernel
// 0: phase-alternating panel step; 1: warp-specialised, 64 own rows per CTA; 2 (default): warp-specialised with 32 own
// rows per CTA while twice the CTAs fit one wave.  (Two further schedules -- tensor warps a full sub-panel ahead of the
// chain, and two scalar warps -- were measured and removed: DESIGN.md section 4.)  Measured on B200 (profiles/r01_panel_variants.txt): potrf N=2048
// 0.721 / 0.657 / 0.538 ms, N=8192 7.88 / 7.72 / 7.68 ms.  MOGP_PANEL_VARIANT overrides the default for A/B runs.
// Programmatic dependent launch of the panel steps: 0 off, 1 (default) inside graph capture only, 2 always.  Measured on B200
// (profiles/r01_pdl_graphs.txt): inside a replayed graph it hides ~2 us of the ~4 us launch gap between consecutive panel
// steps (cfg2 0.902 -> 0.846 ms, cfg4 3.47 -> 3.34 ms); eagerly the event waits between the steps cancel it.
static int g_panel_pdl = std::getenv("MOGP_PANEL_PDL") ? std::atoi(std::getenv("MOGP_PANEL_PDL")) : 1;
extern "C" int mogp_set_panel_pdl(int v) { g_panel_pdl = v; ++g_mogp_cfg_epoch; return 0; }
extern "C" int mogp_get_panel_pdl(void) { return g_panel_pdl; }
static int g_panel_variant = std::getenv("MOGP_PANEL_VARIANT") ? std::atoi(std::getenv("MOGP_PANEL_VARIANT")) : 2;
extern "C" int mogp_set_panel_variant(int v) { g_panel_variant = v; ++g_mogp_cfg_epoch; return 0; }
extern "C" int mogp_panel_debug(long long* out_host /*19*/) {
    if (!g_panel_dbg) {
        if (cudaMalloc(&g_panel_dbg, 64 * 8) != cudaSuccess) return -2;
        cudaMemset(g_panel_dbg, 0, 64 * 8);
        return 1;
    }
    cudaDeviceSynchronize();
    cudaError_t ce = cudaMemcpy(out_host, g_panel_dbg, 64 * 8, cudaMemcpyDeviceToHost);
    return ce == cudaSuccess ? 0 : -2;
}

// ============================================================================ blocked Cholesky
// In-place lower Cholesky of the padded Np x Np matrix A (row-major, lda), right-looking with
// one step of look-ahead over two streams:
//   S1 (caller's stream): panel step s = previous panel's update of column block s (fused, on
//                         the tensor pipe) + factor + solve of the rows below;
//   S2 (handle's stream): bulk(s) = rank-64 update of the column blocks >= s+2 with panel s,
//                         as one lower-triangular GEMM.
// panel(s+1) runs concurrently with bulk(s); panel(s+2) waits for bulk(s).  For small matrices
// the step time is the panel chain, for large ones the chain hides behind the GEMMs.
// Ltmp is an Np x Np scratch whose diagonal blocks are used; diagonal blocks of Linv get inv(L_kk).
// One warp-specialised panel step; n_cta_rows = number of 8*OT-row blocks below the diagonal block.  Steps that consume a
// previous panel are launched with the programmatic-serialization attribute when g_panel_pdl asks for it (see above).
template <int OT, bool SPLIT>
static void launch_panel_ws(double* A, long long ld, double* Ltmp, long long ldt, int k, int n_cta_rows, int has_prev,
                            int32_t* info, long long* dbgp, cudaStream_t s_) {
    const unsigned grid = (unsigned)std::max(1, n_cta_rows);
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    if (g_panel_pdl == 1) cudaStreamIsCapturing(s_, &cap);
    if (has_prev && (g_panel_pdl == 2 || (g_panel_pdl == 1 && cap == cudaStreamCaptureStatusActive))) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(grid); cfg.blockDim = dim3(256);
        cfg.dynamicSmemBytes = WsCfg<OT>::SMEM; cfg.stream = s_;
        cudaLaunchAttribute at[2];
        at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[0].val.programmaticStreamSerializationAllowed = 1;
        at[1].id = cudaLaunchAttributePriority;
        at[1].val.priority = -5;
        cfg.attrs = at; cfg.numAttrs = g_launch_prio ? 2 : 1;
        if (cudaLaunchKernelEx(&cfg, potrf_panel_ws_kernel<OT, SPLIT>, A, ld, Ltmp, ldt, k, n_cta_rows, has_prev, info, dbgp) == cudaSuccess)
            return;
        cudaGetLastError();               // not supported in this context: plain launches from now on
        g_panel_pdl = 0;
    }
    if (g_launch_prio) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(grid); cfg.blockDim = dim3(256);
        cfg.dynamicSmemBytes = WsCfg<OT>::SMEM; cfg.stream = s_;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributePriority;
        at[0].val.priority = -5;
        cfg.attrs = at; cfg.numAttrs = 1;
        if (cudaLaunchKernelEx(&cfg, potrf_panel_ws_kernel<OT, SPLIT>, A, ld, Ltmp, ldt, k, n_cta_rows, has_prev, info, dbgp) == cudaSuccess)
            return;
        cudaGetLastError();
    }
    potrf_panel_ws_kernel<OT, SPLIT><<<grid, 256, WsCfg<OT>::SMEM, s_>>>(A, ld, Ltmp, ldt, k, n_cta_rows, has_prev, info, dbgp);
}

// Operations of the pipelined triangular inverse, in issue order (see build_inverse_plan).
// kind 0: inverse of diagonal block lo; 1: T = L_BA Linv_AA; 2: Linv_BA = -Linv_BB T  (A = [lo, mid), B = [mid, hi)).
// `ready` = index of the last panel step the operation depends on; `wait` = completion event of the operation that
// finishes the sub-range it needs (-1: none), `done` = its own completion event, `level` = log2 of the pair size.
struct InvOp { int kind, lo, mid, hi, ready, wait, done, level; };
// Block doubling as a recursion over 64-row block ranges: mid - lo = the largest power of two below hi - lo (the
// same pairs the level-batched trtri_padded forms).  Depth-first order is dependency order and non-decreasing in
// `ready`, so the operations can be released panel by panel; every doubling level gets its own stream (its pairs
// are ordered by readiness anyway) so that independent levels do not serialise.  Returns the completion event of
// the whole range.
static int build_inverse_plan(int lo, int hi, std::vector<InvOp>& ops, int& nev) {
    if (hi - lo == 1) { ops.push_back({0, lo, lo, hi, lo, -1, nev, 0}); return nev++; }
    int s = 1, lvl = 0;
    while (2 * s < hi - lo) { s *= 2; ++lvl; }
    const int mid = lo + s;
    const int ea = build_inverse_plan(lo, mid, ops, nev);
    ops.push_back({1, lo, mid, hi, mid - 1, ea, nev++, lvl});      // needs Linv_AA (which implies panel mid-1)
    const int eb = build_inverse_plan(mid, hi, ops, nev);
    ops.push_back({2, lo, mid, hi, hi - 1, eb, nev, lvl});         // follows its own T on the level's stream
    return nev++;
}
// sizes above this take the two-level sweep (K = 256 trailing updates).  Measured (profiles/r01_two_level_threshold.txt):
// N = 4096 potrf 1.73 ms single-level, 1.65 ms two-level; N = 2048 0.52 ms single-level, 0.61 ms two-level.
static long long g_two_level_above = std::getenv("MOGP_TWO_LEVEL_ABOVE") ? std::atoll(std::getenv("MOGP_TWO_LEVEL_ABOVE")) : 2048;
extern "C" int mogp_set_two_level_above(long long v) { g_two_level_above = v; ++g_mogp_cfg_epoch; return 0; }
// timing experiments only: skip the bulk trailing updates (wrong factor, shows the bare panel chain)
static int g_skip_bulk = std::getenv("MOGP_SKIP_BULK") ? std::atoi(std::getenv("MOGP_SKIP_BULK")) : 0;
extern "C" int mogp_set_skip_bulk(int v) { g_skip_bulk = v; ++g_mogp_cfg_epoch; return 0; }
// Host self-check hook (tests/test_host_math.py): the plan for nb blocks, 8 int32 per operation
// [kind, lo, mid, hi, ready, wait, done, level]; returns the number of operations (or -1 if cap is too small).
extern "C" int mogp_host_inverse_plan(int nb, int32_t* out, int cap) {
    if (nb < 1) return -1;
    std::vector<InvOp> ops;
    int ne = 0;
    build_inverse_plan(0, nb, ops, ne);
    if ((int)ops.size() > cap) return -1;
    for (size_t i = 0; i < ops.size(); ++i) {
        const InvOp& o = ops[i];
        const int32_t v[8] = {o.kind, o.lo, o.mid, o.hi, o.ready, o.wait, o.done, o.level};
        for (int j = 0; j < 8; ++j) out[8 * i + j] = v[j];
    }
    return (int)ops.size();
}
// Pipelined inverse on by default (measured: cfg2 0.938 -> 0.899 ms, cfg4 3.62 -> 3.55 ms; profiles/r01_panel_variants.txt).
static int g_trtri_pipe = std::getenv("MOGP_TRTRI_PIPE") ? std::atoi(std::getenv("MOGP_TRTRI_PIPE")) : 1;
extern "C" int mogp_set_trtri_pipe(int v) { g_trtri_pipe = v; ++g_mogp_cfg_epoch; return 0; }

// Row-wise pipeline for the small sizes, whose step time is the panel chain: the block-doubling plan above leaves the last pair
// of EVERY level (and then the whole K^-1 = Linv^T Linv product) for after the last panel step.  Here Linv is built by row
// groups of G 64-blocks instead: X_II (the group's diagonal block, by doubling inside the group) as soon as the group's panels
// are done, X[I, 0:r0) = -X_II T[I, 0:r0) with T[I, :] = sum_{J<I} L[I,J] X[J, :] kept up to date by one rank-(64 G) update per
// finished group (right-looking, into Ltmp below the group diagonal), and K^-1 += X[I, :]^T X[I, :] (lower tiles, rank 64 G)
// accumulated into Kacc as the rows complete.  What is left after the last panel step is the last group's diagonal inverse, one
// thin GEMM and one rank-(64 G) update.
static int g_rowpipe = std::getenv("MOGP_ROWPIPE") ? std::atoi(std::getenv("MOGP_ROWPIPE")) : 1;
static long long g_rowpipe_max_np = std::getenv("MOGP_ROWPIPE_MAX_NP") ? std::atoll(std::getenv("MOGP_ROWPIPE_MAX_NP")) : 4096;
static int g_rowpipe_group = std::getenv("MOGP_ROWPIPE_GROUP") ? std::atoi(std::getenv("MOGP_ROWPIPE_GROUP")) : 1;
extern "C" int mogp_set_rowpipe(int on, long long max_np, int group) {
    if (group != 1 && group != 2 && group != 4 && group != 8) return -1;
    g_rowpipe = on; g_rowpipe_max_np = max_np; g_rowpipe_group = group; ++g_mogp_cfg_epoch;
    return 0;
}
extern "C" int mogp_get_rowpipe(void) { return g_rowpipe; }
// 0: only Linv row-wise, K^-1 = Linv^T Linv afterwards as one product; 1: K^-1 accumulated behind the chain per super-group;
// 2: accumulated in halving chunks (see issue_group_ops)
int g_rowpipe_kinv = std::getenv("MOGP_ROWPIPE_KINV") ? std::atoi(std::getenv("MOGP_ROWPIPE_KINV")) : 0;
static int g_rowpipe_wmin = std::getenv("MOGP_ROWPIPE_WMIN") ? std::atoi(std::getenv("MOGP_ROWPIPE_WMIN")) : 4;
extern "C" int mogp_set_rowpipe_kinv(int mode) { g_rowpipe_kinv = mode; ++g_mogp_cfg_epoch; return 0; }
extern "C" int mogp_set_rowpipe_wmin(int blocks) { if (blocks < 1) return -1; g_rowpipe_wmin = blocks; ++g_mogp_cfg_epoch; return 0; }
// Two levels: the rank-(64 G) updates of a finished group only reach the rows of its own super-group (S blocks); rows beyond
// it and K^-1 get ONE rank-(64 S) update per super-group (a rank-64 update of a 2048^2 matrix runs at ~14 TFLOP/s, a rank-256
// one at ~25: profiles/r01_gemm_sweep.txt).  With `taper` the last super-groups shrink (S/2, S/4, ...) so that the update
// left for after the last panel step is a thin one.
static int g_rowpipe_super = std::getenv("MOGP_ROWPIPE_SUPER") ? std::atoi(std::getenv("MOGP_ROWPIPE_SUPER")) : 1;
static int g_rowpipe_taper = std::getenv("MOGP_ROWPIPE_TAPER") ? std::atoi(std::getenv("MOGP_ROWPIPE_TAPER")) : 0;
extern "C" int mogp_set_rowpipe_super(int super_blocks, int taper) {
    if (super_blocks < 1 || super_blocks > 32) return -1;
    g_rowpipe_super = super_blocks; g_rowpipe_taper = taper; ++g_mogp_cfg_epoch;
    return 0;
}
bool rowpipe_applies(int64_t Np) { return g_rowpipe != 0 && g_trtri_pipe != 0 && Np >= 256 && Np <= g_rowpipe_max_np; }
struct RowGroup { int lo, hi, ev_inv, slo, shi; };      // 64-blocks [lo, hi) of super-group [slo, shi)
// Host self-check hook: the group / super-group partition for nb blocks, 4 int32 per group [lo, hi, slo, shi]
static void rowpipe_partition(int nb, int G, int S, int taper, std::vector<RowGroup>& groups) {
    S = std::max(G, S / G * G);
    for (int pos = 0; pos < nb;) {
        const int rem = nb - pos;
        int size = S;
        if (rem <= S) size = taper ? std::max(G, (rem / 2 + G - 1) / G * G) : rem;
        size = std::min(size, rem);
        for (int lo = pos; lo < pos + size; lo += G) groups.push_back({lo, std::min(pos + size, lo + G), -1, pos, pos + size});
        pos += size;
    }
}
// Chunks of the K^-1 accumulation (mode 2): halves of what is left (nb/2, nb/4, ...) down to wmin blocks, multiples of G.
// Returns the first block of the chunk that ends at block `hi`, or -1 if no chunk ends there.
static int kinv_chunk_start(int nb, int G, int wmin, int hi) {
    wmin = std::max(wmin, G);
    for (int lo = 0; lo < nb;) {
        const int rem = nb - lo;
        int size = rem / 2;
        if (size < wmin) size = std::min(rem, wmin);
        size = std::min(rem, (size + G - 1) / G * G);
        if (lo + size == hi) return lo;
        if (lo + size > hi) return -1;
        lo += size;
    }
    return -1;
}
extern "C" int mogp_host_kinv_chunk_start(int nb, int G, int wmin, int hi) { return kinv_chunk_start(nb, G, wmin, hi); }
extern "C" int mogp_host_rowpipe_partition(int nb, int G, int S, int taper, int32_t* out, int cap) {
    std::vector<RowGroup> g;
    if (nb < 1 || G < 1 || S < 1) return -1;
    rowpipe_partition(nb, G, S, taper, g);
    if ((int)g.size() > cap) return -1;
    for (size_t i = 0; i < g.size(); ++i) { out[4 * i] = g[i].lo; out[4 * i + 1] = g[i].hi; out[4 * i + 2] = g[i].slo; out[4 * i + 3] = g[i].shi; }
    return (int)g.size();
}

cudaError_t potrf_padded(double* A, long long ld, double* Linv, long long ldi, double* Ltmp, long long ldt,
                         int64_t Np, double* logdet_part, int32_t* info, cudaStream_t st, const PotrfStreams* ps,
                         bool* fused_inverse, I8Plan* i8, int i8_slices, double* Kacc, bool* fused_kinv, ZChain* zc) {
    if (fused_inverse) *fused_inverse = false;
    if (fused_kinv) *fused_kinv = false;
    if (zc) zc->done = false;
    cudaError_t e = cudaMemsetAsync(info, 0, sizeof(int32_t), st);
    if (e != cudaSuccess) return e;
    const size_t smem_p = (size_t)(128 * PS + 128 * PZ + 64 + 128 * 8) * sizeof(double);
    const size_t smem_d = (size_t)(3 * 64 * LP) * sizeof(double);
    if ((ld | ldt) & 1) return cudaErrorInvalidValue;          // 16-byte row accesses
    static PerDeviceOnce once;
    const int n_sm = once.sms();
    // variant 0: phase-alternating kernel; 1: warp-specialised, 64 own rows per CTA; 2: warp-specialised, 32 own
    // rows per CTA while twice the CTAs still fit one wave (one CTA per SM), 64 otherwise
    auto launch_panel = [&](int64_t k, int nrb, int has_prev, long long* dbgp, cudaStream_t s_) {
        if (g_panel_variant >= 3 && 2 * nrb <= n_sm)
            launch_panel_ws<4, true>(A, ld, Ltmp, ldt, (int)k, 2 * nrb, has_prev, info, dbgp, s_);
        else if (g_panel_variant >= 3)
            launch_panel_ws<8, true>(A, ld, Ltmp, ldt, (int)k, nrb, has_prev, info, dbgp, s_);
        else if (g_panel_variant >= 2 && 2 * nrb <= n_sm)
            launch_panel_ws<4, false>(A, ld, Ltmp, ldt, (int)k, 2 * nrb, has_prev, info, dbgp, s_);
        else if (g_panel_variant >= 1)
            launch_panel_ws<8, false>(A, ld, Ltmp, ldt, (int)k, nrb, has_prev, info, dbgp, s_);
        else
            potrf_panel_kernel<<<std::max(1, nrb), 256, smem_p, s_>>>(A, ld, Ltmp, ldt, (int)k, nrb, has_prev, info, dbgp);
    };
    if (OnceGuard og{once}; og.needed()) {
        {
            struct KA { const void* f; int smem; };
            const KA ks[4] = {{(const void*)potrf_panel_ws_kernel<8, false>, (int)WsCfg<8>::SMEM},
                              {(const void*)potrf_panel_ws_kernel<4, false>, (int)WsCfg<4>::SMEM},
                              {(const void*)potrf_panel_ws_kernel<8, true>, (int)WsCfg<8>::SMEM},
                              {(const void*)potrf_panel_ws_kernel<4, true>, (int)WsCfg<4>::SMEM}};
            for (const KA& ka : ks) {
                e = cudaFuncSetAttribute(ka.f, cudaFuncAttributeMaxDynamicSharedMemorySize, ka.smem);
                if (e != cudaSuccess) return e;
                e = cudaFuncSetAttribute(ka.f, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
                if (e != cudaSuccess) return e;
            }
        }
        e = cudaFuncSetAttribute(potrf_panel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_p);
        if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(potrf_panel_kernel, cudaFuncAttributePreferredSharedMemoryCarveout,
                                 cudaSharedmemCarveoutMaxShared);
        if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(diag_finish_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_d);
        if (e != cudaSuccess) return e;
    }
    const int nb = (int)(Np / MOGP_NB);
    const bool two = ps != nullptr && ps->s2 != nullptr && ps->s1 != nullptr && nb > 2 && nb <= ps->nev;
    // The panel chain runs on a high-priority stream so that its CTAs are scheduled ahead of the bulk
    // GEMM's (same-priority kernels would serialise behind a saturating GEMM grid).
    cudaStream_t user = st;
    cudaStream_t s2 = two ? ps->s2 : st;
    if (two) {
        if ((e = cudaEventRecord(ps->ev1[0], user)) != cudaSuccess) return e;      // order both streams after the caller's prior work
        if ((e = cudaStreamWaitEvent(ps->s1, ps->ev1[0], 0)) != cudaSuccess) return e;
        if ((e = cudaStreamWaitEvent(s2, ps->ev1[0], 0)) != cudaSuccess) return e;
        st = ps->s1;
    }
    auto finish = [&](cudaEvent_t last_bulk_ev) -> cudaError_t {
        cudaError_t ee;
        if (two) {
            if (last_bulk_ev && (ee = cudaStreamWaitEvent(user, last_bulk_ev, 0)) != cudaSuccess) return ee;
            if ((ee = cudaEventRecord(ps->ev1[ps->nev + 1], st)) != cudaSuccess) return ee;
            if ((ee = cudaStreamWaitEvent(user, ps->ev1[ps->nev + 1], 0)) != cudaSuccess) return ee;
        }
        diag_finish_kernel<<<(unsigned)(Np / 64), 256, smem_d, user>>>(A, ld, Ltmp, ldt, Linv, ldi, logdet_part);
        MOGP_COUNT(1);
        return cudaGetLastError();
    };
    // Pipelined inverse: the diagonal-block inverses and the GEMMs of the block
    // doubling are issued on a fourth stream as soon as the panel steps they depend on are done, so that only
    // the last pair of every level is left when the factorisation ends.
    const bool pipe = fused_inverse != nullptr && g_trtri_pipe != 0 && two && ps->s4 != nullptr && ps->evp != nullptr &&
                      ps->evq != nullptr && ldi == ld && ldt == ld && 3 * nb + 16 <= ps->nevq && nb <= 128 &&
                      (Np <= 4096 || g_trtri_pipe >= 2);      // measured at N = 8192: no gain over the level-batched trtri_padded
                                                              // (13.33 ms against 7.71 + 5.80 ms): the machine is already full
    std::vector<InvOp> plan;
    size_t next_op = 0, next_group = 0;
    bool level_used[8] = {};
    bool s4_used = false;
    // (Kacc == NULL: Linv row-wise only, K^-1 is left to the caller)
    const bool rowp = pipe && (Kacc == nullptr || (fused_kinv != nullptr && Kacc != Ltmp && ps->ev_kinv != nullptr)) &&
                      rowpipe_applies(Np) && 4 * nb + 16 <= ps->nevq;
    std::vector<RowGroup> groups;
    const bool xfuse = rowp && g_xfuse != 0 && g_rowpipe_group == 1;
    if (rowp) {
        int ne = 0;
        rowpipe_partition(nb, g_rowpipe_group, g_rowpipe_super, g_rowpipe_taper, groups);
        for (RowGroup& gr : groups) gr.ev_inv = build_inverse_plan(gr.lo, gr.hi, plan, ne);
    } else if (pipe) {
        int ne = 0;
        build_inverse_plan(0, nb, plan, ne);
    }
    // group-level operations of the row-wise pipeline, issued once the group's last panel step is in the chain
    auto issue_group_ops = [&](const RowGroup& gr, int gi) -> cudaError_t {
        cudaError_t ee;
        cudaStream_t sX = ps->sl[6], sW = ps->sl[7];
        level_used[6] = true;
        level_used[7] = Kacc != nullptr;
        const long long r0 = (long long)gr.lo * 64, R = (long long)(gr.hi - gr.lo) * 64, c1 = (long long)gr.hi * 64;
        cudaEvent_t ex = ps->evq[3 * nb + gi];
        if (xfuse) {                 // diagonal inverse + X[I, 0:r0) in one launch, right behind the group's (one) panel step
            if ((ee = cudaStreamWaitEvent(sX, ps->evp[gr.hi - 1], 0)) != cudaSuccess) return ee;
            if ((ee = launch_xrow_fused(A, ld, Ltmp, ldt, Linv, ldi, logdet_part, gr.lo, sX)) != cudaSuccess) return ee;
        } else if ((ee = cudaStreamWaitEvent(sX, ps->evq[gr.ev_inv], 0)) != cudaSuccess) return ee;
        if (r0 > 0 && !xfuse) {      // X[I, 0:r0) = -X_II T[I, 0:r0)   (X_II lower triangular: k ends at the row tile)
            GemmArgs g{};
            g.A = Linv + r0 * ld + r0; g.lda = ld;
            g.B = Ltmp + r0 * ld; g.ldb = ld;
            g.C = Linv + r0 * ld; g.ldc = ld;
            g.M = (int)R; g.N = (int)r0; g.K = (int)R;
            g.khi_mode = 1; g.alpha = -1.0; g.beta = 0.0; g.prio = 5;
            if ((ee = launch_gemm(0, 0, g, 1, sX)) != cudaSuccess) return ee;
        }
        if ((ee = cudaEventRecord(ex, sX)) != cudaSuccess) return ee;
        if (zc) {                    // z = Linv y for the rows that are complete now; behind the last ones, the early loss
            cudaStream_t sZ = ps->sl[3];
            level_used[3] = true;
            if ((ee = cudaStreamWaitEvent(sZ, ex, 0)) != cudaSuccess) return ee;
            if ((ee = launch_trmv_rows(Linv, ld, zc->ypad, zc->z, r0, c1, sZ)) != cudaSuccess) return ee;
            if (c1 == Np) {
                if (zc->early_host &&
                    (ee = launch_lml_early(zc->z, logdet_part, info, zc->N, Np, zc->early_host, zc->early_ctr, sZ)) != cudaSuccess)
                    return ee;
                zc->done = true;
            }
        }
        const long long s0 = (long long)gr.slo * 64, s1 = (long long)gr.shi * 64;       // the super-group's rows / columns
        if (c1 < s1) {               // fine: T[rest of the super-group, 0:c1) += L[those rows, group columns] X[I, 0:c1)
            GemmArgs g{};            // (first touch of the group's own columns)
            g.A = A + c1 * ld + r0; g.lda = ld;
            g.B = Linv + r0 * ld; g.ldb = ld;
            g.C = Ltmp + c1 * ld; g.ldc = ld;
            g.M = (int)(s1 - c1); g.N = (int)c1; g.K = (int)R;
            g.alpha = 1.0; g.beta = 1.0; g.first_touch_col1 = (int)r0 + 1; g.prio = 4;
            if ((ee = launch_gemm(0, 0, g, 1, sX)) != cudaSuccess) return ee;
        }
        if (gr.hi == gr.shi) {       // the super-group is complete: one rank-(s1 - s0) update of everything beyond it
            if (s1 < Np) {           // coarse: T[rows beyond, 0:s1) += L[rows beyond, super-group columns] X[super-group rows, 0:s1)
                GemmArgs g{};
                g.A = A + s1 * ld + s0; g.lda = ld;
                g.B = Linv + s0 * ld; g.ldb = ld;
                g.C = Ltmp + s1 * ld; g.ldc = ld;
                g.M = (int)(Np - s1); g.N = (int)s1; g.K = (int)(s1 - s0);
                g.alpha = 1.0; g.beta = 1.0; g.first_touch_col1 = (int)s0 + 1; g.prio = 2;
                if ((ee = launch_gemm(0, 0, g, 1, sX)) != cudaSuccess) return ee;
            }
        }
        if (!Kacc) return cudaSuccess;
        // K^-1[0:w1, 0:w1) (lower) += X[chunk rows, 0:w1)^T X[chunk rows, 0:w1) once the rows [w0, w1) of a chunk are complete
        // (first touch of rows >= w0).  Chunks: the super-groups (mode 1) or halves of what is left (mode 2: nb/2, nb/4, ...
        // down to g_rowpipe_wmin blocks), which keeps these products deep while leaving a thin one for after the chain.
        long long w0 = -1, w1 = (long long)gr.hi * 64;
        if (g_rowpipe_kinv == 2) {
            const int lo = kinv_chunk_start(nb, g_rowpipe_group, g_rowpipe_wmin, gr.hi);
            if (lo >= 0) w0 = (long long)lo * 64;
        } else if (gr.hi == gr.shi) {
            w0 = s0;
        }
        if (w0 >= 0) {
            if ((ee = cudaStreamWaitEvent(sW, ex, 0)) != cudaSuccess) return ee;
            GemmArgs g{};
            g.A = Linv + w0 * ld; g.lda = ld;
            g.B = Linv + w0 * ld; g.ldb = ld;
            g.C = Kacc; g.ldc = ld;
            g.M = (int)w1; g.N = (int)w1; g.K = (int)(w1 - w0);
            if (w0 == 0) g.klo_mode = 2;     // X[0:w1, 0:w1) is lower triangular: k starts at the row tile
            g.lower = 1; g.alpha = 1.0; g.beta = 1.0; g.first_touch_row1 = (int)w0 + 1; g.prio = 1;
            if ((ee = launch_gemm(1, 0, g, 1, sW)) != cudaSuccess) return ee;
        }
        return cudaSuccess;
    };
    auto issue_inverse_ops = [&](int s) -> cudaError_t {
        cudaError_t ee;
        if ((ee = cudaEventRecord(ps->evp[s], st)) != cudaSuccess) return ee;
        for (; next_op < plan.size() && plan[next_op].ready <= s; ++next_op) {
            const InvOp& op = plan[next_op];
            if (op.kind == 0) {
                if (xfuse) continue;          // the diagonal inverse is part of the group's fused launch (issue_group_ops)
                if ((ee = cudaStreamWaitEvent(ps->s4, ps->evp[s], 0)) != cudaSuccess) return ee;
                s4_used = true;
                diag_inv_kernel<<<1, 64, 0, ps->s4>>>(A, ld, Ltmp, ldt, Linv, ldi, logdet_part, op.lo);
                MOGP_COUNT(1);
                if ((ee = cudaGetLastError()) != cudaSuccess) return ee;
                if ((ee = cudaEventRecord(ps->evq[op.done], ps->s4)) != cudaSuccess) return ee;
                continue;
            }
            cudaStream_t sv = ps->sl[op.level];
            level_used[op.level] = true;
            if ((ee = cudaStreamWaitEvent(sv, ps->evq[op.wait], 0)) != cudaSuccess) return ee;
            const long long o = (long long)op.lo * 64, S = (long long)(op.mid - op.lo) * 64, MB = (long long)(op.hi - op.mid) * 64;
            GemmArgs g{};
            if (op.kind == 1) {          // T = L_BA * Linv_AA   (Linv_AA lower triangular: k starts at the column tile)
                g.A = A + (o + S) * ld + o; g.lda = ld;
                g.B = Linv + o * ld + o; g.ldb = ld;
                g.C = Ltmp + (o + S) * ld + o; g.ldc = ld;
                g.M = (int)MB; g.N = (int)S; g.K = (int)S;
                g.klo_mode = 1; g.pair = 1; g.alpha = 1.0; g.beta = 0.0;
                if (rowp) g.prio = 5;
            } else {                     // Linv_BA = -Linv_BB * T   (Linv_BB lower triangular: k ends at the row tile)
                g.A = Linv + (o + S) * ld + (o + S); g.lda = ld;
                g.B = Ltmp + (o + S) * ld + o; g.ldb = ld;
                g.C = Linv + (o + S) * ld + o; g.ldc = ld;
                g.M = (int)MB; g.N = (int)S; g.K = (int)MB;
                g.khi_mode = 1; g.pair = 2; g.alpha = -1.0; g.beta = 0.0;
                if (rowp) g.prio = 5;
            }
            if ((ee = launch_gemm(0, 0, g, 1, sv)) != cudaSuccess) return ee;
            if ((ee = cudaEventRecord(ps->evq[op.done], sv)) != cudaSuccess) return ee;
        }
        for (; next_group < groups.size() && groups[next_group].hi - 1 <= s; ++next_group)
            if ((ee = issue_group_ops(groups[next_group], (int)next_group)) != cudaSuccess) return ee;
        return cudaSuccess;
    };
    // join of the pipelined inverse: bulk stream, panel chain and every inverse stream back into the caller's stream
    auto finish_pipe = [&](cudaEvent_t last_bulk_ev) -> cudaError_t {
        cudaError_t ee;
        if (last_bulk_ev && (ee = cudaStreamWaitEvent(user, last_bulk_ev, 0)) != cudaSuccess) return ee;
        if ((ee = cudaEventRecord(ps->ev1[ps->nev + 1], st)) != cudaSuccess) return ee;
        if ((ee = cudaStreamWaitEvent(user, ps->ev1[ps->nev + 1], 0)) != cudaSuccess) return ee;
        if (s4_used) {               // (not forked at all when the diagonal inverses are part of the fused row launches)
            if ((ee = cudaEventRecord(ps->evp[ps->nev + 1], ps->s4)) != cudaSuccess) return ee;
            if ((ee = cudaStreamWaitEvent(user, ps->evp[ps->nev + 1], 0)) != cudaSuccess) return ee;
        }
        for (int l = 0; l < 8; ++l) {
            if (!level_used[l]) continue;
            if (rowp && Kacc && l == 7) {        // the K^-1 accumulation is joined by the caller (the solves do not need it)
                if ((ee = cudaEventRecord(ps->ev_kinv, ps->sl[l])) != cudaSuccess) return ee;
                *fused_kinv = true;
                continue;
            }
            cudaEvent_t ej = ps->evq[ps->nevq - 8 + l];
            if ((ee = cudaEventRecord(ej, ps->sl[l])) != cudaSuccess) return ee;
            if ((ee = cudaStreamWaitEvent(user, ej, 0)) != cudaSuccess) return ee;
        }
        *fused_inverse = true;
        return cudaSuccess;
    };
    if (Np > g_two_level_above) {
        // Large matrices: two-level updates (K = 64 inside a 256-column outer panel, one K = 256 SYRK per
        // outer panel) keep the trailing-matrix traffic down.  The SYRK is split into the next outer
        // panel's columns (priority) and the rest; both run on S2 while S1 factors the next outer panel.
        cudaEvent_t last = nullptr;
        const bool two3 = two && ps->s3 != nullptr;
        // three levels for the largest matrices: inside a super-panel of SW = 1024 columns the rank-256 DMMA updates only
        // reach the super-panel's own columns; what lies beyond is updated once per super-panel with K = 1024 on the int8 pipe
        const int64_t SW = 1024;
        const bool super = i8 != nullptr && g_i8_potrf_min > 0 && Np >= g_i8_potrf_min && Np % SW == 0;
        int J = 0;
        for (int64_t K0 = 0; K0 < Np; K0 += MOGP_NB_OUT, ++J) {
            const int64_t Wd = std::min<int64_t>(MOGP_NB_OUT, Np - K0), Kend = K0 + Wd;
            const int64_t Kse = super ? (K0 / SW + 1) * SW : Np;          // first column beyond this super-panel
            if (two && J >= 1 && (e = cudaStreamWaitEvent(st, ps->ev2[2 * (J - 1)], 0)) != cudaSuccess) return e;
            cudaEvent_t last_inner = nullptr;
            for (int64_t k = K0; k < Kend; k += MOGP_NB) {
                const int nrb = (int)((Np - k - MOGP_NB) / MOGP_NB);
                const int ks = (int)(k / MOGP_NB);
                // panel(k) reads its column updated by the inner update issued two steps earlier (third stream)
                if (two3 && k >= K0 + 2 * MOGP_NB && (e = cudaStreamWaitEvent(st, ps->ev3[ks - 2], 0)) != cudaSuccess) return e;
                launch_panel(k, nrb, k > K0 ? 1 : 0, nullptr, st);
                MOGP_COUNT(1);
                if (pipe && (e = issue_inverse_ops(ks)) != cudaSuccess) return e;
                const int64_t c0 = k + 2 * MOGP_NB, Nc = Kend - c0, M = Np - c0;
                if (Nc > 0 && M > 0) {                      // remaining columns of this outer panel
                    cudaStream_t si = st;
                    if (two3) {                             // off the panel chain: overlaps the next panel step
                        if ((e = cudaEventRecord(ps->ev3[ps->nev + 1], st)) != cudaSuccess) return e;
                        if ((e = cudaStreamWaitEvent(ps->s3, ps->ev3[ps->nev + 1], 0)) != cudaSuccess) return e;
                        si = ps->s3;
                    }
                    GemmArgs u{};
                    u.A = A + c0 * ld + k; u.lda = ld;
                    u.B = A + c0 * ld + k; u.ldb = ld;
                    u.C = A + c0 * ld + c0; u.ldc = ld;
                    u.M = (int)M; u.N = (int)Nc; u.K = MOGP_NB;
                    u.lower = 1; u.alpha = -1.0; u.beta = 1.0;
                    if ((e = launch_gemm(0, 1, u, 1, si)) != cudaSuccess) return e;
                    if (two3) {
                        if ((e = cudaEventRecord(ps->ev3[ks], si)) != cudaSuccess) return e;
                        last_inner = ps->ev3[ks];
                    }
                }
            }
            if (two3 && last_inner && (e = cudaStreamWaitEvent(st, last_inner, 0)) != cudaSuccess) return e;   // join s3
            if (Kend >= Np) break;
            if (two) {
                if ((e = cudaEventRecord(ps->ev1[J + 1], st)) != cudaSuccess) return e;
                if ((e = cudaStreamWaitEvent(s2, ps->ev1[J + 1], 0)) != cudaSuccess) return e;
            }
            const int64_t W2 = std::min<int64_t>(MOGP_NB_OUT, Np - Kend);
            if (super && Kend == Kse) {
                // end of a super-panel: ONE rank-1024 update of everything below / right of it on the int8 tensor pipe
                // (i8mm.cu) replaces the four rank-256 DMMA updates beyond the super-panel; same split (next super-panel's
                // columns first, then the rest) and the same events as the DMMA updates below
                if ((e = i8_syrk_update(i8, A, ld, Kse, Kse - SW, Np, 0, i8_slices, s2)) != cudaSuccess) return e;
                if (two && (e = cudaEventRecord(ps->ev2[2 * J], s2)) != cudaSuccess) return e;
                last = two ? ps->ev2[2 * J] : nullptr;
                if (Np - Kse - SW > 0) {
                    if ((e = i8_syrk_update(i8, A, ld, Kse, Kse - SW, Np, 1, i8_slices, s2)) != cudaSuccess) return e;
                    if (two && (e = cudaEventRecord(ps->ev2[2 * J + 1], s2)) != cudaSuccess) return e;
                    last = two ? ps->ev2[2 * J + 1] : nullptr;
                }
                continue;
            }
            {                                               // priority: columns of the next outer panel
                GemmArgs u{};
                u.A = A + Kend * ld + K0; u.lda = ld;
                u.B = A + Kend * ld + K0; u.ldb = ld;
                u.C = A + Kend * ld + Kend; u.ldc = ld;
                u.M = (int)(Np - Kend); u.N = (int)W2; u.K = (int)Wd;
                u.lower = 1; u.alpha = -1.0; u.beta = 1.0;
                if ((e = launch_gemm(0, 1, u, 1, s2)) != cudaSuccess) return e;
                if (two && (e = cudaEventRecord(ps->ev2[2 * J], s2)) != cudaSuccess) return e;
                last = two ? ps->ev2[2 * J] : nullptr;
            }
            const int64_t R0 = Kend + W2, MR = Np - R0, NR = std::min<int64_t>(Np, Kse) - R0;
            if (MR > 0 && NR > 0) {                         // the rest of the trailing matrix (of this super-panel)
                GemmArgs u{};
                u.A = A + R0 * ld + K0; u.lda = ld;
                u.B = A + R0 * ld + K0; u.ldb = ld;
                u.C = A + R0 * ld + R0; u.ldc = ld;
                u.M = (int)MR; u.N = (int)NR; u.K = (int)Wd;
                u.lower = 1; u.alpha = -1.0; u.beta = 1.0;
                if ((e = launch_gemm(0, 1, u, 1, s2)) != cudaSuccess) return e;
                if (two && (e = cudaEventRecord(ps->ev2[2 * J + 1], s2)) != cudaSuccess) return e;
                last = two ? ps->ev2[2 * J + 1] : nullptr;
            }
        }
        return pipe ? finish_pipe(last) : finish(last);
    }
    int last_bulk = -1;
    for (int s = 0; s < nb; ++s) {
        const int64_t k = (int64_t)s * MOGP_NB;
        const int nrb = nb - s - 1;
        if (two && s >= 2 && (e = cudaStreamWaitEvent(st, ps->ev2[s - 2], 0)) != cudaSuccess) return e;
        launch_panel(k, nrb, s > 0 ? 1 : 0, (s == 1) ? g_panel_dbg : nullptr, st);
        MOGP_COUNT(1);
        if (pipe && (e = issue_inverse_ops(s)) != cudaSuccess) return e;
        const int64_t c0 = k + 2 * MOGP_NB, M = Np - c0;                          // column blocks >= s+2
        if (M > 0) {
            if (two) {
                if ((e = cudaEventRecord(ps->ev1[s + 1], st)) != cudaSuccess) return e;
                if ((e = cudaStreamWaitEvent(s2, ps->ev1[s + 1], 0)) != cudaSuccess) return e;
            }
            GemmArgs u{};
            u.A = A + c0 * ld + k; u.lda = ld;
            u.B = A + c0 * ld + k; u.ldb = ld;
            u.C = A + c0 * ld + c0; u.ldc = ld;
            u.M = (int)M; u.N = (int)M; u.K = MOGP_NB;
            u.lower = 1; u.alpha = -1.0; u.beta = 1.0;
            if (rowp) u.prio = 4;
            e = g_skip_bulk ? cudaSuccess : launch_gemm(0, 1, u, 1, s2);
            if (e != cudaSuccess) return e;
            if (two) {
                if ((e = cudaEventRecord(ps->ev2[s], s2)) != cudaSuccess) return e;
                last_bulk = s;
            }
        }
    }
    if (pipe) return finish_pipe(last_bulk >= 0 ? ps->ev2[last_bulk] : nullptr);
    return finish(two && last_bulk >= 0 ? ps->ev2[last_bulk] : nullptr);
}

// ============================================================================ recursive factor + inverse (large sizes)
// chol(A) for A = [[A11, .], [A21, A22]] (halves h):  L11 = chol(A11), X11 = L11^-1;  L21 = A21 X11^T;  A22 -= L21 L21^T;
// L22 = chol(A22), X22 = L22^-1;  X21 = -X22 (L21 X11).  The recursion stops at `leaf` rows, which the blocked sweep above
// factors and inverts (row-wise pipeline); every product above the leaves has K = its block size and runs on the int8 tensor
// pipe (i8_blk_first / i8_blk_second) -- at N = 8192 that is 94 % of the 2 N^3 / 3 flops of factor + inverse, against the
// rank-64 / 256 / 1024 updates of the blocked sweep.  The multiplication by the explicit inverse X11 in place of a triangular
// solve costs the same flops; its error is bounded by cond(L11) eps like that of the solve.
static int g_rchol = std::getenv("MOGP_RCHOL") ? std::atoi(std::getenv("MOGP_RCHOL")) : 1;
static long long g_rchol_min_np = std::getenv("MOGP_RCHOL_MIN_NP") ? std::atoll(std::getenv("MOGP_RCHOL_MIN_NP")) : 4096;
static long long g_rchol_leaf = std::getenv("MOGP_RCHOL_LEAF") ? std::atoll(std::getenv("MOGP_RCHOL_LEAF")) : 2048;
static int g_rchol_overlap = std::getenv("MOGP_RCHOL_OVERLAP") ? std::atoi(std::getenv("MOGP_RCHOL_OVERLAP")) : 1;
extern "C" int mogp_set_rchol_overlap(int on) { g_rchol_overlap = on; ++g_mogp_cfg_epoch; return 0; }
extern "C" int mogp_set_rchol(int on, long long min_np, long long leaf) {
    if (leaf < 1024 || (leaf & (leaf - 1)) != 0) return -1;
    g_rchol = on; g_rchol_min_np = min_np; g_rchol_leaf = leaf; ++g_mogp_cfg_epoch;
    return 0;
}
extern "C" int mogp_get_rchol(long long* min_np, long long* leaf) {
    if (min_np) *min_np = g_rchol_min_np;
    if (leaf) *leaf = g_rchol_leaf;
    return g_rchol;
}
extern "C" int mogp_rchol_applies(long long Np) { return rchol_applies(Np) ? 1 : 0; }
// Leaf size for a padded size: Np / 2^k (k = 1..3) with the leaf a multiple of 128 rows between half and 5/4 of the nominal
// leaf (1024 .. 2560 for 2048); the largest such leaf is taken.  0: the recursive scheme does not apply (odd multiples of 128, ...).
int64_t rchol_leaf_for(int64_t Np) {
    if (!g_rchol || Np < g_rchol_min_np) return 0;
    const int64_t lo = g_rchol_leaf / 2, hi = g_rchol_leaf + g_rchol_leaf / 4;
    for (int k = 1; k <= 3; ++k) {
        if (Np % ((int64_t)128 << k) != 0) break;
        const int64_t leaf = Np >> k;
        if (leaf >= lo && leaf <= hi) return leaf;
        if (leaf < lo) break;
    }
    return 0;
}
extern "C" long long mogp_rchol_leaf_for(long long Np) { return rchol_leaf_for(Np); }
bool rchol_applies(int64_t Np) { return rchol_leaf_for(Np) > 0; }
// info[0] = first failing leaf's pivot index (1-based, global); leaf_info[i] are relative to leaf i
__global__ void rchol_info_kernel(int32_t* info, int nleaf, int leaf_rows) {
    int32_t v = 0;
    for (int i = 0; i < nleaf && v == 0; ++i)
        if (info[1 + i] != 0) v = info[1 + i] + i * leaf_rows;
    info[0] = v;
}
cudaError_t rchol_padded(double* A, long long ld, double* Linv, double* Ltmp, int64_t Np, double* logdet_part, int32_t* info,
                         cudaStream_t st, const PotrfStreams* ps, I8Plan* i8, int i8_slices, int want_inverse) {
    const int64_t leaf = rchol_leaf_for(Np);
    if (leaf <= 0 || !i8 || !i8_blk_ok(i8, Np, ld, i8_slices, leaf)) return cudaErrorNotSupported;
    int leaf_idx = 0;
    cudaError_t err = cudaSuccess;
    // overlap (g_rchol_overlap): the products of a block that its second half does not need at once (T = L21 X11, the A22 update
    // below the second half's first leaf) run on a side stream while the second half's leaf chains -- which leave most of the
    // machine idle -- proceed; one side stream / event triple per recursion depth (0: blocks whose halves are leaves)
    const bool ovl = g_rchol_overlap != 0 && ps && ps->sl[4] && ps->sl[5] && ps->evq && ps->nevq >= 4 * 34 + 48;
    auto rec = [&](auto&& self, int64_t o, int64_t n, int need_inv, cudaEvent_t pending) -> void {
        if (err != cudaSuccess) return;
        if (n <= leaf) {
            bool fused = false;
            double* As = A + o * (ld + 1);
            double* Xs = Linv + o * (ld + 1);
            double* Ts = Ltmp + o * (ld + 1);
            err = potrf_padded(As, ld, Xs, ld, Ts, ld, n, logdet_part + o / 64, info + 1 + leaf_idx, st, ps, &fused, nullptr, i8_slices,
                               nullptr, nullptr);
            if (err == cudaSuccess && !fused) err = trtri_padded(As, Xs, Ts, n, ld, st, nullptr, i8_slices);
            ++leaf_idx;
            return;
        }
        const int64_t h = n / 2;
        self(self, o, h, 1, pending);
        if (err != cudaSuccess) return;
        // (`pending`: the parent's A22 update below this block's first leaf runs on the parent's side stream; it must be done
        //  before this block's products read those rows)
        if (pending && (err = cudaStreamWaitEvent(st, pending, 0)) != cudaSuccess) return;
        const int depth = h > leaf ? 1 : 0;
        const bool deep = h > 2 * leaf;                      // a third level would share the side resources of depth 1: sequential
        I8BlkAsync as{};
        const bool use_as = ovl && !deep;
        if (use_as) {
            as.side = ps->sl[4 + depth];
            as.ev_fork = ps->evq[ps->nevq - 40 + 3 * depth];
            as.ev_rest = ps->evq[ps->nevq - 40 + 3 * depth + 1];
            as.ev_T = ps->evq[ps->nevq - 40 + 3 * depth + 2];
            as.split_rows = depth ? leaf : 0;
            as.own_ops = depth;
        }
        if ((err = i8_blk_first(i8, A, Linv, Ltmp, ld, o, h, need_inv, i8_slices, st, use_as ? &as : nullptr)) != cudaSuccess) return;
        const bool forked = use_as && (need_inv || as.split_rows > 0);
        self(self, o + h, h, need_inv, (forked && as.split_rows > 0) ? as.ev_rest : nullptr);
        if (err != cudaSuccess) return;
        if (forked && (err = cudaStreamWaitEvent(st, as.ev_T, 0)) != cudaSuccess) return;       // joins the side stream
        if (!need_inv) return;
        err = i8_blk_second(i8, Linv, Ltmp, ld, o, h, i8_slices, st);
    };
    rec(rec, 0, Np, want_inverse, nullptr);
    if (err != cudaSuccess) return err;
    rchol_info_kernel<<<1, 1, 0, st>>>(info, leaf_idx, (int)leaf);
    MOGP_COUNT(1);
    return cudaGetLastError();
}

// ============================================================================ triangular inverse
// Linv = L^-1 by level-batched block doubling: at level s (in 64-blocks) every pair of
// adjacent diagonal super-blocks (A, B) gets Linv_BA = -Linv_BB * (L_BA * Linv_AA).
// Two batched GEMMs per level (+2 for a ragged last pair); `scratch` holds L_BA*Linv_AA.
cudaError_t trtri_padded(double* L, double* Linv, double* scratch, int64_t Np, long long ld, cudaStream_t st, I8Plan* i8,
                         int i8_slices) {
    const int64_t nblk = Np / 64;
    for (int64_t s = 1; s < nblk; s *= 2) {
        const int64_t S = s * 64;
        if (i8) {          // the large levels (almost all of the flops) on the int8 tensor pipe (i8mm.cu)
            cudaError_t e8 = i8_trtri_level(i8, L, Linv, scratch, Np, ld, S, i8_slices, st);
            if (e8 == cudaSuccess) continue;
            if (e8 != cudaErrorNotSupported) return e8;
        }
        const int64_t nfull = nblk / (2 * s);
        const int64_t o_rem = nfull * 2 * S;
        const int64_t remB = (nblk - nfull * 2 * s > s) ? (nblk - nfull * 2 * s - s) * 64 : 0;
        for (int pass = 0; pass < 2; ++pass) {
            const int64_t o = pass == 0 ? 0 : o_rem;
            const int64_t MB = pass == 0 ? S : remB;
            const int batch = pass == 0 ? (int)nfull : 1;
            if (MB <= 0 || batch <= 0) continue;
            GemmArgs a{};
            a.A = L + (o + S) * ld + o; a.lda = ld; a.strideA = 2 * S * (ld + 1);
            a.B = Linv + o * ld + o; a.ldb = ld; a.strideB = 2 * S * (ld + 1);
            a.C = scratch + (o + S) * ld + o; a.ldc = ld; a.strideC = 2 * S * (ld + 1);
            a.M = (int)MB; a.N = (int)S; a.K = (int)S;
            a.klo_mode = 1; a.pair = 1; a.alpha = 1.0; a.beta = 0.0;
            cudaError_t e = launch_gemm(0, 0, a, batch, st);
            if (e != cudaSuccess) return e;
            GemmArgs b{};
            b.A = Linv + (o + S) * ld + (o + S); b.lda = ld; b.strideA = 2 * S * (ld + 1);
            b.B = scratch + (o + S) * ld + o; b.ldb = ld; b.strideB = 2 * S * (ld + 1);
            b.C = Linv + (o + S) * ld + o; b.ldc = ld; b.strideC = 2 * S * (ld + 1);
            b.M = (int)MB; b.N = (int)S; b.K = (int)MB;
            b.khi_mode = 1; b.pair = 2; b.alpha = -1.0; b.beta = 0.0;
            e = launch_gemm(0, 0, b, batch, st);
            if (e != cudaSuccess) return e;
        }
    }
    return cudaSuccess;
}

// W(lower) = 0.5 * (Linv^T Linv - avec avec^T)   (avec == NULL -> plain K^-1 lower)
__global__ void zero_vec_kernel(double* v, int64_t n) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) v[i] = 0.0;
}

cudaError_t kinv_padded(const double* Linv, double* W, int64_t Np, long long ld, const double* avec, cudaStream_t st) {
    // K^-1[i][j] = sum_{k >= max(i,j)} Linv[k][i] Linv[k][j]: one TN GEMM over the lower tiles with the k-range
    // clipped at the row tile.  Measured alternatives that were slower and are not used: accumulating over
    // uniform-K row panels (0.25 vs 0.17 ms at N=2048, 6.8 vs 5.7 ms at N=8192) and pairing heavy with light
    // rows in one CTA (pair mode 3: 0.25 ms at N=2048) -- a lone 4-warp CTA reaches only ~60% of an SM's DMMA
    // rate, so longer CTAs lengthen the critical path more than the balance gains.
    GemmArgs g{};
    g.A = Linv; g.lda = ld;
    g.B = Linv; g.ldb = ld;
    g.C = W; g.ldc = ld;
    g.M = (int)Np; g.N = (int)Np; g.K = (int)Np;
    g.lower = 1; g.klo_mode = 2;
    g.alpha = 1.0; g.beta = 0.0;
    if (avec) { g.epi = 1; g.avec = avec; }
    return launch_gemm(1, 0, g, 1, st);
}

// ============================================================================ mat-vecs
// z[r] = sum_{c<=r} Linv[r][c] * y[c]; one warp per row.
__global__ void __launch_bounds__(256) trmv_lower_kernel(const double* __restrict__ Linv, long long ld,
                                                         const double* __restrict__ y, double* __restrict__ z,
                                                         int64_t row0, int64_t Np) {
    const int lane = threadIdx.x & 31;
    const int64_t r = row0 + blockIdx.x * 8 + (threadIdx.x >> 5);
    if (r >= Np) return;
    const double* row = Linv + r * ld;
    double acc = 0.0;
    for (int64_t c = lane; c <= r; c += 32) acc += row[c] * y[c];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) z[r] = acc;
}

cudaError_t launch_trmv_lower(const double* Linv, long long ld, const double* y, double* z, int64_t Np, cudaStream_t st) {
    trmv_lower_kernel<<<(unsigned)((Np + 7) / 8), 256, 0, st>>>(Linv, ld, y, z, 0, Np);
    MOGP_COUNT(1);
    return cudaGetLastError();
}
// rows [row0, row1) only (each row is summed exactly as in the full launch)
cudaError_t launch_trmv_rows(const double* Linv, long long ld, const double* y, double* z, int64_t row0, int64_t row1,
                             cudaStream_t st) {
    if (row1 <= row0) return cudaSuccess;
    trmv_lower_kernel<<<(unsigned)((row1 - row0 + 7) / 8), 256, 0, st>>>(Linv, ld, y, z, row0, row1);
    MOGP_COUNT(1);
    return cudaGetLastError();
}

// column pass, two deterministic phases.
__global__ void __launch_bounds__(256) colpass_kernel(const double* __restrict__ M, long long ld,
                                                      const double* __restrict__ v, int64_t rows, int64_t cols,
                                                      int64_t rows_per, double* __restrict__ part) {
    __shared__ double sd[4][64], sq[4][64];
    const int tx = threadIdx.x & 63, ty = threadIdx.x >> 6;
    const int64_t c = blockIdx.x * 64 + tx;
    const int64_t rb = blockIdx.y * rows_per, re = min(rows, rb + rows_per);
    double ad = 0.0, aq = 0.0;
    if (c < cols)
        for (int64_t r = rb + ty; r < re; r += 4) {
            const double m = M[r * ld + c];
            ad += m * v[r];
            aq += m * m;
        }
    sd[ty][tx] = ad;
    sq[ty][tx] = aq;
    __syncthreads();
    if (ty == 0 && c < cols) {
        ad = (sd[0][tx] + sd[1][tx]) + (sd[2][tx] + sd[3][tx]);
        aq = (sq[0][tx] + sq[1][tx]) + (sq[2][tx] + sq[3][tx]);
        part[(blockIdx.y * 2 + 0) * cols + c] = ad;
        part[(blockIdx.y * 2 + 1) * cols + c] = aq;
    }
}
__global__ void colpass_reduce_kernel(const double* __restrict__ part, int nsplit, int64_t cols,
                                      double* __restrict__ out_dot, double* __restrict__ out_sq) {
    const int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (c >= cols) return;
    double ad = 0.0, aq = 0.0;
    for (int s = 0; s < nsplit; ++s) {
        ad += part[(s * 2 + 0) * cols + c];
        aq += part[(s * 2 + 1) * cols + c];
    }
    if (out_dot) out_dot[c] = ad;
    if (out_sq) out_sq[c] = aq;
}

cudaError_t launch_colpass(const double* M, long long ld, const double* v, int64_t rows, int64_t cols, double* part,
                           size_t part_cap, double* out_dot, double* out_sq, cudaStream_t st) {
    int64_t nsplit = std::max<int64_t>(1, std::min<int64_t>(64, rows / 128));
    while ((size_t)(nsplit * 2 * cols) > part_cap && nsplit > 1) nsplit /= 2;
    if ((size_t)(nsplit * 2 * cols) > part_cap) return cudaErrorInvalidValue;
    const int64_t rows_per = (rows + nsplit - 1) / nsplit;
    dim3 grid((unsigned)((cols + 63) / 64), (unsigned)nsplit);
    colpass_kernel<<<grid, 256, 0, st>>>(M, ld, v, rows, cols, rows_per, part);
    colpass_reduce_kernel<<<(unsigned)((cols + 255) / 256), 256, 0, st>>>(part, (int)nsplit, cols, out_dot, out_sq);
    MOGP_COUNT(2);
    return cudaGetLastError();
}

__global__ void pad_copy_kernel(const double* __restrict__ src, int64_t n, double* __restrict__ dst, int64_t np) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < np) dst[i] = i < n ? src[i] : 0.0;
}
cudaError_t launch_pad_copy(const double* src, int64_t n, double* dst, int64_t np, cudaStream_t st) {
    pad_copy_kernel<<<(unsigned)((np + 255) / 256), 256, 0, st>>>(src, n, dst, np);
    MOGP_COUNT(1);
    return cudaGetLastError();
}

// dir 0: lower triangle of user (n x n) -> padded work (np x np, identity padding)
// dir 1: lower triangle of work -> user
__global__ void copy_tri_kernel(int dir, double* user, long long ldu, double* work, long long ldw, int64_t n, int64_t np) {
    const int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int64_t r = blockIdx.y;
    if (c > r || r >= np) return;
    if (dir == 0) {
        double v = (r < n) ? user[r * ldu + c] : (r == c ? 1.0 : 0.0);
        work[r * ldw + c] = v;
    } else if (r < n) {
        user[r * ldu + c] = work[r * ldw + c];
    }
}
cudaError_t launch_copy_tri(int dir, double* user, long long ldu, double* work, long long ldw, int64_t n, int64_t np,
                            cudaStream_t st) {
    dim3 grid((unsigned)((np + 255) / 256), (unsigned)np);
    copy_tri_kernel<<<grid, 256, 0, st>>>(dir, user, ldu, work, ldw, n, np);
    MOGP_COUNT(1);
    return cudaGetLastError();
}

__global__ void pred_var_kernel(const double* __restrict__ chanbuf, int C, const int32_t* __restrict__ chan_s,
                                const double* __restrict__ colsq, int64_t M, double* __restrict__ var,
                                const double* __restrict__ kss) {
    const int64_t s = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (s >= M) return;
    if (kss) { var[s] = kss[s] - colsq[s]; return; }
    int c = 0;
    while (c + 1 < C && s >= chan_s[c + 1]) ++c;
    var[s] = chanbuf[C + c] - colsq[s];
}
cudaError_t launch_pred_var(const double* chanbuf, int C, const int32_t* chan_s_dev, const double* colsq, int64_t M,
                            double* var, cudaStream_t st, const double* kss) {
    pred_var_kernel<<<(unsigned)((M + 255) / 256), 256, 0, st>>>(chanbuf, C, chan_s_dev, colsq, M, var, kss);
    MOGP_COUNT(1);
    return cudaGetLastError();
}

// ============================================================================ fp64 peak probes
__global__ void __launch_bounds__(256) peak_dmma_kernel(double* out, int iters) {
    double c[8][2];
#pragma unroll
    for (int i = 0; i < 8; ++i) c[i][0] = c[i][1] = 0.0;
    double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) dmma884(c[i][0], c[i][1], a, b);
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
    if (s == 123.456) out[0] = s;
}
__global__ void __launch_bounds__(256) peak_dfma_kernel(double* out, int iters) {
    double c[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) c[i] = i;
    double a = 1.0 + threadIdx.x * 1e-9, b = 1e-9;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) c[i] = fma(c[i], a, b);
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += c[i];
    if (s == 123.456) out[0] = s;
}

cudaError_t run_peak_fp64(double* dmma_tflops, double* dfma_tflops) {
    double* d = nullptr;
    cudaError_t e = cudaMalloc(&d, 8);
    if (e != cudaSuccess) return e;
    cudaDeviceProp prop;
    int dev = 0;
    cudaGetDevice(&dev);
    cudaGetDeviceProperties(&prop, dev);
    const int blocks = prop.multiProcessorCount * 4, threads = 256;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    float ms = 0.f;
    const int iters = 20000;
    for (int rep = 0; rep < 2; ++rep) {
        cudaEventRecord(e0);
        peak_dmma_kernel<<<blocks, threads>>>(d, iters);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
    }
    // one m8n8k4 = 8*8*4 FMA = 512 flop per warp
    *dmma_tflops = (double)blocks * (threads / 32) * (double)iters * 8.0 * 512.0 / (ms * 1e-3) / 1e12;
    for (int rep = 0; rep < 2; ++rep) {
        cudaEventRecord(e0);
        peak_dfma_kernel<<<blocks, threads>>>(d, iters);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
    }
    *dfma_tflops = (double)blocks * threads * (double)iters * 8.0 * 2.0 / (ms * 1e-3) / 1e12;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    e = cudaGetLastError();
    cudaFree(d);
    return e;
}

// ============================================================================ latency probes
// Single-warp dependent-chain latencies (cycles per op), used to budget the Cholesky panel chain.
__global__ void latency_probe_kernel(double* out, double seed) {
    __shared__ double sbuf[64];
    const int n = 256;
    double x = seed + threadIdx.x * 1e-9, y = 1.0000001, z = 1e-9;
    long long t0, t1;
    sbuf[threadIdx.x] = x; sbuf[threadIdx.x + 32] = y;
    __syncthreads();
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < n; ++i) x = fma(x, y, z);
    t1 = clock64();
    double r0 = (double)(t1 - t0) / n;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < n; ++i) x = x * y;
    t1 = clock64();
    double r1 = (double)(t1 - t0) / n;
    x = fabs(x) + 2.0;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < n; ++i) x = rsqrt(x) + 2.0;
    t1 = clock64();
    double r2 = (double)(t1 - t0) / n;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < n; ++i) x = sqrt(x) + 2.0;
    t1 = clock64();
    double r3 = (double)(t1 - t0) / n;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < n; ++i) x = 1.0 / x + 2.0;
    t1 = clock64();
    double r4 = (double)(t1 - t0) / n;
    int idx = threadIdx.x;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < n; ++i) idx = (int)sbuf[idx & 63] & 31;
    t1 = clock64();
    double r5 = (double)(t1 - t0) / n;
    double c0 = x, c1 = y;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < n; ++i) dmma884(c0, c1, y, z);
    t1 = clock64();
    double r6 = (double)(t1 - t0) / n;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < n; ++i) x = __shfl_xor_sync(0xffffffffu, x, 1) + 1.0;
    t1 = clock64();
    double r7 = (double)(t1 - t0) / n;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < n; ++i) __syncthreads();
    t1 = clock64();
    double r8 = (double)(t1 - t0) / n;
    // tensor-pipe wake-up: cost of a short DMMA burst right after `gap` cycles of scalar fp64 work
    double wake[4];
    {
        const int gaps[4] = {0, 64, 256, 1024};
        for (int gi = 0; gi < 4; ++gi) {
            double w = x;
            for (int rep = 0; rep < 3; ++rep) {
                for (int i = 0; i < gaps[gi]; ++i) w = fma(w, y, z);       // ~8 cycles each
                t0 = clock64();
#pragma unroll
                for (int i = 0; i < 8; ++i) dmma884(c0, c1, w * 1e-30 + y, z);
                t1 = clock64();
                wake[gi] = (double)(t1 - t0);
                x += w * 1e-300;
            }
        }
    }
    if (threadIdx.x == 0) {
        out[10] = wake[0]; out[11] = wake[1]; out[12] = wake[2]; out[13] = wake[3];
        out[0] = r0; out[1] = r1; out[2] = r2; out[3] = r3; out[4] = r4; out[5] = r5; out[6] = r6; out[7] = r7; out[8] = r8;
        out[9] = x + idx + c0 + c1;
    }
}
extern "C" int mogp_probe_latency(double* out_host /*16*/) {
    double* d = nullptr;
    if (cudaMalloc(&d, 16 * 8) != cudaSuccess) return -2;
    latency_probe_kernel<<<1, 32>>>(d, 1.5);
    latency_probe_kernel<<<1, 32>>>(d, 1.5);
    cudaError_t e = cudaMemcpy(out_host, d, 16 * 8, cudaMemcpyDeviceToHost);
    cudaFree(d);
    return e == cudaSuccess ? 0 : -2;
}

// Issue-rate probe: `nw` warps of sub-partition 0 (warps 0, 4, 8, ...) each run 8 independent DFMA chains.
// out[0] = cycles per DFMA instruction seen by warp 0 (a lone warp reaches about half of the pipe's rate).
__global__ void issue_probe_kernel(double* out, int iters, int nw, double seed) {
    const int warp = threadIdx.x >> 5;
    double c[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) c[i] = seed + i + threadIdx.x * 1e-9;
    const double a = 1.0000001, bb = 1e-9;
    __syncthreads();
    if ((warp & 3) == 0 && (warp >> 2) < nw) {
        const long long t0 = clock64();
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int i = 0; i < 8; ++i) c[i] = fma(c[i], a, bb);
        }
        const long long t1 = clock64();
        if (threadIdx.x == 0) out[0] = (double)(t1 - t0) / ((double)iters * 8);
    }
    double sum = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) sum += c[i];
    if (sum == 123.456) out[1] = sum;
}
extern "C" int mogp_probe_issue(double* out_host /*4*/) {
    double* d = nullptr;
    if (cudaMalloc(&d, 4 * 8) != cudaSuccess) return -2;
    for (int nw = 1; nw <= 4; ++nw) {
        for (int rep = 0; rep < 2; ++rep) issue_probe_kernel<<<1, 512>>>(d, 512, nw, 1.5);
        if (cudaMemcpy(out_host + (nw - 1), d, 8, cudaMemcpyDeviceToHost) != cudaSuccess) { cudaFree(d); return -2; }
    }
    cudaFree(d);
    return 0;
}

// Contention probe: warps 0..3 run a dependent DFMA chain (the shape of the Cholesky pivot chain) while warps
// 4..7 (one per SM sub-partition, sharing the fp64 pipe with them) stream DMMAs on `nacc` independent
// accumulators.  out[0] = chain cycles per DFMA, out[1] = cycles per DMMA of one tensor warp.
template <int NACC>
__global__ void contention_probe_kernel(double* out, int chain_ops, int dmma_iters, double seed) {
    const int warp = threadIdx.x >> 5;
    double y = 1.0000001, z = 1e-9;
    if (warp < 4) {
        double x = seed + threadIdx.x * 1e-9;
        __syncthreads();
        const long long t0 = clock64();
#pragma unroll 16
        for (int i = 0; i < chain_ops; ++i) x = fma(x, y, z);
        const long long t1 = clock64();
        if (threadIdx.x == 0) { out[0] = (double)(t1 - t0) / chain_ops; out[2] = x; }
    } else {
        double c[NACC > 0 ? NACC : 1][2];
#pragma unroll
        for (int a = 0; a < (NACC > 0 ? NACC : 1); ++a) { c[a][0] = seed; c[a][1] = seed * 0.5; }
        __syncthreads();
        const long long t0 = clock64();
        if (NACC > 0) {
            for (int i = 0; i < dmma_iters; ++i) {
#pragma unroll
                for (int a = 0; a < NACC; ++a) dmma884(c[a][0], c[a][1], y, z);
            }
        }
        const long long t1 = clock64();
        if (threadIdx.x == 128) {
            double s = 0;
            for (int a = 0; a < (NACC > 0 ? NACC : 1); ++a) s += c[a][0] + c[a][1];
            out[1] = NACC > 0 ? (double)(t1 - t0) / ((double)dmma_iters * NACC) : 0.0;
            out[3] = s;
        }
    }
}
// Same probe with the layout of the warp-specialised panel step: chain on warp 0 only, DMMA streams (3
// accumulators) on warps 1,2,3,5,6,7, warp 4 idle.  out[0] = chain cycles per DFMA, out[1] = cycles per DMMA.
__global__ void contention_split_kernel(double* out, int chain_ops, int dmma_iters, double seed) {
    const int warp = threadIdx.x >> 5;
    double y = 1.0000001, z = 1e-9;
    __syncthreads();
    if (warp == 0) {
        double x = seed + threadIdx.x * 1e-9;
        const long long t0 = clock64();
#pragma unroll 16
        for (int i = 0; i < chain_ops; ++i) x = fma(x, y, z);
        const long long t1 = clock64();
        if (threadIdx.x == 0) { out[0] = (double)(t1 - t0) / chain_ops; out[2] = x; }
    } else if (warp != 4) {
        double c[3][2] = {{seed, seed}, {seed, seed}, {seed, seed}};
        const long long t0 = clock64();
        for (int i = 0; i < dmma_iters; ++i) {
#pragma unroll
            for (int a = 0; a < 3; ++a) dmma884(c[a][0], c[a][1], y, z);
        }
        const long long t1 = clock64();
        if (threadIdx.x == 32) { out[1] = (double)(t1 - t0) / ((double)dmma_iters * 3); out[3] = c[0][0] + c[1][1] + c[2][0]; }
    }
}
extern "C" int mogp_probe_contention(double* out_host /*15*/) {
    double* d = nullptr;
    if (cudaMalloc(&d, 16 * 8) != cudaSuccess) return -2;
    const int chain = 2048;
    for (int mode = 0; mode < 5; ++mode) {
        cudaMemset(d, 0, 16 * 8);
        for (int rep = 0; rep < 2; ++rep) {
            if (mode == 4) contention_split_kernel<<<1, 256>>>(d, chain, 1500, 1.5);
            if (mode == 0) contention_probe_kernel<0><<<1, 256>>>(d, chain, 0, 1.5);
            if (mode == 1) contention_probe_kernel<1><<<1, 256>>>(d, chain, 4000, 1.5);
            if (mode == 2) contention_probe_kernel<2><<<1, 256>>>(d, chain, 2000, 1.5);
            if (mode == 3) contention_probe_kernel<4><<<1, 256>>>(d, chain, 1000, 1.5);
        }
        double h[4];
        if (cudaMemcpy(h, d, 4 * 8, cudaMemcpyDeviceToHost) != cudaSuccess) { cudaFree(d); return -2; }
        out_host[3 * mode] = h[0];
        out_host[3 * mode + 1] = h[1];
        out_host[3 * mode + 2] = 0.0;
    }
    cudaFree(d);
    return 0;
}
