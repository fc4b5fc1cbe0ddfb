// fp64 GEMM on the int8 tensor pipe of sm_100a (tcgen05.mma kind::i8, int32 accumulators in TMEM) by Ozaki-style
// slicing -- the tcgen05 path of the O(N^3) stages for large problems (there is no fp64 tcgen05.mma).
//
//   x[i, :] = 2^ex[i] * sum_s X_s[i, :] 2^(-7 (s+1)),   X_s signed 7-bit digits stored as int8, one exponent per row
//   C[i, j] = alpha * 2^(ea[i] + eb[j]) * sum_{d < S} 2^(-7 (d+2)) * ( sum_{s+t=d} A_s B_t^T )[i, j]   (+ beta C)
//
// Every digit product is exact in int32 (|digit| <= 64, K * 2^12 * S < 2^31 up to K = 65536 at S = 8), so the only
// rounding is the truncation of the operands to 7 S bits below their row maximum plus the fp64 sums over d: with
// S = 7 a product carries ~1e-14 of the largest entry, with S = 8 it is as accurate as an fp64 GEMM
// (profiles/r02_i8_bringup_v1.log, profiles/r01_ozaki_accuracy_study.txt).
//
// Kernel structure (one CTA per 128 x 64 output tile, 6 warps):
//   * all S digit planes of the tile's A rows and B rows for one 32-deep K chunk form one pipeline stage (42 KB at
//     S = 7), so a chunk is read from L2 ONCE for its S (S+1) / 2 MMAs (the bring-up kernel streamed the operands
//     once per digit pair and was L2-bound at 0.5 POP/s);
//   * the S anti-diagonal sums live in S x 64 TMEM columns at the same time (448 of 512 at S = 7);
//   * warp 0 lane 0: producer -- 1-D TMA bulk copies (cp.async.bulk, mbarrier complete_tx) of pre-tiled digits;
//     warp 1 lane 0: issues the tcgen05.mma stream and releases stages with tcgen05.commit;
//     warps 2-5: epilogue -- tcgen05.ld of the S accumulators, fp64 recombination, scaling, store.
//   * the sliced operands are stored in global memory in exactly the shared-memory image the MMA wants (no-swizzle
//     K-major core matrices: 8 rows x 16 bytes contiguous), tiled as [k chunk][row tile of 128][slice][k16][row][16 B],
//     so a stage is one 4 S KB copy for A and 2 S copies of 1 KB for B.
//   * per-tile K ranges (triangular operands) come from a host-built tile list.
#include "common.cuh"
#include <algorithm>
#include <cstdio>
#include <cstdlib>

#define I8_TM 128          // tile rows (UMMA M)
#define I8_TN 64           // tile columns (UMMA N)
#define I8_KC 32           // K bytes per stage = one UMMA K step for 8-bit operands
#define I8_STAGES 4
#define I8_SMAX 8
#define I8_BITS 7
#define I8_PANEL 1024      // width of the Cholesky super-panel whose rank-1024 trailing update runs here

__device__ __forceinline__ uint32_t i8_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ------------------------------------------------------------------ slicing
// Row maxima -> exponents.  The operand has R rows in `batch` groups of Rb = R / batch rows: element (row r, k) is
// src[(r / Rb) * bstride + (r % Rb) * rs + k * cs].  tri = 1: rows are zero for k < (r % Rb) rounded down to 64 (a
// lower-triangular block read column-wise); tri = 2: zero for k > r % Rb (a lower-triangular block read row-wise).
__global__ void __launch_bounds__(256) i8_rowmax_kernel(const double* __restrict__ src, long long rs, long long cs,
                                                        long long bstride, int R, int Rb, int K, int tri,
                                                        int* __restrict__ ex_bits) {
    // block: 64 operand rows x 4 k-phases; grid.y splits K
    const int r = blockIdx.x * 64 + (threadIdx.x & 63);
    const int ph = threadIdx.x >> 6;
    const int kper = (K + gridDim.y - 1) / gridDim.y;
    int k0 = blockIdx.y * kper, k1 = min(K, k0 + kper);
    const int rl0 = (int)((blockIdx.x * 64) % Rb);
    if (tri == 1) k0 = max(k0, rl0);
    if (tri == 2) k1 = min(k1, rl0 + 64);
    double m = 0.0;
    if (r < R) {
        const double* x = src + (long long)(r / Rb) * bstride + (long long)(r % Rb) * rs;
        for (int k = k0 + ph; k < k1; k += 4) m = fmax(m, fabs(x[(long long)k * cs]));
    }
    __shared__ double red[256];
    red[threadIdx.x] = m;
    __syncthreads();
    if (ph == 0 && r < R) {
        m = fmax(fmax(red[threadIdx.x], red[threadIdx.x + 64]), fmax(red[threadIdx.x + 128], red[threadIdx.x + 192]));
        // non-negative doubles order like their bit patterns: keep the high word (sign + exponent + 20 mantissa bits)
        if (m > 0.0) atomicMax(ex_bits + r, __double2hiint(m));
    }
}

// digits[kc][rt][s][k16][row % 128][k % 16], kc = k / 32, rt = row / 128.  One block per (kc, rt): 128 rows x 32 k.
// Thread (row, k16) slices 16 consecutive k of one row and writes one 16-byte vector per digit plane.
// Blocks that lie entirely in the zero part of a triangular operand (see i8_rowmax_kernel) are never read by the
// tile lists and are skipped.
__global__ void __launch_bounds__(256) i8_slice_tiled_kernel(const double* __restrict__ src, long long rs, long long cs,
                                                             long long bstride, int R, int Rb, int K, int S, int tri,
                                                             const int* __restrict__ ex_bits, int8_t* __restrict__ digits,
                                                             int nrt) {
    const int kc = blockIdx.x, rt = blockIdx.y;
    const int rl0 = (rt * I8_TM) % Rb;
    if (tri == 1 && kc * I8_KC + I8_KC <= rl0) return;
    if (tri == 2 && kc * I8_KC >= rl0 + I8_TM) return;
    const int row = rt * I8_TM + (threadIdx.x & 127), k16 = threadIdx.x >> 7;
    int e = 0;
    bool nz = false;
    if (row < R) {
        const int hi = ex_bits[row];
        if (hi > 0) { nz = true; e = ((hi >> 20) & 0x7ff) - 1022 + 1; }     // |x| < 2^(e-1): x 2^-e in (-1/2, 1/2)
    }
    const double* x = src + (long long)((row < R ? row : 0) / Rb) * bstride + (long long)((row < R ? row : 0) % Rb) * rs;
    double v[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        const int k = kc * I8_KC + k16 * 16 + i;
        v[i] = (nz && k < K) ? x[(long long)k * cs] : 0.0;
    }
    const double sc = nz ? __hiloint2double((1023 - e) << 20, 0) : 0.0;      // 2^-e
    int8_t* base = digits + (((long long)kc * nrt + rt) * S) * (2 * I8_TM * 16) + (long long)(k16 * I8_TM + (threadIdx.x & 127)) * 16;
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] *= sc;
    for (int s = 0; s < S; ++s) {
        uint32_t w[4] = {0u, 0u, 0u, 0u};
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const double r = v[i] * 128.0;
            const double d = rint(r);                                         // |r| <= 64 -> digit in [-64, 64]
            v[i] = r - d;                                                     // exact
            w[i >> 2] |= ((uint32_t)(int)d & 0xffu) << ((i & 3) * 8);
        }
        *reinterpret_cast<uint4*>(base + (long long)s * (2 * I8_TM * 16)) = make_uint4(w[0], w[1], w[2], w[3]);
    }
}

// ex[r] (int) from the bit pattern of the row maximum; rows of zeros get INT_MIN / 4 (their products scale to 0)
__global__ void i8_exponent_kernel(const int* __restrict__ ex_bits, int* __restrict__ ex, int R) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= R) return;
    const int hi = ex_bits[r];
    ex[r] = hi > 0 ? ((hi >> 20) & 0x7ff) - 1022 + 1 : -(1 << 20);
}

// ------------------------------------------------------------------ tcgen05 helpers
__device__ __forceinline__ void i8_mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(i8_smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void i8_mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(i8_smem_u32(bar)), "r"(parity)
        : "memory");
}
__device__ __forceinline__ void i8_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(i8_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void i8_bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(i8_smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(i8_smem_u32(bar))
                 : "memory");
}
// K-major, no swizzle: core matrix = 8 rows x 16 bytes contiguous; SBO = stride between 8-row groups, LBO = stride
// between the two 16-byte K chunks of one K = 32 instruction (cute::UMMA::SmemDescriptor, version 1).
__device__ __forceinline__ uint64_t i8_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) | (1ull << 46);
}
__device__ __forceinline__ void i8_mma(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
        : "memory");
}
// A operand from tensor memory (lane = row, 8 columns = the 32 K bytes of one UMMA K step), B from shared memory
__device__ __forceinline__ void i8_mma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d), "r"(tmem_a), "l"(db), "r"(idesc), "r"(accumulate)
        : "memory");
}
// shared memory -> tensor memory: 128 rows x 256 bits described by a matrix descriptor
__device__ __forceinline__ void i8_cp_128x256b(uint32_t tmem_dst, uint64_t sdesc) {
    asm volatile("tcgen05.cp.cta_group::1.128x256b [%0], %1;" ::"r"(tmem_dst), "l"(sdesc) : "memory");
}
__device__ __forceinline__ void i8_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(i8_smem_u32(bar)) : "memory");
}

// One 128 x 64 output tile: first row of the A operand, first row of the B operand (a multiple of 64), K chunks
// [kc0, kc1), and the element offset of C[0][0] of the tile.
struct I8Tile { int m0, n0, kc0, kc1; long long c_off; };

struct __align__(128) I8Smem {
    int8_t a[I8_STAGES][I8_SMAX][2][I8_TM][16];      // [stage][slice][k16][row][16 B]   4 KB per slice
    int8_t b[I8_STAGES][I8_SMAX][2][I8_TN][16];      //                                   2 KB per slice
    uint64_t full[I8_STAGES], empty[I8_STAGES], acc_full;
    uint32_t tmem_base;
};

struct I8Args {
    const int8_t* A; const int* ea; int nrtA;       // tiled digits / exponents / row tiles of the A operand
    const int8_t* B; const int* eb; int nrtB;
    const I8Tile* tiles;
    int S;
    double alpha, beta;
    double* C; long long ldc;
};

// TS = true: the S digit planes of the A tile are first copied from shared into tensor memory (tcgen05.cp, 8 columns per
// plane next to the accumulators: 7 * 64 + 7 * 8 = 504 of 512 columns at S = 7) and every MMA takes A from there.  With
// both operands in shared memory (TS = false) an M = 128, N = 64, K = 32 MMA reads 6 KB of shared memory for 32 cycles of
// math, i.e. 192 B/cycle against the ~128 B/cycle an SM's shared memory delivers -- together with the TMA writes that is
// what held the first version at ~51 % of the int8 peak; from tensor memory the A plane is read from shared memory once
// per K chunk instead of once per MMA.
template <int S, bool TS>
__global__ void __launch_bounds__(192, 1) i8_gemm_tiles_kernel(I8Args g) {
    extern __shared__ __align__(128) unsigned char i8_raw[];
    I8Smem& sm = *reinterpret_cast<I8Smem*>(i8_raw);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const I8Tile t = g.tiles[blockIdx.x];
    const int nk = t.kc1 - t.kc0;
    constexpr uint32_t TMEM_COLS = 512;             // S * 64 <= 512 rounded up to a power of two (S = 7, 8)

    if (tid == 0) {
        for (int i = 0; i < I8_STAGES; ++i) { i8_mbar_init(&sm.full[i], 1); i8_mbar_init(&sm.empty[i], 1); }
        i8_mbar_init(&sm.acc_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        __syncwarp();
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(i8_smem_u32(&sm.tmem_base)), "n"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem0 = sm.tmem_base;

    if (warp == 0) {
        // ------------------------------------------------------------------ producer
        if (lane == 0) {
            const int rtA = t.m0 / I8_TM, rtB = t.n0 / I8_TM, hb = (t.n0 % I8_TM) / I8_TN;
            constexpr uint32_t A_BYTES = S * 2 * I8_TM * 16, B_BYTES = S * 2 * I8_TN * 16;
            int stage = 0;
            uint32_t phase = 0;
            for (int kc = t.kc0; kc < t.kc1; ++kc) {
                i8_mbar_wait(&sm.empty[stage], phase ^ 1);                   // passes immediately on the first lap
                i8_mbar_expect_tx(&sm.full[stage], A_BYTES + B_BYTES);
                const int8_t* ga = g.A + ((long long)kc * g.nrtA + rtA) * (long long)A_BYTES;
                i8_bulk_g2s(&sm.a[stage][0][0][0][0], ga, A_BYTES, &sm.full[stage]);
                const int8_t* gb = g.B + ((long long)kc * g.nrtB + rtB) * (long long)A_BYTES + (long long)hb * I8_TN * 16;
#pragma unroll
                for (int s = 0; s < S; ++s)
#pragma unroll
                    for (int k16 = 0; k16 < 2; ++k16)
                        i8_bulk_g2s(&sm.b[stage][s][k16][0][0], gb + (long long)(s * 2 + k16) * (I8_TM * 16), I8_TN * 16,
                                    &sm.full[stage]);
                if (++stage == I8_STAGES) { stage = 0; phase ^= 1; }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ------------------------------------------------------------------ MMA issuer
        if (lane == 0) {
            // D = S32 (2 << 4), A = B = signed 8 bit (1 << 7, 1 << 10), K-major both, N >> 3 at bit 17, M >> 4 at bit 24
            constexpr uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(I8_TN >> 3) << 17) | ((uint32_t)(I8_TM >> 4) << 24);
            int stage = 0;
            uint32_t phase = 0;
            for (int i = 0; i < nk; ++i) {
                i8_mbar_wait(&sm.full[stage], phase);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t a0 = i8_smem_u32(&sm.a[stage][0][0][0][0]), b0 = i8_smem_u32(&sm.b[stage][0][0][0][0]);
                if (TS) {
                    // the tensor pipe executes tcgen05.cp and tcgen05.mma in issue order: these copies wait for the
                    // previous chunk's MMAs (which read the same columns) and the MMAs below wait for the copies
                    const uint32_t ta = tmem0 + (uint32_t)(S * I8_TN);
#pragma unroll
                    for (int s = 0; s < S; ++s)
                        i8_cp_128x256b(ta + (uint32_t)(s * 8), i8_desc(a0 + s * (2 * I8_TM * 16), I8_TM * 16, 128));
#pragma unroll
                    for (int s = 0; s < S; ++s)
#pragma unroll
                        for (int t2 = 0; t2 < S - s; ++t2) {
                            const uint64_t db = i8_desc(b0 + t2 * (2 * I8_TN * 16), I8_TN * 16, 128);
                            i8_mma_ts(tmem0 + (uint32_t)((s + t2) * I8_TN), ta + (uint32_t)(s * 8), db, idesc, (i > 0 || s > 0) ? 1u : 0u);
                        }
                } else {
#pragma unroll
                    for (int d = 0; d < S; ++d)
#pragma unroll
                        for (int s = 0; s <= d; ++s) {
                            const uint64_t da = i8_desc(a0 + s * (2 * I8_TM * 16), I8_TM * 16, 128);
                            const uint64_t db = i8_desc(b0 + (d - s) * (2 * I8_TN * 16), I8_TN * 16, 128);
                            i8_mma(tmem0 + (uint32_t)(d * I8_TN), da, db, idesc, (i > 0 || s > 0) ? 1u : 0u);
                        }
                }
                i8_commit(&sm.empty[stage]);                                  // stage free once these MMAs have read it
                if (++stage == I8_STAGES) { stage = 0; phase ^= 1; }
            }
            i8_commit(&sm.acc_full);
        }
        __syncwarp();
    } else {
        // ------------------------------------------------------------------ epilogue: thread = one output row
        const int q = warp & 3;                                               // TMEM lane quarter this warp may read
        const int row = t.m0 + q * 32 + lane;
        i8_mbar_wait(&sm.acc_full, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int er = g.ea[row];
        double* crow = g.C + t.c_off + (long long)(q * 32 + lane) * g.ldc;
#pragma unroll 1
        for (int c0 = 0; c0 < I8_TN; c0 += 16) {
            double acc[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) acc[j] = 0.0;
#pragma unroll
            for (int d = S - 1; d >= 0; --d) {                                // smallest weights first
                uint32_t v[16];
                const uint32_t taddr = tmem0 + ((uint32_t)(q * 32) << 16) + (uint32_t)(d * I8_TN + c0);
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                    : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                      "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                    : "r"(taddr)
                    : "memory");
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                const double w = __hiloint2double((1023 - I8_BITS * (d + 2)) << 20, 0);     // 2^(-7 (d + 2))
#pragma unroll
                for (int j = 0; j < 16; ++j) acc[j] = fma((double)(int32_t)v[j], w, acc[j]);
            }
#pragma unroll
            for (int j = 0; j < 16; j += 2) {
                const int e0 = er + g.eb[t.n0 + c0 + j], e1 = er + g.eb[t.n0 + c0 + j + 1];
                double x0 = g.alpha * ldexp(acc[j], e0), x1 = g.alpha * ldexp(acc[j + 1], e1);
                double2* p = reinterpret_cast<double2*>(crow + c0 + j);
                if (g.beta != 0.0) { const double2 o = *p; x0 = fma(g.beta, o.x, x0); x1 = fma(g.beta, o.y, x1); }
                *p = make_double2(x0, x1);
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem0), "n"(TMEM_COLS) : "memory");
}

// ------------------------------------------------------------------ wide variant: 128 x 128 tiles, two passes
// The 128 x 64 kernel above keeps all S anti-diagonal accumulators in tensor memory at once, which limits N to 64 and makes
// every MMA read 6 KB of shared memory for 32 cycles of math.  Here a launch only computes the anti-diagonals
// d in [DLO, DHI] (at most 4: 4 x 128 = 512 TMEM columns), so N = 128 fits: an MMA reads 8 KB for 64 cycles (128 B/cycle
// instead of 192).  A product takes two launches -- d = 0..3 (10 digit pairs, planes 0..3) and d = 4..S-1 (18 pairs at
// S = 7, all planes), the second accumulating onto the first's output -- and streams the operands twice from L2, which
// has the headroom (the narrow kernel ran it at ~18 %).
template <int DLO, int DHI>
struct I8Wide {
    static constexpr int NACC = DHI - DLO + 1;
    static constexpr int PLANES = DHI + 1;                        // digit planes 0 .. DHI of both operands are needed
    static constexpr uint32_t STAGE_BYTES = PLANES * 2 * (2 * I8_TM * 16);     // A planes + B planes (128 rows each)
    static constexpr int STAGES = (200 * 1024) / STAGE_BYTES >= 4 ? 4 : (200 * 1024) / STAGE_BYTES;
    static constexpr uint32_t SMEM = STAGES * STAGE_BYTES + 256;
    static constexpr uint32_t TMEM_COLS = NACC * 128 > 256 ? 512 : 256;
};

template <int DLO, int DHI>
__global__ void __launch_bounds__(192, 1) i8_gemm_wide_kernel(I8Args g, int splanes) {
    using W = I8Wide<DLO, DHI>;
    constexpr int TN = 128;
    extern __shared__ __align__(128) unsigned char i8_raw[];
    unsigned char* stages = i8_raw;
    uint64_t* full = reinterpret_cast<uint64_t*>(i8_raw + W::STAGES * W::STAGE_BYTES);
    uint64_t* empty = full + W::STAGES;
    uint64_t* acc_full = empty + W::STAGES;
    uint32_t* tmem_base = reinterpret_cast<uint32_t*>(acc_full + 1);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const I8Tile t = g.tiles[blockIdx.x];
    const int nk = t.kc1 - t.kc0;
    constexpr uint32_t PLANE_BYTES = 2 * I8_TM * 16;              // 4 KB: one digit plane of 128 rows x 32 K bytes
    constexpr uint32_t OP_BYTES = W::PLANES * PLANE_BYTES;

    if (tid == 0) {
        for (int i = 0; i < W::STAGES; ++i) { i8_mbar_init(&full[i], 1); i8_mbar_init(&empty[i], 1); }
        i8_mbar_init(acc_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        __syncwarp();
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(i8_smem_u32(tmem_base)), "n"(W::TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem0 = *tmem_base;

    if (warp == 0) {
        if (lane == 0) {
            // a (K chunk, row tile) block of an operand holds all `splanes` planes; planes 0 .. DHI are its first OP_BYTES
            const long long blockA = (long long)splanes * PLANE_BYTES;
            const int rtA = t.m0 / I8_TM, rtB = t.n0 / I8_TM;
            int stage = 0;
            uint32_t phase = 0;
            for (int kc = t.kc0; kc < t.kc1; ++kc) {
                i8_mbar_wait(&empty[stage], phase ^ 1);
                i8_mbar_expect_tx(&full[stage], 2 * OP_BYTES);
                unsigned char* sp = stages + (size_t)stage * W::STAGE_BYTES;
                i8_bulk_g2s(sp, g.A + ((long long)kc * g.nrtA + rtA) * blockA, OP_BYTES, &full[stage]);
                i8_bulk_g2s(sp + OP_BYTES, g.B + ((long long)kc * g.nrtB + rtB) * blockA, OP_BYTES, &full[stage]);
                if (++stage == W::STAGES) { stage = 0; phase ^= 1; }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(TN >> 3) << 17) | ((uint32_t)(I8_TM >> 4) << 24);
            int stage = 0;
            uint32_t phase = 0;
            for (int i = 0; i < nk; ++i) {
                i8_mbar_wait(&full[stage], phase);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t a0 = i8_smem_u32(stages + (size_t)stage * W::STAGE_BYTES), b0 = a0 + OP_BYTES;
#pragma unroll
                for (int d = DLO; d <= DHI; ++d)
#pragma unroll
                    for (int s = 0; s <= d; ++s) {
                        const uint64_t da = i8_desc(a0 + s * PLANE_BYTES, I8_TM * 16, 128);
                        const uint64_t db = i8_desc(b0 + (d - s) * PLANE_BYTES, I8_TM * 16, 128);
                        i8_mma(tmem0 + (uint32_t)((d - DLO) * TN), da, db, idesc, (i > 0 || s > 0) ? 1u : 0u);
                    }
                i8_commit(&empty[stage]);
                if (++stage == W::STAGES) { stage = 0; phase ^= 1; }
            }
            i8_commit(acc_full);
        }
        __syncwarp();
    } else {
        const int q = warp & 3;
        const int row = t.m0 + q * 32 + lane;
        i8_mbar_wait(acc_full, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int er = g.ea[row];
        double* crow = g.C + t.c_off + (long long)(q * 32 + lane) * g.ldc;
#pragma unroll 1
        for (int c0 = 0; c0 < TN; c0 += 16) {
            double acc[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) acc[j] = 0.0;
#pragma unroll
            for (int d = DHI; d >= DLO; --d) {
                uint32_t v[16];
                const uint32_t taddr = tmem0 + ((uint32_t)(q * 32) << 16) + (uint32_t)((d - DLO) * TN + c0);
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                    : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                      "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                    : "r"(taddr)
                    : "memory");
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                const double w = __hiloint2double((1023 - I8_BITS * (d + 2)) << 20, 0);
#pragma unroll
                for (int j = 0; j < 16; ++j) acc[j] = fma((double)(int32_t)v[j], w, acc[j]);
            }
#pragma unroll
            for (int j = 0; j < 16; j += 2) {
                const int e0 = er + g.eb[t.n0 + c0 + j], e1 = er + g.eb[t.n0 + c0 + j + 1];
                double x0 = g.alpha * ldexp(acc[j], e0), x1 = g.alpha * ldexp(acc[j + 1], e1);
                double2* p = reinterpret_cast<double2*>(crow + c0 + j);
                if (g.beta != 0.0) { const double2 o = *p; x0 = fma(g.beta, o.x, x0); x1 = fma(g.beta, o.y, x1); }
                *p = make_double2(x0, x1);
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem0), "n"(W::TMEM_COLS) : "memory");
}

// ------------------------------------------------------------------ host side
struct I8Operand {                 // a sliced operand in device memory
    int8_t* digits = nullptr; size_t digits_cap = 0;
    int* ex_bits = nullptr; int* ex = nullptr; size_t rows_cap = 0;
    int R = 0, K = 0, nrt = 0, nkc = 0;
};
struct I8List { size_t first = 0, count = 0; };
struct I8Lists {                                 // the tile lists of one tile width (64 or 128 columns)
    I8List kinv;                                 // K^-1 = L^-T L^-1
    std::vector<I8List> lvl_a, lvl_b;            // per doubling level (index = log2 of the block size in 64-blocks): the two GEMMs
    // trailing update of the blocked Cholesky, coordinates relative to the first row below the super-panel: master lists
    // ordered by row tile, so that the list for a smaller trailing matrix is a prefix (prefix counts per row-tile count)
    I8List syrk_a, syrk_b;
    std::vector<size_t> syrk_a_count, syrk_b_count;
    // recursive Cholesky + inverse (i8_blk_*): the four products of one 2h x 2h diagonal block, coordinates relative to the
    // block (index j: h = leaf * 2^j): A21 X11^T, A22 -= L21 L21^T, T = L21 X11, X21 = -X22 T
    std::vector<I8List> blk_trsm, blk_syrk, blk_pa, blk_pb;
};
struct I8Plan {
    I8Operand opA, opB, opP;                 // opP: the 1024-column super-panel of the blocked Cholesky
    I8Operand opC, opD;                      // recursive scheme: operands of products that overlap the second half's recursion
    I8Tile* tiles = nullptr; size_t tiles_cap = 0;
    std::vector<I8Tile> host_tiles;          // all tile lists back to back
    I8Lists L[2];                            // [0]: 128 x 64 tiles (one-pass kernel), [1]: 128 x 128 tiles (two-pass kernel)
    long long tiles_key = -1;
    int64_t Np = 0; long long ld = 0; int S = 0;     // what the lists were built for
    int64_t leaf = 0;                                // leaf of the recursive scheme the blk_* lists were built for (0: none)
    bool ready(int64_t Np_, long long ld_, int S_) const { return tiles_key >= 0 && Np == Np_ && ld == ld_ && S == S_; }
};

static cudaError_t i8_reserve(I8Operand& op, int R, int K, int S) {
    const int nrt = (R + I8_TM - 1) / I8_TM, nkc = (K + I8_KC - 1) / I8_KC;
    const size_t need = (size_t)nkc * nrt * S * 2 * I8_TM * 16;
    cudaError_t e;
    if (need > op.digits_cap) {
        if (op.digits) cudaFree(op.digits);
        op.digits = nullptr; op.digits_cap = 0;
        if ((e = cudaMalloc(&op.digits, need)) != cudaSuccess) return e;
        op.digits_cap = need;
    }
    if ((size_t)nrt * I8_TM > op.rows_cap) {
        if (op.ex_bits) cudaFree(op.ex_bits);
        if (op.ex) cudaFree(op.ex);
        op.ex_bits = op.ex = nullptr; op.rows_cap = 0;
        if ((e = cudaMalloc(&op.ex_bits, (size_t)nrt * I8_TM * 4)) != cudaSuccess) return e;
        if ((e = cudaMalloc(&op.ex, (size_t)nrt * I8_TM * 4)) != cudaSuccess) return e;
        op.rows_cap = (size_t)nrt * I8_TM;
    }
    op.R = R; op.K = K; op.nrt = nrt; op.nkc = nkc;
    return cudaSuccess;
}

// Slice an operand of R rows (batch groups of R / batch rows, see i8_rowmax_kernel), K deep.
static cudaError_t i8_slice(I8Operand& op, const double* src, long long rs, long long cs, long long bstride, int batch,
                            int R, int K, int S, int tri, cudaStream_t st) {
    cudaError_t e = i8_reserve(op, R, K, S);
    if (e != cudaSuccess) return e;
    if ((e = cudaMemsetAsync(op.ex_bits, 0, (size_t)op.nrt * I8_TM * 4, st)) != cudaSuccess) return e;
    const int ksplit = std::max(1, std::min(32, K / 256));
    const int Rb = R / std::max(1, batch);
    i8_rowmax_kernel<<<dim3((R + 63) / 64, ksplit), 256, 0, st>>>(src, rs, cs, bstride, R, Rb, K, tri, op.ex_bits);
    i8_exponent_kernel<<<(op.nrt * I8_TM + 255) / 256, 256, 0, st>>>(op.ex_bits, op.ex, op.nrt * I8_TM);
    i8_slice_tiled_kernel<<<dim3(op.nkc, op.nrt), 256, 0, st>>>(src, rs, cs, bstride, R, Rb, K, S, tri, op.ex_bits, op.digits, op.nrt);
    MOGP_COUNT(3);
    return cudaGetLastError();
}

// 1: A planes through tensor memory (S = 7 only: 8 planes do not fit beside 8 accumulators), 0: both operands from shared memory
int g_i8_ts = std::getenv("MOGP_I8_TS") ? std::atoi(std::getenv("MOGP_I8_TS")) : 0;
// 128 x 128 tiles in two passes over the anti-diagonals (S = 7) instead of 128 x 64 tiles in one pass.  Measured
// (profiles/r02_i8mm_variants.log): 8192^3 99.5 vs 82.9 TFLOP/s fp64-equivalent, but slower when K is short (K = 1024:
// 52.4 vs 59.7) because a product is two launches.  0: never; 1: the K^-1 product for N >= 8192 (K up to N; at N = 4096 the
// one-pass kernel is faster: 0.466 vs 0.491 ms); 2: also the levels of the triangular inverse with block size >= 4096
// (cfg3 step 14.19 -> 14.06 ms); 3: everything (self-test).
int g_i8_wide = std::getenv("MOGP_I8_WIDE") ? std::atoi(std::getenv("MOGP_I8_WIDE")) : 2;
enum { I8_USE_KINV = 1, I8_USE_TRTRI_TOP = 2, I8_USE_OTHER = 3 };
static int i8_width(int S, int use) { return (S == 7 && g_i8_wide >= use) ? 1 : 0; }

// tiles_dev: a list built for the tile width `width` (0: 128 x 64, 1: 128 x 128)
static cudaError_t i8_launch(const I8Operand& A, const I8Operand& B, const I8Tile* tiles_dev, int ntiles, int S, double alpha,
                             double beta, double* C, long long ldc, cudaStream_t st, int width) {
    static PerDeviceOnce once;
    if (OnceGuard og{once}; og.needed()) {
        cudaError_t e = cudaFuncSetAttribute(i8_gemm_tiles_kernel<7, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(I8Smem));
        if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(i8_gemm_tiles_kernel<7, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(I8Smem));
        if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(i8_gemm_tiles_kernel<8, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(I8Smem));
        if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(i8_gemm_wide_kernel<0, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)I8Wide<0, 3>::SMEM);
        if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(i8_gemm_wide_kernel<4, 6>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)I8Wide<4, 6>::SMEM);
        if (e != cudaSuccess) return e;
    }
    if (ntiles <= 0) return cudaSuccess;
    I8Args g{};
    g.A = A.digits; g.ea = A.ex; g.nrtA = A.nrt;
    g.B = B.digits; g.eb = B.ex; g.nrtB = B.nrt;
    g.tiles = tiles_dev; g.S = S; g.alpha = alpha; g.beta = beta; g.C = C; g.ldc = ldc;
    if (width == 1) {
        i8_gemm_wide_kernel<0, 3><<<ntiles, 192, I8Wide<0, 3>::SMEM, st>>>(g, S);
        g.beta = 1.0;                                   // the second pass accumulates onto the first
        i8_gemm_wide_kernel<4, 6><<<ntiles, 192, I8Wide<4, 6>::SMEM, st>>>(g, S);
        MOGP_COUNT(2);
        return cudaGetLastError();
    }
    if (S == 7 && g_i8_ts) i8_gemm_tiles_kernel<7, true><<<ntiles, 192, sizeof(I8Smem), st>>>(g);
    else if (S == 7) i8_gemm_tiles_kernel<7, false><<<ntiles, 192, sizeof(I8Smem), st>>>(g);
    else if (S == 8) i8_gemm_tiles_kernel<8, false><<<ntiles, 192, sizeof(I8Smem), st>>>(g);
    else return cudaErrorInvalidValue;
    MOGP_COUNT(1);
    return cudaGetLastError();
}

static cudaError_t i8_upload_tiles(I8Plan& p, cudaStream_t st) {
    const size_t n = p.host_tiles.size();
    if (n > p.tiles_cap) {
        if (p.tiles) cudaFree(p.tiles);
        p.tiles = nullptr; p.tiles_cap = 0;
        cudaError_t e = cudaMalloc(&p.tiles, n * sizeof(I8Tile));
        if (e != cudaSuccess) return e;
        p.tiles_cap = n;
    }
    return cudaMemcpyAsync(p.tiles, p.host_tiles.data(), n * sizeof(I8Tile), cudaMemcpyHostToDevice, st);
}

I8Plan* i8_plan_create() { return new I8Plan(); }
void i8_plan_destroy(I8Plan* p) {
    if (!p) return;
    for (I8Operand* op : {&p->opA, &p->opB, &p->opP, &p->opC, &p->opD}) {
        if (op->digits) cudaFree(op->digits);
        if (op->ex_bits) cudaFree(op->ex_bits);
        if (op->ex) cudaFree(op->ex);
    }
    if (p->tiles) cudaFree(p->tiles);
    delete p;
}

// Smallest block size (rows) of a doubling level of the triangular inverse that runs on the int8 pipe
long long g_i8_trtri_min = 1024;

static bool i8_level_ok(int64_t Np, int64_t S) {
    return g_i8_trtri_min > 0 && S >= g_i8_trtri_min && S % I8_TM == 0 && Np % (2 * S) == 0;     // no ragged last pair
}

// Host-side preparation (allocations, tile lists) for a given padded size: never inside graph capture.
cudaError_t i8_prepare(I8Plan* p, int64_t Np, long long ld, int S, cudaStream_t st, bool* changed) {
    if (changed) *changed = false;
    cudaError_t e = i8_reserve(p->opA, (int)Np, (int)Np, S);
    if (e != cudaSuccess) return e;
    if ((e = i8_reserve(p->opB, (int)Np, (int)(Np / 2), S)) != cudaSuccess) return e;
    const int64_t leaf = rchol_leaf_for(Np);
    const long long key = (Np * 16 + S) * 65536 + g_i8_trtri_min % 65536 + ld * 1000003ll + leaf * 7919ll;
    if (p->tiles_key == key) return cudaSuccess;
    if (changed) *changed = true;            // captured graphs that replay the old lists must be re-captured
    if ((e = cudaDeviceSynchronize()) != cudaSuccess) return e;      // nothing may still read the lists being replaced
    if ((e = i8_reserve(p->opP, (int)Np, I8_PANEL, S)) != cudaSuccess) return e;
    if (leaf > 0 && Np >= 4 * leaf) {        // (blocks above the lowest recursion level exist from 4 leaves on)
        if ((e = i8_reserve(p->opC, (int)(Np / 2), (int)(Np / 2), S)) != cudaSuccess) return e;
        if ((e = i8_reserve(p->opD, (int)(Np / 2), (int)(Np / 2), S)) != cudaSuccess) return e;
    }
    p->host_tiles.clear();
    const int nkc = (int)(Np / I8_KC);
    for (int w = 0; w < 2; ++w) {
        I8Lists& L = p->L[w];
        const int TN = w == 0 ? I8_TN : 128;
        std::vector<I8Tile>& T = p->host_tiles;
        L.lvl_a.assign(32, I8List());
        L.lvl_b.assign(32, I8List());
        L.kinv.first = T.size();
        for (int ti = 0; ti < (int)(Np / I8_TM); ++ti)                         // longest K ranges first
            for (int tj = 0; tj * TN < (ti + 1) * I8_TM; ++tj)
                T.push_back({ti * I8_TM, tj * TN, ti * I8_TM / I8_KC, nkc, (long long)ti * I8_TM * ld + (long long)tj * TN});
        L.kinv.count = T.size() - L.kinv.first;
        // doubling levels of the triangular inverse: pairs of adjacent S x S diagonal blocks (A, B) at offset o = 2 S pair:
        //   GEMM a: T = L_BA Linv_AA        operand A rows = rows of L_BA, operand B rows = columns of Linv_AA (nonzero k >= n)
        //   GEMM b: Linv_BA = -Linv_BB T    operand A rows = rows of Linv_BB (nonzero k <= m), operand B rows = columns of T
        // both operands of a level are the stacked blocks of all pairs (row index = pair * S + local row).
        int lev = 0;
        for (int64_t S_ = 64; S_ < Np; S_ *= 2, ++lev) {
            if (!i8_level_ok(Np, S_)) continue;
            const int npair = (int)(Np / (2 * S_)), Sn = (int)S_;
            L.lvl_a[lev].first = T.size();
            for (int tn = 0; tn < Sn / TN; ++tn)                                // longest K ranges (small n) first
                for (int pr = 0; pr < npair; ++pr)
                    for (int tm = 0; tm < Sn / I8_TM; ++tm) {
                        const long long o = (long long)pr * 2 * S_;
                        T.push_back({pr * Sn + tm * I8_TM, pr * Sn + tn * TN, tn * TN / I8_KC, Sn / I8_KC,
                                     (o + S_ + (long long)tm * I8_TM) * ld + o + (long long)tn * TN});
                    }
            L.lvl_a[lev].count = T.size() - L.lvl_a[lev].first;
            L.lvl_b[lev].first = T.size();
            for (int tm = Sn / I8_TM - 1; tm >= 0; --tm)                        // longest K ranges (large m) first
                for (int pr = 0; pr < npair; ++pr)
                    for (int tn = 0; tn < Sn / TN; ++tn) {
                        const long long o = (long long)pr * 2 * S_;
                        T.push_back({pr * Sn + tm * I8_TM, pr * Sn + tn * TN, 0, (tm + 1) * I8_TM / I8_KC,
                                     (o + S_ + (long long)tm * I8_TM) * ld + o + (long long)tn * TN});
                    }
            L.lvl_b[lev].count = T.size() - L.lvl_b[lev].first;
        }
        // recursive Cholesky + inverse: one 2h x 2h diagonal block [[L11, .], [A21, A22]] with X11 = L11^-1, X22 = L22^-1 (all h x h);
        // output coordinates relative to the h x h block each product writes
        L.blk_trsm.assign(32, I8List()); L.blk_syrk.assign(32, I8List());
        L.blk_pa.assign(32, I8List()); L.blk_pb.assign(32, I8List());
        lev = 0;
        for (int64_t h = leaf; leaf > 0 && 2 * h <= Np; h *= 2, ++lev) {
            const int hn = (int)h, ntm = hn / I8_TM, ntn = hn / TN, hk = hn / I8_KC;
            // L21 = A21 X11^T: operand A rows = rows of A21, operand B rows = rows of X11 (nonzero k <= n); longest K (large n) first
            L.blk_trsm[lev].first = T.size();
            for (int tn = ntn - 1; tn >= 0; --tn)
                for (int tm = 0; tm < ntm; ++tm)
                    T.push_back({tm * I8_TM, tn * TN, 0, (tn + 1) * TN / I8_KC, (long long)tm * I8_TM * ld + (long long)tn * TN});
            L.blk_trsm[lev].count = T.size() - L.blk_trsm[lev].first;
            // A22 -= L21 L21^T (lower tiles)
            L.blk_syrk[lev].first = T.size();
            for (int tm = 0; tm < ntm; ++tm)
                for (int tn = 0; tn * TN < (tm + 1) * I8_TM; ++tn)
                    T.push_back({tm * I8_TM, tn * TN, 0, hk, (long long)tm * I8_TM * ld + (long long)tn * TN});
            L.blk_syrk[lev].count = T.size() - L.blk_syrk[lev].first;
            // T = L21 X11: operand B rows = columns of X11 (nonzero k >= n); longest K (small n) first
            L.blk_pa[lev].first = T.size();
            for (int tn = 0; tn < ntn; ++tn)
                for (int tm = 0; tm < ntm; ++tm)
                    T.push_back({tm * I8_TM, tn * TN, tn * TN / I8_KC, hk, (long long)tm * I8_TM * ld + (long long)tn * TN});
            L.blk_pa[lev].count = T.size() - L.blk_pa[lev].first;
            // X21 = -X22 T: operand A rows = rows of X22 (nonzero k <= m), operand B rows = columns of T; longest K (large m) first
            L.blk_pb[lev].first = T.size();
            for (int tm = ntm - 1; tm >= 0; --tm)
                for (int tn = 0; tn < ntn; ++tn)
                    T.push_back({tm * I8_TM, tn * TN, 0, (tm + 1) * I8_TM / I8_KC, (long long)tm * I8_TM * ld + (long long)tn * TN});
            L.blk_pb[lev].count = T.size() - L.blk_pb[lev].first;
        }
        // trailing updates C -= P P^T (K = I8_PANEL): (a) the I8_PANEL columns of the next super-panel, (b) the rest (lower tiles)
        const int ntm = (int)(Np / I8_TM);
        L.syrk_a.first = T.size();
        L.syrk_a_count.assign(ntm + 1, 0);
        for (int tm = 0; tm < ntm; ++tm) {
            for (int tn = 0; tn < I8_PANEL / TN && tn * TN < (tm + 1) * I8_TM; ++tn)
                T.push_back({tm * I8_TM, tn * TN, 0, I8_PANEL / I8_KC, (long long)tm * I8_TM * ld + (long long)tn * TN});
            L.syrk_a_count[tm + 1] = T.size() - L.syrk_a.first;
        }
        L.syrk_a.count = T.size() - L.syrk_a.first;
        L.syrk_b.first = T.size();
        L.syrk_b_count.assign(ntm + 1, 0);
        for (int tm = 0; tm < ntm; ++tm) {
            for (int tn = I8_PANEL / TN; tn * TN < (tm + 1) * I8_TM; ++tn)
                T.push_back({tm * I8_TM, tn * TN, 0, I8_PANEL / I8_KC, (long long)tm * I8_TM * ld + (long long)tn * TN});
            L.syrk_b_count[tm + 1] = T.size() - L.syrk_b.first;
        }
        L.syrk_b.count = T.size() - L.syrk_b.first;
    }
    e = i8_upload_tiles(*p, st);
    if (e != cudaSuccess) return e;
    p->tiles_key = key;
    p->Np = Np; p->ld = ld; p->S = S; p->leaf = leaf;
    return cudaStreamSynchronize(st);
}

// Rank-1024 trailing update of the blocked Cholesky on the int8 pipe: with P = A[r0 .. Np) x [k0, k0 + 1024) (a finished
// super-panel below its diagonal block), part 0 slices P and updates the next super-panel's 1024 columns
// A[r0.., r0 .. r0+1024) -= P P^T (lower tiles), part 1 updates the rest A[r0.., r0+1024..) -= P P^T (lower tiles).
// r0 a multiple of 1024.  Pure enqueue; cudaErrorNotSupported when this size is not prepared.
cudaError_t i8_syrk_update(I8Plan* p, double* A, long long ld, int64_t r0, int64_t k0, int64_t Np, int part, int S,
                           cudaStream_t st) {
    if (!p || (r0 % I8_PANEL) != 0) return cudaErrorNotSupported;
    const int width = i8_width(S, I8_USE_OTHER);
    const I8Lists& L = p->L[width];
    if (!p->ready(Np, ld, S) || L.syrk_a.count == 0) return cudaErrorInvalidValue;
    const int M = (int)(Np - r0), ntm = M / I8_TM;
    cudaError_t e;
    if (part == 0) {
        if ((e = i8_slice(p->opP, A + r0 * ld + k0, ld, 1, 0, 1, M, I8_PANEL, S, 0, st)) != cudaSuccess) return e;
        return i8_launch(p->opP, p->opP, p->tiles + L.syrk_a.first, (int)L.syrk_a_count[ntm], S, -1.0, 1.0,
                         A + r0 * (ld + 1), ld, st, width);
    }
    return i8_launch(p->opP, p->opP, p->tiles + L.syrk_b.first, (int)L.syrk_b_count[ntm], S, -1.0, 1.0, A + r0 * (ld + 1), ld, st, width);
}

// W(lower tiles) = Linv^T Linv with Linv lower triangular (zero above the diagonal): K^-1 of the factorised matrix.
// Operand row i = column i of Linv (element (i, k) = Linv[k * ld + i]), nonzero for k >= i: K range of output tile
// (ti, tj), tj <= ti, starts at the tile's first row.  Pure enqueue (capturable) after i8_prepare.
cudaError_t i8_kinv(I8Plan* p, const double* Linv, double* W, int64_t Np, long long ld, int S, cudaStream_t st) {
    const int width = (Np >= 8192 || g_i8_wide >= I8_USE_OTHER) ? i8_width(S, I8_USE_KINV) : 0;
    const I8Lists& L = p->L[width];
    if (!p->ready(Np, ld, S) || L.kinv.count == 0) return cudaErrorInvalidValue;      // i8_prepare was not run for this size
    cudaError_t e = i8_slice(p->opA, Linv, 1, ld, 0, 1, (int)Np, (int)Np, S, 1, st);
    if (e != cudaSuccess) return e;
    return i8_launch(p->opA, p->opA, p->tiles + L.kinv.first, (int)L.kinv.count, S, 1.0, 0.0, W, ld, st, width);
}

// One doubling level (block size S_ rows) of Linv = L^-1 on the int8 pipe; returns cudaErrorNotSupported when the level
// is not eligible (the caller then runs the DMMA GEMMs).  scratch receives T = L_BA Linv_AA (as in trtri_padded).
cudaError_t i8_trtri_level(I8Plan* p, const double* L, double* Linv, double* scratch, int64_t Np, long long ld, int64_t S_,
                           int S, cudaStream_t st) {
    if (!p || !i8_level_ok(Np, S_)) return cudaErrorNotSupported;
    if (!p->ready(Np, ld, S)) return cudaErrorInvalidValue;
    int lev = 0;
    for (int64_t x = 64; x < S_; x *= 2) ++lev;
    const int width = i8_width(S, S_ >= 4096 ? I8_USE_TRTRI_TOP : I8_USE_OTHER);
    const I8Lists& TL = p->L[width];
    if (TL.lvl_a[lev].count == 0) return cudaErrorNotSupported;
    const int npair = (int)(Np / (2 * S_)), R = (int)(npair * S_), K = (int)S_;
    const long long bs = 2 * S_ * (ld + 1);
    cudaError_t e;
    // GEMM a
    if ((e = i8_slice(p->opA, L + S_ * ld, ld, 1, bs, npair, R, K, S, 0, st)) != cudaSuccess) return e;
    if ((e = i8_slice(p->opB, Linv, 1, ld, bs, npair, R, K, S, 1, st)) != cudaSuccess) return e;
    if ((e = i8_launch(p->opA, p->opB, p->tiles + TL.lvl_a[lev].first, (int)TL.lvl_a[lev].count, S, 1.0, 0.0, scratch, ld, st, width)) != cudaSuccess) return e;
    // GEMM b
    if ((e = i8_slice(p->opA, Linv + S_ * ld + S_, ld, 1, bs, npair, R, K, S, 2, st)) != cudaSuccess) return e;
    if ((e = i8_slice(p->opB, scratch + S_ * ld, 1, ld, bs, npair, R, K, S, 0, st)) != cudaSuccess) return e;
    return i8_launch(p->opA, p->opB, p->tiles + TL.lvl_b[lev].first, (int)TL.lvl_b[lev].count, S, -1.0, 0.0, Linv, ld, st, width);
}

// ------------------------------------------------------------------ recursive Cholesky + inverse: the products of one block
// The 2h x 2h diagonal block at origin o of A (factor, in place) and Linv: A = [[L11, .], [A21, A22]], X11 = L11^-1 is complete.
//   i8_blk_first : A21 <- L21 = A21 X11^T;  A22 -= L21 L21^T (lower);  if want_inv: scratch21 <- T = L21 X11
//   i8_blk_second: (X22 = L22^-1 complete)  Linv21 <- X21 = -X22 T
// Replaces the panel-by-panel trailing updates of the blocked sweep above a leaf size by products whose K is the block size,
// i.e. almost all of the N^3/3 + N^3/3 flops of factor + inverse run as large int8-pipe GEMMs.  Pure enqueue after i8_prepare.
static int i8_blk_level(const I8Plan* p, int64_t h) { int lev = 0; for (int64_t x = p->leaf; x < h; x *= 2) ++lev; return lev; }
bool i8_blk_ok(const I8Plan* p, int64_t Np, long long ld, int S, int64_t leaf) {
    if (!p || !p->ready(Np, ld, S) || leaf <= 0 || p->leaf != leaf || leaf % I8_TM != 0 || 2 * leaf > Np) return false;
    return p->L[0].blk_trsm[0].count > 0;
}
// as != NULL: the products the second half's recursion does not wait for run on as->side concurrently with it -- T = L21 X11
// always, and, when the second half is itself a recursion (as->split_rows = rows of its first leaf), the part of the A22 update
// below those rows; their operands then live in opC / opD (as->own_ops), which nothing else touches until i8_blk_second.
// The caller waits for as->ev_rest before the second half reads A22 below split_rows and for as->ev_T before i8_blk_second.
cudaError_t i8_blk_first(I8Plan* p, double* A, const double* Linv, double* scratch, long long ld, int64_t o, int64_t h, int want_inv,
                         int S, cudaStream_t st, const I8BlkAsync* as) {
    const int lev = i8_blk_level(p, h), hn = (int)h;
    const int width = i8_width(S, h >= 4096 ? I8_USE_TRTRI_TOP : I8_USE_OTHER);
    const I8Lists& TL = p->L[width];
    const int TN = width == 0 ? I8_TN : 128;
    double* A21 = A + (o + h) * ld + o;
    double* A22 = A + (o + h) * (ld + 1);
    const bool own = as && as->own_ops && p->opC.digits_cap > 0 && p->opD.digits_cap > 0;
    I8Operand& opL = own ? p->opC : p->opA;          // L21
    I8Operand& opX = own ? p->opD : p->opB;          // columns of X11
    cudaError_t e;
    if ((e = i8_slice(p->opA, A21, ld, 1, 0, 1, hn, hn, S, 0, st)) != cudaSuccess) return e;
    if ((e = i8_slice(p->opB, Linv + o * (ld + 1), ld, 1, 0, 1, hn, hn, S, 2, st)) != cudaSuccess) return e;
    if ((e = i8_launch(p->opA, p->opB, p->tiles + TL.blk_trsm[lev].first, (int)TL.blk_trsm[lev].count, S, 1.0, 0.0, A21, ld, st, width)) != cudaSuccess) return e;
    if ((e = i8_slice(opL, A21, ld, 1, 0, 1, hn, hn, S, 0, st)) != cudaSuccess) return e;
    // A22 -= L21 L21^T: the list is ordered by row tile, so the rows the second half needs first are a prefix
    int n_all = (int)TL.blk_syrk[lev].count, n_first = n_all;
    if (as && as->split_rows > 0 && as->split_rows < h) {
        n_first = 0;
        for (int tm = 0; tm < (int)(as->split_rows / I8_TM); ++tm) n_first += ((tm + 1) * I8_TM + TN - 1) / TN;
    }
    if ((e = i8_launch(opL, opL, p->tiles + TL.blk_syrk[lev].first, n_first, S, -1.0, 1.0, A22, ld, st, width)) != cudaSuccess) return e;
    const bool fork = as && (n_first < n_all || want_inv);
    cudaStream_t s2 = fork ? as->side : st;
    if (fork) {
        if ((e = cudaEventRecord(as->ev_fork, st)) != cudaSuccess) return e;
        if ((e = cudaStreamWaitEvent(s2, as->ev_fork, 0)) != cudaSuccess) return e;
    }
    if (n_first < n_all) {
        if ((e = i8_launch(opL, opL, p->tiles + TL.blk_syrk[lev].first + n_first, n_all - n_first, S, -1.0, 1.0, A22, ld, s2, width)) != cudaSuccess) return e;
        if (fork && (e = cudaEventRecord(as->ev_rest, s2)) != cudaSuccess) return e;
    }
    if (want_inv) {
        if ((e = i8_slice(opX, Linv + o * (ld + 1), 1, ld, 0, 1, hn, hn, S, 1, s2)) != cudaSuccess) return e;
        if ((e = i8_launch(opL, opX, p->tiles + TL.blk_pa[lev].first, (int)TL.blk_pa[lev].count, S, 1.0, 0.0, scratch + (o + h) * ld + o, ld, s2, width)) != cudaSuccess) return e;
    }
    if (fork && (e = cudaEventRecord(as->ev_T, s2)) != cudaSuccess) return e;      // everything on the side stream is done
    return cudaSuccess;
}
cudaError_t i8_blk_second(I8Plan* p, double* Linv, const double* scratch, long long ld, int64_t o, int64_t h, int S, cudaStream_t st) {
    const int lev = i8_blk_level(p, h), hn = (int)h;
    const int width = i8_width(S, h >= 4096 ? I8_USE_TRTRI_TOP : I8_USE_OTHER);
    const I8Lists& TL = p->L[width];
    cudaError_t e;
    if ((e = i8_slice(p->opA, Linv + (o + h) * (ld + 1), ld, 1, 0, 1, hn, hn, S, 2, st)) != cudaSuccess) return e;
    if ((e = i8_slice(p->opB, scratch + (o + h) * ld + o, 1, ld, 0, 1, hn, hn, S, 0, st)) != cudaSuccess) return e;
    return i8_launch(p->opA, p->opB, p->tiles + TL.blk_pb[lev].first, (int)TL.blk_pb[lev].count, S, -1.0, 0.0, Linv + (o + h) * ld + o, ld, st, width);
}

// ------------------------------------------------------------------ self-test hooks (tests, tools/gpu_diag.py)
// General NT product against the DMMA GEMM on random data: out[0] = max |C_i8 - C_dmma| / max |C_dmma|, out[1] = ms of
// the int8 path (slicing included), out[2] = ms of the MMA kernel alone, out[3] = ms of the DMMA GEMM.
__global__ void i8t_fill_kernel(double* x, long long n, unsigned seed) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    unsigned h = (unsigned)i * 2654435761u ^ seed;
    h ^= h >> 15; h *= 2246822519u; h ^= h >> 13; h *= 3266489917u; h ^= h >> 16;
    const double u = (double)h / 4294967296.0 - 0.5;
    x[i] = u * exp2((double)((int)(h % 13) - 6));
}
__global__ void i8t_maxdiff_kernel(const double* a, const double* b, long long n, double* out) {
    __shared__ double sd[256], sm_[256];
    double d = 0.0, m = 0.0;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const double x = fabs(a[i] - b[i]);
        d = (x == x) ? fmax(d, x) : 1e300;
        m = fmax(m, fabs(b[i]));
    }
    sd[threadIdx.x] = d; sm_[threadIdx.x] = m;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) { sd[threadIdx.x] = fmax(sd[threadIdx.x], sd[threadIdx.x + o]); sm_[threadIdx.x] = fmax(sm_[threadIdx.x], sm_[threadIdx.x + o]); }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        atomicMax(reinterpret_cast<unsigned long long*>(out), (unsigned long long)__double_as_longlong(sd[0]));
        atomicMax(reinterpret_cast<unsigned long long*>(out + 1), (unsigned long long)__double_as_longlong(sm_[0]));
    }
}

extern "C" int mogp_i8_selftest(int M, int N, int K, int S, double* out_host /*4*/) {
    if (M % I8_TM || N % I8_TM || K % I8_KC || (S != 7 && S != 8)) return -1;
    double *A = nullptr, *B = nullptr, *C1 = nullptr, *C2 = nullptr, *res = nullptr;
    if (cudaMalloc(&A, (size_t)M * K * 8) || cudaMalloc(&B, (size_t)N * K * 8) || cudaMalloc(&C1, (size_t)M * N * 8) ||
        cudaMalloc(&C2, (size_t)M * N * 8) || cudaMalloc(&res, 16))
        return -2;
    i8t_fill_kernel<<<(unsigned)(((long long)M * K + 255) / 256), 256>>>(A, (long long)M * K, 17u);
    i8t_fill_kernel<<<(unsigned)(((long long)N * K + 255) / 256), 256>>>(B, (long long)N * K, 91u);
    cudaMemset(res, 0, 16);
    I8Plan* p = i8_plan_create();
    int rc = 0;
    const int wsel = i8_width(S, I8_USE_OTHER);          // the self-test follows the "everything" setting (g_i8_wide = 3)
    const int TNs = wsel == 1 ? 128 : I8_TN;
    for (int ti = 0; ti < M / I8_TM; ++ti)
        for (int tj = 0; tj < N / TNs; ++tj)
            p->host_tiles.push_back({ti * I8_TM, tj * TNs, 0, K / I8_KC, (long long)ti * I8_TM * N + (long long)tj * TNs});
    if (i8_upload_tiles(*p, nullptr) != cudaSuccess) rc = -2;
    cudaEvent_t e0, e1, e2, e3;
    cudaEventCreate(&e0); cudaEventCreate(&e1); cudaEventCreate(&e2); cudaEventCreate(&e3);
    for (int rep = 0; rep < 2 && rc == 0; ++rep) {
        cudaEventRecord(e0);
        if (i8_slice(p->opA, A, K, 1, 0, 1, M, K, S, 0, nullptr) != cudaSuccess) rc = -3;
        if (rc == 0 && i8_slice(p->opB, B, K, 1, 0, 1, N, K, S, 0, nullptr) != cudaSuccess) rc = -3;
        cudaEventRecord(e1);
        if (rc == 0 && i8_launch(p->opA, p->opB, p->tiles, (int)p->host_tiles.size(), S, 1.0, 0.0, C1, N, nullptr, wsel) != cudaSuccess) rc = -4;
        cudaEventRecord(e2);
        GemmArgs g{};
        g.A = A; g.lda = K; g.B = B; g.ldb = K; g.C = C2; g.ldc = N;
        g.M = M; g.N = N; g.K = K; g.alpha = 1.0; g.beta = 0.0;
        if (rc == 0 && launch_gemm(0, 1, g, 1, nullptr) != cudaSuccess) rc = -5;
        cudaEventRecord(e3);
    }
    if (rc == 0 && cudaDeviceSynchronize() != cudaSuccess) rc = -6;
    if (rc == 0) {
        i8t_maxdiff_kernel<<<256, 256>>>(C1, C2, (long long)M * N, res);
        double h[2];
        if (cudaMemcpy(h, res, 16, cudaMemcpyDeviceToHost) != cudaSuccess) rc = -7;
        else {
            float t01 = 0.f, t12 = 0.f, t23 = 0.f;
            cudaEventElapsedTime(&t01, e0, e1); cudaEventElapsedTime(&t12, e1, e2); cudaEventElapsedTime(&t23, e2, e3);
            out_host[0] = h[1] > 0.0 ? h[0] / h[1] : -1.0;
            out_host[1] = t01 + t12; out_host[2] = t12; out_host[3] = t23;
        }
    }
    cudaDeviceSynchronize();
    cudaGetLastError();
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaEventDestroy(e2); cudaEventDestroy(e3);
    i8_plan_destroy(p);
    cudaFree(A); cudaFree(B); cudaFree(C1); cudaFree(C2); cudaFree(res);
    return rc;
}
