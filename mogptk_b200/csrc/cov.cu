// Covariance kernels of the exact-GP step (sm_100a):
//   prep        : constrained parameters -> per channel-pair component table, diagonal constants
//   kbuild      : tiled Gram / cross-covariance build (replaces MultiOutputKernel.K + Ksub,
//                 mogptk/gpr/kernel.py:446-481, gpr/multioutput.py:178-204,531-547, gpr/singleoutput.py:594-600)
//   grad_reduce : sum_{r,s} W_rs dK_rs/d(derived constants) without materialising dK/dtheta
//   finalize    : log-marginal likelihood, noise gradient, chain rule to the packed parameters
// The x tile of every CTA is staged into shared memory with a 1-D TMA bulk copy
// (cp.async.bulk + mbarrier); outputs are written with 16-byte vector stores.
#include "covmath.cuh"
#include <cstdio>
#include <cstdlib>

#define RC 8   // components processed per shared-memory trig table

// ------------------------------------------------------------------ exp for non-positive arguments
// The covariance kernels are bound by the fp64 pipe (one exp per element and component), so the library exp
// (~28 fp64-pipe instructions with its special-case handling) is replaced by: x = (64 m + j) ln2/64 + r,
// exp(x) = 2^m * 2^(j/64) * p(r) with a 64-entry table (staged in shared memory) and a degree-5 polynomial on
// |r| <= ln2/128 (~11 fp64 instructions, max relative error 4e-16).  Arguments are always <= 0 here; results
// below ~1e-307 flush to 0, NaN propagates.
__device__ const double EXP2_TABLE[64] = {
    1.00000000000000000e+00, 1.01088928605170048e+00, 1.02189714865411663e+00, 1.03302487902122841e+00,
    1.04427378242741375e+00, 1.05564517836055716e+00, 1.06714040067682370e+00, 1.07876079775711986e+00,
    1.09050773266525769e+00, 1.10238258330784089e+00, 1.11438674259589243e+00, 1.12652161860824185e+00,
    1.13878863475669156e+00, 1.15118922995298267e+00, 1.16372485877757748e+00, 1.17639699165028122e+00,
    1.18920711500272103e+00, 1.20215673145270308e+00, 1.21524735998046896e+00, 1.22848053610687002e+00,
    1.24185781207348400e+00, 1.25538075702469110e+00, 1.26905095719173322e+00, 1.28287001607877826e+00,
    1.29683955465100964e+00, 1.31096121152476441e+00, 1.32523664315974132e+00, 1.33966752405330292e+00,
    1.35425554693689265e+00, 1.36900242297459052e+00, 1.38390988196383202e+00, 1.39897967253831124e+00,
    1.41421356237309515e+00, 1.42961333839197002e+00, 1.44518080697704665e+00, 1.46091779418064704e+00,
    1.47682614593949935e+00, 1.49290772829126484e+00, 1.50916442759342284e+00, 1.52559815074453842e+00,
    1.54221082540794074e+00, 1.55900440023783693e+00, 1.57598084510788650e+00, 1.59314215134226700e+00,
    1.61049033194925428e+00, 1.62802742185734783e+00, 1.64575547815396495e+00, 1.66367658032673638e+00,
    1.68179283050742900e+00, 1.70010635371852348e+00, 1.71861929812247793e+00, 1.73733383527370622e+00,
    1.75625216037329945e+00, 1.77537649252652119e+00, 1.79470907500310717e+00, 1.81425217550039886e+00,
    1.83400808640934243e+00, 1.85397912508338547e+00, 1.87416763411029996e+00, 1.89457598158696561e+00,
    1.91520656139714740e+00, 1.93606179349229435e+00, 1.95714412417540018e+00, 1.97845602638795093e+00};

__device__ __forceinline__ double exp_nonpos(double x, const double* __restrict__ tab /* shared */) {
    const double INV = 9.23324826168936567683e+01;        // 64 / ln 2
    const double C_HI = 1.08304246932675596327e-02;       // ln 2 / 64, 32 significant bits
    const double C_LO = 2.98158582698529328128e-12;
    const double MAGIC = 6755399441055744.0;              // 1.5 * 2^52: rounds to nearest integer
    const double t = fma(x, INV, MAGIC);
    const int k = __double2loint(t);
    const double kf = t - MAGIC;
    double r = fma(kf, -C_HI, x);
    r = fma(kf, -C_LO, r);
    double p = fma(r, 8.3333333333333332e-03, 4.1666666666666664e-02);
    p = fma(p, r, 1.6666666666666666e-01);
    p = fma(p, r, 0.5);
    p = fma(p, r, 1.0);
    p = fma(p, r, 1.0);
    const double res = tab[k & 63] * p;
    const int hi = __double2hiint(res) + ((k >> 6) << 20);          // scale by 2^(k>>6)
    const double scaled = __hiloint2double(hi, __double2loint(res));
    return (x >= -708.0) ? scaled : ((x < -708.0) ? 0.0 : x);
}

// ------------------------------------------------------------------ TMA bulk helpers
__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
    unsigned a = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(a), "r"(count));
}
__device__ __forceinline__ void mbar_init_fence() { asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, unsigned bytes) {
    unsigned a = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(a), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void* dst, const void* src, unsigned bytes, uint64_t* bar) {
    unsigned d = (unsigned)__cvta_generic_to_shared(dst);
    unsigned b = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(d),
                 "l"(src), "r"(bytes), "r"(b)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
    unsigned a = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra.uni WAIT_DONE;\n"
        "bra.uni WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(a),
        "r"(parity)
        : "memory");
}

// Stage the row and column coordinates of a tile (na / nb rows x D doubles) into shared memory.  When both
// sources are 16-byte aligned with sizes that are multiples of 16 bytes, thread 0 issues two TMA bulk copies
// that complete ONE mbarrier phase (a single expect_tx covering both), and everybody waits on that phase; a
// second, separate phase could complete before a slow thread has observed the first (parity aliasing) and
// dead-lock it.  Otherwise plain loads.  Ends with the data visible to all threads.
__device__ __forceinline__ void stage_xy(double* dsta, const double* srca, int na, double* dstb, const double* srcb,
                                         int nb, int D, uint64_t* bar, unsigned& phase, int tid, int nthreads) {
    const unsigned ba = (unsigned)(na * D * sizeof(double)), bb = (unsigned)(nb * D * sizeof(double));
    const bool tma_ok = ((reinterpret_cast<uintptr_t>(srca) & 15) == 0) && ((ba & 15) == 0) && ba > 0 &&
                        ((reinterpret_cast<uintptr_t>(srcb) & 15) == 0) && ((bb & 15) == 0) && bb > 0;
    if (tma_ok) {
        if (tid == 0) {
            mbar_expect_tx(bar, ba + bb);
            tma_bulk_g2s(dsta, srca, ba, bar);
            tma_bulk_g2s(dstb, srcb, bb, bar);
        }
        mbar_wait(bar, phase);
        phase ^= 1;
    } else {
        for (int i = tid; i < na * D; i += nthreads) dsta[i] = srca[i];
        for (int i = tid; i < nb * D; i += nthreads) dstb[i] = srcb[i];
    }
    __syncthreads();
}

// ------------------------------------------------------------------ prep
// chanbuf: [0..C) kdiag_gram | [C..2C) kdiag_api | [2C..3C) sigma^2 | [3C] jitter_add
// Window kinds (MOHSM): the Gram diagonal depends on the row, K_rr = sum_q alpha_q E_q(x_r), E_q(x) = exp(-1/2 l_q sum_d (x_d - c_qd)^2);
// winsum[(c R + q)(2 + D)] = [sum_r E_q, sum_r E_q sum_d s_d^2, sum_r E_q s_d] over the rows of channel c (the relative jitter
// and its gradient need them), chanbuf[c] = sum over the channel's rows of K_rr.
__global__ void __launch_bounds__(256) prep_kernel(KernSpec s, const double* __restrict__ params,
                                                   const double* __restrict__ sigma, const double* __restrict__ data_var,
                                                   const int32_t* __restrict__ chan, long long N, double jitter_rel,
                                                   double* __restrict__ comps, double* __restrict__ chanbuf,
                                                   const double* __restrict__ x, double* __restrict__ winsum) {
    __shared__ double red[256];
    const int tid = threadIdx.x;
    const int st = s.st;
    const int total = s.C * s.C * s.R;
    for (int e = tid; e < total; e += 256) {
        const int r = e % s.R, pj = e / s.R, j = pj % s.C, i = pj / s.C;
        pair_comp(s.kind, s.C, s.Q, s.D, params, i, j, r, comps + (size_t)e * st);
    }
    double dv = 0.0;
    if (data_var)
        for (long long r = tid; r < N; r += 256) dv += data_var[r];
    red[tid] = dv;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (tid < o) red[tid] += red[tid + o];
        __syncthreads();
    }
    if (s.window) {
        const int warp = tid >> 5, lane = tid & 31, D = s.D, wst = 2 + D;
        for (int e = warp; e < s.C * s.R; e += 8) {
            const int c = e / s.R, r = e % s.R;
            const double* cp = comps + (size_t)((c * s.C + c) * s.R + r) * st;
            const double ell = cp[2 + 3 * D];
            double b0 = 0.0, b5 = 0.0, b6[MOGP_MAX_D];
            for (int d = 0; d < D; ++d) b6[d] = 0.0;
            for (long long row = chan[c] + lane; row < chan[c + 1]; row += 32) {
                double ss = 0.0, sd[MOGP_MAX_D];
                for (int d = 0; d < D; ++d) { sd[d] = x[row * D + d] - cp[3 + 3 * D + d]; ss = fma(sd[d], sd[d], ss); }
                const double E = exp(-0.5 * ell * ss);
                b0 += E; b5 += E * ss;
                for (int d = 0; d < D; ++d) b6[d] += E * sd[d];
            }
            for (int o = 16; o > 0; o >>= 1) {
                b0 += __shfl_xor_sync(0xffffffffu, b0, o);
                b5 += __shfl_xor_sync(0xffffffffu, b5, o);
                for (int d = 0; d < D; ++d) b6[d] += __shfl_xor_sync(0xffffffffu, b6[d], o);
            }
            if (lane == 0) {
                winsum[(size_t)e * wst] = b0; winsum[(size_t)e * wst + 1] = b5;
                for (int d = 0; d < D; ++d) winsum[(size_t)e * wst + 2 + d] = b6[d];
            }
        }
        __syncthreads();
    }
    if (tid < s.C) {
        const int c = tid;
        double kd = 0.0;
        for (int r = 0; r < s.R; ++r)
            kd += comps[(size_t)((c * s.C + c) * s.R + r) * st] * (s.window ? winsum[(size_t)(c * s.R + r) * (2 + s.D)] : 1.0);
        chanbuf[c] = kd;
        chanbuf[s.C + c] = kdiag_api_value(s.kind, s.C, s.Q, s.D, params, comps, s.R, c);
        chanbuf[2 * s.C + c] = sigma ? sigma[c] * sigma[c] : 0.0;
    }
    __syncthreads();
    if (tid == 0) {
        double tot = red[0];
        for (int c = 0; c < s.C; ++c) {
            const double n_c = (double)(chan[c + 1] - chan[c]);
            tot += (s.window ? chanbuf[c] : n_c * chanbuf[c]) + n_c * chanbuf[2 * s.C + c];
        }
        chanbuf[3 * s.C] = jitter_rel * tot / (double)N;
    }
}

cudaError_t launch_prep(const KernSpec& s, const double* params, const double* sigma, const double* data_var,
                        const int32_t* chan_dev, int64_t N, double jitter_rel, double* comps, double* chanbuf,
                        cudaStream_t st, const double* x, double* winsum) {
    if (s.window && (!x || !winsum)) return cudaErrorInvalidValue;
    prep_kernel<<<1, 256, 0, st>>>(s, params, sigma, data_var, chan_dev, (long long)N, jitter_rel, comps, chanbuf, x, winsum);
    MOGP_COUNT(1);
    return cudaGetLastError();
}

// ------------------------------------------------------------------ shared tile machinery
// Thread (ty, tx) of 256 owns rows ty + 16*i (i<4) and columns tx + 16*j (j<4) of the 64x64 tile: for a fixed j the 16
// lanes of a half-warp read consecutive doubles of the per-column tables (conflict-free; the round-1 mapping tx*4 + j
// put them 32 bytes apart: 4-way bank conflicts, 35 M per N = 8192 build) and write 128 contiguous bytes of an output row.
struct TileSmem {
    double xa[MOGP_TILE * MOGP_MAX_D];
    double xb[MOGP_TILE * MOGP_MAX_D];
    double rc[RC][MOGP_TILE], rs[RC][MOGP_TILE];     // cos/sin of the row angles
    double cc[RC][MOGP_TILE], cs[RC][MOGP_TILE];     // cos/sin of the column angles
    double comp[RC][MOGP_MAX_STRIDE];
    double x0[MOGP_MAX_D];                           // the shift applied to xa / xb (window kinds shift their centre by it)
    double expt[64];                                 // 2^(j/64), see exp_nonpos
    uint64_t bar;
};

template <int DT>
__device__ __forceinline__ int dims(int D) { return DT > 0 ? DT : D; }

// Loads x rows/cols of the tile (shifted by the tile's first column coordinate) into smem.
template <int DT>
__device__ __forceinline__ void load_tile_x(TileSmem& sm, const CovTile& t, const double* __restrict__ x1,
                                            const double* __restrict__ x2, int Drt, unsigned& phase, int tid) {
    const int D = dims<DT>(Drt);
    if (tid < 64) sm.expt[tid] = EXP2_TABLE[tid];
    stage_xy(sm.xa, x1 + (size_t)t.r0 * D, t.nr, sm.xb, x2 + (size_t)t.c0 * D, t.nc, D, &sm.bar, phase, tid, 256);
    // shift by x0 = first column point so that the trig arguments stay small
    double x0[MOGP_MAX_D];
#pragma unroll
    for (int d = 0; d < (DT > 0 ? DT : MOGP_MAX_D); ++d)
        if (d < D) x0[d] = sm.xb[d];
    __syncthreads();
    if (tid < D) sm.x0[tid] = x0[tid];
    for (int e = tid; e < MOGP_TILE * D; e += 256) {
        const int d = e % D, r = e / D;
        sm.xa[e] = (r < t.nr) ? sm.xa[e] - x0[d] : 0.0;
        sm.xb[e] = (r < t.nc) ? sm.xb[e] - x0[d] : 0.0;
    }
}

// Fills the comp records and trig tables for components [rbase, rbase+nr) of `pair`.
template <int DT, bool COS>
__device__ __forceinline__ void fill_tables(TileSmem& sm, const double* __restrict__ comps, int pair, int R, int rbase,
                                            int nrc, int Drt, int st, int tid) {
    const int D = dims<DT>(Drt);
    for (int e = tid; e < nrc * st; e += 256) sm.comp[e / st][e % st] = comps[(size_t)(pair * R + rbase) * st + e];
    __syncthreads();
    if (COS) {
        for (int e = tid; e < nrc * 2 * MOGP_TILE; e += 256) {
            const int r = e / (2 * MOGP_TILE), w = e % (2 * MOGP_TILE);
            const double* cp = sm.comp[r];
            const double* m = cp + 2 + D;
            const double* th = cp + 2 + 2 * D;
            double sn, cs;
            if (w < MOGP_TILE) {
                double a = cp[1];
                for (int d = 0; d < D; ++d) a += m[d] * (sm.xa[w * D + d] + th[d]);
                sincospi(2.0 * a, &sn, &cs);
                sm.rc[r][w] = cs; sm.rs[r][w] = sn;
            } else {
                const int c = w - MOGP_TILE;
                double a = 0.0;
                for (int d = 0; d < D; ++d) a += m[d] * sm.xb[c * D + d];
                sincospi(2.0 * a, &sn, &cs);
                sm.cc[r][c] = cs; sm.cs[r][c] = sn;
            }
        }
    }
    __syncthreads();
}

// Prior variance of a window kind at one input: sum_q alpha_q exp(-1/2 l_q sum_d (x_d - c_qd)^2) over the records of the
// diagonal pair (c, c).  Shared by the Gram diagonal of kbuild and by kdiag_x so that K_diag == diag(K) bit for bit.
__device__ __forceinline__ double win_diag_value(const double* __restrict__ cpair, int R, int D, int st,
                                                 const double* __restrict__ xrow, const double* __restrict__ tab) {
    double acc = 0.0;
    for (int r = 0; r < R; ++r) {
        const double* cp = cpair + (size_t)r * st;
        double ss = 0.0;
        for (int d = 0; d < D; ++d) { const double sd = xrow[d] - cp[3 + 3 * D + d]; ss = fma(sd, sd, ss); }
        acc = fma(cp[0], exp_nonpos(-0.5 * (cp[2 + 3 * D] * ss), tab), acc);
    }
    return acc;
}

// ------------------------------------------------------------------ kbuild
// mode 0: Gram lower tiles into the padded factor buffer (+ padding rows);  mode 1: Gram, every
// lower tile is also written transposed (full symmetric output);  mode 2: cross-covariance.
template <int DT, bool COS, int MINB, bool WIN>
__global__ void __launch_bounds__(256, MINB) kbuild_kernel(KernSpec s, const CovTile* __restrict__ tiles, int ntiles, int mode,
                                                     const double* __restrict__ comps, const double* __restrict__ chanbuf,
                                                     const double* __restrict__ x1, const double* __restrict__ x2,
                                                     const double* __restrict__ data_var, int add_diag,
                                                     double* __restrict__ K, long long ldk, long long N, long long Np) {
    extern __shared__ __align__(16) unsigned char smraw[];
    TileSmem& sm = *reinterpret_cast<TileSmem*>(smraw);
    double* tbuf = reinterpret_cast<double*>(smraw + sizeof(TileSmem));   // 64 x 65 transpose staging (mode 1)
    const int tid = threadIdx.x;

    if ((int)blockIdx.x >= ntiles) {                  // padding row of the factor buffer: unit diagonal
        const long long r = N + (blockIdx.x - ntiles);
        if (r < Np) {
            for (long long c = tid; c <= r; c += 256) K[r * ldk + c] = (c == r) ? 1.0 : 0.0;
        }
        return;
    }
    const CovTile t = tiles[blockIdx.x];
    const int D = dims<DT>(s.D);
    const int ty = tid >> 4, tx = tid & 15;
    const int pi = t.pair / s.C, pj = t.pair % s.C;
    const bool zero_block = (s.kind == MOGP_KIND_SM) && (pi != pj);

    double acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.0;

    if (!zero_block) {
        if (tid == 0) { mbar_init(&sm.bar, 1); mbar_init_fence(); }
        __syncthreads();
        unsigned phase = 0;
        load_tile_x<DT>(sm, t, x1, x2 ? x2 : x1, s.D, phase, tid);
        for (int rbase = 0; rbase < s.R; rbase += RC) {
            const int nrc = min(RC, s.R - rbase);
            __syncthreads();
            fill_tables<DT, COS>(sm, comps, t.pair, s.R, rbase, nrc, s.D, s.st, tid);
            double xb[4][DT > 0 ? DT : MOGP_MAX_D];
#pragma unroll
            for (int j = 0; j < 4; ++j)
#pragma unroll
                for (int d = 0; d < (DT > 0 ? DT : MOGP_MAX_D); ++d)
                    if (d < D) xb[j][d] = sm.xb[(tx + 16 * j) * D + d];
            for (int r = 0; r < nrc; ++r) {
                const double* cp = sm.comp[r];
                const double alpha = cp[0];
                const double ell = WIN ? cp[2 + 3 * D] : 0.0;
                double cB[4], sB[4];
                if (COS) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) { cB[j] = sm.cc[r][tx + 16 * j]; sB[j] = sm.cs[r][tx + 16 * j]; }
                }
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int row = ty + 16 * i;
                    double ua[DT > 0 ? DT : MOGP_MAX_D], ha[DT > 0 ? DT : MOGP_MAX_D];
#pragma unroll
                    for (int d = 0; d < (DT > 0 ? DT : MOGP_MAX_D); ++d)
                        if (d < D) {
                            ua[d] = sm.xa[row * D + d] + cp[2 + 2 * D + d];
                            // mid-point minus centre in the tile's shifted coordinates, (xa + xb)/2 - (c - x0); the sum first, so
                            // that the value is bit-for-bit symmetric in (a, b) (the diagonal tiles compute both halves)
                            if (WIN) ha[d] = sm.xa[row * D + d];
                        }
                    double cA = 1.0, sA = 0.0;
                    if (COS) { cA = sm.rc[r][row]; sA = sm.rs[r][row]; }
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        double e = 0.0;
#pragma unroll
                        for (int d = 0; d < (DT > 0 ? DT : MOGP_MAX_D); ++d)
                            if (d < D) {
                                const double u = ua[d] - xb[j][d]; e = fma(cp[2 + d] * u, u, e);
                                if (WIN) { const double sd = fma(0.5, ha[d] + xb[j][d], sm.x0[d] - cp[3 + 3 * D + d]); e = fma(ell * sd, sd, e); }
                            }
                        double val = alpha * exp_nonpos(-0.5 * e, sm.expt);
                        if (COS) val *= fma(cA, cB[j], sA * sB[j]);
                        acc[i][j] += val;
                    }
                }
            }
        }
    }

    // diagonal of the Gram matrix: exact K_diag value (+ noise, data variance, relative jitter)
    if (mode != 2 && (t.flags & 1)) {
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int row = ty + 16 * i, col = tx + 16 * j;
                if (row == col && row < t.nr) {
                    double v = WIN ? win_diag_value(comps + (size_t)(pi * s.C + pi) * s.R * s.st, s.R, D, s.st,
                                                    x1 + (size_t)(t.r0 + row) * D, sm.expt)
                                   : chanbuf[pi];
                    if (add_diag) {
                        v += chanbuf[2 * s.C + pi];
                        if (data_var) v += data_var[t.r0 + row];
                        v += chanbuf[3 * s.C];
                    }
                    acc[i][j] = v;
                }
            }
    }

#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int row = ty + 16 * i;
        if (row >= t.nr) continue;
        double* dst = K + (long long)(t.r0 + row) * ldk + t.c0;
#pragma unroll
        for (int j = 0; j < 4; ++j)
            if (tx + 16 * j < t.nc) dst[tx + 16 * j] = acc[i][j];
    }
    if (mode == 1 && !(t.flags & 1)) {                // mirrored copy, transposed through shared memory
        __syncthreads();
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) tbuf[(ty + 16 * i) * 65 + tx + 16 * j] = acc[i][j];
        __syncthreads();
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int orow = ty + 16 * i;            // output row = original column
            if (orow >= t.nc) continue;
            double* dst = K + (long long)(t.c0 + orow) * ldk + t.r0;
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (tx + 16 * j < t.nr) dst[tx + 16 * j] = tbuf[(tx + 16 * j) * 65 + orow];
        }
    }
}

// resident CTAs per SM the covariance kernels are compiled for: 2 (125 registers, no spills) or 3 (80 registers, a few
// spilled words); selectable at run time so that both can be measured (MOGP_COV_MINB / mogp_set_cov_minb)
static int g_cov_minb = std::getenv("MOGP_COV_MINB") ? std::atoi(std::getenv("MOGP_COV_MINB")) : 2;
extern "C" int mogp_set_cov_minb(int v) { g_cov_minb = v == 3 ? 3 : 2; ++g_mogp_cfg_epoch; return 0; }

template <int DT, bool COS, int MINB, bool WIN = false>
static cudaError_t launch_kbuild_t(const KernSpec& s, const TileList& tl, int nblocks, const double* comps,
                                   const double* chanbuf, const double* x1, const double* x2, const double* data_var,
                                   int add_diag, double* K, long long ldk, int64_t N, int64_t Np, cudaStream_t st) {
    const size_t smem = sizeof(TileSmem) + 64 * 65 * sizeof(double);
    auto kern = kbuild_kernel<DT, COS, MINB, WIN>;
    static PerDeviceOnce once;
    if (OnceGuard og{once}; og.needed()) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    if (nblocks <= 0) return cudaSuccess;
    kern<<<nblocks, 256, smem, st>>>(s, tl.dev, tl.n, tl.mode, comps, chanbuf, x1, x2, data_var, add_diag, K, ldk,
                                     (long long)N, (long long)Np);
    MOGP_COUNT(1);
    return cudaGetLastError();
}

cudaError_t launch_kbuild(const KernSpec& s, const TileList& tl, const double* comps, const double* chanbuf,
                          const double* x1, const double* x2, const int32_t* chan1_dev, const double* data_var,
                          int add_diag, double* K, long long ldk, int64_t N, int64_t Np, cudaStream_t st) {
    (void)chan1_dev;
    const int nblocks = tl.n + (tl.mode == 0 ? (int)(Np - N) : 0);
#define KB_ARGS s, tl, nblocks, comps, chanbuf, x1, x2, data_var, add_diag, K, ldk, N, Np, st
    if (s.window) return s.D == 1 ? launch_kbuild_t<1, true, 2, true>(KB_ARGS) : launch_kbuild_t<0, true, 2, true>(KB_ARGS);
    if (g_cov_minb == 3) {
        if (s.D == 1) return s.has_cos ? launch_kbuild_t<1, true, 3>(KB_ARGS) : launch_kbuild_t<1, false, 3>(KB_ARGS);
        return s.has_cos ? launch_kbuild_t<0, true, 3>(KB_ARGS) : launch_kbuild_t<0, false, 3>(KB_ARGS);
    }
    if (s.D == 1) return s.has_cos ? launch_kbuild_t<1, true, 2>(KB_ARGS) : launch_kbuild_t<1, false, 2>(KB_ARGS);
    return s.has_cos ? launch_kbuild_t<0, true, 2>(KB_ARGS) : launch_kbuild_t<0, false, 2>(KB_ARGS);
#undef KB_ARGS
}

__global__ void kdiag_kernel(int C, const double* __restrict__ chanbuf, const int32_t* __restrict__ chan, long long N,
                             double* __restrict__ out) {
    const long long r = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (r >= N) return;
    int c = 0;
    while (c + 1 < C && r >= chan[c + 1]) ++c;
    out[r] = chanbuf[C + c];
}
// window kinds: the prior variance depends on the input (gpr/multioutput.py:389-395)
__global__ void __launch_bounds__(256) kdiag_win_kernel(KernSpec s, const double* __restrict__ comps,
                                                        const int32_t* __restrict__ chan, const double* __restrict__ x,
                                                        long long N, double* __restrict__ out) {
    __shared__ double expt[64];
    if (threadIdx.x < 64) expt[threadIdx.x] = EXP2_TABLE[threadIdx.x];
    __syncthreads();
    const long long r = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (r >= N) return;
    int c = 0;
    while (c + 1 < s.C && r >= chan[c + 1]) ++c;
    out[r] = win_diag_value(comps + (size_t)(c * s.C + c) * s.R * s.st, s.R, s.D, s.st, x + (size_t)r * s.D, expt);
}
cudaError_t launch_kdiag(const KernSpec& s, const double* chanbuf, const int32_t* chan_dev, int64_t N, double* out,
                         cudaStream_t st, const double* comps, const double* x) {
    if (s.window) {
        if (!comps || !x) return cudaErrorInvalidValue;
        kdiag_win_kernel<<<(unsigned)((N + 255) / 256), 256, 0, st>>>(s, comps, chan_dev, x, (long long)N, out);
    } else {
        kdiag_kernel<<<(unsigned)((N + 255) / 256), 256, 0, st>>>(s.C, chanbuf, chan_dev, (long long)N, out);
    }
    MOGP_COUNT(1);
    return cudaGetLastError();
}

// ------------------------------------------------------------------ grad_reduce
// Per tile and component: [S0, S4, S1[D], S2[D], S3[D]] (+ [S5, S6[D]] for the window kinds) with the symmetric weight folded
// into W.
template <int DT, bool COS, int MINB, bool WIN>
__global__ void __launch_bounds__(256, MINB) grad_reduce_kernel(KernSpec s, const CovTile* __restrict__ tiles,
                                                          const double* __restrict__ comps, const double* __restrict__ x,
                                                          const double* __restrict__ W, long long ldw,
                                                          const double* __restrict__ avec,
                                                          double* __restrict__ tile_part) {
    extern __shared__ __align__(16) unsigned char smraw[];
    TileSmem& sm = *reinterpret_cast<TileSmem*>(smraw);
    double* wpart = reinterpret_cast<double*>(smraw + sizeof(TileSmem));   // [8 warps][RC][stride]
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const CovTile t = tiles[blockIdx.x];
    const int D = dims<DT>(s.D);
    const int st = s.st;
    const int ty = tid >> 4, tx = tid & 15;
    const int pi = t.pair / s.C, pj = t.pair % s.C;
    double* outp = tile_part + (size_t)blockIdx.x * s.R * st;
    if (s.kind == MOGP_KIND_SM && pi != pj) {
        for (int e = tid; e < s.R * st; e += 256) outp[e] = 0.0;
        return;
    }
    if (tid == 0) { mbar_init(&sm.bar, 1); mbar_init_fence(); }
    __syncthreads();
    unsigned phase = 0;
    load_tile_x<DT>(sm, t, x, x, s.D, phase, tid);

    // weighted W values of this thread's 4x4 patch (lower triangle of the global matrix only)
    double wv[4][4];
    const bool diag_tile = (t.flags & 1) != 0;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int row = ty + 16 * i, col = tx + 16 * j;
            double v = 0.0;
            if (row < t.nr && col < t.nc) {
                double wgt = 2.0;
                if (diag_tile) wgt = row > col ? 2.0 : (row == col ? 1.0 : 0.0);
                if (wgt != 0.0) {
                    v = W[(long long)(t.r0 + row) * ldw + t.c0 + col];
                    // W holds K^-1 when avec is given: form (K^-1 - a a^T)/2 on the fly (lets the K^-1 GEMM run
                    // concurrently with the solves that produce a)
                    if (avec) v = 0.5 * (v - avec[t.r0 + row] * avec[t.c0 + col]);
                    v *= wgt;
                }
            }
            wv[i][j] = v;
        }

    for (int rbase = 0; rbase < s.R; rbase += RC) {
        const int nrc = min(RC, s.R - rbase);
        __syncthreads();
        fill_tables<DT, COS>(sm, comps, t.pair, s.R, rbase, nrc, s.D, st, tid);
        double xb[4][DT > 0 ? DT : MOGP_MAX_D];
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int d = 0; d < (DT > 0 ? DT : MOGP_MAX_D); ++d)
                if (d < D) xb[j][d] = sm.xb[(tx + 16 * j) * D + d];
        for (int r = 0; r < nrc; ++r) {
            const double* cp = sm.comp[r];
            const double ell = WIN ? cp[2 + 3 * D] : 0.0;
            double s0 = 0.0, s4 = 0.0, s5 = 0.0;
            double s1[DT > 0 ? DT : MOGP_MAX_D], s2[DT > 0 ? DT : MOGP_MAX_D], s3[DT > 0 ? DT : MOGP_MAX_D];
            double s6[DT > 0 ? DT : MOGP_MAX_D];
#pragma unroll
            for (int d = 0; d < (DT > 0 ? DT : MOGP_MAX_D); ++d) s1[d] = s2[d] = s3[d] = s6[d] = 0.0;
            double cB[4], sB[4];
            if (COS) {
#pragma unroll
                for (int j = 0; j < 4; ++j) { cB[j] = sm.cc[r][tx + 16 * j]; sB[j] = sm.cs[r][tx + 16 * j]; }
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int row = ty + 16 * i;
                double ua[DT > 0 ? DT : MOGP_MAX_D], ha[DT > 0 ? DT : MOGP_MAX_D];
#pragma unroll
                for (int d = 0; d < (DT > 0 ? DT : MOGP_MAX_D); ++d)
                    if (d < D) {
                        ua[d] = sm.xa[row * D + d] + cp[2 + 2 * D + d];
                        if (WIN) ha[d] = sm.xa[row * D + d];
                    }
                double cA = 1.0, sA = 0.0;
                if (COS) { cA = sm.rc[r][row]; sA = sm.rs[r][row]; }
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    double u[DT > 0 ? DT : MOGP_MAX_D], sd[DT > 0 ? DT : MOGP_MAX_D];
                    double e = 0.0, ss = 0.0;
#pragma unroll
                    for (int d = 0; d < (DT > 0 ? DT : MOGP_MAX_D); ++d)
                        if (d < D) {
                            u[d] = ua[d] - xb[j][d]; e = fma(cp[2 + d] * u[d], u[d], e);
                            if (WIN) { sd[d] = fma(0.5, ha[d] + xb[j][d], sm.x0[d] - cp[3 + 3 * D + d]); ss = fma(sd[d], sd[d], ss); }
                        }
                    if (WIN) e = fma(ell, ss, e);
                    const double we = wv[i][j] * exp_nonpos(-0.5 * e, sm.expt);
                    double wec = we, wes = 0.0;
                    if (COS) {
                        wec = we * fma(cA, cB[j], sA * sB[j]);      // cos(A - B)
                        wes = we * fma(sA, cB[j], -cA * sB[j]);     // sin(A - B)
                    }
                    s0 += wec;
                    s4 += wes;
#pragma unroll
                    for (int d = 0; d < (DT > 0 ? DT : MOGP_MAX_D); ++d)
                        if (d < D) {
                            const double wu = wec * u[d];
                            s3[d] += wu;
                            s1[d] = fma(wu, u[d], s1[d]);
                            s2[d] = fma(wes, u[d], s2[d]);
                            if (WIN) s6[d] = fma(wec, sd[d], s6[d]);
                        }
                    if (WIN) s5 = fma(wec, ss, s5);
                }
            }
            // warp reduction, lane 0 parks the warp's partial in shared memory
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                s0 += __shfl_xor_sync(0xffffffffu, s0, o);
                s4 += __shfl_xor_sync(0xffffffffu, s4, o);
#pragma unroll
                for (int d = 0; d < (DT > 0 ? DT : MOGP_MAX_D); ++d)
                    if (d < D) {
                        s1[d] += __shfl_xor_sync(0xffffffffu, s1[d], o);
                        s2[d] += __shfl_xor_sync(0xffffffffu, s2[d], o);
                        s3[d] += __shfl_xor_sync(0xffffffffu, s3[d], o);
                        if (WIN) s6[d] += __shfl_xor_sync(0xffffffffu, s6[d], o);
                    }
                if (WIN) s5 += __shfl_xor_sync(0xffffffffu, s5, o);
            }
            if (lane == 0) {
                double* wp = wpart + (size_t)(warp * RC + r) * st;
                wp[0] = s0; wp[1] = s4;
#pragma unroll
                for (int d = 0; d < (DT > 0 ? DT : MOGP_MAX_D); ++d)
                    if (d < D) {
                        wp[2 + d] = s1[d]; wp[2 + D + d] = s2[d]; wp[2 + 2 * D + d] = s3[d];
                        if (WIN) wp[3 + 3 * D + d] = s6[d];
                    }
                if (WIN) wp[2 + 3 * D] = s5;
            }
        }
        __syncthreads();
        for (int e = tid; e < nrc * st; e += 256) {
            double a = 0.0;
#pragma unroll
            for (int w = 0; w < 8; ++w) a += wpart[(size_t)(w * RC) * st + e];
            outp[(size_t)rbase * st + e] = a;
        }
    }
}

cudaError_t launch_grad_reduce(const KernSpec& s, const TileList& tl, const double* comps, const double* x,
                               const double* W, long long ldw, const double* avec, double* tile_part, cudaStream_t st) {
    const size_t smem = sizeof(TileSmem) + (size_t)8 * RC * MOGP_MAX_STRIDE * sizeof(double);
    if (tl.n <= 0) return cudaSuccess;
#define LAUNCH_GR(DT, COS)                                                                                         \
    do {                                                                                                           \
        if (g_cov_minb == 3) { LAUNCH_GR_M(DT, COS, 3); } else { LAUNCH_GR_M(DT, COS, 2); }                       \
    } while (0)
#define LAUNCH_GR_M(DT, COS, MB) LAUNCH_GR_W(DT, COS, MB, false)
#define LAUNCH_GR_W(DT, COS, MB, WIN)                                                                              \
    do {                                                                                                           \
        auto kern = grad_reduce_kernel<DT, COS, MB, WIN>;                                                            \
        static PerDeviceOnce once;                                                                                 \
        if (OnceGuard og{once}; og.needed()) {                                                                                        \
            cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);    \
            if (e != cudaSuccess) return e;                                                                        \
        }                                                                                                          \
        kern<<<tl.n, 256, smem, st>>>(s, tl.dev, comps, x, W, ldw, avec, tile_part);                                     \
        MOGP_COUNT(1);                                                                                             \
    } while (0)
    if (s.window) { if (s.D == 1) LAUNCH_GR_W(1, true, 2, true); else LAUNCH_GR_W(0, true, 2, true); }
    else if (s.D == 1) { if (s.has_cos) LAUNCH_GR(1, true); else LAUNCH_GR(1, false); }
    else { if (s.has_cos) LAUNCH_GR(0, true); else LAUNCH_GR(0, false); }
#undef LAUNCH_GR
#undef LAUNCH_GR_M
#undef LAUNCH_GR_W
    return cudaGetLastError();
}

// ------------------------------------------------------------------ finalize
// pairsum: gsum[pl][r][k] = sum over the tiles of lower pair pl, in tile order (deterministic).
__global__ void pairsum_kernel(int R, int st, const int32_t* __restrict__ pair_first, const double* __restrict__ tile_part,
                               double* __restrict__ gsum) {
    const int pl = blockIdx.x;
    const int t0 = pair_first[pl], t1 = pair_first[pl + 1];
    for (int e = threadIdx.x; e < R * st; e += blockDim.x) {
        double a = 0.0;
        for (int t = t0; t < t1; ++t) a += tile_part[(size_t)t * R * st + e];
        gsum[(size_t)pl * R * st + e] = a;
    }
}

__device__ double block_sum_256(double v, double* red) {
    const int tid = threadIdx.x;
    __syncthreads();
    red[tid] = v;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (tid < o) red[tid] += red[tid + o];
        __syncthreads();
    }
    return red[0];
}

// The log marginal likelihood as soon as the solves are done (same arithmetic as finalize_kernel below), written straight into
// mapped pinned host memory with a sequence number behind it: the host can return from loss() while K^-1, the gradient
// reduction and the chain rule are still running -- whatever the caller enqueues next (the optimiser's kernels) is ordered
// behind them by the stream, as for any asynchronous CUDA work.
__global__ void __launch_bounds__(256) lml_early_kernel(const double* __restrict__ z, const double* __restrict__ logdet_part,
                                                        const int32_t* __restrict__ info, long long N, long long Np,
                                                        volatile double* __restrict__ host_out,
                                                        unsigned long long* __restrict__ counter) {
    __shared__ double red[256];
    const int tid = threadIdx.x;
    double zz = 0.0;
    for (long long r = tid; r < Np; r += 256) zz += z[r] * z[r];
    zz = block_sum_256(zz, red);
    double ld = 0.0;
    for (long long b = tid; b < Np / 64; b += 256) ld += logdet_part[b];
    ld = block_sum_256(ld, red);
    if (tid == 0) {
        const unsigned long long seq = ++(*counter);
        host_out[0] = -0.5 * (double)N * log(2.0 * MOGP_PI) - ld - 0.5 * zz;
        host_out[1] = (double)info[0];
        __threadfence_system();
        host_out[2] = (double)seq;
        __threadfence_system();
    }
}
cudaError_t launch_lml_early(const double* z, const double* logdet_part, const int32_t* info, int64_t N, int64_t Np,
                             double* host_out_dev, unsigned long long* counter, cudaStream_t st) {
    lml_early_kernel<<<1, 256, 0, st>>>(z, logdet_part, info, (long long)N, (long long)Np, host_out_dev, counter);
    MOGP_COUNT(1);
    return cudaGetLastError();
}

__global__ void __launch_bounds__(256) finalize_kernel(KernSpec s, int want_grad, const double* __restrict__ params,
                                                       const double* __restrict__ sigma, const double* __restrict__ comps,
                                                       double* __restrict__ gsum, const double* __restrict__ winsum,
                                                       const double* __restrict__ z,
                                                       const double* __restrict__ alpha, const double* __restrict__ kinv_diag,
                                                       const double* __restrict__ logdet_part, const int32_t* __restrict__ info,
                                                       const int32_t* __restrict__ chan, long long N, long long Np,
                                                       double jitter_rel, double* __restrict__ out) {
    __shared__ double red[256];
    __shared__ double adj[64], csum[64];
    __shared__ double trW_s;
    const int tid = threadIdx.x;
    double zz = 0.0;
    for (long long r = tid; r < Np; r += 256) zz += z[r] * z[r];
    zz = block_sum_256(zz, red);
    double ld = 0.0;
    for (long long b = tid; b < Np / 64; b += 256) ld += logdet_part[b];
    ld = block_sum_256(ld, red);
    if (tid == 0) {
        out[0] = -0.5 * (double)N * log(2.0 * MOGP_PI) - ld - 0.5 * zz;
        out[1] = (double)info[0];
    }
    if (!want_grad) return;
    double tr = 0.0;
    for (int c = 0; c < s.C; ++c) {
        double a = 0.0;
        for (long long r = chan[c] + tid; r < chan[c + 1]; r += 256) a += 0.5 * (kinv_diag[r] - alpha[r] * alpha[r]);
        a = block_sum_256(a, red);
        if (tid == 0) csum[c] = a;
        tr += a;
    }
    if (tid == 0) trW_s = tr;
    __syncthreads();
    const double trW = trW_s;
    if (tid < s.C) adj[tid] = jitter_rel / (double)N * trW * (double)(chan[tid + 1] - chan[tid]);
    __syncthreads();
    if (tid < s.C) out[2 + s.P + tid] = 2.0 * sigma[tid] * (csum[tid] + adj[tid]);
    if (s.window) {
        // row-dependent Gram diagonal: d (jitter term) / d (alpha_q, l_q, c_q) of the diagonal pair (c, c) comes from the row sums
        // of prep (winsum) instead of the row count; fold it into the pair's gradient sums [S0, S5, S6] and clear adj
        const double f = jitter_rel / (double)N * trW;
        const int wst = 2 + s.D;
        for (int e = tid; e < s.C * s.R; e += 256) {
            const int c = e / s.R, r = e % s.R;
            double* S = gsum + (size_t)((c * (c + 1) / 2 + c) * s.R + r) * s.st;
            S[0] += f * winsum[(size_t)e * wst];
            S[2 + 3 * s.D] += f * winsum[(size_t)e * wst + 1];
            for (int d = 0; d < s.D; ++d) S[3 + 3 * s.D + d] += f * winsum[(size_t)e * wst + 2 + d];
        }
        __syncthreads();
        if (tid < s.C) adj[tid] = 0.0;
        __syncthreads();
    }
    const int owners = n_chain_owners(s.kind, s.C, s.Q);
    for (int o = tid; o < owners; o += 256) chain_owner(s.kind, s.C, s.Q, s.D, params, comps, gsum, adj, o, out + 2);
}

cudaError_t launch_finalize(const KernSpec& s, const TileList* tl, int want_grad, const double* params,
                            const double* sigma, const double* comps, const double* chanbuf, const double* tile_part,
                            const double* z, const double* alpha, const double* kinv_diag, const double* logdet_part,
                            const int32_t* info, const int32_t* chan_dev, int64_t N, int64_t Np, double jitter_rel,
                            double* out, cudaStream_t st, const double* winsum) {
    (void)chanbuf;
    if (s.C > 64) return cudaErrorInvalidValue;
    if (s.window && want_grad && !winsum) return cudaErrorInvalidValue;
    const int stc = s.st;
    double* gsum = nullptr;
    if (want_grad) {
        const int npl = s.C * (s.C + 1) / 2;
        gsum = const_cast<double*>(tile_part) + (size_t)tl->n * s.R * stc;
        pairsum_kernel<<<npl, 128, 0, st>>>(s.R, stc, tl->pair_first_dev, tile_part, gsum);
        MOGP_COUNT(1);
    }
    finalize_kernel<<<1, 256, 0, st>>>(s, want_grad, params, sigma, comps, gsum, winsum, z, alpha, kinv_diag, logdet_part, info,
                                       chan_dev, (long long)N, (long long)Np, jitter_rel, out);
    MOGP_COUNT(1);
    return cudaGetLastError();
}

// ------------------------------------------------------------------ host-side self-check hooks
// Run the same closed forms on the CPU (used by the CPU test-suite; no device involved).
extern "C" int mogp_host_pair_comps(int kind, int C, int Q, int D, const double* params, double* comps_out) {
    KernSpec s;
    if (spec_init(s, kind, C, Q, D)) return -1;
    const int st = comp_stride(kind, D);
    for (int i = 0; i < C; ++i)
        for (int j = 0; j < C; ++j)
            for (int r = 0; r < s.R; ++r) pair_comp(kind, C, Q, D, params, i, j, r, comps_out + (size_t)((i * C + j) * s.R + r) * st);
    return s.R;
}
extern "C" int mogp_host_chain(int kind, int C, int Q, int D, const double* params, const double* gsum,
                               const double* adj, double* grad_out) {
    KernSpec s;
    if (spec_init(s, kind, C, Q, D)) return -1;
    const int st = comp_stride(kind, D);
    std::vector<double> comps((size_t)C * C * s.R * st);
    mogp_host_pair_comps(kind, C, Q, D, params, comps.data());
    const int owners = n_chain_owners(kind, C, Q);
    for (int o = 0; o < owners; ++o) chain_owner(kind, C, Q, D, params, comps.data(), gsum, adj, o, grad_out);
    return s.P;
}
