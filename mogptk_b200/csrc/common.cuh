// Internal declarations shared by the translation units of libmogp_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <atomic>
#include <mutex>
#include <string>
#include <vector>
#include "../../include/mogp_b200.h"

#define MOGP_TILE 64          // covariance tile edge (rows/cols of a channel-pair block)
#define MOGP_NB 64            // Cholesky leaf / inner panel width
#define MOGP_NB_OUT 256       // Cholesky outer panel width
#define MOGP_PAD 128          // internal matrices are padded to a multiple of this

// ------------------------------------------------------------------ block-component table
// Every supported kernel is K_ij[a,b] = sum_r alpha_r exp(-1/2 sum_d v_rd u_d^2) cos(2 pi (sum_d m_rd u_d + phi_r)),
// u_d = x_a,d - x_b,d + theta_rd, with per channel-pair (i,j) derived constants.
// comp record layout (doubles): [alpha, phi, v[D], m[D], theta[D]]; the harmonizable family (MOHSM) multiplies every
// component by a Gaussian window in the mid-point, exp(-1/2 l sum_d ((x_a,d + x_b,d)/2 - c_d)^2), and appends [l, c[D]].
__host__ __device__ inline int kind_family(int kind) { return kind & 0xff; }
__host__ __device__ inline int kind_rq(int kind) { return (kind >> 8) > 0 ? (kind >> 8) : 1; }
__host__ __device__ inline bool kind_window(int kind) { return kind_family(kind) == MOGP_KIND_MOHSM; }
__host__ __device__ inline int comp_stride(int kind, int D) { return kind_window(kind) ? 3 + 4 * D : 2 + 3 * D; }
#define MOGP_MAX_STRIDE (3 + 4 * MOGP_MAX_D)

struct KernSpec {
    int kind, C, Q, D;   // kind = family | Rq << 8 (MOGP_KIND_WITH_RQ)
    int R;            // components per channel pair (MOSM: Q, SM: Q*D, CONV: Q, CSM: Q*Rq, SMLMC: Q*D, UMOSM: Q)
    int P;            // packed constrained kernel parameters
    bool has_cos;     // false for CONV (m = phi = 0)
    int Rq;           // sub-components of CSM / SM-LMC (1 otherwise)
    bool window;      // MOHSM: non-stationary mid-point window (row-dependent Gram diagonal)
    int st;           // comp / gradient-sum record length
};

inline int spec_init(KernSpec& s, int kind, int C, int Q, int D) {
    if (C < 1 || Q < 1 || D < 1 || D > MOGP_MAX_D || kind < 0) return -1;
    const int Rq = kind_rq(kind);
    if (Rq > 64) return -1;
    s.kind = kind; s.C = C; s.Q = Q; s.D = D; s.Rq = Rq;
    switch (kind_family(kind)) {
        case MOGP_KIND_MOSM:  s.R = Q;      s.P = C * Q * (2 + 3 * D);                 s.has_cos = true;  break;
        case MOGP_KIND_SM:    s.R = Q * D;  s.P = C * Q * (1 + 2 * D);                 s.has_cos = true;  break;
        case MOGP_KIND_CONV:  s.R = Q;      s.P = Q * (C + C * D + D);                 s.has_cos = false; break;
        case MOGP_KIND_CSM:   s.R = Q * Rq; s.P = 2 * Q * C * Rq + 2 * Q * D;          s.has_cos = true;  break;
        case MOGP_KIND_SMLMC: s.R = Q * D;  s.P = C * Q * Rq + Q + 2 * Q * D;          s.has_cos = true;  break;
        case MOGP_KIND_UMOSM: s.R = Q;      s.P = Q * C * C + 3 * Q * C * D + Q * C;   s.has_cos = true;  break;
        case MOGP_KIND_MOHSM: s.R = Q;      s.P = Q * (3 * C + 3 * C * D + D);         s.has_cos = true;  break;
        default: return -1;
    }
    s.window = kind_window(kind);
    s.st = comp_stride(kind, D);
    if (kind_family(kind) < MOGP_KIND_CSM && (kind >> 8) != 0) return -1;
    return 0;
}

// One 64x64 (or ragged) tile of a channel-pair block.
struct CovTile {
    int pair;      // i*C + j
    int r0, c0;    // global first row / col
    int nr, nc;    // valid rows / cols (<= MOGP_TILE)
    int flags;     // bit0: tile lies on the global diagonal (r0 == c0, Gram mode)
};

struct TileList {
    std::vector<int32_t> off1, off2;   // key
    int mode;                          // 0 = Gram lower, 1 = Gram full (mirrored), 2 = cross
    std::vector<CovTile> host;
    CovTile* dev = nullptr;
    int n = 0;
    std::vector<int32_t> pair_first;   // Gram-lower: first tile of each lower pair (for the finalize pass)
    int32_t* pair_first_dev = nullptr;
};

// ------------------------------------------------------------------ GEMM descriptor
struct GemmArgs {
    const double* A; const double* B; double* C;
    long long lda, ldb, ldc;
    long long strideA, strideB, strideC;   // batch strides (elements)
    int M, N, K;
    int lower;      // 1: skip tiles strictly above the diagonal ((tm+1)*BM <= tn*BN)
    int klo_mode;   // 0: 0        1: tn*BN     2: tm*BM
    int khi_mode;   // 0: K        1: min(K,(tm+1)*BM)
    int epi;        // 0: C = alpha*acc + beta*C      1: C = 0.5*((alpha*acc + beta*C) - avec[row]*avec[col])
    int pair;       // 0: one tile per CTA   1: also tile (ntn-1-tn) (with klo_mode 1)   2: also tile (ntm-1-tm) (khi_mode 1)
    int first_touch_row1; // 0: off; else tiles whose first row is >= (value - 1) use beta = 0 (first touch)
    int first_touch_col1; // same for tiles whose first column is >= (value - 1)
    int prio;       // 0: the stream's priority; p > 0: launch attribute priority -(p - 1) (1 = lowest ... 6 = highest), which
                    // -- unlike a stream priority -- is also recorded in a captured graph's kernel node
    double alpha, beta;
    const double* avec;
};

// transa: 0 -> A stored [m][k] (k contiguous), 1 -> A stored [k][m]
// transb: 0 -> B stored [k][n] (n contiguous), 1 -> B stored [n][k]
cudaError_t launch_gemm(int transa, int transb, const GemmArgs& g, int batch, cudaStream_t s);

// ------------------------------------------------------------------ handle
struct I8Plan;                 // int8 tensor-pipe GEMM state (i8mm.cu)
struct StepGraph {             // one captured exact-GP step (see capi.cu)
    int kind = 0, C = 0, Q = 0, D = 0, want_grad = 0, has_dv = 0, uses = 0;
    int64_t N = 0;
    double jitter = 0.0;
    long long launches = 0, epoch = 0, cfg_epoch = 0;
    std::vector<int32_t> off;
    cudaGraphExec_t exec = nullptr;
};
struct PotrfStreams {          // look-ahead resources owned by the handle
    cudaStream_t s1 = nullptr;    // high priority: the panel chain
    cudaStream_t s2 = nullptr;    // low priority: bulk trailing updates
    cudaStream_t s3 = nullptr;    // high priority: inner updates of the two-level variant
    cudaEvent_t* ev3 = nullptr;   // [nev + 2] recorded on s3
    cudaEvent_t* ev1 = nullptr;   // [nev + 2] recorded on the caller's stream after panel steps
    cudaEvent_t* ev2 = nullptr;   // [nev + 2] recorded on s2 after bulk updates
    cudaStream_t s4 = nullptr;    // medium priority: triangular inverse pipelined behind the panel chain
    cudaEvent_t* evp = nullptr;   // [nev + 2] recorded on s1 after every panel step (pipelined inverse)
    cudaStream_t sl[8] = {};      // medium priority: one stream per doubling level of the pipelined inverse
    cudaEvent_t* evq = nullptr;   // [nevq] completion events of the pipelined inverse's operations (+ 8 join events)
    cudaEvent_t ev_kinv = nullptr; // row-wise pipeline: K^-1 (accumulated behind the panel chain) is complete
    int nevq = 0;
    int nev = 0;
};
struct mogp_handle_s {
    int device = 0;
    int64_t max_n = 0, np_max = 0;
    double *A = nullptr, *Linv = nullptr, *W = nullptr;      // np_max^2 each
    double *T = nullptr; size_t T_cap = 0;                    // scratch of the row-wise pipelined inverse (Np^2, small sizes only)
    double *comps = nullptr; size_t comps_cap = 0;            // C*C*R*stride (state of the last lml_grad)
    double *chanbuf = nullptr;                                // per-channel scalars of the last lml_grad
    double *comps2 = nullptr; size_t comps2_cap = 0;          // same, scratch of mogp_kbuild / mogp_kdiag
    double *winsum = nullptr; size_t winsum_cap = 0;          // window kinds: row sums of the Gram diagonal (prep -> finalize)
    double *winsum2 = nullptr; size_t winsum2_cap = 0;        // same, scratch of mogp_kbuild
    double *chanbuf2 = nullptr;
    int64_t linv_np = 0;                                      // layout (ld) Linv was last zero-initialised for
    double *vec = nullptr;                                    // 8 * np_max doubles of vector scratch
    double *colpart = nullptr; size_t colpart_cap = 0;        // column-pass partials
    double *tile_part = nullptr; size_t tile_part_cap = 0;    // gradient tile partials
    double *xbuf = nullptr; size_t xbuf_cap = 0;              // copy of training x (N*D)
    double *pbuf = nullptr; size_t pbuf_cap = 0;              // host-entry staging (params, sigma, y, dv)
    double *out_dev = nullptr; size_t out_cap = 0;
    double *pred_K = nullptr, *pred_V = nullptr, *pred_S = nullptr; size_t pred_cap = 0, pred_s_cap = 0;
    double *logdet_part = nullptr;                            // np_max/64
    int32_t *info = nullptr;
    int32_t *chan_dev = nullptr;                              // device copy of chan_off (C+1)
    std::vector<TileList*> tiles;
    // state of the last factorisation (for predict)
    bool have_factor = false;
    KernSpec spec{};
    std::vector<int32_t> chan_off;
    int64_t N = 0, Np = 0;
    std::string err;
    PotrfStreams ps;
    std::vector<StepGraph*> graphs;
    long long realloc_epoch = 0;
    void *pent_dev = nullptr, *pent_host = nullptr; size_t pent_cap = 0; int pent_n = 0;   // parameter-entry table
    cudaStream_t hs = nullptr;                                // the step runs here in graph mode
    cudaEvent_t ev_in = nullptr, ev_out = nullptr, ev_f1 = nullptr, ev_f2 = nullptr;
    std::vector<int32_t> chan_uploaded;                       // content of chan_dev slot 0
    double *gbuf = nullptr; size_t gbuf_cap = 0;              // graph staging: params | sigma | y | data_var | out
    I8Plan* i8 = nullptr;                                     // int8 tensor-pipe GEMM state (large problems)
    // early loss: [lml, info, seq] in mapped pinned host memory, written by the step right after the solves (mogp_early_loss)
    double* early_host = nullptr; unsigned long long* early_ctr = nullptr; bool early_on = false;
    unsigned long long early_expected = 0;
    // optional stage timing (mogp_set_profile): events at the stage boundaries of mogp_lml_grad
    bool profile = false;
    cudaEvent_t ev[8] = {};
    int n_ev = 0;
};

#define MOGP_CHECK(h, expr)                                                          \
    do {                                                                             \
        cudaError_t _e = (expr);                                                     \
        if (_e != cudaSuccess) {                                                     \
            (h)->err = std::string(#expr) + ": " + cudaGetErrorString(_e);           \
            return -2;                                                               \
        }                                                                            \
    } while (0)

// Kernel attributes (dynamic shared-memory opt-in, carve-out) are per DEVICE: a process that drives several GPUs
// (gpr.use_gpu(n), one Engine per device) has to set them once on each.  `static PerDeviceOnce once; if (once.first()) ...`
struct PerDeviceOnce {
    std::mutex mu;
    bool done[64] = {};
    int sm_count[64] = {};
    int device() const { int d = 0; cudaGetDevice(&d); return d & 63; }
    int sms() {
        const int d = device();
        std::lock_guard<std::mutex> lk(mu);
        if (sm_count[d] <= 0 && (cudaDeviceGetAttribute(&sm_count[d], cudaDevAttrMultiProcessorCount, d) != cudaSuccess || sm_count[d] <= 0))
            sm_count[d] = 148;
        return sm_count[d];
    }
};
// `if (OnceGuard og{once}; og.needed()) { set the attributes }`: the lock is held while they are being set, so that a second
// host thread (replicas.train_restarts runs one per model) cannot launch the kernel before its attributes exist.
struct OnceGuard {
    PerDeviceOnce& o; int d; bool need;
    explicit OnceGuard(PerDeviceOnce& o_) : o(o_), d(o_.device()) { o.mu.lock(); need = !o.done[d]; }
    bool needed() const { return need; }
    ~OnceGuard() { if (need) o.done[d] = true; o.mu.unlock(); }
};

// number of kernels launched by this library since load (bench.py reports it as gpu_launches)
extern std::atomic<long long> g_mogp_launches;
extern long long g_mogp_cfg_epoch;
#define MOGP_COUNT(n) (g_mogp_launches += (n))

// ------------------------------------------------------------------ covariance kernels (cov.cu)
// chanbuf layout (doubles): [0..C) kdiag_gram | [C..2C) kdiag_api | [2C..3C) noise variance | [3C] jitter add | [3C+1] trW scratch
// window kinds (MOHSM) also need the N x D inputs and a C * R * (2 + D) buffer for the row sums of the Gram diagonal
cudaError_t launch_prep(const KernSpec& s, const double* params, const double* sigma, const double* data_var,
                        const int32_t* chan_dev, int64_t N, double jitter_rel, double* comps, double* chanbuf,
                        cudaStream_t st, const double* x = nullptr, double* winsum = nullptr);
// mode 0: Gram lower into padded A (ld = Np), also initialises the padding; mode 1: Gram full; mode 2: cross
cudaError_t launch_kbuild(const KernSpec& s, const TileList& tl, const double* comps, const double* chanbuf,
                          const double* x1, const double* x2, const int32_t* chan1_dev, const double* data_var,
                          int add_diag, double* K, long long ldk, int64_t N, int64_t Np, cudaStream_t st);
cudaError_t launch_kdiag(const KernSpec& s, const double* chanbuf, const int32_t* chan_dev, int64_t N, double* out,
                         cudaStream_t st, const double* comps = nullptr, const double* x = nullptr);
// avec != NULL: W holds K^-1 and the kernel forms (K^-1 - avec avec^T)/2 while loading
cudaError_t launch_grad_reduce(const KernSpec& s, const TileList& tl, const double* comps, const double* x,
                               const double* W, long long ldw, const double* avec, double* tile_part, cudaStream_t st);
// out: [0]=lml [1]=info [2..2+P) grad params [2+P..2+P+C) grad sigma
cudaError_t launch_finalize(const KernSpec& s, const TileList* tl, int want_grad, const double* params,
                            const double* sigma, const double* comps, const double* chanbuf,
                            const double* tile_part, const double* z, const double* alpha, const double* kinv_diag,
                            const double* logdet_part, const int32_t* info, const int32_t* chan_dev,
                            int64_t N, int64_t Np, double jitter_rel, double* out, cudaStream_t st,
                            const double* winsum = nullptr);

// ------------------------------------------------------------------ dense linear algebra (linalg.cu)
// fused_inverse: NULL -> factor only (diagonal blocks of Linv get inv(L_kk)); else the triangular inverse may be
// pipelined behind the panel chain (needs Ltmp == the scratch trtri_padded would use, ldt == ldi == lda); on return
// *fused_inverse tells whether Linv is complete (true) or trtri_padded still has to run (false).
// z = L^-1 y carried along the row-wise pipeline: as soon as a row block of Linv is complete its entries of z follow, and behind
// the last block the early loss [lml, info, seq] (see launch_lml_early) -- instead of one triangular mat-vec after the join.
struct ZChain {
    const double* ypad; double* z;       // Np each (padding rows of y zero)
    int64_t N;                           // true row count (constant of the log marginal likelihood)
    double* early_host; unsigned long long* early_ctr;   // NULL: no early loss
    bool done;                           // out: z (and the early loss) were produced inside potrf_padded
};
// Kacc != NULL (a further Np x Np buffer, ld = lda, distinct from Ltmp) allows the row-wise pipeline (small sizes): Linv
// is built row group by row group behind the panel chain and K^-1 = Linv^T Linv (lower) is accumulated into Kacc by
// rank updates as the rows complete; *fused_kinv then says that Kacc is (will be) complete once ps->ev_kinv -- recorded
// before the call returns -- has been reached, and the caller must make its stream wait for that event.
cudaError_t potrf_padded(double* A, long long lda, double* Linv, long long ldi, double* Ltmp, long long ldt,
                         int64_t Np, double* logdet_part, int32_t* info, cudaStream_t st, const PotrfStreams* ps,
                         bool* fused_inverse = nullptr, I8Plan* i8 = nullptr, int i8_slices = 7, double* Kacc = nullptr,
                         bool* fused_kinv = nullptr, ZChain* zc = nullptr);
cudaError_t launch_trmv_rows(const double* Linv, long long ld, const double* y, double* z, int64_t row0, int64_t row1,
                             cudaStream_t st);
bool rowpipe_applies(int64_t Np);
cudaError_t trtri_padded(double* A /*L*/, double* Linv, double* scratch, int64_t Np, long long ld, cudaStream_t st,
                         I8Plan* i8 = nullptr, int i8_slices = 7);
cudaError_t kinv_padded(const double* Linv, double* W, int64_t Np, long long ld, const double* avec, cudaStream_t st);
// z = Linv * y (lower-triangular mat-vec), rows [0,Np)
cudaError_t launch_trmv_lower(const double* Linv, long long ld, const double* y, double* z, int64_t Np, cudaStream_t st);
// column pass: out_dot[c] = sum_r M[r][c]*v[r], out_sq[c] = sum_r M[r][c]^2, rows [0,rows), cols [0,cols)
cudaError_t launch_colpass(const double* M, long long ld, const double* v, int64_t rows, int64_t cols,
                           double* part, size_t part_cap, double* out_dot, double* out_sq, cudaStream_t st);
cudaError_t launch_pad_copy(const double* src, int64_t n, double* dst, int64_t np, cudaStream_t st);
cudaError_t launch_copy_tri(int dir, double* user, long long ldu, double* work, long long ldw, int64_t n, int64_t np,
                            cudaStream_t st);
// var = prior variance - colsq; the prior variance is the per-channel constant chanbuf[C + c] or, if kss != NULL, kss[m]
cudaError_t launch_pred_var(const double* chanbuf, int C, const int32_t* chan_s_dev, const double* colsq, int64_t M,
                            double* var, cudaStream_t st, const double* kss = nullptr);
cudaError_t run_peak_fp64(double* dmma_tflops, double* dfma_tflops);
// [lml, info, sequence number] into mapped pinned host memory right after the solves (see cov.cu)
cudaError_t launch_lml_early(const double* z, const double* logdet_part, const int32_t* info, int64_t N, int64_t Np,
                             double* host_out_dev, unsigned long long* counter, cudaStream_t st);
cudaError_t launch_stamp(int slot, cudaStream_t st);      // diagnostics: global-timer stamp of a point of the step (mogp_set_stamps)

// ------------------------------------------------------------------ fp64 GEMM on the int8 tensor pipe (i8mm.cu)
struct I8Plan;
I8Plan* i8_plan_create();
void i8_plan_destroy(I8Plan* p);
// host-side preparation (allocation, tile lists) for a given padded size / leading dimension: call outside graph capture
// (*changed = true when the device-side tile lists were rebuilt: graphs that replay the old ones are stale)
cudaError_t i8_prepare(I8Plan* p, int64_t Np, long long ld, int S, cudaStream_t st, bool* changed = nullptr);
// one doubling level (block size S_ rows) of Linv = L^-1; cudaErrorNotSupported -> run the DMMA GEMMs instead
cudaError_t i8_trtri_level(I8Plan* p, const double* L, double* Linv, double* scratch, int64_t Np, long long ld, int64_t S_,
                           int S, cudaStream_t st);
extern long long g_i8_trtri_min;
// rank-256 trailing update of the blocked Cholesky (see i8mm.cu); smallest padded size that uses it (0 = never)
cudaError_t i8_syrk_update(I8Plan* p, double* A, long long ld, int64_t r0, int64_t k0, int64_t Np, int part, int S,
                           cudaStream_t st);
extern long long g_i8_potrf_min;
// recursive Cholesky + inverse (linalg.cu: rchol_padded): the four int8-pipe products of one 2h x 2h diagonal block (i8mm.cu)
bool i8_blk_ok(const I8Plan* p, int64_t Np, long long ld, int S, int64_t h);
struct I8BlkAsync {                // overlap of a block's trailing products with the recursion into its second half
    cudaStream_t side; cudaEvent_t ev_fork, ev_rest, ev_T;
    int64_t split_rows;            // > 0: the second half needs only these rows of A22 at once (its first leaf)
    int own_ops;                   // 1: operands in buffers of their own (the second half uses the shared ones)
};
cudaError_t i8_blk_first(I8Plan* p, double* A, const double* Linv, double* scratch, long long ld, int64_t o, int64_t h,
                         int want_inv, int S, cudaStream_t st, const I8BlkAsync* as = nullptr);
cudaError_t i8_blk_second(I8Plan* p, double* Linv, const double* scratch, long long ld, int64_t o, int64_t h, int S,
                          cudaStream_t st);
// Recursive factor (+ inverse) for padded sizes leaf * 2^k: leaves by the blocked sweep with its pipelined inverse, everything above
// them by the products above.  want_inverse = 0: only what the factor itself needs (Linv is partially filled).  Returns
// cudaErrorNotSupported when the size / plan does not qualify (the caller runs the blocked sweep).
cudaError_t rchol_padded(double* A, long long ld, double* Linv, double* Ltmp, int64_t Np, double* logdet_part, int32_t* info,
                         cudaStream_t st, const PotrfStreams* ps, I8Plan* i8, int i8_slices, int want_inverse);
bool rchol_applies(int64_t Np);
int64_t rchol_leaf_for(int64_t Np);      // leaf rows of the recursive scheme for a padded size (0: does not apply)
// W(lower tiles) = Linv^T Linv via tcgen05.mma kind::i8 (S digit planes of 7 bits); pure enqueue
cudaError_t i8_kinv(I8Plan* p, const double* Linv, double* W, int64_t Np, long long ld, int S, cudaStream_t st);
// Smallest padded size that takes the int8 path (0 = never) and the number of digit planes (7 or 8)
extern long long g_i8_min_np;
extern int g_i8_slices;
