// C ABI of libmogp_b200.so: workspace handle, tile lists, the fused exact-GP step.
// See include/mogp_b200.h for the contract and the reference interfaces each entry replaces.
#include "common.cuh"
#include <algorithm>
#include <cstring>
#include <cstdlib>

#define MOGP_VERSION 1000

static int64_t round_up(int64_t a, int64_t b) { return (a + b - 1) / b * b; }

#define H_ARG(h, cond, msg)             \
    do {                                \
        if (!(cond)) {                  \
            (h)->err = (msg);           \
            return -1;                  \
        }                               \
    } while (0)

template <typename T>
static int ensure(mogp_handle_s* h, T*& ptr, size_t& cap, size_t need) {
    if (need <= cap && ptr) return 0;
    if (ptr) MOGP_CHECK(h, cudaFree(ptr));
    ptr = nullptr; cap = 0;
    size_t want = need + need / 4 + 64;
    MOGP_CHECK(h, cudaMalloc(&ptr, want * sizeof(T)));
    cap = want;
    h->realloc_epoch++;          // captured graphs hold the old pointer: they are re-captured on next use
    return 0;
}

// Padded size of an N-row problem (padding rows carry a unit diagonal).  Normally the next multiple of 128; large problems get a few
// more rows (at most ~3 %) when that makes the size leaf * 2^k and thereby eligible for the recursive factor + inverse, which is
// 12-25 % faster than the blocked sweep (profiles/r02_recursive_cholesky.txt) against (1.03)^3 = 9 % more flops at worst.
extern long long g_i8_min_np;
static int g_pad_for_rchol = std::getenv("MOGP_PAD_FOR_RCHOL") ? std::atoi(std::getenv("MOGP_PAD_FOR_RCHOL")) : 1;
extern "C" int mogp_set_pad_for_rchol(int on) { g_pad_for_rchol = on; ++g_mogp_cfg_epoch; return 0; }
static int64_t padded_size(int64_t N, int64_t limit, int64_t min_np = 0) {
    const int64_t Np = round_up(N, MOGP_PAD);
    if (!g_pad_for_rchol || g_i8_min_np <= 0 || rchol_leaf_for(Np) > 0) return Np;
    for (int64_t q : {256, 512, 1024}) {
        const int64_t c = round_up(N, q);
        if (c >= min_np && c >= g_i8_min_np && (limit <= 0 || c <= limit) && (c - N) * 33 <= N && rchol_leaf_for(c) > 0) return c;
    }
    return Np;
}
extern "C" long long mogp_padded_size(long long N) { return padded_size(N, 0); }

extern "C" int mogp_version(void) { return MOGP_VERSION; }

extern "C" int mogp_num_params(int kind, int C, int Q, int D) {
    KernSpec s;
    if (spec_init(s, kind, C, Q, D)) return -1;
    return s.P;
}

extern "C" int mogp_create(int device, int64_t max_n, mogp_handle_t* out) {
    if (!out || max_n < 1) return -1;
    *out = nullptr;
    cudaError_t e = cudaSetDevice(device);
    if (e != cudaSuccess) return -2;
    mogp_handle_s* h = new mogp_handle_s();
    h->device = device;
    h->max_n = max_n;
    h->np_max = padded_size(max_n, 0);
    const size_t nn = (size_t)h->np_max * h->np_max;
    auto fail = [&](cudaError_t err) { (void)err; mogp_destroy(h); return -2; };
    if ((e = cudaMalloc(&h->A, nn * 8)) != cudaSuccess) return fail(e);
    if ((e = cudaMalloc(&h->Linv, nn * 8)) != cudaSuccess) return fail(e);
    if ((e = cudaMalloc(&h->W, nn * 8)) != cudaSuccess) return fail(e);
    // invariant: everything above the diagonal blocks of Linv is zero (GEMM k-clipping relies on it)
    if ((e = cudaMemset(h->Linv, 0, nn * 8)) != cudaSuccess) return fail(e);
    if ((e = cudaMalloc(&h->vec, 8 * (size_t)h->np_max * 8)) != cudaSuccess) return fail(e);
    if ((e = cudaMalloc(&h->chanbuf, 1024 * 8)) != cudaSuccess) return fail(e);
    if ((e = cudaMalloc(&h->chanbuf2, 1024 * 8)) != cudaSuccess) return fail(e);
    h->linv_np = h->np_max;
    if ((e = cudaMalloc(&h->logdet_part, (size_t)(h->np_max / 64 + 1) * 8)) != cudaSuccess) return fail(e);
    if ((e = cudaMalloc(&h->info, 64)) != cudaSuccess) return fail(e);
    if ((e = cudaMalloc(&h->chan_dev, 4 * 260 * 2)) != cudaSuccess) return fail(e);
    if ((e = cudaStreamCreateWithFlags(&h->hs, cudaStreamNonBlocking)) != cudaSuccess) return fail(e);
    if ((e = cudaEventCreateWithFlags(&h->ev_in, cudaEventDisableTiming)) != cudaSuccess) return fail(e);
    if ((e = cudaEventCreateWithFlags(&h->ev_out, cudaEventDisableTiming)) != cudaSuccess) return fail(e);
    if ((e = cudaEventCreateWithFlags(&h->ev_f1, cudaEventDisableTiming)) != cudaSuccess) return fail(e);
    if ((e = cudaEventCreateWithFlags(&h->ev_f2, cudaEventDisableTiming)) != cudaSuccess) return fail(e);
    {   // look-ahead stream + events for the blocked Cholesky
        const int nev = (int)(h->np_max / 64) + 2;
        int plo = 0, phi = 0;
        cudaDeviceGetStreamPriorityRange(&plo, &phi);      // bulk updates yield to the panel chain
        if ((e = cudaStreamCreateWithPriority(&h->ps.s2, cudaStreamNonBlocking, plo)) != cudaSuccess) return fail(e);
        if ((e = cudaStreamCreateWithPriority(&h->ps.s1, cudaStreamNonBlocking, phi)) != cudaSuccess) return fail(e);
        if ((e = cudaStreamCreateWithPriority(&h->ps.s3, cudaStreamNonBlocking, phi)) != cudaSuccess) return fail(e);
        h->ps.ev1 = new cudaEvent_t[nev + 2]();
        h->ps.ev2 = new cudaEvent_t[nev + 2]();
        h->ps.ev3 = new cudaEvent_t[nev + 2]();
        h->ps.evp = new cudaEvent_t[nev + 2]();
        if ((e = cudaStreamCreateWithPriority(&h->ps.s4, cudaStreamNonBlocking, plo + (phi - plo) / 2)) != cudaSuccess) return fail(e);
        for (int i = 0; i < 8; ++i)
            if ((e = cudaStreamCreateWithPriority(&h->ps.sl[i], cudaStreamNonBlocking, plo + (phi - plo) / 2)) != cudaSuccess) return fail(e);
        h->ps.nevq = 4 * nev + 16;
        if ((e = cudaEventCreateWithFlags(&h->ps.ev_kinv, cudaEventDisableTiming)) != cudaSuccess) return fail(e);
        h->ps.evq = new cudaEvent_t[h->ps.nevq]();
        for (int i = 0; i < h->ps.nevq; ++i)
            if ((e = cudaEventCreateWithFlags(&h->ps.evq[i], cudaEventDisableTiming)) != cudaSuccess) return fail(e);
        for (int i = 0; i < nev + 2; ++i) {
            if ((e = cudaEventCreateWithFlags(&h->ps.evp[i], cudaEventDisableTiming)) != cudaSuccess) return fail(e);
            if ((e = cudaEventCreateWithFlags(&h->ps.ev1[i], cudaEventDisableTiming)) != cudaSuccess) return fail(e);
            if ((e = cudaEventCreateWithFlags(&h->ps.ev2[i], cudaEventDisableTiming)) != cudaSuccess) return fail(e);
            if ((e = cudaEventCreateWithFlags(&h->ps.ev3[i], cudaEventDisableTiming)) != cudaSuccess) return fail(e);
        }
        h->ps.nev = nev;
    }
    h->colpart_cap = (size_t)64 * 2 * h->np_max;
    if ((e = cudaMalloc(&h->colpart, h->colpart_cap * 8)) != cudaSuccess) return fail(e);
    *out = h;
    return 0;
}

extern "C" int mogp_destroy(mogp_handle_t h) {
    if (!h) return 0;
    cudaSetDevice(h->device);
    cudaDeviceSynchronize();
    for (TileList* t : h->tiles) {
        if (t->dev) cudaFree(t->dev);
        if (t->pair_first_dev) cudaFree(t->pair_first_dev);
        delete t;
    }
    if (h->i8) i8_plan_destroy(h->i8);
    if (h->pent_dev) cudaFree(h->pent_dev);
    if (h->pent_host) cudaFreeHost(h->pent_host);
    if (h->early_host) cudaFreeHost(h->early_host);
    if (h->early_ctr) cudaFree(h->early_ctr);
    for (StepGraph* g : h->graphs) {
        if (g->exec) cudaGraphExecDestroy(g->exec);
        delete g;
    }
    if (h->ps.ev1)
        for (int i = 0; i < h->ps.nev + 2; ++i) {
            if (h->ps.ev1[i]) cudaEventDestroy(h->ps.ev1[i]);
            if (h->ps.ev2[i]) cudaEventDestroy(h->ps.ev2[i]);
            if (h->ps.ev3 && h->ps.ev3[i]) cudaEventDestroy(h->ps.ev3[i]);
            if (h->ps.evp && h->ps.evp[i]) cudaEventDestroy(h->ps.evp[i]);
        }
    delete[] h->ps.evp;
    if (h->ps.evq)
        for (int i = 0; i < h->ps.nevq; ++i)
            if (h->ps.evq[i]) cudaEventDestroy(h->ps.evq[i]);
    delete[] h->ps.evq;
    if (h->ps.ev_kinv) cudaEventDestroy(h->ps.ev_kinv);
    for (int i = 0; i < 8; ++i)
        if (h->ps.sl[i]) cudaStreamDestroy(h->ps.sl[i]);
    delete[] h->ps.ev1;
    delete[] h->ps.ev2;
    delete[] h->ps.ev3;
    if (h->hs) cudaStreamDestroy(h->hs);
    if (h->ev_in) cudaEventDestroy(h->ev_in);
    if (h->ev_out) cudaEventDestroy(h->ev_out);
    if (h->ev_f1) cudaEventDestroy(h->ev_f1);
    if (h->ev_f2) cudaEventDestroy(h->ev_f2);
    if (h->ps.s2) cudaStreamDestroy(h->ps.s2);
    if (h->ps.s1) cudaStreamDestroy(h->ps.s1);
    if (h->ps.s3) cudaStreamDestroy(h->ps.s3);
    if (h->ps.s4) cudaStreamDestroy(h->ps.s4);
    void* ptrs[] = {h->T, h->A, h->Linv, h->W, h->vec, h->chanbuf, h->logdet_part, h->info, h->chan_dev, h->colpart,
                    h->comps, h->comps2, h->chanbuf2, h->winsum, h->winsum2, h->gbuf, h->tile_part, h->xbuf, h->pbuf, h->out_dev, h->pred_K, h->pred_V, h->pred_S};
    for (void* p : ptrs)
        if (p) cudaFree(p);
    delete h;
    return 0;
}

extern "C" const char* mogp_last_error(mogp_handle_t h) { return h ? h->err.c_str() : "null handle"; }

// ------------------------------------------------------------------ tile lists
static TileList* get_tiles(mogp_handle_s* h, int C, const int32_t* off1, const int32_t* off2, int mode, cudaStream_t st) {
    std::vector<int32_t> o1(off1, off1 + C + 1), o2;
    if (off2) o2.assign(off2, off2 + C + 1);
    for (size_t k = 0; k < h->tiles.size(); ++k) {
        TileList* t = h->tiles[k];
        if (t->mode == mode && t->off1 == o1 && t->off2 == o2) {
            // least-recently-used order: a hit moves the list to the back
            h->tiles.erase(h->tiles.begin() + k);
            h->tiles.push_back(t);
            return t;
        }
    }
    TileList* t = new TileList();
    t->mode = mode; t->off1 = o1; t->off2 = o2;
    const int T = MOGP_TILE;
    if (mode == 2) {
        for (int i = 0; i < C; ++i)
            for (int j = 0; j < C; ++j)
                for (int r = o1[i]; r < o1[i + 1]; r += T)
                    for (int c = o2[j]; c < o2[j + 1]; c += T)
                        t->host.push_back({i * C + j, r, c, std::min(T, o1[i + 1] - r), std::min(T, o2[j + 1] - c), 0});
    } else {
        for (int i = 0; i < C; ++i)
            for (int j = 0; j <= i; ++j) {
                t->pair_first.push_back((int32_t)t->host.size());
                for (int r = o1[i]; r < o1[i + 1]; r += T)
                    for (int c = o1[j]; c < o1[j + 1]; c += T) {
                        if (i == j && c > r) continue;
                        t->host.push_back({i * C + j, r, c, std::min(T, o1[i + 1] - r), std::min(T, o1[j + 1] - c),
                                           (i == j && r == c) ? 1 : 0});
                    }
            }
        t->pair_first.push_back((int32_t)t->host.size());
    }
    t->n = (int)t->host.size();
    if (t->n > 0) {
        if (cudaMalloc(&t->dev, t->host.size() * sizeof(CovTile)) != cudaSuccess) { delete t; return nullptr; }
        cudaMemcpyAsync(t->dev, t->host.data(), t->host.size() * sizeof(CovTile), cudaMemcpyHostToDevice, st);
    }
    if (!t->pair_first.empty()) {
        if (cudaMalloc(&t->pair_first_dev, t->pair_first.size() * 4) != cudaSuccess) { delete t; return nullptr; }
        cudaMemcpyAsync(t->pair_first_dev, t->pair_first.data(), t->pair_first.size() * 4, cudaMemcpyHostToDevice, st);
    }
    if (h->tiles.size() >= 64) {
        // Bounded cache, least recently used first.  Captured step graphs (StepGraph) have the device arrays of the
        // training lists (mode 0) baked in, so those are evicted last, and when one does go every captured graph is
        // invalidated through realloc_epoch (re-captured on next use) -- replaying a graph over freed tiles would
        // silently compute a wrong K.
        size_t victim = h->tiles.size();
        for (size_t k = 0; k + 1 < h->tiles.size(); ++k)     // (never the most recently used one: the caller may hold it)
            if (h->tiles[k]->mode != 0) { victim = k; break; }
        if (victim == h->tiles.size()) victim = 0;
        TileList* old = h->tiles[victim];
        cudaStreamSynchronize(st);
        cudaDeviceSynchronize();              // kernels on the handle's own streams may still read the list
        if (old->mode == 0) h->realloc_epoch++;
        if (old->dev) cudaFree(old->dev);
        if (old->pair_first_dev) cudaFree(old->pair_first_dev);
        delete old;
        h->tiles.erase(h->tiles.begin() + victim);
    }
    h->tiles.push_back(t);
    return t;
}

static int check_offsets(mogp_handle_s* h, int C, const int32_t* off) {
    H_ARG(h, off != nullptr, "chan_off is NULL");
    H_ARG(h, off[0] == 0, "chan_off[0] must be 0");
    for (int c = 0; c < C; ++c) H_ARG(h, off[c + 1] >= off[c], "chan_off must be non-decreasing");
    return 0;
}

static int upload_chan(mogp_handle_s* h, int C, const int32_t* off, int slot, cudaStream_t st) {
    MOGP_CHECK(h, cudaMemcpyAsync(h->chan_dev + slot * 260, off, (C + 1) * 4, cudaMemcpyHostToDevice, st));
    return 0;
}

static int prep_common(mogp_handle_s* h, KernSpec& s, int kind, int C, int Q, int D, const double* params_dev,
                       const double* sigma_dev, const double* data_var_dev, const int32_t* off, int64_t N,
                       double jitter_rel, bool scratch, cudaStream_t st, const double* x_dev = nullptr) {
    H_ARG(h, spec_init(s, kind, C, Q, D) == 0, "bad kernel spec (kind, C, Q, D)");
    H_ARG(h, C <= 64, "at most 64 channels");
    H_ARG(h, (size_t)(3 * C + 2) <= 1024, "too many channels");
    if (check_offsets(h, C, off)) return -1;
    if (upload_chan(h, C, off, 0, st)) return -2;
    h->chan_uploaded.assign(off, off + C + 1);
    double*& comps = scratch ? h->comps2 : h->comps;
    size_t& cap = scratch ? h->comps2_cap : h->comps_cap;
    if (ensure(h, comps, cap, (size_t)C * C * s.R * s.st)) return -2;
    double*& ws = scratch ? h->winsum2 : h->winsum;
    size_t& wcap = scratch ? h->winsum2_cap : h->winsum_cap;
    if (s.window) {
        H_ARG(h, x_dev != nullptr || N == 0, "this kernel is non-stationary (MOHSM): the inputs are required");
        if (ensure(h, ws, wcap, (size_t)C * s.R * (2 + D))) return -2;
    }
    MOGP_CHECK(h, launch_prep(s, params_dev, sigma_dev, data_var_dev, h->chan_dev, N, jitter_rel, comps,
                              scratch ? h->chanbuf2 : h->chanbuf, st, x_dev, s.window ? ws : nullptr));
    return 0;
}

// Linv relies on "zero above the diagonal blocks" for its current leading dimension.
static int zero_linv_for(mogp_handle_s* h, int64_t Np, cudaStream_t st) {
    if (h->linv_np == Np) return 0;
    MOGP_CHECK(h, cudaMemsetAsync(h->Linv, 0, (size_t)Np * Np * 8, st));
    h->linv_np = Np;
    h->have_factor = false;
    return 0;
}

// ------------------------------------------------------------------ K(X1, X2), K_diag
extern "C" int mogp_kbuild(mogp_handle_t h, int kind, int C, int Q, int D, const double* params_dev,
                           const double* x1_dev, const int32_t* chan_off1_host, const double* x2_dev,
                           const int32_t* chan_off2_host, const double* noise_sigma_dev, const double* data_var_dev,
                           double jitter_rel, double* K_dev, int64_t ldk, void* stream) {
    if (!h) return -1;
    cudaStream_t st = (cudaStream_t)stream;
    MOGP_CHECK(h, cudaSetDevice(h->device));
    KernSpec s;
    if (check_offsets(h, C, chan_off1_host)) return -1;
    const int64_t N1 = chan_off1_host[C];
    const bool gram = (x2_dev == nullptr);
    if (!gram && check_offsets(h, C, chan_off2_host)) return -1;
    const int add_diag = gram && (noise_sigma_dev || data_var_dev || jitter_rel != 0.0);
    int rc = prep_common(h, s, kind, C, Q, D, params_dev, noise_sigma_dev, gram ? data_var_dev : nullptr, chan_off1_host,
                         N1, gram ? jitter_rel : 0.0, true, st, x1_dev);
    if (rc) return rc;
    TileList* tl = get_tiles(h, C, chan_off1_host, gram ? nullptr : chan_off2_host, gram ? 1 : 2, st);
    H_ARG(h, tl != nullptr, "tile list allocation failed");
    MOGP_CHECK(h, launch_kbuild(s, *tl, h->comps2, h->chanbuf2, x1_dev, gram ? nullptr : x2_dev, h->chan_dev,
                                gram ? data_var_dev : nullptr, add_diag, K_dev, ldk, N1, N1, st));
    return 0;
}

extern "C" int mogp_kdiag_x(mogp_handle_t h, int kind, int C, int Q, int D, const double* params_dev, const double* x_dev,
                            const int32_t* chan_off_host, double* out_dev, void* stream) {
    if (!h) return -1;
    cudaStream_t st = (cudaStream_t)stream;
    MOGP_CHECK(h, cudaSetDevice(h->device));
    KernSpec s;
    if (check_offsets(h, C, chan_off_host)) return -1;
    const int64_t N = chan_off_host[C];
    int rc = prep_common(h, s, kind, C, Q, D, params_dev, nullptr, nullptr, chan_off_host, std::max<int64_t>(N, 1), 0.0, true, st,
                         x_dev);
    if (rc) return rc;
    if (N > 0) MOGP_CHECK(h, launch_kdiag(s, h->chanbuf2, h->chan_dev, N, out_dev, st, h->comps2, x_dev));
    return 0;
}
extern "C" int mogp_kdiag(mogp_handle_t h, int kind, int C, int Q, int D, const double* params_dev,
                          const int32_t* chan_off_host, double* out_dev, void* stream) {
    return mogp_kdiag_x(h, kind, C, Q, D, params_dev, nullptr, chan_off_host, out_dev, stream);
}

// fp64-on-int8 tcgen05 path (i8mm.cu) for the K^-1 = L^-T L^-1 product: padded sizes >= g_i8_min_np use it (0 = never).
// (measured at N = 2048, cfg2: K^-1 stage 0.126 -> 0.114 ms, step 0.847 -> 0.822 ms; below that the DMMA GEMM wins)
long long g_i8_min_np = std::getenv("MOGP_I8_MIN_NP") ? std::atoll(std::getenv("MOGP_I8_MIN_NP")) : 2048;
int g_i8_slices = std::getenv("MOGP_I8_SLICES") ? std::atoi(std::getenv("MOGP_I8_SLICES")) : 7;
extern "C" int mogp_set_i8(long long min_np, int slices) {
    if (slices != 7 && slices != 8) return -1;
    g_i8_min_np = min_np; g_i8_slices = slices; ++g_mogp_cfg_epoch;
    return 0;
}
// Three-level Cholesky for the largest matrices: super-panels of 1024 columns whose rank-1024 trailing update runs on the
// int8 tensor pipe (padded sizes >= this, multiples of 1024; 0 = never).  (A first version with rank-256 int8 updates was
// slower than DMMA: at K = 256 a tile has 8 K chunks and the kernel's fixed costs dominate -- cfg3 potrf 7.6 -> 11.5 ms,
// profiles/r02_i8_potrf_updates.txt.)
// Measured (profiles/r02_i8_potrf_three_level.txt): cfg3 (N = 8192) potrf 7.60 -> 6.71 ms, cfg4 (N = 4096) 2.42 -> 2.61 ms: on from 8192.
long long g_i8_potrf_min = std::getenv("MOGP_I8_POTRF_MIN") ? std::atoll(std::getenv("MOGP_I8_POTRF_MIN")) : 8192;
extern "C" int mogp_set_i8_potrf_min(long long np) { g_i8_potrf_min = np; ++g_mogp_cfg_epoch; return 0; }
extern int g_i8_ts, g_i8_wide;
extern "C" int mogp_set_i8_wide(int level) { g_i8_wide = level < 0 ? 0 : (level > 3 ? 3 : level); ++g_mogp_cfg_epoch; return 0; }
extern "C" int mogp_get_i8_wide(void) { return g_i8_wide; }
extern "C" int mogp_set_i8_ts(int on) { g_i8_ts = on ? 1 : 0; ++g_mogp_cfg_epoch; return 0; }
extern "C" int mogp_get_i8_ts(void) { return g_i8_ts; }
// smallest doubling-level block size of the triangular inverse that runs on the int8 pipe (0 = none)
extern "C" int mogp_set_i8_trtri_min(long long rows) { g_i8_trtri_min = rows; ++g_mogp_cfg_epoch; return 0; }
extern "C" long long mogp_get_i8_min_np(void) { return g_i8_min_np; }
extern "C" int mogp_get_i8_slices(void) { return g_i8_slices; }
extern "C" long long mogp_get_i8_trtri_min(void) { return g_i8_trtri_min; }
extern int g_rowpipe_kinv;      // linalg.cu
static bool use_i8(int64_t Np) { return g_i8_min_np > 0 && Np >= g_i8_min_np; }
// host-side preparation of the int8 path for a padded size (never inside capture); invalidates captured graphs when the
// tile lists had to be rebuilt (another size / leading dimension / slice count used this handle in between)
static int i8_ready(mogp_handle_s* h, int64_t Np, long long ld, cudaStream_t st) {
    if (!h->i8) h->i8 = i8_plan_create();
    bool changed = false;
    MOGP_CHECK(h, i8_prepare(h->i8, Np, ld, g_i8_slices, st, &changed));
    if (changed) h->realloc_epoch++;
    return 0;
}
static cudaError_t kinv_dispatch(mogp_handle_s* h, int64_t Np, long long ld, cudaStream_t st) {
    if (use_i8(Np) && h->i8) return i8_kinv(h->i8, h->Linv, h->W, Np, ld, g_i8_slices, st);
    return kinv_padded(h->Linv, h->W, Np, ld, nullptr, st);
}

// ------------------------------------------------------------------ Cholesky
extern "C" int mogp_potrf(mogp_handle_t h, double* A_dev, int64_t n, int64_t lda, int32_t* info_dev, void* stream) {
    if (!h) return -1;
    cudaStream_t st = (cudaStream_t)stream;
    MOGP_CHECK(h, cudaSetDevice(h->device));
    H_ARG(h, n >= 1 && n <= h->max_n, "n exceeds the handle's max_n");
    H_ARG(h, lda >= n, "lda < n");
    const int64_t Np = padded_size(n, h->np_max, 8192);      // (the factor alone takes the recursive scheme from 8192 rows)
    if (zero_linv_for(h, Np, st)) return -2;
    h->have_factor = false;
    const bool direct = n == Np && (lda % 2) == 0 && (reinterpret_cast<uintptr_t>(A_dev) % 16) == 0;
    // int8 trailing updates (large n) only for the handle's standard leading dimension: one set of tile lists per handle
    const bool i8 = use_i8(Np) && (!direct || lda == Np);
    if (i8 && i8_ready(h, Np, Np, st)) return -2;
    // (large sizes: the recursive scheme, which needs the inverses of its leading blocks: Linv and W are its scratch)
    auto factor = [&](double* Af, long long ldf) -> cudaError_t {
        // (measured, profiles/r02_recursive_cholesky.txt: factor alone 6.74 -> 6.21 ms at N = 8192, but 1.63 -> 1.77 ms at 4096, where
        //  the inverses of the leading blocks are pure overhead for a caller that only wants L)
        cudaError_t er = (i8 && ldf == Np && Np >= 8192) ? rchol_padded(Af, ldf, h->Linv, h->W, Np, h->logdet_part, h->info, st, &h->ps, h->i8,
                                                          g_i8_slices, 0)
                                           : cudaErrorNotSupported;
        if (er != cudaErrorNotSupported) return er;
        return potrf_padded(Af, ldf, h->Linv, Np, h->W, Np, Np, h->logdet_part, h->info, st, &h->ps, nullptr, i8 ? h->i8 : nullptr,
                            g_i8_slices);
    };
    if (direct) {
        MOGP_CHECK(h, factor(A_dev, lda));
    } else {
        MOGP_CHECK(h, launch_copy_tri(0, A_dev, lda, h->A, Np, n, Np, st));
        MOGP_CHECK(h, factor(h->A, Np));
        MOGP_CHECK(h, launch_copy_tri(1, A_dev, lda, h->A, Np, n, Np, st));
    }
    if (info_dev) MOGP_CHECK(h, cudaMemcpyAsync(info_dev, h->info, 4, cudaMemcpyDeviceToDevice, st));
    return 0;
}

extern "C" int mogp_trtri_kinv(mogp_handle_t h, double* A_dev, double* Linv_dev, double* Kinv_dev, int64_t n,
                               int32_t* info_dev, void* stream) {
    if (!h) return -1;
    cudaStream_t st = (cudaStream_t)stream;
    MOGP_CHECK(h, cudaSetDevice(h->device));
    H_ARG(h, n % MOGP_PAD == 0 && n <= h->np_max, "n must be a multiple of 128 within max_n");
    MOGP_CHECK(h, cudaMemsetAsync(Linv_dev, 0, (size_t)n * n * 8, st));
    const bool i8 = use_i8(n);           // same dispatch as the fused step (int8 tensor pipe for large n)
    if (i8 && i8_ready(h, n, n, st)) return -2;
    bool fused_inverse = false, fused_kinv = false;
    // same dispatch as the fused step: row-wise pipeline for small n; a scratch of its own only when K^-1 is accumulated behind
    // the chain as well (then Kinv_dev cannot double as the scratch)
    const bool rowp = rowpipe_applies(n) && g_rowpipe_kinv != 0;
    if (rowp && ensure(h, h->T, h->T_cap, (size_t)n * n)) return -2;
    MOGP_CHECK(h, potrf_padded(A_dev, n, Linv_dev, n, rowp ? h->T : Kinv_dev, n, n, h->logdet_part, h->info, st, &h->ps,
                               &fused_inverse, i8 ? h->i8 : nullptr, g_i8_slices, rowp ? Kinv_dev : nullptr, &fused_kinv));
    if (!fused_inverse) MOGP_CHECK(h, trtri_padded(A_dev, Linv_dev, Kinv_dev, n, n, st, i8 ? h->i8 : nullptr, g_i8_slices));
    if (fused_kinv) MOGP_CHECK(h, cudaStreamWaitEvent(st, h->ps.ev_kinv, 0));
    else if (i8) MOGP_CHECK(h, i8_kinv(h->i8, Linv_dev, Kinv_dev, n, n, g_i8_slices, st));
    else MOGP_CHECK(h, kinv_padded(Linv_dev, Kinv_dev, n, n, nullptr, st));
    if (info_dev) MOGP_CHECK(h, cudaMemcpyAsync(info_dev, h->info, 4, cudaMemcpyDeviceToDevice, st));
    h->have_factor = false;
    return 0;
}

// ------------------------------------------------------------------ the exact-GP step
// vec layout (np_max each): 0 y_pad | 1 z | 2 alpha | 3 kinv_diag | 4 colsq | 5 mu_tmp | 6.. staging (see below)
//
// The launch sequence of one evaluation (~100 kernels over three streams) is fixed for a given problem shape,
// so it is captured into a CUDA graph once and replayed (launch latency dominates the small
// configurations).  Graph kernels only touch handle-owned buffers; the caller's params / sigma / y / x / data_var
// are copied into them by one small staging kernel before the graph, and the output block is copied out after it.
__global__ void stage_inputs_kernel(const double* __restrict__ params, int P, const double* __restrict__ sigma, int C,
                                    const double* __restrict__ y, const double* __restrict__ dv,
                                    const double* __restrict__ x, long long N, int D, double* __restrict__ gp,
                                    double* __restrict__ gs, double* __restrict__ gy, double* __restrict__ gdv,
                                    double* __restrict__ gx) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i < P) gp[i] = params[i];
    if (i < C) gs[i] = sigma[i];
    if (i < N) { gy[i] = y[i]; if (dv) gdv[i] = dv[i]; }
    if (i < N * D && x != gx) gx[i] = x[i];
}
__global__ void copy_out_kernel(const double* __restrict__ src, double* __restrict__ dst, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = src[i];
}

// Pure enqueue of the step (kernel launches + event fork/join only: capturable).
static int enqueue_step(mogp_handle_s* h, const KernSpec& s, TileList* tl, int64_t N, int64_t Np, const double* params,
                        const double* sigma, const double* y, const double* dv, double jitter_rel, int want_grad,
                        double* out, cudaStream_t st) {
    const long long ld = Np;
    double* ypad = h->vec;
    double* z = h->vec + h->np_max;
    double* alpha = h->vec + 2 * h->np_max;
    double* kdiag = h->vec + 3 * h->np_max;
#define STAGE_MARK()                                                            \
    do {                                                                        \
        if (h->profile && h->n_ev < 8) cudaEventRecord(h->ev[h->n_ev++], st);   \
    } while (0)
    h->n_ev = 0;
    STAGE_MARK();
    MOGP_CHECK(h, launch_stamp(0, st));
    MOGP_CHECK(h, launch_prep(s, params, sigma, dv, h->chan_dev, N, jitter_rel, h->comps, h->chanbuf, st, h->xbuf,
                              s.window ? h->winsum : nullptr));
    // K~ (lower) -> L, diag blocks of Linv
    MOGP_CHECK(h, launch_kbuild(s, *tl, h->comps, h->chanbuf, h->xbuf, nullptr, h->chan_dev, dv, 1, h->A, ld, N, Np, st));
    STAGE_MARK();
    MOGP_CHECK(h, launch_stamp(1, st));
    // (with the pipelined inverse the GEMMs of Linv = L^-1 are issued behind the panel chain, inside potrf_padded)
    // Small sizes (row-wise pipeline, linalg.cu): Linv AND K^-1 (into W) are built behind the panel chain; the scratch is T then.
    bool fused_inverse = false, fused_kinv = false;
    // (`rowp`: K^-1 accumulated behind the chain as well, which needs the scratch T beside W; off by default -- the row-wise
    //  pipeline of Linv alone runs with W as its scratch)
    const bool rowp = want_grad && g_rowpipe_kinv != 0 && rowpipe_applies(Np) && h->T != nullptr && h->T_cap >= (size_t)Np * Np;
    MOGP_CHECK(h, launch_pad_copy(y, N, ypad, Np, st));
    ZChain zc{ypad, z, N, h->early_on ? h->early_host : nullptr, h->early_on ? h->early_ctr : nullptr, false};
    // (z along the pipeline: measured -1.3 % on the device-timed step for no visible end-to-end gain -- 32 more launches beside
    //  the chain; off by default, MOGP_ZCHAIN=1)
    static const int zchain_on = std::getenv("MOGP_ZCHAIN") ? std::atoi(std::getenv("MOGP_ZCHAIN")) : 0;
    // Large sizes: recursive factor + inverse, everything above the 2048-row leaves on the int8 tensor pipe
    cudaError_t er = use_i8(Np) ? rchol_padded(h->A, ld, h->Linv, h->W, Np, h->logdet_part, h->info, st, &h->ps, h->i8, g_i8_slices, 1)
                                : cudaErrorNotSupported;
    if (er == cudaSuccess) fused_inverse = true;
    else if (er != cudaErrorNotSupported) MOGP_CHECK(h, er);
    else
    MOGP_CHECK(h, potrf_padded(h->A, ld, h->Linv, ld, rowp ? h->T : h->W, ld, Np, h->logdet_part, h->info, st, &h->ps,
                               &fused_inverse, use_i8(Np) ? h->i8 : nullptr, g_i8_slices, rowp ? h->W : nullptr, &fused_kinv, zchain_on ? &zc : nullptr));
    STAGE_MARK();
    MOGP_CHECK(h, launch_stamp(2, st));
    // Linv, z = Linv y, alpha = Linv^T z, diag(K^-1)
    if (!fused_inverse) MOGP_CHECK(h, trtri_padded(h->A, h->Linv, h->W, Np, ld, st, use_i8(Np) ? h->i8 : nullptr, g_i8_slices));
    STAGE_MARK();
    // K^-1 = Linv^T Linv does not need alpha: it runs on a second stream concurrently with the solves
    // (z = Linv y, alpha = Linv^T z, diag K^-1); the gradient kernel subtracts alpha alpha^T while loading.
    // (Profiling keeps the stages sequential so that the stage timers stay meaningful.)
    const bool fork = want_grad && !fused_kinv && !h->profile && h->ps.s3 != nullptr;
    if (fork) {
        MOGP_CHECK(h, cudaEventRecord(h->ev_f1, st));
        MOGP_CHECK(h, cudaStreamWaitEvent(h->ps.s3, h->ev_f1, 0));
        MOGP_CHECK(h, kinv_dispatch(h, Np, ld, h->ps.s3));
        MOGP_CHECK(h, cudaEventRecord(h->ev_f2, h->ps.s3));
    }
    if (!zc.done) {      // (with the row-wise pipeline z and the early loss came out of potrf_padded)
        MOGP_CHECK(h, launch_trmv_lower(h->Linv, ld, ypad, z, Np, st));
        if (h->early_on) MOGP_CHECK(h, launch_lml_early(z, h->logdet_part, h->info, N, Np, h->early_host, h->early_ctr, st));
    }
    MOGP_CHECK(h, launch_colpass(h->Linv, ld, z, Np, Np, h->colpart, h->colpart_cap, alpha, kdiag, st));
    STAGE_MARK();
    MOGP_CHECK(h, launch_stamp(3, st));
    if (want_grad) {
        if (fused_kinv) MOGP_CHECK(h, cudaStreamWaitEvent(st, h->ps.ev_kinv, 0));
        else if (fork) MOGP_CHECK(h, cudaStreamWaitEvent(st, h->ev_f2, 0));
        else MOGP_CHECK(h, kinv_dispatch(h, Np, ld, st));
        STAGE_MARK();
        MOGP_CHECK(h, launch_stamp(4, st));
        MOGP_CHECK(h, launch_grad_reduce(s, *tl, h->comps, h->xbuf, h->W, ld, alpha, h->tile_part, st));
    }
    MOGP_CHECK(h, launch_finalize(s, tl, want_grad, params, sigma, h->comps, h->chanbuf, h->tile_part, z, alpha, kdiag,
                                  h->logdet_part, h->info, h->chan_dev, N, Np, jitter_rel, out, st, s.window ? h->winsum : nullptr));
    STAGE_MARK();
    MOGP_CHECK(h, launch_stamp(5, st));
#undef STAGE_MARK
    return 0;
}

extern "C" int mogp_set_panel_pdl(int v);      // linalg.cu
static int g_fail_capture = 0;                 // test hook: treat the next captures as failed (exercises the eager fall-back)
extern "C" void mogp_test_fail_capture(int on) { g_fail_capture = on; }
static int g_use_graphs = -1;
// Largest padded size whose step is replayed as a CUDA graph.  With programmatic dependent launches between the panel steps
// the replayed graph wins at every size we run (cfg4 N=4096: 3.34 ms against 3.53 ms eagerly; cfg3 N=8192: equal).
static long long g_graph_max_np = std::getenv("MOGP_GRAPH_MAX_NP") ? std::atoll(std::getenv("MOGP_GRAPH_MAX_NP")) : (1ll << 20);
extern "C" void mogp_set_graph_max_np(long long v) { g_graph_max_np = v; }
extern "C" void mogp_set_graphs(int on) { g_use_graphs = on ? 1 : 0; }

extern "C" int mogp_lml_grad(mogp_handle_t h, int kind, int C, int Q, int D, const double* params_dev,
                             const double* x_dev, const int32_t* chan_off_host, const double* y_dev,
                             const double* noise_sigma_dev, const double* data_var_dev, double jitter_rel, int want_grad,
                             double* out_dev, void* stream) {
    if (!h) return -1;
    cudaStream_t st = (cudaStream_t)stream;
    MOGP_CHECK(h, cudaSetDevice(h->device));
    H_ARG(h, params_dev && x_dev && y_dev && noise_sigma_dev && out_dev, "NULL argument");
    if (check_offsets(h, C, chan_off_host)) return -1;
    const int64_t N = chan_off_host[C];
    H_ARG(h, N >= 1 && N <= h->max_n, "N out of range for this handle");
    KernSpec s;
    H_ARG(h, spec_init(s, kind, C, Q, D) == 0, "bad kernel spec (kind, C, Q, D)");
    H_ARG(h, C <= 64, "at most 64 channels");
    const int64_t Np = padded_size(N, h->np_max);
    if (g_use_graphs < 0) {
        const char* e = getenv("MOGP_GRAPH");
        g_use_graphs = (e && atoi(e) == 0) ? 0 : 1;
    }
    // ---- host-side preparation (allocations, uploads): never captured
    std::vector<int32_t> off(chan_off_host, chan_off_host + C + 1);
    if (off != h->chan_uploaded) {
        if (upload_chan(h, C, chan_off_host, 0, st)) return -2;
        h->chan_uploaded = off;
    }
    if (ensure(h, h->comps, h->comps_cap, (size_t)C * C * s.R * s.st)) return -2;
    if (s.window && ensure(h, h->winsum, h->winsum_cap, (size_t)C * s.R * (2 + D))) return -2;
    if (zero_linv_for(h, Np, st)) return -2;
    TileList* tl = get_tiles(h, C, chan_off_host, nullptr, 0, st);
    H_ARG(h, tl != nullptr, "tile list allocation failed");
    if (ensure(h, h->xbuf, h->xbuf_cap, (size_t)N * D)) return -2;
    if (want_grad) {
        const size_t need = ((size_t)tl->n + (size_t)C * (C + 1) / 2) * s.R * s.st;
        if (ensure(h, h->tile_part, h->tile_part_cap, need)) return -2;
    }
    if (use_i8(Np) && i8_ready(h, Np, Np, st)) return -2;
    if (want_grad && g_rowpipe_kinv != 0 && rowpipe_applies(Np) && !(use_i8(Np) && rchol_applies(Np)) &&
        ensure(h, h->T, h->T_cap, (size_t)Np * Np)) return -2;
    const size_t nout = 2 + (size_t)s.P + C;
    const size_t stage_need = (size_t)s.P + C + 2 * (size_t)N + nout + 16;
    if (ensure(h, h->gbuf, h->gbuf_cap, stage_need)) return -2;
    double* gp = h->gbuf;
    double* gs = gp + s.P;
    double* gy = gs + C;
    double* gdv = gy + N;
    double* gout = gdv + N;

    // (Before the panel steps were chained with programmatic dependent launches, graphs were only used up to Np = 3072:
    // a replayed graph loses the stream priorities of the look-ahead, which cost 7% at N = 4096.)
    const bool graphs = g_use_graphs == 1 && !h->profile && Np <= g_graph_max_np;
    if (!graphs) {
        if (x_dev != h->xbuf)
            MOGP_CHECK(h, cudaMemcpyAsync(h->xbuf, x_dev, (size_t)N * D * 8, cudaMemcpyDeviceToDevice, st));
        int rc = enqueue_step(h, s, tl, N, Np, params_dev, noise_sigma_dev, y_dev, data_var_dev, jitter_rel, want_grad,
                              out_dev, st);
        if (rc) return rc;
    } else {
        // stage the caller's buffers, then replay (or first run / capture) the step on handle-owned buffers.
        // Everything runs on the handle's own stream (capture is not allowed on the legacy default stream that
        // PyTorch uses by default), forked from / joined to the caller's stream with events.
        cudaStream_t user = st;
        MOGP_CHECK(h, cudaEventRecord(h->ev_in, user));
        st = h->hs;
        MOGP_CHECK(h, cudaStreamWaitEvent(st, h->ev_in, 0));
        const long long nmax = std::max<long long>(std::max<long long>(s.P, C), N * D);
        stage_inputs_kernel<<<(unsigned)((nmax + 255) / 256), 256, 0, st>>>(params_dev, s.P, noise_sigma_dev, C, y_dev,
                                                                            data_var_dev, x_dev, N, D, gp, gs, gy, gdv,
                                                                            h->xbuf);
        MOGP_COUNT(1);
        StepGraph* sg = nullptr;
        for (StepGraph* c : h->graphs)
            if (c->kind == kind && c->C == C && c->Q == Q && c->D == D && c->N == N && c->want_grad == want_grad &&
                c->jitter == jitter_rel && c->has_dv == (data_var_dev != nullptr) && c->off == off) { sg = c; break; }
        if (!sg) {
            sg = new StepGraph();
            sg->kind = kind; sg->C = C; sg->Q = Q; sg->D = D; sg->N = N; sg->want_grad = want_grad;
            sg->jitter = jitter_rel; sg->has_dv = data_var_dev != nullptr; sg->off = off;
            if (h->graphs.size() >= 16) {
                StepGraph* old = h->graphs.front();
                if (old->exec) cudaGraphExecDestroy(old->exec);
                delete old;
                h->graphs.erase(h->graphs.begin());
            }
            h->graphs.push_back(sg);
        }
        const double* dvp = data_var_dev ? gdv : nullptr;
        if (sg->exec && (sg->epoch != h->realloc_epoch || sg->cfg_epoch != g_mogp_cfg_epoch)) {   // a workspace buffer moved (or a tuning knob changed) since the capture
            cudaGraphExecDestroy(sg->exec);
            sg->exec = nullptr;
        }
        if (sg->exec) {
            MOGP_CHECK(h, cudaGraphLaunch(sg->exec, st));
            MOGP_COUNT(sg->launches);
        } else if (sg->uses == 0) {                  // first use: plain run (also sets kernel attributes, warms caches)
            int rc = enqueue_step(h, s, tl, N, Np, gp, gs, gy, dvp, jitter_rel, want_grad, gout, st);
            if (rc) return rc;
        } else {                                     // second use: capture, instantiate, launch
            const long long l0 = g_mogp_launches;
            MOGP_CHECK(h, cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
            int rc = enqueue_step(h, s, tl, N, Np, gp, gs, gy, dvp, jitter_rel, want_grad, gout, st);
            cudaGraph_t graph = nullptr;
            cudaError_t ce = cudaStreamEndCapture(st, &graph);
            bool captured = rc == 0 && ce == cudaSuccess && graph != nullptr && !g_fail_capture;
            if (captured) {
                sg->launches = g_mogp_launches - l0;
                sg->epoch = h->realloc_epoch;
                sg->cfg_epoch = g_mogp_cfg_epoch;
                ce = cudaGraphInstantiate(&sg->exec, graph, 0);
                if (ce != cudaSuccess) { sg->exec = nullptr; captured = false; }
            }
            if (graph) cudaGraphDestroy(graph);
            if (captured) {
                MOGP_CHECK(h, cudaGraphLaunch(sg->exec, st));
            } else {
                // Capture or instantiation is not possible in this context (e.g. a driver without programmatic
                // dependent launch edges in graphs): no graphs and plain launch dependencies from now on, and
                // this evaluation runs eagerly -- the caller never sees the failed attempt.
                cudaGetLastError();
                if (rc) return rc;                   // a genuine argument / launch error of the step itself
                g_use_graphs = 0;
                mogp_set_panel_pdl(0);
                rc = enqueue_step(h, s, tl, N, Np, gp, gs, gy, dvp, jitter_rel, want_grad, gout, st);
                if (rc) return rc;
            }
        }
        sg->uses++;
        copy_out_kernel<<<(unsigned)((nout + 255) / 256), 256, 0, st>>>(gout, out_dev, (int)(want_grad ? nout : 2));
        MOGP_COUNT(1);
        MOGP_CHECK(h, cudaEventRecord(h->ev_out, st));
        MOGP_CHECK(h, cudaStreamWaitEvent(user, h->ev_out, 0));
    }
    if (h->early_on) ++h->early_expected;      // exactly one lml_early_kernel was enqueued above (eager run or graph launch)
    h->have_factor = true;
    h->spec = s;
    h->chan_off.assign(chan_off_host, chan_off_host + C + 1);
    h->N = N;
    h->Np = Np;
    return 0;
}

// Early loss: the step writes [lml, info, seq] into mapped pinned host memory as soon as the solves are done (long before the
// gradient is), seq counting the evaluations of this handle since the feature was switched on.  A caller that only needs the
// loss value to return (gpr.Model.loss: the value and a possible CholeskyException, mogptk/gpr/model.py:279-292) polls
// host_buf[2] for the value of mogp_early_expected() instead of synchronising the stream; the gradient buffers are then
// complete in stream order, like the result of any asynchronous CUDA call.
extern "C" int mogp_early_loss(mogp_handle_t h, int on, double** host_buf) {
    if (!h) return -1;
    MOGP_CHECK(h, cudaSetDevice(h->device));
    if (on && !h->early_host) {
        MOGP_CHECK(h, cudaHostAlloc(&h->early_host, 8 * sizeof(double), cudaHostAllocMapped | cudaHostAllocPortable));
        MOGP_CHECK(h, cudaMalloc(&h->early_ctr, sizeof(unsigned long long)));
        for (int i = 0; i < 8; ++i) h->early_host[i] = 0.0;
    }
    if (on) {
        MOGP_CHECK(h, cudaDeviceSynchronize());
        MOGP_CHECK(h, cudaMemset(h->early_ctr, 0, sizeof(unsigned long long)));
        h->early_expected = 0;
        h->early_host[2] = 0.0;
    }
    if ((on != 0) != h->early_on) { h->early_on = on != 0; h->realloc_epoch++; }      // captured steps change: re-capture
    if (host_buf) *host_buf = h->early_host;
    return 0;
}
extern "C" unsigned long long mogp_early_expected(mogp_handle_t h) { return h ? h->early_expected : 0; }

extern "C" int mogp_lml_grad_host(mogp_handle_t h, int kind, int C, int Q, int D, const double* params_host,
                                  const double* x_host, const int32_t* chan_off_host, const double* y_host,
                                  const double* noise_sigma_host, const double* data_var_host, double jitter_rel,
                                  int want_grad, double* out_host) {
    if (!h) return -1;
    MOGP_CHECK(h, cudaSetDevice(h->device));
    KernSpec s;
    H_ARG(h, spec_init(s, kind, C, Q, D) == 0, "bad kernel spec (kind, C, Q, D)");
    if (check_offsets(h, C, chan_off_host)) return -1;
    const int64_t N = chan_off_host[C];
    H_ARG(h, N >= 1 && N <= h->max_n, "N out of range for this handle");
    const size_t nP = s.P, nout = 2 + s.P + C;
    const size_t need = nP + C + (size_t)N * (1 + D) + (data_var_host ? N : 0) + 8;
    if (ensure(h, h->pbuf, h->pbuf_cap, need)) return -2;
    if (ensure(h, h->out_dev, h->out_cap, nout)) return -2;
    if (ensure(h, h->xbuf, h->xbuf_cap, (size_t)N * D)) return -2;
    cudaStream_t st = 0;
    double* p_dev = h->pbuf;
    double* sig_dev = p_dev + nP;
    double* y_dev = sig_dev + C;
    double* dv_dev = data_var_host ? y_dev + N : nullptr;
    MOGP_CHECK(h, cudaMemcpyAsync(p_dev, params_host, nP * 8, cudaMemcpyHostToDevice, st));
    MOGP_CHECK(h, cudaMemcpyAsync(sig_dev, noise_sigma_host, C * 8, cudaMemcpyHostToDevice, st));
    MOGP_CHECK(h, cudaMemcpyAsync(y_dev, y_host, N * 8, cudaMemcpyHostToDevice, st));
    MOGP_CHECK(h, cudaMemcpyAsync(h->xbuf, x_host, (size_t)N * D * 8, cudaMemcpyHostToDevice, st));
    if (dv_dev) MOGP_CHECK(h, cudaMemcpyAsync(dv_dev, data_var_host, N * 8, cudaMemcpyHostToDevice, st));
    int rc = mogp_lml_grad(h, kind, C, Q, D, p_dev, h->xbuf, chan_off_host, y_dev, sig_dev, dv_dev, jitter_rel, want_grad,
                           h->out_dev, (void*)st);
    if (rc) return rc;
    MOGP_CHECK(h, cudaMemcpyAsync(out_host, h->out_dev, (want_grad ? nout : 2) * 8, cudaMemcpyDeviceToHost, st));
    MOGP_CHECK(h, cudaStreamSynchronize(st));
    return 0;
}

// ------------------------------------------------------------------ prediction
extern "C" int mogp_predict(mogp_handle_t h, const double* xs_dev, const int32_t* chan_off_s_host, int full,
                            double* mu_dev, double* var_dev, void* stream) {
    if (!h) return -1;
    cudaStream_t st = (cudaStream_t)stream;
    MOGP_CHECK(h, cudaSetDevice(h->device));
    H_ARG(h, h->have_factor, "mogp_predict needs a preceding mogp_lml_grad on this handle");
    const KernSpec s = h->spec;
    if (check_offsets(h, s.C, chan_off_s_host)) return -1;
    const int64_t M = chan_off_s_host[s.C];
    H_ARG(h, M >= 1, "no test points");
    const int64_t Mp = round_up(M, 128), Np = h->Np, N = h->N;
    const size_t need = (size_t)Np * Mp;
    if (need > h->pred_cap) {
        if (h->pred_K) cudaFree(h->pred_K);
        if (h->pred_V) cudaFree(h->pred_V);
        h->pred_K = h->pred_V = nullptr; h->pred_cap = 0;
        MOGP_CHECK(h, cudaMalloc(&h->pred_K, need * 8));
        MOGP_CHECK(h, cudaMalloc(&h->pred_V, need * 8));
        h->pred_cap = need;
    }
    if (upload_chan(h, s.C, chan_off_s_host, 1, st)) return -2;
    const int32_t* chan_s_dev = h->chan_dev + 260;
    TileList* tl = get_tiles(h, s.C, h->chan_off.data(), chan_off_s_host, 2, st);
    H_ARG(h, tl != nullptr, "tile list allocation failed");
    double* alpha = h->vec + 2 * h->np_max;
    double* colsq = h->vec + 4 * h->np_max;      // needs Mp <= np_max ... guarded below
    double* mu_tmp = h->vec + 5 * h->np_max;
    H_ARG(h, Mp <= h->np_max, "too many test points for one call (chunk on the host side)");
    if ((size_t)(64 * 2 * Mp) > h->colpart_cap) { h->err = "column-pass scratch too small"; return -1; }

    MOGP_CHECK(h, cudaMemsetAsync(h->pred_K, 0, need * 8, st));
    MOGP_CHECK(h, launch_kbuild(s, *tl, h->comps, h->chanbuf, h->xbuf, xs_dev, h->chan_dev, nullptr, 0, h->pred_K, Mp, N,
                                N, st));
    // mu = Kfs^T alpha
    MOGP_CHECK(h, launch_colpass(h->pred_K, Mp, alpha, Np, Mp, h->colpart, h->colpart_cap, mu_tmp, nullptr, st));
    MOGP_CHECK(h, cudaMemcpyAsync(mu_dev, mu_tmp, M * 8, cudaMemcpyDeviceToDevice, st));
    // V = Linv * Kfs
    GemmArgs g{};
    g.A = h->Linv; g.lda = Np;
    g.B = h->pred_K; g.ldb = Mp;
    g.C = h->pred_V; g.ldc = Mp;
    g.M = (int)Np; g.N = (int)Mp; g.K = (int)Np;
    g.khi_mode = 1; g.alpha = 1.0; g.beta = 0.0;
    MOGP_CHECK(h, launch_gemm(0, 0, g, 1, st));
    if (!full) {
        MOGP_CHECK(h, launch_colpass(h->pred_V, Mp, alpha, Np, Mp, h->colpart, h->colpart_cap, nullptr, colsq, st));
        const double* kss = nullptr;
        if (s.window) {          // the prior variance at the test points depends on the input
            MOGP_CHECK(h, launch_kdiag(s, h->chanbuf, chan_s_dev, M, mu_tmp, st, h->comps, xs_dev));
            kss = mu_tmp;        // (mu has been copied out above)
        }
        MOGP_CHECK(h, launch_pred_var(h->chanbuf, s.C, chan_s_dev, colsq, M, var_dev, st, kss));
    } else {
        const size_t sneed = (size_t)Mp * Mp;
        if (sneed > h->pred_s_cap) {
            if (h->pred_S) cudaFree(h->pred_S);
            h->pred_S = nullptr; h->pred_s_cap = 0;
            MOGP_CHECK(h, cudaMalloc(&h->pred_S, sneed * 8));
            h->pred_s_cap = sneed;
        }
        MOGP_CHECK(h, cudaMemsetAsync(h->pred_S, 0, sneed * 8, st));
        TileList* tg = get_tiles(h, s.C, chan_off_s_host, nullptr, 1, st);
        H_ARG(h, tg != nullptr, "tile list allocation failed");
        MOGP_CHECK(h, launch_kbuild(s, *tg, h->comps, h->chanbuf, xs_dev, nullptr, chan_s_dev, nullptr, 0, h->pred_S, Mp, M,
                                    M, st));
        GemmArgs c{};
        c.A = h->pred_V; c.lda = Mp;
        c.B = h->pred_V; c.ldb = Mp;
        c.C = h->pred_S; c.ldc = Mp;
        c.M = (int)Mp; c.N = (int)Mp; c.K = (int)Np;
        c.alpha = -1.0; c.beta = 1.0;
        MOGP_CHECK(h, launch_gemm(1, 0, c, 1, st));
        MOGP_CHECK(h, cudaMemcpy2DAsync(var_dev, M * 8, h->pred_S, Mp * 8, M * 8, M, cudaMemcpyDeviceToDevice, st));
    }
    return 0;
}

// alpha = K~^-1 y of the last evaluation (rows in the channel-sorted order the caller passed): d LML / d y = -alpha,
// which is what the host layer needs to back-propagate into a trainable mean function (gpr/model.py:445-452).
extern "C" int mogp_alpha(mogp_handle_t h, double* alpha_dev, void* stream) {
    if (!h) return -1;
    MOGP_CHECK(h, cudaSetDevice(h->device));
    H_ARG(h, h->have_factor, "mogp_alpha needs a preceding mogp_lml_grad on this handle");
    H_ARG(h, alpha_dev != nullptr, "NULL argument");
    MOGP_CHECK(h, cudaMemcpyAsync(alpha_dev, h->vec + 2 * h->np_max, (size_t)h->N * 8, cudaMemcpyDeviceToDevice,
                                  (cudaStream_t)stream));
    return 0;
}

// ------------------------------------------------------------------ building blocks
extern "C" int mogp_dgemm(mogp_handle_t h, int transa, int transb, int M, int N, int K, double alpha,
                          const double* A_dev, int64_t lda, const double* B_dev, int64_t ldb, double beta, double* C_dev,
                          int64_t ldc, void* stream) {
    if (!h) return -1;
    MOGP_CHECK(h, cudaSetDevice(h->device));
    H_ARG(h, M % 64 == 0 && N % 64 == 0 && K % 16 == 0, "M, N multiples of 64 and K multiple of 16 required");
    H_ARG(h, lda % 2 == 0 && ldb % 2 == 0 && ldc % 2 == 0, "leading dimensions must be even");
    GemmArgs g{};
    g.A = A_dev; g.lda = lda; g.B = B_dev; g.ldb = ldb; g.C = C_dev; g.ldc = ldc;
    g.M = M; g.N = N; g.K = K; g.alpha = alpha; g.beta = beta;
    // header convention: transa=0 -> A is (M x K) row-major; transb=0 -> B is (K x N) row-major
    MOGP_CHECK(h, launch_gemm(transa ? 1 : 0, transb ? 1 : 0, g, 1, (cudaStream_t)stream));
    return 0;
}

extern "C" long long mogp_launch_count(void) { return g_mogp_launches; }

// Stage timing of mogp_lml_grad (diagnostics for bench.py): with want_grad the stages are
// [kbuild, potrf, trtri, solves, kinv, grad+finalize]; returns the number of stages written.
extern "C" int mogp_set_profile(mogp_handle_t h, int on) {
    if (!h) return -1;
    MOGP_CHECK(h, cudaSetDevice(h->device));
    if (on && !h->ev[0])
        for (int i = 0; i < 8; ++i) MOGP_CHECK(h, cudaEventCreate(&h->ev[i]));
    h->profile = on != 0;
    h->n_ev = 0;
    return 0;
}
extern "C" int mogp_stage_times(mogp_handle_t h, float* ms_out) {
    if (!h || !h->profile || h->n_ev < 2) return 0;
    cudaEventSynchronize(h->ev[h->n_ev - 1]);
    for (int i = 0; i + 1 < h->n_ev; ++i) cudaEventElapsedTime(&ms_out[i], h->ev[i], h->ev[i + 1]);
    return h->n_ev - 1;
}

extern "C" int mogp_peak_fp64(mogp_handle_t h, double* dmma_tflops_host, double* dfma_tflops_host) {
    if (!h) return -1;
    MOGP_CHECK(h, cudaSetDevice(h->device));
    MOGP_CHECK(h, run_peak_fp64(dmma_tflops_host, dfma_tflops_host));
    return 0;
}
