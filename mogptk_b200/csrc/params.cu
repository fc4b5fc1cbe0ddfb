// Constrained-parameter transforms of the reference on the device (keep the
// training iteration device-resident).  Restates mogptk/gpr/parameter.py:30-96:
//   Softplus: y = lower + softplus_beta(x) with torch's threshold rule (beta*x > 20 -> x), dy/dx = sigmoid(beta x)
//   Sigmoid : y = lower + (upper - lower) * sigmoid(x),                                     dy/dx = (upper-lower) s (1-s)
// forward : raw leaves -> packed constrained vector (+ the derivative of every element)
// backward: d loss / d constrained (from mogp_lml_grad) -> raw-space gradients written into the p.grad buffers.
#include "common.cuh"
#include <cstring>

struct DevEntry {
    const double* raw; double* grad; const double* lower; const double* upper;
    long long n, off; int type, lower_n, upper_n, pad; double beta;
};

__global__ void params_forward_kernel(const DevEntry* __restrict__ ent, double* __restrict__ packed,
                                      double* __restrict__ dcons) {
    const DevEntry e = ent[blockIdx.x];
    for (long long i = threadIdx.x; i < e.n; i += blockDim.x) {
        const double x = e.raw[i];
        double y = x, d = 1.0;
        if (e.type == 1) {                                   // softplus with slope beta
            const double lo = e.lower[e.lower_n == 1 ? 0 : i];
            const double bx = e.beta * x;
            if (bx > 20.0) { y = lo + x; d = 1.0; }
            else { y = lo + log1p(exp(bx)) / e.beta; d = 1.0 / (1.0 + exp(-bx)); }
        } else if (e.type == 2) {                            // sigmoid between lower and upper
            const double lo = e.lower[e.lower_n == 1 ? 0 : i], up = e.upper[e.upper_n == 1 ? 0 : i];
            const double s = 1.0 / (1.0 + exp(-x));
            y = lo + (up - lo) * s;
            d = (up - lo) * s * (1.0 - s);
        }
        packed[e.off + i] = y;
        dcons[e.off + i] = d;
    }
}

__global__ void params_backward_kernel(const DevEntry* __restrict__ ent, const double* __restrict__ gcons,
                                       const double* __restrict__ dcons, const double* __restrict__ lml,
                                       double* __restrict__ loss_out) {
    const DevEntry e = ent[blockIdx.x];
    if (e.grad)
        for (long long i = threadIdx.x; i < e.n; i += blockDim.x) e.grad[i] = gcons[e.off + i] * dcons[e.off + i];
    if (blockIdx.x == 0 && threadIdx.x == 0 && loss_out) loss_out[0] = -lml[0];
}

static int upload_entries(mogp_handle_s* h, const mogp_param_entry* ent, int n, cudaStream_t st) {
    if (n < 1 || n > 4096) { h->err = "bad number of parameter entries"; return -1; }
    const size_t bytes = (size_t)n * sizeof(DevEntry);
    if (bytes > h->pent_cap) {
        if (h->pent_dev) cudaFree(h->pent_dev);
        if (h->pent_host) cudaFreeHost(h->pent_host);
        h->pent_dev = nullptr; h->pent_host = nullptr; h->pent_cap = 0;
        MOGP_CHECK(h, cudaMalloc(&h->pent_dev, bytes * 2));
        MOGP_CHECK(h, cudaMallocHost(&h->pent_host, bytes * 2));
        h->pent_cap = bytes * 2;
    }
    static_assert(sizeof(DevEntry) == sizeof(mogp_param_entry), "entry layout");
    if (n == h->pent_n && memcmp(h->pent_host, ent, bytes) == 0) return 0;       // unchanged since the last call
    MOGP_CHECK(h, cudaStreamSynchronize(st));                                   // the staging copy may still be in flight
    memcpy(h->pent_host, ent, bytes);
    MOGP_CHECK(h, cudaMemcpyAsync(h->pent_dev, h->pent_host, bytes, cudaMemcpyHostToDevice, st));
    h->pent_n = n;
    return 0;
}

extern "C" int mogp_params_forward(mogp_handle_t h, const mogp_param_entry* entries_host, int n_entries,
                                   double* packed_dev, double* dcons_dev, void* stream) {
    if (!h) return -1;
    cudaStream_t st = (cudaStream_t)stream;
    MOGP_CHECK(h, cudaSetDevice(h->device));
    int rc = upload_entries(h, entries_host, n_entries, st);
    if (rc) return rc;
    params_forward_kernel<<<n_entries, 128, 0, st>>>((const DevEntry*)h->pent_dev, packed_dev, dcons_dev);
    MOGP_COUNT(1);
    MOGP_CHECK(h, cudaGetLastError());
    return 0;
}

extern "C" int mogp_params_backward(mogp_handle_t h, const mogp_param_entry* entries_host, int n_entries,
                                    const double* gcons_dev, const double* dcons_dev, const double* lml_dev,
                                    double* loss_out_dev, void* stream) {
    if (!h) return -1;
    cudaStream_t st = (cudaStream_t)stream;
    MOGP_CHECK(h, cudaSetDevice(h->device));
    int rc = upload_entries(h, entries_host, n_entries, st);
    if (rc) return rc;
    params_backward_kernel<<<n_entries, 128, 0, st>>>((const DevEntry*)h->pent_dev, gcons_dev, dcons_dev, lml_dev,
                                                     loss_out_dev);
    MOGP_COUNT(1);
    MOGP_CHECK(h, cudaGetLastError());
    return 0;
}

// One training-iteration evaluation in a single call: raw leaves -> constrained values, fused exact-GP step, chain rule
// into the p.grad buffers (what gpr.Model.loss does, mogptk/gpr/model.py:279-292).  work_dev: 3 * (2 + P + C) doubles laid
// out as packed | d constrained / d raw | out(lml, info, gradient block); loss_out_dev receives -lml.
extern "C" int mogp_loss_grad(mogp_handle_t h, int kind, int C, int Q, int D, const mogp_param_entry* entries_host,
                              int n_entries, const double* x_dev, const int32_t* chan_off_host, const double* y_dev,
                              const double* data_var_dev, double jitter_rel, double* work_dev, double* loss_out_dev,
                              void* stream) {
    if (!h) return -1;
    cudaStream_t st = (cudaStream_t)stream;
    MOGP_CHECK(h, cudaSetDevice(h->device));
    const int P = mogp_num_params(kind, C, Q, D);
    if (P < 0 || !entries_host || !work_dev) { h->err = "mogp_loss_grad: bad argument"; return -1; }
    const size_t n = 2 + (size_t)P + C;
    double* packed = work_dev;
    double* dcons = work_dev + n;
    double* out = work_dev + 2 * n;
    int rc = upload_entries(h, entries_host, n_entries, st);
    if (rc) return rc;
    params_forward_kernel<<<n_entries, 128, 0, st>>>((const DevEntry*)h->pent_dev, packed, dcons);
    MOGP_COUNT(1);
    rc = mogp_lml_grad(h, kind, C, Q, D, packed, x_dev, chan_off_host, y_dev, packed + P, data_var_dev, jitter_rel, 1, out, stream);
    if (rc) return rc;
    params_backward_kernel<<<n_entries, 128, 0, st>>>((const DevEntry*)h->pent_dev, out + 2, dcons, out, loss_out_dev);
    MOGP_COUNT(1);
    MOGP_CHECK(h, cudaGetLastError());
    return 0;
}

// ------------------------------------------------------------------ device-resident Adam training
// Replaces the body of the reference's training loop (mogptk/model.py:563-565: `progress(i, self.loss())`,
// `optimizer.step()` with torch.optim.Adam) for `iters` iterations without a host synchronisation: per iteration
//   raw leaves -> constrained (params_forward) -> fused exact-GP step (replayed CUDA graph) -> chain rule into the
//   p.grad buffers and the Adam update of the raw leaves in one kernel.
// The update restates torch.optim.Adam's single-tensor formula (amsgrad = False, weight_decay = 0, maximize = False):
//   m <- m + (1 - b1) (g - m);  v <- b2 v + (1 - b2) g^2;  p <- p - (lr / (1 - b1^t)) * m / (sqrt(v) / sqrt(1 - b2^t) + eps)
// A Cholesky failure (info != 0) freezes the parameters from that iteration on (sticky flag), so that the host can
// re-evaluate at exactly the failing parameters and raise the reference's CholeskyException there.
__global__ void params_backward_adam_kernel(const DevEntry* __restrict__ ent, const double* __restrict__ out /* lml, info, grads */,
                                            const double* __restrict__ dcons, double* __restrict__ exp_avg,
                                            double* __restrict__ exp_avg_sq, double one_minus_b1, double b2,
                                            double step_size, double bc2_sqrt, double eps, double* __restrict__ loss_slot,
                                            int32_t* __restrict__ fail /* [0] = info, [1] = iteration (1-based) */, int iter1) {
    const DevEntry e = ent[blockIdx.x];
    const double info = out[1];
    const bool failed = (fail[0] != 0) || (info != 0.0) || !(info == info);
    const double* gcons = out + 2;
    for (long long i = threadIdx.x; i < e.n; i += blockDim.x) {
        const double g = gcons[e.off + i] * dcons[e.off + i];
        if (e.grad) e.grad[i] = g;
        if (failed) continue;
        double m = exp_avg[e.off + i], v = exp_avg_sq[e.off + i];
        // torch's lerp_(grad, w = 1 - beta1): m + w (g - m) for w < 0.5, g - (g - m) (1 - w) otherwise
        m = one_minus_b1 < 0.5 ? m + one_minus_b1 * (g - m) : g - (g - m) * (1.0 - one_minus_b1);
        v = v * b2 + (1.0 - b2) * g * g;                      // mul_(beta2).addcmul_(grad, grad, value = 1 - beta2)
        exp_avg[e.off + i] = m;
        exp_avg_sq[e.off + i] = v;
        const double denom = sqrt(v) / bc2_sqrt + eps;
        const_cast<double*>(e.raw)[i] = e.raw[i] - step_size * (m / denom);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        loss_slot[0] = -out[0];
        // every block reads fail[] before this write can matter: the flag is only consulted by LATER launches
        if (failed && fail[0] == 0) { fail[1] = iter1; }
    }
}
// second tiny kernel so that the sticky flag is written after every block of the update kernel has read it
__global__ void adam_flag_kernel(const double* __restrict__ out, int32_t* __restrict__ fail) {
    const double info = out[1];
    if (fail[0] == 0 && ((info != 0.0) || !(info == info))) fail[0] = (info == info && info > 0.0) ? (int32_t)info : -1;
}

extern "C" int mogp_train_adam(mogp_handle_t h, int kind, int C, int Q, int D, const mogp_param_entry* entries_host,
                               int n_entries, const double* x_dev, const int32_t* chan_off_host, const double* y_dev,
                               const double* data_var_dev, double jitter_rel, double* work_dev /* 3 * (2 + P + C) */,
                               double* exp_avg_dev, double* exp_avg_sq_dev, long long step0, int iters, double lr,
                               double beta1, double beta2, double eps, double* losses_dev, int32_t* fail_dev,
                               void* stream) {
    if (!h) return -1;
    cudaStream_t st = (cudaStream_t)stream;
    MOGP_CHECK(h, cudaSetDevice(h->device));
    if (!entries_host || !work_dev || !exp_avg_dev || !exp_avg_sq_dev || !losses_dev || !fail_dev || iters < 0 ||
        !(beta1 >= 0.0 && beta1 < 0.5 + 0.5) || !(beta2 >= 0.0 && beta2 < 1.0) || step0 < 0) {
        h->err = "mogp_train_adam: bad argument";
        return -1;
    }
    const int P = mogp_num_params(kind, C, Q, D);
    if (P < 0) { h->err = "bad kernel spec (kind, C, Q, D)"; return -1; }
    const size_t n = 2 + (size_t)P + C;
    double* packed = work_dev;
    double* dcons = work_dev + n;
    double* out = work_dev + 2 * n;
    int rc = upload_entries(h, entries_host, n_entries, st);
    if (rc) return rc;
    for (int i = 0; i < iters; ++i) {
        params_forward_kernel<<<n_entries, 128, 0, st>>>((const DevEntry*)h->pent_dev, packed, dcons);
        MOGP_COUNT(1);
        rc = mogp_lml_grad(h, kind, C, Q, D, packed, x_dev, chan_off_host, y_dev, packed + P, data_var_dev, jitter_rel, 1, out,
                           stream);
        if (rc) return rc;
        const double t = (double)(step0 + i + 1);
        const double bc1 = 1.0 - pow(beta1, t), bc2 = 1.0 - pow(beta2, t);
        params_backward_adam_kernel<<<n_entries, 128, 0, st>>>((const DevEntry*)h->pent_dev, out, dcons, exp_avg_dev,
                                                              exp_avg_sq_dev, 1.0 - beta1, beta2, lr / bc1, sqrt(bc2), eps,
                                                              losses_dev + i, fail_dev, i + 1);
        adam_flag_kernel<<<1, 1, 0, st>>>(out, fail_dev);
        MOGP_COUNT(2);
    }
    MOGP_CHECK(h, cudaGetLastError());
    return 0;
}
