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
