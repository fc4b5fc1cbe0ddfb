// Host self-check hooks for csrc/covmath_next.cuh (round-2 groundwork; linked only into libmogp_b200_exp.so).
#include "covmath_next.cuh"

extern "C" int mogp_exp_num_params(int kind, int C, int Q, int Rq, int D) { return next_num_params(kind, C, Q, Rq, D); }

// comps_out: C*C*R records of comp_stride(D) doubles; returns R
extern "C" int mogp_exp_host_pair_comps(int kind, int C, int Q, int Rq, int D, const double* params, double* comps_out) {
    if ((kind != MOGP_KIND_CSM && kind != MOGP_KIND_SMLMC && kind != MOGP_KIND_UMOSM) || D < 1 || D > MOGP_MAX_D || C > 64) return -1;
    const int R = next_num_comps(kind, Q, Rq, D), st = comp_stride(D);
    for (int i = 0; i < C; ++i)
        for (int j = 0; j < C; ++j)
            for (int r = 0; r < R; ++r) pair_comp_next(kind, C, Q, Rq, D, params, i, j, r, comps_out + (size_t)((i * C + j) * R + r) * st);
    return R;
}

// gsum: per lower pair and component the weighted sums (see covmath.cuh); grad_out: packed gradient; returns P
extern "C" int mogp_exp_host_chain(int kind, int C, int Q, int Rq, int D, const double* params, const double* gsum,
                                   const double* adj, double* grad_out) {
    if ((kind != MOGP_KIND_CSM && kind != MOGP_KIND_SMLMC && kind != MOGP_KIND_UMOSM) || D < 1 || D > MOGP_MAX_D || C > 64) return -1;
    const int R = next_num_comps(kind, Q, Rq, D), st = comp_stride(D);
    std::vector<double> comps((size_t)C * C * R * st);
    mogp_exp_host_pair_comps(kind, C, Q, Rq, D, params, comps.data());
    const int owners = n_chain_owners_next(kind, C, Q, Rq);
    for (int o = 0; o < owners; ++o) chain_owner_next(kind, C, Q, Rq, D, params, comps.data(), gsum, adj, o, grad_out);
    return next_num_params(kind, C, Q, Rq, D);
}
