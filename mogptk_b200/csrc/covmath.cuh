// Host/device closed forms shared by the CUDA kernels and the host-side self-checks:
//   * derived per channel-pair constants of the MOSM / SM / CONV kernels,
//   * the chain rule from per-pair weighted sums back to the constrained parameters.
// Formulas restate mogptk/gpr/multioutput.py:178-210 (MOSM), gpr/singleoutput.py:594-605 (SM),
// gpr/multioutput.py:531-553 (CONV); the derivatives are ours (the reference uses autograd).
#pragma once
#include "common.cuh"
#include <math.h>

#define MOGP_PI 3.14159265358979323846

// Packed-parameter offsets -------------------------------------------------------------
struct MosmOff { int w, mu, var, th, ph; };
__host__ __device__ inline MosmOff mosm_off(int C, int Q, int D) {
    MosmOff o; o.w = 0; o.mu = C * Q; o.var = o.mu + C * Q * D; o.th = o.var + C * Q * D; o.ph = o.th + C * Q * D;
    return o;
}
struct SmOff { int mag, mu, var; };
__host__ __device__ inline SmOff sm_off(int C, int Q, int D) {
    SmOff o; o.mag = 0; o.mu = C * Q; o.var = o.mu + C * Q * D; return o;
}
struct ConvOff { int w, var, base; };
__host__ __device__ inline ConvOff conv_off(int C, int Q, int D) {
    ConvOff o; o.w = 0; o.var = Q * C; o.base = o.var + Q * C * D; return o;
}

// gsum record of lower pair (i, j), component r (see the chain rule below)
__host__ __device__ inline const double* gs_rec(const double* gsum, int i, int j, int R, int st, int r) {
    const int pl = i * (i + 1) / 2 + j;
    return gsum + (size_t)(pl * R + r) * st;
}
#include "covmath_next.cuh"     // CSM / SM-LMC / uMOSM / MOHSM tables (families >= MOGP_KIND_CSM)

// comp record: [alpha, phi, v[D], m[D], theta[D]] (+ [l, c[D]] for MOHSM).  kind = family | Rq << 8.
__host__ __device__ inline void pair_comp(int kind, int C, int Q, int D, const double* __restrict__ p, int i, int j,
                                          int r, double* __restrict__ out) {
    if (kind_family(kind) >= MOGP_KIND_CSM) {
        pair_comp_next(kind_family(kind), C, Q, kind_rq(kind), D, p, i, j, r, out);
        return;
    }
    double alpha = 0.0, phi = 0.0;
    double* v = out + 2;
    double* m = out + 2 + D;
    double* th = out + 2 + 2 * D;
    for (int d = 0; d < D; ++d) v[d] = m[d] = th[d] = 0.0;
    if (kind == MOGP_KIND_MOSM) {
        const MosmOff o = mosm_off(C, Q, D);
        const int q = r;
        const double twopi_pow = pow(2.0 * MOGP_PI, 0.5 * (double)D);
        const double* mui = p + o.mu + (i * Q + q) * D; const double* muj = p + o.mu + (j * Q + q) * D;
        const double* si = p + o.var + (i * Q + q) * D; const double* sj = p + o.var + (j * Q + q) * D;
        if (i == j) {                                   // multioutput.py:182-187
            double prod = 1.0;
            for (int d = 0; d < D; ++d) { v[d] = si[d]; m[d] = mui[d]; prod *= si[d]; }
            const double w = p[o.w + i * Q + q];
            alpha = w * w * twopi_pow * sqrt(prod);
        } else {                                        // multioutput.py:189-199
            double esum = 0.0, prod = 1.0;
            for (int d = 0; d < D; ++d) {
                const double iv = 1.0 / (si[d] + sj[d]);
                const double dm = mui[d] - muj[d];
                esum += dm * iv * dm;
                m[d] = iv * (si[d] * muj[d] + sj[d] * mui[d]);
                v[d] = 2.0 * si[d] * iv * sj[d];
                th[d] = p[o.th + (i * Q + q) * D + d] - p[o.th + (j * Q + q) * D + d];
                prod *= v[d];
            }
            const double mag = p[o.w + i * Q + q] * p[o.w + j * Q + q] * exp(-MOGP_PI * MOGP_PI * esum);
            alpha = mag * twopi_pow * sqrt(prod);
            phi = p[o.ph + i * Q + q] - p[o.ph + j * Q + q];
        }
    } else if (kind == MOGP_KIND_SM) {                  // singleoutput.py:598-600: one term per (q, d)
        const SmOff o = sm_off(C, Q, D);
        const int q = r / D, d = r % D;
        if (i == j) {
            alpha = p[o.mag + i * Q + q];
            v[d] = 4.0 * MOGP_PI * MOGP_PI * p[o.var + (i * Q + q) * D + d];
            m[d] = p[o.mu + (i * Q + q) * D + d];
        }
    } else {                                            // CONV, multioutput.py:538-547
        const ConvOff o = conv_off(C, Q, D);
        const int q = r;
        double pb = 1.0, pv = 1.0;
        for (int d = 0; d < D; ++d) {
            const double b = p[o.base + q * D + d];
            const double V = p[o.var + (q * C + i) * D + d] + p[o.var + (q * C + j) * D + d] + b;
            v[d] = 1.0 / V;
            pb *= b; pv *= V;
        }
        alpha = p[o.w + q * C + i] * p[o.w + q * C + j] * sqrt(pb / pv);
    }
    out[0] = alpha;
    out[1] = phi;
}

// K_diag as the reference API returns it (multioutput.py:206-210, singleoutput.py:602-605,
// multioutput.py:549-553): for SM this is sum_q magnitude_q whatever D is.
__host__ __device__ inline double kdiag_api_value(int kind, int C, int Q, int D, const double* p,
                                                  const double* comps, int R, int c) {
    const int st = comp_stride(kind, D);
    double s = 0.0;
    if (kind == MOGP_KIND_SM) {
        const SmOff o = sm_off(C, Q, D);
        for (int q = 0; q < Q; ++q) s += p[o.mag + c * Q + q];
    } else if (kind_family(kind) == MOGP_KIND_SMLMC) {
        // LMC.Ksub_diag over SpectralKernel.K_diag (multioutput.py:497-502, singleoutput.py:558-561): sum_q (sum_s w^2) magnitude_q,
        // whatever D is (like SM, the reference's K_diag does not carry the factor D its K has for D > 1)
        const int Rq = kind_rq(kind);
        const LmcOff o = lmc_off(C, Q, Rq, D);
        for (int q = 0; q < Q; ++q) {
            double w = 0.0;
            for (int t = 0; t < Rq; ++t) w += p[o.w + (c * Q + q) * Rq + t] * p[o.w + (c * Q + q) * Rq + t];
            s += w * p[o.mag + q];
        }
    } else {
        for (int r = 0; r < R; ++r) s += comps[(size_t)((c * C + c) * R + r) * st];
    }
    return s;
}

// ---- chain rule ----------------------------------------------------------------------
// gsum: per lower pair pl = i(i+1)/2 + j, per component r, a record of comp_stride(kind, D) sums
//   [S0 = sum W E C, S4 = sum W E Sn, S1[d] = sum W E C u_d^2, S2[d] = sum W E Sn u_d, S3[d] = sum W E C u_d]
// (W already carries the symmetric weight).  adj[c] is added to S0 of the diagonal pair (c,c)
// (relative-jitter term).  `owner` enumerates disjoint output slices, see n_chain_owners().
__host__ __device__ inline int n_chain_owners(int kind, int C, int Q) {
    if (kind_family(kind) >= MOGP_KIND_CSM) return n_chain_owners_next(kind_family(kind), C, Q, kind_rq(kind));
    return kind == MOGP_KIND_CONV ? Q * C + Q : C * Q;
}

__host__ __device__ inline void chain_owner(int kind, int C, int Q, int D, const double* __restrict__ p,
                                            const double* __restrict__ comps, const double* __restrict__ gsum,
                                            const double* __restrict__ adj, int owner, double* __restrict__ g) {
    if (kind_family(kind) >= MOGP_KIND_CSM) {
        chain_owner_next(kind_family(kind), C, Q, kind_rq(kind), D, p, comps, gsum, adj, owner, g);
        return;
    }
    const int st = comp_stride(kind, D);
    const double PI2 = MOGP_PI * MOGP_PI;
    if (kind == MOGP_KIND_MOSM) {
        const MosmOff o = mosm_off(C, Q, D);
        const int R = Q, c = owner / Q, q = owner % Q;
        double gw = 0.0, gph = 0.0;
        double gmu[MOGP_MAX_D], gs[MOGP_MAX_D], gth[MOGP_MAX_D];
        for (int d = 0; d < D; ++d) gmu[d] = gs[d] = gth[d] = 0.0;
        for (int other = 0; other < C; ++other) {
            const int i = c > other ? c : other, j = c > other ? other : c;
            const double* S = gs_rec(gsum, i, j, R, st, q);
            const double* cp = comps + (size_t)((i * C + j) * R + q) * st;
            const double alpha = cp[0];
            const double* v = cp + 2; const double* m = cp + 2 + D;
            double S0 = S[0];
            const double S4 = S[1];
            if (i == j) S0 += adj[c];
            const double aGa = alpha * S0;
            const double Gph = -2.0 * MOGP_PI * alpha * S4;
            const double wi = p[o.w + i * Q + q], wj = p[o.w + j * Q + q];
            for (int side = 0; side < 2; ++side) {
                // side 0: c plays the row channel i, side 1: c plays the column channel j
                if (i != j && ((side == 0) != (c == i))) continue;
                gw += aGa / (side == 0 ? wi : wj);
                gph += side == 0 ? Gph : -Gph;
                for (int d = 0; d < D; ++d) {
                    const double si = p[o.var + (i * Q + q) * D + d], sj = p[o.var + (j * Q + q) * D + d];
                    const double mui = p[o.mu + (i * Q + q) * D + d], muj = p[o.mu + (j * Q + q) * D + d];
                    const double iv = 1.0 / (si + sj), dm = mui - muj;
                    const double Gv = -0.5 * alpha * S[2 + d];
                    const double Gm = -2.0 * MOGP_PI * alpha * S[2 + D + d];
                    const double Gth = alpha * (-v[d] * S[2 + 2 * D + d] - 2.0 * MOGP_PI * m[d] * S4);
                    if (side == 0) {
                        gmu[d] += aGa * (-2.0 * PI2 * dm * iv) + Gm * sj * iv;
                        gs[d] += aGa * (PI2 * dm * dm * iv * iv + 0.5 * sj * iv / si) + Gv * (2.0 * sj * sj * iv * iv)
                                 + Gm * (-sj * dm * iv * iv);
                        gth[d] += Gth;
                    } else {
                        gmu[d] += aGa * (2.0 * PI2 * dm * iv) + Gm * si * iv;
                        gs[d] += aGa * (PI2 * dm * dm * iv * iv + 0.5 * si * iv / sj) + Gv * (2.0 * si * si * iv * iv)
                                 + Gm * (si * dm * iv * iv);
                        gth[d] -= Gth;
                    }
                }
            }
        }
        g[o.w + c * Q + q] = gw;
        g[o.ph + c * Q + q] = gph;
        for (int d = 0; d < D; ++d) {
            g[o.mu + (c * Q + q) * D + d] = gmu[d];
            g[o.var + (c * Q + q) * D + d] = gs[d];
            g[o.th + (c * Q + q) * D + d] = gth[d];
        }
    } else if (kind == MOGP_KIND_SM) {
        const SmOff o = sm_off(C, Q, D);
        const int R = Q * D, c = owner / Q, q = owner % Q;
        double gmag = 0.0;
        for (int d = 0; d < D; ++d) {
            const int r = q * D + d;
            const double* S = gs_rec(gsum, c, c, R, st, r);
            const double alpha = comps[(size_t)((c * C + c) * R + r) * st];
            gmag += S[0] + adj[c];
            g[o.var + (c * Q + q) * D + d] = -0.5 * alpha * S[2 + d] * 4.0 * PI2;
            g[o.mu + (c * Q + q) * D + d] = -2.0 * MOGP_PI * alpha * S[2 + D + d];
        }
        g[o.mag + c * Q + q] = gmag;
    } else {
        const ConvOff o = conv_off(C, Q, D);
        const int R = Q;
        if (owner < Q * C) {
            const int q = owner / C, c = owner % C;
            double gw = 0.0, gv[MOGP_MAX_D];
            for (int d = 0; d < D; ++d) gv[d] = 0.0;
            for (int other = 0; other < C; ++other) {
                const int i = c > other ? c : other, j = c > other ? other : c;
                const double* S = gs_rec(gsum, i, j, R, st, q);
                const double* cp = comps + (size_t)((i * C + j) * R + q) * st;
                const double alpha = cp[0];
                double S0 = S[0];
                if (i == j) S0 += adj[c];
                const double aGa = alpha * S0;
                const double mult = (i == j) ? 2.0 : 1.0;   // c is both the row and the column channel
                gw += mult * aGa / p[o.w + q * C + c];
                for (int d = 0; d < D; ++d) {
                    const double vinv = cp[2 + d];           // 1/V_d
                    const double Gv = -0.5 * alpha * S[2 + d];
                    gv[d] += mult * (aGa * (-0.5 * vinv) + Gv * (-vinv * vinv));
                }
            }
            g[o.w + q * C + c] = gw;
            for (int d = 0; d < D; ++d) g[o.var + (q * C + c) * D + d] = gv[d];
        } else {
            const int q = owner - Q * C;
            double gb[MOGP_MAX_D];
            for (int d = 0; d < D; ++d) gb[d] = 0.0;
            for (int i = 0; i < C; ++i)
                for (int j = 0; j <= i; ++j) {
                    const double* S = gs_rec(gsum, i, j, R, st, q);
                    const double* cp = comps + (size_t)((i * C + j) * R + q) * st;
                    const double alpha = cp[0];
                    double S0 = S[0];
                    if (i == j) S0 += adj[i];
                    const double aGa = alpha * S0;
                    for (int d = 0; d < D; ++d) {
                        const double vinv = cp[2 + d];
                        const double Gv = -0.5 * alpha * S[2 + d];
                        gb[d] += aGa * (0.5 / p[o.base + q * D + d] - 0.5 * vinv) + Gv * (-vinv * vinv);
                    }
                }
            for (int d = 0; d < D; ++d) g[o.base + q * D + d] = gb[d];
        }
    }
}
