// Per channel-pair component tables and chain rules of the CSM, SM-LMC and uMOSM kernel families, in the same derived
// form and with the same record / owner conventions as the MOSM / SM / CONV tables of covmath.cuh, which includes this
// file and dispatches to it from pair_comp() / chain_owner() (families >= MOGP_KIND_CSM).
//   CSM    MixtureKernel of Q CrossSpectralKernel (mogptk/gpr/multioutput.py:428-454):
//          packed  amplitude (Q,C,Rq) | mean (Q,D) | variance (Q,D) | shift (Q,C,Rq);  R = Q*Rq components r = q*Rq + s
//   SMLMC  LinearModelOfCoregionalizationKernel of Q SpectralKernel (gpr/multioutput.py:490-502, singleoutput.py:550-561):
//          packed  weight (C,Q,Rq) | magnitude (Q) | mean (Q,D) | variance (Q,D);       R = Q*D  components r = q*D + d
//   UMOSM  MixtureKernel of Q UncoupledMultiOutputSpectralKernel (gpr/multioutput.py:261-293):
//          packed  weight (Q,C,C; lower triangle used) | mean (Q,C,D) | variance (Q,C,D) | delay (Q,C,D) | phase (Q,C)
//   MOHSM  MixtureKernel of Q MultiOutputHarmonizableSpectralKernel (gpr/multioutput.py:353-395): the MOSM-like stationary
//          factor times a Gaussian window exp(-1/2 l_ij sum_d ((x_d + x'_d)/2 - c_d)^2) in the mid-point; the comp record
//          carries [l_ij, c[D]] after theta, the gradient-sum record [S5 = sum W E C sum_d s_d^2, S6[d] = sum W E C s_d]
//          packed  weight (Q,C) | mean (Q,C,D) | variance (Q,C,D) | lengthscale (Q,C) | center (Q,D) | delay (Q,C,D) | phase (Q,C)
#pragma once

struct CsmOff { int amp, mu, var, sh; };
__host__ __device__ inline CsmOff csm_off(int C, int Q, int Rq, int D) {
    CsmOff o; o.amp = 0; o.mu = Q * C * Rq; o.var = o.mu + Q * D; o.sh = o.var + Q * D; return o;
}
struct LmcOff { int w, mag, mu, var; };
__host__ __device__ inline LmcOff lmc_off(int C, int Q, int Rq, int D) {
    LmcOff o; o.w = 0; o.mag = C * Q * Rq; o.mu = o.mag + Q; o.var = o.mu + Q * D; return o;
}
struct UmosmOff { int w, mu, var, th, ph; };
__host__ __device__ inline UmosmOff umosm_off(int C, int Q, int D) {
    UmosmOff o; o.w = 0; o.mu = Q * C * C; o.var = o.mu + Q * C * D; o.th = o.var + Q * C * D; o.ph = o.th + Q * C * D; return o;
}
struct MohsmOff { int w, mu, var, ls, ctr, th, ph; };
__host__ __device__ inline MohsmOff mohsm_off(int C, int Q, int D) {
    MohsmOff o; o.w = 0; o.mu = Q * C; o.var = o.mu + Q * C * D; o.ls = o.var + Q * C * D; o.ctr = o.ls + Q * C;
    o.th = o.ctr + Q * D; o.ph = o.th + Q * C * D; return o;
}
__host__ __device__ inline int next_num_params(int kind, int C, int Q, int Rq, int D) {
    if (kind == MOGP_KIND_MOHSM) return Q * (3 * C + 3 * C * D + D);
    if (kind == MOGP_KIND_UMOSM) return Q * C * C + 3 * Q * C * D + Q * C;
    return kind == MOGP_KIND_CSM ? 2 * Q * C * Rq + 2 * Q * D : C * Q * Rq + Q + 2 * Q * D;
}
__host__ __device__ inline int next_num_comps(int kind, int Q, int Rq, int D) {
    return (kind == MOGP_KIND_UMOSM || kind == MOGP_KIND_MOHSM) ? Q : (kind == MOGP_KIND_CSM ? Q * Rq : Q * D);
}
// (L L^T)_ij of the lower triangle L of the q-th C x C weight matrix
__host__ __device__ inline double umosm_mag(const double* w, int C, int i, int j) {
    double m = 0.0;
    const int kmax = i < j ? i : j;
    for (int k = 0; k <= kmax; ++k) m += w[i * C + k] * w[j * C + k];
    return m;
}

// comp record: [alpha, phi, v[D], m[D], theta[D]]
__host__ __device__ inline void pair_comp_next(int kind, int C, int Q, int Rq, int D, const double* __restrict__ p, int i,
                                               int j, int r, double* __restrict__ out) {
    double* v = out + 2;
    double* m = out + 2 + D;
    double* th = out + 2 + 2 * D;
    for (int d = 0; d < D; ++d) v[d] = m[d] = th[d] = 0.0;
    if (kind == MOGP_KIND_MOHSM) {                       // multioutput.py:358-387; the i == j branch (:359-367) is the same formula
        const MohsmOff o = mohsm_off(C, Q, D);
        const int q = r;
        const double* mui = p + o.mu + (q * C + i) * D; const double* muj = p + o.mu + (q * C + j) * D;
        const double* si = p + o.var + (q * C + i) * D; const double* sj = p + o.var + (q * C + j) * D;
        double esum = 0.0, prod = 1.0;
        for (int d = 0; d < D; ++d) {
            const double iv = 1.0 / (si[d] + sj[d]);
            const double dm = mui[d] - muj[d];
            esum += dm * iv * dm;
            m[d] = iv * (si[d] * muj[d] + sj[d] * mui[d]);
            v[d] = 2.0 * si[d] * iv * sj[d];
            th[d] = p[o.th + (q * C + i) * D + d] - p[o.th + (q * C + j) * D + d];
            prod *= v[d];
            out[3 + 3 * D + d] = p[o.ctr + q * D + d];
        }
        const double li = p[o.ls + q * C + i] * p[o.ls + q * C + i], lj = p[o.ls + q * C + j] * p[o.ls + q * C + j];
        const double ell = 2.0 * li * (1.0 / (li + lj)) * lj;                                           // :379
        out[0] = p[o.w + q * C + i] * p[o.w + q * C + j] * exp(-MOGP_PI * MOGP_PI * esum) * pow(2.0 * MOGP_PI, (double)D) *
                 sqrt(prod) * pow(sqrt(ell), (double)D);                                                // :375,382 (twopi = (2 pi)^D, :350)
        out[1] = (p[o.ph + q * C + i] - p[o.ph + q * C + j]) / (2.0 * MOGP_PI);                         // phase outside the 2 pi factor (:385)
        out[2 + 3 * D] = ell;
    } else if (kind == MOGP_KIND_UMOSM) {                // multioutput.py:266-286; the i == j branch is the same formula
        const UmosmOff o = umosm_off(C, Q, D);
        const int q = r;
        const double* mui = p + o.mu + (q * C + i) * D; const double* muj = p + o.mu + (q * C + j) * D;
        const double* si = p + o.var + (q * C + i) * D; const double* sj = p + o.var + (q * C + j) * D;
        double esum = 0.0, prod = 1.0;
        for (int d = 0; d < D; ++d) {
            const double iv = 1.0 / (si[d] + sj[d]);
            const double dm = mui[d] - muj[d];
            esum += dm * iv * dm;
            m[d] = iv * (si[d] * muj[d] + sj[d] * mui[d]);
            v[d] = 2.0 * si[d] * iv * sj[d];
            th[d] = p[o.th + (q * C + i) * D + d] - p[o.th + (q * C + j) * D + d];
            prod *= v[d];
        }
        out[0] = umosm_mag(p + o.w + q * C * C, C, i, j) * exp(-MOGP_PI * MOGP_PI * esum) * pow(2.0 * MOGP_PI, 0.5 * (double)D) * sqrt(prod);
        out[1] = (p[o.ph + q * C + i] - p[o.ph + q * C + j]) / (2.0 * MOGP_PI);      // the phase sits outside the 2 pi factor (:285)
    } else if (kind == MOGP_KIND_CSM) {                  // multioutput.py:432-447; i == j is the same formula
        const CsmOff o = csm_off(C, Q, Rq, D);
        const int q = r / Rq, s = r % Rq;
        out[0] = sqrt(p[o.amp + (q * C + i) * Rq + s] * p[o.amp + (q * C + j) * Rq + s]);
        out[1] = p[o.sh + (q * C + i) * Rq + s] - p[o.sh + (q * C + j) * Rq + s];
        for (int d = 0; d < D; ++d) { v[d] = p[o.var + q * D + d]; m[d] = p[o.mu + q * D + d]; }
    } else {                                             // multioutput.py:493-495 over singleoutput.py:554-556
        const LmcOff o = lmc_off(C, Q, Rq, D);
        const int q = r / D, d = r % D;
        double w = 0.0;
        for (int s = 0; s < Rq; ++s) w += p[o.w + (i * Q + q) * Rq + s] * p[o.w + (j * Q + q) * Rq + s];
        out[0] = w * p[o.mag + q];
        out[1] = 0.0;
        v[d] = 4.0 * MOGP_PI * MOGP_PI * p[o.var + q * D + d];
        m[d] = p[o.mu + q * D + d];
    }
}

// owners: CSM: [0, Q*C*Rq) one (q, c, s) each (amplitude, shift), then Q owners (mean, variance of q);
//         SMLMC: [0, C*Q*Rq) one (c, q, s) each (weight), then Q owners (magnitude, mean, variance of q)
//         UMOSM: Q*C owners (q, c): row c of the weight matrix, mean, variance, delay, phase of channel c
//         MOHSM: Q*C owners (q, c): weight, mean, variance, lengthscale, delay, phase of channel c, then Q owners (center of q)
__host__ __device__ inline int n_chain_owners_next(int kind, int C, int Q, int Rq) {
    if (kind == MOGP_KIND_MOHSM) return Q * C + Q;
    if (kind == MOGP_KIND_UMOSM) return Q * C;
    return (kind == MOGP_KIND_CSM ? Q * C * Rq : C * Q * Rq) + Q;
}

__host__ __device__ inline void chain_owner_next(int kind, int C, int Q, int Rq, int D, const double* __restrict__ p,
                                                 const double* __restrict__ comps, const double* __restrict__ gsum,
                                                 const double* __restrict__ adj, int owner, double* __restrict__ g) {
    const int st = comp_stride(kind, D);
    const int R = next_num_comps(kind, Q, Rq, D);
    if (kind == MOGP_KIND_MOHSM) {
        // (the relative-jitter term of the row-dependent Gram diagonal is folded into gsum by the caller: adj is not used)
        const MohsmOff o = mohsm_off(C, Q, D);
        const double PI2 = MOGP_PI * MOGP_PI;
        if (owner >= Q * C) {                            // center of component q: d K / d c_d = K l s_d
            const int q = owner - Q * C;
            double gc[MOGP_MAX_D];
            for (int d = 0; d < D; ++d) gc[d] = 0.0;
            for (int i = 0; i < C; ++i)
                for (int j = 0; j <= i; ++j) {
                    const double* S = gs_rec(gsum, i, j, R, st, q);
                    const double* cp = comps + (size_t)((i * C + j) * R + q) * st;
                    for (int d = 0; d < D; ++d) gc[d] += cp[0] * cp[2 + 3 * D] * S[3 + 3 * D + d];
                }
            for (int d = 0; d < D; ++d) g[o.ctr + q * D + d] = gc[d];
            return;
        }
        const int q = owner / C, c = owner % C;
        double gw = 0.0, gph = 0.0, gls = 0.0, gmu[MOGP_MAX_D], gs[MOGP_MAX_D], gth[MOGP_MAX_D];
        for (int d = 0; d < D; ++d) gmu[d] = gs[d] = gth[d] = 0.0;
        for (int other = 0; other < C; ++other) {
            const int i = c > other ? c : other, j = c > other ? other : c;
            const double* S = gs_rec(gsum, i, j, R, st, q);
            const double* cp = comps + (size_t)((i * C + j) * R + q) * st;
            const double alpha = cp[0];
            const double* v = cp + 2; const double* m = cp + 2 + D;
            const double ell = cp[2 + 3 * D];
            const double S0 = S[0], S4 = S[1], S5 = S[2 + 3 * D];
            const double aGa = alpha * S0;
            const double Gph = -alpha * S4;                  // d loss / d (phase_i - phase_j)
            const double Gell = aGa * 0.5 * (double)D / ell - 0.5 * alpha * S5;      // alpha ~ l^(D/2); window exp(-1/2 l sum s^2)
            const double wi = p[o.w + q * C + i], wj = p[o.w + q * C + j];
            const double lsi = p[o.ls + q * C + i], lsj = p[o.ls + q * C + j];
            const double li = lsi * lsi, lj = lsj * lsj, il = 1.0 / (li + lj);
            for (int side = 0; side < 2; ++side) {
                if (i != j && ((side == 0) != (c == i))) continue;
                gw += aGa / (side == 0 ? wi : wj);
                gph += side == 0 ? Gph : -Gph;
                // l_ij = 2 li lj / (li + lj), li = ls_i^2: d l_ij / d ls_i = 2 lj^2 / (li + lj)^2 * 2 ls_i
                gls += Gell * (side == 0 ? 2.0 * lj * lj * il * il * 2.0 * lsi : 2.0 * li * li * il * il * 2.0 * lsj);
                for (int d = 0; d < D; ++d) {
                    const double si = p[o.var + (q * C + i) * D + d], sj = p[o.var + (q * C + j) * D + d];
                    const double mui = p[o.mu + (q * C + i) * D + d], muj = p[o.mu + (q * C + j) * D + d];
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
        g[o.w + q * C + c] = gw;
        g[o.ph + q * C + c] = gph;
        g[o.ls + q * C + c] = gls;
        for (int d = 0; d < D; ++d) {
            g[o.mu + (q * C + c) * D + d] = gmu[d];
            g[o.var + (q * C + c) * D + d] = gs[d];
            g[o.th + (q * C + c) * D + d] = gth[d];
        }
    } else if (kind == MOGP_KIND_UMOSM) {
        const UmosmOff o = umosm_off(C, Q, D);
        const double PI2 = MOGP_PI * MOGP_PI;
        const int q = owner / C, c = owner % C;
        const double* wq = p + o.w + q * C * C;
        double gph = 0.0, gL[64], gmu[MOGP_MAX_D], gs[MOGP_MAX_D], gth[MOGP_MAX_D];
        for (int k = 0; k < C; ++k) gL[k] = 0.0;
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
            // alpha = M_ij * rest: d loss / d M_ij = S0 * rest, rest recomputed so that M_ij = 0 is harmless
            double esum = 0.0, prod = 1.0;
            for (int d = 0; d < D; ++d) {
                const double si = p[o.var + (q * C + i) * D + d], sj = p[o.var + (q * C + j) * D + d];
                const double dm = p[o.mu + (q * C + i) * D + d] - p[o.mu + (q * C + j) * D + d];
                esum += dm * dm / (si + sj);
                prod *= v[d];
            }
            const double GM = S0 * exp(-PI2 * esum) * pow(2.0 * MOGP_PI, 0.5 * (double)D) * sqrt(prod);
            const int kmax = c < other ? c : other;
            for (int k = 0; k <= kmax; ++k) gL[k] += GM * (other == c ? 2.0 * wq[c * C + k] : wq[other * C + k]);
            const double Gph = -alpha * S4;                  // d loss / d (phase_i - phase_j): -2 pi alpha S4 / (2 pi)
            for (int side = 0; side < 2; ++side) {
                if (i != j && ((side == 0) != (c == i))) continue;
                gph += side == 0 ? Gph : -Gph;
                for (int d = 0; d < D; ++d) {
                    const double si = p[o.var + (q * C + i) * D + d], sj = p[o.var + (q * C + j) * D + d];
                    const double mui = p[o.mu + (q * C + i) * D + d], muj = p[o.mu + (q * C + j) * D + d];
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
        for (int k = 0; k < C; ++k) g[o.w + (q * C + c) * C + k] = k <= c ? gL[k] : 0.0;
        g[o.ph + q * C + c] = gph;
        for (int d = 0; d < D; ++d) {
            g[o.mu + (q * C + c) * D + d] = gmu[d];
            g[o.var + (q * C + c) * D + d] = gs[d];
            g[o.th + (q * C + c) * D + d] = gth[d];
        }
    } else if (kind == MOGP_KIND_CSM) {
        const CsmOff o = csm_off(C, Q, Rq, D);
        if (owner < Q * C * Rq) {
            const int q = owner / (C * Rq), c = (owner / Rq) % C, s = owner % Rq, r = q * Rq + s;
            double ga = 0.0, gsh = 0.0;
            for (int other = 0; other < C; ++other) {
                const int i = c > other ? c : other, j = c > other ? other : c;
                const double* S = gs_rec(gsum, i, j, R, st, r);
                const double alpha = comps[(size_t)((i * C + j) * R + r) * st];
                double S0 = S[0];
                if (i == j) S0 += adj[c];
                // alpha = sqrt(a_i a_j): d alpha / d a_c = alpha / (2 a_c), twice that on the diagonal pair (= 1)
                ga += (i == j ? 2.0 : 1.0) * S0 * alpha / (2.0 * p[o.amp + (q * C + c) * Rq + s]);
                if (i != j) gsh += (c == i ? 1.0 : -1.0) * (-2.0 * MOGP_PI * alpha * S[1]);
            }
            g[o.amp + (q * C + c) * Rq + s] = ga;
            g[o.sh + (q * C + c) * Rq + s] = gsh;
        } else {
            const int q = owner - Q * C * Rq;
            double gm[MOGP_MAX_D], gv[MOGP_MAX_D];
            for (int d = 0; d < D; ++d) gm[d] = gv[d] = 0.0;
            for (int i = 0; i < C; ++i)
                for (int j = 0; j <= i; ++j)
                    for (int s = 0; s < Rq; ++s) {
                        const int r = q * Rq + s;
                        const double* S = gs_rec(gsum, i, j, R, st, r);
                        const double alpha = comps[(size_t)((i * C + j) * R + r) * st];
                        for (int d = 0; d < D; ++d) {
                            gv[d] += -0.5 * alpha * S[2 + d];
                            gm[d] += -2.0 * MOGP_PI * alpha * S[2 + D + d];
                        }
                    }
            for (int d = 0; d < D; ++d) { g[o.mu + q * D + d] = gm[d]; g[o.var + q * D + d] = gv[d]; }
        }
    } else {
        const LmcOff o = lmc_off(C, Q, Rq, D);
        if (owner < C * Q * Rq) {
            const int c = owner / (Q * Rq), q = (owner / Rq) % Q, s = owner % Rq;
            double gw = 0.0;
            for (int other = 0; other < C; ++other) {
                const int i = c > other ? c : other, j = c > other ? other : c;
                double Ga = 0.0;                                         // d loss / d (sum_s w_i w_j) / magnitude
                for (int d = 0; d < D; ++d) {
                    const double* S = gs_rec(gsum, i, j, R, st, q * D + d);
                    Ga += S[0] + (i == j ? adj[c] : 0.0);
                }
                gw += (i == j ? 2.0 : 1.0) * Ga * p[o.mag + q] * p[o.w + (other * Q + q) * Rq + s];
            }
            g[o.w + (c * Q + q) * Rq + s] = gw;
        } else {
            const int q = owner - C * Q * Rq;
            double gmag = 0.0, gm[MOGP_MAX_D], gv[MOGP_MAX_D];
            for (int d = 0; d < D; ++d) gm[d] = gv[d] = 0.0;
            for (int i = 0; i < C; ++i)
                for (int j = 0; j <= i; ++j) {
                    double w = 0.0;
                    for (int s = 0; s < Rq; ++s) w += p[o.w + (i * Q + q) * Rq + s] * p[o.w + (j * Q + q) * Rq + s];
                    for (int d = 0; d < D; ++d) {
                        const double* S = gs_rec(gsum, i, j, R, st, q * D + d);
                        const double alpha = comps[(size_t)((i * C + j) * R + q * D + d) * st];
                        gmag += (S[0] + (i == j ? adj[i] : 0.0)) * w;
                        gv[d] += -0.5 * alpha * S[2 + d] * 4.0 * MOGP_PI * MOGP_PI;
                        gm[d] += -2.0 * MOGP_PI * alpha * S[2 + D + d];
                    }
                }
            g[o.mag + q] = gmag;
            for (int d = 0; d < D; ++d) { g[o.mu + q * D + d] = gm[d]; g[o.var + q * D + d] = gv[d]; }
        }
    }
}
