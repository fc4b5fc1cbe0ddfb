#!/usr/bin/env python
"""CPU prototype (numpy, exact integer arithmetic) of fp64 GEMM emulation on an int8 tensor pipe -- groundwork for
moving the O(N^3) stages (trailing updates, L^-1, K^-1 = L^-T L^-1) from DMMA (37 TFLOP/s measured) to tcgen05
kind::i8 (dense int8 peak 4.5 POP/s on B200).  Ozaki-style slicing with a shared exponent per row of A / column of B:

    A[i, :] = 2^ea[i] * sum_s  As[s][i, :] * 2^(-b (s+1)),   As[s] int8 in [-2^(b-1) .. 2^(b-1)]  (signed digits)
    C = sum_{s+t < S} 2^(-b (s+t+2)) * (As[s] @ Bs[t])  scaled by 2^(ea[i] + eb[j]),  every int8 GEMM exact in int32

S slices -> S (S + 1) / 2 int8 GEMMs.  Reports the error of K^-1 = L^-T L^-1 for GP covariances of increasing
condition number against an extended-precision reference, next to the error of the plain fp64 GEMM, as a function of
S and b; with --pipeline the whole O(N^3) structure of the step (blocked Cholesky updates, block-doubling inverse,
K^-1) runs on the emulated GEMM and LML / gradient-like quantities are compared end to end.
Usage: python tools/ozaki_proto.py [N] | --pipeline [N]"""
import sys

import numpy as np



def split_rows(A, S, b):
    """Signed-digit slices of every row of A with one exponent per row: returns (digits [S, m, k] int64, exponents [m])."""
    amax = np.abs(A).max(axis=1)
    e = np.where(amax > 0, np.floor(np.log2(np.where(amax > 0, amax, 1.0))) + 1, 0).astype(np.int64)   # |A| < 2^e
    R = A / np.exp2(e)[:, None]                         # in (-1, 1)
    digs = np.zeros((S,) + A.shape, dtype=np.int64)
    for s in range(S):
        R = R * (1 << b)
        d = np.rint(R)                                  # round to nearest: digits in [-2^(b-1), 2^(b-1)]
        digs[s] = d.astype(np.int64)
        R = R - d                                       # exact in fp64 (the remainder has fewer significant bits)
    return digs, e


def ozaki_gemm(A, B, S, b, full=False):
    """C ~= A @ B from S slices of b bits; pairs with s + t < S (or all S^2 pairs when full)."""
    da, ea = split_rows(A, S, b)
    db, eb = split_rows(B.T, S, b)
    m, n = A.shape[0], B.shape[1]
    C = np.zeros((m, n), dtype=np.longdouble)
    n_gemm = 0
    for s in range(S):
        for t in range(S):
            if not full and s + t >= S:
                continue
            P = da[s] @ db[t].T                         # exact: |P| <= k 2^(2b-2) < 2^31 for k <= 2^(33-2b)
            assert np.abs(P).max() < 2 ** 31
            C += P.astype(np.longdouble) * np.longdouble(2.0) ** (-b * (s + t + 2))
            n_gemm += 1
    C = C * (np.longdouble(2.0) ** ea)[:, None] * (np.longdouble(2.0) ** eb)[None, :]
    return C.astype(np.float64), n_gemm


def gp_covariance(n, noise, seed=0):
    """A GP covariance of the kind the path factorises: two-channel spectral-mixture-like kernel on sorted 1-D inputs
    plus a noise diagonal (smaller noise = worse conditioning)."""
    rng = np.random.default_rng(seed)
    x = np.sort(rng.uniform(0.0, 10.0, n))
    tau = x[:, None] - x[None, :]
    K = 1.3 * np.exp(-0.5 * tau ** 2 / 0.4 ** 2) * np.cos(2 * np.pi * 0.7 * tau) + 0.5 * np.exp(-0.5 * tau ** 2 / 2.0 ** 2)
    K[np.arange(n), np.arange(n)] += noise
    return K


def ozaki_gemm_fast(A, B, S, b):
    """Same arithmetic as ozaki_gemm (pairs s + t < S) with the exact integer products done by the fp64 BLAS
    (every partial sum is an integer below 2^53, so the float matmul is exact); anti-diagonals share an accumulator."""
    da, ea = split_rows(A, S, b)
    db, eb = split_rows(B.T, S, b)
    da = da.astype(np.float64)
    db = db.astype(np.float64)
    C = np.zeros((A.shape[0], B.shape[1]))
    for d in range(S - 1, -1, -1):                    # smallest weights first
        P = np.zeros_like(C)
        for s in range(d + 1):
            P += da[s] @ db[d - s].T
        C += P * 2.0 ** (-b * (d + 2))
    return C * np.exp2(ea)[:, None] * np.exp2(eb)[None, :]


def exact_gp_step(K, y, gemm, nb=64):
    """The O(N^3) structure of the product path with a pluggable GEMM: blocked right-looking Cholesky (panel work in
    fp64, trailing updates through `gemm`), L^-1 by block doubling, K^-1 = L^-T L^-1; returns (lml, tr W, W-weighted
    probe) where W = (K^-1 - a a^T) / 2 is what the gradient kernel consumes."""
    n = K.shape[0]
    A = K.copy()
    for k in range(0, n, nb):
        e = min(k + nb, n)
        A[k:e, k:e] = np.linalg.cholesky(A[k:e, k:e])
        if e < n:
            A[e:, k:e] = np.linalg.solve(A[k:e, k:e], A[e:, k:e].T).T
            A[e:, e:] -= gemm(A[e:, k:e], A[e:, k:e].T)
    L = np.tril(A)
    Linv = np.zeros_like(L)
    for k in range(0, n, nb):
        e = min(k + nb, n)
        Linv[k:e, k:e] = np.linalg.inv(L[k:e, k:e])
    s = nb
    while s < n:
        for o in range(0, n, 2 * s):
            m, h = o + s, min(o + 2 * s, n)
            if m >= n:
                continue
            T = gemm(L[m:h, o:m], Linv[o:m, o:m])
            Linv[m:h, o:m] = -gemm(Linv[m:h, m:h], T)
        s *= 2
    Kinv = gemm(np.ascontiguousarray(Linv.T), Linv)
    z = Linv @ y
    alpha = Linv.T @ z
    lml = -0.5 * n * np.log(2 * np.pi) - np.log(np.diag(L)).sum() - 0.5 * float(z @ z)
    W = 0.5 * (Kinv - np.outer(alpha, alpha))
    idx = np.arange(n)
    probe = float((W * np.cos(0.37 * (idx[:, None] - idx[None, :]))).sum())    # a dK/dtheta-like weighting
    return lml, float(np.trace(W)), probe


def pipeline(n):
    rng = np.random.default_rng(1)
    for noise in (1.0, 1e-2, 1e-4, 1e-6):
        K = gp_covariance(n, noise)
        y = rng.standard_normal(n)
        ref = exact_gp_step(K.astype(np.longdouble).astype(np.float64), y, lambda a, b_: (a.astype(np.longdouble) @ b_.astype(np.longdouble)).astype(np.float64))
        f64 = exact_gp_step(K, y, lambda a, b_: a @ b_)
        print("N=%d noise=%.0e cond(K)=%.1e: fp64 GEMMs: lml rel %.1e  trW rel %.1e  probe rel %.1e" % (
            n, noise, np.linalg.cond(K), abs(f64[0] - ref[0]) / abs(ref[0]), abs(f64[1] - ref[1]) / abs(ref[1]),
            abs(f64[2] - ref[2]) / abs(ref[2])))
        for S in (5, 6, 7, 8):
            got = exact_gp_step(K, y, lambda a, b_: ozaki_gemm_fast(a, b_, S, 7))
            print("   int8 slices S=%d (%2d GEMMs): lml rel %.1e  trW rel %.1e  probe rel %.1e" % (
                S, S * (S + 1) // 2, abs(got[0] - ref[0]) / abs(ref[0]), abs(got[1] - ref[1]) / abs(ref[1]),
                abs(got[2] - ref[2]) / abs(ref[2])))


def main(argv):
    if argv and argv[0] == "--pipeline":
        return pipeline(int(argv[1]) if len(argv) > 1 else 1024)
    n = int(argv[0]) if argv else 768
    for noise in (1.0, 1e-2, 1e-4):
        K = gp_covariance(n, noise)
        L = np.linalg.cholesky(K)
        Linv = np.linalg.inv(L)
        ref = (Linv.T.astype(np.longdouble) @ Linv.astype(np.longdouble)).astype(np.float64)
        f64 = Linv.T @ Linv
        scale = np.abs(ref).max()
        e64 = np.abs(f64 - ref).max()
        print("N=%d noise=%.0e cond(K)=%.1e  max|K^-1|=%.2e  fp64 GEMM error %.2e of the largest entry" % (
            n, noise, np.linalg.cond(K), scale, e64 / scale))
        for b in (6, 7):
            for S in (6, 7, 8, 9):
                C, ng = ozaki_gemm(np.ascontiguousarray(Linv.T), Linv, S, b)
                err = np.abs(C - ref).max()
                print("   b=%d S=%d (%2d int8 GEMMs): max error %.2e of the largest entry  (%.1f x the fp64 GEMM error)" % (
                    b, S, ng, err / scale, err / max(e64, 1e-300)))


if __name__ == "__main__":
    main(sys.argv[1:])
