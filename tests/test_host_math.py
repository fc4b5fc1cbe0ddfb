"""CPU: the closed forms shared by the CUDA kernels (csrc/covmath.cuh), run on the host
through the library's self-check hooks, against the oracle:
  * derived per channel-pair constants  -> K rebuilt in numpy equals the oracle K,
  * analytic chain rule                 -> gradient equals the oracle's autograd gradient."""
import ctypes as C

import numpy as np
import pytest
import torch

from conftest import golden_names, load_golden
from mogptk_b200 import _cabi
from mogptk_b200.engine import pack_params, unpack_grads, kernel_dims
from oracle import mogp_oracle as orc

CASES = [n for n in golden_names() if not n.startswith("cfg")]


def _ptr(a):
    return C.c_void_p(a.ctypes.data)


def comps_of(lib, g):
    kind = g["kind"]
    Cn, Q, D = kernel_dims(kind, g["params"])
    p = pack_params(kind, g["params"]).numpy().copy()
    st = 2 + 3 * D
    R = Q * D if kind == "SM" else Q
    comps = np.zeros(Cn * Cn * R * st)
    r = lib.mogp_host_pair_comps(_cabi.KIND[kind], Cn, Q, D, _ptr(p), _ptr(comps))
    assert r == R
    return p, comps.reshape(Cn, Cn, R, st), (Cn, Q, D, R, st)


def block_terms(comp, xa, xb, D):
    """E*C, E*S and u for one component of one pair; xa (n,D), xb (m,D)."""
    alpha, phi = comp[0], comp[1]
    v, m, th = comp[2:2 + D], comp[2 + D:2 + 2 * D], comp[2 + 2 * D:2 + 3 * D]
    u = xa[:, None, :] - xb[None, :, :] + th[None, None, :]
    E = np.exp(-0.5 * (u ** 2 * v).sum(-1))
    ang = 2 * np.pi * ((u * m).sum(-1) + phi)
    return alpha, E * np.cos(ang), E * np.sin(ang), u


@pytest.mark.parametrize("name", CASES)
def test_pair_constants_rebuild_K(lib, name):
    g = load_golden(name)
    p, comps, (Cn, Q, D, R, st) = comps_of(lib, g)
    X = g["X"]
    order = np.argsort(X[:, 0], kind="stable")
    Xs = X[order]
    off = np.concatenate([[0], np.cumsum(np.bincount(Xs[:, 0].astype(int), minlength=Cn))])
    N = X.shape[0]
    K = np.zeros((N, N))
    for i in range(Cn):
        for j in range(Cn):
            xa, xb = Xs[off[i]:off[i + 1], 1:], Xs[off[j]:off[j + 1], 1:]
            blk = np.zeros((xa.shape[0], xb.shape[0]))
            for r in range(R):
                a, EC, _, _ = block_terms(comps[i, j, r], xa, xb, D)
                blk += a * EC
            K[off[i]:off[i + 1], off[j]:off[j + 1]] = blk
    Kref = orc.K(g["kind"], g["params"], torch.tensor(Xs)).numpy()
    assert np.abs(K - Kref).max() <= 1e-13 * np.abs(Kref).max()


@pytest.mark.parametrize("name", CASES)
def test_chain_rule_matches_autograd(lib, name):
    g = load_golden(name)
    kind = g["kind"]
    p, comps, (Cn, Q, D, R, st) = comps_of(lib, g)
    X, y = g["X"], g["y"]
    order = np.argsort(X[:, 0], kind="stable")
    Xs, ys = X[order], y[order]
    dv = g["data_var"][order] if "data_var" in g else None
    off = np.concatenate([[0], np.cumsum(np.bincount(Xs[:, 0].astype(int), minlength=Cn))])
    N = X.shape[0]
    Xt = torch.tensor(Xs)
    Kt = orc._noisy_gram(kind, g["params"], g["sigma_t"], Xt, g["jitter"], dv).numpy()
    Kinv = np.linalg.inv(Kt)
    al = Kinv @ ys
    W = 0.5 * (Kinv - np.outer(al, al))
    # weighted sums per lower pair: [S0, S4, S1[D], S2[D], S3[D]]
    npl = Cn * (Cn + 1) // 2
    gsum = np.zeros((npl, R, st))
    for i in range(Cn):
        for j in range(i + 1):
            xa, xb = Xs[off[i]:off[i + 1], 1:], Xs[off[j]:off[j + 1], 1:]
            Wb = W[off[i]:off[i + 1], off[j]:off[j + 1]] * (1.0 if i == j else 2.0)
            for r in range(R):
                _, EC, ES, u = block_terms(comps[i, j, r], xa, xb, D)
                rec = gsum[i * (i + 1) // 2 + j, r]
                rec[0] = (Wb * EC).sum()
                rec[1] = (Wb * ES).sum()
                for d in range(D):
                    rec[2 + d] = (Wb * EC * u[..., d] ** 2).sum()
                    rec[2 + D + d] = (Wb * ES * u[..., d]).sum()
                    rec[2 + 2 * D + d] = (Wb * EC * u[..., d]).sum()
    trW = np.trace(W)
    nc = np.diff(off)
    adj = g["jitter"] / N * trW * nc.astype(float)
    grad = np.zeros(p.size)
    P = lib.mogp_host_chain(_cabi.KIND[kind], Cn, Q, D, _ptr(p), _ptr(gsum), _ptr(adj), _ptr(grad))
    assert P == p.size
    got = unpack_grads(kind, Cn, Q, D, torch.tensor(grad))
    _, ref = orc.loss_and_grad(kind, g["params"], g["sigma_t"], Xt, ys, g["jitter"], dv)
    for k, v in got.items():
        scale = max(float(ref[k].abs().max()), 1e-12)
        assert float((v - ref[k]).abs().max()) <= 1e-7 * scale, (k, v, ref[k])
    # noise gradient formula used by the finalize kernel
    sig = g["sigma"]
    gs = np.array([2 * sig[c] * (np.trace(W[off[c]:off[c + 1], off[c]:off[c + 1]]) + adj[c]) for c in range(Cn)])
    assert np.abs(gs - ref["sigma"].numpy()).max() <= 1e-7 * max(np.abs(ref["sigma"].numpy()).max(), 1e-12)


# ---------------------------------------------------------------------------------------------------------------
# Pipelined triangular inverse: the host-side plan (csrc/linalg.cu::build_inverse_plan) that potrf_padded releases
# panel by panel.  Replaces the level-batched loop of trtri_padded for N <= 4096; the arithmetic it schedules is
# torch.linalg.solve_triangular / cholesky_solve of the reference (mogptk/gpr/model.py:452,470).
def _inverse_plan(lib, nb):
    import ctypes as C
    cap = 3 * nb + 8
    buf = (C.c_int32 * (8 * cap))()
    n = lib.mogp_host_inverse_plan(nb, buf, cap)
    assert n > 0
    return [tuple(buf[8 * i + j] for j in range(8)) for i in range(n)]


@pytest.mark.parametrize("nb", [1, 2, 3, 4, 6, 8, 10, 32, 34, 64, 128])
def test_inverse_plan_is_a_valid_schedule(lib, nb):
    ops = _inverse_plan(lib, nb)
    assert len(ops) == nb + 2 * (nb - 1)
    ready = [o[4] for o in ops]
    assert ready == sorted(ready), "operations must be releasable panel by panel"
    assert ready[-1] == nb - 1
    done_at = {}
    for idx, (kind, lo, mid, hi, rdy, wait, done, level) in enumerate(ops):
        assert done not in done_at
        done_at[done] = idx
        if kind == 0:
            assert hi == lo + 1 and rdy == lo and wait == -1
        else:
            assert lo < mid < hi <= nb and (mid - lo) == 1 << level and (mid - lo) >= (hi - mid)
            assert wait in done_at and done_at[wait] < idx, "a dependency is issued before its consumer"
            assert rdy == (mid - 1 if kind == 1 else hi - 1)
    # the T product of a pair is issued before its second GEMM, on the same level
    for idx, o in enumerate(ops):
        if o[0] == 2:
            assert any(p[0] == 1 and p[1:4] == o[1:4] and p[7] == o[7] for p in ops[:idx])


@pytest.mark.parametrize("nb", [1, 2, 5, 8, 13])
def test_inverse_plan_computes_the_inverse(lib, nb):
    """Run the plan with numpy blocks (block size 3) in issue order, only ever touching what an operation may touch."""
    rng = np.random.default_rng(nb)
    bs, n = 3, 3 * nb
    L = np.tril(rng.standard_normal((n, n))) + 4.0 * np.eye(n)
    Linv = np.zeros((n, n))
    T = np.zeros((n, n))
    for kind, lo, mid, hi, rdy, wait, done, level in _inverse_plan(lib, nb):
        a, m, h = lo * bs, mid * bs, hi * bs
        if kind == 0:
            Linv[a:h, a:h] = np.linalg.inv(L[a:h, a:h])
        elif kind == 1:
            T[m:h, a:m] = L[m:h, a:m] @ Linv[a:m, a:m]
        else:
            Linv[m:h, a:m] = -Linv[m:h, m:h] @ T[m:h, a:m]
    assert np.allclose(Linv, np.linalg.inv(L), rtol=1e-10, atol=1e-12)


# ---------------------------------------------------------------------------------------------------------------
# host-side schedules of the pipelined / recursive factorisation (pure index arithmetic, checked without a GPU)
def _partition(lib, nb, G, S, taper):
    out = (C.c_int32 * (4 * 256))()
    n = lib.mogp_host_rowpipe_partition(nb, G, S, taper, out, 256)
    assert n > 0
    return np.array(out[:4 * n]).reshape(n, 4)


@pytest.mark.parametrize("nb", [2, 5, 6, 7, 16, 32, 33, 64])
@pytest.mark.parametrize("G,S,taper", [(1, 1, 0), (1, 4, 1), (2, 4, 0), (2, 8, 1), (4, 8, 1), (8, 8, 0), (1, 32, 1)])
def test_rowwise_pipeline_partition(lib, nb, G, S, taper):
    """Groups tile [0, nb) in order, every group lies inside its super-group, super-groups are whole groups, only the last
    group of a super-group may be ragged, and with taper the last super-groups do not grow."""
    a = _partition(lib, nb, G, S, taper)
    assert a[0, 0] == 0 and a[-1, 1] == nb
    assert all(a[i, 1] == a[i + 1, 0] for i in range(len(a) - 1))
    for lo, hi, slo, shi in a:
        assert slo <= lo < hi <= shi and hi - lo <= G
        assert (lo - slo) % G == 0
    sup = sorted(set((int(r[2]), int(r[3])) for r in a))
    assert sup[0][0] == 0 and sup[-1][1] == nb and all(sup[i][1] == sup[i + 1][0] for i in range(len(sup) - 1))
    if taper and len(sup) > 2:
        sizes = [b - a_ for a_, b in sup]
        assert sizes[-1] <= sizes[-2] <= max(sizes)


@pytest.mark.parametrize("nb,G,wmin", [(32, 1, 4), (32, 2, 4), (6, 1, 4), (5, 1, 2), (64, 1, 8), (32, 1, 1), (7, 2, 4), (2, 1, 4)])
def test_kinv_accumulation_chunks(lib, nb, G, wmin):
    """The halving chunks of the progressive K^-1 accumulation tile [0, nb), end at group boundaries and shrink towards the end."""
    ch = [(lib.mogp_host_kinv_chunk_start(nb, G, wmin, hi), hi) for hi in range(1, nb + 1)]
    ch = [(lo, hi) for lo, hi in ch if lo >= 0]
    assert ch[0][0] == 0 and ch[-1][1] == nb
    assert all(ch[i][1] == ch[i + 1][0] for i in range(len(ch) - 1))
    assert all(hi % G == 0 or hi == nb for _, hi in ch)
    sizes = [hi - lo for lo, hi in ch]
    assert all(sizes[i] >= sizes[i + 1] or sizes[i + 1] <= max(wmin, G) for i in range(len(sizes) - 1))


def test_recursive_scheme_applies_to_leaf_times_power_of_two(lib):
    """Padded sizes leaf * 2^k (k = 1..3) with the leaf a multiple of 128 rows between half and 5/4 of the nominal leaf."""
    assert lib.mogp_set_rchol(1, 4096, 2048) == 0
    want = {2048: 0, 4096: 2048, 4224: 0, 4352: 2176, 5120: 2560, 5248: 0, 6144: 1536, 7168: 1792, 8192: 2048, 10240: 2560,
            16384: 2048, 20480: 2560, 32768: 0}
    for n, leaf in want.items():
        assert lib.mogp_rchol_leaf_for(n) == leaf, (n, lib.mogp_rchol_leaf_for(n), leaf)
        assert int(lib.mogp_rchol_applies(n)) == int(leaf > 0)
        if leaf:
            k = (n // leaf).bit_length() - 1
            assert leaf << k == n and 1 <= k <= 3 and leaf % 128 == 0
    assert lib.mogp_set_rchol(1, 4096, 1024) == 0
    assert [int(lib.mogp_rchol_leaf_for(n)) for n in (2048, 4096, 8192, 16384)] == [0, 1024, 1024, 0]      # at most 8 leaves
    assert lib.mogp_set_rchol(0, 4096, 2048) == 0 and lib.mogp_rchol_applies(8192) == 0
    assert lib.mogp_set_rchol(1, 4096, 1000) == -1
    assert lib.mogp_set_rchol(1, 4096, 2048) == 0


def test_mohsm_parameter_count_and_kdiag_needs_inputs(lib):
    Cn, Q, D = 3, 2, 2
    assert lib.mogp_num_params(_cabi.KIND["MOHSM"], Cn, Q, D) == Q * (3 * Cn + 3 * Cn * D + D)


def test_padding_policy_for_the_recursive_scheme(lib):
    """A few more padding rows (<= ~3 %) when that makes the padded size leaf * 2^k; never for small problems, never when the
    128-row padding already qualifies, and the result is always a multiple of 128 that is >= N."""
    assert lib.mogp_set_rchol(1, 4096, 2048) == 0
    want = {2048: 2048, 3000: 3072, 4096: 4096, 4224: 4352, 5000: 5120, 5130: 5248, 6000: 6144, 7500: 7680, 8000: 8192, 8192: 8192}
    for n, np_ in want.items():
        got = lib.mogp_padded_size(n)
        assert got == np_, (n, got, np_)
        assert got % 128 == 0 and got >= n
    for n in range(4097, 9000, 37):
        got = lib.mogp_padded_size(n)
        base = (n + 127) // 128 * 128
        assert got % 128 == 0 and base <= got and (got - n) * 33 <= n or got == base
        if got != base:
            assert lib.mogp_rchol_leaf_for(got) > 0 and lib.mogp_rchol_leaf_for(base) == 0
    assert lib.mogp_set_pad_for_rchol(0) == 0
    assert lib.mogp_padded_size(6000) == 6016
    assert lib.mogp_set_pad_for_rchol(1) == 0
