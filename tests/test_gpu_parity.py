"""GPU parity tests: the CUDA path (through the C ABI) against the golden vectors generated
from the live reference, against the oracle on fresh random cases, and through
size-independent properties at the BASELINE.json sizes.

Tolerances (BASELINE.json north_star): log-marginal likelihood rtol 1e-8, posterior mean /
variance rtol 1e-6; additionally K max-abs <= 1e-12 * max|K| and gradients <= 1e-6 of the
largest entry of each parameter tensor."""
import os

import numpy as np
import pytest
import torch

from conftest import golden_names, load_golden, next_golden_names

pytestmark = pytest.mark.gpu

ALL = golden_names()
SMALLISH = [n for n in ALL if n != "cfg3"]


def rel(a, b):
    a = a.detach().cpu().numpy() if isinstance(a, torch.Tensor) else np.asarray(a, float)
    b = b.detach().cpu().numpy() if isinstance(b, torch.Tensor) else np.asarray(b, float)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


# ------------------------------------------------------------------ dense building blocks
@pytest.mark.parametrize("ta,tb", [(0, 0), (0, 1), (1, 0), (1, 1)])
@pytest.mark.parametrize("M,N,K", [(64, 64, 16), (128, 64, 64), (192, 128, 256), (448, 320, 96)])
def test_dgemm_against_torch_fp64(engine, ta, tb, M, N, K):
    g = torch.Generator(device="cpu").manual_seed(M * 7 + N * 3 + K + ta * 2 + tb)
    A = torch.randn((K, M) if ta else (M, K), generator=g, dtype=torch.float64).cuda()
    B = torch.randn((N, K) if tb else (K, N), generator=g, dtype=torch.float64).cuda()
    C0 = torch.randn((M, N), generator=g, dtype=torch.float64).cuda()
    ref = 0.7 * (A.T if ta else A) @ (B.T if tb else B) - 1.3 * C0
    out = engine.dgemm(ta, tb, 0.7, A, B, -1.3, C0.clone())
    torch.cuda.synchronize()
    assert rel(out, ref) < 1e-13


@pytest.mark.parametrize("n", [64, 128, 200, 777, 1024, 2048])
def test_potrf_against_lapack(engine, n):
    g = torch.Generator().manual_seed(n)
    B = torch.randn((n, n + 8), generator=g, dtype=torch.float64)
    A = B @ B.T / n + 0.5 * torch.eye(n, dtype=torch.float64)
    Lref = torch.linalg.cholesky(A)
    Ad = A.cuda().clone()
    info = engine.potrf_(Ad)
    assert info == 0
    L = torch.tril(Ad).cpu()
    assert rel(L, Lref) < 1e-11
    assert rel(L @ L.T, A) < 1e-13


@pytest.mark.parametrize("bad", [150, 151, 7, 64, 0, 299])
def test_potrf_reports_first_bad_pivot(engine, bad):
    """LAPACK convention: info = k means the leading minor of order k is not positive definite.  Even and odd
    columns take different branches of the 2x2 pivot step; 64 is the first column of the second panel."""
    n = 300
    g = torch.Generator().manual_seed(5)
    B = torch.randn((n, n), generator=g, dtype=torch.float64)
    A = B @ B.T / n + 0.5 * torch.eye(n, dtype=torch.float64)
    A[bad, bad] = -1.0                      # leading minor bad + 1 is not positive definite
    info = engine.potrf_(A.cuda().clone())
    assert info == bad + 1


@pytest.mark.parametrize("n", [128, 384, 1024])
def test_trtri_kinv(engine, n):
    g = torch.Generator().manual_seed(n + 1)
    B = torch.randn((n, n + 8), generator=g, dtype=torch.float64)
    A = B @ B.T / n + 0.3 * torch.eye(n, dtype=torch.float64)
    Ad = A.cuda().clone()
    Linv, Kinv, info = engine.trtri_kinv_(Ad)
    assert info == 0
    Lref = torch.linalg.cholesky(A)
    Linv_ref = torch.linalg.inv(Lref)
    assert rel(torch.tril(Linv).cpu(), Linv_ref) < 1e-10
    Kinv_ref = torch.linalg.inv(A)
    assert rel(torch.tril(Kinv).cpu(), torch.tril(Kinv_ref)) < 1e-10


# ------------------------------------------------------------------ kernel matrices
@pytest.mark.parametrize("name", ALL)
def test_K_matches_reference(engine, name):
    g = load_golden(name)
    K = engine.K(g["kind"], g["params"], g["X"]).cpu().numpy()
    assert np.array_equal(K, K.T)
    if "K_full" in g:
        assert rel(K, g["K_full"]) < 1e-12
    else:
        idx = g["K_idx"]
        scale = np.abs(g["K_val"]).max()
        assert np.abs(K[idx[:, 0], idx[:, 1]] - g["K_val"]).max() <= 1e-12 * scale
        assert np.abs(K[:, 0] - g["K_rowsum0"]).max() <= 1e-12 * scale
        assert abs(np.sqrt((K ** 2).sum()) - float(g["K_fro"])) <= 1e-11 * float(g["K_fro"])
    # reference unit test (tests/unit/test_kernels.py:43-57): K_diag == diag(K), bit-exact
    kd = engine.K_diag(g["kind"], g["params"], g["X"]).cpu().numpy()
    if not (g["kind"] == "SM" and g["D"] > 1):      # the reference's own SM K_diag ignores D
        assert np.array_equal(kd, np.diagonal(K))
    assert rel(kd, g["K_diag"]) < 1e-14


@pytest.mark.parametrize("name", [n for n in SMALLISH if n not in ("cfg4",)])
def test_cross_K_matches_reference(engine, name):
    g = load_golden(name)
    Kfs = engine.K(g["kind"], g["params"], g["X"], g["Xs"]).cpu().numpy()
    stride = int(g["Kfs_row_stride"])
    ref = g["Kfs_rows"]
    scale = max(np.abs(ref).max(), 1e-300)
    assert np.abs(Kfs[::stride] - ref).max() <= 1e-12 * max(scale, np.abs(g["K_diag"]).max())


def test_gram_with_noise_and_jitter_diagonal(engine):
    g = load_golden("mosm_datavar")
    from oracle import mogp_oracle as orc
    X = torch.tensor(g["X"])
    ref = orc._noisy_gram(g["kind"], g["params"], g["sigma_t"], X, g["jitter"], g["data_var"]).numpy()
    K = engine.K(g["kind"], g["params"], g["X"], sigma=g["sigma"], data_var=g["data_var"], jitter=g["jitter"])
    assert rel(K, ref) < 1e-13


# ------------------------------------------------------------------ LML + gradient
@pytest.mark.parametrize("name", ALL)
def test_lml_and_gradient_match_reference(engine, name):
    g = load_golden(name)
    res = engine.lml_grad(g["kind"], g["params"], g["sigma"], g["X"], g["y"], g["jitter"], True,
                          data_var=g.get("data_var"))
    assert res["info"] == 0
    lml_ref = float(g["lml"])
    assert abs(res["lml"] - lml_ref) <= 1e-8 * abs(lml_ref), (res["lml"], lml_ref)
    for k, got in res["grad"].items():
        ref = g["gc_" + k]
        scale = max(np.abs(ref).max(), 1e-12)
        err = np.abs(got.numpy().reshape(ref.shape) - ref).max()
        assert err <= 1e-6 * scale, (k, err / scale)


@pytest.mark.parametrize("name", ["mosm_mid", "cfg2"])
def test_lml_only_and_host_entry_agree(engine, name):
    from mogptk_b200.engine import pack_params
    g = load_golden(name)
    full = engine.lml_grad(g["kind"], g["params"], g["sigma"], g["X"], g["y"], g["jitter"], True)
    only = engine.lml_grad(g["kind"], g["params"], g["sigma"], g["X"], g["y"], g["jitter"], False)
    assert full["lml"] == only["lml"]
    rows = engine.prepare(g["kind"], g["params"], g["X"], g["y"])
    p = pack_params(g["kind"], g["params"]).numpy().copy()
    yh = rows.y.cpu().numpy().copy()
    out = engine.lml_grad_host(g["kind"], rows.dims, p, rows.x_host, rows.chan_off, yh, g["sigma"].copy(), g["jitter"])
    assert out[0] == full["lml"]
    flat = np.concatenate([full["grad"][k].numpy().ravel() for k in
                           ("weight", "mean", "variance", "delay", "phase", "sigma")])
    assert np.array_equal(out[2:], flat)


def test_not_positive_definite_is_reported(engine):
    from mogptk_b200.engine import NotPositiveDefiniteError
    g = load_golden("mosm_small")
    y = g["y"].copy()
    p = {k: v.clone() for k, v in g["params"].items()}
    p["weight"][:] = float("nan")
    with pytest.raises(NotPositiveDefiniteError):
        engine.lml_grad(g["kind"], p, g["sigma"], g["X"], y, g["jitter"], True)


# ------------------------------------------------------------------ prediction
@pytest.mark.parametrize("name", ALL)          # incl. cfg3 (N = 8192): predictions appended by make_golden --add-pred
def test_predict_f_matches_reference(engine, name):
    g = load_golden(name)
    engine.lml_grad(g["kind"], g["params"], g["sigma"], g["X"], g["y"], g["jitter"], False, data_var=g.get("data_var"))
    mu, var = engine.predict(g["Xs"])
    assert rel(mu, g["pred_mu"]) < 1e-6
    assert np.abs(var.cpu().numpy() - g["pred_var"]).max() <= 1e-6 * np.abs(g["pred_var"]).max()
    if "pred_cov" in g:
        mu2, cov = engine.predict(g["Xs"], full=True)
        assert rel(cov, g["pred_cov"]) < 1e-6
        assert rel(mu2, g["pred_mu"]) < 1e-6


# ------------------------------------------------------------------ fresh random cases vs the oracle
@pytest.mark.parametrize("kind,C,ns,Q,D,seed", [
    ("MOSM", 2, [130, 65], 4, 1, 21), ("MOSM", 5, [31, 64, 1, 129, 77], 9, 1, 22), ("MOSM", 3, [40, 50, 60], 2, 3, 23),
    ("SM", 3, [100, 3, 64], 4, 1, 24), ("CONV", 4, [64, 64, 64, 64], 3, 1, 25), ("CONV", 2, [90, 45], 1, 2, 26),
    ("MOSM", 3, [0, 80, 50], 2, 1, 27),
    ("MOSM", 1, [1], 1, 1, 28), ("SM", 2, [1, 2], 2, 1, 29), ("CONV", 2, [3, 2], 1, 5, 30), ("MOSM", 2, [70, 60], 2, 8, 31),
])
def test_random_cases_against_oracle(engine, kind, C, ns, Q, D, seed):
    from mogptk_b200 import synth
    from oracle import mogp_oracle as orc
    X, y = synth.make_data(C, ns, seed, D)
    p, _ = synth.make_params(kind, C, Q, D, seed, random_delay_phase=True)
    sigma = torch.tensor(np.random.default_rng(seed).uniform(0.1, 0.6, C))
    Xt = torch.tensor(X)
    assert rel(engine.K(kind, p, X), orc.K(kind, p, Xt)) < 1e-12
    res = engine.lml_grad(kind, p, sigma, X, y, 1e-8, True)
    loss, gref = orc.loss_and_grad(kind, p, sigma, Xt, y, 1e-8)
    assert abs(res["lml"] + float(loss)) <= 1e-8 * abs(float(loss))
    gmax = max(float(v.abs().max()) for v in gref.values())
    for k, got in res["grad"].items():
        # tensors whose gradient is pure round-off next to the others (e.g. delays between channels that are
        # uncorrelated in 8 input dimensions) are compared on the scale of the largest gradient
        scale = max(float(gref[k].abs().max()), 1e-7 * gmax, 1e-12)
        assert float((got.reshape(gref[k].shape) - gref[k]).abs().max()) <= 1e-6 * scale, k
    rng = np.random.default_rng(seed)
    Xs = np.concatenate([rng.integers(0, C, 37).astype(float)[:, None], rng.uniform(0, 10, (37, D))], axis=1)
    mu, var = engine.predict(Xs)
    mu_r, var_r = orc.predict_f(kind, p, sigma, Xt, y, Xs, 1e-8)
    assert rel(mu, mu_r.ravel()) < 1e-6
    assert float((var.cpu() - var_r.ravel()).abs().max()) <= 1e-6 * float(var_r.abs().max())


# ------------------------------------------------------------------ properties at full size
def test_properties_at_cfg2_size(engine):
    g = load_golden("cfg2")
    a = engine.lml_grad(g["kind"], g["params"], g["sigma"], g["X"], g["y"], g["jitter"], True)
    b = engine.lml_grad(g["kind"], g["params"], g["sigma"], g["X"], g["y"], g["jitter"], True)
    assert a["lml"] == b["lml"]                                         # deterministic
    for k in a["grad"]:
        assert torch.equal(a["grad"][k], b["grad"][k])
    perm = np.random.default_rng(0).permutation(g["X"].shape[0])       # row order does not matter
    c = engine.lml_grad(g["kind"], g["params"], g["sigma"], g["X"][perm], g["y"][perm], g["jitter"], True)
    assert abs(c["lml"] - a["lml"]) <= 1e-12 * abs(a["lml"])
    # directional finite difference of the LML along the analytic gradient (weights only)
    p2 = {k: v.clone() for k, v in g["params"].items()}
    d = a["grad"]["weight"] / a["grad"]["weight"].norm()
    eps = 1e-5
    p2["weight"] = g["params"]["weight"] + eps * d
    up = engine.lml_grad(g["kind"], p2, g["sigma"], g["X"], g["y"], g["jitter"], False)["lml"]
    p2["weight"] = g["params"]["weight"] - eps * d
    dn = engine.lml_grad(g["kind"], p2, g["sigma"], g["X"], g["y"], g["jitter"], False)["lml"]
    fd = -(up - dn) / (2 * eps)                                          # d(-LML)/d direction
    an = float((a["grad"]["weight"] * d).sum())
    assert abs(fd - an) <= 1e-5 * abs(an)


# ------------------------------------------------------------------ CUDA-graph replay of the step
def test_repeated_evaluations_replay_the_graph_and_survive_reallocation(engine):
    """The first evaluation of a configuration runs plain, the second is captured into a CUDA graph, later ones
    replay it; a larger problem in between reallocates workspace buffers and must invalidate the capture."""
    small = load_golden("mosm_datavar")          # exercises the data-variance pointer inside the graph
    mid = load_golden("mosm_mid")
    big = load_golden("cfg2")

    def ev(g, grad=True):
        return engine.lml_grad(g["kind"], g["params"], g["sigma"], g["X"], g["y"], g["jitter"], grad,
                               data_var=g.get("data_var"))

    first = {n: ev(g) for n, g in (("small", small), ("mid", mid))}
    for rep in range(3):                          # capture on rep 0, replay afterwards
        for n, g in (("small", small), ("mid", mid)):
            r = ev(g)
            assert r["lml"] == first[n]["lml"] and r["info"] == 0
            for k in r["grad"]:
                assert torch.equal(r["grad"][k], first[n]["grad"][k])
    ref_big = ev(big)                             # bigger problem: workspace buffers move
    assert abs(ref_big["lml"] - float(big["lml"])) <= 1e-8 * abs(float(big["lml"]))
    for n, g in (("small", small), ("mid", mid)):
        for rep in range(2):
            r = ev(g)
            assert r["lml"] == first[n]["lml"]
            for k in r["grad"]:
                assert torch.equal(r["grad"][k], first[n]["grad"][k])
    mu, var = engine.predict(mid["Xs"])           # the factor of a replayed step serves predictions
    assert rel(mu, mid["pred_mu"]) < 1e-6
    p2 = {k: v.clone() for k, v in mid["params"].items()}
    p2["weight"] = p2["weight"] * 1.01            # new parameter values through the same captured graph
    a = engine.lml_grad(mid["kind"], p2, mid["sigma"], mid["X"], mid["y"], mid["jitter"], True)
    from oracle import mogp_oracle as orc
    ref = float(orc.lml(mid["kind"], p2, mid["sigma_t"], torch.tensor(mid["X"]), mid["y"], mid["jitter"]))
    assert abs(a["lml"] - ref) <= 1e-8 * abs(ref)


def test_prediction_beyond_the_workspace_size_is_chunked():
    """More test points than the engine's max_n: predict() works through slices (diagonal variances only)."""
    from mogptk_b200.engine import Engine
    from oracle import mogp_oracle as orc
    g = load_golden("mosm_small")
    eng = Engine(device=0, max_n=64)
    try:
        eng.lml_grad(g["kind"], g["params"], g["sigma"], g["X"], g["y"], g["jitter"], False)
        rng = np.random.default_rng(3)
        Xs = np.concatenate([rng.integers(0, g["C"], 150).astype(float)[:, None], rng.uniform(0, 10, (150, 1))], axis=1)
        mu, var = eng.predict(Xs)
        mu_r, var_r = orc.predict_f(g["kind"], g["params"], g["sigma_t"], torch.tensor(g["X"]), g["y"], Xs, g["jitter"])
        assert rel(mu, mu_r.ravel()) < 1e-6 and rel(var, var_r.ravel()) < 1e-6
        with pytest.raises(ValueError):
            eng.predict(Xs, full=True)
    finally:
        eng.close()


def test_tile_cache_eviction_does_not_break_the_captured_training_graph():
    """VERDICT r1 / ADVICE: 65+ distinct test-set layouts through mogp_predict used to evict (and free) the training
    tile list that a captured step graph still referenced.  Cycle 70 layouts between training steps: every
    evaluation must keep returning the golden LML, gradient and predictions."""
    from mogptk_b200.engine import Engine
    g = load_golden("mosm_mid")
    eng = Engine(device=0, max_n=512)
    try:
        rows = eng.prepare(g["kind"], g["params"], g["X"], g["y"])
        from mogptk_b200.engine import pack_params
        packed = pack_params(g["kind"], g["params"], eng.device)
        sig = g["sigma_t"].to(eng.device)
        C_ = g["C"]
        rng = np.random.default_rng(0)
        ref = None
        for it in range(70):
            out = eng.lml_grad_prepared(rows, packed, sig, g["jitter"], True)       # replayed graph after the 2nd call
            cur = out.cpu().numpy()
            assert abs(cur[0] - g["lml"]) <= 1e-8 * abs(g["lml"])
            if ref is None:
                ref = cur
            assert np.array_equal(cur, ref), it                                      # bit-for-bit, every time
            n_s = 3 + it                                                             # a new test layout every time
            Xs = np.stack([rng.integers(0, C_, n_s).astype(np.float64), rng.uniform(0, 10, n_s)], axis=1)
            mu, var = eng.predict(Xs, full=(it % 7 == 0))
            assert torch.isfinite(mu).all() and torch.isfinite(var).all()
        mu, var = eng.predict(g["Xs"])
        assert rel(mu, g["pred_mu"]) < 1e-6
    finally:
        eng.close()


# ------------------------------------------------------------------ further kernel families
def _next_case(name):
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", name + ".npz"), allow_pickle=False)
    g = {k: z[k] for k in z.files}
    fam = str(g["kind"])
    kind = fam if fam in ("UMOSM", "MOHSM") else "%s:%d" % (fam, int(g["Rq"]))
    p = {k[2:]: torch.tensor(v, dtype=torch.float64) for k, v in g.items() if k.startswith("p_")}
    return g, kind, p


@pytest.mark.parametrize("name", next_golden_names())
def test_further_kernel_families_match_the_reference(engine, name):
    """CSM (gpr/multioutput.py:397-454), SM-LMC (:456-502), uMOSM (:212-293) and MOHSM (:295-395, non-stationary: mid-point
    window, row-dependent Gram diagonal) on the same tile kernels: K, K_diag, LML (rtol 1e-8), constrained-space gradients
    (1e-6) and predictions (1e-6) against the live-reference fixtures."""
    g, kind, p = _next_case(name)
    X, y, sigma, jitter = g["X"], g["y"], torch.tensor(g["sigma"]), float(g["jitter"])
    K = engine.K(kind, p, X)
    assert rel(K, g["K"]) < 1e-12
    assert torch.equal(K, K.T)
    if not (kind.startswith("SMLMC") and int(g["D"]) > 1):
        assert torch.equal(engine.K_diag(kind, p, X), K.diagonal())
    # (for D > 1 the reference's own SM-LMC K_diag, sum_q w^2 magnitude_q, lacks the factor D its K has on the diagonal --
    #  SpectralKernel.K sums over the input dimensions, gpr/singleoutput.py:550-561 -- and the engine mirrors both)
    assert rel(engine.K_diag(kind, p, X), g["K_diag"]) < 1e-12
    Kx = engine.K(kind, p, X, g["Xs"])
    from oracle import next_kernels as nk
    orc = nk.register()
    assert rel(Kx, orc.K(kind.partition(":")[0], p, X, g["Xs"])) < 1e-12
    res = engine.lml_grad(kind, p, sigma, X, y, jitter, True)
    assert abs(res["lml"] - float(g["lml"])) <= 1e-8 * abs(float(g["lml"]))
    for k, got in res["grad"].items():
        ref = g["gc_" + k]
        assert np.abs(got.cpu().numpy().reshape(ref.shape) - ref).max() <= 1e-6 * max(np.abs(ref).max(), 1e-12), k
    mu, var = engine.predict(g["Xs"])
    assert rel(mu, g["pred_mu"]) < 1e-6
    assert np.abs(var.cpu().numpy() - g["pred_var"]).max() <= 1e-6 * np.abs(g["pred_var"]).max()


@pytest.mark.parametrize("C_,Q,D,ns,seed", [(3, 2, 1, [150, 97, 131], 0), (2, 3, 2, [90, 140], 1), (1, 2, 1, [200], 2)])
def test_mohsm_multi_tile_case_against_the_oracle(engine, C_, Q, D, ns, seed):
    """MOHSM beyond one tile per channel pair (ragged tiles, several components, D = 2, a single channel): K, K_diag (bit-equal
    to diag K), cross-covariance, LML, every gradient (incl. lengthscale and center, and the relative-jitter term of the
    row-dependent diagonal: jitter 1e-3 makes it visible) and predictions against the oracle restatement."""
    from oracle import next_kernels as nk
    orc = nk.register()
    rng = np.random.default_rng(100 + seed)
    gen = torch.Generator().manual_seed(seed)
    r = lambda *s: torch.rand(*s, generator=gen, dtype=torch.float64)
    p = {"weight": r(Q, C_) + 0.3, "mean": r(Q, C_, D) * 0.8 + 0.05, "variance": r(Q, C_, D) * 0.5 + 0.1,
         "lengthscale": r(Q, C_) * 0.3 + 0.15, "center": r(Q, D) * 3.0 + 1.0,
         "delay": 0.2 * (r(Q, C_, D) - 0.5), "phase": 0.6 * (r(Q, C_) - 0.5)}
    X = np.concatenate([np.concatenate([np.full((n, 1), float(c)), np.sort(rng.uniform(0, 5, (n, D)), axis=0)], axis=1)
                        for c, n in enumerate(ns)])
    y = rng.standard_normal(X.shape[0])
    # (the reference's cross-channel MOHSM formula is not positive semi-definite for arbitrary parameters -- smallest eigenvalue
    #  -0.54 for the first case -- so the noise is chosen large enough for the Gram matrix to be factorisable)
    sigma = torch.tensor(0.9 + 0.4 * rng.uniform(size=C_))
    jitter = 1e-3
    K = engine.K("MOHSM", p, X)
    Kref = orc.K("MOHSM", p, torch.tensor(X))
    assert rel(K, Kref) < 1e-12
    assert torch.equal(K, K.T)
    assert torch.equal(engine.K_diag("MOHSM", p, X), K.diagonal())
    assert rel(engine.K_diag("MOHSM", p, X), orc.K_diag("MOHSM", p, torch.tensor(X))) < 1e-12
    Xs = np.concatenate([np.concatenate([np.full((7, 1), float(c)), rng.uniform(0, 5, (7, D))], axis=1) for c in range(C_)])
    assert rel(engine.K("MOHSM", p, X, Xs), orc.K("MOHSM", p, torch.tensor(X), torch.tensor(Xs))) < 1e-12
    for _ in range(3):                    # plain run, graph capture, replay
        res = engine.lml_grad("MOHSM", p, sigma, X, y, jitter, True)
    lml_ref = float(orc.lml("MOHSM", p, sigma, torch.tensor(X), torch.tensor(y), jitter))
    assert abs(res["lml"] - lml_ref) <= 1e-8 * abs(lml_ref)
    _, g_ref = orc.loss_and_grad("MOHSM", p, sigma, torch.tensor(X), torch.tensor(y), jitter)
    for k, got in res["grad"].items():
        ref = g_ref[k].numpy()
        assert np.abs(got.cpu().numpy().reshape(ref.shape) - ref).max() <= 1e-6 * max(np.abs(ref).max(), 1e-12), k
    mu, var = engine.predict(Xs)
    mu_ref, var_ref = orc.predict_f("MOHSM", p, sigma, torch.tensor(X), torch.tensor(y), torch.tensor(Xs), jitter)
    assert rel(mu, mu_ref.reshape(-1)) < 1e-6
    assert np.abs(var.cpu().numpy() - var_ref.numpy().reshape(-1)).max() <= 1e-6 * float(var_ref.abs().max())
    _, cov = engine.predict(Xs, full=True)
    _, cov_ref = orc.predict_f("MOHSM", p, sigma, torch.tensor(X), torch.tensor(y), torch.tensor(Xs), jitter, full=True)
    assert rel(cov, cov_ref) < 1e-6
