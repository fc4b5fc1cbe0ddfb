"""GPU: the UNMODIFIED reference drives the CUDA engine through its own builder seam.

``mogptk.MOSM/SM/CONV(dataset, Q, inference=mogptk_b200.B200Exact())`` (mogptk/model.py:181,231) ->
``train('Adam')`` / ``train('LBFGS')`` (mogptk/model.py:546-566) -> ``predict()`` (:608-664) ->
``save()`` / ``LoadModel`` (:62-75,323-340), compared on the same box with the stock ``mogptk.Exact``
running the reference's own PyTorch-CPU path.  The reference package is the byte-for-byte copy in
oracle/_ref (oracle/build_ref.py); nothing here reads /root/reference.

Tolerances: losses rtol 1e-8, predictions rtol 1e-6 (BASELINE.json north_star).
"""
import contextlib
import os

import numpy as np
import pytest
import torch

from conftest import load_golden

pytestmark = pytest.mark.gpu


def test_reference_copy_is_present_and_unmodified():
    """oracle/_ref must have travelled to the GPU box and must still match the manifest of its source."""
    from oracle import ref_loader
    assert os.path.isdir(ref_loader.REF_COPY), "oracle/_ref is missing: run __graft_entry__.build() in the build container"
    assert ref_loader.verify_copy()


@pytest.fixture(scope="module")
def mogptk():
    from oracle import ref_loader
    if not ref_loader.available():
        pytest.skip("reference copy not available (see test_reference_copy_is_present_and_unmodified)")
    m = ref_loader.import_reference("cuda")
    yield m
    ref_loader.set_device(m, "cuda")


@contextlib.contextmanager
def on(mogptk, device):
    from oracle import ref_loader
    ref_loader.set_device(mogptk, device)
    try:
        yield
    finally:
        ref_loader.set_device(mogptk, "cuda")


def dataset(mogptk, X, y, C):
    ds = mogptk.DataSet()
    for c in range(C):
        msk = X[:, 0] == c
        ds.append(mogptk.Data(X[msk, 1:], y[msk], name=str(c)))
    return ds


def make_pair(mogptk, family, C, ns, Q, seed, D=1, **kw):
    """(stock reference model on the CPU, the same model through B200Exact on the GPU) with identical raw
    parameters (the constructors draw torch.rand on their own device, so the raw leaves are copied over)."""
    import mogptk_b200 as mb
    from mogptk_b200 import synth
    X, y = synth.make_data(C, ns, seed=seed, D=D)
    cls = getattr(mogptk, family)
    with on(mogptk, "cpu"):
        torch.manual_seed(seed)
        a = cls(dataset(mogptk, X, y, C), Q=Q, **kw)
        g = torch.Generator().manual_seed(seed + 1)
        for p in a.gpr.parameters():            # move off the constructor's dead-mean start
            if p._name.endswith(".mean"):
                p.assign(0.1 + 1.2 * torch.rand(p.shape, generator=g, dtype=torch.float64))
    b = cls(dataset(mogptk, X, y, C), Q=Q, inference=mb.B200Exact(), **kw)
    pa, pb = list(a.gpr.named_parameters()), list(b.gpr.named_parameters())
    assert [n for n, _ in pa] == [n for n, _ in pb]
    for (_, u), (_, v) in zip(pa, pb):
        assert v.is_cuda and type(v).__module__.startswith("mogptk.")      # the reference's own Parameter objects
        v.data = u.data.detach().clone().to(v.device)
    assert type(b.gpr).__module__ == "mogptk_b200.gpr"
    return a, b


def close(u, v, rtol):
    u, v = np.asarray(u, dtype=np.float64), np.asarray(v, dtype=np.float64)
    return np.abs(u - v).max() <= rtol * max(np.abs(u).max(), 1e-12)


@pytest.mark.parametrize("family,C,ns,Q", [("MOSM", 3, [60, 41, 80], 2), ("SM", 2, [70, 50], 3), ("CONV", 3, [50, 64, 40], 2)])
def test_adam_training_and_prediction_through_the_reference(mogptk, family, C, ns, Q):
    a, b = make_pair(mogptk, family, C, ns, Q, seed=3)
    with on(mogptk, "cpu"):
        la, _ = a.train(method="Adam", iters=25, lr=0.05, verbose=False, jit=False)
        lml_a = a.log_marginal_likelihood()
        _, Ma, La, Ua = a.predict()
    lb, _ = b.train(method="Adam", iters=25, lr=0.05, verbose=False, jit=False)
    assert len(la) == len(lb) == 26
    assert close(la, lb, 1e-8), np.abs(la - lb).max()
    assert abs(lml_a - b.log_marginal_likelihood()) <= 1e-8 * abs(lml_a)
    _, Mb, Lb, Ub = b.predict()
    for u, v in zip(Ma + La + Ua, Mb + Lb + Ub):
        assert close(u, v, 1e-6)
    assert a.num_parameters() == b.num_parameters()
    # training moved the reference's own Parameter objects (the optimiser is the reference's torch.optim.Adam)
    for (_, u), (_, v) in zip(a.gpr.named_parameters(), b.gpr.named_parameters()):
        assert close(u.detach().numpy(), v.detach().cpu().numpy(), 1e-6)


def test_lbfgs_closure_resume_and_save_load(mogptk, tmp_path):
    a, b = make_pair(mogptk, "MOSM", 2, [45, 38], 2, seed=7)
    with on(mogptk, "cpu"):
        la, _ = a.train(method="LBFGS", iters=8, verbose=False, jit=False)
    lb, _ = b.train(method="LBFGS", iters=8, verbose=False, jit=False)
    assert len(la) == len(lb)
    assert close(la, lb, 1e-7), np.abs(la - lb).max()
    with on(mogptk, "cpu"):
        a.train(method="Adam", iters=3, lr=0.02, verbose=False, jit=True)
    b.train(method="Adam", iters=3, lr=0.02, verbose=False, jit=True)       # resumes; compile() is a no-op for the engine
    assert len(b.losses) == len(lb) + 3 and close(a.losses, b.losses, 1e-7)
    path = str(tmp_path / "model")
    b.save(path)
    c = mogptk.LoadModel(path)
    _, Mb, Lb, Ub = b.predict()
    _, Mc, Lc, Uc = c.predict()
    for u, v in zip(Mb + Lb + Ub, Mc + Lc + Uc):
        assert close(u, v, 1e-12)
    assert abs(b.log_marginal_likelihood() - c.log_marginal_likelihood()) <= 1e-12 * abs(b.log_marginal_likelihood())


def reference_kernel(mogptk, g):
    """The reference's own kernel objects (as oracle/make_golden.py builds them) on the GPU, raw values from the fixture."""
    gp = mogptk.gpr
    kind, C, Q, D = g["kind"], g["C"], g["Q"], g["D"]
    if kind == "MOSM":
        k = gp.MultiOutputSpectralMixtureKernel(Q=Q, output_dims=C, input_dims=D)
        plist = {n: [getattr(k, n)] for n in ("weight", "mean", "variance", "delay", "phase")}
    elif kind == "SM":
        k = gp.IndependentMultiOutputKernel([gp.SpectralMixtureKernel(Q=Q, input_dims=D) for _ in range(C)], output_dims=C)
        plist = {n: [getattr(k[c], n) for c in range(C)] for n in ("magnitude", "mean", "variance")}
    else:
        k = gp.MixtureKernel(gp.GaussianConvolutionProcessKernel(output_dims=C, input_dims=D), Q)
        plist = {n: [getattr(k[q], n) for q in range(Q)] for n in ("weight", "variance", "base_variance")}
    return k, plist


@pytest.mark.parametrize("name", ["mosm_small", "mosm_shuffled", "mosm_datavar", "mosm_small_d2", "sm_small", "sm_small_d2",
                                  "conv_small", "conv_small_d2", "cfg1", "cfg2", "cfg2_rdp", "cfg4", "cfg3"])
def test_reference_objects_through_the_builder_seam_match_the_golden_vectors(mogptk, name):
    """B200Exact._build(kernel, x, y, y_err) with the reference's Parameter / kernel / transform objects (duck-typed
    by kernel_spec, transforms matched in _fast_table) against the live-reference fixtures, up to N = 8192."""
    import mogptk_b200 as mb
    g = load_golden(name)
    k, plist = reference_kernel(mogptk, g)
    y_err = np.sqrt(g["data_var"]) if "data_var" in g else None
    m = mb.B200Exact(variance=(g["sigma"] ** 2).tolist(), jitter=g["jitter"])._build(k, g["X"], g["y"].reshape(-1, 1), y_err)
    assert type(m.likelihood).__module__.startswith("mogptk.")
    plist["sigma"] = [m.likelihood.scale]
    for n, lst in plist.items():
        raw = torch.tensor(g["r_" + n], dtype=torch.float64)
        for i, prm in enumerate(lst):
            prm.data = (raw if len(lst) == 1 else raw[i]).clone().reshape(prm.shape).to(prm.device)
    loss = m.loss()
    assert m._fast_table() is not None                 # the device-resident path took the reference's transforms
    assert abs(float(loss) - float(g["loss"])) <= 1e-8 * abs(float(g["loss"]))
    for n, lst in plist.items():
        ref = torch.tensor(g["gr_" + n])
        got = (lst[0].grad if len(lst) == 1 else torch.stack([p.grad for p in lst])).cpu()
        assert float((got.reshape(ref.shape) - ref).abs().max()) <= 1e-6 * max(float(ref.abs().max()), 1e-12), n
    mu, var = m.predict_f(g["Xs"])
    assert close(g["pred_mu"], mu.cpu().numpy().ravel(), 1e-6)
    assert close(g["pred_var"], var.cpu().numpy().ravel(), 1e-6)


def test_cfg2_size_training_against_the_reference_cuda_path(mogptk):
    """BASELINE configs[1] (MOSM 4 x 512, Q = 5) through mogptk.MOSM(...).train(): the plug-in against the stock
    reference running its own PyTorch-CUDA path (gpr/config.py:51-62) on the same GPU."""
    import mogptk_b200 as mb
    from mogptk_b200 import synth
    X, y = synth.make_data(4, 512, seed=0)
    torch.manual_seed(0)
    a = mogptk.MOSM(dataset(mogptk, X, y, 4), Q=5)
    a.gpr.kernel.mean.assign(torch.rand(4, 5, 1) * 2.0 + 0.05)
    b = mogptk.MOSM(dataset(mogptk, X, y, 4), Q=5, inference=mb.B200Exact())
    for (_, u), (_, v) in zip(a.gpr.named_parameters(), b.gpr.named_parameters()):
        v.data = u.data.detach().clone()
    la, _ = a.train(method="Adam", iters=10, lr=0.02, verbose=False, jit=False)
    lb, _ = b.train(method="Adam", iters=10, lr=0.02, verbose=False, jit=False)
    assert close(la, lb, 1e-8), np.abs(la - lb).max()
    _, Ma, _, _ = a.predict()
    _, Mb, _, _ = b.predict()
    for u, v in zip(Ma, Mb):
        assert close(u, v, 1e-6)


def test_cholesky_failure_raises_the_reference_exception(mogptk):
    import mogptk_b200 as mb
    g = load_golden("mosm_small")
    k, _ = reference_kernel(mogptk, g)
    m = mb.B200Exact(jitter=g["jitter"])._build(k, g["X"], g["y"].reshape(-1, 1))
    k.weight.data.fill_(float("nan"))
    with pytest.raises(mogptk.gpr.CholeskyException):
        m.loss()


def test_trainable_mean_function_trains_like_the_reference(mogptk):
    """mean=LinearMean: the reference back-propagates through y - mean(X) (gpr/model.py:445-452); the plug-in
    chains d LML / d y = -alpha (mogp_alpha) into the same parameters."""
    import mogptk_b200 as mb
    from mogptk_b200 import synth
    X, y = synth.make_data(2, [50, 42], seed=12)
    y = y + 0.4 * X[:, 1] - 1.0

    def model(**kw):
        torch.manual_seed(5)
        m = mogptk.MOSM(dataset(mogptk, X, y, 2), Q=2, mean=mogptk.gpr.LinearMean(input_dims=2), **kw)
        return m

    with on(mogptk, "cpu"):
        a = model()
        a.gpr.kernel.mean.assign(torch.full((2, 2, 1), 0.45, dtype=torch.float64))
    b = model(inference=mb.B200Exact())
    for (na, u), (nb, v) in zip(a.gpr.named_parameters(), b.gpr.named_parameters()):
        assert na == nb
        v.data = u.data.detach().clone().to(v.device)
    assert any("LinearMean" in n or "mean." in n for n, _ in b.gpr.named_parameters())
    with on(mogptk, "cpu"):
        la, _ = a.train(method="Adam", iters=15, lr=0.05, verbose=False, jit=False)
        _, Ma, _, _ = a.predict()
    lb, _ = b.train(method="Adam", iters=15, lr=0.05, verbose=False, jit=False)
    assert close(la, lb, 1e-8), np.abs(la - lb).max()
    assert abs(float(b.gpr.mean.bias())) > 1e-3                   # the mean really moved
    assert close(a.gpr.mean.slope().detach().numpy(), b.gpr.mean.slope().detach().cpu().numpy(), 1e-6)
    _, Mb, _, _ = b.predict()
    for u, v in zip(Ma, Mb):
        assert close(u, v, 1e-6)


def test_device_resident_adam_matches_the_reference_for_100_iterations(mogptk):
    """mogptk_b200.install() routes mogptk.Model.train('Adam') to mogp_train_adam (K iterations per synchronisation):
    the loss trajectory must equal the stock reference's (its own CPU path + torch.optim.Adam) to 1e-7 over 100
    iterations, and losses / times / iters must keep the reference's bookkeeping, including resume."""
    import mogptk_b200 as mb
    a, b = make_pair(mogptk, "MOSM", 3, [64, 50, 70], 2, seed=11)
    with on(mogptk, "cpu"):
        la, _ = a.train(method="Adam", iters=100, lr=0.02, verbose=False, jit=False)
        la2, _ = a.train(method="Adam", iters=5, lr=0.01, verbose=False, jit=False)
        _, Ma, _, _ = a.predict()
    mb.install(mogptk)
    try:
        calls = []
        orig = b.gpr.loss
        b.gpr.loss = lambda: (calls.append(1), orig())[1]
        lb, eb = b.train(method="Adam", iters=100, lr=0.02, verbose=False, jit=False, sync_every=32)
        assert len(calls) == 1                               # only the closing evaluation went through loss()
        assert len(lb) == 101 and len(eb) == 101 and b.iters == 100 and b.times.shape == (101,)
        assert np.all(np.diff(b.times) >= 0)
        assert close(la, lb, 1e-7), np.abs(la - lb).max()
        lb2, _ = b.train(method="Adam", iters=5, lr=0.01, verbose=False, jit=False)      # resume: fresh optimiser, appended history
        assert len(lb2) == 106 and b.iters == 105 and close(la2, lb2, 1e-7)
        _, Mb, _, _ = b.predict()
        for u, v in zip(Ma, Mb):
            assert close(u, v, 1e-6)
        # requests the device loop does not cover fall through to the reference's own train()
        b.train(method="LBFGS", iters=3, verbose=False, jit=False)
        b.train(method="Adam", iters=2, lr=0.01, amsgrad=True, verbose=False, jit=False)
        assert len(calls) > 3
    finally:
        mb.uninstall(mogptk)
    assert mogptk.Model.train.__name__ == "train"


def test_device_resident_adam_stops_at_a_cholesky_failure(mogptk):
    import mogptk_b200 as mb
    from mogptk_b200 import train as fused
    g = load_golden("mosm_small")
    k, _ = reference_kernel(mogptk, g)
    m = mb.B200Exact(jitter=g["jitter"])._build(k, g["X"], g["y"].reshape(-1, 1))
    losses, _ = fused.fit_adam(m, 6, lr=0.01, sync_every=4)
    assert np.isfinite(losses).all() and losses[-1] < losses[0]
    raw_before = k.variance.data.clone()
    k.weight.data[0, 0] = float("nan")
    with pytest.raises(mogptk.gpr.CholeskyException):
        fused.fit_adam(m, 8, lr=0.01, sync_every=8)
    assert torch.equal(k.variance.data, raw_before)          # frozen at the failing iteration


@pytest.mark.parametrize("family,kw", [("CSM", dict(Q=2, Rq=2)), ("SM_LMC", dict(Q=3, Rq=2)), ("MOHSM", dict(P=2, Q=2))])
def test_csm_and_sm_lmc_models_through_the_reference(mogptk, family, kw):
    """mogptk.CSM / mogptk.SM_LMC / mogptk.MOHSM (mogptk/models/csm.py, sm_lmc.py, mohsm.py) with inference=B200Exact() against
    the stock reference on the CPU: Adam training, predictions, and the device-resident loop after install()."""
    import mogptk_b200 as mb
    from mogptk_b200 import synth
    X, y = synth.make_data(3, [60, 44, 72], seed=13)
    cls = getattr(mogptk, family)
    with on(mogptk, "cpu"):
        torch.manual_seed(6)
        a = cls(dataset(mogptk, X, y, 3), **kw)
        g = torch.Generator().manual_seed(2)
        for n, p in a.gpr.named_parameters():
            if n.endswith(".mean"):
                p.assign(0.2 + torch.rand(p.shape, generator=g, dtype=torch.float64))
            elif n.endswith(".shift"):
                p.assign(0.3 * torch.randn(p.shape, generator=g, dtype=torch.float64))
            elif n.endswith(".center"):
                p.assign(2.0 + 6.0 * torch.rand(p.shape, generator=g, dtype=torch.float64))
    b = cls(dataset(mogptk, X, y, 3), inference=mb.B200Exact(), **kw)
    c = cls(dataset(mogptk, X, y, 3), inference=mb.B200Exact(), **kw)
    for (na, u), (nb, v), (nc, w) in zip(a.gpr.named_parameters(), b.gpr.named_parameters(), c.gpr.named_parameters()):
        assert na == nb == nc
        v.data = u.data.detach().clone().to(v.device)
        w.data = u.data.detach().clone().to(w.device)
    with on(mogptk, "cpu"):
        la, _ = a.train(method="Adam", iters=30, lr=0.03, verbose=False, jit=False)
        _, Ma, _, _ = a.predict()
    lb, _ = b.train(method="Adam", iters=30, lr=0.03, verbose=False, jit=False)
    assert close(la, lb, 1e-8), np.abs(la - lb).max()
    _, Mb, _, _ = b.predict()
    for u, v in zip(Ma, Mb):
        assert close(u, v, 1e-6)
    mb.install(mogptk)
    try:
        lc, _ = c.train(method="Adam", iters=30, lr=0.03, verbose=False, jit=False)
    finally:
        mb.uninstall(mogptk)
    assert close(la, lc, 1e-7), np.abs(la - lc).max()


# ------------------------------------------------------------------ initialisers that run exact GPs
def _signal(n, seed):
    rng = np.random.default_rng(seed)
    x = np.sort(rng.uniform(0.0, 12.0, n))
    y = np.sin(2 * np.pi * 0.45 * x) + 0.6 * np.sin(2 * np.pi * 1.2 * x + 0.4) + 0.1 * rng.standard_normal(n)
    return x, y


def test_bnse_on_the_engine_matches_the_reference(mogptk):
    """mogptk_b200.init.BNSE (training, Gram, factor and the N^2 n products through the C ABI) against the reference's
    mogptk.init.BNSE (mogptk/init.py:5-126) on the CPU: same grid, PSD mean / variance to 1e-5 / 1e-4 of their maxima."""
    import mogptk_b200 as mb
    x, y = _signal(150, 21)
    with on(mogptk, "cpu"):
        w0, m0, v0 = mogptk.init.BNSE(x.copy(), y.copy(), n=400, iters=30, jit=False)
    w1, m1, v1 = mb.init.BNSE(x.copy(), y.copy(), n=400, iters=30)
    assert np.array_equal(w0, w1) or np.abs(w0 - w1).max() <= 1e-14 * np.abs(w0).max()
    assert np.abs(m0 - m1).max() <= 1e-5 * np.abs(m0).max(), np.abs(m0 - m1).max() / np.abs(m0).max()
    assert np.abs(v0 - v1).max() <= 1e-4 * np.abs(v0).max()
    # with measurement errors (data variance in the training model only, init.py:40)
    ye = 0.05 + 0.05 * np.cos(x) ** 2
    with on(mogptk, "cpu"):
        _, m2, _ = mogptk.init.BNSE(x.copy(), y.copy(), y_err=ye, n=200, iters=15, jit=False)
    _, m3, _ = mb.init.BNSE(x.copy(), y.copy(), y_err=ye, n=200, iters=15)
    assert np.abs(m2 - m3).max() <= 1e-5 * np.abs(m2).max()


@pytest.mark.parametrize("method", ["BNSE", "SM"])
def test_init_parameters_through_the_engine(mogptk, method):
    """mogptk.MOSM(...).init_parameters(method) (mogptk/models/mosm.py:62-113) after mogptk_b200.install(): the per-channel
    exact GPs of BNSE / Data.get_sm_estimation run on the engine (channels concurrently for BNSE) and give the reference's
    starting values."""
    import mogptk_b200 as mb
    from mogptk_b200 import synth
    X, y = synth.make_data(3, [90, 70, 110], seed=17)
    with on(mogptk, "cpu"):
        torch.manual_seed(3)
        a = mogptk.MOSM(dataset(mogptk, X, y, 3), Q=2)
        a.init_parameters(method, iters=25)
    torch.manual_seed(3)
    b = mogptk.MOSM(dataset(mogptk, X, y, 3), Q=2, inference=mb.B200Exact())
    mb.install(mogptk)
    try:
        b.init_parameters(method, iters=25)
    finally:
        mb.uninstall(mogptk)
    for name in ("weight", "mean", "variance"):
        u = getattr(a.gpr.kernel, name).numpy()
        v = getattr(b.gpr.kernel, name).numpy()
        assert np.abs(u - v).max() <= 1e-4 * max(np.abs(u).max(), 1e-12), (name, u, v)
    assert close(a.gpr.likelihood.scale.numpy(), b.gpr.likelihood.scale.numpy(), 1e-10)
    assert mogptk.init.BNSE.__module__ == "mogptk.init"          # uninstall() restored the reference's functions
