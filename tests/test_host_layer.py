"""CPU: host-side logic of the reference-facing layer (mogptk_b200.gpr.Exact, B200Exact) with the
oracle-backed test double in place of the CUDA engine -- parameter packing, the chain rule into
raw-space p.grad, row permutations, factor caching, error mapping, pickling; and, when the
reference is importable (build container only), the real drop-in:
mogptk.MOSM(dataset, Q, inference=B200Exact(...)) -> train() / predict()."""
import os
import pickle
import sys

import numpy as np
import pytest
import torch

from conftest import load_golden
from fake_engine import FakeEngine
import mogptk_b200 as mb
from mogptk_b200 import gpr

HAVE_REF = os.path.isdir("/root/reference/mogptk")


def build_mirror(g, engine, raw_from_golden=True):
    kind, C, Q, D = g["kind"], g["C"], g["Q"], g["D"]
    if kind == "MOSM":
        k = gpr.MultiOutputSpectralMixtureKernel(Q=Q, output_dims=C, input_dims=D)
        plist = {n: [getattr(k, n)] for n in ("weight", "mean", "variance", "delay", "phase")}
    elif kind == "SM":
        k = gpr.IndependentMultiOutputKernel([gpr.SpectralMixtureKernel(Q=Q, input_dims=D) for _ in range(C)], output_dims=C)
        plist = {n: [getattr(k[c], n) for c in range(C)] for n in ("magnitude", "mean", "variance")}
    else:
        k = gpr.MixtureKernel(gpr.GaussianConvolutionProcessKernel(output_dims=C, input_dims=D), Q)
        plist = {n: [getattr(k[q], n) for q in range(Q)] for n in ("weight", "variance", "base_variance")}
    m = gpr.Exact(k, g["X"], g["y"], variance=(g["sigma"] ** 2).tolist(), data_variance=g.get("data_var"),
                  jitter=g["jitter"], engine=engine)
    plist["sigma"] = [m.likelihood.scale]
    if raw_from_golden:              # the reference's raw values (its transform inverse is not exact)
        for n, lst in plist.items():
            raw = torch.tensor(g["r_" + n], dtype=torch.float64)
            for i, prm in enumerate(lst):
                prm.data = (raw if len(lst) == 1 else raw[i]).clone().reshape(prm.shape).to(prm.device)
    return m, plist


@pytest.mark.parametrize("name", ["mosm_small", "mosm_shuffled", "mosm_datavar", "mosm_c1", "sm_small", "conv_small",
                                  "mosm_small_d2"])
def test_loss_fills_raw_gradients_like_the_reference(name):
    g = load_golden(name)
    m, plist = build_mirror(g, FakeEngine())
    loss = m.loss()
    assert abs(float(loss) - float(g["loss"])) <= 1e-10 * abs(float(g["loss"]))
    for n, lst in plist.items():
        ref = torch.tensor(g["gr_" + n])
        got = lst[0].grad if len(lst) == 1 else torch.stack([p.grad for p in lst])
        assert float((got.reshape(ref.shape) - ref).abs().max()) <= 1e-8 * max(float(ref.abs().max()), 1e-12), n
    mu, var = m.predict_f(g["Xs"])
    assert np.abs(mu.numpy().ravel() - g["pred_mu"]).max() <= 1e-8 * np.abs(g["pred_mu"]).max()
    assert np.abs(var.numpy().ravel() - g["pred_var"]).max() <= 1e-8 * np.abs(g["pred_var"]).max()


def test_parameter_semantics():
    p = gpr.Parameter(torch.tensor([1.0, 2.0]), lower=1e-8)
    assert abs(float(p()[0]) - 1.0000000995) < 1e-9            # reference quirk: inverse is not exact
    with pytest.raises(ValueError):
        p.assign(torch.ones(3))
    p.assign(upper=5.0)                                         # bounds-only assign re-reads the raw tensor
    assert p.transform.kind == "sigmoid"
    q = pickle.loads(pickle.dumps(p))
    assert torch.equal(q.data, p.data) and q._name == p._name and float(q.upper) == 5.0
    k = gpr.MultiOutputSpectralMixtureKernel(Q=2, output_dims=1)
    assert k.delay.train is False and k.phase.train is False   # gpr/multioutput.py:172-174
    with pytest.raises(AttributeError):
        k.weight = 3.0                                          # read-only, use assign()
    assert k.weight._name == "MultiOutputSpectralMixtureKernel.weight"


def test_unsupported_kernels_and_inputs_raise():
    class Other(gpr.Kernel):
        pass
    with pytest.raises(NotImplementedError):
        gpr.Exact(Other(), np.zeros((3, 2)), np.zeros(3), engine=FakeEngine())
    k = gpr.MultiOutputSpectralMixtureKernel(Q=1, output_dims=2)
    with pytest.raises(ValueError):
        gpr.Exact(k, np.zeros((3, 2)), np.zeros(4), engine=FakeEngine())
    with pytest.raises(ValueError):
        gpr.Exact(k, np.zeros((3, 2)), np.zeros(3), variance=[1.0, 1.0, 1.0], engine=FakeEngine())
    m = gpr.Exact(k, np.array([[0, 1.0], [1, 2.0], [5, 3.0]]), np.zeros(3), engine=FakeEngine())
    with pytest.raises(ValueError):                             # channel id 5 >= output_dims
        m.loss()


def test_factor_is_reused_until_parameters_change_and_model_pickles():
    g = load_golden("mosm_small")
    eng = FakeEngine()
    m, _ = build_mirror(g, eng)
    m.loss()
    n = eng.calls
    m.predict_f(g["Xs"]); m.predict_f(g["Xs"], full=True)
    assert eng.calls == n                                       # cached factor, unlike gpr/model.py:463-469
    m.kernel.weight.assign(m.kernel.weight() * 1.1)
    m.predict_f(g["Xs"])
    assert eng.calls == n + 1
    m2 = pickle.loads(pickle.dumps(m))
    assert m2._engine is None and torch.equal(m2.kernel.weight.data, m.kernel.weight.data)


def test_cholesky_failure_maps_to_exception():
    g = load_golden("mosm_small")
    m, _ = build_mirror(g, FakeEngine())
    m.kernel.weight.data.fill_(float("nan"))
    with pytest.raises(gpr.CholeskyException):
        m.loss()


def test_adam_training_follows_the_oracle_trajectory():
    from oracle import mogp_oracle as orc
    g = load_golden("mosm_small")
    m, plist = build_mirror(g, FakeEngine())
    ref = orc.RawModel(g["kind"], g["params"], g["sigma_t"], g["X"], g["y"], g["jitter"])
    for k in list(ref.raw):
        ref.raw[k] = torch.tensor(g["r_" + k], dtype=torch.float64).requires_grad_(True)
    oa = torch.optim.Adam(m.parameters(), lr=0.05)
    ob = torch.optim.Adam(list(ref.raw.values()), lr=0.05)
    for _ in range(5):
        la = m.loss(); oa.step()
        lb = ref.loss(); ob.step()
        assert abs(float(la) - float(lb)) <= 1e-9 * abs(float(lb))


@pytest.mark.skipif(not HAVE_REF, reason="reference checkout only exists in the build container")
def test_drop_in_under_the_reference_model_classes():
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
    from make_golden import import_reference
    mogptk = import_reference()
    from mogptk_b200 import synth
    rng_X, rng_y = synth.make_data(3, [40, 25, 33], seed=4)
    def dataset():
        ds = mogptk.DataSet()
        for c in range(3):
            msk = rng_X[:, 0] == c
            ds.append(mogptk.Data(rng_X[msk, 1], rng_y[msk], name=str(c)))
        return ds
    torch.manual_seed(0)
    a = mogptk.MOSM(dataset(), Q=2)                                        # stock reference path
    torch.manual_seed(0)
    b = mogptk.MOSM(dataset(), Q=2, inference=mb.B200Exact(engine=FakeEngine()))   # through the plug-in
    for mdl in (a, b):
        mdl.gpr.kernel.mean.assign(torch.full((3, 2, 1), 0.4))
    assert type(b.gpr).__name__ == "Exact" and type(b.gpr).__module__ == "mogptk_b200.gpr"
    la, _ = a.train(method="Adam", iters=4, lr=0.05, verbose=False, jit=False)
    lb, _ = b.train(method="Adam", iters=4, lr=0.05, verbose=False, jit=False)
    assert np.abs(la - lb).max() <= 1e-8 * np.abs(la).max()
    Xa, Ma, La, Ua = a.predict()
    Xb, Mb, Lb, Ub = b.predict()
    for u, v in zip(Ma + La + Ua, Mb + Lb + Ub):
        assert np.abs(u - v).max() <= 1e-7 * max(np.abs(u).max(), 1e-12)
    assert abs(a.log_marginal_likelihood() - b.log_marginal_likelihood()) <= 1e-8 * abs(a.log_marginal_likelihood())
    assert a.num_parameters() == b.num_parameters()
    b.print_parameters()


@pytest.mark.skipif(not HAVE_REF, reason="reference checkout only exists in the build container")
@pytest.mark.parametrize("family", ["SM", "CONV"])
def test_drop_in_for_the_other_two_model_families(family):
    """mogptk.SM (independent SM kernels) and mogptk.CONV (mixture of CONV kernels) through the plug-in."""
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
    from make_golden import import_reference
    mogptk = import_reference()
    from mogptk_b200 import synth
    Xd, yd = synth.make_data(2, [30, 24], seed=6)

    def dataset():
        ds = mogptk.DataSet()
        for c in range(2):
            msk = Xd[:, 0] == c
            ds.append(mogptk.Data(Xd[msk, 1], yd[msk], name=str(c)))
        return ds

    cls = getattr(mogptk, family)
    torch.manual_seed(1)
    a = cls(dataset(), Q=2)
    torch.manual_seed(1)
    b = cls(dataset(), Q=2, inference=mb.B200Exact(engine=FakeEngine()))
    if family == "SM":                                   # the constructor pins `mean`: assign explicitly
        for mdl in (a, b):
            for c in range(2):
                mdl.gpr.kernel[c].mean.assign(torch.tensor([[0.3], [0.8]]))
    la, _ = a.train(method="Adam", iters=3, lr=0.05, verbose=False, jit=False)
    lb, _ = b.train(method="Adam", iters=3, lr=0.05, verbose=False, jit=False)
    assert np.abs(la - lb).max() <= 1e-8 * np.abs(la).max()
    _, Ma, _, _ = a.predict()
    _, Mb, _, _ = b.predict()
    for u, v in zip(Ma, Mb):
        assert np.abs(u - v).max() <= 1e-7 * max(np.abs(u).max(), 1e-12)


@pytest.mark.skipif(not HAVE_REF, reason="reference checkout only exists in the build container")
def test_lbfgs_closure_and_resume_through_the_plugin():
    """mogptk.Model.train(method='LBFGS') wraps loss() in a closure (mogptk/model.py:546-553); a second train()
    call resumes and appends to the loss history (:501-509)."""
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
    from make_golden import import_reference
    mogptk = import_reference()
    from mogptk_b200 import synth
    Xd, yd = synth.make_data(2, [26, 20], seed=9)

    def model(**kw):
        ds = mogptk.DataSet()
        for c in range(2):
            msk = Xd[:, 0] == c
            ds.append(mogptk.Data(Xd[msk, 1], yd[msk], name=str(c)))
        torch.manual_seed(2)
        m = mogptk.MOSM(ds, Q=2, **kw)
        m.gpr.kernel.mean.assign(torch.full((2, 2, 1), 0.5))
        return m

    a, b = model(), model(inference=mb.B200Exact(engine=FakeEngine()))
    la, _ = a.train(method="LBFGS", iters=4, verbose=False, jit=False)
    lb, _ = b.train(method="LBFGS", iters=4, verbose=False, jit=False)
    assert len(la) == len(lb) and np.abs(la - lb).max() <= 1e-7 * np.abs(la).max()
    b.train(method="Adam", iters=2, lr=0.01, verbose=False, jit=True)       # jit=True -> compile() is a no-op here
    assert len(b.losses) == len(lb) + 2
    import pickle
    pickle.loads(pickle.dumps(b.gpr))                                       # what model.save() relies on


def test_trainable_mean_function_receives_gradients():
    """The reference back-propagates the LML through y - mean(X) (gpr/model.py:445-452); here d LML / d y = -alpha
    from the engine is chained into the mean's parameters (they silently stayed untrained)."""
    from oracle import mogp_oracle as orc
    g = load_golden("mosm_shuffled")

    class LinMean(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.bias = gpr.Parameter(0.3)
            self.slope = gpr.Parameter(torch.tensor([-0.2, 0.05], dtype=torch.float64))

        def forward(self, X):
            return self.bias() + X.mm(self.slope().reshape(-1, 1))

    eng = FakeEngine()
    m, plist = build_mirror(g, eng)
    mean = LinMean()
    m.mean = mean
    loss = m.loss()
    # oracle: same LML on y - mean(X), autograd into (bias, slope)
    b = torch.tensor(0.3, dtype=torch.float64, requires_grad=True)
    s = torch.tensor([-0.2, 0.05], dtype=torch.float64, requires_grad=True)
    X = torch.tensor(g["X"])
    yt = torch.tensor(g["y"]).reshape(-1, 1) - (b + X.mm(s.reshape(-1, 1)))
    p = {k: orc.softplus_forward(torch.tensor(g["r_" + k]), float(g["lower_" + k])) if bool(g["has_lower_" + k])
         else torch.tensor(g["r_" + k]) for k in g["params"]}
    sig = orc.softplus_forward(torch.tensor(g["r_sigma"]), float(g["lower_sigma"]))
    ref = -orc.lml(g["kind"], p, sig, X, yt, g["jitter"])
    ref.backward()
    assert abs(float(loss) - float(ref)) <= 1e-10 * abs(float(ref))
    assert mean.bias.grad is not None and mean.slope.grad is not None
    assert abs(float(mean.bias.grad) - float(b.grad)) <= 1e-8 * abs(float(b.grad))
    assert float((mean.slope.grad - s.grad).abs().max()) <= 1e-8 * float(s.grad.abs().max())
    mu, _ = m.predict_f(g["Xs"])                                # mean is added back (gpr/model.py:473-474)
    assert mu.shape == (g["Xs"].shape[0], 1)


def test_active_dims_other_than_identity_are_refused():
    k = gpr.MultiOutputSpectralMixtureKernel(Q=1, output_dims=2, input_dims=2)
    k.active_dims = torch.tensor([0, 1])
    gpr.kernel_spec(k)                                          # identity selection is fine
    k.active_dims = torch.tensor([1, 0])
    with pytest.raises(NotImplementedError):
        gpr.kernel_spec(k)
    sub = [gpr.SpectralMixtureKernel(Q=1, input_dims=1) for _ in range(2)]
    sub[1].active_dims = [1]                                    # out of range after the channel column is stripped
    with pytest.raises(NotImplementedError):
        gpr.kernel_spec(gpr.IndependentMultiOutputKernel(sub, output_dims=2))


@pytest.mark.skipif(not HAVE_REF, reason="reference checkout only exists in the build container")
@pytest.mark.parametrize("family,kw", [("CSM", dict(Q=2, Rq=2)), ("SM_LMC", dict(Q=2, Rq=2)), ("CSM", dict(Q=3, Rq=1))])
def test_drop_in_for_csm_and_sm_lmc(family, kw):
    """mogptk.CSM (mixture of CrossSpectralKernel) and mogptk.SM_LMC (LMC of SpectralKernel) through the plug-in:
    kernel_spec's packing of the reference's parameter objects and the gradient routing back into them."""
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
    from make_golden import import_reference
    mogptk = import_reference()
    from mogptk_b200 import synth
    Xd, yd = synth.make_data(3, [22, 30, 18], seed=8)

    def dataset():
        ds = mogptk.DataSet()
        for c in range(3):
            msk = Xd[:, 0] == c
            ds.append(mogptk.Data(Xd[msk, 1], yd[msk], name=str(c)))
        return ds

    cls = getattr(mogptk, family)
    torch.manual_seed(4)
    a = cls(dataset(), **kw)
    torch.manual_seed(4)
    b = cls(dataset(), inference=mb.B200Exact(engine=FakeEngine()), **kw)
    g = torch.Generator().manual_seed(1)
    for (na, pa), (nb, pb) in zip(a.gpr.named_parameters(), b.gpr.named_parameters()):
        assert na == nb
        if na.endswith(".mean"):                     # off the constructor's dead-mean start
            pa.assign(0.2 + torch.rand(pa.shape, generator=g, dtype=torch.float64))
        elif na.endswith(".shift"):
            pa.assign(0.3 * torch.randn(pa.shape, generator=g, dtype=torch.float64))
        pb.data = pa.data.detach().clone()
    assert b.gpr._kind.startswith("CSM" if family == "CSM" else "SMLMC")
    la, _ = a.train(method="Adam", iters=4, lr=0.05, verbose=False, jit=False)
    lb, _ = b.train(method="Adam", iters=4, lr=0.05, verbose=False, jit=False)
    assert np.abs(la - lb).max() <= 1e-8 * np.abs(la).max()
    for (na, pa), (nb, pb) in zip(a.gpr.named_parameters(), b.gpr.named_parameters()):
        assert float((pa.grad - pb.grad).abs().max()) <= 1e-7 * max(float(pa.grad.abs().max()), 1e-10), na
    _, Ma, _, _ = a.predict()
    _, Mb, _, _ = b.predict()
    for u, v in zip(Ma, Mb):
        assert np.abs(u - v).max() <= 1e-7 * max(np.abs(u).max(), 1e-12)
