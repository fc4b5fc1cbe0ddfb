"""GPU: the reference-facing model layer (mogptk_b200.gpr.Exact on the real CUDA engine) against
the golden vectors of the live reference: loss() fills raw-space p.grad, predict_f / predict_y,
K / K_diag, CholeskyException, pickling, a short Adam run against the oracle trajectory."""
import pickle

import numpy as np
import pytest
import torch

from conftest import load_golden

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gpr():
    from mogptk_b200 import gpr as g
    g.use_gpu(0)
    return g


def build(gpr, g):
    from test_host_layer import build_mirror
    return build_mirror(g, None)


@pytest.mark.parametrize("name", ["mosm_small", "mosm_shuffled", "mosm_datavar", "mosm_c1", "mosm_mid", "sm_small",
                                  "sm_small_d2", "conv_small", "conv_small_d2", "mosm_small_d2", "cfg1", "cfg2_rdp"])
def test_loss_and_raw_gradients(gpr, name):
    g = load_golden(name)
    m, plist = build(gpr, g)
    loss = m.loss()
    assert all(p.grad.is_cuda for p in m.parameters())     # raw-space gradients were written on the device
    assert abs(float(loss) - float(g["loss"])) <= 1e-8 * abs(float(g["loss"]))       # LML rtol 1e-8
    for n, lst in plist.items():
        ref = torch.tensor(g["gr_" + n])
        got = (lst[0].grad if len(lst) == 1 else torch.stack([p.grad for p in lst])).cpu()
        assert float((got.reshape(ref.shape) - ref).abs().max()) <= 1e-6 * max(float(ref.abs().max()), 1e-12), n
    lml = m.log_marginal_likelihood()
    assert abs(float(lml) - float(g["lml"])) <= 1e-8 * abs(float(g["lml"]))
    mu, var = m.predict_f(g["Xs"])
    assert mu.shape == (g["Xs"].shape[0], 1) and var.shape == (g["Xs"].shape[0], 1)
    assert np.abs(mu.cpu().numpy().ravel() - g["pred_mu"]).max() <= 1e-6 * np.abs(g["pred_mu"]).max()
    assert np.abs(var.cpu().numpy().ravel() - g["pred_var"]).max() <= 1e-6 * np.abs(g["pred_var"]).max()


def test_predict_y_and_kernel_calls(gpr):
    g = load_golden("mosm_small")
    m, _ = build(gpr, g)
    mu, lo, up = m.predict_y(g["Xs"], sigma=2.0)
    scale = torch.tensor(g["sigma"])[torch.tensor(g["Xs"][:, 0]).long()].reshape(-1, 1)
    assert np.abs(mu.cpu().numpy().ravel() - g["pred_mu"]).max() <= 1e-6 * np.abs(g["pred_mu"]).max()
    # reference quirk (gpr/likelihood.py:355-367): multi-output band = mu -/+ sigma * scale_c
    assert torch.allclose((up - lo).cpu(), 4.0 * scale.to(torch.float64), rtol=1e-12)
    K = m.K(g["X"])
    assert np.abs(K.cpu().numpy() - g["K_full"]).max() <= 1e-12 * np.abs(g["K_full"]).max()
    assert torch.equal(m.kernel.K_diag(torch.tensor(g["X"], device=K.device)), K.diagonal())
    with pytest.raises(ValueError):
        m.kernel(torch.tensor([[7.0, 1.0]]))                    # channel id out of range
    s = m.sample_y(g["Xs"], n=3)
    assert s.shape == (3, g["Xs"].shape[0]) and torch.isfinite(s).all()


def test_cholesky_exception_and_pickle(gpr):
    g = load_golden("mosm_small")
    m, _ = build(gpr, g)
    m.loss()
    m2 = pickle.loads(pickle.dumps(m))
    assert abs(float(m2.log_marginal_likelihood()) - float(g["lml"])) <= 1e-8 * abs(float(g["lml"]))
    m.kernel.weight.data.fill_(float("nan"))
    with pytest.raises(gpr.CholeskyException) as ei:
        m.loss()
    assert "not positive-definite" in str(ei.value) and ei.value.K is not None and ei.value.model is m


def test_adam_training_matches_the_oracle_trajectory(gpr):
    from oracle import mogp_oracle as orc
    g = load_golden("mosm_mid")
    m, _ = build(gpr, g)
    ref = orc.RawModel(g["kind"], g["params"], g["sigma_t"], g["X"], g["y"], g["jitter"])
    for k in list(ref.raw):
        ref.raw[k] = torch.tensor(g["r_" + k], dtype=torch.float64).requires_grad_(True)
    oa = torch.optim.Adam(m.parameters(), lr=0.05)
    ob = torch.optim.Adam(list(ref.raw.values()), lr=0.05)
    first = None
    for _ in range(8):
        la = m.loss(); oa.step()
        lb = ref.loss(); ob.step()
        first = first if first is not None else float(lb)
        assert abs(float(la) - float(lb)) <= 1e-7 * abs(float(lb))
    assert float(lb) < first


def test_plugin_builder_returns_the_engine_model(gpr):
    import mogptk_b200 as mb
    g = load_golden("conv_small")
    k = gpr.MixtureKernel(gpr.GaussianConvolutionProcessKernel(output_dims=g["C"], input_dims=g["D"]), g["Q"])
    for q in range(g["Q"]):
        k[q].weight.assign(g["params"]["weight"][q])
        k[q].variance.assign(g["params"]["variance"][q])
        k[q].base_variance.assign(g["params"]["base_variance"][q])
    model = mb.B200Exact(variance=(g["sigma"] ** 2).tolist(), jitter=g["jitter"])._build(k, g["X"], g["y"].reshape(-1, 1))
    assert isinstance(model, gpr.Exact) and model.likelihood.scale.shape == (g["C"],)
    lml = float(model.log_marginal_likelihood())
    assert abs(lml - float(g["lml"])) <= 1e-6 * abs(float(g["lml"]))    # assign() round trip moves params by ~1e-7


@pytest.mark.parametrize("name", ["mosm_mid", "sm_small", "conv_small"])
def test_device_resident_iteration_equals_the_autograd_path(gpr, name):
    """loss() normally runs transforms + chain rule on the device (mogp_params_forward/backward); forcing the
    general torch-autograd path must give the same loss and raw gradients."""
    g = load_golden(name)
    m, plist = build(gpr, g)
    assert m._fast_table() is not None
    la = m.loss()
    fast = {n: [p.grad.clone() for p in lst] for n, lst in plist.items()}
    m._fast_table = lambda: None                      # general path: torch autograd through Softplus / Sigmoid
    lb = m.loss()
    assert abs(float(la) - float(lb)) <= 1e-13 * abs(float(lb))
    for n, lst in plist.items():
        for p, ga in zip(lst, fast[n]):
            scale = max(float(p.grad.abs().max()), 1e-12)
            assert float((p.grad - ga).abs().max()) <= 1e-12 * scale, n


def test_sigmoid_bounded_parameter_on_the_device_path(gpr):
    """An upper bound turns the transform into a Sigmoid (what mogptk.MOSM does with the Nyquist frequency)."""
    g = load_golden("mosm_small")
    m, plist = build(gpr, g)
    m.kernel.mean.assign(m.kernel.mean().detach(), upper=torch.full_like(m.kernel.mean().detach(), 5.0))
    assert m.kernel.mean.transform.kind == "sigmoid"
    la = m.loss()
    ga = m.kernel.mean.grad.clone()
    m._fast_table = lambda: None
    lb = m.loss()
    assert abs(float(la) - float(lb)) <= 1e-13 * abs(float(lb))
    assert float((m.kernel.mean.grad - ga).abs().max()) <= 1e-12 * max(float(ga.abs().max()), 1e-12)


def test_fused_adam_equals_loss_plus_torch_adam(gpr):
    """mogp_train_adam (one C call for K iterations) against loss() + torch.optim.Adam.step() on a twin model."""
    from mogptk_b200 import fit_adam
    g = load_golden("mosm_mid")
    a, _ = build(gpr, g)
    b, _ = build(gpr, g)
    opt = torch.optim.Adam(a.parameters(), lr=0.03, betas=(0.8, 0.95), eps=1e-7)
    ref = []
    for _ in range(40):
        ref.append(float(a.loss()))
        opt.step()
    got, times = fit_adam(b, 40, lr=0.03, betas=(0.8, 0.95), eps=1e-7, sync_every=16)
    assert got.shape == (40,) and np.all(np.diff(times) > 0)
    assert np.abs(got - np.array(ref)).max() <= 1e-9 * np.abs(ref).max()
    for p, q in zip(a.parameters(), b.parameters()):
        assert float((p.data - q.data).abs().max()) <= 1e-9 * max(float(p.data.abs().max()), 1.0)
    assert abs(float(a.loss()) - float(b.loss())) <= 1e-9 * abs(float(a.loss()))
    mu_a, _ = a.predict_f(g["Xs"])
    mu_b, _ = b.predict_f(g["Xs"])                              # factor cache was invalidated by the in-place updates
    assert float((mu_a - mu_b).abs().max()) <= 1e-8 * float(mu_a.abs().max())


def test_concurrent_restarts_on_one_gpu_equal_sequential_training(gpr):
    """replicas.train_restarts: several models, each with its own handle / stream / host thread, train concurrently and end
    exactly where they end when trained one after the other (bit-for-bit: the per-model arithmetic does not change)."""
    from mogptk_b200 import fit_adam, replicas
    from mogptk_b200.engine import Engine
    g = load_golden("mosm_mid")
    engines = [Engine(device=0, max_n=512) for _ in range(3)]
    try:
        seq, con = [], []
        for i in range(3):
            a, _ = build_with_engine(gpr, g, engines[i])
            a.kernel.weight.data += 0.01 * i
            seq.append(a)
            b, _ = build_with_engine(gpr, g, engines[i])
            b.kernel.weight.data += 0.01 * i
            con.append(b)
        ref = [fit_adam(m, 24, lr=0.02, sync_every=8)[0] for m in seq]
        got = replicas.train_restarts(con, 24, lr=0.02, sync_every=8)
        for r, h in zip(ref, got):
            assert np.array_equal(r, h)
        for a, b in zip(seq, con):
            for p, q in zip(a.parameters(), b.parameters()):
                assert torch.equal(p.data, q.data)
    finally:
        for e in engines:
            e.close()


def build_with_engine(gpr, g, engine):
    from test_host_layer import build_mirror
    return build_mirror(g, engine)


@pytest.mark.parametrize("name", ["mosm_mid", "cfg2_rdp"])
def test_early_loss_equals_the_synchronous_path(gpr, name, monkeypatch):
    """loss() returns as soon as the step has published [lml, info] (mapped pinned memory, right after the solves), while K^-1,
    the gradient reduction and the chain rule are still running: the value is bit-identical to the synchronous path's, p.grad
    read afterwards (stream order) is complete, an optimiser step enqueued right behind it sees the whole gradient, and a
    Cholesky failure still raises inside loss()."""
    g = load_golden(name)
    m, plist = build(gpr, g)
    eng = m._eng()
    assert eng.early_loss_buffer() is not None
    losses, grads = [], []
    for _ in range(4):                                   # plain run, capture, replays
        l = m.loss()
        losses.append(float(l))
        grads.append([p.grad.clone() for p in m.parameters()])
    assert all(v == losses[0] for v in losses)
    for gl in grads[1:]:
        assert all(torch.equal(a, b) for a, b in zip(gl, grads[0]))
    # the synchronous path on a fresh model / engine state
    monkeypatch.setenv("MOGP_EARLY_LOSS", "0")
    from mogptk_b200.engine import Engine
    eng2 = Engine(device=0, max_n=g["X"].shape[0])
    try:
        m2, _ = build(gpr, g)
        m2._engine = eng2
        l2 = m2.loss()
        assert eng2.early_loss_buffer() is None
        assert float(l2) == losses[0]
        for a, b in zip(m2.parameters(), grads[0]):
            assert torch.equal(a.grad, b)
    finally:
        eng2.close()
    monkeypatch.delenv("MOGP_EARLY_LOSS")
    # an optimiser step enqueued right behind the early return uses the complete gradient
    opt = torch.optim.SGD(m.parameters(), lr=1e-3)
    before = [p.detach().clone() for p in m.parameters()]
    m.loss()
    opt.step()
    for p, b0, g0 in zip(m.parameters(), before, grads[0]):
        assert torch.allclose(p.detach(), b0 - 1e-3 * g0, rtol=0, atol=1e-12 * max(1.0, float(b0.abs().max())))
